"""Warp-stall summary of one profiled launch from `ncu -i X.ncu-rep --page source --csv` (needs -lineinfo and
--import-source on): totals per stall reason and the instructions that collect the most samples.

    python benchmarks/ncu_stalls.py <source.csv> <out.json> [top_n]
"""
import csv
import json
import sys

REASONS = ["stall_long_sb", "stall_short_sb", "stall_lg", "stall_mio", "stall_math", "stall_wait", "stall_barrier",
           "stall_branch_resolving", "stall_dispatch", "stall_no_inst", "stall_not_selected", "stall_selected",
           "stall_drain", "stall_membar", "stall_sleep", "stall_tex", "stall_misc"]


def main():
    src, out = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 12
    rows = list(csv.reader(open(src)))
    kernel = rows[0][1] if rows and len(rows[0]) > 1 else ""
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]

    def num(r, k):
        try:
            return int(r[idx[k]] or 0)
        except (KeyError, ValueError):
            return 0

    total = sum(num(r, "# Samples") for r in data)
    reasons = {k: sum(num(r, k) for r in data) for k in REASONS if k in idx}
    top = sorted(data, key=lambda r: -num(r, "# Samples"))[:top_n]
    rec = {"kernel": kernel, "samples": total, "sass_instructions": len(data),
           "stall_share": {k: round(v / max(total, 1), 4) for k, v in sorted(reasons.items(), key=lambda kv: -kv[1]) if v},
           "top_instructions": [{"sass": r[idx["Source"]].strip(), "samples": num(r, "# Samples"),
                                 "share": round(num(r, "# Samples") / max(total, 1), 4),
                                 "main_reason": max(REASONS, key=lambda k: num(r, k))} for r in top]}
    with open(out, "w") as fh:
        json.dump(rec, fh, indent=1)
    print(json.dumps(rec["stall_share"]))


if __name__ == "__main__":
    main()
