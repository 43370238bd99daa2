"""Turn `ncu -i X.ncu-rep --page raw --csv` dumps into the small, committed evidence under profiles/.

    python benchmarks/ncu_summarize.py <tag> <raw.csv> [<raw.csv> ...]

Writes profiles/<tag>_<kernel>.json (picked metrics of each profiled launch) and refreshes
profiles/ncu_summary.json, which bench.py reads for `roofline.traffic` (DRAM bytes per launch).
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from benchmarks.ncu_pick import KEYS  # noqa: E402

EXTRA = ["launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
         "launch__occupancy_limit_registers", "sm__maximum_warps_per_active_cycle_pct",
         "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
         "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
         "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
         "smsp__cycles_active.avg", "sm__cycles_active.avg"]
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def main():
    tag, files = sys.argv[1], sys.argv[2:]
    out_dir = os.path.join(ROOT, "profiles")
    os.makedirs(out_dir, exist_ok=True)
    summary_path = os.path.join(out_dir, "ncu_summary.json")
    summary = json.load(open(summary_path)) if os.path.exists(summary_path) else {}
    for path in files:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            short = re.match(r"(?:void\s+)?(\w+)", name).group(1)
            rec = {"kernel": name, "source": os.path.basename(path), "tag": tag,
                   "grid": r[hdr.index("Grid Size")] if "Grid Size" in hdr else None,
                   "block": r[hdr.index("Block Size")] if "Block Size" in hdr else None, "metrics": {}}
            for k in KEYS + EXTRA:
                if k in hdr:
                    i = hdr.index(k)
                    try:
                        v = float(r[i].replace(",", ""))
                    except ValueError:
                        v = r[i]
                    rec["metrics"][k] = {"value": v, "unit": units[i]}
            m = rec["metrics"]

            def in_bytes(key):
                return m[key]["value"] * UNIT_SCALE.get(m[key]["unit"], 1.0)

            dram = in_bytes("dram__bytes_read.sum") + in_bytes("dram__bytes_write.sum")
            rec["dram_bytes_per_launch"] = dram
            with open(os.path.join(out_dir, f"{tag}_{short}.json"), "w") as fh:
                json.dump(rec, fh, indent=1)
            summary[short] = {"dram_bytes_per_launch": dram, "tag": tag, "kernel": name,
                              "duration_us_under_ncu": m["gpu__time_duration.sum"]["value"] *
                              (1e3 if m["gpu__time_duration.sum"]["unit"] == "ms" else 1.0),
                              "l1_data_pipe_pct": m.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", {}).get("value"),
                              "lts_pct": m.get("lts__throughput.avg.pct_of_peak_sustained_elapsed", {}).get("value"),
                              "dram_pct": m.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", {}).get("value")}
            print(short, summary[short])
    with open(summary_path, "w") as fh:
        json.dump(summary, fh, indent=1)


if __name__ == "__main__":
    main()
