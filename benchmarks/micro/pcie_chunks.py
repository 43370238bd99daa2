"""Pinned-host <-> device copy rate with both directions busy, the slab of each direction split over 1 / 2 / 4 streams
(does more than one copy engine per direction help the end-to-end arm of bench.py?).  326 MB per direction per step."""
import json
import time

import torch

n = 325754880 // 4
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_in = torch.empty(n, dtype=torch.float32, device="cuda")
d_out = torch.randn(n, dtype=torch.float32, device="cuda")
res = {}
for chunks in (1, 2, 4):
    s_in = [torch.cuda.Stream() for _ in range(chunks)]
    s_out = [torch.cuda.Stream() for _ in range(chunks)]
    step = (n + chunks - 1) // chunks

    def one():
        for c in range(chunks):
            a, b = c * step, min(n, (c + 1) * step)
            with torch.cuda.stream(s_in[c]):
                d_in[a:b].copy_(h_in[a:b], non_blocking=True)
            with torch.cuda.stream(s_out[c]):
                h_out[a:b].copy_(d_out[a:b], non_blocking=True)
    for _ in range(3):
        one()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        one()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    res[f"{chunks} stream(s) per direction"] = {"ms_per_step": round(dt * 1e3, 3), "GBps_per_direction": round(n * 4 / dt / 1e9, 2)}
print(json.dumps(res))
