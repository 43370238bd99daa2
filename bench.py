"""bench.py -- the contract benchmark of the temporal MSDeformAttn hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype fp32|bf16] [--dist local|uniform]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One STEP = one encoder-layer temporal attention over one clip, forward + backward, at the DeVIS R50 T=6
YouTube-VIS shape (BASELINE.json configs[1]; SURVEY.md section 8d "unit U"): T=6 frames, levels
(45,80),(23,40),(12,20),(6,10) -> S=4820 rows per frame, 8 heads x 32 channels, one query per pixel,
4 current + 5x4x4 temporal points -> 96 taps per (query, head).  It is what the reference does with 12
MSDeformAttnFunction calls + 6 gather copies of value (modules/ms_deform_attn.py:435-460) and what this
repository does with one forward launch and one backward launch.

Rank 0 prints ONE JSON line.  `value` is whole-job layer-clips per second with inputs resident in HBM;
`e2e` is the same unit of work driven through the public autograd API from pinned HOST buffers, every
step paying the host->device copy of its inputs and the device->host copy of its results.
Multi-GPU: clips are independent, so each rank works on its own clip (weak scaling, no collective on
the op path -- SURVEY.md section 8e).

`--impl reference` times the reference's CPU implementation of the same unit of work (the oracle's
PyTorch restatement of ms_deform_attn_core_pytorch + autograd, all host threads) on a bounded sample.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "temporal_msda_fwd_bwd_layer_clips_per_sec"
UNIT = "layer-clips/s"
WORKLOAD = "DeVIS R50 T=6 encoder layer-clip: S=4820, M=8, D=32, Lq=4820/frame, K=96 taps, fwd+bwd"
T_FRAMES, S_ROWS, HEADS, CH, K_TAPS = 6, 4820, 8, 32, 96


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe), through NVML
    (the same counters nvidia-smi prints) every 5 ms from a helper thread."""
    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20),
               ("hw_thermal_slowdown", 0x40), ("hw_power_brake_slowdown", 0x80))

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []          # (perf_counter, sm MHz, reason mask, watts)
        self.window = None         # (t0, t1): the timed region; samples outside it are dropped
        self.max_mhz = None
        self._stop = threading.Event()
        self.thread = None
        self.err = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = int(visible.split(",")[self.gpu]) if visible and visible.split(",")[self.gpu].isdigit() else self.gpu
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception as exc:  # noqa: BLE001
            self.err = f"nvml unavailable: {exc}"
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                watts = nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                self.samples.append((time.perf_counter(), mhz, mask, watts))
            except Exception as exc:  # noqa: BLE001
                self.err = str(exc)
                return
            time.sleep(0.002)

    def stop(self):
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=1)
        t0, t1 = self.window if self.window else (float("-inf"), float("inf"))
        inside = [x for x in self.samples if t0 <= x[0] <= t1]
        reasons = sorted({name for _, _, mask, _ in inside for name, bit in self.REASONS if mask & bit})
        out = {"sm_mhz": statistics.median(x[1] for x in inside) if inside else None, "sm_max_mhz": self.max_mhz,
               "samples": len(inside), "reasons": reasons,
               "power_w_max": max(x[3] for x in inside) if inside else None}
        if self.err:
            out["note"] = self.err
        return out


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel_key):
    """per-launch DRAM traffic of the dominant kernel from the committed ncu --set full capture"""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as fh:
            return json.load(fh)[kernel_key]["dram_bytes_per_launch"]
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline
# ----------------------------------------------------------------------------------------------------
def cpu_reference_sample(reps, dist, seed=0):
    """Times the reference's CPU path (oracle PyTorch restatement: grid_sample forward + autograd backward)
    on ONE of the T=6 query frames of the workload: its current-frame call, the gather copy of the 5
    other frames' value and the temporal call (2 of the 12 op calls of a layer-clip).  Returns
    (seconds per sample [median], cores, description)."""
    import torch
    from devis_b200 import synthetic
    from oracle import msda_torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    clip = synthetic.make_clip(dist=dist, seed=seed, device="cpu")
    shapes = torch.tensor(clip["shapes"])
    t = T_FRAMES // 2
    frames = clip["frame_table"][t]
    tshapes = shapes.repeat(len(frames), 1)
    times = []
    for _ in range(reps):
        v = clip["value"].clone().requires_grad_(True)
        lc, ac = clip["loc_curr"][t][None].clone().requires_grad_(True), clip["aw_curr"][t][None].clone().requires_grad_(True)
        lt, at = clip["loc_temporal"][t][None].clone().requires_grad_(True), clip["aw_temporal"][t][None].clone().requires_grad_(True)
        t0 = time.perf_counter()
        cur = msda_torch.msda_forward_torch(v[t][None], shapes, lc, ac)
        stacked = v[frames].flatten(0, 1)[None]
        tmp = msda_torch.msda_forward_torch(stacked, tshapes, lt, at)
        (cur + tmp).backward(clip["grad_out"][t][None])
        times.append(time.perf_counter() - t0)
    return statistics.median(times), cores, "1 of 6 query frames (current call + gather copy + temporal call), fwd+autograd bwd, fp32"


def run_reference(args, rank):
    if rank != 0:
        return
    t_wall = time.perf_counter()
    for _ in range(min(args.warmup, 1)):
        cpu_reference_sample(1, args.dist)
    sec, cores, sample = cpu_reference_sample(max(1, min(args.steps, 5)), args.dist)
    per_clip = sec * T_FRAMES
    value = 1.0 / per_clip
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_clip * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "dist": args.dist, "note": "CPU path; value extrapolated x6 from the sample"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_wall,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist_mod
    from devis_b200 import _lib, clip_geometry, synthetic, temporal_ms_deform_attn

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist_mod.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    dtype = {"fp32": torch.float32, "bf16": torch.bfloat16}[args.dtype]
    half_acc = args.bf16_accumulate and dtype == torch.bfloat16
    if half_acc:
        from devis_b200 import MultiScaleDeformableAttention as MSDA
        MSDA.set_bf16_grad_value_accumulation(True)
    bwd_flags = _lib.FLAG_BF16_GRAD_VALUE if half_acc else 0
    clip = synthetic.make_clip(dist=args.dist, dtype=dtype, seed=100 + rank, device=dev)
    geom = clip_geometry.ClipGeometry(clip["shapes"], T_FRAMES, clip["frame_table"])
    order = geom.tile_order(dev, 8, 8)
    names = ("value", "loc_curr", "aw_curr", "loc_temporal", "aw_temporal")
    leaves = [clip[k].requires_grad_(True) for k in names]
    gout = clip["grad_out"]

    def step():
        for x in leaves:
            x.grad = None
        out = temporal_ms_deform_attn(*leaves, geom, order)
        out.backward(gout)
        return out

    def barrier():
        if world > 1:
            dist_mod.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()        # started before the warm-up so that NVML's slow first calls are over by the timed region
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_region0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    sampler.window = (t_region0, time.perf_counter())
    ms_total = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist_mod.all_reduce(t, op=dist_mod.ReduceOp.MAX)
        ms_total = float(t.item())
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist_mod.all_reduce(lt, op=dist_mod.ReduceOp.SUM)
        launches = int(lt.item())
    ms_step = ms_total / args.steps
    value = world / (ms_step * 1e-3)

    # ---- per-kernel durations (events around each launch, same stream, same inputs) for the roofline
    from benchmarks.sweep import RawClip
    raw = RawClip({k: (v.detach() if hasattr(v, "detach") else v) for k, v in clip.items()}, order)
    f_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    b_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for _ in range(3):
        raw.fwd(); raw.bwd(bwd_flags)
    torch.cuda.synchronize()
    for (fa, fb), (ba, bb) in zip(f_ev, b_ev):
        fa.record(); raw.fwd(); fb.record()
        ba.record(); raw.bwd(bwd_flags); bb.record()
    torch.cuda.synchronize()
    us_fwd = statistics.mean(a.elapsed_time(b) for a, b in f_ev) * 1e3
    us_bwd = statistics.mean(a.elapsed_time(b) for a, b in b_ev) * 1e3
    elem = clip["value"].element_size()
    bytes_f, bytes_b = synthetic.algorithmic_bytes(T_FRAMES, S_ROWS, HEADS, CH, S_ROWS, K_TAPS, elem=elem)
    peak, peak_src = measured_peak_gbs()

    # ---- secondary rows (N = 1 only, not part of the headline): the same unit of work with bf16 value, and the
    # fused-prologue form the encoder module runs (raw Linear outputs in, softmax + location arithmetic in the kernels)
    extra = {}
    if world == 1 and dtype == torch.float32:
        try:
            from benchmarks.sweep import time_us
            clip16 = synthetic.make_clip(dist=args.dist, dtype=torch.bfloat16, seed=100 + rank, device=dev)
            raw16 = RawClip(clip16, order)
            f16, b16 = time_us(raw16.fwd, 20), time_us(raw16.bwd, 20)
            extra["bf16_value"] = {"us_fwd": round(f16, 1), "us_bwd": round(b16, 1),
                                   "layer_clips_per_sec": round(1e6 / (f16 + b16), 1)}
            del clip16, raw16
            from devis_b200 import TemporalMSDeformAttnFusedFunction
            g = torch.Generator(device=dev).manual_seed(7)
            nl, wt = len(clip["shapes"]), T_FRAMES - 1
            ref = synthetic.pixel_reference_points(clip["shapes"], T_FRAMES, dev)
            shp = (T_FRAMES, S_ROWS, HEADS)
            ops = [clip["value"].detach().clone().requires_grad_(True), ref,
                   (2.0 * torch.randn(*shp, nl, 4, 2, generator=g, device=dev)).requires_grad_(True),
                   torch.randn(*shp, nl * 4, generator=g, device=dev).requires_grad_(True),
                   (2.0 * torch.randn(*shp, wt * nl, 4, 2, generator=g, device=dev)).requires_grad_(True),
                   torch.randn(*shp, wt * nl * 4, generator=g, device=dev).requires_grad_(True)]

            def fused_fwd():
                with torch.no_grad():
                    TemporalMSDeformAttnFusedFunction.apply(*ops, geom, order)

            def fused_fwd_bwd():
                for x in ops:
                    x.grad = None
                TemporalMSDeformAttnFusedFunction.apply(*ops, geom, order).backward(gout)

            extra["fused_prologue_f32"] = {"us_fwd": round(time_us(fused_fwd, 20), 1),
                                           "us_fwd_bwd": round(time_us(fused_fwd_bwd, 20), 1)}
            del ops
        except Exception as exc:   # noqa: BLE001 -- secondary rows must never take the headline down
            extra["error"] = str(exc)[:200]
    fwd_kernel = "msda_fwdc_kernel" if dtype == torch.float32 else "msda_fwd8_kernel"   # what the launcher picks at D = 32

    # ---- the on-chip resources that actually bind: every tap gathers 4 value rows through the SM's
    # L1 data pipe (128 B/clk/SM), every live corner leaves the SM as one row reduction (5.16 cycles per 128-B row,
    # benchmarks/micro/tma_reduce.cu).  Counted from this step's own taps, timed live.
    def live_corners():
        n = 0
        for loc, reps in ((clip["loc_curr"], 1), (clip["loc_temporal"], T_FRAMES - 1)):
            size = torch.tensor([[w, h] for h, w in clip["shapes"]], device=dev, dtype=torch.float32).repeat(reps, 1)
            pix = loc.detach().float() * size[None, None, None, :, None, :] - 0.5
            x, y = pix[..., 0], pix[..., 1]
            wl, hl = size[:, 0][None, None, None, :, None], size[:, 1][None, None, None, :, None]
            inr = (x > -1) & (y > -1) & (x < wl) & (y < hl)
            x0, y0 = torch.floor(x), torch.floor(y)
            lft, rgt, top, bot = x0 >= 0, x0 + 1 <= wl - 1, y0 >= 0, y0 + 1 <= hl - 1
            n += int(((inr & top & lft).sum() + (inr & top & rgt).sum() + (inr & bot & lft).sum() + (inr & bot & rgt).sum()).item())
        return n

    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    sm_hz = 1e6 * float((clocks or {}).get("sm_mhz") or 1965.0)
    row_bytes = CH * elem
    gather_bytes = T_FRAMES * S_ROWS * HEADS * K_TAPS * 4 * row_bytes
    l1_peak = sm_count * 128 * sm_hz                      # bytes per second through the L1/shared data pipes
    red_rows = live_corners()
    on_chip = {
        "fwd_l1_gather": {"bytes": gather_bytes, "achieved_TBps": gather_bytes / us_fwd / 1e6, "peak_TBps": l1_peak / 1e12,
                          "frac": gather_bytes / (us_fwd * 1e-6) / l1_peak},
        "bwd_reduction_egress": {"row_reductions": red_rows, "cycles_per_row": 5.16,
                                 "busy_frac": red_rows / sm_count * 5.16 / (us_bwd * 1e-6 * sm_hz)},
        "note": "gathered value rows through the SMs' L1 data pipes (128 B/clk/SM at the sampled SM clock) and grad_value row "
                "reductions leaving the SMs (5.16 cycles per 128-B row measured); these, not HBM, bound the kernels",
    }

    # ---- e2e: public autograd API driven from pinned HOST buffers.  Every step copies its six operands host->device
    # and its six results device->host; copies of neighbouring steps overlap the kernels (three streams, two
    # buffer sets), as any host-fed pipeline would run it.  Timed with events on the compute stream + a final sync.
    host_in = [x.detach().cpu().pin_memory() for x in leaves] + [gout.cpu().pin_memory()]
    n_buf = 2
    dev_in = [[torch.empty_like(h, device=dev) for h in host_in] for _ in range(n_buf)]
    host_out = [None] * n_buf
    s_h2d, s_d2h, s_comp = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.current_stream()
    ev_in = [torch.cuda.Event() for _ in range(n_buf)]       # inputs of buffer b are on the device
    ev_free = [torch.cuda.Event() for _ in range(n_buf)]     # compute no longer needs buffer b's inputs
    ev_out = [torch.cuda.Event() for _ in range(n_buf)]      # results of buffer b are computed
    ev_copied = [torch.cuda.Event() for _ in range(n_buf)]   # results of buffer b are on the host
    results = [None] * n_buf

    def e2e_pipeline(n_steps):
        for b in range(n_buf):
            ev_free[b].record(s_comp)
            ev_copied[b].record(s_d2h)
        for i in range(n_steps):
            b = i % n_buf
            with torch.cuda.stream(s_h2d):
                s_h2d.wait_event(ev_free[b])
                for h, d in zip(host_in, dev_in[b]):
                    d.copy_(h, non_blocking=True)
                ev_in[b].record(s_h2d)
            s_comp.wait_event(ev_in[b])
            s_comp.wait_event(ev_copied[b])                  # the previous results held in slot b have left
            ins = [d.requires_grad_(True) for d in dev_in[b][:5]]
            out = temporal_ms_deform_attn(*ins, geom, order)
            out.backward(dev_in[b][5])
            results[b] = [out.detach()] + [x.grad for x in ins]
            for d in dev_in[b][:5]:
                d.requires_grad_(False)
                d.grad = None
            ev_free[b].record(s_comp)
            ev_out[b].record(s_comp)
            with torch.cuda.stream(s_d2h):
                s_d2h.wait_event(ev_out[b])
                if host_out[b] is None:
                    host_out[b] = [torch.empty(r.shape, dtype=r.dtype).pin_memory() for r in results[b]]
                for h, r in zip(host_out[b], results[b]):
                    h.copy_(r, non_blocking=True)
                    r.record_stream(s_d2h)
                ev_copied[b].record(s_d2h)

    # Timed e2e_steps at a time, three times over, best run reported (all runs listed): the number is PCIe-bound
    # (2 x 326 MB per step, ~43 GB/s per direction with both directions busy on these boxes) and the first run after the
    # warm-up can still pay for the caching allocator growing its pool (a cudaMalloc serialises the three streams).
    e2e_steps = max(4, min(args.steps, 20))
    e2e_pipeline(6)
    e2e_runs = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        e2e_pipeline(e2e_steps)
        torch.cuda.synchronize()
        run_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps  # wall clock incl. the last D2H; device events cannot span 3 streams
        barrier()
        if world > 1:
            t = torch.tensor([run_ms], device=dev, dtype=torch.float64)
            dist_mod.all_reduce(t, op=dist_mod.ReduceOp.MAX)
            run_ms = float(t.item())
        e2e_runs.append(run_ms)
    e2e_ms = min(e2e_runs)
    h2d = sum(h.numel() * h.element_size() for h in host_in)
    d2h = sum(h.numel() * h.element_size() for h in host_out[0])

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if dtype == torch.float32 else "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "dist": args.dist, "taps": "boundary-safe",
                       "l2_policy": "inputs larger than L2 (326 MB of operands per step vs 126 MB L2); no flush",
                       "clips_per_rank_per_step": 1, "parallelism": f"clip-sharded x{world}, no collective",
                       "grad_value_accumulation": "bf16 (opt-in flag)" if half_acc else "f32"},
            "us_fwd": us_fwd, "us_bwd": us_bwd,
            "roofline": {"bound": "hbm", "kernel": "msda_bwd_kernel", "achieved": bytes_b / us_bwd / 1e3, "peak": peak,
                         "unit": "GB/s", "frac": bytes_b / us_bwd / 1e3 / peak, "traffic": ncu_traffic("msda_bwd_kernel"),
                         "algorithmic_bytes": bytes_b, "peak_source": peak_src,
                         "fwd": {"kernel": fwd_kernel, "achieved": bytes_f / us_fwd / 1e3,
                                 "frac": bytes_f / us_fwd / 1e3 / peak, "algorithmic_bytes": bytes_f,
                                 "traffic": ncu_traffic(fwd_kernel)},
                         "fwd_bwd": {"achieved": (bytes_f + bytes_b) / (us_fwd + us_bwd) / 1e3,
                                     "frac": (bytes_f + bytes_b) / (us_fwd + us_bwd) / 1e3 / peak},
                         "note": "contract roofline (HBM). Binding resources measured with ncu + microbenchmarks: forward = SM "
                                 "L1/shared data pipe (~77 % of peak), backward = SM reduction egress (~25 B/clk/SM, 79 % busy) together with "
                                 "the data pipe (71 %); "
                                 "DESIGN.md 3.6, profiles/README.md",
                         "on_chip": on_chip},
            "e2e": {"value": world / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "runs_ms_per_step": [round(x, 3) for x in e2e_runs], "steps_per_run": e2e_steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "clocks": clocks,
        }
        if extra:
            line["also"] = extra
        if world == 1 and not args.no_cpu_baseline:
            sec, cores, sample = cpu_reference_sample(2, args.dist)
            line["cpu_baseline"] = {"value": 1.0 / (sec * T_FRAMES), "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": sample + f"; {sec:.2f} s per sample, x6 per layer-clip"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist_mod.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--dist", default="local", choices=["local", "uniform"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bf16-accumulate", action="store_true",
                    help="with --dtype bf16: accumulate grad_value in bf16 (DEVIS_MSDA_FLAG_BF16_GRAD_VALUE, opt-in)")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
