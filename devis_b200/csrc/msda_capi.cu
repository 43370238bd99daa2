// msda_capi.cu -- C ABI (include/devis_msda.h) of the sm_100a multi-scale deformable attention library.
//
// Host side of the drop-in boundary: validates the operands the way the reference's C++ host does
// (cuda/ms_deform_attn_cuda.cu:28-52,93-119), picks a kernel, launches on the caller's stream and
// reports launch failures as error codes.  No torch types, no allocation, no synchronisation.
#include <atomic>
#include <cstring>

#include "../../include/devis_msda.h"
#include "msda_bwd.cuh"
#include "msda_bwd_win.cuh"
#include "msda_bwd_sort.cuh"
#include "msda_common.cuh"
#include "msda_fwd.cuh"
#include "msda_generic.cuh"
#include "tmsda_fused.cuh"

#include "capi_common.h"

using namespace devis;

namespace {

std::atomic<uint64_t> g_launches{0};
thread_local int t_last_cuda_error = 0;
std::atomic<int> g_tuning[16];

}  // namespace

// shared with the other translation units of the library (capi_common.h)
int devis_capi_cuda_fail(cudaError_t e)
{
    t_last_cuda_error = (int)e;
    return DEVIS_MSDA_ERR_CUDA;
}

int devis_capi_check_launch()
{
    g_launches.fetch_add(1, std::memory_order_relaxed);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? DEVIS_MSDA_OK : devis_capi_cuda_fail(e);
}

namespace {

int cuda_fail(cudaError_t e) { return devis_capi_cuda_fail(e); }
int check_launch() { return devis_capi_check_launch(); }

size_t elem_size(int dtype) { return dtype == DEVIS_MSDA_F64 ? 8 : dtype == DEVIS_MSDA_BF16 ? 2 : 4; }

// channel counts served by the grouped-lane kernels: D = 4 * LPG.  They address value with 32-bit byte
// offsets, so tensors of 4 GiB and more go to the generic kernels.
int lanes_per_group(int dtype, const OpDims &d)
{
    if (dtype == DEVIS_MSDA_F64) return 0;
    const unsigned long long bytes = (unsigned long long)d.outer * d.S * d.M * d.D * elem_size(dtype);
    if (bytes >= (1ull << 32)) return 0;
    return d.D == 32 ? 8 : d.D == 16 ? 4 : 0;
}

size_t exchange_bytes(int lpg, int threads)
{
    return (size_t)(threads / 32) * (lpg == 8 ? TapExchange<8>::kBytesPerWarp : TapExchange<4>::kBytesPerWarp);
}

struct LaunchShape {
    int threads, qpg;
};

// Launch shape of the grouped-lane kernels.  Defaults from benchmarks/sweep.py at the DeVIS layer-clip shape
// (profiles/r1f_sweep_*.json): the backward wants 128-thread blocks (84 registers -> 5 blocks = 20 warps per SM
// against 2 x 8 warps with 256 threads: 1428 vs 1536 us), the fp32 forward 256 threads x 2 queries per group
// (527 vs 549 us), the bf16 forward 128 threads.  Tuning keys override (developer tools only); max_qpg is the largest
// QPG the chosen kernel is instantiated with -- anything else would launch a grid that does not cover the queries.
enum ShapeKind { kShapeFwdF32, kShapeFwdBf16, kShapeBwd };

LaunchShape pick_shape(int Lq, ShapeKind kind, int max_qpg, int key_threads, int key_qpg)
{
    LaunchShape s;
    s.threads = g_tuning[key_threads].load();
    s.qpg = g_tuning[key_qpg].load();
    if (s.threads == 0) {
        const int big = kind == kShapeFwdF32 ? 256 : 128;
        s.threads = Lq >= 1024 ? big : Lq >= 128 ? 128 : 64;
    }
    if (s.qpg == 0) s.qpg = (kind == kShapeFwdF32 && Lq >= 2048) ? 2 : 1;
    if (s.threads < 32) s.threads = 32;
    if (s.threads > 256) s.threads = 256;
    s.threads = (s.threads / 32) * 32;
    if (s.qpg != 1 && s.qpg != 2 && s.qpg != 4) s.qpg = 1;
    if (s.qpg > max_qpg) s.qpg = max_qpg;
    return s;
}

template <class SlotSrc>
int launch_forward(const FwdArgs<SlotSrc> &a, int dtype, cudaStream_t st)
{
    const OpDims &d = a.d;
    if (d.outer == 0 || d.Lq == 0) return DEVIS_MSDA_OK;
    size_t smem = (size_t)a.n_slots_total * sizeof(int4);
    const int lpg = lanes_per_group(dtype, d);
    // D = 32 with P % 4 == 0 in every segment can use 4 lanes x 8 channels with 16-byte tap records
    // (msda_fwd8_kernel).  Measured at the DeVIS shape: bf16 -10 % (64-B rows: 0.75 instead of 1.0 data-pipe cycles per
    // row), fp32 +-1 % local / +10 % uniform taps (LDG.256 costs 1.18 wavefronts per 128-B row against 1.01 for
    // LDG.128, which eats what the smaller tap record saves) -> default: bf16 only.  Tuning key 4: 1 never, 2 always.
    const int wide_mode = g_tuning[4].load();
    bool wide = lpg == 8 && (wide_mode == 2 || (wide_mode == 0 && dtype == DEVIS_MSDA_BF16));
    for (int sg = 0; sg < a.n_seg; ++sg) wide = wide && (a.seg[sg].P % 4 == 0);
    if (wide) {
        const LaunchShape s = pick_shape(d.Lq, dtype == DEVIS_MSDA_BF16 ? kShapeFwdBf16 : kShapeFwdF32, 2, 0, 1);
        smem += (size_t)(s.threads / 32) * Tap16::kBytesPerWarp;
        const int qc = s.threads / 4;
        const long long chunks = ((long long)d.Lq + (long long)qc * s.qpg - 1) / ((long long)qc * s.qpg);
        if (chunks * d.M > 0x7fffffffLL) return DEVIS_MSDA_ERR_TOO_LARGE;
        const dim3 grid((unsigned)(chunks * d.M), (unsigned)d.outer);
#define DEVIS_FWD8(BF, QPG) msda_fwd8_kernel<BF, QPG, SlotSrc><<<grid, s.threads, smem, st>>>(a)
        if (dtype == DEVIS_MSDA_BF16) {
            if (s.qpg == 2) DEVIS_FWD8(true, 2);
            else DEVIS_FWD8(true, 1);
        } else {
            if (s.qpg == 2) DEVIS_FWD8(false, 2);
            else DEVIS_FWD8(false, 1);
        }
#undef DEVIS_FWD8
        return check_launch();
    }
    // 8 lanes x 4 channels with 16-byte tap records (msda_fwdc_kernel), D = 32 and P % 4 == 0: -4..5 % against the
    // 32-byte records for fp32 value (default for fp32).  Tuning key 5: 1 always, 2 never.
    const int compact_mode = g_tuning[5].load();
    bool compact = lpg == 8 && (compact_mode == 1 || (compact_mode == 0 && dtype == DEVIS_MSDA_F32));
    for (int sg = 0; sg < a.n_seg; ++sg) compact = compact && (a.seg[sg].P % 4 == 0);
    if (compact) {
        const LaunchShape s = pick_shape(d.Lq, dtype == DEVIS_MSDA_BF16 ? kShapeFwdBf16 : kShapeFwdF32,
                                         dtype == DEVIS_MSDA_BF16 ? 2 : 4, 0, 1);
        smem += (size_t)(s.threads / 32) * Tap16x8::kBytesPerWarp;
        const int qc = s.threads / 8;
        const long long chunks = ((long long)d.Lq + (long long)qc * s.qpg - 1) / ((long long)qc * s.qpg);
        if (chunks * d.M > 0x7fffffffLL) return DEVIS_MSDA_ERR_TOO_LARGE;
        const dim3 grid((unsigned)(chunks * d.M), (unsigned)d.outer);
        if (dtype == DEVIS_MSDA_BF16) {
            if (s.qpg == 2) msda_fwdc_kernel<true, 2, SlotSrc><<<grid, s.threads, smem, st>>>(a);
            else msda_fwdc_kernel<true, 1, SlotSrc><<<grid, s.threads, smem, st>>>(a);
        } else {
            if (s.qpg == 4) msda_fwdc_kernel<false, 4, SlotSrc><<<grid, s.threads, smem, st>>>(a);
            else if (s.qpg == 2) msda_fwdc_kernel<false, 2, SlotSrc><<<grid, s.threads, smem, st>>>(a);
            else msda_fwdc_kernel<false, 1, SlotSrc><<<grid, s.threads, smem, st>>>(a);
        }
        return check_launch();
    }
    if (lpg) {
        const LaunchShape s = pick_shape(d.Lq, dtype == DEVIS_MSDA_BF16 ? kShapeFwdBf16 : kShapeFwdF32, 4, 0, 1);
        smem += exchange_bytes(lpg, s.threads);
        const int qc = s.threads / lpg;
        const long long chunks = ((long long)d.Lq + (long long)qc * s.qpg - 1) / ((long long)qc * s.qpg);
        if (chunks * d.M > 0x7fffffffLL) return DEVIS_MSDA_ERR_TOO_LARGE;
        const dim3 grid((unsigned)(chunks * d.M), (unsigned)d.outer);
#define DEVIS_FWD(BF, LPG, QPG) msda_fwd_kernel<BF, LPG, QPG, SlotSrc><<<grid, s.threads, smem, st>>>(a)
#define DEVIS_FWD_Q(BF, LPG)                        \
    do {                                            \
        if (s.qpg == 4) DEVIS_FWD(BF, LPG, 4);      \
        else if (s.qpg == 2) DEVIS_FWD(BF, LPG, 2); \
        else DEVIS_FWD(BF, LPG, 1);                 \
    } while (0)
        if (dtype == DEVIS_MSDA_BF16) {
            if (lpg == 8) DEVIS_FWD_Q(true, 8);
            else DEVIS_FWD_Q(true, 4);
        } else {
            if (lpg == 8) DEVIS_FWD_Q(false, 8);
            else DEVIS_FWD_Q(false, 4);
        }
#undef DEVIS_FWD_Q
#undef DEVIS_FWD
        return check_launch();
    }
    const int threads = 128, wpc = threads / 32;
    const long long blocks = ((long long)d.Lq * d.M + wpc - 1) / wpc;
    if (blocks > 0x7fffffffLL) return DEVIS_MSDA_ERR_TOO_LARGE;
    const dim3 grid((unsigned)blocks, (unsigned)d.outer);
    if (dtype == DEVIS_MSDA_F32) msda_fwd_generic_kernel<float, SlotSrc><<<grid, threads, smem, st>>>(a);
    else if (dtype == DEVIS_MSDA_F64) msda_fwd_generic_kernel<double, SlotSrc><<<grid, threads, smem, st>>>(a);
    else msda_fwd_generic_kernel<__nv_bfloat16, SlotSrc><<<grid, threads, smem, st>>>(a);
    return check_launch();
}

// deterministic-mode workspace: [int64 accumulators: outer*S*M*D][2 x u32 max slots, padded to 16 B]
size_t det_workspace_bytes(long long outer, long long S, long long M, long long D)
{
    return (size_t)(outer * S * M * D) * sizeof(long long) + 16;
}

// sorted whole-clip backward (msda_bwd_sort.cuh), defined below
bool sort_applicable(const BwdArgs<ClipTable> &a, int dtype, unsigned flags, bool det);
int launch_backward_sort(const BwdArgs<ClipTable> &a, int dtype, bool det, cudaStream_t st);
inline bool sort_applicable(const BwdArgs<DeviceLevels> &, int, unsigned, bool) { return false; }
inline int launch_backward_sort(const BwdArgs<DeviceLevels> &, int, bool, cudaStream_t) { return DEVIS_MSDA_ERR_UNSUPPORTED; }

template <class SlotSrc>
int launch_backward(BwdArgs<SlotSrc> a, int dtype, unsigned flags, void *workspace, size_t workspace_bytes,
                    cudaStream_t st)
{
    const OpDims &d = a.d;
    const bool det = (flags & DEVIS_MSDA_FLAG_DETERMINISTIC) && a.grad_value != nullptr;
    // bf16 grad_value accumulated with packed bf16 reductions: bf16 value, grouped-lane kernels, default mode only
    const bool half_acc = (flags & DEVIS_MSDA_FLAG_BF16_GRAD_VALUE) && a.grad_value != nullptr;
    if (half_acc && (dtype != DEVIS_MSDA_BF16 || det || lanes_per_group(dtype, d) == 0)) return DEVIS_MSDA_ERR_UNSUPPORTED;
    const size_t n_value = (size_t)d.outer * d.S * d.M * d.D;
    float *final_grad_value = reinterpret_cast<float *>(a.grad_value);
    if (det) {
        if (dtype == DEVIS_MSDA_F64) return DEVIS_MSDA_ERR_UNSUPPORTED;
        const size_t need = det_workspace_bytes(d.outer, d.S, d.M, d.D);
        if (!workspace || workspace_bytes < need) return DEVIS_MSDA_ERR_WORKSPACE;
        cudaError_t e = cudaMemsetAsync(workspace, 0, need, st);
        if (e != cudaSuccess) return cuda_fail(e);
        a.det.acc = reinterpret_cast<long long *>(workspace);
        unsigned *slots = reinterpret_cast<unsigned *>(reinterpret_cast<char *>(workspace) + n_value * sizeof(long long));
        a.det.max_bits = slots;
        if (d.outer > 0 && d.Lq > 0) {
            const size_t n_go = (size_t)d.outer * d.Lq * d.M * d.D;
            const int blocks = 148 * 8;
            if (dtype == DEVIS_MSDA_BF16) absmax_kernel<true><<<blocks, 256, 0, st>>>(a.grad_out, n_go, slots);
            else absmax_kernel<false><<<blocks, 256, 0, st>>>(a.grad_out, n_go, slots);
            int rc = check_launch();
            if (rc) return rc;
            for (int sg = 0; sg < a.n_seg; ++sg) {
                const size_t n_aw = (size_t)d.outer * d.Lq * d.M * a.seg[sg].n_slots * a.seg[sg].P;
                absmax_kernel<false><<<blocks, 256, 0, st>>>(a.seg[sg].aw, n_aw, slots + 1);
                rc = check_launch();
                if (rc) return rc;
            }
        }
        a.grad_value = nullptr;  // the float reductions are replaced by the fixed-point ones
    }
    if (a.grad_value) {  // the reference's at::zeros_like(value), ms_deform_attn_cuda.cu:121
        const size_t bytes = (size_t)d.outer * d.S * d.M * d.D * (dtype == DEVIS_MSDA_F64 ? 8 : half_acc ? 2 : 4);
        if (bytes) {
            const cudaError_t e = cudaMemsetAsync(a.grad_value, 0, bytes, st);
            if (e != cudaSuccess) return cuda_fail(e);
        }
    }
    const int lpg = lanes_per_group(dtype, d);
    auto finalize = [&]() -> int {
        if (!det || n_value == 0) return DEVIS_MSDA_OK;
        det_finalize_kernel<<<148 * 8, 256, 0, st>>>(a.det.acc, final_grad_value, n_value, a.det.max_bits, lpg);
        return check_launch();
    };
    if (d.outer == 0 || d.Lq == 0) return finalize();
    size_t smem = (size_t)a.n_slots_total * sizeof(int4);
    if (lpg && det && sort_applicable(a, dtype, flags, true)) {
        // deterministic mode, encoder form: sorted pre-aggregation in 64-bit fixed point (bit-identical to the direct
        // deterministic scatter, ~4x faster: 5x fewer 64-bit reductions leave the SM)
        const int rc = launch_backward_sort(a, dtype, true, st);
        return rc ? rc : finalize();
    }
    if (lpg) {
        const LaunchShape s = pick_shape(d.Lq, kShapeBwd, half_acc ? 1 : 2, 2, 3);
        smem += exchange_bytes(lpg, s.threads);
        const int qc = s.threads / lpg;
        const long long chunks = ((long long)d.Lq + (long long)qc * s.qpg - 1) / ((long long)qc * s.qpg);
        if (chunks * d.M > 0x7fffffffLL) return DEVIS_MSDA_ERR_TOO_LARGE;
        const dim3 grid((unsigned)(chunks * d.M), (unsigned)d.outer);
#define DEVIS_BWD(BF, LPG, QPG) msda_bwd_kernel<BF, LPG, QPG, SlotSrc><<<grid, s.threads, smem, st>>>(a)
#define DEVIS_BWD_Q(BF, LPG)                        \
    do {                                            \
        if (s.qpg == 2) DEVIS_BWD(BF, LPG, 2);      \
        else DEVIS_BWD(BF, LPG, 1);                 \
    } while (0)
        if (half_acc) {
            if (lpg == 8) msda_bwd_kernel<true, 8, 1, SlotSrc, true><<<grid, s.threads, smem, st>>>(a);
            else msda_bwd_kernel<true, 4, 1, SlotSrc, true><<<grid, s.threads, smem, st>>>(a);
        } else if (dtype == DEVIS_MSDA_BF16) {
            if (lpg == 8) DEVIS_BWD_Q(true, 8);
            else DEVIS_BWD_Q(true, 4);
        } else {
            if (lpg == 8) DEVIS_BWD_Q(false, 8);
            else DEVIS_BWD_Q(false, 4);
        }
#undef DEVIS_BWD_Q
#undef DEVIS_BWD
        const int rc = check_launch();
        return rc ? rc : finalize();
    }
    const int threads = 128, wpc = threads / 32;
    const long long blocks = ((long long)d.Lq * d.M + wpc - 1) / wpc;
    if (blocks > 0x7fffffffLL) return DEVIS_MSDA_ERR_TOO_LARGE;
    const dim3 grid((unsigned)blocks, (unsigned)d.outer);
    if (dtype == DEVIS_MSDA_F32) msda_bwd_generic_kernel<float, SlotSrc><<<grid, threads, smem, st>>>(a);
    else if (dtype == DEVIS_MSDA_F64) msda_bwd_generic_kernel<double, SlotSrc><<<grid, threads, smem, st>>>(a);
    else msda_bwd_generic_kernel<__nv_bfloat16, SlotSrc><<<grid, threads, smem, st>>>(a);
    const int rc = check_launch();
    return rc ? rc : finalize();
}

// ---- windowed whole-clip backward (msda_bwd_win.cuh) -----------------------------------------------------------
constexpr size_t kWinWorkspaceBytes = 16;   // [u32 unused][u32 bits of max|attn weight|]

size_t window_smem_bytes(int n_slots_total, int budget_rows)
{
    return (size_t)n_slots_total * sizeof(int4) + 5 * kMaxLevels * sizeof(int) + 16 +
           (size_t)(kWinThreads / 32) * (TapExchange<8>::kBytesPerWarp + 64 * sizeof(unsigned)) +
           (size_t)budget_rows * 32 * sizeof(int);
}

// The windowed kernel serves the encoder form: one query per pyramid pixel (the caller says so by passing a
// query_order), D = 32, fp32 / bf16 value with float grad_value, levels stored back to back, every sampled frame's
// taps a whole number of 8-tap chunks.  EXPERIMENTAL and off by default (tuning key 6 = 2 switches it on): at the DeVIS
// shape it removes 34 - 41 % of the global reductions but executes 33 % more instructions and keeps the L1 data pipe at
// 71 %, so it runs in 1.76 ms against 1.43 ms for msda_bwd_kernel (profiles/README.md, round 1h).
bool window_applicable(const BwdArgs<ClipTable> &a, int dtype, unsigned flags, const void *workspace, size_t workspace_bytes)
{
    if (g_tuning[6].load() != 2) return false;
    if (!a.q_perm || !a.grad_value || (flags & (DEVIS_MSDA_FLAG_DETERMINISTIC | DEVIS_MSDA_FLAG_BF16_GRAD_VALUE))) return false;
    if (!workspace || workspace_bytes < kWinWorkspaceBytes) return false;
    if (lanes_per_group(dtype, a.d) != 8 || a.d.Lq != a.d.S || a.d.outer == 0) return false;
    const ClipTable &tb = a.src;
    int next = 0;
    for (int l = 0; l < tb.L; ++l) {
        if (tb.lsi[l] != next) return false;
        next += tb.H[l] * tb.W[l];
    }
    if (next != a.d.S) return false;
    for (int sg = 0; sg < a.n_seg; ++sg)
        if ((tb.L * a.seg[sg].P) % 8 != 0 || a.seg[sg].n_slots % tb.L != 0) return false;
    return true;
}

int launch_backward_window(const BwdArgs<ClipTable> &a, int dtype, void *workspace, cudaStream_t st)
{
    const OpDims &d = a.d;
    WinArgs w{};
    w.b = a;
    int tiles = 0;
    for (int l = 0; l < a.src.L; ++l) {
        w.tiles_x[l] = (a.src.W[l] + kWinTile - 1) / kWinTile;
        w.tile_start[l] = tiles;
        tiles += w.tiles_x[l] * ((a.src.H[l] + kWinTile - 1) / kWinTile);
    }
    w.n_tiles = tiles;
    w.margin = g_tuning[7].load() > 0 ? g_tuning[7].load() : 6;
    w.budget_rows = g_tuning[8].load() > 0 ? g_tuning[8].load() : 384;
    if (w.budget_rows > 1536) w.budget_rows = 1536;
    unsigned *slots = reinterpret_cast<unsigned *>(workspace);
    w.aw_max_bits = slots + 1;
    cudaError_t e = cudaMemsetAsync(workspace, 0, kWinWorkspaceBytes, st);
    if (e != cudaSuccess) return cuda_fail(e);
    e = cudaMemsetAsync(a.grad_value, 0, (size_t)d.outer * d.S * d.M * d.D * sizeof(float), st);
    if (e != cudaSuccess) return cuda_fail(e);
    for (int sg = 0; sg < a.n_seg; ++sg) {
        const size_t n_aw = (size_t)d.outer * d.Lq * d.M * a.seg[sg].n_slots * a.seg[sg].P;
        absmax_kernel<false><<<148 * 8, 256, 0, st>>>(a.seg[sg].aw, n_aw, slots + 1);
        const int rc = check_launch();
        if (rc) return rc;
    }
    const size_t smem = window_smem_bytes(a.n_slots_total, w.budget_rows);
    const dim3 grid((unsigned)(tiles * d.M), (unsigned)d.outer);
    if (dtype == DEVIS_MSDA_BF16) {
        e = cudaFuncSetAttribute(msda_bwdw_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e);
        msda_bwdw_kernel<true><<<grid, kWinThreads, smem, st>>>(w);
    } else {
        e = cudaFuncSetAttribute(msda_bwdw_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e);
        msda_bwdw_kernel<false><<<grid, kWinThreads, smem, st>>>(w);
    }
    return check_launch();
}

// ---- sorted whole-clip backward (msda_bwd_sort.cuh) ---------------------------------------------------------------
// Serves the encoder form: one query per pyramid pixel (the caller says so by passing a query_order), D = 32, fp32 /
// bf16 value with float grad_value, levels stored back to back, an even number of levels, 4 points per slot.
// Default mode: slower than the direct scatter (data-pipe bound) -> opt-in, tuning key 6 = 3.
// Deterministic mode: ~4x faster than the direct 64-bit scatter and bit-identical to it -> on unless key 6 = 1.
// Key 7 = window margin in pixels (default 6); key 9 = first level with a window + 1 (default: all levels).
bool sort_applicable(const BwdArgs<ClipTable> &a, int dtype, unsigned flags, bool det)
{
    const int mode = g_tuning[6].load();
    if (det ? mode == 1 : mode != 3) return false;          // deterministic mode: on by default; else opt-in
    if (det ? !a.det.acc : (!a.grad_value || (flags & DEVIS_MSDA_FLAG_DETERMINISTIC))) return false;
    if (!a.q_perm || (flags & DEVIS_MSDA_FLAG_BF16_GRAD_VALUE)) return false;
    if (lanes_per_group(dtype, a.d) != 8 || a.d.Lq != a.d.S || a.d.outer == 0) return false;
    const ClipTable &tb = a.src;
    if (tb.L % 2 != 0) return false;
    int next = 0;
    for (int l = 0; l < tb.L; ++l) {
        if (tb.lsi[l] != next || tb.H[l] * (long long)tb.W[l] > 65536) return false;
        next += tb.H[l] * tb.W[l];
    }
    if (next != a.d.S) return false;
    for (int sg = 0; sg < a.n_seg; ++sg)
        if (a.seg[sg].P != 4 || a.seg[sg].n_slots % tb.L != 0) return false;
    if (sort_smem_bytes(a.n_slots_total, (tb.L / 2) * kSortMaxKeys) > 200 * 1024) return false;
    return true;
}

template <bool BF, bool DET>
int launch_sort_kernel(const SortArgs &w, dim3 grid, size_t smem, cudaStream_t st)
{
    const cudaError_t e = cudaFuncSetAttribute(msda_bwds_kernel<BF, DET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e);
    msda_bwds_kernel<BF, DET><<<grid, kSortThreads, smem, st>>>(w);
    return check_launch();
}

// det: the caller (launch_backward) has set up a.det and zero-filled the accumulators, and runs the finalize pass
int launch_backward_sort(const BwdArgs<ClipTable> &a, int dtype, bool det, cudaStream_t st)
{
    const OpDims &d = a.d;
    SortArgs w{};
    w.b = a;
    int tiles = 0;
    for (int l = 0; l < a.src.L; ++l) {
        w.tiles_x[l] = (a.src.W[l] + kSortTile - 1) / kSortTile;
        w.tile_start[l] = tiles;
        tiles += w.tiles_x[l] * ((a.src.H[l] + kSortTile - 1) / kSortTile);
    }
    w.n_tiles = tiles;
    w.margin = g_tuning[7].load() > 0 ? g_tuning[7].load() : 6;
    w.min_level = g_tuning[9].load() > 0 ? g_tuning[9].load() - 1 : 0;
    w.lut_entries = (a.src.L / 2) * kSortMaxKeys;
    if (!det) {
        const cudaError_t e = cudaMemsetAsync(a.grad_value, 0, (size_t)d.outer * d.S * d.M * d.D * sizeof(float), st);
        if (e != cudaSuccess) return cuda_fail(e);
    }
    const size_t smem = sort_smem_bytes(a.n_slots_total, w.lut_entries);
    if ((long long)tiles * d.M > 0x7fffffffLL) return DEVIS_MSDA_ERR_TOO_LARGE;
    const dim3 grid((unsigned)(tiles * d.M), (unsigned)d.outer);
    if (dtype == DEVIS_MSDA_BF16) return det ? launch_sort_kernel<true, true>(w, grid, smem, st) : launch_sort_kernel<true, false>(w, grid, smem, st);
    return det ? launch_sort_kernel<false, true>(w, grid, smem, st) : launch_sort_kernel<false, false>(w, grid, smem, st);
}

int check_common(int outer, int S, int M, int D, int L, int Lq, int dtype)
{
    if (dtype != DEVIS_MSDA_F32 && dtype != DEVIS_MSDA_F64 && dtype != DEVIS_MSDA_BF16) return DEVIS_MSDA_ERR_BAD_DTYPE;
    if (outer < 0 || S < 0 || M <= 0 || D <= 0 || L <= 0 || Lq < 0) return DEVIS_MSDA_ERR_BAD_SHAPE;
    if (outer > 65535) return DEVIS_MSDA_ERR_TOO_LARGE;
    // value rows are indexed with 31-bit integers in every kernel
    if ((long long)outer * S >= (1LL << 31)) return DEVIS_MSDA_ERR_TOO_LARGE;
    return DEVIS_MSDA_OK;
}

int fill_clip_table(ClipTable &tb, const int64_t *shapes, const int64_t *lsi, const int32_t *frames, int T, int S,
                    int L, int Wt)
{
    if (L > kMaxLevels || (long long)T * Wt > kMaxFrameTable || T > 255) return DEVIS_MSDA_ERR_TOO_LARGE;
    std::memset(&tb, 0, sizeof(tb));
    tb.L = L;
    tb.Wt = Wt;
    for (int l = 0; l < L; ++l) {
        const int64_t H = shapes[2 * l], W = shapes[2 * l + 1], st = lsi[l];
        if (H <= 0 || W <= 0 || st < 0 || st + H * W > S) return DEVIS_MSDA_ERR_BAD_SHAPE;
        tb.H[l] = (int)H;
        tb.W[l] = (int)W;
        tb.lsi[l] = (int)st;
    }
    for (int i = 0; i < T * Wt; ++i) {
        if (frames[i] < 0 || frames[i] >= T) return DEVIS_MSDA_ERR_BAD_FRAME_TABLE;
        tb.frame[i] = (uint8_t)frames[i];
    }
    return DEVIS_MSDA_OK;
}

}  // namespace

extern "C" {

int devis_msda_abi_version(void) { return DEVIS_MSDA_ABI_VERSION; }

const char *devis_msda_error_string(int code)
{
    switch (code) {
        case DEVIS_MSDA_OK: return "ok";
        case DEVIS_MSDA_ERR_NULL_POINTER: return "null pointer for a non-empty tensor";
        case DEVIS_MSDA_ERR_BAD_SHAPE: return "invalid dimension";
        case DEVIS_MSDA_ERR_BAD_DTYPE: return "unsupported dtype code";
        case DEVIS_MSDA_ERR_BATCH_STEP: return "batch must be divisible by min(batch, im2col_step)";
        case DEVIS_MSDA_ERR_TOO_LARGE: return "tensor too large for the kernels' indexing";
        case DEVIS_MSDA_ERR_WORKSPACE: return "workspace missing or too small";
        case DEVIS_MSDA_ERR_CUDA: return "CUDA runtime error (see devis_msda_last_cuda_error)";
        case DEVIS_MSDA_ERR_UNSUPPORTED: return "no kernel for this request in this build";
        case DEVIS_MSDA_ERR_BAD_FRAME_TABLE: return "frame table entry outside [0, num_frames)";
        default: return "unknown error code";
    }
}

int devis_msda_last_cuda_error(void) { return t_last_cuda_error; }
uint64_t devis_msda_launch_count(void) { return g_launches.load(); }

int devis_msda_set_tuning(int key, int value)
{
    if (key < 0 || key >= 16) return DEVIS_MSDA_ERR_BAD_SHAPE;
    g_tuning[key].store(value);
    return DEVIS_MSDA_OK;
}

int devis_msda_forward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                       const void *sampling_loc, const void *attn_weight, void *output, int batch,
                       int spatial_size, int num_heads, int channels, int num_levels, int num_query,
                       int num_point, int im2col_step, int dtype, void *stream)
{
    int rc = check_common(batch, spatial_size, num_heads, channels, num_levels, num_query, dtype);
    if (rc) return rc;
    if (num_point <= 0 || im2col_step <= 0 || num_levels > kMaxSlots) return DEVIS_MSDA_ERR_BAD_SHAPE;
    if (batch > 0 && batch % (batch < im2col_step ? batch : im2col_step) != 0) return DEVIS_MSDA_ERR_BATCH_STEP;
    const bool empty = batch == 0 || num_query == 0;
    if (!empty && (!value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !output))
        return DEVIS_MSDA_ERR_NULL_POINTER;
    FwdArgs<DeviceLevels> a{};
    a.value = value;
    a.out = output;
    a.seg[0] = Segment{sampling_loc, attn_weight, nullptr, nullptr, num_levels, num_point};
    a.n_seg = 1;
    a.n_slots_total = num_levels;
    a.src = DeviceLevels{spatial_shapes, level_start_index};
    a.d = OpDims{batch, spatial_size, num_heads, channels, num_query};
    a.q_perm = nullptr;
    return launch_forward(a, dtype, (cudaStream_t)stream);
}

size_t devis_msda_backward_workspace_bytes(int batch, int spatial_size, int num_heads, int channels, int, int, int,
                                           int dtype, unsigned flags)
{
    if (!(flags & DEVIS_MSDA_FLAG_DETERMINISTIC) || (flags & DEVIS_MSDA_FLAG_NO_GRAD_VALUE) || dtype == DEVIS_MSDA_F64)
        return 0;
    return det_workspace_bytes(batch, spatial_size, num_heads, channels);
}

int devis_msda_backward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                        const void *sampling_loc, const void *attn_weight, const void *grad_output,
                        void *grad_value, void *grad_sampling_loc, void *grad_attn_weight, int batch,
                        int spatial_size, int num_heads, int channels, int num_levels, int num_query,
                        int num_point, int im2col_step, int dtype, unsigned flags, void *workspace,
                        size_t workspace_bytes, void *stream)
{
    int rc = check_common(batch, spatial_size, num_heads, channels, num_levels, num_query, dtype);
    if (rc) return rc;
    if (num_point <= 0 || im2col_step <= 0 || num_levels > kMaxSlots) return DEVIS_MSDA_ERR_BAD_SHAPE;
    if (batch > 0 && batch % (batch < im2col_step ? batch : im2col_step) != 0) return DEVIS_MSDA_ERR_BATCH_STEP;
    const bool want_gv = !(flags & DEVIS_MSDA_FLAG_NO_GRAD_VALUE);
    const bool empty = batch == 0 || num_query == 0;
    if (!empty && (!value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight ||
                   !grad_output || !grad_sampling_loc || !grad_attn_weight))
        return DEVIS_MSDA_ERR_NULL_POINTER;
    if (want_gv && !grad_value && (size_t)batch * spatial_size > 0) return DEVIS_MSDA_ERR_NULL_POINTER;
    BwdArgs<DeviceLevels> a{};
    a.value = value;
    a.grad_out = grad_output;
    a.grad_value = want_gv ? grad_value : nullptr;
    a.seg[0] = Segment{sampling_loc, attn_weight, grad_sampling_loc, grad_attn_weight, num_levels, num_point};
    a.n_seg = 1;
    a.n_slots_total = num_levels;
    a.src = DeviceLevels{spatial_shapes, level_start_index};
    a.d = OpDims{batch, spatial_size, num_heads, channels, num_query};
    a.q_perm = nullptr;
    return launch_backward(a, dtype, flags, workspace, workspace_bytes, (cudaStream_t)stream);
}

int devis_tmsda_forward(const void *value, const int64_t *spatial_shapes_host,
                        const int64_t *level_start_index_host, const int32_t *frame_table_host,
                        const void *loc_curr, const void *aw_curr, const void *loc_temporal,
                        const void *aw_temporal, void *output, const int32_t *query_order, int num_frames,
                        int spatial_size, int num_heads, int channels, int num_levels, int num_query,
                        int n_curr_points, int n_temporal_points, int t_window, int dtype, void *stream)
{
    int rc = check_common(num_frames, spatial_size, num_heads, channels, num_levels, num_query, dtype);
    if (rc) return rc;
    if (n_curr_points <= 0 || n_temporal_points < 0 || t_window < 0) return DEVIS_MSDA_ERR_BAD_SHAPE;
    const bool temporal = t_window > 0 && n_temporal_points > 0;
    if (!spatial_shapes_host || !level_start_index_host || (temporal && !frame_table_host))
        return DEVIS_MSDA_ERR_NULL_POINTER;
    const bool empty = num_frames == 0 || num_query == 0;
    if (!empty && (!value || !loc_curr || !aw_curr || !output || (temporal && (!loc_temporal || !aw_temporal))))
        return DEVIS_MSDA_ERR_NULL_POINTER;
    FwdArgs<ClipTable> a{};
    rc = fill_clip_table(a.src, spatial_shapes_host, level_start_index_host, frame_table_host, num_frames,
                         spatial_size, num_levels, temporal ? t_window : 0);
    if (rc) return rc;
    a.value = value;
    a.out = output;
    a.seg[0] = Segment{loc_curr, aw_curr, nullptr, nullptr, num_levels, n_curr_points};
    a.n_seg = 1;
    a.n_slots_total = num_levels;
    if (temporal) {
        a.seg[1] = Segment{loc_temporal, aw_temporal, nullptr, nullptr, t_window * num_levels, n_temporal_points};
        a.n_seg = 2;
        a.n_slots_total += t_window * num_levels;
    }
    if (a.n_slots_total > kMaxSlots) return DEVIS_MSDA_ERR_TOO_LARGE;
    a.d = OpDims{num_frames, spatial_size, num_heads, channels, num_query};
    a.q_perm = query_order;
    return launch_forward(a, dtype, (cudaStream_t)stream);
}

size_t devis_tmsda_backward_workspace_bytes(int num_frames, int spatial_size, int num_heads, int channels, int, int,
                                            int, int, int, int dtype, unsigned flags)
{
    if ((flags & DEVIS_MSDA_FLAG_NO_GRAD_VALUE) || dtype == DEVIS_MSDA_F64) return 0;
    if (!(flags & DEVIS_MSDA_FLAG_DETERMINISTIC))   // the experimental windowed kernel keeps max|attn weight| there
        return g_tuning[6].load() == 2 ? kWinWorkspaceBytes : 0;
    return det_workspace_bytes(num_frames, spatial_size, num_heads, channels);
}

int devis_tmsda_backward(const void *value, const int64_t *spatial_shapes_host,
                         const int64_t *level_start_index_host, const int32_t *frame_table_host,
                         const void *loc_curr, const void *aw_curr, const void *loc_temporal,
                         const void *aw_temporal, const void *grad_output, void *grad_value,
                         void *grad_loc_curr, void *grad_aw_curr, void *grad_loc_temporal,
                         void *grad_aw_temporal, const int32_t *query_order, int num_frames, int spatial_size,
                         int num_heads, int channels, int num_levels, int num_query, int n_curr_points,
                         int n_temporal_points, int t_window, int dtype, unsigned flags, void *workspace,
                         size_t workspace_bytes, void *stream)
{
    int rc = check_common(num_frames, spatial_size, num_heads, channels, num_levels, num_query, dtype);
    if (rc) return rc;
    if (n_curr_points <= 0 || n_temporal_points < 0 || t_window < 0) return DEVIS_MSDA_ERR_BAD_SHAPE;
    const bool temporal = t_window > 0 && n_temporal_points > 0;
    if (!spatial_shapes_host || !level_start_index_host || (temporal && !frame_table_host))
        return DEVIS_MSDA_ERR_NULL_POINTER;
    const bool want_gv = !(flags & DEVIS_MSDA_FLAG_NO_GRAD_VALUE);
    const bool empty = num_frames == 0 || num_query == 0;
    if (!empty && (!value || !loc_curr || !aw_curr || !grad_output || !grad_loc_curr || !grad_aw_curr ||
                   (temporal && (!loc_temporal || !aw_temporal || !grad_loc_temporal || !grad_aw_temporal))))
        return DEVIS_MSDA_ERR_NULL_POINTER;
    if (want_gv && !grad_value && (size_t)num_frames * spatial_size > 0) return DEVIS_MSDA_ERR_NULL_POINTER;
    BwdArgs<ClipTable> a{};
    rc = fill_clip_table(a.src, spatial_shapes_host, level_start_index_host, frame_table_host, num_frames,
                         spatial_size, num_levels, temporal ? t_window : 0);
    if (rc) return rc;
    a.value = value;
    a.grad_out = grad_output;
    a.grad_value = want_gv ? grad_value : nullptr;
    a.seg[0] = Segment{loc_curr, aw_curr, grad_loc_curr, grad_aw_curr, num_levels, n_curr_points};
    a.n_seg = 1;
    a.n_slots_total = num_levels;
    if (temporal) {
        a.seg[1] = Segment{loc_temporal, aw_temporal, grad_loc_temporal, grad_aw_temporal, t_window * num_levels,
                           n_temporal_points};
        a.n_seg = 2;
        a.n_slots_total += t_window * num_levels;
    }
    if (a.n_slots_total > kMaxSlots) return DEVIS_MSDA_ERR_TOO_LARGE;
    a.d = OpDims{num_frames, spatial_size, num_heads, channels, num_query};
    a.q_perm = query_order;
    if (sort_applicable(a, dtype, flags, false)) return launch_backward_sort(a, dtype, false, (cudaStream_t)stream);
    if (window_applicable(a, dtype, flags, workspace, workspace_bytes))
        return launch_backward_window(a, dtype, workspace, (cudaStream_t)stream);
    return launch_backward(a, dtype, flags, workspace, workspace_bytes, (cudaStream_t)stream);
}

static int fill_fused(FusedArgs &a, const void *value, const int64_t *shapes, const int64_t *lsi, const int32_t *frames,
                      const void *ref, const void *off_c, const void *logit_c, const void *off_t, const void *logit_t,
                      const int32_t *order, int T, int S, int M, int D, int L, int Lq, int Pc, int Pt, int Wt, int dtype)
{
    int rc = check_common(T, S, M, D, L, Lq, dtype);
    if (rc) return rc;
    if (Pc <= 0 || Pt < 0 || Wt < 0) return DEVIS_MSDA_ERR_BAD_SHAPE;
    if (dtype == DEVIS_MSDA_F64 || D != 32 || Pc % 4 != 0 || (Wt > 0 && Pt > 0 && Pt % 4 != 0))
        return DEVIS_MSDA_ERR_UNSUPPORTED;
    if ((unsigned long long)T * S * M * D * elem_size(dtype) >= (1ull << 32)) return DEVIS_MSDA_ERR_UNSUPPORTED;
    const bool temporal = Wt > 0 && Pt > 0;
    if (!shapes || !lsi || (temporal && !frames)) return DEVIS_MSDA_ERR_NULL_POINTER;
    const bool empty = T == 0 || Lq == 0;
    if (!empty && (!value || !ref || !off_c || !logit_c || (temporal && (!off_t || !logit_t))))
        return DEVIS_MSDA_ERR_NULL_POINTER;
    rc = fill_clip_table(a.src, shapes, lsi, frames, T, S, L, temporal ? Wt : 0);
    if (rc) return rc;
    a.value = value;
    a.ref = reinterpret_cast<const float *>(ref);
    a.off[0] = reinterpret_cast<const float *>(off_c);
    a.logit[0] = reinterpret_cast<const float *>(logit_c);
    a.off[1] = reinterpret_cast<const float *>(off_t);
    a.logit[1] = reinterpret_cast<const float *>(logit_t);
    a.n_slots[0] = L;
    a.P[0] = Pc;
    a.n_slots[1] = temporal ? Wt * L : 0;
    a.P[1] = temporal ? Pt : 1;
    a.n_seg = temporal ? 2 : 1;
    if (a.n_slots[0] + a.n_slots[1] > kMaxSlots) return DEVIS_MSDA_ERR_TOO_LARGE;
    a.d = OpDims{T, S, M, D, Lq};
    a.q_perm = order;
    return DEVIS_MSDA_OK;
}

static int fused_grid(const FusedArgs &a, int key_threads, dim3 &grid, int &threads, size_t &smem, int qpg = 1)
{
    threads = g_tuning[key_threads].load();
    if (threads == 0) threads = a.d.Lq >= 1024 ? (key_threads == 2 ? 128 : 256) : a.d.Lq >= 128 ? 128 : 64;   // see pick_shape
    threads = threads < 32 ? 32 : threads > 256 ? 256 : (threads / 32) * 32;
    const long long per_block = (long long)(threads / 8) * qpg;
    const long long chunks = ((long long)a.d.Lq + per_block - 1) / per_block;
    if (chunks * a.d.M > 0x7fffffffLL) return DEVIS_MSDA_ERR_TOO_LARGE;
    grid = dim3((unsigned)(chunks * a.d.M), (unsigned)a.d.outer);
    smem = (size_t)(a.n_slots[0] + a.n_slots[1]) * sizeof(int4) + exchange_bytes(8, threads);
    return DEVIS_MSDA_OK;
}

int devis_tmsda_fused_forward(const void *value, const int64_t *spatial_shapes_host,
                              const int64_t *level_start_index_host, const int32_t *frame_table_host,
                              const void *ref, const void *off_curr, const void *logit_curr,
                              const void *off_temporal, const void *logit_temporal, void *output,
                              const int32_t *query_order, int num_frames, int spatial_size, int num_heads,
                              int channels, int num_levels, int num_query, int n_curr_points,
                              int n_temporal_points, int t_window, int dtype, void *stream)
{
    FusedArgs a{};
    int rc = fill_fused(a, value, spatial_shapes_host, level_start_index_host, frame_table_host, ref, off_curr,
                        logit_curr, off_temporal, logit_temporal, query_order, num_frames, spatial_size, num_heads,
                        channels, num_levels, num_query, n_curr_points, n_temporal_points, t_window, dtype);
    if (rc) return rc;
    if (num_frames == 0 || num_query == 0) return DEVIS_MSDA_OK;
    if (!output) return DEVIS_MSDA_ERR_NULL_POINTER;
    a.out = output;
    dim3 grid;
    int threads;
    size_t smem;
    // bf16 value: four lanes x 8 channels per (query, head) like msda_fwd8_kernel (tuning key 4: 1 never, 2 always)
    const int wide_mode = g_tuning[4].load();
    if (wide_mode == 2 || (wide_mode == 0 && dtype == DEVIS_MSDA_BF16)) {
        threads = g_tuning[0].load();
        if (threads == 0) threads = num_query >= 1024 ? 128 : 64;
        threads = threads < 32 ? 32 : threads > 256 ? 256 : (threads / 32) * 32;
        const long long chunks = ((long long)num_query + threads / 4 - 1) / (threads / 4);
        if (chunks * num_heads > 0x7fffffffLL) return DEVIS_MSDA_ERR_TOO_LARGE;
        grid = dim3((unsigned)(chunks * num_heads), (unsigned)num_frames);
        smem = (size_t)(a.n_slots[0] + a.n_slots[1]) * sizeof(int4) + (size_t)(threads / 32) * Tap16::kBytesPerWarp;
        if (dtype == DEVIS_MSDA_BF16) tmsda_fused_fwd8_kernel<true><<<grid, threads, smem, (cudaStream_t)stream>>>(a);
        else tmsda_fused_fwd8_kernel<false><<<grid, threads, smem, (cudaStream_t)stream>>>(a);
        return check_launch();
    }
    // two queries per lane group for fp32 at encoder sizes, like msda_fwdc_kernel (tuning key 1 overrides)
    int qpg = g_tuning[1].load();
    if (qpg != 1 && qpg != 2) qpg = (dtype == DEVIS_MSDA_F32 && num_query >= 2048) ? 2 : 1;
    rc = fused_grid(a, 0, grid, threads, smem, qpg);
    if (rc) return rc;
    if (dtype == DEVIS_MSDA_BF16) {
        if (qpg == 2) tmsda_fused_fwd_kernel<true, 2><<<grid, threads, smem, (cudaStream_t)stream>>>(a);
        else tmsda_fused_fwd_kernel<true, 1><<<grid, threads, smem, (cudaStream_t)stream>>>(a);
    } else {
        if (qpg == 2) tmsda_fused_fwd_kernel<false, 2><<<grid, threads, smem, (cudaStream_t)stream>>>(a);
        else tmsda_fused_fwd_kernel<false, 1><<<grid, threads, smem, (cudaStream_t)stream>>>(a);
    }
    return check_launch();
}

int devis_tmsda_fused_backward(const void *value, const int64_t *spatial_shapes_host,
                               const int64_t *level_start_index_host, const int32_t *frame_table_host,
                               const void *ref, const void *off_curr, const void *logit_curr,
                               const void *off_temporal, const void *logit_temporal, const void *grad_output,
                               void *grad_value, void *grad_off_curr, void *grad_logit_curr,
                               void *grad_off_temporal, void *grad_logit_temporal, const int32_t *query_order,
                               int num_frames, int spatial_size, int num_heads, int channels, int num_levels,
                               int num_query, int n_curr_points, int n_temporal_points, int t_window, int dtype,
                               unsigned flags, void *stream)
{
    FusedArgs a{};
    int rc = fill_fused(a, value, spatial_shapes_host, level_start_index_host, frame_table_host, ref, off_curr,
                        logit_curr, off_temporal, logit_temporal, query_order, num_frames, spatial_size, num_heads,
                        channels, num_levels, num_query, n_curr_points, n_temporal_points, t_window, dtype);
    if (rc) return rc;
    if (flags & DEVIS_MSDA_FLAG_DETERMINISTIC) return DEVIS_MSDA_ERR_UNSUPPORTED;
    const bool want_gv = !(flags & DEVIS_MSDA_FLAG_NO_GRAD_VALUE);
    const bool half_acc = want_gv && (flags & DEVIS_MSDA_FLAG_BF16_GRAD_VALUE);
    if (half_acc && dtype != DEVIS_MSDA_BF16) return DEVIS_MSDA_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    if (want_gv) {
        const size_t bytes = (size_t)num_frames * spatial_size * num_heads * channels * (half_acc ? 2 : sizeof(float));
        if (bytes && !grad_value) return DEVIS_MSDA_ERR_NULL_POINTER;
        if (bytes) {
            const cudaError_t e = cudaMemsetAsync(grad_value, 0, bytes, st);
            if (e != cudaSuccess) return cuda_fail(e);
        }
    }
    if (num_frames == 0 || num_query == 0) return DEVIS_MSDA_OK;
    if (!grad_output || !grad_off_curr || !grad_logit_curr || (a.n_seg > 1 && (!grad_off_temporal || !grad_logit_temporal)))
        return DEVIS_MSDA_ERR_NULL_POINTER;
    a.grad_out = grad_output;
    a.grad_value = want_gv ? grad_value : nullptr;
    a.grad_off[0] = reinterpret_cast<float *>(grad_off_curr);
    a.grad_logit[0] = reinterpret_cast<float *>(grad_logit_curr);
    a.grad_off[1] = reinterpret_cast<float *>(grad_off_temporal);
    a.grad_logit[1] = reinterpret_cast<float *>(grad_logit_temporal);
    dim3 grid;
    int threads;
    size_t smem;
    rc = fused_grid(a, 2, grid, threads, smem);
    if (rc) return rc;
    {   // (weight, d/d weight) of every tap parked in shared memory until the row's softmax-backward sum is known
        int iters = 0;
        for (int sg = 0; sg < a.n_seg; ++sg) iters += (a.n_slots[sg] * a.P[sg] + 7) / 8;
        const size_t park = (size_t)iters * threads * sizeof(float2);
        if (smem + park <= 48 * 1024) {
            a.park_iters = iters;
            smem += park;
        }
    }
    if (half_acc) tmsda_fused_bwd_kernel<true, true><<<grid, threads, smem, st>>>(a);
    else if (dtype == DEVIS_MSDA_BF16) tmsda_fused_bwd_kernel<true><<<grid, threads, smem, st>>>(a);
    else tmsda_fused_bwd_kernel<false><<<grid, threads, smem, st>>>(a);
    return check_launch();
}

}  // extern "C"
