// Microbenchmark: what does it cost an SM to ACCUMULATE a 128-byte row (32 floats) on chip instead of sending it to
// the L2 as a reduction?  Answers whether grad_value contributions can be pre-aggregated per thread block
// (DESIGN.md 3.6 / 7).  Every mode adds one row per 8-lane group per iteration (4 rows per warp instruction group),
// random rows of a per-CTA window, all warps of the CTA hitting the same window (the access pattern of msda_bwd_kernel).
//   mode 0  red.global.add.v4.f32                      (REDG.E.ADD.F32x4)     -- the shipped scatter
//   mode 1  red.global.add.noftz.v2.bf16x2, 64-B rows  (REDG.E.ADD.BF16x4)    -- DEVIS_MSDA_FLAG_BF16_GRAD_VALUE
//   mode 2  atomicAdd(float) on shared memory          (ATOMS.CAST.SPIN loop)
//   mode 3  atomicAdd(int) on shared memory            (ATOMS.ADD, fixed point)
//   mode 4  LDS.128 + FADD + STS.128, NOT atomic       (lower bound of any shared-memory read-modify-write)
//   mode 5  red.shared::cluster.add.f32 to the CTA's own window through its cluster address (ATOM.E.ADD.F32 generic)
//   mode 6  red.shared::cluster.add.f32 to the PEER CTA's window (cluster of 2)
//   mode 7  atomicAdd(int) with the four words of a lane rotated by its group index: the 4 rows of a warp instruction
//           hit 4 different bank octets (conflict-free)
//   mode 8  mode 7 with the float -> fixed-point conversion of the real kernel (FFMA with a 1.5*2^23 magic + IADD)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_accumulate smem_accumulate.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace cg = cooperative_groups;

constexpr int ITERS = 512;
constexpr int WROWS = 1024;      // window rows per CTA (128 KB of float accumulators)
constexpr int GROWS = 4096;      // global rows per CTA region

__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}

template <int MODE>
__global__ void __launch_bounds__(512) k(float *gbuf, long long *cycles, float *sink)
{
    extern __shared__ __align__(128) float win[];   // [WROWS][32]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 3, j = lane & 7;
    for (int i = threadIdx.x; i < WROWS * 32; i += blockDim.x) win[i] = 0.f;
    unsigned peer = 0;
    if (MODE >= 5) {
        cg::cluster_group cl = cg::this_cluster();
        peer = (MODE == 6) ? (cl.block_rank() ^ 1u) : cl.block_rank();
        cl.sync();
    } else {
        __syncthreads();
    }
    float *gbase = gbuf + (size_t)blockIdx.x * GROWS * 32;
    unsigned s = (blockIdx.x * 64u + warp * 8u + g + 1u) * 2654435761u;
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        s = s * 1664525u + 1013904223u;
        const float a = (float)(s & 7u);
        if (MODE == 0) {
            float *p = gbase + (size_t)((s >> 9) % GROWS) * 32 + j * 4;
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
        } else if (MODE == 1) {
            char *p = reinterpret_cast<char *>(gbase) + (size_t)((s >> 9) % (GROWS * 2)) * 64 + j * 8;
            const __nv_bfloat162 lo = __floats2bfloat162_rn(a, 2.f), hi = __floats2bfloat162_rn(3.f, 4.f);
            asm volatile("red.global.add.noftz.v2.bf16x2 [%0], {%1,%2};" ::"l"(p),
                         "r"(*reinterpret_cast<const unsigned *>(&lo)), "r"(*reinterpret_cast<const unsigned *>(&hi)) : "memory");
        } else {
            float *p = win + ((s >> 9) % WROWS) * 32 + j * 4;
            if (MODE == 2) {
                atomicAdd(p + 0, a); atomicAdd(p + 1, 2.f); atomicAdd(p + 2, 3.f); atomicAdd(p + 3, 4.f);
            } else if (MODE == 3) {
                int *q = reinterpret_cast<int *>(p);
                atomicAdd(q + 0, (int)a); atomicAdd(q + 1, 2); atomicAdd(q + 2, 3); atomicAdd(q + 3, 4);
            } else if (MODE == 7 || MODE == 8) {
                int *q = reinterpret_cast<int *>(win + ((s >> 9) % WROWS) * 32) + j;
                int v0 = (int)(s & 7u), v1 = 2, v2 = 3, v3 = 4;
                if (MODE == 8) {
                    const float sc = (float)(s & 15u);
                    v0 = __float_as_int(fmaf(a, sc, 12582912.f)) - 0x4B400000;
                    v1 = __float_as_int(fmaf(2.f, sc, 12582912.f)) - 0x4B400000;
                    v2 = __float_as_int(fmaf(3.f, sc, 12582912.f)) - 0x4B400000;
                    v3 = __float_as_int(fmaf(4.f, sc, 12582912.f)) - 0x4B400000;
                }
                atomicAdd(q + ((0 + g) & 3) * 8, v0);
                atomicAdd(q + ((1 + g) & 3) * 8, v1);
                atomicAdd(q + ((2 + g) & 3) * 8, v2);
                atomicAdd(q + ((3 + g) & 3) * 8, v3);
            } else if (MODE == 4) {
                float4 v = *reinterpret_cast<float4 *>(p);
                v.x += a; v.y += 2.f; v.z += 3.f; v.w += 4.f;
                *reinterpret_cast<float4 *>(p) = v;
            } else {
                const unsigned addr = mapa((unsigned)__cvta_generic_to_shared(p), peer);
                asm volatile("red.relaxed.cluster.shared::cluster.add.f32 [%0], %1;" ::"r"(addr), "f"(a) : "memory");
                asm volatile("red.relaxed.cluster.shared::cluster.add.f32 [%0], %1;" ::"r"(addr + 4), "f"(2.f) : "memory");
                asm volatile("red.relaxed.cluster.shared::cluster.add.f32 [%0], %1;" ::"r"(addr + 8), "f"(3.f) : "memory");
                asm volatile("red.relaxed.cluster.shared::cluster.add.f32 [%0], %1;" ::"r"(addr + 12), "f"(4.f) : "memory");
            }
        }
    }
    if (MODE >= 5) cg::this_cluster().sync();
    else __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (MODE >= 2) {   // keep the window live
        float acc = 0.f;
        for (int i = threadIdx.x; i < WROWS * 32; i += blockDim.x) acc += win[i];
        if (acc == 12345.678f) sink[0] = acc;
    }
}

template <int MODE>
void run(const char *name, float *gbuf, long long *cyc, float *sink, int threads, int grid)
{
    const size_t smem = (size_t)WROWS * 32 * 4;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = MODE == 6 ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaSuccess;
    for (int rep = 0; rep < 2 && e == cudaSuccess; ++rep) {
        e = cudaLaunchKernelEx(&cfg, k<MODE>, gbuf, cyc, sink);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
    }
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < grid; ++i) avg += (double)h[i];
    avg /= grid;
    const double rows = (double)ITERS * (threads / 32) * 4;
    printf("%-66s grid=%3d threads=%3d  %6.2f cyc per row per SM  (%s)\n", name, grid, threads, avg / rows,
           cudaGetErrorString(e));
}

int main()
{
    float *gbuf, *sink;
    long long *cyc;
    cudaMalloc(&gbuf, (size_t)148 * GROWS * 32 * 4);
    cudaMalloc(&cyc, 148 * 8);
    cudaMalloc(&sink, 16);
    cudaMemset(gbuf, 0, (size_t)148 * GROWS * 32 * 4);
    for (int threads : {256, 512}) {
        const int grid = 148;
        run<0>("red.global.add.v4.f32, 128-B rows", gbuf, cyc, sink, threads, grid);
        run<1>("red.global.add.v2.bf16x2, 64-B rows", gbuf, cyc, sink, threads, grid);
        run<2>("shared atomicAdd(float) (ATOMS.CAST.SPIN)", gbuf, cyc, sink, threads, grid);
        run<3>("shared atomicAdd(int) (ATOMS.ADD)", gbuf, cyc, sink, threads, grid);
        run<4>("shared LDS.128 + FADD + STS.128, not atomic", gbuf, cyc, sink, threads, grid);
        run<7>("shared atomicAdd(int), bank-rotated (conflict-free)", gbuf, cyc, sink, threads, grid);
        run<8>("shared atomicAdd(int), bank-rotated + float->fixed conversion", gbuf, cyc, sink, threads, grid);
        run<5>("red.shared::cluster.add.f32 -> own CTA (ATOM.E.ADD.F32 generic)", gbuf, cyc, sink, threads, grid);
        run<6>("red.shared::cluster.add.f32 -> peer CTA of a 2-cluster", gbuf, cyc, sink, threads, grid);
    }
    return 0;
}
