// Microbenchmark: rate of 128-B row reductions into global memory per SM, via
//   (a) red.global.add.v4.f32 (REDG.E.ADD.F32x4)             -- what msda_bwd_kernel does
//   (b) cp.reduce.async.bulk.global.shared::cta .add.f32      -- TMA bulk reduction from a shared-memory staging row
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_reduce tma_reduce.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 1024;
constexpr int ROWS = 4096;   // 128-B rows per CTA region (512 KB), random targets

__device__ __forceinline__ unsigned hash(unsigned x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// MODE 0: REDG v4, 8 lanes per row, 4 rows per warp instruction
// MODE 1: TMA bulk reduce, one 128-B row per 8-lane group per iteration (4 per warp), ring of DEPTH staging rows
// MODE 2: TMA bulk reduce, 256-B (two adjacent rows) per group per iteration
template <int MODE, int DEPTH>
__global__ void __launch_bounds__(1024) k(float *redbuf, long long *cycles, int bytes_per_op)
{
    extern __shared__ __align__(128) float stage[];   // [warp][DEPTH][4 groups][64 floats]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 3, j = lane & 7;
    float *base = redbuf + (size_t)blockIdx.x * ROWS * 32;
    unsigned s8 = hash(blockIdx.x * 64u + warp * 8u + g + 1u);
    float *mystage = stage + (size_t)warp * DEPTH * 4 * 64;
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        s8 = s8 * 1664525u + 1013904223u;
        const unsigned row = (s8 >> 9) % (ROWS - 1);
        if (MODE == 3) {   // both paths in the same iteration: REDG for one row, TMA bulk reduce for another
            float *p = base + (size_t)row * 32 + j * 4;
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
        }
        if (MODE == 0) {
            float *p = base + (size_t)row * 32 + j * 4;
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
        } else {
            const unsigned row = (MODE == 3) ? ((s8 >> 5) % (ROWS - 1)) : ((s8 >> 9) % (ROWS - 1));
            const int slot = it % DEPTH;
            if (it >= DEPTH && j == 0) {   // the staging row is reused: wait until at most DEPTH-1 groups are pending
                asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory");
            }
            __syncwarp();
            float *srow = mystage + (slot * 4 + g) * 64;
            *reinterpret_cast<float4 *>(srow + j * 4) = make_float4(1.f, 2.f, 3.f, 4.f);
            if (MODE == 2) *reinterpret_cast<float4 *>(srow + 32 + j * 4) = make_float4(1.f, 2.f, 3.f, 4.f);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (j == 0) {
                const unsigned saddr = (unsigned)__cvta_generic_to_shared(srow);
                float *gptr = base + (size_t)row * 32;
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                             ::"l"(gptr), "r"(saddr), "r"(bytes_per_op) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    if (MODE != 0 && j == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE, int DEPTH>
void run(const char *name, float *red, long long *cyc, int threads, int grid, int bytes)
{
    const size_t smem = (size_t)(threads / 32) * DEPTH * 4 * 64 * 4;
    cudaFuncSetAttribute(k<MODE, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<MODE, DEPTH><<<grid, threads, smem>>>(red, cyc, bytes);
    cudaDeviceSynchronize();
    k<MODE, DEPTH><<<grid, threads, smem>>>(red, cyc, bytes);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < grid; ++i) avg += (double)h[i];
    avg /= grid;
    const double rows = (double)ITERS * (threads / 32) * 4 * (bytes / 128) * (MODE == 3 ? 2 : 1);
    printf("%-58s grid=%3d threads=%4d depth=%d  %6.2f cyc per 128-B row per SM  (%s)\n", name, grid, threads, DEPTH,
           avg / rows, cudaGetErrorString(e));
}

int main()
{
    float *red;
    long long *cyc;
    cudaMalloc(&red, (size_t)148 * ROWS * 32 * 4);
    cudaMalloc(&cyc, 148 * 8);
    cudaMemset(red, 0, (size_t)148 * ROWS * 32 * 4);
    for (int grid : {16, 148}) {
        for (int threads : {512, 1024}) {
            run<0, 1>("REDG v4.f32, 4 rows / warp instr", red, cyc, threads, grid, 128);
            run<1, 2>("TMA bulk reduce add.f32, 128 B per op", red, cyc, threads, grid, 128);
            run<1, 4>("TMA bulk reduce add.f32, 128 B per op", red, cyc, threads, grid, 128);
            run<1, 8>("TMA bulk reduce add.f32, 128 B per op", red, cyc, threads, grid, 128);
            run<2, 4>("TMA bulk reduce add.f32, 256 B per op", red, cyc, threads, grid, 256);
            run<3, 4>("REDG row + TMA 128-B row per iteration (2 rows)", red, cyc, threads, grid, 128);
        }
    }
    return 0;
}
