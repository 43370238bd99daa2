// Microbenchmark: cost (SM cycles per warp instruction at saturation) of the L1/shared access patterns the
// deformable-attention kernels can choose between.  All data is L1/L2 resident; 148 CTAs x 1024 threads.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1_patterns l1_patterns.cu && ./l1_patterns
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;
constexpr int ROWS = 512;  // rows of 128 B -> 64 KB footprint per CTA, L1 resident

__device__ __forceinline__ unsigned hash(unsigned x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

template <int MODE>
__global__ void __launch_bounds__(1024) k(const float *__restrict__ g, float *out, float *redbuf, long long *cycles)
{
    __shared__ float sm[ROWS * 32 / 4];  // 16 KB
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < ROWS * 8; i += blockDim.x) sm[i] = g[i];
    __syncthreads();
    const char *base = reinterpret_cast<const char *>(g) + (size_t)blockIdx.x * ROWS * 128;
    float acc = 0.f;
    // per-GROUP pseudo-random row stream, 2 ALU ops per iteration (LCG + bit-field extract)
    const int gid8 = lane >> 3, gid4 = lane >> 2, gid16 = lane >> 4;
    unsigned s8 = hash(blockIdx.x * 64u + warp * 8u + gid8 + 1u), s4 = hash(blockIdx.x * 640u + warp * 8u + gid4 + 7u);
    unsigned s16 = hash(blockIdx.x * 77u + warp * 2u + gid16 + 3u), s1 = hash(blockIdx.x * 99u + warp + 5u);
    unsigned seed = s1;
    const long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; ++it) {
        s8 = s8 * 1664525u + 1013904223u; s4 = s4 * 1664525u + 1013904223u;
        s16 = s16 * 1664525u + 1013904223u; s1 = s1 * 1664525u + 1013904223u; seed = s1;
#define ROW8 ((s8 >> 9) % ROWS)
#define ROW4 ((s4 >> 9) % ROWS)
#define ROW16 ((s16 >> 9) % ROWS)
#define ROW1 ((s1 >> 9) % ROWS)
        if (MODE == 0) {  // LDG.128, 8 lanes per row, 4 distinct rows per warp (the value gather)
            const unsigned row = ROW8;
            const float4 v = __ldg(reinterpret_cast<const float4 *>(base + row * 128) + (lane & 7));
            acc += v.x + v.y + v.z + v.w;
        } else if (MODE == 1) {  // LDG.64 broadcast: 8-lane groups read the same 8 B, 4 rows
            const unsigned row = ROW8;
            const float2 v = __ldg(reinterpret_cast<const float2 *>(base + row * 128));
            acc += v.x + v.y;
        } else if (MODE == 2) {  // LDG.64 broadcast: 4-lane groups, 8 rows
            const unsigned row = ROW4;
            const float2 v = __ldg(reinterpret_cast<const float2 *>(base + row * 128));
            acc += v.x + v.y;
        } else if (MODE == 3) {  // LDG.32 broadcast: 4-lane groups, 8 rows
            const unsigned row = ROW4;
            acc += __ldg(reinterpret_cast<const float *>(base + row * 128));
        } else if (MODE == 4) {  // LDG.256: 4 lanes per row, 8 rows per warp
            const unsigned row = ROW4;
            float a0, a1, a2, a3, a4, a5, a6, a7;
            asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(a0), "=f"(a1), "=f"(a2), "=f"(a3), "=f"(a4), "=f"(a5), "=f"(a6), "=f"(a7)
                         : "l"(base + row * 128 + (lane & 3) * 32));
            acc += a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
        } else if (MODE == 5) {  // LDG.128: 4 lanes cover HALF a row (64 B), 8 rows per warp
            const unsigned row = ROW4;
            const float4 v = __ldg(reinterpret_cast<const float4 *>(base + row * 128) + (lane & 3));
            acc += v.x + v.y + v.z + v.w;
        } else if (MODE == 6) {  // LDS.128 broadcast: 8-lane groups, 4 distinct 16 B
            const unsigned r = ((s8 >> 9) % (ROWS * 2));
            const float4 v = *reinterpret_cast<const float4 *>(sm + r * 4);
            acc += v.x + v.y + v.z + v.w;
        } else if (MODE == 7) {  // LDS.32 broadcast: 8-lane groups, 4 distinct words
            const unsigned r = ((s8 >> 9) % (ROWS * 8));
            acc += sm[r];
        } else if (MODE == 8) {  // LDS.128 broadcast: 4-lane groups, 8 distinct 16 B
            const unsigned r = ((s4 >> 9) % (ROWS * 2));
            const float4 v = *reinterpret_cast<const float4 *>(sm + r * 4);
            acc += v.x + v.y + v.z + v.w;
        } else if (MODE == 9) {  // SHFL width 8
            acc += __shfl_sync(0xffffffffu, acc + (float)seed, it & 7, 8);
        } else if (MODE == 10) {  // LDS.128 no broadcast, conflict-free (512 B per warp)
            const float4 v = *reinterpret_cast<const float4 *>(sm + ((it * 32 + lane) % (ROWS * 2)) * 4);
            acc += v.x + v.y + v.z + v.w;
        } else if (MODE == 11) {  // RED.128: 8 lanes per row, 4 rows (the grad_value scatter)
            const unsigned row = ROW8;
            float *p = redbuf + ((size_t)blockIdx.x * ROWS + row) * 32 + (lane & 7) * 4;
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
        } else if (MODE == 12) {  // RED.64 (v2): 16 lanes per row, 2 rows
            const unsigned row = ROW16;
            float *p = redbuf + ((size_t)blockIdx.x * ROWS + row) * 32 + (lane & 15) * 2;
            asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(1.f), "f"(2.f) : "memory");
        } else if (MODE == 13) {  // RED.32: 32 lanes one row
            const unsigned row = ROW1;
            float *p = redbuf + ((size_t)blockIdx.x * ROWS + row) * 32 + lane;
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(1.f) : "memory");
        } else if (MODE == 14) {  // LDG.128 full row, all 4 groups the SAME row (1 line per warp)
            const unsigned row = ROW1;
            const float4 v = __ldg(reinterpret_cast<const float4 *>(base + row * 128) + (lane & 7));
            acc += v.x + v.y + v.z + v.w;
        } else if (MODE == 15) {  // LDG.64, 8 lanes per row-quarter: 4 rows x 64 B contiguous (bf16 value gather)
            const unsigned row = ROW8;
            const float2 v = __ldg(reinterpret_cast<const float2 *>(base + row * 128) + (lane & 7));
            acc += v.x + v.y;
        } else if (MODE == 17) {  // LDS.128 gather: 8 lanes per 128 B row, 4 random rows (smem-staged value tile)
            const unsigned r = ((s8 >> 9) % (ROWS / 4));
            const float4 v = *reinterpret_cast<const float4 *>(sm + r * 32 + (lane & 7) * 4);
            acc += v.x + v.y + v.z + v.w;
        } else if (MODE == 18) {  // 2 x LDS.128 gather: 4 lanes per 128 B row (32 B per lane), 8 random rows
            const unsigned r = ((s4 >> 9) % (ROWS / 4));
            const float4 v = *reinterpret_cast<const float4 *>(sm + r * 32 + (lane & 3) * 4);
            const float4 u = *reinterpret_cast<const float4 *>(sm + r * 32 + 16 + (lane & 3) * 4);
            acc += v.x + v.y + v.z + v.w + u.x + u.y + u.z + u.w;
        } else if (MODE == 19) {  // LDG.128 4 rows + LDS.128 gather 4 rows in the same iteration (do the pipes add up?)
            const unsigned row = ROW8;
            const float4 v = __ldg(reinterpret_cast<const float4 *>(base + row * 128) + (lane & 7));
            const unsigned r = ((s4 >> 9) % (ROWS / 4));
            const float4 u = *reinterpret_cast<const float4 *>(sm + r * 32 + (lane & 7) * 4);
            acc += v.x + v.y + v.z + v.w + u.x + u.y + u.z + u.w;
        } else if (MODE == 20) {  // 2 x LDG.128: 4 lanes per row, lane reads bytes [32j,32j+16) and [32j+16,32j+32): 8 rows
            const unsigned row = ROW4;
            const float4 *p = reinterpret_cast<const float4 *>(base + row * 128 + (lane & 3) * 32);
            const float4 v = __ldg(p), u = __ldg(p + 1);
            acc += v.x + v.y + v.z + v.w + u.x + u.y + u.z + u.w;
        } else if (MODE >= 21 && MODE <= 24) {  // LDG.128 8 lanes/row with (MODE - 20) of the 4 lane groups predicated OFF
            const unsigned row = ROW8;
            float4 v = make_float4(acc, acc, acc, acc);
            const unsigned live = gid8 >= (MODE - 20);       // groups 0 .. MODE-21 are dead
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];\n\t}"
                         : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
                         : "l"(base + row * 128 + (lane & 7) * 16), "r"(live));
            acc += v.x + v.y + v.z + v.w;
        } else if (MODE == 25) {  // as 22 (two groups dead), but which two varies per iteration and warp
            const unsigned row = ROW8;
            float4 v = make_float4(acc, acc, acc, acc);
            const unsigned live = ((s1 >> (12 + gid8)) & 1u);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];\n\t}"
                         : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
                         : "l"(base + row * 128 + (lane & 7) * 16), "r"(live));
            acc += v.x + v.y + v.z + v.w;
        } else if (MODE == 16) {  // LDG.128, 4 lanes x 16 B = 64 B contiguous per row, 8 rows (bf16, 8 ch / lane)
            const unsigned row = ROW4;
            const float4 v = __ldg(reinterpret_cast<const float4 *>(base + row * 128 + ((row & 1) ? 64 : 0)) + (lane & 3));
            acc += v.x + v.y + v.z + v.w;
        }
    }
    const long long t1 = clock64();
    if (acc == 123.456f) out[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static int g_grid = 148;

template <int MODE>
void run(const char *name, const float *g, float *out, float *red, long long *cyc, int threads)
{
    cudaMemset(cyc, 0, 148 * 8);
    k<MODE><<<g_grid, threads>>>(g, out, red, cyc);
    cudaDeviceSynchronize();
    k<MODE><<<g_grid, threads>>>(g, out, red, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < g_grid; ++i) avg += (double)h[i];
    avg /= g_grid;
    const double per = avg / ((double)ITERS * (threads / 32));
    printf("%-62s grid=%3d threads=%4d  %7.2f cyc / warp-instr   (%s)\n", name, g_grid, threads, per, cudaGetErrorString(e));
}

int main(int argc, char **argv)
{
    const bool red_only = argc > 1 && argv[1][0] != 'p';
    float *g, *out, *red;
    long long *cyc;
    const size_t n = (size_t)148 * ROWS * 32;
    cudaMalloc(&g, n * 4);
    cudaMalloc(&red, n * 4);
    cudaMalloc(&out, 16);
    cudaMalloc(&cyc, 148 * 8);
    cudaMemset(g, 0, n * 4);
    cudaMemset(red, 0, n * 4);
    if (red_only) {
        for (int grid : {8, 18, 37, 74, 148}) {
            g_grid = grid;
            run<11>("RED.128  8 lanes/row, 4 rows", g, out, red, cyc, 1024);
            run<13>("RED.32   32 lanes, 1 row", g, out, red, cyc, 1024);
            run<0>("LDG.128  4 full rows (L1 hits)", g, out, red, cyc, 1024);
        }
        return 0;
    }
    if (argc > 1 && argv[1][0] == 'p') {        // ./l1_patterns p : predicated-off lane groups
        for (int threads : {256, 1024}) {
            run<0>("LDG.128  8 lanes/row, 4 live rows", g, out, red, cyc, threads);
            run<21>("LDG.128  8 lanes/row, 3 live rows (group 0 predicated off)", g, out, red, cyc, threads);
            run<22>("LDG.128  8 lanes/row, 2 live rows", g, out, red, cyc, threads);
            run<23>("LDG.128  8 lanes/row, 1 live row", g, out, red, cyc, threads);
            run<24>("LDG.128  8 lanes/row, 0 live rows (all predicated off)", g, out, red, cyc, threads);
            run<25>("LDG.128  8 lanes/row, random half of the groups off", g, out, red, cyc, threads);
        }
        return 0;
    }
    for (int threads : {512, 1024}) {
        run<0>("LDG.128  8 lanes/row, 4 full rows        [value gather now]", g, out, red, cyc, threads);
        run<14>("LDG.128  all lanes same row (1 line)", g, out, red, cyc, threads);
        run<1>("LDG.64   broadcast, 8-lane groups, 4 rows", g, out, red, cyc, threads);
        run<2>("LDG.64   broadcast, 4-lane groups, 8 rows", g, out, red, cyc, threads);
        run<3>("LDG.32   broadcast, 4-lane groups, 8 rows", g, out, red, cyc, threads);
        run<4>("LDG.256  4 lanes/row, 8 full rows", g, out, red, cyc, threads);
        run<5>("LDG.128  4 lanes/half-row, 8 rows", g, out, red, cyc, threads);
        run<15>("LDG.64   8 lanes x 8 B = 64 B/row, 4 rows   [bf16 gather now]", g, out, red, cyc, threads);
        run<16>("LDG.128  4 lanes x 16 B = 64 B/row, 8 rows  [bf16, 8 ch/lane]", g, out, red, cyc, threads);
        run<17>("LDS.128  gather 8 lanes/row, 4 rows         [smem value tile]", g, out, red, cyc, threads);
        run<18>("2xLDS.128 gather 4 lanes/row, 8 rows", g, out, red, cyc, threads);
        run<19>("LDG.128 4 rows + LDS.128 4 rows (mixed)", g, out, red, cyc, threads);
        run<20>("2xLDG.128 4 lanes/row (32 B per lane), 8 rows", g, out, red, cyc, threads);
        run<6>("LDS.128  broadcast, 8-lane groups, 4 distinct", g, out, red, cyc, threads);
        run<8>("LDS.128  broadcast, 4-lane groups, 8 distinct", g, out, red, cyc, threads);
        run<7>("LDS.32   broadcast, 8-lane groups, 4 distinct", g, out, red, cyc, threads);
        run<10>("LDS.128  no broadcast, 512 B/warp", g, out, red, cyc, threads);
        run<9>("SHFL     width 8", g, out, red, cyc, threads);
        run<11>("RED.128  8 lanes/row, 4 rows             [grad_value now]", g, out, red, cyc, threads);
        run<12>("RED.64   16 lanes/row, 2 rows", g, out, red, cyc, threads);
        run<13>("RED.32   32 lanes, 1 row", g, out, red, cyc, threads);
    }
    return 0;
}
