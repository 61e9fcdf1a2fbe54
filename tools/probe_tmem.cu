// Micro-probe for the extrapolation GEMM's epilogue (tools/, not product): what one B200 SM sustains for
//   A  tcgen05.ld 32x32b (.x32 / .x64 / .x128), 4 / 8 / 16 warps       -> TMEM read bytes per clock per SM
//   B  FFMA2 / FFMA register chains                                      -> fp32 FMAs per clock per SM
//   C  the epilogue's inner loop without its barriers: ld x32 x2 + 32 FFMA2 per 64 columns, weights from shared memory
//      (LDS.128 broadcast) or from registers
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/probe_tmem tools/probe_tmem.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD32(taddr, r)                                                                                                    \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),        \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),  \
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),             \
                   "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),             \
                   "=r"(r[30]), "=r"(r[31])                                                                               \
                 : "r"(taddr)                                                                                             \
                 : "memory")
#define LD16(taddr, r)                                                                                                    \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                                \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"                         \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),        \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])   \
                 : "r"(taddr)                                                                                             \
                 : "memory")
#define LDWAIT() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1)
{
    asm volatile("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%0, %1};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
        : "+f"(d0), "+f"(d1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void lds_f4(uint32_t addr, float* dst)
{
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(dst[0]), "=f"(dst[1]), "=f"(dst[2]), "=f"(dst[3]) : "r"(addr));
}

struct Tmem {
    uint32_t base;
};
__device__ __forceinline__ uint32_t tmem_alloc_all(uint32_t* slot)
{
    if ((threadIdx.x >> 5) == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return *slot;
}
__device__ __forceinline__ void tmem_free_all(uint32_t base)
{
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512) : "memory");
}

// A: MODE 0: one x32 load then wait; 1: two x32 loads in flight then wait; 2: four x32 in flight; 3: x16 pairs
template <int MODE>
__global__ void k_tmem_ld(int iters, long long* cyc, uint32_t* sink)
{
    __shared__ uint32_t slot;
    const uint32_t base = tmem_alloc_all(&slot);
    const int warp = threadIdx.x >> 5;
    const uint32_t row = base + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t c0 = (uint32_t)((it * 128 + (warp >> 2) * 64) & 511);
        if (MODE == 0) {
            uint32_t v[32];
            LD32(row + (c0 & 480), v);
            LDWAIT();
            acc ^= v[0] ^ v[31];
        } else if (MODE == 1) {
            uint32_t v[2][32];
            LD32(row + (c0 & 448), v[0]);
            LD32(row + (c0 & 448) + 32, v[1]);
            LDWAIT();
            acc ^= v[0][0] ^ v[1][31];
        } else if (MODE == 2) {
            uint32_t v[4][32];
            LD32(row + (c0 & 384), v[0]);
            LD32(row + (c0 & 384) + 32, v[1]);
            LD32(row + (c0 & 384) + 64, v[2]);
            LD32(row + (c0 & 384) + 96, v[3]);
            LDWAIT();
            acc ^= v[0][0] ^ v[1][31] ^ v[2][5] ^ v[3][7];
        } else {
            uint32_t v[2][16];
            LD16(row + (c0 & 480), v[0]);
            LD16(row + (c0 & 480) + 16, v[1]);
            LDWAIT();
            acc ^= v[0][0] ^ v[1][15];
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    tmem_free_all(base);
}

// B: MODE 0 = FFMA2 16 chains (8 pairs), 1 = FFMA 16 chains, 2 = FFMA2 + one LDS.128 per 2 FFMA2
template <int MODE>
__global__ void k_fma(int iters, long long* cyc, float* sink, float seed)
{
    __shared__ float wsm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) wsm[i] = seed + i * 1e-6f;
    __syncthreads();
    float d[16], a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { d[i] = seed * i; a[i] = seed + threadIdx.x * 1e-3f + i; }
    const uint32_t wb = smem_u32(wsm);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 16; i += 2) ffma2(d[i], d[i + 1], a[i], a[i + 1], a[(i + 2) & 15], a[(i + 3) & 15]);
        } else if (MODE == 1) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(d[i]) : "f"(a[i]), "f"(a[(i + 1) & 15]));
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    float w[4];
                    lds_f4(wb + (uint32_t)(((it * 16 + u * 4 + i) * 4) & 4095), w);
                    ffma2(d[i], d[i + 1], a[i], a[i + 1], w[0], w[1]);
                    ffma2(d[i + 2], d[i + 3], a[i + 2], a[i + 3], w[2], w[3]);
                }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += d[i];
    if (s == 1.2345f) sink[0] = s;
}

// C: epilogue inner loop: per iteration 64 columns of this warp's 32 rows: two x32 loads, wait, 32 FFMA2.
//    WREG = 0: weights by LDS.128 broadcast (16 per 64 columns); 1: weights in registers (same 64 every iteration)
template <int WREG>
__global__ void k_epi(int iters, long long* cyc, float* sink, float seed)
{
    __shared__ uint32_t slot;
    __shared__ __align__(16) float wsm[16][128];
    for (int i = threadIdx.x; i < 16 * 128; i += blockDim.x) (&wsm[0][0])[i] = seed + i * 1e-6f;
    const uint32_t base = tmem_alloc_all(&slot);
    const int warp = threadIdx.x >> 5;
    const uint32_t row = base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t wb = smem_u32(&wsm[warp & 15][0]);
    float wr[64];
    if (WREG) {
#pragma unroll
        for (int i = 0; i < 64; ++i) wr[i] = wsm[warp & 15][i];
    }
    float dot[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) dot[i] = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t c0 = (uint32_t)((it * 128 + (warp >> 2) * 64) & 448);
        uint32_t v[2][32];
        LD32(row + c0, v[0]);
        LD32(row + c0 + 32, v[1]);
        LDWAIT();
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
                float w[8];
                if (WREG) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) w[i] = wr[32 * h + 8 * g8 + i];
                } else {
                    lds_f4(wb + (uint32_t)((32 * h + 8 * g8) * 4), &w[0]);
                    lds_f4(wb + (uint32_t)((32 * h + 8 * g8 + 4) * 4), &w[4]);
                }
#pragma unroll
                for (int i = 0; i < 8; i += 2)
                    ffma2(dot[i], dot[i + 1], __uint_as_float(v[h][8 * g8 + i]), __uint_as_float(v[h][8 * g8 + i + 1]), w[i], w[i + 1]);
            }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += dot[i];
    if (s == 1.2345f) sink[0] = s;
    tmem_free_all(base);
}

template <typename F>
static void run(const char* name, F launch, int warps, int iters, double units_per_warp_iter, const char* unit)
{
    long long* cyc;
    cudaMalloc(&cyc, 148 * sizeof(long long));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch(cyc);   // warm-up
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch(cyc);
    cudaEventRecord(e1);
    cudaError_t rc = cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const double total = units_per_warp_iter * warps * iters;
    printf("%-44s warps=%2d  %9lld cyc  %8.1f %s/clk/SM  (%.3f ms, %s)\n", name, warps, mx, total / (double)mx, unit, ms,
           rc == cudaSuccess ? "ok" : cudaGetErrorString(rc));
    cudaFree(cyc);
}

int main()
{
    uint32_t* sink;
    cudaMalloc(&sink, 64);
    const int iters = 20000;
    for (int warps : {4, 8, 16}) {
        run("A0 tcgen05.ld x32, wait each", [&](long long* c) { k_tmem_ld<0><<<148, warps * 32>>>(iters, c, sink); }, warps, iters, 32 * 32 * 4.0, "B");
        run("A1 tcgen05.ld x32 x2 in flight", [&](long long* c) { k_tmem_ld<1><<<148, warps * 32>>>(iters, c, sink); }, warps, iters, 2 * 32 * 32 * 4.0, "B");
        run("A2 tcgen05.ld x32 x4 in flight", [&](long long* c) { k_tmem_ld<2><<<148, warps * 32>>>(iters, c, sink); }, warps, iters, 4 * 32 * 32 * 4.0, "B");
        run("A3 tcgen05.ld x16 x2 in flight", [&](long long* c) { k_tmem_ld<3><<<148, warps * 32>>>(iters, c, sink); }, warps, iters, 2 * 16 * 32 * 4.0, "B");
    }
    for (int warps : {4, 8, 16}) {
        run("B0 FFMA2 register chains", [&](long long* c) { k_fma<0><<<148, warps * 32>>>(iters, c, (float*)sink, 1.0f); }, warps, iters, 4 * 8 * 2 * 32.0, "FMA");
        run("B1 FFMA register chains", [&](long long* c) { k_fma<1><<<148, warps * 32>>>(iters, c, (float*)sink, 1.0f); }, warps, iters, 4 * 16 * 32.0, "FMA");
        run("B2 FFMA2 + LDS.128 per 2", [&](long long* c) { k_fma<2><<<148, warps * 32>>>(iters, c, (float*)sink, 1.0f); }, warps, iters, 4 * 8 * 2 * 32.0, "FMA");
    }
    for (int warps : {4, 8, 16}) {
        run("C0 epilogue loop, weights LDS (cols x rows)", [&](long long* c) { k_epi<0><<<148, warps * 32>>>(iters, c, (float*)sink, 1.0f); }, warps, iters, 64 * 32.0, "elem");
        run("C1 epilogue loop, weights in registers", [&](long long* c) { k_epi<1><<<148, warps * 32>>>(iters, c, (float*)sink, 1.0f); }, warps, iters, 64 * 32.0, "elem");
    }
    return 0;
}
