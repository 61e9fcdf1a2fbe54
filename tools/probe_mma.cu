// Micro-probe (tools/, not product): does a running tcgen05.mma slow tcgen05.ld down, and what do short accumulation chains cost?
//   one CTA per SM; warp 1 issues MMAs (128 x N x 16, fp16, operands = whatever sits in shared memory) into accumulator columns
//   [0, N); warps 4.. drain columns [256, 512) with tcgen05.ld x32 (two in flight), the extrapolation GEMM's epilogue pattern.
//   modes: MMA alone, LD alone, both; and "tiles": K_STEPS MMAs per accumulator then commit, round-robin over NACC accumulators,
//   with the epilogue warps waiting for each commit, draining the accumulator (optionally with the FFMA work) and handing it back.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I image-processing-graph-laplacian_b200/csrc -o tools/_bin/probe_mma tools/probe_mma.cu
#include "tc_common.cuh"

void gl_set_error(const char*, ...) {}

using namespace tc;

__device__ __forceinline__ void ffma2_(float& d0, float& d1, float a0, float a1, float b0, float b1) { ffma2(d0, d1, a0, a1, b0, b1); }

// mode bit 0: MMAs run, bit 1: loads run
template <int N>
__global__ void __launch_bounds__(384, 1) k_contend(int mode, int mma_iters, int ld_iters, long long* cyc_mma, long long* cyc_ld)
{
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;                  // 128 x 64 fp16, SW128 K-major
    uint8_t* sb = smem + 16384;          // 256 x 64 fp16
    uint64_t* bars = (uint64_t*)(smem + 16384 + 32768);
    uint32_t* slot = (uint32_t*)(bars + 8);
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003800u ^ (uint32_t)(i * 2654435761u & 0x03ff03ffu);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 1 && lane == 0) {
        mbar_init(smem_u32(bars), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *slot;
    if (warp == 1 && (mode & 1)) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(128, N, 0);
            const uint64_t da = make_smem_desc(smem_u32(sa)), db = make_smem_desc(smem_u32(sb));
            uint32_t ph = 0;
            const long long t0 = clock64();
            for (int it = 0; it < mma_iters; ++it) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(tmem, da + 2 * k, db + 2 * k, idesc, 1u);
                if ((it & 15) == 15) {   // bound the queue: wait for the last 64 MMAs
                    umma_commit(smem_u32(bars));
                    while (!mbar_try_wait(smem_u32(bars), ph)) {}
                    ph ^= 1;
                }
            }
            umma_commit(smem_u32(bars));
            while (!mbar_try_wait(smem_u32(bars), ph)) {}
            cyc_mma[blockIdx.x] = clock64() - t0;
        }
    } else if (warp >= 4 && (mode & 2)) {
        const uint32_t row = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256u + (uint32_t)(((warp - 4) >> 2) * 128);
        uint32_t acc = 0;
        const long long t0 = clock64();
        for (int it = 0; it < ld_iters; ++it) {
            uint32_t v[2][32];
            tmem_ld_32x32b_x32(row + (uint32_t)((it & 1) * 64), v[0]);
            tmem_ld_32x32b_x32(row + (uint32_t)((it & 1) * 64 + 32), v[1]);
            tmem_ld_wait();
            acc ^= v[0][3] ^ v[1][17];
        }
        if (lane == 0 && warp == 4) cyc_ld[blockIdx.x] = clock64() - t0;
        if (acc == 0x1234567u) cyc_ld[0] = 0;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

// "tiles": the pipeline of the fused kernel without operand loads.  NACC accumulators of N columns (NACC * N <= 512); the issuer waits
// for accumulator (t % NACC) to be free, issues KSTEPS MMAs of 128 x N x 16, commits; 8 epilogue warps wait for the commit, drain the
// accumulator (WORK = 0: loads only; 1: + FFMA2 with weights from shared memory; 2: + FFMA with weights in registers), hand it back.
template <int N, int NACC, int WORK>
__global__ void __launch_bounds__(384, 1) k_tiles(int tiles, int ksteps, long long* cyc, float* sink)
{
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;
    uint8_t* sb = smem + 16384;
    uint64_t* bars = (uint64_t*)(smem + 16384 + 32768);   // [0,NACC) full, [NACC, 2 NACC) empty
    uint32_t* slot = (uint32_t*)(bars + 16);
    float* wsm = (float*)(bars + 32);                      // 256 weights
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003800u ^ (uint32_t)(i * 2654435761u & 0x03ff03ffu);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) wsm[i] = 1.0f + i * 1e-3f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 1 && lane == 0) {
        for (int a = 0; a < NACC; ++a) {
            mbar_init(smem_u32(bars + a), 1);
            mbar_init(smem_u32(bars + NACC + a), 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *slot;
    const long long t0 = clock64();
    if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(128, N, 0);
            const uint64_t da = make_smem_desc(smem_u32(sa)), db = make_smem_desc(smem_u32(sb));
            for (int t = 0; t < tiles; ++t) {
                const int a = t % NACC;
                const uint32_t ph = (uint32_t)((t / NACC) & 1);
                while (!mbar_try_wait(smem_u32(bars + NACC + a), ph ^ 1)) {}
                tcgen05_fence_after();
                for (int k = 0; k < ksteps; ++k) umma_f16(tmem + (uint32_t)(a * N), da + 2 * (k & 3), db + 2 * (k & 3), idesc, (uint32_t)(k != 0));
                umma_commit(smem_u32(bars + a));
            }
        }
    } else if (warp >= 4) {
        const int share = (warp - 4) >> 2;    // two warps per lane quarter: each half of the columns
        constexpr int COLS = N / 2;
        float wr[WORK == 2 ? COLS : 1];
        if (WORK == 2) {
#pragma unroll
            for (int i = 0; i < COLS; ++i) wr[i] = wsm[share * COLS + i];
        }
        float dot[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) dot[i] = 0.f;
        const uint32_t wb = smem_u32(wsm + share * COLS);
        for (int t = 0; t < tiles; ++t) {
            const int a = t % NACC;
            const uint32_t ph = (uint32_t)((t / NACC) & 1);
            while (!mbar_try_wait(smem_u32(bars + a), ph)) {}
            tcgen05_fence_after();
            const uint32_t row = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(a * N + share * COLS);
#pragma unroll
            for (int c0 = 0; c0 < COLS; c0 += 64) {
                uint32_t v[2][32];
                tmem_ld_32x32b_x32(row + (uint32_t)c0, v[0]);
                tmem_ld_32x32b_x32(row + (uint32_t)(c0 + 32), v[1]);
                tmem_ld_wait();
                if (WORK == 0) {
                    dot[0] += __uint_as_float(v[0][0]) + __uint_as_float(v[1][31]);
                } else {
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int g8 = 0; g8 < 4; ++g8) {
                            if (WORK == 1) {
                                float w[8];
                                lds_f4(wb + (uint32_t)((c0 + 32 * h + 8 * g8) * 4), &w[0]);
                                lds_f4(wb + (uint32_t)((c0 + 32 * h + 8 * g8 + 4) * 4), &w[4]);
#pragma unroll
                                for (int i = 0; i < 8; i += 2)
                                    ffma2_(dot[i], dot[i + 1], __uint_as_float(v[h][8 * g8 + i]), __uint_as_float(v[h][8 * g8 + i + 1]), w[i], w[i + 1]);
                            } else {
#pragma unroll
                                for (int i = 0; i < 8; ++i) dot[i] = fmaf(__uint_as_float(v[h][8 * g8 + i]), wr[c0 + 32 * h + 8 * g8 + i], dot[i]);
                            }
                        }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(bars + NACC + a));
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += dot[i];
        if (s == 1.2345f) sink[0] = s;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

static long long maxof(long long* d);

// "tiles2": as k_tiles with N = 256 and two accumulators, but TWO issuer threads (warps 1 and 3, issuer w owns accumulator w) and
// G epilogue groups of 8 warps (G = 2: group g owns accumulator g; G = 1: one group alternates).  ATONCE = 1: the epilogue requests all
// of its columns back to back, waits once, hands the accumulator back and only then multiplies.
template <int G, int WORK, int ATONCE, int ISSUERS>
__global__ void __launch_bounds__(128 + 256 * G, 1) k_tiles2(int tiles, int ksteps, long long* cyc, float* sink)
{
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;
    uint8_t* sb = smem + 16384;
    uint64_t* bars = (uint64_t*)(smem + 16384 + 32768);   // [0,2) full, [2,4) empty
    uint32_t* slot = (uint32_t*)(bars + 16);
    float* wsm = (float*)(bars + 32);
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003800u ^ (uint32_t)(i * 2654435761u & 0x03ff03ffu);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) wsm[i] = 1.0f + i * 1e-3f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 1 && lane == 0) {
        for (int a = 0; a < 2; ++a) { mbar_init(smem_u32(bars + a), 1); mbar_init(smem_u32(bars + 2 + a), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *slot;
    const long long t0 = clock64();
    if ((warp == 1 || (warp == 3 && ISSUERS == 2)) && lane == 0) {
        const int me = warp == 1 ? 0 : 1;
        const uint32_t idesc = make_idesc(128, 256, 0);
        const uint64_t da = make_smem_desc(smem_u32(sa)), db = make_smem_desc(smem_u32(sb));
        for (int t = 0; t < tiles; ++t) {
            const int a = t & 1;
            if (ISSUERS == 2 && a != me) continue;
            const uint32_t ph = (uint32_t)((t >> 1) & 1);
            while (!mbar_try_wait(smem_u32(bars + 2 + a), ph ^ 1)) {}
            tcgen05_fence_after();
            for (int k = 0; k < ksteps; ++k) umma_f16(tmem + (uint32_t)(a * 256), da + 2 * (k & 3), db + 2 * (k & 3), idesc, (uint32_t)(k != 0));
            umma_commit(smem_u32(bars + a));
        }
    } else if (warp >= 4) {
        const int ew = warp - 4, grp = ew >> 3, share = (ew >> 2) & 1;
        float dot[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) dot[i] = 0.f;
        const uint32_t wb = smem_u32(wsm + share * 128);
        for (int t = 0; t < tiles; ++t) {
            const int a = t & 1;
            if (G == 2 && a != grp) continue;
            const uint32_t ph = (uint32_t)((t >> 1) & 1);
            while (!mbar_try_wait(smem_u32(bars + a), ph)) {}
            tcgen05_fence_after();
            const uint32_t row = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(a * 256 + share * 128);
            if (ATONCE) {
                uint32_t v[4][32];
#pragma unroll
                for (int k = 0; k < 4; ++k) tmem_ld_32x32b_x32(row + (uint32_t)(32 * k), v[k]);
                tmem_ld_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(bars + 2 + a));
                if (WORK) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int g8 = 0; g8 < 4; ++g8) {
                            float w[8];
                            lds_f4(wb + (uint32_t)((32 * k + 8 * g8) * 4), &w[0]);
                            lds_f4(wb + (uint32_t)((32 * k + 8 * g8 + 4) * 4), &w[4]);
#pragma unroll
                            for (int i = 0; i < 8; i += 2)
                                ffma2_(dot[i], dot[i + 1], __uint_as_float(v[k][8 * g8 + i]), __uint_as_float(v[k][8 * g8 + i + 1]), w[i], w[i + 1]);
                        }
                } else dot[0] += __uint_as_float(v[0][0]) + __uint_as_float(v[3][31]);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(row + (uint32_t)(32 * k), v);
                    tmem_ld_wait();
                    if (k == 3) {
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(bars + 2 + a));
                    }
                    if (WORK) {
#pragma unroll
                        for (int g8 = 0; g8 < 4; ++g8) {
                            float w[8];
                            lds_f4(wb + (uint32_t)((32 * k + 8 * g8) * 4), &w[0]);
                            lds_f4(wb + (uint32_t)((32 * k + 8 * g8 + 4) * 4), &w[4]);
#pragma unroll
                            for (int i = 0; i < 8; i += 2)
                                ffma2_(dot[i], dot[i + 1], __uint_as_float(v[8 * g8 + i]), __uint_as_float(v[8 * g8 + i + 1]), w[i], w[i + 1]);
                        }
                    } else dot[0] += __uint_as_float(v[0]);
                }
            }
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += dot[i];
        if (s == 1.2345f) sink[0] = s;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

template <int G, int WORK, int ATONCE, int ISSUERS>
static void tiles2_case(int ksteps)
{
    long long* c;
    float* sink;
    cudaMalloc(&c, 148 * 8);
    cudaMalloc(&sink, 64);
    const int SM = 16384 + 32768 + 4096;
    cudaFuncSetAttribute(k_tiles2<G, WORK, ATONCE, ISSUERS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);
    const int tiles = 4000;
    k_tiles2<G, WORK, ATONCE, ISSUERS><<<148, 128 + 256 * G, SM>>>(tiles, ksteps, c, sink);
    cudaDeviceSynchronize();
    k_tiles2<G, WORK, ATONCE, ISSUERS><<<148, 128 + 256 * G, SM>>>(tiles, ksteps, c, sink);
    cudaError_t rc = cudaDeviceSynchronize();
    const long long t = maxof(c);
    printf("tiles2 issuers=%d groups=%d at_once=%d work=%d ksteps=%d: %7.1f cyc per 128x256 tile  [%s]\n", ISSUERS, G, ATONCE, WORK, ksteps,
           (double)t / tiles, rc == cudaSuccess ? "ok" : cudaGetErrorString(rc));
    fflush(stdout);
    cudaFree(c);
    cudaFree(sink);
}

static long long maxof(long long* d)
{
    long long h[148], mx = 0;
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    return mx;
}

template <int N>
static void contend()
{
    long long *cm, *cl;
    cudaMalloc(&cm, 148 * 8);
    cudaMalloc(&cl, 148 * 8);
    const int SM = 16384 + 32768 + 4096;
    cudaFuncSetAttribute(k_contend<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);
    const int mi = 4000, li = 20000;
    for (int mode = 1; mode <= 3; ++mode) {
        cudaMemset(cm, 0, 148 * 8);
        cudaMemset(cl, 0, 148 * 8);
        k_contend<N><<<148, 384, SM>>>(mode, mi, li, cm, cl);
        cudaError_t rc = cudaDeviceSynchronize();
        const long long a = maxof(cm), b = maxof(cl);
        printf("contend N=%3d mode=%d (%s): MMA %8lld cyc = %6.1f cyc per 128xNx16 (floor %d) | LD %8lld cyc = %6.1f B/clk/SM  [%s]\n", N, mode,
               mode == 1 ? "mma only" : mode == 2 ? "ld only " : "both    ", a, a ? (double)a / (4.0 * mi) : 0.0, 128 * N / 256, b,
               b ? 8.0 * li * 2 * 32 * 32 * 4 / (double)b : 0.0, rc == cudaSuccess ? "ok" : cudaGetErrorString(rc));
    }
    cudaFree(cm);
    cudaFree(cl);
}

template <int N, int NACC, int WORK>
static void tiles_case(int ksteps)
{
    long long* c;
    float* sink;
    cudaMalloc(&c, 148 * 8);
    cudaMalloc(&sink, 64);
    const int SM = 16384 + 32768 + 4096;
    cudaFuncSetAttribute(k_tiles<N, NACC, WORK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);
    const int tiles = 4000;
    k_tiles<N, NACC, WORK><<<148, 384, SM>>>(tiles, ksteps, c, sink);
    cudaDeviceSynchronize();
    k_tiles<N, NACC, WORK><<<148, 384, SM>>>(tiles, ksteps, c, sink);
    cudaError_t rc = cudaDeviceSynchronize();
    const long long t = maxof(c);
    printf("tiles N=%3d NACC=%d ksteps=%2d work=%d: %7.1f cyc per tile = %6.1f cyc per 128x256-equivalent (MMA floor %4d, per 256 cols)  [%s]\n", N, NACC,
           ksteps, WORK, (double)t / tiles, (double)t / tiles * 256 / N, ksteps * 128, rc == cudaSuccess ? "ok" : cudaGetErrorString(rc));
    cudaFree(c);
    cudaFree(sink);
}

int main(int argc, char** argv)
{
    if (argc > 1) {   // pipeline-structure study
        for (int ks : {1, 2}) {
            tiles2_case<1, 0, 0, 1>(ks); tiles2_case<1, 1, 0, 1>(ks); tiles2_case<1, 1, 1, 1>(ks);
            tiles2_case<1, 1, 0, 2>(ks); tiles2_case<1, 1, 1, 2>(ks);
            tiles2_case<2, 0, 0, 1>(ks); tiles2_case<2, 1, 0, 1>(ks); tiles2_case<2, 1, 0, 2>(ks); tiles2_case<2, 0, 0, 2>(ks);
        }
        return 0;
    }
    contend<256>();
    contend<128>();
    for (int ks : {2, 4, 7, 16}) {
        tiles_case<256, 2, 0>(ks);
        tiles_case<256, 2, 1>(ks);
        tiles_case<256, 2, 2>(ks);
        tiles_case<128, 4, 0>(ks);
        tiles_case<128, 4, 1>(ks);
        tiles_case<128, 4, 2>(ks);
        tiles_case<128, 2, 1>(ks);
    }
    return 0;
}
