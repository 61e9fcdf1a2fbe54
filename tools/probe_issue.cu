// Micro-probe (tools/, not product): what do tcgen05.mma issue and tcgen05.commit cost the ISSUING thread?
// One CTA per SM; thread 32 issues, per tile, KS MMAs (128 x 256 x 16) and NC commits to distinct mbarriers, then waits for the last
// commit (so the tensor pipe is idle when the next tile starts, like a short-K pipeline); clock64 around each part.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I image-processing-graph-laplacian_b200/csrc -o tools/_bin/probe_issue tools/probe_issue.cu
#include "tc_common.cuh"
void gl_set_error(const char*, ...) {}
using namespace tc;

template <int KS, int NC, int WAIT_EACH>
__global__ void __launch_bounds__(128, 1) k_issue(int tiles, long long* out)
{
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + 49152);
    uint32_t* slot = (uint32_t*)(bars + 8);
    for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003800u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(smem_u32(bars + i), 1);
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
    if (warp == 1 && lane == 0) {
        const uint32_t idesc = make_idesc(128, 256, 0);
        const uint64_t da = make_smem_desc(smem_u32(smem)), db = make_smem_desc(smem_u32(smem + 16384));
        long long t_mma = 0, t_commit = 0, t_wait = 0;
        uint32_t ph[4] = {0, 0, 0, 0};
        for (int t = 0; t < tiles; ++t) {
            const long long a = clock64();
#pragma unroll
            for (int k = 0; k < KS; ++k) umma_f16(tmem + (uint32_t)((t & 1) * 256), da + 2 * k, db + 2 * k, idesc, (uint32_t)(k != 0));
            const long long b = clock64();
#pragma unroll
            for (int c = 0; c < NC; ++c) umma_commit(smem_u32(bars + (WAIT_EACH ? c : ((t & 1) * 2 + (c & 1)))));
            const long long c0 = clock64();
            if (!WAIT_EACH && t > 0) {   // one tile of slack: wait for the PREVIOUS tile's commits (its own barriers), then go on
                const int o = ((t - 1) & 1) * 2;
                for (int c = 0; c < (NC < 2 ? NC : 2); ++c) { while (!mbar_try_wait(smem_u32(bars + o + c), ph[o + c])) {} ph[o + c] ^= 1; }
            }
            if (WAIT_EACH) {
#pragma unroll
                for (int c = 0; c < NC; ++c) { while (!mbar_try_wait(smem_u32(bars + c), ph[c])) {} ph[c] ^= 1; }
            }
            const long long d = clock64();
            t_mma += b - a; t_commit += c0 - b; t_wait += d - c0;
        }
        if (!WAIT_EACH) {   // drain the last tile
            const int o = ((tiles - 1) & 1) * 2;
            for (int c = 0; c < (NC < 2 ? NC : 2); ++c) { while (!mbar_try_wait(smem_u32(bars + o + c), ph[o + c])) {} ph[o + c] ^= 1; }
        }
        if (blockIdx.x == 0) { out[0] = t_mma; out[1] = t_commit; out[2] = t_wait; }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

template <int KS, int NC, int WE>
static void run()
{
    long long* d;
    cudaMalloc(&d, 64);
    const int SM = 49152 + 2048;
    cudaFuncSetAttribute(k_issue<KS, NC, WE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);
    const int tiles = 4000;
    k_issue<KS, NC, WE><<<148, 128, SM>>>(tiles, d);
    k_issue<KS, NC, WE><<<148, 128, SM>>>(tiles, d);
    cudaError_t rc = cudaDeviceSynchronize();
    long long h[3];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("KS=%d commits=%d wait_each=%d: per tile  mma issue %6.1f  commits %6.1f  wait %6.1f  cycles  [%s]\n", KS, NC, WE, (double)h[0] / tiles,
           (double)h[1] / tiles, (double)h[2] / tiles, rc == cudaSuccess ? "ok" : cudaGetErrorString(rc));
    fflush(stdout);
    cudaFree(d);
}

int main()
{
    run<1, 1, 1>(); run<2, 1, 1>(); run<2, 2, 1>(); run<2, 3, 1>(); run<4, 1, 1>();
    run<1, 1, 0>(); run<2, 1, 0>(); run<2, 2, 0>(); run<4, 1, 0>();
    fflush(stdout);
    return 0;
}
