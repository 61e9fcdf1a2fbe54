// Non-local-means patch affinity, the fourth plugin of the reference's Python prototype
// (python/affinity_methods/NLM.py:9-34, registered in python/affinity_methods/__init__.py:8-13):
//     K(s, q) = exp(-sum_k (G_k (P_s[k] - P_q[k]))^2 / h^2)
// over the 7x7 patches P around the two pixels of the symmetrically padded image (np.pad 'symmetric', NLM.py:16), G the
// Gaussian patch weights of sigma 1.2 normalised to sum 1 (python/utils.py:19-32, NLM.py:17-19), h = 3 in the reference
// (NLM.py:12; here the h_val argument).  One channel (the reference filters the luminance).
//
// Rows and columns are raster indices.  The reference's columns come out in column-major pixel order (im2col of the transposed
// image, NLM.py:21) although its callers read them as raster indices; tests/test_oracle.py pins the oracle to the reference
// through that index map and the CUDA path to the oracle.
//
// Same outputs and conventions as affinity.cu: K_A fp64 p x p, K_B in the blocked fp16 layout [block][512 pixels][64 slots]
// (no spatial term, so every block is stored), D and T from the fp32 kernel values before rounding, fixed-order reductions.
// The contraction has inner dimension 49: here it runs on the CUDA cores as 49 x (subtract, FMA) per pair with the
// weighted patches staged in shared memory -- a tensor-core formulation (|a-b|^2 = |a|^2 + |b|^2 - 2ab) needs split
// operands to survive the cancellation and is left for later.
#include <cmath>

#include "common.cuh"

#define NLM_K 7
#define NLM_KK 49
#define NLM_RAD 3
#define NLM_THREADS 256
#define NLM_TP 512      // pixels per tile, as in affinity.cu
#define NLM_PASS 8      // pixels per thread and pass (two passes cover the 16 pixels a thread owns)

struct NlmWeights {
    float g[NLM_KK];
};

// index into the symmetrically padded image: -1 -> 0, -2 -> 1, n -> n - 1, n + 1 -> n - 2 (numpy 'symmetric')
__device__ __forceinline__ int nlm_reflect(int i, int n) { return i < 0 ? -i - 1 : (i >= n ? 2 * n - i - 1 : i); }

// weighted patches of the samples in INTERNAL order, SoA [49][p_int]; empty slots carry 1e18 so that K == 0
__global__ void k_nlm_sample_patches(const uint8_t* __restrict__ img, const uint32_t* __restrict__ samples, const uint32_t* __restrict__ perm,
                                     int p_int, int width, int height, NlmWeights w, float* __restrict__ sp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p_int) return;
    const uint32_t j = perm[i];
    if (j == 0xffffffffu) {
        for (int k = 0; k < NLM_KK; ++k) sp[(size_t)k * p_int + i] = 1e18f;
        return;
    }
    const uint32_t q = samples[j];
    const int r = (int)(q / width), c = (int)(q % width);
    for (int k = 0; k < NLM_KK; ++k) {
        const int rr = nlm_reflect(r + k / NLM_K - NLM_RAD, height), cc = nlm_reflect(c + k % NLM_K - NLM_RAD, width);
        sp[(size_t)k * p_int + i] = w.g[k] * (float)img[(size_t)rr * width + cc];
    }
}

// K_A in fp64, and (K_A y_S) beside it for the filter's projection from the affinity sums
__global__ void k_nlm_affinity_A(const uint8_t* __restrict__ img, const uint32_t* __restrict__ samples, int p, int width, int height,
                                 NlmWeights w, double inv_h2, double* __restrict__ KA)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= p) return;
    const uint32_t a = samples[i], b = samples[j];
    const int ra = (int)(a / width), ca = (int)(a % width), rb = (int)(b / width), cb = (int)(b % width);
    double d2 = 0.0;
    for (int k = 0; k < NLM_KK; ++k) {
        const int dr = k / NLM_K - NLM_RAD, dc = k % NLM_K - NLM_RAD;
        const double va = (double)img[(size_t)nlm_reflect(ra + dr, height) * width + nlm_reflect(ca + dc, width)];
        const double vb = (double)img[(size_t)nlm_reflect(rb + dr, height) * width + nlm_reflect(cb + dc, width)];
        const double t = (double)w.g[k] * (va - vb);
        d2 += t * t;
    }
    KA[(size_t)i * p + j] = exp(-d2 * inv_h2);
}

// K_B tile kernel.  Thread (tx = tid % 8, ty = tid / 8): 8 consecutive slots of the current block x pixels ty + 32 i of the
// tile, eight pixels per pass.  Shared memory: the tile's weighted pixel patches [49][512], the block's weighted sample
// patches [49][64], the CTA's running sums [2][p_int] and the per-warp partials of a block.
__global__ void __launch_bounds__(NLM_THREADS, 1)
k_nlm_affinity_B(const uint8_t* __restrict__ img, const float* __restrict__ sp /* [49][p_int] */, int p_int, int width, int height,
                 int64_t q0, int64_t q1, NlmWeights w, float b2 /* -log2(e) / h^2 */, const int4* __restrict__ tab,
                 const int* __restrict__ starts, __half* __restrict__ KB /* [block][512][64] */,
                 float* __restrict__ partial /* [gridDim.x][2][p_int] */)
{
    extern __shared__ float nlm_smem[];
    float* cta_sum = nlm_smem;                         // [2][p_int]: D and T
    float* ws = cta_sum + 2 * p_int;                   // [2][8 warps][64]
    float* px = ws + 2 * 8 * 64;                       // [49][512] weighted pixel patches
    float* sb_ = px + NLM_KK * NLM_TP;                 // [49][64] weighted sample patches of the current block
    float* pv = sb_ + NLM_KK * 64;                     // [512] the pixels' own values (for T)
    const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 2 * p_int; i += NLM_THREADS) cta_sum[i] = 0.f;

    const int64_t tiles = (q1 - q0 + NLM_TP - 1) / NLM_TP;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t base = q0 + tile * NLM_TP;
        __syncthreads();   // the previous tile's readers are done
        for (int e = tid; e < NLM_KK * NLM_TP; e += NLM_THREADS) {
            const int k = e / NLM_TP, i = e % NLM_TP;
            const int64_t q = min(base + i, q1 - 1);
            const int r = (int)(q / width), c = (int)(q % width);
            const int rr = nlm_reflect(r + k / NLM_K - NLM_RAD, height), cc = nlm_reflect(c + k % NLM_K - NLM_RAD, width);
            const float v = (float)img[(size_t)rr * width + cc];
            px[e] = w.g[k] * v;
            if (k == NLM_KK / 2) pv[i] = v;
        }
        const int4 tl = tab[tile];
        for (int ci = 0; ci < tl.y; ++ci) {
            const int sb = starts[tl.x + ci];
            __syncthreads();   // px ready (first block) / the previous block's readers of sb_ and ws are done
            for (int e = tid; e < NLM_KK * 64; e += NLM_THREADS) sb_[e] = sp[(size_t)(e >> 6) * p_int + sb + (e & 63)];
            __syncthreads();
            __half* kb_blk = KB + ((size_t)(tl.z + ci) * NLM_TP) * 64 + (tx << 3);
            float acc[8], tacc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = tacc[j] = 0.f;
            for (int pass = 0; pass < 2; ++pass) {
                float d[NLM_PASS][8];
#pragma unroll
                for (int i = 0; i < NLM_PASS; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) d[i][j] = 0.f;
#pragma unroll 7
                for (int k = 0; k < NLM_KK; ++k) {
                    float b[8];
                    *(float4*)&b[0] = *(const float4*)&sb_[k * 64 + (tx << 3)];
                    *(float4*)&b[4] = *(const float4*)&sb_[k * 64 + (tx << 3) + 4];
#pragma unroll
                    for (int i = 0; i < NLM_PASS; ++i) {
                        const float a = px[k * NLM_TP + ty + 32 * (pass * NLM_PASS + i)];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float t = a - b[j];
                            d[i][j] = fmaf(t, t, d[i][j]);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < NLM_PASS; ++i) {
                    const int pi = ty + 32 * (pass * NLM_PASS + i);
                    if (base + pi < q1) {
                        float kv[8];
                        const float y = pv[pi];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            kv[j] = exp2f(d[i][j] * b2);
                            acc[j] += kv[j];
                            tacc[j] = fmaf(kv[j], y, tacc[j]);
                        }
                        __half2 h0 = __floats2half2_rn(kv[0], kv[1]), h1 = __floats2half2_rn(kv[2], kv[3]);
                        __half2 h2 = __floats2half2_rn(kv[4], kv[5]), h3 = __floats2half2_rn(kv[6], kv[7]);
                        uint4 pk;
                        pk.x = *(uint32_t*)&h0; pk.y = *(uint32_t*)&h1; pk.z = *(uint32_t*)&h2; pk.w = *(uint32_t*)&h3;
                        *(uint4*)&kb_blk[(size_t)pi * 64] = pk;
                    }
                }
            }
            // row sums: lanes sharing tx (fixed shuffle tree), then the 8 warps in order (deterministic)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
#pragma unroll
                for (int o = 8; o < 32; o <<= 1) {
                    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
                    tacc[j] += __shfl_xor_sync(0xffffffffu, tacc[j], o);
                }
            }
            if (lane < 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    ws[warp * 64 + (lane << 3) + j] = acc[j];
                    ws[8 * 64 + warp * 64 + (lane << 3) + j] = tacc[j];
                }
            }
            __syncthreads();
            if (tid < 128) {
                const int which = tid >> 6, sidx = tid & 63;
                float sum = 0.f;
#pragma unroll
                for (int wi = 0; wi < 8; ++wi) sum += ws[which * 8 * 64 + wi * 64 + sidx];
                cta_sum[which * p_int + sb + sidx] += sum;
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < 2 * p_int; i += NLM_THREADS) partial[(size_t)blockIdx.x * 2 * p_int + i] = cta_sum[i];
}

static NlmWeights nlm_weights_host()
{
    // matlab_style_gauss2D((7,7), 1.2) (python/utils.py:19-32), normalised again as NLM.py:19 does
    double h[NLM_KK], mx = 0.0, sum = 0.0;
    for (int a = 0; a < NLM_K; ++a)
        for (int b = 0; b < NLM_K; ++b) {
            const double y = a - NLM_RAD, x = b - NLM_RAD;
            h[a * NLM_K + b] = exp(-(x * x + y * y) / (2.0 * 1.2 * 1.2));
            mx = fmax(mx, h[a * NLM_K + b]);
        }
    for (int k = 0; k < NLM_KK; ++k) {
        if (h[k] < 2.220446049250313e-16 * mx) h[k] = 0.0;
        sum += h[k];
    }
    double sum2 = 0.0;
    for (int k = 0; k < NLM_KK; ++k) { h[k] /= sum; sum2 += h[k]; }
    NlmWeights w;
    for (int k = 0; k < NLM_KK; ++k) w.g[k] = (float)(h[k] / sum2);
    return w;
}

size_t gl_nlm_smem_bytes(int p_int) { return sizeof(float) * ((size_t)2 * p_int + 2 * 8 * 64 + NLM_KK * NLM_TP + NLM_KK * 64 + NLM_TP); }

// K_A and the K_B blocks + partial sums for the NLM kind; the caller (gl_impl_affinity) owns layout, buffers and reductions
int gl_nlm_affinity_launch(gl_ctx* ctx, double h, int p_int, double* KA, const int4* tab, const int* starts, const uint32_t* perm, __half* KB,
                           float* partial, int grid)
{
    const int p = (int)ctx->p;
    GL_REQUIRE(ctx->channels == 1, "affinity: the NLM kind is defined on one channel");
    GL_REQUIRE(ctx->width >= NLM_RAD && ctx->height >= NLM_RAD, "affinity: the NLM patches need an image of at least 3 x 3 pixels");
    GL_REQUIRE(ctx->tile_kbs == 64, "affinity: the NLM kind stores 64-slot blocks (option kb_block)");
    const NlmWeights w = nlm_weights_host();
    gl_buf* sp = nullptr;
    GL_CHECK(gl_alloc(ctx, sizeof(float) * (size_t)NLM_KK * p_int, &sp));
    int rc = GL_OK;
    do {
        k_nlm_sample_patches<<<(unsigned)ceil_div(p_int, 128), 128, 0, ctx->stream>>>(
            (const uint8_t*)ctx->img->ptr, (const uint32_t*)ctx->samples->ptr, perm, p_int, ctx->width, ctx->height, w, (float*)sp->ptr);
        ctx->launches++;
        dim3 ga((unsigned)ceil_div(p, 128), (unsigned)p);
        k_nlm_affinity_A<<<ga, 128, 0, ctx->stream>>>((const uint8_t*)ctx->img->ptr, (const uint32_t*)ctx->samples->ptr, p, ctx->width,
                                                     ctx->height, w, 1.0 / (h * h), KA);
        ctx->launches++;
        const size_t smem = gl_nlm_smem_bytes(p_int);
        GL_CUDA_BREAK(rc, cudaFuncSetAttribute(k_nlm_affinity_B, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        {
            StageTimer kt(ctx, GL_T_K_AFFINITY_B);
            k_nlm_affinity_B<<<grid, NLM_THREADS, smem, ctx->stream>>>((const uint8_t*)ctx->img->ptr, (const float*)sp->ptr, p_int, ctx->width,
                                                                      ctx->height, ctx->q0, ctx->q1, w, (float)(-1.4426950408889634 / (h * h)),
                                                                      tab, starts, KB, partial);
        }
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) { gl_set_error("affinity: the NLM kernels failed to launch"); rc = GL_ERR_CUDA; }
    } while (0);
    gl_buf_release(sp);
    return rc;
}
