// a-2: affinity blocks K_A (p x p, fp64) and K_B (band pixels x p_pad, fp16, pixel-major) with the
// row sums D = K_A.1 + K_B.1 taken from the fp32 kernel values before rounding, and the
// image-weighted row sums T = [K_A K_B] y (one per channel) from the same fp32 values: the filter's
// projection c = Phi^T y = U^T y_S + W^T (K_B y_B) is then a p-sized product instead of a pass over Phi,
// and -- more important -- it is free of the fp16 rounding of W, which the heavy cancellation in Phi^T y
// (|c| ~ 1 against |y| ~ 1e4) would amplify to a ~1e-2 error of z - y (measured).
// Replaces ComputeAffinityMatrices / ComputeDistance / ComputeBilateralFilter,
// hpc/affinity.c:129-262,115-122,59-113 (photometric :8-17, spatial :19-57).
//
// Layout choices (B200-first, not the reference's):
//   * K_B is stored TRANSPOSED (one row per pixel, samples contiguous): it is the K-major "A" operand
//     of the extrapolation GEMM (nystroem_gemm.cu) and each pixel row is one 16-byte-vector store target.
//   * every band pixel gets a row, sample pixels included: their rows of Phi are overwritten with the
//     eigenvectors afterwards, which removes the reference's Permutation pass (hpc/utils.c:134-173), and
//     the row sum over ALL pixels is exactly rowsum(K_A) + rowsum(K_B) (hpc/laplacian.c:18-20).
//   * coordinates and grey values are integers: differences are exact in fp32, the two bandwidth
//     factors are applied to the exact squared distances, one ex2 per pair (the reference takes two exps
//     and multiplies, hpc/affinity.c:99,107,110 -- equal up to rounding).
//   * K_B is stored in BLOCKS of [512 pixels][64 samples] fp16 (a pixel row of a block = 128 contiguous bytes), and only
//     the blocks that can hold a non-zero are stored: the samples are in ascending raster order, hence sorted by image
//     row, so for a 512-pixel tile covering rows [ra, rb] the samples within R rows form a contiguous range, and
//     everything outside it has exp(-dr^2/h_loc^2) < 2^-25, which fp16 storage flushes to zero anyway
//     (R = floor(h_loc sqrt(25 ln 2)) + 1).  The tile table {first block, block count, block offset} is
//     built on the host from the sample indices and cached.  Photometric affinity has no cutoff: every block is stored
//     and the layout degenerates to a dense blocked matrix.  At 4K / p=1000 / h_loc=40 this keeps ~21 % of the blocks:
//     the kernel evaluations, the K_B bytes and the extrapolation GEMM's K loop all shrink by that factor.
#include <algorithm>
#include <cmath>

#include "common.cuh"

#define AFF_THREADS 256
#define AFF_TP 512               // pixels per tile
#define AFF_PPT (AFF_TP / 32)    // pixels per thread per 64-sample chunk

// sample features, SoA: [0] row, [1] col, [2..2+C) values; padded samples carry 1e18 so that K == 0
__global__ void k_sample_features(const uint8_t* __restrict__ img, const uint32_t* __restrict__ samples, int p, int p_pad,
                                  int width, int channels, float* __restrict__ sf)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p_pad) return;
    if (i < p) {
        uint32_t q = samples[i];
        sf[i] = (float)(q / width);
        sf[p_pad + i] = (float)(q % width);
        for (int ch = 0; ch < channels; ++ch) sf[(2 + ch) * p_pad + i] = (float)img[(size_t)q * channels + ch];
    } else {
        for (int k = 0; k < 2 + channels; ++k) sf[k * p_pad + i] = 1e18f;
    }
}

// K_A in fp64 exactly as the reference writes it: exp(-d2/h_loc^2) * exp(-dv2/h_val^2)
template <int KIND, int C>
__global__ void k_affinity_A(const uint8_t* __restrict__ img, const uint32_t* __restrict__ samples, int p, int width,
                             double inv_hl2, double inv_hv2, double* __restrict__ KA)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int i = blockIdx.y;
    if (j >= p) return;
    uint32_t a = samples[i], b = samples[j];
    double k = 1.0;
    if (KIND != GL_PHOTOMETRIC) {
        double dr = (double)(a / width) - (double)(b / width), dc = (double)(a % width) - (double)(b % width);
        k *= exp(-(dr * dr + dc * dc) * inv_hl2);
    }
    if (KIND != GL_SPATIAL) {
        double d2 = 0.0;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
            double dv = (double)img[(size_t)a * C + ch] - (double)img[(size_t)b * C + ch];
            d2 += dv * dv;
        }
        k *= exp(-d2 * inv_hv2);
    }
    KA[(size_t)i * p + j] = k;
}

// K_B tile kernel.  Thread (tx = tid % 8, ty = tid / 8): 8 consecutive samples of the current 64-sample
// chunk x pixels ty, ty+32, ... of the tile.  A warp stores 4 pixel rows x 128 contiguous bytes.
template <int KIND, int C>
__global__ void __launch_bounds__(AFF_THREADS, 2)
k_affinity_B(const uint8_t* __restrict__ img, const float* __restrict__ sf, int p_pad, int width, int64_t q0, int64_t q1,
             float a2, float b2,  // -log2(e)/h_loc^2, -log2(e)/h_val^2
             const int4* __restrict__ tab /* per tile: first block, block count, block offset */,
             __half* __restrict__ KB /* [block][512][64] */, float* __restrict__ partial /* [gridDim.x][1 + C][p_pad] */)
{
    extern __shared__ float aff_smem[];
    constexpr int NS = 1 + C;                  // sums per sample: D and T[ch]
    float* cta_sum = aff_smem;                 // [NS][p_pad]
    float* ws = cta_sum + NS * p_pad;          // [2][NS][8 warps][64]
    float* px = ws + 2 * NS * 8 * 64;          // [(2 + C)][AFF_TP] pixel features
    const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < NS * p_pad; i += AFF_THREADS) cta_sum[i] = 0.f;

    const int64_t n_band = q1 - q0;
    const int64_t tiles = (n_band + AFF_TP - 1) / AFF_TP;
    int flip = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t base = q0 + tile * AFF_TP;
        __syncthreads();  // previous tile's readers of px are done
        for (int i = tid; i < AFF_TP; i += AFF_THREADS) {
            int64_t q = base + i;
            bool in = q < q1;
            int64_t qq = in ? q : q1 - 1;
            px[i] = (float)(qq / width);
            px[AFF_TP + i] = (float)(qq % width);
#pragma unroll
            for (int ch = 0; ch < C; ++ch) px[(2 + ch) * AFF_TP + i] = (float)img[(size_t)qq * C + ch];
        }
        __syncthreads();
        const int4 tl = tab[tile];
        for (int ci = 0; ci < tl.y; ++ci) {
            const int ck = tl.x + ci;
            const int s0 = (ck << 6) + (tx << 3);
            __half* kb_blk = KB + ((size_t)(tl.z + ci) * AFF_TP) * 64 + (tx << 3);
            float sr[8], sc[8], sv[C][8];
            if (KIND != GL_PHOTOMETRIC) {
                *(float4*)&sr[0] = *(const float4*)&sf[s0];
                *(float4*)&sr[4] = *(const float4*)&sf[s0 + 4];
                *(float4*)&sc[0] = *(const float4*)&sf[p_pad + s0];
                *(float4*)&sc[4] = *(const float4*)&sf[p_pad + s0 + 4];
            }
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
                *(float4*)&sv[ch][0] = *(const float4*)&sf[(2 + ch) * p_pad + s0];
                *(float4*)&sv[ch][4] = *(const float4*)&sf[(2 + ch) * p_pad + s0 + 4];
            }
            float acc[8], tacc[C][8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                acc[k] = 0.f;
#pragma unroll
                for (int ch = 0; ch < C; ++ch) tacc[ch][k] = 0.f;
            }
#pragma unroll 2
            for (int i = 0; i < AFF_PPT; ++i) {
                const int pi = ty + (i << 5);
                const int64_t q = base + pi;
                const float pr = px[pi], pc = px[AFF_TP + pi];
                float pv[C];
#pragma unroll
                for (int ch = 0; ch < C; ++ch) pv[ch] = px[(2 + ch) * AFF_TP + pi];
                float kv[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float x = 0.f;
                    if (KIND != GL_SPATIAL) {
                        float d = pv[0] - sv[0][k];
                        float t = d * d;
#pragma unroll
                        for (int ch = 1; ch < C; ++ch) {
                            d = pv[ch] - sv[ch][k];
                            t = fmaf(d, d, t);
                        }
                        x = t * b2;
                    }
                    if (KIND != GL_PHOTOMETRIC) {
                        float dr = pr - sr[k], dc = pc - sc[k];
                        float t = fmaf(dc, dc, dr * dr);
                        x = fmaf(t, a2, x);
                    }
                    kv[k] = fast_exp2(x);
                }
                if (q < q1) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        acc[k] += kv[k];
#pragma unroll
                        for (int ch = 0; ch < C; ++ch) tacc[ch][k] = fmaf(kv[k], pv[ch], tacc[ch][k]);
                    }
                    __half2 h0 = __floats2half2_rn(kv[0], kv[1]), h1 = __floats2half2_rn(kv[2], kv[3]);
                    __half2 h2 = __floats2half2_rn(kv[4], kv[5]), h3 = __floats2half2_rn(kv[6], kv[7]);
                    uint4 pk;
                    pk.x = *(uint32_t*)&h0; pk.y = *(uint32_t*)&h1; pk.z = *(uint32_t*)&h2; pk.w = *(uint32_t*)&h3;
                    *(uint4*)&kb_blk[(size_t)pi * 64] = pk;
                }
            }
            // row sums: fixed-order reduction (deterministic): lanes sharing tx, then the 8 warps
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 8);
                acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 16);
#pragma unroll
                for (int ch = 0; ch < C; ++ch) {
                    tacc[ch][k] += __shfl_xor_sync(0xffffffffu, tacc[ch][k], 8);
                    tacc[ch][k] += __shfl_xor_sync(0xffffffffu, tacc[ch][k], 16);
                }
            }
            float* w = ws + flip * NS * 512;
            if (lane < 8) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    w[warp * 64 + (lane << 3) + k] = acc[k];
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) w[(1 + ch) * 512 + warp * 64 + (lane << 3) + k] = tacc[ch][k];
                }
            }
            __syncthreads();
            for (int i = tid; i < NS * 64; i += AFF_THREADS) {
                const int which = i >> 6, sidx = i & 63;
                float sum = 0.f;
#pragma unroll
                for (int wi = 0; wi < 8; ++wi) sum += w[which * 512 + wi * 64 + sidx];
                cta_sum[which * p_pad + (ck << 6) + sidx] += sum;
            }
            flip ^= 1;
        }
    }
    __syncthreads();
    for (int i = tid; i < NS * p_pad; i += AFF_THREADS) partial[(size_t)blockIdx.x * NS * p_pad + i] = cta_sum[i];
}

// DT[which][s] = sum over CTAs (fixed order, fp64); which 0 = D, 1.. = T[ch]; padding samples get 0
__global__ void k_reduce_partials(const float* __restrict__ partial, int nblocks, int p_pad, int ns, double* __restrict__ DT)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ns * p_pad) return;
    double acc = 0.0;
    for (int b = 0; b < nblocks; ++b) acc += (double)partial[(size_t)b * ns * p_pad + i];
    DT[i] = acc;
}

// Tile table of K_B (see the header): one int4 {first 64-sample block, block count, block offset, 0} per 512-pixel tile
// of this rank's band.  Built on the host (tiles x 2 binary searches over the sorted sample rows), uploaded only when
// it differs from the cached one.
static int build_tile_table(gl_ctx* ctx, int kind, double h_loc)
{
    const int p = (int)ctx->p, nblk = ctx->p_pad >> 6, W = ctx->width;
    const int64_t n_band = ctx->q1 - ctx->q0;
    const int64_t tiles = (n_band + AFF_TP - 1) / AFF_TP;
    if (!ctx->h_samples_valid) {
        ctx->h_samples.resize(p);
        GL_CUDA_CHECK(cudaMemcpyAsync(ctx->h_samples.data(), ctx->samples->ptr, sizeof(uint32_t) * p, cudaMemcpyDeviceToHost, ctx->stream));
        GL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        ctx->h_samples_valid = true;
    }
    const bool cut = ctx->kb_cutoff && kind != GL_PHOTOMETRIC;
    // |dr| > h_loc sqrt(25 ln 2)  =>  exp(-dr^2/h_loc^2) < 2^-25  =>  the fp16 value is 0; one more row for rounding slack
    const double rr = std::floor(h_loc * std::sqrt(25.0 * 0.6931471805599453)) + 1.0;
    const int64_t R = rr < 1e9 ? (int64_t)rr : (int64_t)1e9;
    const int64_t key[6] = {W, ctx->q0, ctx->q1, p, R, cut ? 1 : 0};
    if (ctx->tile_tab && !memcmp(key, ctx->tab_key, sizeof(key)) && ctx->tab_samples.size() == (size_t)p &&
        !memcmp(ctx->tab_samples.data(), ctx->h_samples.data(), sizeof(uint32_t) * p))
        return GL_OK;  // same geometry, samples and cutoff as last time: the cached table stands
    std::vector<int> srow(p);
    for (int i = 0; i < p; ++i) srow[i] = (int)(ctx->h_samples[i] / (uint32_t)W);
    std::vector<int4> tab((size_t)tiles);
    int64_t off = 0;
    for (int64_t t = 0; t < tiles; ++t) {
        int lo = 0, cnt = nblk;
        if (cut) {
            const int64_t qa = ctx->q0 + t * AFF_TP, qb = std::min(ctx->q1, qa + AFF_TP) - 1;
            const int64_t ra = qa / W - R, rb = qb / W + R;
            const int s_lo = (int)(std::lower_bound(srow.begin(), srow.end(), (int)std::max<int64_t>(ra, -1)) - srow.begin());
            const int s_hi = (int)(std::upper_bound(srow.begin(), srow.end(), (int)std::min<int64_t>(rb, 0x7fffffff)) - srow.begin());
            lo = s_lo >> 6;
            cnt = ((s_hi + 63) >> 6) - lo;
            if (cnt < 1) { lo = std::min(lo, nblk - 1); cnt = 1; }   // keep one block so that the GEMM writes zeros
        }
        tab[(size_t)t] = make_int4(lo, cnt, (int)off, 0);
        off += cnt;
        GL_REQUIRE(off < 0x7fffffff / 512, "affinity: K_B has too many blocks for 32-bit tile coordinates");
    }
    const bool same = ctx->tile_tab && ctx->h_tile_tab.size() == tab.size() &&
                      !memcmp(ctx->h_tile_tab.data(), tab.data(), sizeof(int4) * tab.size());
    if (!same) {
        const size_t bytes = sizeof(int4) * tab.size();
        if (ctx->tile_tab) gl_buf_release(ctx->tile_tab);
        ctx->tile_tab = nullptr;
        GL_CHECK(gl_alloc(ctx, bytes, &ctx->tile_tab));
        GL_CHECK(gl_ensure_pinned(ctx, bytes));
        GL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // the pinned block may still feed an earlier copy
        memcpy(ctx->pinned, tab.data(), bytes);
        GL_CUDA_CHECK(cudaMemcpyAsync(ctx->tile_tab->ptr, ctx->pinned, bytes, cudaMemcpyHostToDevice, ctx->stream));
        ctx->h_tile_tab.swap(tab);
        ctx->tile_total_blocks = off;
    }
    memcpy(ctx->tab_key, key, sizeof(key));
    ctx->tab_samples = ctx->h_samples;
    return GL_OK;
}

template <int KIND, int C>
static int launch_affinity(gl_ctx* ctx, double h_loc, double h_val, const float* sf, double* KA, const int4* tab, __half* KB,
                           float* partial, int grid)
{
    const int p = (int)ctx->p, p_pad = ctx->p_pad;
    dim3 ga((unsigned)ceil_div(p, 128), (unsigned)p);
    k_affinity_A<KIND, C><<<ga, 128, 0, ctx->stream>>>((const uint8_t*)ctx->img->ptr, (const uint32_t*)ctx->samples->ptr, p,
                                                      ctx->width, 1.0 / (h_loc * h_loc), 1.0 / (h_val * h_val), KA);
    GL_LAUNCH_CHECK(ctx);
    const size_t smem = sizeof(float) * ((size_t)(1 + C) * p_pad + 2 * (1 + C) * 8 * 64 + (size_t)(2 + C) * AFF_TP);
    GL_CUDA_CHECK(cudaFuncSetAttribute(k_affinity_B<KIND, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const float log2e = 1.4426950408889634f;
    StageTimer kt(ctx, GL_T_K_AFFINITY_B);
    k_affinity_B<KIND, C><<<grid, AFF_THREADS, smem, ctx->stream>>>(
        (const uint8_t*)ctx->img->ptr, sf, p_pad, ctx->width, ctx->q0, ctx->q1, (float)(-log2e / (h_loc * h_loc)),
        (float)(-log2e / (h_val * h_val)), tab, KB, partial);
    GL_LAUNCH_CHECK(ctx);
    return GL_OK;
}

int gl_impl_affinity(gl_ctx* ctx, int kind, double h_loc, double h_val, gl_mat** K_A_out, gl_mat** K_B_out)
{
    const int p = (int)ctx->p, p_pad = ctx->p_pad, C = ctx->channels;
    const int64_t n_band = ctx->q1 - ctx->q0;
    GL_REQUIRE(p_pad <= 16384, "affinity: p = %d too large", p);

    gl_mat* KA = gl_mat_new(ctx, GL_MAT_KA);
    gl_mat* KB = gl_mat_new(ctx, GL_MAT_KB);
    gl_buf *sf = nullptr, *partial = nullptr;
    const int grid = ctx->sm_count * 2;
    int rc = GL_OK;
    do {
        KA->rows = KA->cols = KA->local_rows = p;
        KA->ld = p;
        KA->elem_bytes = 8;
        if ((rc = gl_alloc(ctx, sizeof(double) * (size_t)p * p, &KA->buf)) != GL_OK) break;
        KB->rows = p;                   // logical K_B: p x (n - p); stored transposed for the whole band
        KB->cols = ctx->n - p;
        KB->local_rows = n_band;
        KB->ld = 64;                    // blocked storage: a pixel row of a block is 64 samples
        KB->elem_bytes = 2;
        KB->p = p;
        KB->p_pad = p_pad;
        KB->q0 = ctx->q0;
        if ((rc = build_tile_table(ctx, kind, h_loc)) != GL_OK) break;
        KB->tiles = ctx->tile_tab;      // shared with the context's cache (a new table is a new buffer)
        KB->tiles->refs++;
        KB->total_blocks = ctx->tile_total_blocks;
        if ((rc = gl_alloc(ctx, sizeof(__half) * (size_t)KB->total_blocks * AFF_TP * 64, &KB->buf)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(double) * (size_t)(1 + C) * p_pad, &KB->aux)) != GL_OK) break;  // [D | T[ch]]
        if ((rc = gl_alloc(ctx, sizeof(float) * (size_t)(2 + C) * p_pad, &sf)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(float) * (size_t)grid * (1 + C) * p_pad, &partial)) != GL_OK) break;

        k_sample_features<<<(unsigned)ceil_div(p_pad, 256), 256, 0, ctx->stream>>>(
            (const uint8_t*)ctx->img->ptr, (const uint32_t*)ctx->samples->ptr, p, p_pad, ctx->width, C, (float*)sf->ptr);
        ctx->launches++;

#define AFF_CASE(K, CC)                                                                                              \
    if (kind == K && C == CC)                                                                                        \
        rc = launch_affinity<K, CC>(ctx, h_loc, h_val, (const float*)sf->ptr, (double*)KA->buf->ptr, (const int4*)KB->tiles->ptr, \
                                    (__half*)KB->buf->ptr, (float*)partial->ptr, grid);
        AFF_CASE(GL_BILATERAL, 1) else AFF_CASE(GL_BILATERAL, 3) else AFF_CASE(GL_PHOTOMETRIC, 1)
        else AFF_CASE(GL_PHOTOMETRIC, 3) else AFF_CASE(GL_SPATIAL, 1) else AFF_CASE(GL_SPATIAL, 3)
#undef AFF_CASE
        if (rc != GL_OK) break;

        k_reduce_partials<<<(unsigned)ceil_div((1 + C) * p_pad, 128), 128, 0, ctx->stream>>>((const float*)partial->ptr, grid, p_pad,
                                                                                             1 + C, (double*)KB->aux->ptr);
        GL_LAUNCH_CHECK(ctx);
        // ONE allreduce of the band-partial sums: D (p doubles) and T (C x p doubles)
        if ((rc = gl_allreduce_f64(ctx, (double*)KB->aux->ptr, (size_t)(1 + C) * p_pad)) != GL_OK) break;
        KB->channels = C;
        KB->image_epoch = ctx->image_epoch;
        KB->aff_kind = kind;
        KB->aff_h_loc = h_loc;
        KB->aff_h_val = h_val;
    } while (0);
    if (sf) gl_buf_release(sf);
    if (partial) gl_buf_release(partial);
    if (rc != GL_OK) {
        gl_mat_destroy(KA);
        gl_mat_destroy(KB);
        return rc;
    }
    *K_A_out = KA;
    *K_B_out = KB;
    return GL_OK;
}
