// f-1: the reference's "-no_approx" mode, matrix-free.
// Replaces ComputeEntireAffinityMatrix (hpc/affinity.c:264-336), ComputeEntireLaplacianMatrix (hpc/laplacian.c:44-65)
// and ComputeResultFromEntireLaplacian (hpc/display.c:128-149):
//     K (n x n),  D = K.1,  alpha = 1 / mean(D),  L = alpha (diag D - K),  z = clip(y - L y, 0, 255).
// The reference materialises K and L as dense n x n fp64 matrices ("very memory consuming", hpc/README.md:19: 550 TB
// for a 4K image).  Here nothing n x n is ever stored:
//     (L y)_i = alpha * sum_j K_ij (y_i - y_j),
// so ONE pass over the pixel pairs accumulates D_i = sum_j K_ij and Q_i = sum_j K_ij (y_i - y_j) per pixel (Q is formed
// from the differences, which removes the cancellation of D_i y_i - (K y)_i), alpha follows from a reduction of D, and
// z = y - alpha Q is elementwise.  The three stage handles the host sees (K, Lapl, result) are views of {D, Q, alpha}.
// With a spatial term, pairs further apart than r_c = h_loc sqrt(9 ln 10) contribute < 1e-9 of a row sum and are
// skipped, so the pass is O(n r_c^2) (photometric affinity has no cutoff: O(n^2), small images only).
// Kernel: CTA = 8 x 32 output pixels (one per thread); source pixels are staged through shared memory in 8 x 128
// chunks and broadcast to all threads; one ex2 per pair; per-chunk partial sums are folded into the totals.
#include <cmath>

#include "common.cuh"

#define FF_TY 8
#define FF_TX 32
#define FF_SR 8
#define FF_SC 128

template <int KIND, int C>
__global__ void __launch_bounds__(FF_TY* FF_TX) k_full_pass(const uint8_t* __restrict__ img, int W, int H, int row0, int row1, int R,
                                                           float a2, float b2,  // -log2(e)/h_loc^2, -log2(e)/h_val^2
                                                           float* __restrict__ Dout, float* __restrict__ Qout /* [band px][C] */)
{
    __shared__ __align__(16) float sv[C][FF_SR * FF_SC];
    const int tx = threadIdx.x % FF_TX, ty = threadIdx.x / FF_TX;
    const int tiles_x = (W + FF_TX - 1) / FF_TX;
    const int tile_r0 = row0 + (blockIdx.x / tiles_x) * FF_TY, tile_c0 = (blockIdx.x % tiles_x) * FF_TX;
    const int pr = tile_r0 + ty, pc = tile_c0 + tx;
    const bool live = pr < row1 && pc < W;
    float pv[C];
#pragma unroll
    for (int ch = 0; ch < C; ++ch) pv[ch] = live ? (float)img[((size_t)pr * W + pc) * C + ch] : 0.f;
    float Dt = 0.f, Qt[C];
#pragma unroll
    for (int ch = 0; ch < C; ++ch) Qt[ch] = 0.f;

    // source window of the whole tile, clipped to the image
    const int wr0 = max(0, tile_r0 - R), wr1 = min(H, min(row1, tile_r0 + FF_TY) + R);
    const int wc0 = max(0, tile_c0 - R), wc1 = min(W, tile_c0 + FF_TX + R);
    const float R2 = (float)R * (float)R;
    for (int sr0 = wr0; sr0 < wr1; sr0 += FF_SR) {
        for (int sc0 = wc0; sc0 < wc1; sc0 += FF_SC) {
            __syncthreads();
            for (int i = threadIdx.x; i < FF_SR * FF_SC; i += FF_TY * FF_TX) {
                const int r = sr0 + i / FF_SC, c = sc0 + i % FF_SC;
                const bool in = r < wr1 && c < wc1;
#pragma unroll
                for (int ch = 0; ch < C; ++ch) sv[ch][i] = in ? (float)img[((size_t)r * W + c) * C + ch] : 0.f;
            }
            __syncthreads();
            const int nr = min(FF_SR, wr1 - sr0), nc = min(FF_SC, wc1 - sc0);
            float Dc = 0.f, Qc[C];
#pragma unroll
            for (int ch = 0; ch < C; ++ch) Qc[ch] = 0.f;
            for (int i = 0; i < nr; ++i) {
                const float dr = (float)(sr0 + i - pr);
                const float er = dr * dr;
                if (KIND != GL_PHOTOMETRIC && er > R2) continue;
                const float* row = &sv[0][i * FF_SC];
                float dc = (float)(sc0 - pc);
                for (int j = 0; j < nc; ++j, dc += 1.f) {
                    float x = 0.f, d0 = 0.f, dvs[C];
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) dvs[ch] = pv[ch] - row[ch * FF_SR * FF_SC + j];
                    if (KIND != GL_SPATIAL) {
                        float t = dvs[0] * dvs[0];
#pragma unroll
                        for (int ch = 1; ch < C; ++ch) t = fmaf(dvs[ch], dvs[ch], t);
                        x = t * b2;
                    }
                    if (KIND != GL_PHOTOMETRIC) {
                        d0 = fmaf(dc, dc, er);
                        x = fmaf(d0, a2, x);
                    }
                    float k = fast_exp2(x);
                    if (KIND != GL_PHOTOMETRIC && d0 > R2) k = 0.f;   // circular window: the same set of pairs for every tiling
                    Dc += k;
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) Qc[ch] = fmaf(k, dvs[ch], Qc[ch]);
                }
            }
            Dt += Dc;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) Qt[ch] += Qc[ch];
        }
    }
    if (live) {
        const size_t o = (size_t)(pr - row0) * W + pc;
        Dout[o] = Dt;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) Qout[o * C + ch] = Qt[ch];
    }
}

// fixed-order two-level sum of D over the band (fp64)
__global__ void k_full_dsum_partial(const float* __restrict__ D, int64_t count, double* __restrict__ part)
{
    __shared__ double sh[32];
    const int64_t per = (count + gridDim.x - 1) / gridDim.x;
    const int64_t a = per * blockIdx.x, b = min(count, a + per);
    double acc = 0.0;
    for (int64_t i = a + threadIdx.x; i < b; i += blockDim.x) acc += (double)D[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
        part[blockIdx.x] = s;
    }
}
__global__ void k_full_dsum_final(const double* __restrict__ part, int nparts, double* __restrict__ out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < nparts; ++i) s += part[i];
        out[0] = s;
    }
}
// alpha = n / sum(D)   (hpc/laplacian.c:57: 1 / mean)
__global__ void k_full_alpha(const double* __restrict__ dsum, double n, double* __restrict__ alpha)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) alpha[0] = n / dsum[0];
}
// z = clip(y - alpha Q, 0, 255)   (hpc/display.c:138-145)
__global__ void k_full_result(const uint8_t* __restrict__ y, const float* __restrict__ Q, const double* __restrict__ alpha, int64_t count,
                              float* __restrict__ z, uint8_t* __restrict__ z8)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float a = (float)alpha[0];
    float v = (float)y[i] - a * Q[i];
    v = fminf(fmaxf(v, 0.f), 255.f);
    z[i] = v;
    if (z8) z8[i] = (uint8_t)v;
}

int gl_impl_full_affinity(gl_ctx* ctx, int kind, double h_loc, double h_val, gl_mat** K_out)
{
    const int C = ctx->channels, W = ctx->width, H = ctx->height;
    const int64_t n_band = ctx->q1 - ctx->q0;
    int R = 1 << 30;
    if (kind != GL_PHOTOMETRIC) {
        const double rc = std::ceil(h_loc * std::sqrt(9.0 * 2.302585092994046));   // exp(-(r_c/h_loc)^2) = 1e-9
        R = rc < 1e9 ? (int)rc : (1 << 30);
    } else {
        GL_REQUIRE(ctx->n <= (int64_t)1 << 18, "full (no-approx) path with photometric affinity is O(n^2): image of %lld pixels is too large",
                   (long long)ctx->n);
    }
    if (R > W + H) R = W + H;
    gl_mat* K = gl_mat_new(ctx, GL_MAT_FULL);
    K->rows = K->cols = ctx->n;
    K->local_rows = n_band;
    K->ld = 1;
    K->elem_bytes = 4;
    K->q0 = ctx->q0;
    K->channels = C;
    K->image_epoch = ctx->image_epoch;
    K->aff_kind = kind;
    K->aff_h_loc = h_loc;
    K->aff_h_val = h_val;
    int rc = gl_alloc(ctx, sizeof(float) * (size_t)n_band, &K->buf);                       // D
    if (rc == GL_OK) rc = gl_alloc(ctx, sizeof(float) * (size_t)n_band * C, &K->aux);      // Q
    if (rc != GL_OK) {
        gl_mat_destroy(K);
        return rc;
    }
    const int tiles_x = (W + FF_TX - 1) / FF_TX, tiles_y = (ctx->row1 - ctx->row0 + FF_TY - 1) / FF_TY;
    const float log2e = 1.4426950408889634f;
    const float a2 = (float)(-log2e / (h_loc * h_loc)), b2 = (float)(-log2e / (h_val * h_val));
    const uint8_t* img = (const uint8_t*)ctx->img->ptr;
    const unsigned grid = (unsigned)(tiles_x * tiles_y);
#define FF_CASE(KK, CC)                                                                                                  \
    if (kind == KK && C == CC)                                                                                           \
        k_full_pass<KK, CC><<<grid, FF_TY * FF_TX, 0, ctx->stream>>>(img, W, H, ctx->row0, ctx->row1, R, a2, b2, (float*)K->buf->ptr, \
                                                                     (float*)K->aux->ptr);
    FF_CASE(GL_BILATERAL, 1) else FF_CASE(GL_BILATERAL, 3) else FF_CASE(GL_PHOTOMETRIC, 1) else FF_CASE(GL_PHOTOMETRIC, 3)
    else FF_CASE(GL_SPATIAL, 1) else FF_CASE(GL_SPATIAL, 3)
#undef FF_CASE
    GL_LAUNCH_CHECK(ctx);
    *K_out = K;
    return GL_OK;
}

int gl_impl_full_laplacian(gl_ctx* ctx, gl_mat* K, gl_mat** L_out)
{
    const int64_t n_band = K->local_rows;
    gl_buf *part = nullptr, *dsum = nullptr, *alpha = nullptr;
    const int nparts = 256;
    int rc = gl_alloc(ctx, sizeof(double) * nparts, &part);
    if (rc == GL_OK) rc = gl_alloc(ctx, sizeof(double), &dsum);
    if (rc == GL_OK) rc = gl_alloc(ctx, sizeof(double), &alpha);
    if (rc == GL_OK) {
        k_full_dsum_partial<<<nparts, 256, 0, ctx->stream>>>((const float*)K->buf->ptr, n_band, (double*)part->ptr);
        ctx->launches++;
        k_full_dsum_final<<<1, 32, 0, ctx->stream>>>((const double*)part->ptr, nparts, (double*)dsum->ptr);
        ctx->launches++;
        rc = gl_allreduce_f64(ctx, (double*)dsum->ptr, 1);   // the one reduction that crosses GPUs
    }
    if (rc == GL_OK) {
        k_full_alpha<<<1, 32, 0, ctx->stream>>>((const double*)dsum->ptr, (double)ctx->n, (double*)alpha->ptr);
        ctx->launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            gl_set_error("full laplacian: %s", cudaGetErrorString(e));
            rc = GL_ERR_CUDA;
        }
    }
    if (part) gl_buf_release(part);
    if (dsum) gl_buf_release(dsum);
    if (rc != GL_OK) {
        if (alpha) gl_buf_release(alpha);
        return rc;
    }
    gl_mat* L = gl_mat_new(ctx, GL_MAT_FULL);
    *L = *K;
    L->refs = 1;
    L->buf->refs++;
    L->aux->refs++;
    L->dscale = alpha;          // alpha on the device
    L->scale_on_host = false;
    *L_out = L;
    return GL_OK;
}

int gl_impl_full_result(gl_ctx* ctx, gl_mat* L, float* z_f32, uint8_t* z_u8)
{
    const int C = L->channels;
    const int64_t count = L->local_rows * C;
    gl_buf *z = nullptr, *z8 = nullptr;
    int rc = gl_alloc(ctx, sizeof(float) * (size_t)count, &z);
    if (rc == GL_OK && z_u8) rc = gl_alloc(ctx, (size_t)count, &z8);
    if (rc == GL_OK) {
        const uint8_t* y = (const uint8_t*)ctx->img->ptr + (size_t)L->q0 * C;
        k_full_result<<<(unsigned)ceil_div(count, 256), 256, 0, ctx->stream>>>(y, (const float*)L->aux->ptr, (const double*)L->dscale->ptr,
                                                                               count, (float*)z->ptr, z8 ? (uint8_t*)z8->ptr : nullptr);
        ctx->launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            gl_set_error("full result: %s", cudaGetErrorString(e));
            rc = GL_ERR_CUDA;
        }
    }
    if (rc == GL_OK && z_f32)
        if (cudaMemcpyAsync(z_f32 + (size_t)L->q0 * C, z->ptr, sizeof(float) * (size_t)count, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = GL_ERR_CUDA;
    if (rc == GL_OK && z_u8)
        if (cudaMemcpyAsync(z_u8 + (size_t)L->q0 * C, z8->ptr, (size_t)count, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = GL_ERR_CUDA;
    if (rc == GL_OK && (z_f32 || z_u8) && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = GL_ERR_CUDA;
    if (rc == GL_ERR_CUDA && !*gl_last_error()) gl_set_error("full result: copy failed");
    if (z) gl_buf_release(z);
    if (z8) gl_buf_release(z8);
    return rc;
}
