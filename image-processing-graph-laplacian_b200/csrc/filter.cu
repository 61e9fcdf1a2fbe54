// a-9: filter application  z = y + gain * Phi (f(lambda) o (Phi^T y)),  z[z > 255] = 255.
// Replaces ComputeResultFromLaplacian (hpc/display.c:58-83) with pngbytes2OneColMat (hpc/utils.c:461-486),
// AboveXSetY (:652-703) and OneColMat2pngbytes (:492-534).
//
// Two bandwidth-bound passes over Phi (fp16, [band pixels][m_pad], one 16-byte load = 8 columns):
//   (i)  c = Phi^T y : thread <-> (row lane, column group); 8*C fp32 accumulators per thread, rows streamed with
//        4 loads in flight per thread; per-CTA partials, fixed-order reduction (deterministic), one small
//        allreduce over ranks;
//   (ii) z = y + Phi w,  w = gain * f(lambda) o c held in registers; row dot products reduced with warp shuffles.
// Algorithmic bytes: 2 * rows * m_pad * 2 (Phi read twice) + rows*C*(1 + 1 + 4) (y twice, z once).
#include "common.cuh"

#define FL_THREADS 256
#define FL_RU 4  // rows in flight per thread

struct FilterGeom {
    int G;      // 16-byte column groups per row = m_pad / 8
    int TPR;    // threads per row = min(G, 256)
    int RL;     // row lanes per CTA = 256 / TPR
    int NG;     // groups per thread = ceil(G / TPR) (1 or 2)
};

static FilterGeom filter_geom(int m_pad)
{
    FilterGeom g;
    g.G = m_pad / 8;
    g.TPR = g.G < FL_THREADS ? g.G : FL_THREADS;
    g.RL = FL_THREADS / g.TPR;
    g.NG = (g.G + g.TPR - 1) / g.TPR;
    return g;
}

__device__ __forceinline__ uint4 ld_stream(const void* p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8])
{
    const float2 a = __half22float2(*(const __half2*)&v.x), b = __half22float2(*(const __half2*)&v.y);
    const float2 c = __half22float2(*(const __half2*)&v.z), d = __half22float2(*(const __half2*)&v.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

// (i) partial[blockIdx][j*C + ch] = sum over this CTA's rows of Phi[row][j] * y[row][ch]
template <int C, int NG>
__global__ void __launch_bounds__(FL_THREADS) k_filter_project(const __half* __restrict__ phi, int64_t rows, int m_pad, int G,
                                                               int TPR, int RL, const uint8_t* __restrict__ y /* band base */,
                                                               float* __restrict__ partial)
{
    __shared__ float red[FL_THREADS * 8 * C];  // only used when RL > 1 (then NG == 1)
    const int tid = threadIdx.x;
    const int rl = tid / TPR, tg = tid - rl * TPR;
    const bool active = rl < RL;
    float acc[NG][8][C];
#pragma unroll
    for (int g = 0; g < NG; ++g)
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int ch = 0; ch < C; ++ch) acc[g][k][ch] = 0.f;

    const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
    const int64_t r_begin = per * blockIdx.x, r_end = min(rows, r_begin + per);
    if (active) {
        for (int64_t r0 = r_begin + rl; r0 < r_end; r0 += (int64_t)RL * FL_RU) {
            uint4 v[FL_RU][NG];
            float yy[FL_RU][C];
#pragma unroll
            for (int u = 0; u < FL_RU; ++u) {
                const int64_t r = r0 + (int64_t)u * RL;
                const bool ok = r < r_end;
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const int grp = tg + g * TPR;
                    v[u][g] = (ok && grp < G) ? ld_stream(phi + (size_t)r * m_pad + (size_t)grp * 8) : make_uint4(0, 0, 0, 0);
                }
#pragma unroll
                for (int ch = 0; ch < C; ++ch) yy[u][ch] = ok ? (float)y[(size_t)r * C + ch] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < FL_RU; ++u)
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    float f[8];
                    unpack8(v[u][g], f);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
#pragma unroll
                        for (int ch = 0; ch < C; ++ch) acc[g][k][ch] = fmaf(f[k], yy[u][ch], acc[g][k][ch]);
                }
        }
    }
    float* out = partial + (size_t)blockIdx.x * m_pad * C;
    if (RL == 1) {
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const int grp = tg + g * TPR;
            if (grp < G)
#pragma unroll
                for (int k = 0; k < 8; ++k)
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) out[(size_t)(grp * 8 + k) * C + ch] = acc[g][k][ch];
        }
    } else {
        // fixed-order reduction over the row lanes
        if (active) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
#pragma unroll
                for (int ch = 0; ch < C; ++ch) red[((size_t)rl * TPR + tg) * 8 * C + k * C + ch] = acc[0][k][ch];
        }
        __syncthreads();
        for (int i = tid; i < TPR * 8 * C; i += FL_THREADS) {
            float s = 0.f;
            for (int l = 0; l < RL; ++l) s += red[(size_t)l * TPR * 8 * C + i];
            out[i] = s;  // i = (grp*8 + k)*C + ch
        }
    }
}

// c[i] = sum_b partial[b][i]  (fixed order)
__global__ void k_filter_reduce(const float* __restrict__ partial, int nblocks, int count, float* __restrict__ c)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += (double)partial[(size_t)b * count + i];
    c[i] = (float)s;
}

// w[j][ch] = gain * f[j] * c[j][ch] for j < m, 0 for padding columns
__global__ void k_filter_weights(const float* __restrict__ c, const double* __restrict__ f, int m, int m_pad, int C, float gain,
                                 float* __restrict__ w)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m_pad * C) return;
    const int j = i / C;
    w[i] = j < m ? (float)((double)gain * f[j] * (double)c[i]) : 0.f;
}

// (ii) z[row][ch] = y + sum_j Phi[row][j] * w[j][ch]; clip; optional u8
template <int C, int NG>
__global__ void __launch_bounds__(FL_THREADS) k_filter_apply(const __half* __restrict__ phi, int64_t rows, int m_pad, int G,
                                                             int TPR, int RL, const uint8_t* __restrict__ y,
                                                             const float* __restrict__ w, int clip_low, float* __restrict__ z,
                                                             uint8_t* __restrict__ z8)
{
    __shared__ float red[FL_RU][FL_THREADS / 32][C];  // cross-warp partials when a row spans several warps
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rl = tid / TPR, tg = tid - rl * TPR;
    const bool active = rl < RL;
    float wr[NG][8][C];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const int grp = tg + g * TPR;
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int ch = 0; ch < C; ++ch) wr[g][k][ch] = (active && grp < G) ? w[(size_t)(grp * 8 + k) * C + ch] : 0.f;
    }
    const int seg = TPR < 32 ? TPR : 32;      // lanes of one warp that share a row
    const int wpr = TPR > 32 ? TPR / 32 : 1;  // warps per row

    const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
    const int64_t r_begin = per * blockIdx.x, r_end = min(rows, r_begin + per);
    const int64_t step = (int64_t)RL * FL_RU;
    const int64_t iters = (r_end - r_begin + step - 1) / step;
    for (int64_t itn = 0; itn < iters; ++itn) {
        const int64_t r0 = r_begin + itn * step + rl;
        float dot[FL_RU][C];
        uint4 v[FL_RU][NG];
#pragma unroll
        for (int u = 0; u < FL_RU; ++u) {
            const int64_t r = r0 + (int64_t)u * RL;
            const bool ok = active && r < r_end;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const int grp = tg + g * TPR;
                v[u][g] = (ok && grp < G) ? ld_stream(phi + (size_t)r * m_pad + (size_t)grp * 8) : make_uint4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int u = 0; u < FL_RU; ++u) {
#pragma unroll
            for (int ch = 0; ch < C; ++ch) dot[u][ch] = 0.f;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                float f[8];
                unpack8(v[u][g], f);
#pragma unroll
                for (int k = 0; k < 8; ++k)
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) dot[u][ch] = fmaf(f[k], wr[g][k][ch], dot[u][ch]);
            }
#pragma unroll
            for (int ch = 0; ch < C; ++ch)
                for (int o = seg >> 1; o > 0; o >>= 1) dot[u][ch] += __shfl_xor_sync(0xffffffffu, dot[u][ch], o);
        }
        if (wpr > 1) {
            // RL is 1 or 2 here; a row's warps are consecutive
            if (lane == 0) {
#pragma unroll
                for (int u = 0; u < FL_RU; ++u)
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) red[u][warp][ch] = dot[u][ch];
            }
            __syncthreads();
            if (active && tg == 0) {
#pragma unroll
                for (int u = 0; u < FL_RU; ++u)
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) {
                        float s = 0.f;
                        for (int k = 0; k < wpr; ++k) s += red[u][rl * wpr + k][ch];
                        dot[u][ch] = s;
                    }
            }
        }
        if (active && tg == 0) {
#pragma unroll
            for (int u = 0; u < FL_RU; ++u) {
                const int64_t r = r0 + (int64_t)u * RL;
                if (r < r_end) {
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) {
                        float val = (float)y[(size_t)r * C + ch] + dot[u][ch];
                        val = fminf(val, 255.f);                  // AboveXSetY(z, 255, 255), display.c:76
                        if (clip_low) val = fmaxf(val, 0.f);
                        if (z) z[(size_t)r * C + ch] = val;
                        if (z8) z8[(size_t)r * C + ch] = (uint8_t)fminf(fmaxf(val, 0.f), 255.f);  // clamp then truncate
                    }
                }
            }
        }
        if (wpr > 1) __syncthreads();
    }
}

// (ii), fast path when a row is a whole number of warps' worth of 16-byte groups (m_pad % 256 == 0): one WARP per
// row, lane l owns groups l, l+32, ...; w lives in registers; no block-level synchronisation at all, 2 rows
// (2 * NGL independent 16-byte loads per lane) in flight.
// A wide Phi (more than 12 / C groups per lane) is processed in column chunks of NGL * 256 columns, one launch each: the
// first chunks leave the partial dot in z (`first` = start from 0, `last` = add y, clip, write the result).
template <int C, int NGL>
__global__ void __launch_bounds__(FL_THREADS) k_filter_apply_warp(const __half* __restrict__ phi, int64_t rows, int m_pad,
                                                                  const uint8_t* __restrict__ y, const float* __restrict__ w,
                                                                  int clip_low, float* __restrict__ z, uint8_t* __restrict__ z8,
                                                                  int col0 = 0, int first = 1, int last = 1)
{
    phi += col0;
    w += (size_t)col0 * C;
    constexpr int RU = 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int WPC = FL_THREADS / 32;
    float wr[NGL][8][C];
#pragma unroll
    for (int g = 0; g < NGL; ++g)
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int ch = 0; ch < C; ++ch) wr[g][k][ch] = w[(size_t)((lane + 32 * g) * 8 + k) * C + ch];
    const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
    const int64_t r_begin = per * blockIdx.x, r_end = min(rows, r_begin + per);
    for (int64_t r0 = r_begin + warp; r0 < r_end; r0 += (int64_t)WPC * RU) {
        uint4 v[RU][NGL];
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            const int64_t r = r0 + (int64_t)u * WPC;
#pragma unroll
            for (int g = 0; g < NGL; ++g)
                v[u][g] = r < r_end ? ld_stream(phi + (size_t)r * m_pad + (size_t)(lane + 32 * g) * 8) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < RU; ++u) {
            float dot[C];
#pragma unroll
            for (int ch = 0; ch < C; ++ch) dot[ch] = 0.f;
#pragma unroll
            for (int g = 0; g < NGL; ++g) {
                float f[8];
                unpack8(v[u][g], f);
#pragma unroll
                for (int k = 0; k < 8; ++k)
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) dot[ch] = fmaf(f[k], wr[g][k][ch], dot[ch]);
            }
#pragma unroll
            for (int ch = 0; ch < C; ++ch) dot[ch] = warp_sum(dot[ch]);
            const int64_t r = r0 + (int64_t)u * WPC;
            if (lane < C && r < r_end) {
                float d = dot[0];
#pragma unroll
                for (int ch = 1; ch < C; ++ch) d = lane == ch ? dot[ch] : d;
                if (!first) d += z[(size_t)r * C + lane];
                if (!last) {
                    z[(size_t)r * C + lane] = d;          // partial dot of the column chunks so far
                } else {
                    float val = (float)y[(size_t)r * C + lane] + d;
                    val = fminf(val, 255.f);                  // AboveXSetY(z, 255, 255), display.c:76
                    if (clip_low) val = fmaxf(val, 0.f);
                    if (z) z[(size_t)r * C + lane] = val;
                    if (z8) z8[(size_t)r * C + lane] = (uint8_t)fminf(fmaxf(val, 0.f), 255.f);
                }
            }
        }
    }
}

// c (fp32 [m_pad][C]) <- proj (fp64)
__global__ void k_filter_c_from_proj(const double* __restrict__ proj, int count, float* __restrict__ c)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) c[i] = (float)proj[i];
}

// ---- filter applied inside the extrapolation GEMM (nystroem_gemm.cu, FC > 0): weights first, partial sums afterwards ----
__global__ void k_filter_weights_from_proj(const double* __restrict__ proj, const double* __restrict__ f, int m, int m_pad, int C,
                                           double gain, float* __restrict__ w)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m_pad * C) return;
    const int j = i / C;
    w[i] = j < m ? (float)(gain * f[j] * (double)(float)proj[i]) : 0.f;   // c is rounded to fp32 exactly as in the staged path
}

int gl_filter_weights_from_proj(gl_ctx* ctx, const double* proj, const double* f, double gain, int m, int m_pad, int C, float* w)
{
    k_filter_weights_from_proj<<<(unsigned)ceil_div(m_pad * C, 256), 256, 0, ctx->stream>>>(proj, f, m, m_pad, C, gain, w);
    GL_LAUNCH_CHECK(ctx);
    return GL_OK;
}

// z[row][ch] = y + sum over parts (fixed order); clip; optional u8
__global__ void k_filter_sum_parts(const float* __restrict__ zpart, int parts, int64_t rows, int C, const uint8_t* __restrict__ y,
                                   int clip_low, float* __restrict__ z, uint8_t* __restrict__ z8)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * C) return;
    float s = 0.f;
    for (int k = 0; k < parts; ++k) s += zpart[(size_t)k * rows * C + i];
    float val = fminf((float)y[i] + s, 255.f);        // AboveXSetY(z, 255, 255), display.c:76
    if (clip_low) val = fmaxf(val, 0.f);
    z[i] = val;
    if (z8) z8[i] = (uint8_t)fminf(fmaxf(val, 0.f), 255.f);
}

// the sample pixels' rows of Phi are Phi_A, not the extrapolation (nystroem.c:25-34): z[s_i] = y + U[i, :] . w
__global__ void k_filter_sample_rows(const float* __restrict__ U, int ldU, int p, int m, const uint32_t* __restrict__ samples,
                                     int64_t q0, int64_t q1, int C, const float* __restrict__ w, const uint8_t* __restrict__ img,
                                     int clip_low, float* __restrict__ z, uint8_t* __restrict__ z8)
{
    const int i = blockIdx.x;
    const int64_t q = samples[i];
    if (q < q0 || q >= q1) return;
    __shared__ float red[3][32];
    float acc[3] = {0.f, 0.f, 0.f};
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        const float u = __half2float(__float2half_rn(U[(size_t)j * ldU + i]));   // the value stored in Phi
        for (int ch = 0; ch < C; ++ch) acc[ch] = fmaf(u, w[(size_t)j * C + ch], acc[ch]);
    }
    for (int ch = 0; ch < C; ++ch) {
        const float v = warp_sum(acc[ch]);
        if ((threadIdx.x & 31) == 0) red[ch][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if ((int)threadIdx.x < C) {
        float s = 0.f;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += red[threadIdx.x][k];
        const size_t o = (size_t)(q - q0) * C + threadIdx.x;
        float val = fminf((float)img[(size_t)q * C + threadIdx.x] + s, 255.f);
        if (clip_low) val = fmaxf(val, 0.f);
        z[o] = val;
        if (z8) z8[o] = (uint8_t)fminf(fmaxf(val, 0.f), 255.f);
    }
}

int gl_filter_fused_finish(gl_ctx* ctx, gl_mat* phi, const float* zpart, int parts, const float* w, const float* U, int ldU,
                           int clip_low, float* z_f32, uint8_t* z_u8, gl_buf* z_dev, gl_buf* z8_dev, uint8_t* z8_direct)
{
    const int C = ctx->channels;
    const int64_t rows = phi->local_rows;
    gl_buf *z = z_dev, *z8 = z8_dev;      // (ownership passes to this function)
    int rc = GL_OK;
    do {
        if (!z && (rc = gl_alloc(ctx, sizeof(float) * (size_t)rows * C, &z)) != GL_OK) break;
        if (z_u8 && !z8 && !z8_direct && (rc = gl_alloc(ctx, (size_t)rows * C, &z8)) != GL_OK) break;
        uint8_t* z8p = z8_direct ? z8_direct : (z8 ? (uint8_t*)z8->ptr : nullptr);
        {
            StageTimer t(ctx, GL_T_FILTER);
            const uint8_t* y = (const uint8_t*)ctx->img->ptr + (size_t)phi->q0 * C;
            if (parts > 0) {
                k_filter_sum_parts<<<(unsigned)ceil_div(rows * C, 256), 256, 0, ctx->stream>>>(zpart, parts, rows, C, y, clip_low, (float*)z->ptr,
                                                                                              z8p);
                GL_LAUNCH_CHECK(ctx);
            }
            k_filter_sample_rows<<<phi->p, 128, 0, ctx->stream>>>(U, ldU, phi->p, phi->m, (const uint32_t*)ctx->samples->ptr, phi->q0,
                                                                 phi->q0 + rows, C, w, (const uint8_t*)ctx->img->ptr, clip_low,
                                                                 (float*)z->ptr, z8p);
            GL_LAUNCH_CHECK(ctx);
        }
        ctx->ev_valid[GL_T_K_FILTER_PROJECT] = false;
        ctx->ev_valid[GL_T_K_FILTER_APPLY] = false;
        {
            StageTimer t(ctx, GL_T_D2H);
            if (z_f32)
                GL_CUDA_BREAK(rc, cudaMemcpyAsync(z_f32 + (size_t)phi->q0 * C, z->ptr, sizeof(float) * (size_t)rows * C, cudaMemcpyDeviceToHost,
                                              ctx->stream));
            if (z_u8 && !z8_direct)
                GL_CUDA_BREAK(rc, cudaMemcpyAsync(z_u8 + (size_t)phi->q0 * C, z8->ptr, (size_t)rows * C, cudaMemcpyDeviceToHost, ctx->stream));
        }
        if (z_f32 || z_u8 || z8_direct) GL_CUDA_BREAK(rc, cudaStreamSynchronize(ctx->stream));
    } while (0);
    if (z) gl_buf_release(z);
    if (z8) gl_buf_release(z8);
    return rc;
}

template <int C, int NG>
static int run_filter(gl_ctx* ctx, gl_mat* phi, const FilterGeom& g, const double* f, double gain, int clip_low, int grid,
                      float* partial, float* c, float* w, float* z, uint8_t* z8, bool use_proj)
{
    const int64_t rows = phi->local_rows;
    const int m = phi->m, m_pad = phi->m_pad;
    const uint8_t* y = (const uint8_t*)ctx->img->ptr + (size_t)phi->q0 * C;
    const __half* P = (const __half*)phi->buf->ptr;
    if (use_proj) {
        // c = Phi^T y was derived from the affinity stage's sums (nystroem_gemm.cu): no pass over Phi, no allreduce
        k_filter_c_from_proj<<<(unsigned)ceil_div(m_pad * C, 256), 256, 0, ctx->stream>>>((const double*)phi->proj->ptr, m_pad * C, c);
        GL_LAUNCH_CHECK(ctx);
    } else {
        {
            StageTimer kt(ctx, GL_T_K_FILTER_PROJECT);
            k_filter_project<C, NG><<<grid, FL_THREADS, 0, ctx->stream>>>(P, rows, m_pad, g.G, g.TPR, g.RL, y, partial);
        }
        GL_LAUNCH_CHECK(ctx);
        k_filter_reduce<<<(unsigned)ceil_div(m_pad * C, 256), 256, 0, ctx->stream>>>(partial, grid, m_pad * C, c);
        GL_LAUNCH_CHECK(ctx);
        GL_CHECK(gl_allreduce_f32(ctx, c, (size_t)m_pad * C));
    }
    k_filter_weights<<<(unsigned)ceil_div(m_pad * C, 256), 256, 0, ctx->stream>>>(c, f, m, m_pad, C, (float)gain, w);
    GL_LAUNCH_CHECK(ctx);
    {
        StageTimer kt(ctx, GL_T_K_FILTER_APPLY);
        const int ngl = (g.G % 32 == 0) ? g.G / 32 : 0;   // 16-byte groups per lane when a warp owns a row
        int wgrid = ctx->sm_count * 8;
        if ((int64_t)wgrid * 16 > rows) wgrid = (int)ceil_div(rows, 16);
        constexpr int MAXG = 12 / C;   // groups per lane whose weights fit in registers
        if (ctx->filter_apply_impl == 1 || ngl == 0) {
            k_filter_apply<C, NG><<<grid, FL_THREADS, 0, ctx->stream>>>(P, rows, m_pad, g.G, g.TPR, g.RL, y, w, clip_low, z, z8);
        } else {
            // column chunks of at most MAXG groups per lane (one chunk unless Phi is very wide or has three channels)
            for (int g0 = 0; g0 < ngl; g0 += MAXG) {
                const int n = ngl - g0 < MAXG ? ngl - g0 : MAXG;
                const int col0 = g0 * 256, first = g0 == 0, last = g0 + n >= ngl;
#define FLW_CASE(N)                                                                                                                  \
    if (n == N) {                                                                                                                      \
        if constexpr (N <= MAXG)                                                                                                       \
            k_filter_apply_warp<C, N><<<wgrid, FL_THREADS, 0, ctx->stream>>>(P, rows, m_pad, y, w, clip_low, z, z8, col0, first, last); \
    }
                FLW_CASE(1) FLW_CASE(2) FLW_CASE(3) FLW_CASE(4) FLW_CASE(5) FLW_CASE(6) FLW_CASE(7) FLW_CASE(8) FLW_CASE(9)
                FLW_CASE(10) FLW_CASE(11) FLW_CASE(12)
#undef FLW_CASE
                if (!last) ctx->launches++;
            }
        }
    }
    GL_LAUNCH_CHECK(ctx);
    return GL_OK;
}

int gl_impl_filter(gl_ctx* ctx, gl_mat* phi, gl_mat* f_eigvals, double gain, int clip_low, float* z_f32, uint8_t* z_u8)
{
    const int C = ctx->channels;
    const int64_t rows = phi->local_rows;
    const int m_pad = phi->m_pad;
    const FilterGeom g = filter_geom(m_pad);
    const bool use_proj_pre = ctx->projection_mode == 0 && phi->proj && phi->channels == C && phi->image_epoch == ctx->image_epoch;
    // the thread-per-column-group kernels (projection pass, generic apply) hold at most two groups per thread (m_pad <= 4096);
    // wider Phi goes through the projection from the affinity sums and the column-chunked warp apply only
    const bool needs_generic = !use_proj_pre || ctx->filter_apply_impl == 1 || g.G % 32 != 0;
    if (g.NG > 2 && needs_generic) {
        gl_set_error("filter: m_pad = %d is too wide for the stand-alone projection / generic apply kernels (<= 4096)", m_pad);
        return GL_ERR_UNSUPPORTED;
    }
    GL_REQUIRE((g.TPR & (g.TPR - 1)) == 0 || g.TPR % 32 == 0, "filter: unsupported column-group count %d", g.G);
    int grid = ctx->sm_count * 4;
    if ((int64_t)grid * g.RL * FL_RU > rows) grid = (int)ceil_div(rows, (int64_t)g.RL * FL_RU);
    if (grid < 1) grid = 1;

    gl_buf *partial = nullptr, *c = nullptr, *w = nullptr, *z = nullptr, *z8 = nullptr;
    int rc = GL_OK;
    const bool use_proj = ctx->projection_mode == 0 && phi->proj && phi->channels == C && phi->image_epoch == ctx->image_epoch;
    if (use_proj) ctx->ev_valid[GL_T_K_FILTER_PROJECT] = false;  // that kernel does not run
    do {
        if ((rc = gl_alloc(ctx, sizeof(float) * (size_t)grid * m_pad * C, &partial)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(float) * (size_t)m_pad * C, &c)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(float) * (size_t)m_pad * C, &w)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(float) * (size_t)rows * C, &z)) != GL_OK) break;
        if (z_u8 && (rc = gl_alloc(ctx, (size_t)rows * C, &z8)) != GL_OK) break;
        {
            StageTimer t(ctx, GL_T_FILTER);
            const double* f = (const double*)f_eigvals->buf->ptr;
            float* zp = (float*)z->ptr;
            uint8_t* z8p = z8 ? (uint8_t*)z8->ptr : nullptr;
#define FL_CASE(CC, NGG)                                                                                                      \
    if (C == CC && (g.NG == NGG || (NGG == 1 && g.NG > 2)))                                                                   \
        rc = run_filter<CC, NGG>(ctx, phi, g, f, gain, clip_low, grid, (float*)partial->ptr, (float*)c->ptr, (float*)w->ptr, zp, z8p, use_proj);
            FL_CASE(1, 1) else FL_CASE(1, 2) else FL_CASE(3, 1) else FL_CASE(3, 2)
#undef FL_CASE
        }
        if (rc != GL_OK) break;
        {
            StageTimer t(ctx, GL_T_D2H);
            if (z_f32)
                GL_CUDA_BREAK(rc, cudaMemcpyAsync(z_f32 + (size_t)phi->q0 * C, z->ptr, sizeof(float) * (size_t)rows * C,
                                              cudaMemcpyDeviceToHost, ctx->stream));
            if (z_u8)
                GL_CUDA_BREAK(rc, cudaMemcpyAsync(z_u8 + (size_t)phi->q0 * C, z8->ptr, (size_t)rows * C, cudaMemcpyDeviceToHost,
                                              ctx->stream));
        }
        if (z_f32 || z_u8) GL_CUDA_BREAK(rc, cudaStreamSynchronize(ctx->stream));
    } while (0);
    if (partial) gl_buf_release(partial);
    if (c) gl_buf_release(c);
    if (w) gl_buf_release(w);
    if (z) gl_buf_release(z);
    if (z8) gl_buf_release(z8);
    return rc;
}
