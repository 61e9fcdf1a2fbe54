// a-8: orthonormalisation of the columns of Phi (n x m).
// Replaces OrthonormaliseVecs / Projection, hpc/gram_schmidt.c:11-64: classical Gram-Schmidt,
//   u_k = v_k - sum_{j<k} <v_k,u_j>/<u_j,u_j> u_j,  u_k /= |u_k|,  norms[k] = |u_k| before normalising,
// which the reference runs as m^2 VecDot calls (two allreduces each).  In exact arithmetic that is the QR
// factorisation Phi = Q R with R upper triangular and a positive diagonal, norms = diag(R).  Computed here as
// a blocked CholeskyQR, which needs ONE reduction over pixels (and over ranks) instead of m^2:
//   1. G = Phi^T Phi        pixel-parallel partial dot products (warp/CTA tiles), fixed-order slab reduction,
//                           one m_pad^2 allreduce;
//   2. R = chol(G), T = R^-1   m x m, fp64, on device (replicated, deterministic);
//   3. Q = Phi + Phi (T - I)   tensor-core GEMM (nystroem_gemm.cu) with the identity part added in the
//                           epilogue, so that only the small correction E = T - I is rounded to fp16.
// Phi is nearly orthonormal on entry (|Phi^T Phi - I|_F ~ 1e-3..1e-2), so G is well
// conditioned and one CholeskyQR pass is stable; the result is orthonormal up to the fp16 storage of Q.
#include "common.cuh"

// ---- 1. Gram matrix: CTA = 64 x 64 tile of G over a slab of rows; 256 threads, 4 x 4 outputs each ----------
__global__ void __launch_bounds__(256) k_gram_tile(const __half* __restrict__ phi, int64_t rows, int m_pad, int slabs,
                                                   float* __restrict__ partial /* [slabs][m_pad][m_pad] */)
{
    __shared__ float As[32][64 + 4], Bs[32][64 + 4];
    const int ti = blockIdx.x, tj = blockIdx.y, slab = blockIdx.z;
    if (tj < ti) return;  // upper triangle only; mirrored by k_gram_reduce
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t per = (rows + slabs - 1) / slabs;
    const int64_t r_begin = per * slab, r_end = min(rows, r_begin + per);
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    for (int64_t r0 = r_begin; r0 < r_end; r0 += 32) {
        // 32 rows x 64 columns of each operand: 256 threads x one 16-byte load (8 bf16) each
        {
            const int rr = threadIdx.x >> 3, cg = threadIdx.x & 7;
            const int64_t r = r0 + rr;
            uint4 va = make_uint4(0, 0, 0, 0), vb = make_uint4(0, 0, 0, 0);
            if (r < r_end) {
                va = *(const uint4*)(phi + (size_t)r * m_pad + ti * 64 + cg * 8);
                vb = *(const uint4*)(phi + (size_t)r * m_pad + tj * 64 + cg * 8);
            }
            const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 fa = __half22float2(*(const __half2*)&wa[k]), fb = __half22float2(*(const __half2*)&wb[k]);
                As[rr][cg * 8 + 2 * k] = fa.x;
                As[rr][cg * 8 + 2 * k + 1] = fa.y;
                Bs[rr][cg * 8 + 2 * k] = fb.x;
                Bs[rr][cg * 8 + 2 * k + 1] = fb.y;
            }
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            const float4 a = *(const float4*)&As[k][ty * 4];
            const float4 b = *(const float4*)&Bs[k][tx * 4];
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(av[x], bv[y], acc[x][y]);
        }
        __syncthreads();
    }
    float* out = partial + (size_t)slab * m_pad * m_pad;
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) out[(size_t)(ti * 64 + ty * 4 + x) * m_pad + tj * 64 + tx * 4 + y] = acc[x][y];
}

// G[i][j] (fp64, full symmetric) = sum over slabs in fixed order; lower triangle mirrored
__global__ void k_gram_reduce(const float* __restrict__ partial, int slabs, int m_pad, double* __restrict__ G)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= m_pad) return;
    const int a = min(i, j) , b = max(i, j);
    // tile (a/64, b/64) was computed iff a/64 <= b/64, always true for a <= b
    double s = 0.0;
    for (int k = 0; k < slabs; ++k) s += (double)partial[(size_t)k * m_pad * m_pad + (size_t)a * m_pad + b];
    G[(size_t)i * m_pad + j] = s;
}

// ---- 2. R = chol(G) (upper, G = R^T R) in place in the upper triangle, one CTA, fp64 ---------------------------
__global__ void __launch_bounds__(1024, 1) k_cholesky_upper(double* __restrict__ G, int m, int ld, int* __restrict__ status)
{
    __shared__ double s_diag;
    __shared__ double red[32];
    for (int k = 0; k < m; ++k) {
        // diagonal: R[k][k] = sqrt(G[k][k] - sum_{i<k} R[i][k]^2)
        double acc = 0.0;
        for (int i = threadIdx.x; i < k; i += blockDim.x) {
            const double v = G[(size_t)i * ld + k];
            acc += v * v;
        }
        acc = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < 32; ++w) s += red[w];
            const double d = G[(size_t)k * ld + k] - s;
            if (!(d > 0.0)) { *status = k + 1; s_diag = 1.0; }
            else s_diag = sqrt(d);
            G[(size_t)k * ld + k] = s_diag;
        }
        __syncthreads();
        const double inv = 1.0 / s_diag;
        // row k right of the diagonal: R[k][j] = (G[k][j] - sum_{i<k} R[i][k] R[i][j]) / R[k][k]
        for (int j = k + 1 + threadIdx.x; j < m; j += blockDim.x) {
            double s = G[(size_t)k * ld + j];
            for (int i = 0; i < k; ++i) s -= G[(size_t)i * ld + k] * G[(size_t)i * ld + j];
            G[(size_t)k * ld + j] = s * inv;
        }
        __syncthreads();
    }
}

// T = R^-1 (upper triangular), one thread per column j; T row-major [m][ld]
__global__ void k_upper_inverse(const double* __restrict__ R, int m, int ld, double* __restrict__ T)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    for (int i = m - 1; i > j; --i) T[(size_t)i * ld + j] = 0.0;
    T[(size_t)j * ld + j] = 1.0 / R[(size_t)j * ld + j];
    for (int i = j - 1; i >= 0; --i) {
        double s = 0.0;
        for (int k = i + 1; k <= j; ++k) s += R[(size_t)i * ld + k] * T[(size_t)k * ld + j];
        T[(size_t)i * ld + j] = -s / R[(size_t)i * ld + i];
    }
}

// Et[j][k] = fp16(2^12 * (T[k][j] - delta_kj)) for k, j < m; 0 in the padding (K-major B operand of Q = Phi + Phi E).
// E is ~1e-3: the 2^12 keeps it in fp16's normal range; the GEMM epilogue multiplies by 2^-12.
#define ET_SCALE_LOG2 12
__global__ void k_build_et(const double* __restrict__ T, int m, int ld, int m_pad, __half* __restrict__ Et,
                           const double* __restrict__ R, double* __restrict__ norms)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (k >= m_pad) return;
    double v = 0.0;
    if (k < m && j < m) v = T[(size_t)k * ld + j] - (k == j ? 1.0 : 0.0);
    v = ldexp(v, ET_SCALE_LOG2);
    v = fmin(fmax(v, -60000.0), 60000.0);
    Et[(size_t)j * m_pad + k] = __float2half_rn((float)v);
    if (norms && k == j && j < m) norms[j] = R[(size_t)j * ld + j];
}

// Q = Phi T  =>  Q^T y = T^T (Phi^T y): out[j][ch] = sum_{k<=j} T[k][j] c[k][ch]
__global__ void k_proj_times_t(const double* __restrict__ T, int m, int ld, int m_pad, int C, const double* __restrict__ c,
                               double* __restrict__ out)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m_pad) return;
    for (int ch = 0; ch < C; ++ch) {
        double s = 0.0;
        if (j < m)
            for (int k = 0; k <= j; ++k) s += T[(size_t)k * ld + j] * c[(size_t)k * C + ch];
        out[(size_t)j * C + ch] = s;
    }
}

__global__ void k_et_scales(float* s)
{
    s[0] = ldexpf(1.f, ET_SCALE_LOG2);
    s[1] = ldexpf(1.f, -ET_SCALE_LOG2);
}

int gl_impl_orthonormalise(gl_ctx* ctx, gl_mat* phi, double* norms_out)
{
    const int m = phi->m, m_pad = phi->m_pad;
    const int64_t rows = phi->local_rows;
    GL_REQUIRE(m_pad % 64 == 0, "orthonormalise: m_pad %d", m_pad);
    const int tiles = m_pad / 64;
    int slabs = (int)((4 * (int64_t)ctx->sm_count) / ((int64_t)tiles * (tiles + 1) / 2) + 1);
    if (slabs > 64) slabs = 64;
    if ((int64_t)slabs * 32 > rows) slabs = (int)ceil_div(rows, 32);
    if (slabs < 1) slabs = 1;

    gl_buf *partial = nullptr, *G = nullptr, *T = nullptr, *Et = nullptr, *Q = nullptr, *st = nullptr, *norms = nullptr, *sc = nullptr;
    int rc = GL_OK;
    do {
        if ((rc = gl_alloc(ctx, sizeof(float) * (size_t)slabs * m_pad * m_pad, &partial)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(double) * (size_t)m_pad * m_pad, &G)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(double) * (size_t)m_pad * m_pad, &T)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(__half) * (size_t)m_pad * m_pad, &Et)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(__half) * (size_t)rows * m_pad, &Q)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(int) * 4, &st)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(double) * (size_t)m_pad, &norms)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(float) * 4, &sc)) != GL_OK) break;
        GL_CUDA_CHECK(cudaMemsetAsync(st->ptr, 0, sizeof(int) * 4, ctx->stream));

        dim3 gg((unsigned)tiles, (unsigned)tiles, (unsigned)slabs);
        k_gram_tile<<<gg, 256, 0, ctx->stream>>>((const __half*)phi->buf->ptr, rows, m_pad, slabs, (float*)partial->ptr);
        GL_LAUNCH_CHECK(ctx);
        dim3 gr((unsigned)ceil_div(m_pad, 128), (unsigned)m_pad);
        k_gram_reduce<<<gr, 128, 0, ctx->stream>>>((const float*)partial->ptr, slabs, m_pad, (double*)G->ptr);
        GL_LAUNCH_CHECK(ctx);
        if ((rc = gl_allreduce_f64(ctx, (double*)G->ptr, (size_t)m_pad * m_pad)) != GL_OK) break;

        k_cholesky_upper<<<1, 1024, 0, ctx->stream>>>((double*)G->ptr, m, m_pad, (int*)st->ptr);
        GL_LAUNCH_CHECK(ctx);
        k_upper_inverse<<<(unsigned)ceil_div(m, 128), 128, 0, ctx->stream>>>((const double*)G->ptr, m, m_pad, (double*)T->ptr);
        GL_LAUNCH_CHECK(ctx);
        dim3 ge((unsigned)ceil_div(m_pad, 128), (unsigned)m_pad);
        k_build_et<<<ge, 128, 0, ctx->stream>>>((const double*)T->ptr, m, m_pad, m_pad, (__half*)Et->ptr,
                                                (const double*)G->ptr, (double*)norms->ptr);
        GL_LAUNCH_CHECK(ctx);
        k_et_scales<<<1, 1, 0, ctx->stream>>>((float*)sc->ptr);
        GL_LAUNCH_CHECK(ctx);
        if ((rc = gl_gemm_kmajor(ctx, phi->buf->ptr, 0, rows, m_pad, Et->ptr, m_pad, (const float*)sc->ptr, phi->buf->ptr,
                                 Q->ptr)) != GL_OK) break;

        GL_CHECK(gl_ensure_pinned(ctx, sizeof(double) * (size_t)m_pad + 64));
        GL_CUDA_CHECK(cudaMemcpyAsync(ctx->pinned, st->ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        GL_CUDA_CHECK(cudaMemcpyAsync((char*)ctx->pinned + 64, norms->ptr, sizeof(double) * (size_t)m, cudaMemcpyDeviceToHost,
                                      ctx->stream));
        GL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        const int status = *(int*)ctx->pinned;
        if (status != 0) {
            gl_set_error("orthonormalise: Gram matrix not positive definite at column %d (Phi rank deficient)", status - 1);
            rc = GL_ERR_NOTCONVERGED;
            break;
        }
        if (norms_out) memcpy(norms_out, (char*)ctx->pinned + 64, sizeof(double) * (size_t)m);
        if (phi->proj) {
            gl_buf* np_ = nullptr;
            if ((rc = gl_alloc(ctx, sizeof(double) * (size_t)m_pad * phi->channels, &np_)) != GL_OK) break;
            k_proj_times_t<<<(unsigned)ceil_div(m_pad, 128), 128, 0, ctx->stream>>>((const double*)T->ptr, m, m_pad, m_pad, phi->channels,
                                                                                    (const double*)phi->proj->ptr, (double*)np_->ptr);
            ctx->launches++;
            gl_buf_release(phi->proj);
            phi->proj = np_;
        }
        // Phi <- Q (swap storage; the handle keeps its identity)
        gl_buf* old = phi->buf;
        phi->buf = Q;
        Q = old;
    } while (0);
    if (partial) gl_buf_release(partial);
    if (G) gl_buf_release(G);
    if (T) gl_buf_release(T);
    if (Et) gl_buf_release(Et);
    if (Q) gl_buf_release(Q);
    if (st) gl_buf_release(st);
    if (norms) gl_buf_release(norms);
    if (sc) gl_buf_release(sc);
    return rc;
}
