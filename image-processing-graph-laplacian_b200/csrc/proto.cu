// Device versions of the experimental blocks of the reference's Python prototype:
//
//   sinkhorn(phi, Pi)             python/image_processing.py:90-107    gl_sinkhorn
//   orthogonalisation(A, B)       python/image_processing.py:110-127   gl_orthogonalisation
//   smoothing_matrix(s, phi, Pi)  python/image_processing.py:151-194   gl_smoothing_matrix
//   smoothing / sharpening        python/image_processing.py:197-241   gl_matrix_filter (polynomial of W = V L V^T applied to y)
//
// None of them is on the hot path (the prototype's main never calls the first two, and the other two are alternatives to its
// filter line), so they are built from what the path already has: Phi stays the fp16 [band pixels][m_pad] matrix in raster order,
// every n x m product with a small matrix runs on the tcgen05 GEMM of nystroem_gemm.cu, the n-vectors (r, c, degrees, iterates)
// are fp64 and the products Phi^T x / Phi w are bandwidth-bound passes with fp64 accumulation, everything p x p is fp64
// (dense_small.cu) and the p x p eigensolves are the block Jacobi.  Rows are in raster order throughout: the prototype's
// "sample rows first" order and its permutation() disappear, the sample rows are addressed through ctx->samples.
// Multi-GPU: pixel bands as everywhere; the m- and p-sized reductions are allreduced.
#include <cfloat>
#include <cmath>

#include "common.cuh"

void gl_phi_describe(gl_ctx* ctx, gl_mat* phi, const gl_mat* L_B, const gl_mat* phi_A);   // nystroem_gemm.cu

namespace {

// ---------------------------------------------------------------------------------------------
// n-vectors (fp64, this rank's band)
// ---------------------------------------------------------------------------------------------
__global__ void k_fill64(double* __restrict__ x, int64_t n, double v)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = v;
}
__global__ void k_channel64(const uint8_t* __restrict__ img, int64_t q0, int64_t rows, int C, int ch, double* __restrict__ y)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) y[i] = (double)img[(size_t)(q0 + i) * C + ch];
}
// out = nan_to_num(1 / u)  (python/image_processing.py:98-101: inf -> the largest double, nan -> 0)
__global__ void k_recip_nan_to_num(const double* __restrict__ u, int64_t n, double* __restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = 1.0 / u[i];
    if (isnan(v)) v = 0.0;
    else if (isinf(v)) v = copysign(DBL_MAX, v);
    out[i] = v;
}
__global__ void k_axpy64(double a, const double* __restrict__ x, int64_t n, double* __restrict__ acc, int first)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) acc[i] = (first ? 0.0 : acc[i]) + a * x[i];
}
__global__ void k_store_channel(const double* __restrict__ z, int64_t rows, int C, int ch, float* __restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) out[(size_t)i * C + ch] = (float)z[i];
}

// partial[b][m_pad] = sum over the rows of block b of Phi[r][:] x[r]  (x == nullptr: ones).  A thread owns one 16-byte column group
// of one row lane; the lanes of a block are summed in shared memory, the blocks by k_sum_partials (fixed order: reproducible)
__global__ void __launch_bounds__(256) k_phiT_x(const __half* __restrict__ phi, int64_t rows, int m_pad, const double* __restrict__ x,
                                                int64_t rows_per_block, double* __restrict__ partial)
{
    __shared__ double red[256 * 8];
    const int G8 = m_pad / 8;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
    for (int g0 = 0; g0 < G8; g0 += 256) {
        const int gs = G8 - g0 < 256 ? G8 - g0 : 256;
        const int RL = 256 / gs;
        const int lane = threadIdx.x / gs, g = threadIdx.x % gs;
        double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (lane < RL) {
            for (int64_t r = r0 + lane; r < r1; r += RL) {
                const uint4 v = *reinterpret_cast<const uint4*>(phi + (size_t)r * m_pad + (size_t)(g0 + g) * 8);
                const __half2* h = reinterpret_cast<const __half2*>(&v);
                const double xv = x ? x[r] : 1.0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 f = __half22float2(h[k]);
                    acc[2 * k] = fma((double)f.x, xv, acc[2 * k]);
                    acc[2 * k + 1] = fma((double)f.y, xv, acc[2 * k + 1]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) red[threadIdx.x * 8 + k] = acc[k];
        __syncthreads();
        if (threadIdx.x < gs) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                double s = 0.0;
                for (int l = 0; l < RL; ++l) s += red[(l * gs + g) * 8 + k];
                partial[(size_t)blockIdx.x * m_pad + (size_t)(g0 + g) * 8 + k] = s;
            }
        }
        __syncthreads();
    }
}
// out[i] = d[i] * sum_b partial[b][i]   (d == nullptr: 1)
__global__ void k_sum_partials(const double* __restrict__ partial, int nb, int count, const double* __restrict__ d, double* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double s = 0.0;
    for (int b = 0; b < nb; ++b) s += partial[(size_t)b * count + i];
    out[i] = d ? d[i] * s : s;
}
__global__ void k_mul64(const double* __restrict__ d, int n, double* __restrict__ t)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) t[i] *= d[i];
}
// u[r] = Phi[r][:] . w   (one warp per row, fp64 accumulation)
__global__ void __launch_bounds__(256) k_phi_w(const __half* __restrict__ phi, int64_t rows, int m_pad, const double* __restrict__ w,
                                               double* __restrict__ u)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < rows; r += nwarps) {
        double s = 0.0;
        for (int g = lane; g < m_pad / 8; g += 32) {
            const uint4 v = *reinterpret_cast<const uint4*>(phi + (size_t)r * m_pad + (size_t)g * 8);
            const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 f = __half22float2(h[k]);
                s = fma((double)f.x, w[g * 8 + 2 * k], s);
                s = fma((double)f.y, w[g * 8 + 2 * k + 1], s);
            }
        }
        s = warp_sum(s);
        if (lane == 0) u[r] = s;
    }
}
// rows of Phi / entries of an n-vector at the sample pixels of this band (zero elsewhere: summed over ranks afterwards)
__global__ void k_gather_rows(const __half* __restrict__ phi, int64_t q0, int64_t rows, int m_pad, const uint32_t* __restrict__ samples,
                              double* __restrict__ out /* [p][m_pad] */)
{
    const int i = blockIdx.x;
    const int64_t q = samples[i];
    const bool mine = q >= q0 && q < q0 + rows;
    for (int j = threadIdx.x; j < m_pad; j += blockDim.x)
        out[(size_t)i * m_pad + j] = mine ? (double)__half2float(phi[(size_t)(q - q0) * m_pad + j]) : 0.0;
}
__global__ void k_gather_vec(const double* __restrict__ x, int64_t q0, int64_t rows, const uint32_t* __restrict__ samples, int p,
                             double* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p) return;
    const int64_t q = samples[i];
    out[i] = (q >= q0 && q < q0 + rows) ? x[q - q0] : 0.0;
}
// out[0] = sum, out[1] = max |.| of an array (one CTA, fixed order)
__global__ void __launch_bounds__(1024) k_sum_absmax(const double* __restrict__ a, int64_t n, double* __restrict__ out)
{
    __shared__ double rs[32], rm[32];
    double s = 0.0, mx = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) { s += a[i]; mx = fmax(mx, fabs(a[i])); }
    s = warp_sum(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s; rm[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0, m = 0.0;
        for (int w = 0; w < 32; ++w) { t += rs[w]; m = fmax(m, rm[w]); }
        out[0] = t;
        out[1] = m;
    }
}

// ---------------------------------------------------------------------------------------------
// small fp64 matrices (row-major)
// ---------------------------------------------------------------------------------------------
__global__ void k_cm32_to_rm64(const float* __restrict__ U, int ld, int rows, int cols, double* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows * cols) out[i] = (double)U[(size_t)(i % cols) * ld + i / cols];
}
__global__ void k_rm64_to_cm32(const double* __restrict__ A, int rows, int cols, int ld, float* __restrict__ U)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows * cols) U[(size_t)(i % cols) * ld + i / cols] = (float)A[i];
}
// A[i][j] *= (rs ? rs[i] : 1) * (cs ? cs[j] : 1) * a
__global__ void k_scale_rc(double* __restrict__ A, int rows, int cols, const double* __restrict__ rs, const double* __restrict__ cs, double a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows * cols) A[i] *= a * (rs ? rs[i / cols] : 1.0) * (cs ? cs[i % cols] : 1.0);
}
// A = a A + b I - c diag(d)
__global__ void k_affine_diag(double* __restrict__ A, int p, double a, double b, double c, const double* __restrict__ d)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p * p) return;
    const int r = i / p, col = i % p;
    A[i] = a * A[i] + (r == col ? b - (d ? c * d[r] : 0.0) : 0.0);
}
__global__ void k_symmetrise(double* __restrict__ A, int p)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p * p) return;
    const int r = i / p, c = i % p;
    if (r < c) { const double v = 0.5 * (A[i] + A[(size_t)c * p + r]); A[i] = v; A[(size_t)c * p + r] = v; }
}
__global__ void k_map_diag(const double* __restrict__ d, int n, int op, double* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = d[i];
    out[i] = op == 0 ? 1.0 / v : (op == 1 ? 1.0 / sqrt(v) : fmin(v, 1.0));
}
// Bt[i][k] (fp16, [n_pad][k_pad]) = s * Mt[i][k] for i < q, k < k_pad; zero rows beyond q
__global__ void k_to_half_scaled(const double* __restrict__ Mt, int q, int k_pad, int n_pad, double s, __half* __restrict__ Bt)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_pad * k_pad) return;
    Bt[i] = __float2half_rn(i / k_pad < q ? (float)(s * Mt[i]) : 0.f);
}
__global__ void k_set_scales(float a, float b, float* __restrict__ scales)
{
    scales[0] = a; scales[1] = b; scales[2] = 0.f; scales[3] = 0.f;
}
// D[r][:] *= rs[r] * inv
__global__ void k_rowscale_half(__half* __restrict__ D, int64_t rows, int n_pad, const double* __restrict__ rs, double inv)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * (n_pad / 2)) return;
    const int64_t r = i / (n_pad / 2);
    const float f = (float)(rs[r] * inv);
    __half2* h = reinterpret_cast<__half2*>(D) + i;
    const float2 v = __half22float2(*h);
    *h = __floats2half2_rn(v.x * f, v.y * f);
}
// rows of the sample pixels <- given fp64 rows (times s)
__global__ void k_override_rows(const double* __restrict__ src /* [p][q] */, int q, const uint32_t* __restrict__ samples, int64_t q0, int64_t rows,
                                int n_pad, double s, __half* __restrict__ D)
{
    const int i = blockIdx.x;
    const int64_t px = samples[i];
    if (px < q0 || px >= q0 + rows) return;
    for (int j = threadIdx.x; j < q; j += blockDim.x) D[(size_t)(px - q0) * n_pad + j] = __float2half_rn((float)(s * src[(size_t)i * q + j]));
}

// Kc[r][i] = K(pixel qa + r, sample i) in fp64 as the reference writes it (hpc/affinity.c:59-122); rows of sample pixels are zeroed
// when `skip_samples` (they belong to K_A, not K_B)
template <int C>
__global__ void k_rows_affinity64(const uint8_t* __restrict__ img, const uint32_t* __restrict__ samples, int p, int width, int kind,
                                  double inv_hl2, double inv_hv2, int64_t qa, int64_t nrows, int skip_samples, double* __restrict__ Kc)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nrows * p) return;
    const int64_t q = qa + idx / p;
    const int i = (int)(idx % p);
    if (skip_samples) {   // binary search of q among the ascending samples
        int lo = 0, hi = p;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int64_t)samples[mid] < q) lo = mid + 1; else hi = mid; }
        if (lo < p && (int64_t)samples[lo] == q) { Kc[idx] = 0.0; return; }
    }
    const int64_t b = samples[i];
    double k = 1.0;
    if (kind != GL_PHOTOMETRIC) {
        const double dr = (double)(q / width) - (double)(b / width), dc = (double)(q % width) - (double)(b % width);
        k *= exp(-(dr * dr + dc * dc) * inv_hl2);
    }
    if (kind != GL_SPATIAL) {
        double d2 = 0.0;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
            const double dv = (double)img[(size_t)q * C + ch] - (double)img[(size_t)b * C + ch];
            d2 += dv * dv;
        }
        k *= exp(-d2 * inv_hv2);
    }
    Kc[idx] = k;
}
__global__ void k_f64_to_half_rows(const double* __restrict__ src, int64_t nrows, int q, int n_pad, __half* __restrict__ dst)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows * n_pad) return;
    const int64_t r = i / n_pad;
    const int c = (int)(i % n_pad);
    dst[i] = __float2half_rn(c < q ? (float)src[r * q + c] : 0.f);
}

// ---------------------------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------------------------
struct Bufs {   // buffers released together
    std::vector<gl_buf*> v;
    ~Bufs() { for (gl_buf* b : v) if (b) gl_buf_release(b); }
    int get(gl_ctx* ctx, size_t bytes, void** out)
    {
        gl_buf* b = nullptr;
        GL_CHECK(gl_alloc(ctx, bytes ? bytes : 8, &b));
        v.push_back(b);
        *out = b->ptr;
        return GL_OK;
    }
    gl_buf* take_last() { gl_buf* b = v.back(); v.pop_back(); return b; }
};
#define ALLOC(var, type, count) type* var = nullptr; GL_CHECK(bufs.get(ctx, sizeof(type) * (size_t)(count), (void**)&var))

inline unsigned nblk(int64_t n, int t = 256) { return (unsigned)ceil_div(n, t); }

int fetch(gl_ctx* ctx, const double* dev, int n, double* host)
{
    GL_CHECK(gl_ensure_pinned(ctx, sizeof(double) * (size_t)n));
    GL_CUDA_CHECK(cudaMemcpyAsync(ctx->pinned, dev, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    GL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    memcpy(host, ctx->pinned, sizeof(double) * (size_t)n);
    return GL_OK;
}

struct PhiOps {   // Phi^T x and Phi w on one materialised Phi
    gl_ctx* ctx;
    const __half* phi;
    int64_t rows;
    int m_pad;
    int nb;
    int64_t rpb;
    double* partial;
    double* t;   // [m_pad]
    int init(gl_ctx* c, const gl_mat* P, Bufs& bufs)
    {
        ctx = c; phi = (const __half*)P->buf->ptr; rows = P->local_rows; m_pad = P->m_pad;
        nb = (int)std::min<int64_t>(ceil_div(rows, 64), (int64_t)ctx->sm_count * 4);
        rpb = ceil_div(rows, nb);
        nb = (int)ceil_div(rows, rpb);
        GL_CHECK(bufs.get(ctx, sizeof(double) * (size_t)nb * m_pad, (void**)&partial));
        GL_CHECK(bufs.get(ctx, sizeof(double) * (size_t)m_pad, (void**)&t));
        return GL_OK;
    }
    // t = d o (Phi^T x), summed over ranks
    int tx(const double* x, const double* d)
    {
        k_phiT_x<<<nb, 256, 0, ctx->stream>>>(phi, rows, m_pad, x, rpb, partial);
        GL_LAUNCH_CHECK(ctx);
        k_sum_partials<<<nblk(m_pad), 256, 0, ctx->stream>>>(partial, nb, m_pad, nullptr, t);
        GL_LAUNCH_CHECK(ctx);
        GL_CHECK(gl_allreduce_f64(ctx, t, (size_t)m_pad));
        if (d) {
            k_mul64<<<nblk(m_pad), 256, 0, ctx->stream>>>(d, m_pad, t);
            GL_LAUNCH_CHECK(ctx);
        }
        return GL_OK;
    }
    int w(const double* wv, double* u)
    {
        const int grid = (int)std::min<int64_t>(ceil_div(rows, 8), (int64_t)ctx->sm_count * 8);
        k_phi_w<<<grid, 256, 0, ctx->stream>>>(phi, rows, m_pad, wv, u);
        GL_LAUNCH_CHECK(ctx);
        return GL_OK;
    }
    // u = Phi (d o (Phi^T x))
    int apply(const double* x, const double* d, double* u)
    {
        GL_CHECK(tx(x, d));
        return w(t, u);
    }
};

// diagonal handle -> fp64 [m_pad] with zero padding
int padded_diag(gl_ctx* ctx, const gl_mat* d, int m, int m_pad, Bufs& bufs, double** out)
{
    GL_CHECK(bufs.get(ctx, sizeof(double) * (size_t)m_pad, (void**)out));
    GL_CUDA_CHECK(cudaMemsetAsync(*out, 0, sizeof(double) * (size_t)m_pad, ctx->stream));
    GL_CUDA_CHECK(cudaMemcpyAsync(*out, d->buf->ptr, sizeof(double) * (size_t)m, cudaMemcpyDeviceToDevice, ctx->stream));
    return GL_OK;
}

gl_mat* new_diag(gl_ctx* ctx, int n)
{
    gl_mat* d = gl_mat_new(ctx, GL_MAT_DIAG);
    d->rows = d->local_rows = d->cols = n;
    d->ld = 1;
    d->elem_bytes = 8;
    return d;
}

// A new Phi-shaped matrix [band rows][gl_m_pad(q)] (fp16) = rowscale o (Phi . Mt^T) on the tcgen05 GEMM.  Mt is fp64 [q][m_pad]
// (row i = column i of the small factor, zero beyond m); the stored values are `pre` times the product (row-scaled), the handle's
// scale undoes it.  `over` (fp64 [p][q], or null) replaces the rows of the sample pixels.
int phi_times_small(gl_ctx* ctx, const gl_mat* P, const double* Mt, int q, const double* rowscale, double rowscale_inv, double pre,
                    const double* over, gl_mat** out)
{
    Bufs bufs;
    const int m_pad = P->m_pad, n_pad = gl_m_pad(q);
    const int64_t rows = P->local_rows;
    ALLOC(red, double, 2);
    k_sum_absmax<<<1, 1024, 0, ctx->stream>>>(Mt, (int64_t)q * m_pad, red);
    GL_LAUNCH_CHECK(ctx);
    double h[2];
    GL_CHECK(fetch(ctx, red, 2, h));
    int e = 0;
    if (h[1] > 0.0 && std::isfinite(h[1])) { int ex; std::frexp(h[1], &ex); e = 13 - ex; }
    e = std::max(-100, std::min(100, e));
    ALLOC(Bt, __half, (size_t)n_pad * m_pad);
    ALLOC(scales, float, 4);
    k_to_half_scaled<<<nblk((int64_t)n_pad * m_pad), 256, 0, ctx->stream>>>(Mt, q, m_pad, n_pad, std::ldexp(1.0, e), Bt);
    GL_LAUNCH_CHECK(ctx);
    k_set_scales<<<1, 1, 0, ctx->stream>>>((float)std::ldexp(1.0, e), (float)(std::ldexp(1.0, -e) * pre), scales);
    GL_LAUNCH_CHECK(ctx);
    gl_mat* R = gl_mat_new(ctx, GL_MAT_PHI);
    R->rows = P->rows; R->cols = q; R->local_rows = rows; R->ld = n_pad; R->elem_bytes = 2;
    R->p = P->p; R->p_pad = P->p_pad; R->m = q; R->m_pad = n_pad; R->q0 = P->q0;
    R->scale = 1.0 / (pre * rowscale_inv);
    int rc = gl_alloc(ctx, sizeof(__half) * (size_t)rows * n_pad, &R->buf);
    if (rc == GL_OK) rc = gl_gemm_kmajor(ctx, P->buf->ptr, 0, rows, m_pad, Bt, n_pad, scales, nullptr, R->buf->ptr);
    if (rc != GL_OK) { gl_mat_destroy(R); return rc; }
    if (rowscale) {
        k_rowscale_half<<<nblk(rows * (n_pad / 2)), 256, 0, ctx->stream>>>((__half*)R->buf->ptr, rows, n_pad, rowscale, rowscale_inv);
        GL_LAUNCH_CHECK(ctx);
    }
    if (over) {
        k_override_rows<<<P->p, 128, 0, ctx->stream>>>(over, q, (const uint32_t*)ctx->samples->ptr, P->q0, rows, n_pad, pre * rowscale_inv,
                                                       (__half*)R->buf->ptr);
        GL_LAUNCH_CHECK(ctx);
    }
    *out = R;
    return GL_OK;
}

int check_phi(gl_ctx* ctx, gl_mat* phi, const gl_mat* Pi, const char* who)
{
    GL_REQUIRE(phi && phi->kind == GL_MAT_PHI && Pi && Pi->kind == GL_MAT_DIAG, "%s: want a Phi and a diagonal handle", who);
    GL_CHECK(gl_phi_materialise(ctx, phi));
    GL_REQUIRE(phi->buf && phi->scale == 1.0, "%s: Phi is not stored", who);
    GL_REQUIRE(Pi->rows == phi->m, "%s: Phi has %d columns but the diagonal %lld entries", who, phi->m, (long long)Pi->rows);
    GL_REQUIRE(ctx->samples && (int)ctx->p == phi->p, "%s: the context's samples are not the ones Phi was built from", who);
    return GL_OK;
}

// the p x p eigensolve of the prototype blocks: every pair, descending, converged as far as the fp32 rotations go
int eig_desc(gl_ctx* ctx, gl_mat* A, gl_mat** U, gl_mat** L, gl_mat** Linv)
{
    const int keep_l = ctx->eig_largest;
    const float keep_t = ctx->jacobi_tol;
    const bool keep_a = ctx->async_mode;
    ctx->eig_largest = 1;
    ctx->jacobi_tol = std::min(keep_t, 2e-6f);
    ctx->async_mode = false;
    const int rc = gl_impl_eigensolve(ctx, A, (int)A->rows, U, L, Linv);
    ctx->eig_largest = keep_l;
    ctx->jacobi_tol = keep_t;
    ctx->async_mode = keep_a;
    return rc;
}

gl_mat* wrap_ka(gl_ctx* ctx, gl_buf* buf, int p)
{
    gl_mat* A = gl_mat_new(ctx, GL_MAT_KA);
    A->rows = A->cols = A->local_rows = p;
    A->ld = p;
    A->elem_bytes = 8;
    A->buf = buf;
    return A;
}
}  // namespace

// ---------------------------------------------------------------------------------------------
// sinkhorn(phi, Pi), python/image_processing.py:90-107
// ---------------------------------------------------------------------------------------------
int gl_impl_sinkhorn(gl_ctx* ctx, gl_mat* phi, gl_mat* Pi, int iterations, gl_mat** W_A_out, gl_mat** W_ABt_out)
{
    GL_CHECK(check_phi(ctx, phi, Pi, "sinkhorn"));
    GL_REQUIRE(iterations >= 0, "sinkhorn: iterations < 0");
    Bufs bufs;
    const int p = phi->p, m = phi->m, m_pad = phi->m_pad;
    const int64_t rows = phi->local_rows;
    PhiOps ops;
    GL_CHECK(ops.init(ctx, phi, bufs));
    double* d = nullptr;
    GL_CHECK(padded_diag(ctx, Pi, m, m_pad, bufs, &d));
    ALLOC(r, double, rows);
    ALLOC(c, double, rows);
    ALLOC(u, double, rows);
    k_fill64<<<nblk(rows), 256, 0, ctx->stream>>>(r, rows, 1.0);
    GL_LAUNCH_CHECK(ctx);
    GL_CUDA_CHECK(cudaMemcpyAsync(c, r, sizeof(double) * (size_t)rows, cudaMemcpyDeviceToDevice, ctx->stream));
    for (int it = 0; it < iterations; ++it) {   // :96-100
        GL_CHECK(ops.apply(r, d, u));
        k_recip_nan_to_num<<<nblk(rows), 256, 0, ctx->stream>>>(u, rows, c);
        GL_LAUNCH_CHECK(ctx);
        GL_CHECK(ops.apply(c, d, u));
        k_recip_nan_to_num<<<nblk(rows), 256, 0, ctx->stream>>>(u, rows, r);
        GL_LAUNCH_CHECK(ctx);
    }
    // the sample rows of Phi and of r, c (replicated)
    const uint32_t* samples = (const uint32_t*)ctx->samples->ptr;
    ALLOC(PS, double, (size_t)p * m_pad);
    ALLOC(rS, double, p);
    ALLOC(cS, double, p);
    k_gather_rows<<<p, 128, 0, ctx->stream>>>(ops.phi, phi->q0, rows, m_pad, samples, PS);
    GL_LAUNCH_CHECK(ctx);
    k_gather_vec<<<nblk(p), 256, 0, ctx->stream>>>(r, phi->q0, rows, samples, p, rS);
    GL_LAUNCH_CHECK(ctx);
    k_gather_vec<<<nblk(p), 256, 0, ctx->stream>>>(c, phi->q0, rows, samples, p, cS);
    GL_LAUNCH_CHECK(ctx);
    GL_CHECK(gl_allreduce_f64(ctx, PS, (size_t)p * m_pad));
    GL_CHECK(gl_allreduce_f64(ctx, rS, (size_t)p));
    GL_CHECK(gl_allreduce_f64(ctx, cS, (size_t)p));
    // Mt[i][:] = r_i Phi_S[i][:] o Pi : row i of W_AB is Mt[i] . (Phi o c)^T  (:101-104)
    ALLOC(Mt, double, (size_t)p * m_pad);
    GL_CUDA_CHECK(cudaMemcpyAsync(Mt, PS, sizeof(double) * (size_t)p * m_pad, cudaMemcpyDeviceToDevice, ctx->stream));
    k_scale_rc<<<nblk((int64_t)p * m_pad), 256, 0, ctx->stream>>>(Mt, p, m_pad, rS, d, 1.0);
    GL_LAUNCH_CHECK(ctx);
    // W_A = (Mt Phi_S^T) diag(c_S)
    gl_buf* wa = nullptr;
    GL_CHECK(gl_alloc(ctx, sizeof(double) * (size_t)p * p, &wa));
    gl_mat* WA = wrap_ka(ctx, wa, p);
    int rc = gl_dgemm(ctx, p, p, m_pad, 1.0, Mt, m_pad, 0, PS, m_pad, 1, 0.0, (double*)wa->ptr, p);
    if (rc == GL_OK) {
        k_scale_rc<<<nblk((int64_t)p * p), 256, 0, ctx->stream>>>((double*)wa->ptr, p, p, nullptr, cS, 1.0);
        ctx->launches++;
    }
    // W_AB^T for every band pixel: c_j (Phi[j] . Mt[i]); stored relative to max r * max c so that fp16 keeps its precision
    double rmax = 1.0, cmax = 1.0;
    ALLOC(red, double, 4);
    ALLOC(mxd, double, 2 * ctx->world);
    if (rc == GL_OK) {
        k_sum_absmax<<<1, 1024, 0, ctx->stream>>>(r, rows, red);
        k_sum_absmax<<<1, 1024, 0, ctx->stream>>>(c, rows, red + 2);
        ctx->launches += 2;
        double h4[4];
        rc = fetch(ctx, red, 4, h4);   // {sum r, max r, sum c, max c}
        rmax = h4[1];
        cmax = h4[3];
    }
    if (rc == GL_OK && ctx->world > 1) {   // the same scales on every rank
        std::vector<double> all(2 * ctx->world, 0.0);
        all[2 * ctx->rank] = rmax;
        all[2 * ctx->rank + 1] = cmax;
        GL_CUDA_CHECK(cudaMemcpyAsync(mxd, all.data(), sizeof(double) * all.size(), cudaMemcpyHostToDevice, ctx->stream));
        GL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        GL_CHECK(gl_allreduce_f64(ctx, mxd, all.size()));
        GL_CHECK(fetch(ctx, mxd, (int)all.size(), all.data()));
        for (int k = 0; k < ctx->world; ++k) { rmax = std::max(rmax, all[2 * k]); cmax = std::max(cmax, all[2 * k + 1]); }
    }
    gl_mat* WB = nullptr;
    if (rc == GL_OK) {
        if (!(rmax > 0 && std::isfinite(rmax))) rmax = 1.0;
        if (!(cmax > 0 && std::isfinite(cmax))) cmax = 1.0;
        int ex;
        std::frexp(rmax, &ex);
        rc = phi_times_small(ctx, phi, Mt, p, c, 1.0 / cmax, std::ldexp(1.0, 12 - ex), nullptr, &WB);
    }
    if (rc != GL_OK) { gl_mat_destroy(WA); if (WB) gl_mat_destroy(WB); return rc; }
    if (W_A_out) *W_A_out = WA; else gl_mat_destroy(WA);
    if (W_ABt_out) *W_ABt_out = WB; else gl_mat_destroy(WB);
    return GL_OK;
}

// ---------------------------------------------------------------------------------------------
// smoothing_matrix(sample_indices, phi, Pi), python/image_processing.py:151-194
// ---------------------------------------------------------------------------------------------
int gl_impl_smoothing_matrix(gl_ctx* ctx, gl_mat* phi, gl_mat* Pi, gl_mat** V_out, gl_mat** L_out)
{
    GL_CHECK(check_phi(ctx, phi, Pi, "smoothing_matrix"));
    Bufs bufs;
    const int p = phi->p, m = phi->m, m_pad = phi->m_pad;
    const int64_t rows = phi->local_rows;
    PhiOps ops;
    GL_CHECK(ops.init(ctx, phi, bufs));
    double* d = nullptr;
    GL_CHECK(padded_diag(ctx, Pi, m, m_pad, bufs, &d));
    // degrees D = K 1 = Phi (Pi o (Phi^T 1)), alpha = 1 / mean(D)  (:154-158)
    ALLOC(D, double, rows);
    GL_CHECK(ops.apply(nullptr, d, D));
    ALLOC(red, double, 2);
    k_sum_absmax<<<1, 1024, 0, ctx->stream>>>(D, rows, red);
    GL_LAUNCH_CHECK(ctx);
    GL_CHECK(gl_allreduce_f64(ctx, red, 1));
    double h[2];
    GL_CHECK(fetch(ctx, red, 2, h));
    GL_REQUIRE(h[0] != 0.0 && std::isfinite(h[0]), "smoothing_matrix: the degrees sum to %g", h[0]);
    const double alpha = (double)ctx->n / h[0];
    const uint32_t* samples = (const uint32_t*)ctx->samples->ptr;
    ALLOC(PS, double, (size_t)p * m_pad);
    ALLOC(DS, double, p);
    k_gather_rows<<<p, 128, 0, ctx->stream>>>(ops.phi, phi->q0, rows, m_pad, samples, PS);
    GL_LAUNCH_CHECK(ctx);
    k_gather_vec<<<nblk(p), 256, 0, ctx->stream>>>(D, phi->q0, rows, samples, p, DS);
    GL_LAUNCH_CHECK(ctx);
    GL_CHECK(gl_allreduce_f64(ctx, PS, (size_t)p * m_pad));
    GL_CHECK(gl_allreduce_f64(ctx, DS, (size_t)p));
    // W_A = I + alpha (Phi_S Pi Phi_S^T - diag D_S)  (:160-161)
    ALLOC(PSd, double, (size_t)p * m_pad);
    GL_CUDA_CHECK(cudaMemcpyAsync(PSd, PS, sizeof(double) * (size_t)p * m_pad, cudaMemcpyDeviceToDevice, ctx->stream));
    k_scale_rc<<<nblk((int64_t)p * m_pad), 256, 0, ctx->stream>>>(PSd, p, m_pad, nullptr, d, 1.0);
    GL_LAUNCH_CHECK(ctx);
    gl_buf* wa = nullptr;
    GL_CHECK(gl_alloc(ctx, sizeof(double) * (size_t)p * p, &wa));
    gl_mat* WA = wrap_ka(ctx, wa, p);
    gl_mat *U = nullptr, *L = nullptr, *Linv = nullptr, *V = nullptr;
    int rc = gl_dgemm(ctx, p, p, m_pad, 1.0, PSd, m_pad, 0, PS, m_pad, 1, 0.0, (double*)wa->ptr, p);
    do {
        if (rc != GL_OK) break;
        k_affine_diag<<<nblk((int64_t)p * p), 256, 0, ctx->stream>>>((double*)wa->ptr, p, alpha, 1.0, alpha, DS);
        k_symmetrise<<<nblk((int64_t)p * p), 256, 0, ctx->stream>>>((double*)wa->ptr, p);
        ctx->launches += 2;
        GL_BREAK(rc, eig_desc(ctx, WA, &U, &L, &Linv));   // :183-185
        // V[j] = Phi[j] . (alpha Pi Phi_S^T U diag(1/L)) for every pixel = W_B^T U / L  (:186-190); sample rows = U
        double *U64 = nullptr, *Mt = nullptr;
        GL_BREAK(rc, bufs.get(ctx, sizeof(double) * (size_t)p * p, (void**)&U64));
        GL_BREAK(rc, bufs.get(ctx, sizeof(double) * (size_t)p * m_pad, (void**)&Mt));
        k_cm32_to_rm64<<<nblk((int64_t)p * p), 256, 0, ctx->stream>>>((const float*)U->buf->ptr, (int)U->ld, p, p, U64);
        ctx->launches++;
        GL_BREAK(rc, gl_dgemm(ctx, p, m_pad, p, 1.0, U64, p, 1, PSd, m_pad, 0, 0.0, Mt, m_pad));   // Mt[i][k] = sum_s U[s][i] Pi_k Phi_S[s][k]
        k_scale_rc<<<nblk((int64_t)p * m_pad), 256, 0, ctx->stream>>>(Mt, p, m_pad, (const double*)Linv->buf->ptr, nullptr, alpha);
        ctx->launches++;
        GL_BREAK(rc, phi_times_small(ctx, phi, Mt, p, nullptr, 1.0, 1.0, U64, &V));
    } while (0);
    gl_mat_destroy(WA);
    if (U) gl_mat_destroy(U);
    if (Linv) gl_mat_destroy(Linv);
    if (rc != GL_OK) { if (L) gl_mat_destroy(L); if (V) gl_mat_destroy(V); return rc; }
    if (V_out) *V_out = V; else gl_mat_destroy(V);
    if (L_out) *L_out = L; else gl_mat_destroy(L);
    return GL_OK;
}

// ---------------------------------------------------------------------------------------------
// z = sum_k coef[k] W^k y with W = V diag(L) V^T: smoothing (:197-219) is coef = {0, 1}, sharpening (:222-241) is
// {0, 0, 1 + beta, -beta}.  Not clipped, like the prototype.  z_f32: host, n * C floats, this rank's band at its raster offset.
// ---------------------------------------------------------------------------------------------
int gl_impl_matrix_filter(gl_ctx* ctx, gl_mat* V, gl_mat* L, const double* coef, int ncoef, float* z_f32)
{
    GL_CHECK(check_phi(ctx, V, L, "matrix_filter"));
    GL_REQUIRE(coef && ncoef >= 1 && ncoef <= 16 && z_f32, "matrix_filter: want 1..16 coefficients and a destination");
    GL_REQUIRE(ctx->img, "matrix_filter: no image");
    Bufs bufs;
    const int m = V->m, m_pad = V->m_pad, C = ctx->channels;
    const int64_t rows = V->local_rows;
    PhiOps ops;
    GL_CHECK(ops.init(ctx, V, bufs));
    double* d = nullptr;
    GL_CHECK(padded_diag(ctx, L, m, m_pad, bufs, &d));
    ALLOC(x, double, rows);
    ALLOC(xn, double, rows);
    ALLOC(acc, double, rows);
    ALLOC(zf, float, (size_t)rows * C);
    for (int ch = 0; ch < C; ++ch) {
        k_channel64<<<nblk(rows), 256, 0, ctx->stream>>>((const uint8_t*)ctx->img->ptr, V->q0, rows, C, ch, x);
        GL_LAUNCH_CHECK(ctx);
        k_axpy64<<<nblk(rows), 256, 0, ctx->stream>>>(coef[0], x, rows, acc, 1);
        GL_LAUNCH_CHECK(ctx);
        for (int k = 1; k < ncoef; ++k) {
            GL_CHECK(ops.apply(x, d, xn));
            std::swap(x, xn);
            if (coef[k] != 0.0) {
                k_axpy64<<<nblk(rows), 256, 0, ctx->stream>>>(coef[k], x, rows, acc, 0);
                GL_LAUNCH_CHECK(ctx);
            }
        }
        k_store_channel<<<nblk(rows), 256, 0, ctx->stream>>>(acc, rows, C, ch, zf);
        GL_LAUNCH_CHECK(ctx);
    }
    GL_CUDA_CHECK(cudaMemcpyAsync(z_f32 + (size_t)V->q0 * C, zf, sizeof(float) * (size_t)rows * C, cudaMemcpyDeviceToHost, ctx->stream));
    GL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return GL_OK;
}

// ---------------------------------------------------------------------------------------------
// orthogonalisation(A, B), python/image_processing.py:110-127: the one-shot orthogonal Nystroem extension.
//   A^-1/2 (:113-115)           coupled Newton-Schulz iteration in fp64 (no eigenvectors of the possibly ill-conditioned A needed)
//   Q = A + A^-1/2 B B^T A^-1/2 (:117)   B B^T accumulated in fp64 from exact affinity rows, chunk by chunk, summed over ranks
//   (Phi_Q, Pi_Q) = eig(Q), descending (:118)   block Jacobi
//   V = [A; B^T] A^-1/2 Phi_Q Pi_Q^-1/2 (:122), Pi = min(Pi_Q, 1) (:123-124)   fp64 rows, stored fp16 in raster order
// K_B is only asked for its parameters (kind, bandwidths, sample set): its fp16 entries are not accurate enough for a factor that
// A^-1/2 amplifies by the square root of A's condition number.
// ---------------------------------------------------------------------------------------------
int gl_impl_orthogonalisation(gl_ctx* ctx, gl_mat* K_A, gl_mat* K_B, gl_mat** V_out, gl_mat** Pi_out)
{
    GL_REQUIRE(K_A && K_A->kind == GL_MAT_KA && K_B && K_B->kind == GL_MAT_KB, "orthogonalisation: want K_A and K_B");
    const int p = (int)K_A->rows;
    GL_REQUIRE(K_B->p == p && (int)ctx->p == p && K_B->sample_epoch == ctx->sample_epoch && K_B->image_epoch == ctx->image_epoch,
               "orthogonalisation: K_B does not belong to the current image and samples");
    GL_REQUIRE(K_B->aff_kind != GL_NLM, "orthogonalisation: not available for the NLM affinity");
    GL_REQUIRE(ctx->channels == 1 || ctx->channels == 3, "orthogonalisation: 1 or 3 channels");
    Bufs bufs;
    const size_t pp = (size_t)p * p;
    const double* A = (const double*)K_A->buf->ptr;
    ALLOC(Y, double, pp);
    ALLOC(Z, double, pp);
    ALLOC(T, double, pp);
    ALLOC(Y2, double, pp);
    ALLOC(Z2, double, pp);
    ALLOC(red, double, 2);
    // Y_0 = A / s, Z_0 = I;  T = (3 I - Z Y) / 2;  Y <- Y T, Z <- T Z;  Y -> (A/s)^1/2, Z -> (A/s)^-1/2
    k_sum_absmax<<<1, 1024, 0, ctx->stream>>>(A, (int64_t)pp, red);
    GL_LAUNCH_CHECK(ctx);
    double h[2];
    GL_CHECK(fetch(ctx, red, 2, h));
    const double s = h[1] * p;   // >= the largest row sum >= the largest eigenvalue
    GL_REQUIRE(s > 0 && std::isfinite(s), "orthogonalisation: K_A is zero or not finite");
    GL_CUDA_CHECK(cudaMemcpyAsync(Y, A, sizeof(double) * pp, cudaMemcpyDeviceToDevice, ctx->stream));
    k_scale_rc<<<nblk((int64_t)pp), 256, 0, ctx->stream>>>(Y, p, p, nullptr, nullptr, 1.0 / s);
    GL_LAUNCH_CHECK(ctx);
    GL_CUDA_CHECK(cudaMemsetAsync(Z, 0, sizeof(double) * pp, ctx->stream));
    k_affine_diag<<<nblk((int64_t)pp), 256, 0, ctx->stream>>>(Z, p, 0.0, 1.0, 0.0, nullptr);
    GL_LAUNCH_CHECK(ctx);
    double resid = 1.0;
    int it = 0;
    for (; it < 200; ++it) {
        GL_CHECK(gl_dgemm(ctx, p, p, p, -0.5, Z, p, 0, Y, p, 0, 0.0, T, p));
        k_affine_diag<<<nblk((int64_t)pp), 256, 0, ctx->stream>>>(T, p, 1.0, 1.5, 0.0, nullptr);      // T = 1.5 I - 0.5 Z Y
        GL_LAUNCH_CHECK(ctx);
        // |I - Z Y|_max = 2 |T - I|_max: read it every few steps
        GL_CHECK(gl_dgemm(ctx, p, p, p, 1.0, Y, p, 0, T, p, 0, 0.0, Y2, p));
        GL_CHECK(gl_dgemm(ctx, p, p, p, 1.0, T, p, 0, Z, p, 0, 0.0, Z2, p));
        std::swap(Y, Y2);
        std::swap(Z, Z2);
        if (it % 4 == 3 || it > 20) {
            k_affine_diag<<<nblk((int64_t)pp), 256, 0, ctx->stream>>>(T, p, 1.0, -1.0, 0.0, nullptr);
            k_sum_absmax<<<1, 1024, 0, ctx->stream>>>(T, (int64_t)pp, red);
            ctx->launches += 2;
            GL_CHECK(fetch(ctx, red, 2, h));
            resid = 2.0 * h[1];
            if (!(resid == resid)) break;
            if (resid < 1e-14) { ++it; break; }
        }
    }
    GL_REQUIRE(resid < 1e-9, "orthogonalisation: the inverse square root of K_A did not converge (residual %g after %d steps; is K_A positive "
               "definite?)", resid, it);
    if (ctx->verbose) fprintf(stderr, "[libglcuda] orthogonalisation: A^-1/2 after %d Newton-Schulz steps (residual %.2g)\n", it, resid);
    double* X = Z;   // A^-1/2 = Z / sqrt(s)
    k_scale_rc<<<nblk((int64_t)pp), 256, 0, ctx->stream>>>(X, p, p, nullptr, nullptr, 1.0 / std::sqrt(s));
    GL_LAUNCH_CHECK(ctx);
    k_symmetrise<<<nblk((int64_t)pp), 256, 0, ctx->stream>>>(X, p);
    GL_LAUNCH_CHECK(ctx);

    // G = B B^T over this band's non-sample pixels
    const int64_t rows = ctx->q1 - ctx->q0;
    const int64_t chunk = std::max<int64_t>(256, std::min<int64_t>(rows, ((int64_t)1 << 27) / std::max(p, 1)));   // <= 1 GB of fp64 rows
    ALLOC(Kc, double, (size_t)chunk * p);
    double* G = T;
    GL_CUDA_CHECK(cudaMemsetAsync(G, 0, sizeof(double) * pp, ctx->stream));
    const double inv_hl2 = 1.0 / (K_B->aff_h_loc * K_B->aff_h_loc), inv_hv2 = 1.0 / (K_B->aff_h_val * K_B->aff_h_val);
    const uint32_t* samples = (const uint32_t*)ctx->samples->ptr;
    const uint8_t* img = (const uint8_t*)ctx->img->ptr;
    auto rows_affinity = [&](int64_t qa, int64_t nr, int skip) {
        if (ctx->channels == 1)
            k_rows_affinity64<1><<<nblk(nr * p), 256, 0, ctx->stream>>>(img, samples, p, ctx->width, K_B->aff_kind, inv_hl2, inv_hv2, qa, nr, skip, Kc);
        else
            k_rows_affinity64<3><<<nblk(nr * p), 256, 0, ctx->stream>>>(img, samples, p, ctx->width, K_B->aff_kind, inv_hl2, inv_hv2, qa, nr, skip, Kc);
        ctx->launches++;
    };
    for (int64_t r0 = 0; r0 < rows; r0 += chunk) {
        const int64_t nr = std::min(chunk, rows - r0);
        rows_affinity(ctx->q0 + r0, nr, 1);
        GL_CHECK(gl_dgemm(ctx, p, p, (int)nr, 1.0, Kc, p, 1, Kc, p, 0, 1.0, G, p));
    }
    GL_CHECK(gl_allreduce_f64(ctx, G, pp));
    // Q = A + X G X
    gl_buf* qb = nullptr;
    GL_CHECK(gl_alloc(ctx, sizeof(double) * pp, &qb));
    gl_mat* Qm = wrap_ka(ctx, qb, p);
    double* Q = (double*)qb->ptr;
    gl_mat *UQ = nullptr, *PQ = nullptr, *V = nullptr, *Pi = nullptr;
    int rc = GL_OK;
    do {
        GL_BREAK(rc, gl_dgemm(ctx, p, p, p, 1.0, X, p, 0, G, p, 0, 0.0, Y, p));
        GL_CUDA_BREAK(rc, cudaMemcpyAsync(Q, A, sizeof(double) * pp, cudaMemcpyDeviceToDevice, ctx->stream));
        GL_BREAK(rc, gl_dgemm(ctx, p, p, p, 1.0, Y, p, 0, X, p, 0, 1.0, Q, p));
        k_symmetrise<<<nblk((int64_t)pp), 256, 0, ctx->stream>>>(Q, p);
        ctx->launches++;
        GL_BREAK(rc, eig_desc(ctx, Qm, &UQ, &PQ, nullptr));
        // M = X Phi_Q Pi_Q^-1/2  (p x p)
        double *U64 = Y2, *M = Z2, *isq = nullptr;
        GL_BREAK(rc, bufs.get(ctx, sizeof(double) * (size_t)p, (void**)&isq));
        k_cm32_to_rm64<<<nblk((int64_t)pp), 256, 0, ctx->stream>>>((const float*)UQ->buf->ptr, (int)UQ->ld, p, p, U64);
        k_map_diag<<<nblk(p), 256, 0, ctx->stream>>>((const double*)PQ->buf->ptr, p, 1, isq);
        ctx->launches += 2;
        GL_BREAK(rc, gl_dgemm(ctx, p, p, p, 1.0, X, p, 0, U64, p, 0, 0.0, M, p));
        k_scale_rc<<<nblk((int64_t)pp), 256, 0, ctx->stream>>>(M, p, p, nullptr, isq, 1.0);
        ctx->launches++;
        // V rows: every band pixel k_j^T M (a sample pixel's k_j is its row of A)
        const int n_pad = gl_m_pad(p);
        V = gl_mat_new(ctx, GL_MAT_PHI);
        V->rows = ctx->n; V->cols = p; V->local_rows = rows; V->ld = n_pad; V->elem_bytes = 2;
        V->p = p; V->p_pad = K_B->p_pad; V->m = p; V->m_pad = n_pad; V->q0 = ctx->q0;
        GL_BREAK(rc, gl_alloc(ctx, sizeof(__half) * (size_t)rows * n_pad, &V->buf));
        double* Vc = nullptr;
        GL_BREAK(rc, bufs.get(ctx, sizeof(double) * (size_t)chunk * p, (void**)&Vc));
        for (int64_t r0 = 0; r0 < rows && rc == GL_OK; r0 += chunk) {
            const int64_t nr = std::min(chunk, rows - r0);
            rows_affinity(ctx->q0 + r0, nr, 0);
            rc = gl_dgemm(ctx, (int)nr, p, p, 1.0, Kc, p, 0, M, p, 0, 0.0, Vc, p);
            if (rc != GL_OK) break;
            k_f64_to_half_rows<<<nblk(nr * n_pad), 256, 0, ctx->stream>>>(Vc, nr, p, n_pad, (__half*)V->buf->ptr + (size_t)r0 * n_pad);
            ctx->launches++;
        }
        if (rc != GL_OK) break;
        Pi = new_diag(ctx, p);
        GL_BREAK(rc, gl_alloc(ctx, sizeof(double) * (size_t)p, &Pi->buf));
        k_map_diag<<<nblk(p), 256, 0, ctx->stream>>>((const double*)PQ->buf->ptr, p, 2, (double*)Pi->buf->ptr);
        ctx->launches++;
    } while (0);
    gl_mat_destroy(Qm);
    if (UQ) gl_mat_destroy(UQ);
    if (PQ) gl_mat_destroy(PQ);
    const cudaError_t ce = cudaStreamSynchronize(ctx->stream);   // the chunk buffers are released below
    if (rc == GL_OK && ce != cudaSuccess) { gl_set_error("orthogonalisation: %s", cudaGetErrorString(ce)); rc = GL_ERR_CUDA; }
    if (rc != GL_OK) { if (V) gl_mat_destroy(V); if (Pi) gl_mat_destroy(Pi); return rc; }
    if (V_out) *V_out = V; else gl_mat_destroy(V);
    if (Pi_out) *Pi_out = Pi; else gl_mat_destroy(Pi);
    return GL_OK;
}
