// Small dense fp64 algebra on the device (p x p and p x m operands of the sample block) and, built on it, the reference's own
// eigensolver: the inverse subspace iteration of hpc/inverse_power_it.c:86-252 with its options -opti_gs and -inv_it_epsilon
// (SURVEY 8f-3).  The default solver of this library stays the converged block Jacobi (eigen_jacobi.cu); this one serves callers who
// want the reference's stopping rule and its partial spectra (m << p) at the cost of a few p x p x m products.
//
//   X_0 = pseudo-random p x m, orthonormalised                                   (:94-95; PETSc's rank-seeded RNG is replaced by a
//                                                                                 fixed hash: SURVEY 8c-iv)
//   r = |(I - X X^T) A X|_F                                                       (ComputeResidualsNorm, :49-80)
//   while r > epsilon:  X <- A^-1 X; keep X; every opti_gs-th step orthonormalise (norms out); r = ...     (:159-181)
//   lambda_i = 1 / norm_i, eigenvector i = the kept column i, normalised          (:186-235)
//
// A^-1 X is exact here: A = L_A is symmetric positive definite, so A = R^T R (Cholesky) and A^-1 = T T^T with T = R^-1 (the blocked
// fp64 kernels of orthonormalise.cu) -- the reference's GMRES + additive Schwarz solves approximate the same product.
// OrthonormaliseVecs (hpc/gram_schmidt.c:29-64) is classical Gram-Schmidt = the QR factorisation with a positive diagonal, computed
// as CholeskyQR: G = X^T X = R^T R, Q = X R^-1, and the "norms before normalisation" are diag(R).
#include <cmath>

#include "common.cuh"

int gl_chol_inverse_upper(gl_ctx* ctx, double* G, int m, int ld, double* T, int* status_dev);   // orthonormalise.cu

// C[M x N] (ldc) = alpha * op(A) op(B) + beta * C, row-major fp64; op = transpose when the flag is set.  16 x 16 tiles.
__global__ void __launch_bounds__(256) k_dgemm(int M, int N, int K, double alpha, const double* __restrict__ A, int lda, int ta,
                                               const double* __restrict__ B, int ldb, int tb, double beta, double* __restrict__ C, int ldc)
{
    __shared__ double As[16][17], Bs[16][17];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int i = blockIdx.y * 16 + ty, j = blockIdx.x * 16 + tx;
    double acc = 0.0;
    for (int k0 = 0; k0 < K; k0 += 16) {
        const int ka = k0 + tx, kb = k0 + ty;
        As[ty][tx] = (i < M && ka < K) ? (ta ? A[(size_t)ka * lda + i] : A[(size_t)i * lda + ka]) : 0.0;
        Bs[ty][tx] = (kb < K && j < N) ? (tb ? B[(size_t)j * ldb + kb] : B[(size_t)kb * ldb + j]) : 0.0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) acc = fma(As[ty][k], Bs[k][tx], acc);
        __syncthreads();
    }
    if (i < M && j < N) C[(size_t)i * ldc + j] = alpha * acc + (beta != 0.0 ? beta * C[(size_t)i * ldc + j] : 0.0);
}

int gl_dgemm(gl_ctx* ctx, int M, int N, int K, double alpha, const double* A, int lda, int ta, const double* B, int ldb, int tb, double beta,
             double* C, int ldc)
{
    dim3 grid((unsigned)ceil_div(N, 16), (unsigned)ceil_div(M, 16));
    k_dgemm<<<grid, 256, 0, ctx->stream>>>(M, N, K, alpha, A, lda, ta, B, ldb, tb, beta, C, ldc);
    GL_LAUNCH_CHECK(ctx);
    return GL_OK;
}

__device__ __forceinline__ uint32_t ds_hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
// X_0[i][j] uniform in (0, 1) from a fixed hash of (i, j): the same start on every rank and in the oracle (oracle_np.inverse_iteration_start)
__global__ void k_ii_start(int p, int m, double* __restrict__ X)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p * m) return;
    X[idx] = ((double)(ds_hash32((uint32_t)idx + 0x9e3779b9u) >> 8) + 0.5) / 16777216.0;
}
// sum of squares of an array -> out[0] (one CTA, fixed order)
__global__ void __launch_bounds__(1024) k_sumsq(const double* __restrict__ a, int n, double* __restrict__ out)
{
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) s += a[i] * a[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 32; ++w) t += red[w];
        out[0] = t;
    }
}
__global__ void k_diag_of(const double* __restrict__ R, int ld, int m, double* __restrict__ d)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) d[j] = R[(size_t)j * ld + j];
}
// eigenpairs out: U[j][i] (fp32 column-major, ld) = Xb[i][j] / |Xb[:, j]|, mu_j = 1 / norm_j, mu_inv_j = norm_j
__global__ void k_ii_extract(const double* __restrict__ Xb, int p, int m, int ld, const double* __restrict__ norms, float* __restrict__ U,
                             double* __restrict__ mu, double* __restrict__ mu_inv)
{
    const int j = blockIdx.x;
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < p; i += blockDim.x) { const double v = Xb[(size_t)i * m + j]; s += v * v; }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        red[0] = 1.0 / sqrt(t);
    }
    __syncthreads();
    const double inv = red[0];
    for (int i = threadIdx.x; i < ld; i += blockDim.x) U[(size_t)j * ld + i] = i < p ? (float)(Xb[(size_t)i * m + j] * inv) : 0.f;
    if (threadIdx.x == 0) { mu[j] = 1.0 / norms[j]; mu_inv[j] = norms[j]; }
}

namespace {
struct IIWork {
    gl_ctx* ctx;
    int p, m;
    double *A, *T, *X, *Xb, *Y, *G, *R, *Tg, *S, *norms, *scal;
    int* st;
};
// X <- X R^-1 with G = X^T X = R^T R; norms = diag(R)  (OrthonormaliseVecs, hpc/gram_schmidt.c:29-64)
int ii_orthonormalise(IIWork& w)
{
    GL_CHECK(gl_dgemm(w.ctx, w.m, w.m, w.p, 1.0, w.X, w.m, 1, w.X, w.m, 0, 0.0, w.G, w.m));
    GL_CHECK(gl_chol_inverse_upper(w.ctx, w.G, w.m, w.m, w.Tg, w.st));
    k_diag_of<<<(unsigned)ceil_div(w.m, 128), 128, 0, w.ctx->stream>>>(w.G, w.m, w.m, w.norms);
    w.ctx->launches++;
    GL_CHECK(gl_dgemm(w.ctx, w.p, w.m, w.m, 1.0, w.X, w.m, 0, w.Tg, w.m, 0, 0.0, w.Y, w.m));
    GL_CUDA_CHECK(cudaMemcpyAsync(w.X, w.Y, sizeof(double) * (size_t)w.p * w.m, cudaMemcpyDeviceToDevice, w.ctx->stream));
    return GL_OK;
}
// |(I - X X^T) A X|_F  (ComputeResidualsNorm, hpc/inverse_power_it.c:49-80); one 8-byte read-back
int ii_residual(IIWork& w, double* r)
{
    GL_CHECK(gl_dgemm(w.ctx, w.p, w.m, w.p, 1.0, w.A, w.p, 0, w.X, w.m, 0, 0.0, w.Y, w.m));          // Y = A X
    GL_CHECK(gl_dgemm(w.ctx, w.m, w.m, w.p, 1.0, w.X, w.m, 1, w.Y, w.m, 0, 0.0, w.S, w.m));          // S = X^T A X
    GL_CHECK(gl_dgemm(w.ctx, w.p, w.m, w.m, -1.0, w.X, w.m, 0, w.S, w.m, 0, 1.0, w.Y, w.m));         // Y -= X S
    k_sumsq<<<1, 1024, 0, w.ctx->stream>>>(w.Y, w.p * w.m, w.scal);
    w.ctx->launches++;
    GL_CHECK(gl_ensure_pinned(w.ctx, 64));
    GL_CUDA_CHECK(cudaMemcpyAsync(w.ctx->pinned, w.scal, sizeof(double), cudaMemcpyDeviceToHost, w.ctx->stream));
    GL_CUDA_CHECK(cudaStreamSynchronize(w.ctx->stream));
    *r = std::sqrt(*(const double*)w.ctx->pinned);
    return GL_OK;
}
}  // namespace

int gl_impl_inverse_iteration(gl_ctx* ctx, gl_mat* L_A, int m, int opti_gs, double epsilon, int max_iterations, gl_mat** eigvecs,
                              gl_mat** eigvals, gl_mat** eigvals_inv, int* iterations_out, double* residual_out)
{
    const int p = (int)L_A->rows;
    if (opti_gs < 1) opti_gs = 1;                                   // hpc/image_processing.c:134-138
    gl_buf* bufs[12] = {};
    auto alloc = [&](int k, size_t bytes) { return gl_alloc(ctx, bytes, &bufs[k]); };
    gl_mat *U = nullptr, *mu = nullptr, *mui = nullptr;
    int rc = GL_OK;
    do {
        const size_t pp = sizeof(double) * (size_t)p * p, pm = sizeof(double) * (size_t)p * m, mm = sizeof(double) * (size_t)m * m;
        GL_BREAK(rc, alloc(0, pp)); GL_BREAK(rc, alloc(1, pp));                    // R (Cholesky of A, in place), T = R^-1
        GL_BREAK(rc, alloc(2, pm)); GL_BREAK(rc, alloc(3, pm)); GL_BREAK(rc, alloc(4, pm));   // X, X before orthonormalisation, Y
        GL_BREAK(rc, alloc(5, mm)); GL_BREAK(rc, alloc(6, mm)); GL_BREAK(rc, alloc(7, mm));   // G, T_G, S
        GL_BREAK(rc, alloc(8, sizeof(double) * (size_t)m)); GL_BREAK(rc, alloc(9, 64)); GL_BREAK(rc, alloc(10, 64));
        IIWork w{ctx, p, m, (double*)L_A->buf->ptr, (double*)bufs[1]->ptr, (double*)bufs[2]->ptr, (double*)bufs[3]->ptr, (double*)bufs[4]->ptr,
                 (double*)bufs[5]->ptr, (double*)bufs[0]->ptr, (double*)bufs[6]->ptr, (double*)bufs[7]->ptr, (double*)bufs[8]->ptr,
                 (double*)bufs[9]->ptr, (int*)bufs[10]->ptr};
        GL_CUDA_BREAK(rc, cudaMemsetAsync(w.st, 0, 64, ctx->stream));
        // A = R^T R, T = R^-1  =>  A^-1 = T T^T
        GL_CUDA_BREAK(rc, cudaMemcpyAsync(w.R, w.A, pp, cudaMemcpyDeviceToDevice, ctx->stream));
        GL_BREAK(rc, gl_chol_inverse_upper(ctx, w.R, p, p, w.T, w.st));
        k_ii_start<<<(unsigned)ceil_div((int64_t)p * m, 256), 256, 0, ctx->stream>>>(p, m, w.X);
        ctx->launches++;
        GL_BREAK(rc, ii_orthonormalise(w));
        GL_CUDA_BREAK(rc, cudaMemcpyAsync(w.Xb, w.X, pm, cudaMemcpyDeviceToDevice, ctx->stream));
        double r = 0.0;
        GL_BREAK(rc, ii_residual(w, &r));
        int it = 0;
        while (r > epsilon && it < max_iterations) {
            ++it;
            GL_BREAK(rc, gl_dgemm(ctx, p, m, p, 1.0, w.T, p, 1, w.X, m, 0, 0.0, w.Y, m));      // Y = T^T X
            GL_BREAK(rc, gl_dgemm(ctx, p, m, p, 1.0, w.T, p, 0, w.Y, m, 0, 0.0, w.X, m));      // X = T Y = A^-1 X   (:163-166)
            GL_CUDA_BREAK(rc, cudaMemcpyAsync(w.Xb, w.X, pm, cudaMemcpyDeviceToDevice, ctx->stream));   // CopyVecs (:169)
            if (it % opti_gs == 0) GL_BREAK(rc, ii_orthonormalise(w));                            // (:172-175)
            GL_BREAK(rc, ii_residual(w, &r));
            if (ctx->verbose) fprintf(stderr, "[libglcuda] inverse iteration %d: residual %.6g\n", it, r);
        }
        if (rc != GL_OK) break;
        if (opti_gs != 1 && (it % opti_gs) != 0) GL_BREAK(rc, ii_orthonormalise(w));              // (:183-186)
        {   // a non-positive pivot means A (or a Gram block) was not positive definite
            GL_BREAK(rc, gl_ensure_pinned(ctx, 64));
            GL_CUDA_BREAK(rc, cudaMemcpyAsync(ctx->pinned, w.st, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            GL_CUDA_BREAK(rc, cudaStreamSynchronize(ctx->stream));
            if (*(const int*)ctx->pinned != 0) {
                gl_set_error("inverse iteration: matrix not positive definite (pivot %d)", *(const int*)ctx->pinned - 1);
                rc = GL_ERR_NOTCONVERGED;
                break;
            }
        }
        if (r > epsilon) {
            gl_set_error("inverse iteration: residual %.3g > %.3g after %d iterations", r, epsilon, it);
            rc = GL_ERR_NOTCONVERGED;
            break;
        }
        U = gl_mat_new(ctx, GL_MAT_EIGVEC);
        U->rows = U->local_rows = p;
        U->cols = m;
        U->ld = round_up(p, 64);
        U->elem_bytes = 4;
        GL_BREAK(rc, gl_alloc(ctx, sizeof(float) * (size_t)U->ld * m, &U->buf));
        mu = gl_mat_new(ctx, GL_MAT_DIAG);
        mu->rows = mu->local_rows = m;
        mu->cols = m;
        mu->ld = 1;
        mu->elem_bytes = 8;
        GL_BREAK(rc, gl_alloc(ctx, sizeof(double) * (size_t)m, &mu->buf));
        mui = gl_mat_new(ctx, GL_MAT_DIAG);
        *mui = *mu;
        mui->buf = nullptr;
        GL_BREAK(rc, gl_alloc(ctx, sizeof(double) * (size_t)m, &mui->buf));
        k_ii_extract<<<m, 256, 0, ctx->stream>>>(w.Xb, p, m, (int)U->ld, w.norms, (float*)U->buf->ptr, (double*)mu->buf->ptr,
                                                 (double*)mui->buf->ptr);
        ctx->launches++;
        if (iterations_out) *iterations_out = it;
        if (residual_out) *residual_out = r;
    } while (0);
    for (gl_buf* b : bufs)
        if (b) gl_buf_release(b);
    if (rc != GL_OK) {
        gl_mat_destroy(U);
        gl_mat_destroy(mu);
        gl_mat_destroy(mui);
        return rc;
    }
    if (eigvecs) *eigvecs = U; else gl_mat_destroy(U);
    if (eigvals) *eigvals = mu; else gl_mat_destroy(mu);
    if (eigvals_inv) *eigvals_inv = mui; else gl_mat_destroy(mui);
    return GL_OK;
}
