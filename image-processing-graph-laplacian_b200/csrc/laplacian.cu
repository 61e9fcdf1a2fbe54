// a-3: graph-Laplacian normalisation.  Replaces ComputeLaplacianMatrix, hpc/laplacian.c:14-42
// (+ MatRowSum hpc/utils.c:364-376, VecMean :378-388):
//   D_A = rowsum(K_A) + rowsum(K_B)   (carried by the K_B handle, summed in the affinity epilogue)
//   alpha = 1 / mean(D_A);  L_A = alpha (diag D_A - K_A);  L_B = -alpha K_B.
// L_B is NOT a second copy (the reference duplicates K_B, laplacian.c:38-39): the handle shares K_B's
// buffer and carries the factor -alpha, which stays on the device (no host round trip) and is folded
// into the extrapolation operand.
// Also the diagonal helpers InverseDiagMat (hpc/utils.c:559-586) and MatPow (hpc/utils.c:705-729).
#include "common.cuh"

__global__ void k_alpha(const double* __restrict__ D, int p, double* __restrict__ out /* [2]: -alpha, alpha */)
{
    __shared__ double sh[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < p; i += blockDim.x) acc += D[i];   // fixed order per thread
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) {
            double alpha = 1.0 / (v / (double)p);
            out[0] = -alpha;
            out[1] = alpha;
        }
    }
}

// also marks which [64 rows x 16 columns] chunks of L_A hold anything but (numerical) zeros: K_A between samples further
// apart than a few h_loc is < 1e-25, and the eigensolver's fp64 Rayleigh pass skips those chunks (eigen_jacobi.cu)
__global__ void k_laplacian_A(const double* __restrict__ KA, const double* __restrict__ D, const double* __restrict__ al, int p,
                              double* __restrict__ LA, unsigned char* __restrict__ nz, int nz_ld)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= p) return;
    const double alpha = al[1];
    const double v = alpha * ((i == j ? D[i] : 0.0) - KA[(size_t)i * p + j]);
    LA[(size_t)i * p + j] = v;
    if (fabs(v) > 1e-25) nz[(size_t)(i >> 6) * nz_ld + (j >> 4)] = 1;
}

int gl_impl_laplacian(gl_ctx* ctx, gl_mat* K_A, gl_mat* K_B, gl_mat** L_A_out, gl_mat** L_B_out)
{
    const int p = (int)K_A->rows;
    GL_REQUIRE(K_B->p == p && K_B->aux, "laplacian: K_A is %d x %d but K_B has %d samples", p, p, K_B->p);
    gl_mat* LA = gl_mat_new(ctx, GL_MAT_KA);
    LA->rows = LA->cols = LA->local_rows = p;
    LA->ld = p;
    LA->elem_bytes = 8;
    gl_buf* al = nullptr;
    const int nz_ld = (int)ceil_div(p, 16), nz_rows = (int)ceil_div(p, 64);
    int rc = gl_alloc(ctx, sizeof(double) * (size_t)p * p, &LA->buf);
    if (rc == GL_OK) rc = gl_alloc(ctx, 2 * sizeof(double), &al);
    if (rc == GL_OK) rc = gl_alloc(ctx, (size_t)nz_ld * nz_rows, &LA->aux);   // chunk occupancy map of L_A
    if (rc != GL_OK) {
        gl_mat_destroy(LA);
        return rc;
    }
    k_alpha<<<1, 1024, 0, ctx->stream>>>((const double*)K_B->aux->ptr, p, (double*)al->ptr);
    GL_LAUNCH_CHECK(ctx);
    GL_CUDA_CHECK(cudaMemsetAsync(LA->aux->ptr, 0, (size_t)nz_ld * nz_rows, ctx->stream));
    dim3 g((unsigned)ceil_div(p, 128), (unsigned)p);
    k_laplacian_A<<<g, 128, 0, ctx->stream>>>((const double*)K_A->buf->ptr, (const double*)K_B->aux->ptr,
                                             (const double*)al->ptr, p, (double*)LA->buf->ptr, (unsigned char*)LA->aux->ptr, nz_ld);
    GL_LAUNCH_CHECK(ctx);

    gl_mat* LB = gl_mat_new(ctx, GL_MAT_KB);
    *LB = *K_B;              // same shape/bookkeeping ...
    LB->refs = 1;
    if (LB->buf) LB->buf->refs++;         // ... sharing the storage
    if (LB->pt_info) LB->pt_info->refs++;
    if (LB->pt_slots) LB->pt_slots->refs++;
    if (LB->pt_buf) LB->pt_buf->refs++;
    if (LB->aux) LB->aux->refs++;
    if (LB->tiles) LB->tiles->refs++;
    if (LB->starts) LB->starts->refs++;
    if (LB->perm) LB->perm->refs++;
    LB->proj = nullptr;
    LB->dscale = al;         // scale = -alpha, device resident
    LB->scale_on_host = false;
    LB->scale = 0.0;
    *L_A_out = LA;
    *L_B_out = LB;
    return GL_OK;
}

// op 0: 1/x (InverseDiagMat); op 1: x^arg (MatPow as intended)
__global__ void k_diag_map(const double* __restrict__ src, int n, int op, double arg, double* __restrict__ dst)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = src[i];
    dst[i] = op == 0 ? 1.0 / v : (arg == 1.0 ? v : pow(v, arg));
}

int gl_impl_diag_map(gl_ctx* ctx, gl_mat* d, int op, double arg, gl_mat** out)
{
    gl_mat* r = gl_mat_new(ctx, GL_MAT_DIAG);
    r->rows = r->local_rows = d->rows;
    r->cols = d->rows;
    r->ld = 1;
    r->elem_bytes = 8;
    int rc = gl_alloc(ctx, sizeof(double) * (size_t)d->rows, &r->buf);
    if (rc != GL_OK) {
        gl_mat_destroy(r);
        return rc;
    }
    k_diag_map<<<(unsigned)ceil_div(d->rows, 256), 256, 0, ctx->stream>>>((const double*)d->buf->ptr, (int)d->rows, op, arg,
                                                                          (double*)r->buf->ptr);
    GL_LAUNCH_CHECK(ctx);
    *out = r;
    return GL_OK;
}
