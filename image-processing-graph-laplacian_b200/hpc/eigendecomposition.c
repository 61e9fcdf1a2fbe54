#include "eigendecomposition.h"

#include "glhost.h"

/* Reference: hpc/eigendecomposition.c:12-124 (SLEPc Krylov-Schur, EPS_HEP).  Both directions run the device
 * block-Jacobi solver, which computes the whole spectrum and keeps the requested end of it. */
static void solve(Mat A, PetscInt n, Mat* vecs, Mat* vals, Mat* vals_inv, int largest)
{
    gl_ctx* ctx = GLHostContext();
    char v[8];
    v[0] = largest ? '1' : '0';
    v[1] = 0;
    gl_ctx_set_option(ctx, "eig_largest", v);
    int rc = gl_eigensolve(ctx, A, n, vecs, vals, vals_inv);
    gl_ctx_set_option(ctx, "eig_largest", "0");
    if (rc != GL_OK) GLHostFatal("Eigendecomposition");
}

void EigendecompositionLargest(Mat A, const PetscInt num_eigenpairs, Mat* eigenvectors, Mat* eigenvalues, Mat* eigenvalues_inv)
{
    solve(A, num_eigenpairs, eigenvectors, eigenvalues, eigenvalues_inv, 1);
}

void EigendecompositionSmallest(Mat A, const PetscInt num_eigenpairs, Mat* eigenvectors, Mat* eigenvalues, Mat* eigenvalues_inv)
{
    solve(A, num_eigenpairs, eigenvectors, eigenvalues, eigenvalues_inv, 0);
}
