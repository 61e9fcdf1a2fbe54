#include "nystroem.h"

#include "glhost.h"

/* Reference: hpc/nystroem.c:5-69.  N = pixels, n = samples, p = eigenpairs (the reference's own argument names).
 * The returned matrix already has every pixel's row at its raster position (the GEMM writes it there), so the
 * Permutation that follows in the reference (hpc/utils.c:134-173) is an identity here. */
Mat Nystroem(Mat B, Mat phi_A, Mat Pi_A_Inv, const unsigned int N, const unsigned int n, const unsigned int p)
{
    (void)N; (void)n; (void)p;
    Mat phi = NULL;
    if (gl_nystroem(GLHostContext(), B, phi_A, Pi_A_Inv, &phi) != GL_OK) GLHostFatal("Nystroem");
    return phi;
}
