#include "laplacian.h"

#include <stdio.h>

#include "glhost.h"

/* Reference: hpc/laplacian.c:14-42.  L_B shares K_B's storage with the factor -alpha attached. */
void ComputeLaplacianMatrix(Mat* L_A, Mat* L_B, Mat K_A, Mat K_B)
{
    if (gl_laplacian(GLHostContext(), K_A, K_B, L_A, L_B) != GL_OK) GLHostFatal("ComputeLaplacianMatrix");
}

/* Reference: hpc/laplacian.c:44-65: alpha = 1 / mean(rowsum K), L = alpha (D - K); matrix-free here. */
void ComputeEntireLaplacianMatrix(Mat* Lapl, Mat K)
{
    if (gl_full_laplacian(GLHostContext(), K, Lapl) != GL_OK) GLHostFatal("ComputeEntireLaplacianMatrix");
}
