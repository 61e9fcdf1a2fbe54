#include "laplacian.h"

#include <stdio.h>

#include "glhost.h"

/* Reference: hpc/laplacian.c:14-42.  L_B shares K_B's storage with the factor -alpha attached. */
void ComputeLaplacianMatrix(Mat* L_A, Mat* L_B, Mat K_A, Mat K_B)
{
    if (gl_laplacian(GLHostContext(), K_A, K_B, L_A, L_B) != GL_OK) GLHostFatal("ComputeLaplacianMatrix");
}

void ComputeEntireLaplacianMatrix(Mat* Lapl, Mat K)
{
    (void)K;
    *Lapl = NULL;
    fprintf(stderr, "ComputeEntireLaplacianMatrix: the -no_approx path is not part of this build\n");
}
