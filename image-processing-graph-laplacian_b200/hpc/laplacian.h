/* Same entry points as the reference's hpc/laplacian.h:3-4. */
#ifndef GLB200_LAPLACIAN_H
#define GLB200_LAPLACIAN_H
#include "petsc_compat.h"
void ComputeLaplacianMatrix(Mat* L_A, Mat* L_B, Mat K_A, Mat K_B);
void ComputeEntireLaplacianMatrix(Mat* Lapl, Mat K);
#endif
