/* Same entry point as the reference's hpc/nystroem.h:3. */
#ifndef GLB200_NYSTROEM_H
#define GLB200_NYSTROEM_H
#include "petsc_compat.h"
Mat Nystroem(Mat B, Mat phi_A, Mat Pi_A_Inv, const unsigned int N, const unsigned int n, const unsigned int p);
#endif
