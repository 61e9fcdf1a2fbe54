/* The helpers of the reference's hpc/utils.h:7-36 that the approximation path calls.  The others are PETSc
 * plumbing (Vecs2Mat, Mat2Vecs, GetFirstCols, pngbytes2OneColMat, AboveXSetY, OneColMat2pngbytes, ...) that the
 * fused device kernels absorb; see INTEGRATION.md for the mapping. */
#ifndef GLB200_UTILS_H
#define GLB200_UTILS_H
#include "petsc_compat.h"

extern const PetscInt ZERO;

unsigned int num2x(const unsigned int num, const unsigned int num_col);
unsigned int num2y(const unsigned int num, const unsigned int num_col);
unsigned int xy2num(const unsigned int x, const unsigned y, const unsigned int num_col);

Mat Permutation(Mat m, const unsigned int* const sample_indices, const unsigned int num_sample_indices);
Vec MatRowSum(Mat A);
PetscScalar VecMean(Vec x);
Mat InverseDiagMat(Mat x);
Mat MatPow(Mat A, PetscScalar x);
Vec DiagMat2Vec(Mat x);
#endif
