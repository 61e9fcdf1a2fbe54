/* Same entry points as the reference's hpc/eigendecomposition.h:3-4. */
#ifndef GLB200_EIGENDECOMPOSITION_H
#define GLB200_EIGENDECOMPOSITION_H
#include "petsc_compat.h"
void EigendecompositionLargest(Mat A, const PetscInt num_eigenpairs, Mat* eigenvectors, Mat* eigenvalues, Mat* eigenvalues_inv);
void EigendecompositionSmallest(Mat A, const PetscInt num_eigenpairs, Mat* eigenvectors, Mat* eigenvalues, Mat* eigenvalues_inv);
#endif
