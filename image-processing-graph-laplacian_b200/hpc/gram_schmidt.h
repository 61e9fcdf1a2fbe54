/* Same entry points as the reference's hpc/gram_schmidt.h:4-5, plus the Phi-sized variant the pipeline uses. */
#ifndef GLB200_GRAM_SCHMIDT_H
#define GLB200_GRAM_SCHMIDT_H
#include "petsc_compat.h"
void OrthonormaliseVecs(Vec* X, const unsigned int n, const unsigned int p, PetscScalar* norms);
void NormaliseVecs(Vec* X, const unsigned int p, PetscScalar* norms);
/* Orthonormalise the columns of an n x m device matrix Phi in place (same arithmetic: classical Gram-Schmidt ==
 * QR with positive diagonal; computed as CholeskyQR on the device, one allreduce over the GPUs). */
void OrthonormaliseMat(Mat phi, PetscScalar* norms);
#endif
