/* Same entry point as the reference's hpc/inverse_power_it.h:3. */
#ifndef GLB200_INVERSE_POWER_IT_H
#define GLB200_INVERSE_POWER_IT_H
#include "petsc_compat.h"
void InversePowerIteration(const Mat A, const unsigned int p, Mat* eigenvectors, Mat* eigenvalues, PetscBool optiGramSchmidt,
                           PetscScalar epsilon);
#endif
