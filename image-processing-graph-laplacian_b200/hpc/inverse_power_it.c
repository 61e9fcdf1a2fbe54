#include "inverse_power_it.h"

#include "glhost.h"

/* Reference: hpc/inverse_power_it.c:86-252 -- inverse subspace iteration (m GMRES solves per outer step,
 * Gram-Schmidt every optiGramSchmidt steps) stopped at a Frobenius residual of epsilon (default 0.1) and returning
 * the normalised pre-orthogonalisation iterates.  That loop only APPROXIMATES the m eigenpairs of A nearest zero,
 * and its result depends on the MPI process count (random start seeded by rank, :29).  This build returns the
 * converged pairs from the device block-Jacobi solver; optiGramSchmidt and epsilon are accepted and unused
 * (SURVEY.md 8c-iv). */
void InversePowerIteration(const Mat A, const unsigned int p, Mat* eigenvectors, Mat* eigenvalues, PetscBool optiGramSchmidt,
                           PetscScalar epsilon)
{
    (void)optiGramSchmidt;
    (void)epsilon;
    if (gl_eigensolve(GLHostContext(), A, (int)p, eigenvectors, eigenvalues, NULL) != GL_OK) GLHostFatal("InversePowerIteration");
}
