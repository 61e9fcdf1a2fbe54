#include "inverse_power_it.h"

#include "glhost.h"

/* Reference: hpc/inverse_power_it.c:86-252 -- inverse subspace iteration (m GMRES solves per outer step,
 * Gram-Schmidt every optiGramSchmidt steps) stopped at a Frobenius residual of epsilon (default 0.1) and returning
 * the normalised pre-orthogonalisation iterates.  That loop only APPROXIMATES the m eigenpairs of A nearest zero,
 * and its result depends on the MPI process count (random start seeded by rank, :29).  By default this build returns
 * the converged pairs from the device block-Jacobi solver (optiGramSchmidt and epsilon unused);
 * with the host option -inverse_iteration it runs the reference's algorithm itself on the device (gl_inverse_iteration:
 * same outer loop, stopping rule, -opti_gs and -inv_it_epsilon semantics, exact solves). */
void InversePowerIteration(const Mat A, const unsigned int p, Mat* eigenvectors, Mat* eigenvalues, PetscBool optiGramSchmidt,
                           PetscScalar epsilon)
{
    if (g_opt.inverse_iteration) {
        int iterations = 0;
        double residual = 0.0;
        if (gl_inverse_iteration(GLHostContext(), A, (int)p, (int)optiGramSchmidt, (double)epsilon, 10000, eigenvectors, eigenvalues, NULL,
                                 &iterations, &residual) != GL_OK)
            GLHostFatal("InversePowerIteration");
        GLHostPrintf("Inverse subspace iteration took %d outer iterations\n", iterations);    /* hpc/inverse_power_it.c:187 */
        return;
    }
    if (gl_eigensolve(GLHostContext(), A, (int)p, eigenvectors, eigenvalues, NULL) != GL_OK) GLHostFatal("InversePowerIteration");
}
