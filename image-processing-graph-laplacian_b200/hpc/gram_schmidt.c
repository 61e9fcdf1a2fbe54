#include "gram_schmidt.h"

#include <math.h>
#include <stdlib.h>

#include "glhost.h"

/* Reference: hpc/gram_schmidt.c:29-64.  X is an array of p host vectors of length n (the reference uses it on the
 * p-length iterates of its inverse iteration only).  u_k = v_k - sum_{j<k} <v_k,u_j>/<u_j,u_j> u_j, then
 * normalise; norms[k] = |u_k| before normalising (gram_schmidt.c:55-59). */
void OrthonormaliseVecs(Vec* X, const unsigned int n, const unsigned int p, PetscScalar* norms)
{
    double* sum = (double*)malloc(sizeof(double) * n);
    for (unsigned int k = 0; k < p; ++k) {
        for (unsigned int i = 0; i < n; ++i) sum[i] = 0.0;
        for (unsigned int j = 0; j < k; ++j) {
            double vu = 0.0, uu = 0.0;
            for (unsigned int i = 0; i < n; ++i) {
                vu += X[k]->data[i] * X[j]->data[i];
                uu += X[j]->data[i] * X[j]->data[i];
            }
            const double f = vu / uu;
            for (unsigned int i = 0; i < n; ++i) sum[i] += f * X[j]->data[i];
        }
        double nrm = 0.0;
        for (unsigned int i = 0; i < n; ++i) {
            X[k]->data[i] -= sum[i];
            nrm += X[k]->data[i] * X[k]->data[i];
        }
        nrm = sqrt(nrm);
        if (norms) norms[k] = nrm;
        for (unsigned int i = 0; i < n; ++i) X[k]->data[i] /= nrm;
    }
    free(sum);
}

/* Reference: hpc/gram_schmidt.c:66-77. */
void NormaliseVecs(Vec* X, const unsigned int p, PetscScalar* norms)
{
    for (unsigned int k = 0; k < p; ++k) {
        double nrm = 0.0;
        for (unsigned int i = 0; i < X[k]->n; ++i) nrm += X[k]->data[i] * X[k]->data[i];
        nrm = sqrt(nrm);
        if (norms) norms[k] = nrm;
        for (unsigned int i = 0; i < X[k]->n; ++i) X[k]->data[i] /= nrm;
    }
}

void OrthonormaliseMat(Mat phi, PetscScalar* norms)
{
    if (gl_orthonormalise(GLHostContext(), phi, norms) != GL_OK) GLHostFatal("OrthonormaliseMat");
}
