#include "utils.h"

#include <stdio.h>
#include <stdlib.h>

#include "glhost.h"

const PetscInt ZERO = 0;

/* hpc/utils.c:11-24: x is the ROW, y the COLUMN of raster index num. */
unsigned int num2x(const unsigned int num, const unsigned int num_col) { return num / num_col; }
unsigned int num2y(const unsigned int num, const unsigned int num_col) { return num % num_col; }
unsigned int xy2num(const unsigned int x, const unsigned int y, const unsigned int num_col) { return x * num_col + y; }

/* hpc/utils.c:134-173 moves the first rows to the sample positions and the others to the remaining raster
 * positions.  Nystroem() here writes rows at their raster position to begin with, so this returns the same matrix
 * (one more reference; the caller destroys both handles as the reference does, image_processing.c:252-254). */
Mat Permutation(Mat m, const unsigned int* const sample_indices, const unsigned int num_sample_indices)
{
    (void)sample_indices;
    (void)num_sample_indices;
    gl_mat_retain(m);
    return m;
}

/* hpc/utils.c:364-376.  For the device-resident K_B the sums were taken in the affinity kernel's epilogue
 * (and already include K_A's, hpc/laplacian.c:18-20); for a p x p matrix they are formed from a download. */
Vec MatRowSum(Mat A)
{
    gl_mat_info info;
    if (gl_mat_info_get(A, &info) != GL_OK) GLHostFatal("MatRowSum");
    Vec v = VecCreateHost((unsigned int)info.rows);
    if (info.kind == GL_MAT_KB) {
        if (gl_mat_rowsums(GLHostContext(), A, v->data, v->n) != GL_OK) GLHostFatal("MatRowSum");
    } else {
        const size_t cnt = (size_t)info.rows * (size_t)info.cols;
        double* tmp = (double*)malloc(sizeof(double) * cnt);
        if (gl_mat_download(GLHostContext(), A, tmp, cnt) != GL_OK) GLHostFatal("MatRowSum");
        for (int64_t i = 0; i < info.rows; ++i) {
            double s = 0.0;
            for (int64_t j = 0; j < info.cols; ++j) s += tmp[i * info.cols + j];
            v->data[i] = s;
        }
        free(tmp);
    }
    return v;
}

/* hpc/utils.c:378-388 */
PetscScalar VecMean(Vec x)
{
    double s = 0.0;
    for (unsigned int i = 0; i < x->n; ++i) s += x->data[i];
    return s / x->n;
}

/* hpc/utils.c:559-586 */
Mat InverseDiagMat(Mat x)
{
    Mat y = NULL;
    if (gl_diag_inverse(GLHostContext(), x, &y) != GL_OK) GLHostFatal("InverseDiagMat");
    return y;
}

/* hpc/utils.c:705-729.  The reference computes pow(value, x) and DISCARDS it (:721), so the returned matrix equals
 * A whatever x is; that behaviour is kept (the filter is then f(lambda) = lambda).  `-filter_pow P` makes MatPow
 * really apply the exponent P (the evidently intended behaviour, image_processing.c:263 asks for 6). */
Mat MatPow(Mat A, PetscScalar x)
{
    (void)x;
    Mat B = NULL;
    const double e = g_opt.filter_pow_set ? g_opt.filter_pow : 1.0;
    if (gl_diag_pow(GLHostContext(), A, e, &B) != GL_OK) GLHostFatal("MatPow");
    return B;
}

/* hpc/utils.c:536-549 (diagonal of a diagonal matrix as a vector) */
Vec DiagMat2Vec(Mat x)
{
    gl_mat_info info;
    if (gl_mat_info_get(x, &info) != GL_OK) GLHostFatal("DiagMat2Vec");
    Vec v = VecCreateHost((unsigned int)info.rows);
    if (gl_mat_download(GLHostContext(), x, v->data, v->n) != GL_OK) GLHostFatal("DiagMat2Vec");
    return v;
}
