/* hpc_api.h -- every stage entry point of the reference's hpc/ program, declared once.
 *
 * The reference spreads these prototypes over one small header per source file (hpc/sampling.h, affinity.h,
 * laplacian.h, eigendecomposition.h, inverse_power_it.h, nystroem.h, gram_schmidt.h, display.h, utils.h, read_img.h,
 * write_img.h); those headers still exist here, each one including this file, so code written against the reference
 * compiles unchanged.  Names, argument order and ownership rules are the reference's: out-params
 * and returned Mat / Vec / png_bytep* belong to the caller, who destroys them with MatDestroy / VecDestroy / free.
 * Types come from petsc_compat.h; every function runs on the GPU through include/gl_cuda.h.
 */
#ifndef GLB200_HPC_API_H
#define GLB200_HPC_API_H
#include "petsc_compat.h"

/* ---- sampling.c (reference hpc/sampling.h:1) ---------------------------------------------------------------- */
/* *sample_size in: requested count, out: actual count; *sample_indices: malloc'd ascending raster indices */
void Sampling(const int width, const int height, unsigned int* const sample_size, unsigned int** const sample_indices);

/* ---- affinity.c (reference hpc/affinity.h:5-7) ---------------------------------------------------------------- */
void ComputeAffinityMatrices(Mat* K_A, Mat* K_B, const png_bytep* const img_bytes, const int width, const int height,
                             const unsigned int sample_size, const unsigned int* sample_indices);
void ComputeEntireAffinityMatrix(Mat* K, const png_bytep* const img_bytes, const int width, const int height);

/* ---- laplacian.c (reference hpc/laplacian.h:3-4) -------------------------------------------------------------- */
void ComputeLaplacianMatrix(Mat* L_A, Mat* L_B, Mat K_A, Mat K_B);
void ComputeEntireLaplacianMatrix(Mat* Lapl, Mat K);

/* ---- eigendecomposition.c, inverse_power_it.c (reference hpc/eigendecomposition.h:3-4, hpc/inverse_power_it.h:3) - */
void EigendecompositionLargest(Mat A, const PetscInt num_eigenpairs, Mat* eigenvectors, Mat* eigenvalues, Mat* eigenvalues_inv);
void EigendecompositionSmallest(Mat A, const PetscInt num_eigenpairs, Mat* eigenvectors, Mat* eigenvalues, Mat* eigenvalues_inv);
void InversePowerIteration(const Mat A, const unsigned int p, Mat* eigenvectors, Mat* eigenvalues, PetscBool optiGramSchmidt,
                           PetscScalar epsilon);

/* ---- nystroem.c (reference hpc/nystroem.h:3) ------------------------------------------------------------------- */
Mat Nystroem(Mat B, Mat phi_A, Mat Pi_A_Inv, const unsigned int N, const unsigned int n, const unsigned int p);

/* ---- gram_schmidt.c (reference hpc/gram_schmidt.h:4-5) --------------------------------------------------------- */
void OrthonormaliseVecs(Vec* X, const unsigned int n, const unsigned int p, PetscScalar* norms);
void NormaliseVecs(Vec* X, const unsigned int p, PetscScalar* norms);
/* the same orthonormalisation for the columns of an n x m device matrix Phi, in place (classical Gram-Schmidt == QR with
 * positive diagonal, computed as CholeskyQR on the device with one allreduce over the GPUs) */
void OrthonormaliseMat(Mat phi, PetscScalar* norms);

/* ---- display.c (reference hpc/display.h:8-14) ------------------------------------------------------------------ */
void WriteVec(Vec v, const char* const filename);
void WriteDiagMat(Mat x, const char* const filename);
void WriteMatCol(Mat x, const unsigned int col_num, const char* const filename);
void WritePngMatCol(Mat x, const unsigned int col_num, const unsigned int width, const unsigned int height, const char* const filename);
png_bytep* ComputeResultFromLaplacian(const png_bytep* const img_bytes, Mat phi, Mat Pi, const unsigned int width, const unsigned int height);
png_bytep* ComputeResultFromEntireLaplacian(const png_bytep* const img_bytes, Mat Lapl, const unsigned int width, const unsigned int height);

/* ---- utils.c: the helpers of the reference's hpc/utils.h:7-36 that the approximation path calls; the others are PETSc
 * plumbing (Vecs2Mat, Mat2Vecs, GetFirstCols, pngbytes2OneColMat, AboveXSetY, OneColMat2pngbytes, ...) that the fused
 * device kernels absorb (INTEGRATION.md has the mapping) -------------------------------------------------------------- */
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

/* ---- read_img.c, write_img.c (reference hpc/read_img.h:3, hpc/write_img.h:4): 0 on success, -1 on failure ---------- */
int read_png(const char* const filename, png_bytep** row_pointers, int* const width, int* const height);
int write_png(const char* const filename, png_bytep* img_bytes, const unsigned int width, const unsigned int height);
/* colour plumbing (-color): interleaved RGB rows of 3 * width bytes */
int read_png_rgb(const char* const filename, png_bytep** row_pointers, int* const width, int* const height, int* const file_is_colour);
int write_png_rgb(const char* const filename, png_bytep* img_bytes, const unsigned int width, const unsigned int height);

#endif
