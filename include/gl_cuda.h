/*
 * gl_cuda.h -- C ABI of libglcuda.so: the B200 (sm_100a) implementation of the
 * Nystroem graph-Laplacian image-filter path.
 *
 * This is the thin layer the C host (hpc/ in this repo, a drop-in for the
 * reference's hpc/image_processing.c) calls instead of PETSc/SLEPc/MPI.  Plain C
 * types and opaque handles only; every entry point returns an int status
 * (GL_OK == 0) and never throws.  There is no CPU fallback: every compute entry
 * fails with GL_ERR_CUDA when no sm_100 device is usable.
 *
 * Reference interfaces replaced (file:line in David-Wobrock/
 * image-processing-graph-laplacian):
 *   gl_sampling_*      <- Sampling/UniformSampling           hpc/sampling.c:6-33, hpc/sampling.h:1
 *                         random_sample                       python/sampling/random.py:8-16
 *   gl_affinity        <- ComputeAffinityMatrices            hpc/affinity.c:129-262, hpc/affinity.h:5
 *   gl_laplacian       <- ComputeLaplacianMatrix             hpc/laplacian.c:14-42, hpc/laplacian.h:3
 *   gl_eigensolve      <- EigendecompositionSmallest         hpc/eigendecomposition.c:121-124
 *                         InversePowerIteration               hpc/inverse_power_it.c:86-252
 *   gl_nystroem        <- Nystroem + Permutation             hpc/nystroem.c:5-69, hpc/utils.c:134-173
 *   gl_orthonormalise  <- OrthonormaliseVecs                 hpc/gram_schmidt.c:29-64
 *   gl_filter          <- ComputeResultFromLaplacian         hpc/display.c:58-83
 *   gl_mat_*           <- Mat/Vec objects, MatDestroy        (PETSc, not in tree)
 *   gl_comm_*          <- MPI_Comm_rank/size, PETSc allreduces hpc/image_processing.c:30-38
 *
 * Threading: a gl_ctx is used from one host thread at a time (the reference is
 * single-threaded SPMD).  Multi-GPU is SPMD too: one process (or thread) per
 * GPU, each with its own gl_ctx of the same `world`, calling every stage in the
 * same order; the stages allreduce internally over NCCL.
 */
#ifndef GL_CUDA_H
#define GL_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GL_API __attribute__((visibility("default")))

typedef struct gl_ctx gl_ctx; /* device, stream, workspace arena, NCCL communicator, current image + samples */
typedef struct gl_mat gl_mat; /* opaque device matrix / vector (refcounted buffer + shape + dtype) */

enum gl_status {
    GL_OK = 0,
    GL_ERR_CUDA = 1,        /* CUDA runtime/driver error, or no sm_100 device */
    GL_ERR_ARG = 2,         /* bad argument / wrong handle kind / stage order */
    GL_ERR_NOMEM = 3,       /* device allocation failed */
    GL_ERR_NCCL = 4,        /* NCCL missing or failed (only when world > 1) */
    GL_ERR_UNSUPPORTED = 5, /* valid request outside what this build handles */
    GL_ERR_NOTCONVERGED = 6 /* eigensolver hit its sweep limit */
};

enum gl_affinity_kind { /* hpc/affinity.c:115-122 picks bilateral; the two others are commented there */
    GL_BILATERAL = 0,   /* exp(-|dpos|^2/h_loc^2) * exp(-|dval|^2/h_val^2) */
    GL_PHOTOMETRIC = 1, /* exp(-|dval|^2/h_val^2) */
    GL_SPATIAL = 2,     /* exp(-|dpos|^2/h_loc^2) */
    GL_NLM = 3          /* non-local means, python/affinity_methods/NLM.py: exp(-|G o (patch_a - patch_b)|^2 / h_val^2) over 7x7 patches
                           of the symmetrically padded image, G = normalised Gaussian weights (sigma 1.2); grey only; the
                           reference's h is 3 */
};

enum gl_mat_kind {
    GL_MAT_KA = 1,      /* p x p fp64 row-major (K_A or L_A) */
    GL_MAT_KB = 2,      /* this rank's pixel band x p fp16: K_B stored pixel-major (transposed) in blocks of
                           [512 pixels][64 samples], all band pixels incl. samples; sample blocks whose entries
                           fp16 flushes to zero (spatial distance) are not stored; carries the fp64 row sums
                           D = K_A.1 + K_B.1 */
    GL_MAT_EIGVEC = 3,  /* p x m fp32 column-major (eigenvectors of L_A in columns) */
    GL_MAT_DIAG = 4,    /* m-vector fp64 standing for a diagonal matrix */
    GL_MAT_PHI = 5,     /* this rank's pixel band x m_pad fp16 row-major, rows in raster order */
    GL_MAT_FULL = 6     /* -no_approx: the n x n K (or L = alpha (D - K)) held matrix-free as the per-pixel vectors
                           D = K.1 and Q = sum_j K_ij (y_i - y_j) of this rank's band (+ alpha for L) */
};

typedef struct gl_mat_info {
    int kind;            /* gl_mat_kind */
    int64_t rows, cols;  /* logical (global) shape as the reference would see it */
    int64_t local_rows;  /* rows held by this rank (pixel band) for KB/PHI, else rows */
    int64_t ld;          /* leading dimension in elements of the stored layout */
    int elem_bytes;
    double scale;        /* logical value = scale * stored value (L_B = -alpha K_B shares K_B's buffer) */
    int64_t stored_blocks; /* KB, blocked layout: [512 x ld] blocks held (ld = 64, or 32 with option kb_block; dense: ceil(local_rows/512) * ceil(p/ld)); else 0 */
    int layout;            /* KB: 0 = blocked storage (csrc/affinity.cu), 1 = patch layout (csrc/patch.cu: gathered sample lists per 64 x 16 pixel patch) */
    int64_t stored_pairs;  /* KB: (pixel, sample slot) pairs held, padding included */
    int64_t mma_pairs;     /* KB: (pixel, sample slot) pairs the extrapolation multiplies (x 2 m_pad flop each) */
} gl_mat_info;

/* Stage indices for gl_ctx_stage_ms (same vocabulary as the reference's stdout timers,
 * hpc/image_processing.c:198-233,249,268). */
enum gl_stage {
    GL_T_H2D = 0, GL_T_SAMPLING, GL_T_AFFINITY, GL_T_LAPLACIAN, GL_T_EIGEN, GL_T_NYSTROEM,
    GL_T_GRAM_SCHMIDT, GL_T_FILTER, GL_T_D2H, GL_T_TOTAL,
    /* single-kernel timers (roofline numerators): the K_B affinity kernel, the extrapolation GEMM kernel,
     * the two filter passes */
    GL_T_K_AFFINITY_B, GL_T_K_GEMM, GL_T_K_FILTER_PROJECT, GL_T_K_FILTER_APPLY, GL_T_K_JACOBI,
    GL_T_COUNT
};

typedef struct gl_params {
    int affinity_kind;      /* gl_affinity_kind; default GL_BILATERAL */
    double h_loc, h_val;    /* hpc/affinity.c:117-118: 40, 30 */
    int sampling_random;    /* 0 = spatially uniform grid (hpc/sampling.c), 1 = random (python/sampling/random.py) */
    uint32_t seed;          /* random sampling seed (np.random.seed) */
    unsigned sample_size;   /* requested p; 0 => 1 % of the pixels (hpc/image_processing.c:187) */
    int num_eigvals;        /* m; <0 or >= p => p-1 (hpc/image_processing.c:96-108) */
    double gain;            /* hpc/display.c:73: 3.0 */
    double power;           /* f(lambda) = lambda^power; MatPow is a no-op in the reference (hpc/utils.c:721) => 1 */
    int gram_schmidt;       /* orthonormalise Phi before filtering (off in the restored reference block) */
    int clip_low;           /* 0: only z>255 is clipped (hpc/display.c:76); 1: also clamp z<0 to 0 */
} gl_params;

/* ---- library / errors ------------------------------------------------------------------ */
GL_API int gl_version(void);
GL_API const char* gl_last_error(void);             /* thread-local text of the last failure */
GL_API void gl_default_params(gl_params* p);
GL_API int gl_device_count(int* count);             /* sm_100 devices visible */
/* Device memory of this context's allocator in bytes: held by live handles and workspaces, cached for reuse, and the peak
 * of `live` since the context was made (or since the last call with reset_peak != 0).  Any pointer may be NULL. */
GL_API int gl_memory_stats(gl_ctx* ctx, size_t* live, size_t* cached, size_t* peak, int reset_peak);
GL_API int gl_kernel_launches(gl_ctx* ctx, long long* count); /* kernels of this library launched on ctx so far */

/* ---- context ----------------------------------------------------------------------------- */
GL_API int gl_ctx_create(gl_ctx** ctx, int device, int rank, int world);
GL_API int gl_ctx_destroy(gl_ctx* ctx);
GL_API int gl_ctx_sync(gl_ctx* ctx);
GL_API int gl_ctx_stage_ms(gl_ctx* ctx, float* ms /* [GL_T_COUNT] */); /* CUDA-event times of the last run of each stage */
GL_API int gl_ctx_set_option(gl_ctx* ctx, const char* key, const char* value); /* tuning knobs, see DESIGN.md */
/* CUDA-event marks on the context stream (slot 0..127) and the device time between two of them. */
GL_API int gl_ctx_mark(gl_ctx* ctx, int slot);
GL_API int gl_ctx_mark_elapsed_ms(gl_ctx* ctx, int slot_a, int slot_b, float* ms);
/* Benchmarks: write `bytes` of scratch on the context stream so that nothing of the previous iteration stays in the L2
 * (0 = twice the device's L2 size). */
GL_API int gl_ctx_flush_l2(gl_ctx* ctx, size_t bytes);
/* NCCL bootstrap (world > 1): rank 0 makes a 128-byte id, the host shares it, every rank joins. */
GL_API int gl_comm_unique_id(void* id128);
GL_API int gl_comm_init(gl_ctx* ctx, const void* id128);

/* ---- image ------------------------------------------------------------------------------- */
/* Whole image on every rank (the reference broadcasts it, hpc/image_processing.c:45-76);
 * interleaved u8, `channels` in {1,3}.  Copies H2D on the context stream. */
GL_API int gl_set_image(gl_ctx* ctx, const uint8_t* pixels, int width, int height, int channels);
/* Same from the reference's png_bytep* row pointers (grey). */
GL_API int gl_set_image_rows(gl_ctx* ctx, const uint8_t* const* rows, int width, int height);
/* Deterministic synthetic image generated on device (bench inputs; oracle_np.synthetic_image). */
GL_API int gl_set_synthetic_image(gl_ctx* ctx, int width, int height, int channels, uint32_t seed);
GL_API int gl_get_image(gl_ctx* ctx, uint8_t* pixels_out);
/* Pixel band [row0,row1) of image rows this rank owns. */
GL_API int gl_get_band(gl_ctx* ctx, int* row0, int* row1);

/* ---- a-1 sampling (on device, bit-exact) ---------------------------------------------------- */
GL_API int gl_sampling_uniform(gl_ctx* ctx, unsigned requested, unsigned* actual);
GL_API int gl_sampling_random(gl_ctx* ctx, unsigned requested, uint32_t seed, unsigned* actual);
GL_API int gl_set_samples(gl_ctx* ctx, const uint32_t* indices, unsigned count); /* ascending raster indices */
GL_API int gl_get_samples(gl_ctx* ctx, uint32_t* indices_out, unsigned cap, unsigned* count);

/* ---- a-2 .. a-9 stages ------------------------------------------------------------------------ */
GL_API int gl_affinity(gl_ctx* ctx, int kind, double h_loc, double h_val, gl_mat** K_A, gl_mat** K_B);
GL_API int gl_laplacian(gl_ctx* ctx, gl_mat* K_A, gl_mat* K_B, gl_mat** L_A, gl_mat** L_B);
/* m smallest eigenpairs of symmetric positive definite L_A, ascending, 1 <= m <= p (m < 0 or m > p: p - 1).
 * eigvecs and/or eigvals_inv may be NULL. */
GL_API int gl_eigensolve(gl_ctx* ctx, gl_mat* L_A, int m, gl_mat** eigvecs, gl_mat** eigvals, gl_mat** eigvals_inv);
/* The reference's own eigensolver, hpc/inverse_power_it.c:86-252 (InversePowerIteration), with its options -opti_gs and
 * -inv_it_epsilon (hpc/image_processing.c:124-152): inverse subspace iteration on m vectors, orthonormalised every opti_gs-th
 * step, until |(I - X X^T) A X|_F <= epsilon; lambda_i = 1 / (norm of iterate i before normalisation), eigenvectors = the
 * normalised iterates, in the order the iteration leaves them (no sort, as in the reference).  The linear solves are exact
 * (Cholesky of the SPD L_A) where the reference runs GMRES; the start is a fixed pseudo-random matrix where the reference seeds
 * PETSc's generator with the MPI rank.  m << p is what this solver is for (-num_eigvals); gl_eigensolve stays the default (all
 * pairs, converged).  GL_ERR_NOTCONVERGED after max_iterations outer steps.  iterations_out / residual_out may be NULL. */
GL_API int gl_inverse_iteration(gl_ctx* ctx, gl_mat* L_A, int m, int opti_gs, double epsilon, int max_iterations, gl_mat** eigvecs,
                                gl_mat** eigvals, gl_mat** eigvals_inv, int* iterations_out, double* residual_out);
/* Phi (n x m): sample rows = phi_A, other rows = L_B^T . phi_A . diag(eigvals_inv), already in raster order.
 * The handle is DEFERRED (option lazy_phi=0: computed at once): it retains its three inputs and the matrix is computed by
 * its first consumer -- gl_filter runs extrapolation and filter as one pass and stores nothing (option keep_phi=1: stores
 * Phi in that pass), gl_mat_download / gl_mat_download_cols / gl_orthonormalise run the plain GEMM and keep the matrix. */
GL_API int gl_nystroem(gl_ctx* ctx, gl_mat* L_B, gl_mat* phi_A, gl_mat* eigvals_inv, gl_mat** phi);
/* gl_nystroem and gl_filter in ONE pass over Phi: the filter weights gain * f o (Phi^T y) are formed first (Phi^T y
 * from the affinity stage's sums) and the extrapolation GEMM accumulates each row's product with the weights in its
 * epilogue, so Phi is never read back.  Needs L_B of the CURRENT image; same outputs as the two calls.  `phi` != NULL:
 * Phi is written as well and returned; NULL: it is not written to memory at all (its tiles only ever exist in tensor memory). */
GL_API int gl_nystroem_filter(gl_ctx* ctx, gl_mat* L_B, gl_mat* phi_A, gl_mat* eigvals_inv, gl_mat* f_eigvals, double gain,
                              int clip_low, gl_mat** phi, float* z_f32, uint8_t* z_u8);
GL_API int gl_orthonormalise(gl_ctx* ctx, gl_mat* phi, double* norms_out /* m, may be NULL */);
/* z = y + gain * Phi (f o (Phi^T y)), z[z>255]=255.  z_f32 (n*C floats) and/or z_u8 (n*C bytes) receive this
 * rank's band at its raster offset (other entries untouched); either may be NULL. */
GL_API int gl_filter(gl_ctx* ctx, gl_mat* phi, gl_mat* f_eigvals, double gain, int clip_low, float* z_f32, uint8_t* z_u8);
/* diag helpers: InverseDiagMat (hpc/utils.c:559-586) and MatPow (hpc/utils.c:705-729, as intended: x^power) */
GL_API int gl_diag_inverse(gl_ctx* ctx, gl_mat* d, gl_mat** out);
GL_API int gl_diag_pow(gl_ctx* ctx, gl_mat* d, double power, gl_mat** out);

/* ---- the experimental blocks of the reference's Python prototype (csrc/proto.cu) ---------------------------------
 * They work on a STORED Phi in raster order (from gl_nystroem -- also on a bare K_B with the eigenpairs of K_A, which is the
 * prototype's nystroem(K_A, K_B), python/image_processing.py:69-88 -- or uploaded with gl_mat_upload(GL_MAT_PHI)) and on the
 * context's current samples; the prototype's "sample rows first" order and its permutation() do not exist here.
 *   gl_sinkhorn          <- sinkhorn(phi, Pi)                python/image_processing.py:90-107
 *   gl_orthogonalisation <- orthogonalisation(A, B)          python/image_processing.py:110-127
 *   gl_smoothing_matrix  <- smoothing_matrix(s, phi, Pi)     python/image_processing.py:151-194
 *   gl_matrix_filter     <- smoothing / sharpening           python/image_processing.py:197-241 */
/* `iterations` (the prototype: 100) alternating scalings r, c of K = Phi diag(Pi) Phi^T from r = 1, then the sample rows of
 * diag(r) K diag(c): W_A (p x p, GL_MAT_KA) and, for every pixel j of this rank's band, W_ABt[j][i] = (diag(r) K diag(c))[s_i][j]
 * (Phi-shaped, band pixels x p; the prototype's W_B is its transpose without the sample pixels' rows).  Either may be NULL. */
GL_API int gl_sinkhorn(gl_ctx* ctx, gl_mat* phi, gl_mat* Pi, int iterations, gl_mat** W_A, gl_mat** W_ABt);
/* One-shot orthogonal Nystroem extension of the affinity blocks gl_affinity returned for the current image and samples:
 * V (Phi-shaped, band pixels x p, orthonormal columns over the whole image, raster order) and Pi = min(eigenvalues of
 * A + A^-1/2 B B^T A^-1/2, 1), descending.  Bilateral / photometric / spatial affinity; K_A must be positive definite. */
GL_API int gl_orthogonalisation(gl_ctx* ctx, gl_mat* K_A, gl_mat* K_B, gl_mat** V, gl_mat** Pi);
/* From the approximation K = Phi diag(Pi) Phi^T: D = K 1, alpha = 1 / mean(D), W = I + alpha (K - diag D); eigenpairs (L, descending)
 * of its sample block and their Nystroem extension V (Phi-shaped, band pixels x p, sample rows = the eigenvectors). */
GL_API int gl_smoothing_matrix(gl_ctx* ctx, gl_mat* phi, gl_mat* Pi, gl_mat** V, gl_mat** L);
/* z = sum_k coef[k] W^k y for W = V diag(L) V^T and the current image y (every channel), not clipped: smoothing() is
 * coef = {0, 1}, sharpening() is {0, 0, 1 + beta, -beta} with beta = 1.5.  z_f32: n * C floats, this rank's band at its raster offset. */
GL_API int gl_matrix_filter(gl_ctx* ctx, gl_mat* V, gl_mat* L, const double* coef, int ncoef, float* z_f32);

/* ---- the reference's -no_approx mode, matrix-free (csrc/full_filter.cu) ----------------------------------
 *   gl_full_affinity  <- ComputeEntireAffinityMatrix       hpc/affinity.c:264-336
 *   gl_full_laplacian <- ComputeEntireLaplacianMatrix      hpc/laplacian.c:44-65
 *   gl_full_result    <- ComputeResultFromEntireLaplacian  hpc/display.c:128-149:  z = clip(y - L y, 0, 255) */
GL_API int gl_full_affinity(gl_ctx* ctx, int kind, double h_loc, double h_val, gl_mat** K);
GL_API int gl_full_laplacian(gl_ctx* ctx, gl_mat* K, gl_mat** L);
GL_API int gl_full_result(gl_ctx* ctx, gl_mat* L, float* z_f32, uint8_t* z_u8);

/* ---- whole path in one call (what hpc/image_processing.c:183-277 sequences) -------------------
 * Phi is a temporary of this call: without -gram_schmidt it is consumed tile by tile by the fused filter and not stored
 * (option keep_phi=1 stores it; z is the same bit for bit). */
/* (world > 1: every rank passes the WHOLE host image; a rank copies only its band of rows and the sampled pixels to its GPU.) */
GL_API int gl_run(gl_ctx* ctx, const uint8_t* pixels, int width, int height, int channels, const gl_params* prm,
                  float* z_f32, uint8_t* z_u8, unsigned* p_out, int* m_out, double* eigvals_out /* may be NULL */,
                  size_t eigvals_cap /* doubles eigvals_out can hold; GL_ERR_ARG if the m eigenvalues do not fit */);
/* Same, image already on the device (gl_set_image / gl_set_synthetic_image); no host copies unless z_* given. */
GL_API int gl_run_resident(gl_ctx* ctx, const gl_params* prm, float* z_f32, uint8_t* z_u8, unsigned* p_out, int* m_out,
                           double* eigvals_out, size_t eigvals_cap);

/* ---- matrices ---------------------------------------------------------------------------------- */
GL_API int gl_mat_info_get(const gl_mat* m, gl_mat_info* info);
GL_API int gl_mat_retain(gl_mat* m);
GL_API int gl_mat_destroy(gl_mat* m);  /* drop one reference (MatDestroy) */
/* Download as fp64, logical values (scale applied), row-major rows x cols; for KB/PHI: this rank's band
 * rows x logical cols.  `cap` = number of doubles `out` can hold. */
GL_API int gl_mat_download(gl_ctx* ctx, const gl_mat* m, double* out, size_t cap);
/* A strip of columns [col0, col0 + ncols) of a Phi (this rank's band rows), eigenvector or p x p matrix, as fp64 row-major
 * rows x ncols (replaces MatGetColumnVector / GetFirstCols+GetLastCols in WriteMatCol / WritePngMatCol, hpc/display.c:85-126);
 * a 4K Phi is 17 GB on the device and 66 GB as fp64, so the eigenvector dumps must not go through gl_mat_download. */
GL_API int gl_mat_download_cols(gl_ctx* ctx, const gl_mat* m, int col0, int ncols, double* out, size_t cap);
/* D = rowsum(K_A)+rowsum(K_B) carried by a KB handle (p doubles), already summed over ranks. */
GL_API int gl_mat_rowsums(gl_ctx* ctx, const gl_mat* K_B, double* out, size_t cap);
/* Upload a host fp64 row-major matrix as GL_MAT_KA / GL_MAT_EIGVEC / GL_MAT_DIAG, or as GL_MAT_PHI (this rank's band rows x cols, raster
 * order; needs the image geometry and a sample set) (tests, host-built inputs). */
GL_API int gl_mat_upload(gl_ctx* ctx, int kind, const double* data, int64_t rows, int64_t cols, gl_mat** out);

/* ---- inspection (host only, no GPU needed) ------------------------------------------------------------------------
 * The storage layout gl_affinity gives K_B for these (strictly ascending) samples over the raster range [q0, q1): per
 * 512-pixel tile `tile_count` blocks of `block_slots` sample slots whose first slots are starts[tile_first ...]; perm[slot] = index
 * of the sample in `samples` (0xffffffff: empty), p_pad + 64 slots with p_pad = p rounded up to 64.  Call once with the
 * arrays NULL to learn n_tiles / n_blocks.  strips = 0 lets the library choose the number of column strips. */
GL_API int gl_kb_layout_host(int width, int64_t q0, int64_t q1, const uint32_t* samples, unsigned p, double h_loc, int cutoff,
                             int strips, int block_slots /* 64 or 32 */, int* strips_out, int64_t* n_tiles, int64_t* n_blocks, int32_t* tile_first,
                             int32_t* tile_count, int32_t* starts, uint32_t* perm);

/* ---- pinned host memory for callers that want async copies ------------------------------------ */
GL_API int gl_host_alloc(void** p, size_t bytes);
GL_API int gl_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* GL_CUDA_H */
