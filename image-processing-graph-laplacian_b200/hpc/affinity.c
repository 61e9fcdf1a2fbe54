#include "affinity.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "glhost.h"

/* Reference: hpc/affinity.c:129-262.  K_A is p x p; K_B is held pixel-major on the device for this rank's band
 * of image rows (see include/gl_cuda.h GL_MAT_KB).  The kernel and its bandwidths are the reference's compile-time
 * choices (hpc/affinity.c:117-121) unless overridden with -affinity / -h_loc / -h_val. */
void ComputeAffinityMatrices(Mat* K_A, Mat* K_B, const png_bytep* const img_bytes, const int width, const int height,
                             const unsigned int sample_size, const unsigned int* sample_indices)
{
    gl_ctx* ctx = GLHostContext();
    if (g_opt.color) {
        /* -color: rows are interleaved RGB (3 * width bytes); the photometric term runs over the three channels */
        uint8_t* flat = (uint8_t*)malloc((size_t)width * height * 3);
        for (int r = 0; r < height; ++r) memcpy(flat + (size_t)r * width * 3, img_bytes[r], (size_t)width * 3);
        int rc = gl_set_image(ctx, flat, width, height, 3);
        if (rc == GL_OK) rc = gl_ctx_sync(ctx);
        free(flat);
        if (rc != GL_OK) GLHostFatal("ComputeAffinityMatrices");
    } else if (gl_set_image_rows(ctx, (const uint8_t* const*)img_bytes, width, height) != GL_OK) GLHostFatal("ComputeAffinityMatrices");
    if (gl_set_samples(ctx, sample_indices, sample_size) != GL_OK) GLHostFatal("ComputeAffinityMatrices");
    if (gl_affinity(ctx, g_opt.affinity_kind, g_opt.h_loc, g_opt.h_val, K_A, K_B) != GL_OK) GLHostFatal("ComputeAffinityMatrices");
}

/* Reference: hpc/affinity.c:264-336 (-no_approx).  The N x N matrix is never formed: the handle stands for K through
 * the per-pixel sums the later stages need (csrc/full_filter.cu). */
void ComputeEntireAffinityMatrix(Mat* K, const png_bytep* const img_bytes, const int width, const int height)
{
    gl_ctx* ctx = GLHostContext();
    *K = NULL;
    if (g_opt.color) {
        uint8_t* flat = (uint8_t*)malloc((size_t)width * height * 3);
        for (int r = 0; r < height; ++r) memcpy(flat + (size_t)r * width * 3, img_bytes[r], (size_t)width * 3);
        int rc = gl_set_image(ctx, flat, width, height, 3);
        if (rc == GL_OK) rc = gl_ctx_sync(ctx);
        free(flat);
        if (rc != GL_OK) GLHostFatal("ComputeEntireAffinityMatrix");
    } else if (gl_set_image_rows(ctx, (const uint8_t* const*)img_bytes, width, height) != GL_OK) GLHostFatal("ComputeEntireAffinityMatrix");
    if (gl_full_affinity(ctx, g_opt.affinity_kind, g_opt.h_loc, g_opt.h_val, K) != GL_OK) GLHostFatal("ComputeEntireAffinityMatrix");
}
