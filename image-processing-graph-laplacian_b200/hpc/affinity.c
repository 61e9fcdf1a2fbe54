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

/* Reference: hpc/affinity.c:264-336 (-no_approx).  The N x N matrix is O(N^2) memory ("very memory consuming",
 * hpc/README.md:19) and outside the accelerated path. */
void ComputeEntireAffinityMatrix(Mat* K, const png_bytep* const img_bytes, const int width, const int height)
{
    (void)img_bytes; (void)width; (void)height;
    *K = NULL;
    fprintf(stderr, "ComputeEntireAffinityMatrix: the -no_approx path is not part of this build\n");
}
