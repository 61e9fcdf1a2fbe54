#include "display.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "glhost.h"
#include "utils.h"

/* hpc/display.c:42-49: PETSc ASCII viewer output of a vector ("Vec Object: ..." header, one value per line). */
void WriteVec(Vec v, const char* const filename)
{
    if (GLHostRank() != 0) return;
    FILE* f = fopen(filename, "w");
    if (!f) {
        fprintf(stderr, "Could not open file %s\n", filename);
        return;
    }
    fprintf(f, "Vec Object: %d MPI processes\n  type: mpi\n", GLHostSize());
    for (unsigned int i = 0; i < v->n; ++i) fprintf(f, "%.18g\n", v->data[i]);
    fclose(f);
}

/* hpc/display.c:51-56 */
void WriteDiagMat(Mat x, const char* const filename)
{
    Vec v = DiagMat2Vec(x);
    WriteVec(v, filename);
    VecDestroy(&v);
}

/* hpc/display.c:58-83: z = y + 3 Phi Pi Phi^T y, values above 255 set to 255, bytes gathered on rank 0.
 * Every rank filters its band of rows into one image shared by the ranks; rank 0 returns it (the others NULL,
 * like OneColMat2pngbytes, hpc/utils.c:508).  The bytes are clamped to [0,255] then truncated: the reference
 * casts a possibly negative double to png_byte (hpc/utils.c:525), which is undefined. */
png_bytep* ComputeResultFromLaplacian(const png_bytep* const img_bytes, Mat phi, Mat Pi, const unsigned int width, const unsigned int height)
{
    (void)img_bytes;  /* the image is already on the device (ComputeAffinityMatrices uploaded it) */
    const unsigned int row_bytes = width * (g_opt.color ? 3u : 1u);
    png_bytep* shared = GLHostSharedImage(row_bytes, height);
    if (!shared) {
        fprintf(stderr, "ComputeResultFromLaplacian: image too large for the shared output buffer\n");
        exit(1);
    }
    if (gl_filter(GLHostContext(), phi, Pi, g_opt.filter_gain, 0, NULL, shared[0]) != GL_OK) GLHostFatal("ComputeResultFromLaplacian");
    GLHostBandDone();
    if (GLHostRank() != 0) return NULL;
    GLHostWaitBands();
    png_bytep* out = (png_bytep*)malloc(sizeof(png_bytep) * height);
    for (unsigned int i = 0; i < height; ++i) {
        out[i] = (png_bytep)malloc(row_bytes);
        memcpy(out[i], shared[i], row_bytes);
    }
    return out;
}

/* hpc/display.c:128-149: z = y - L y, clipped to [0, 255] on both sides. */
png_bytep* ComputeResultFromEntireLaplacian(const png_bytep* const img_bytes, Mat Lapl, const unsigned int width, const unsigned int height)
{
    (void)img_bytes;
    const unsigned int row_bytes = width * (g_opt.color ? 3u : 1u);
    png_bytep* shared = GLHostSharedImage(row_bytes, height);
    if (!shared) {
        fprintf(stderr, "ComputeResultFromEntireLaplacian: image too large for the shared output buffer\n");
        exit(1);
    }
    if (gl_full_result(GLHostContext(), Lapl, NULL, shared[0]) != GL_OK) GLHostFatal("ComputeResultFromEntireLaplacian");
    GLHostBandDone();
    if (GLHostRank() != 0) return NULL;
    GLHostWaitBands();
    png_bytep* out = (png_bytep*)malloc(sizeof(png_bytep) * height);
    for (unsigned int i = 0; i < height; ++i) {
        out[i] = (png_bytep)malloc(row_bytes);
        memcpy(out[i], shared[i], row_bytes);
    }
    return out;
}
