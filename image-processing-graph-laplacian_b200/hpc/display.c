#include "display.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "glhost.h"
#include "utils.h"
#include "write_img.h"

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

/* Column `col_num` of x gathered on rank 0 (NULL elsewhere).  Phi is split into bands of rows over the ranks: each one
 * downloads its part of the column into the mapping the ranks share, the way the u8 result image travels; the small
 * replicated matrices (eigenvectors of L_A) are read by rank 0 alone.  Replaces MatGetColumnVector + VecScatterCreateToZero
 * (hpc/display.c:95, hpc/utils.c:498-506). */
static Vec GatherMatCol(Mat x, const unsigned int col_num)
{
    gl_mat_info info;
    if (gl_mat_info_get(x, &info) != GL_OK) GLHostFatal("GatherMatCol");
    const int banded = info.kind == GL_MAT_PHI;
    const unsigned int n = (unsigned int)info.rows;
    if (!banded) {
        if (GLHostRank() != 0) return NULL;
        Vec v = VecCreateHost(n);
        if (gl_mat_download_cols(GLHostContext(), x, (int)col_num, 1, v->data, v->n) != GL_OK) GLHostFatal("GatherMatCol");
        return v;
    }
    /* the shared mapping seen as n doubles (8 n bytes as n/512+1 "rows" of 4096 bytes) */
    png_bytep* rows = GLHostSharedImage(4096, (unsigned int)(((size_t)n * sizeof(double) + 4095) / 4096));
    if (!rows) {
        fprintf(stderr, "GatherMatCol: column too long for the shared buffer\n");
        exit(1);
    }
    double* shared = (double*)rows[0];
    int r0 = 0, r1 = 0;
    if (gl_get_band(GLHostContext(), &r0, &r1) != GL_OK) GLHostFatal("GatherMatCol");
    const size_t width = r1 > r0 ? (size_t)info.local_rows / (size_t)(r1 - r0) : 0;   /* band = image rows r0..r1 */
    GLHostSharedBegin();
    if (info.local_rows > 0 &&
        gl_mat_download_cols(GLHostContext(), x, (int)col_num, 1, shared + (size_t)r0 * width, (size_t)info.local_rows) != GL_OK)
        GLHostFatal("GatherMatCol");
    GLHostBandDone();
    Vec v = NULL;
    if (GLHostRank() == 0) {
        GLHostWaitBands();
        v = VecCreateHost(n);
        memcpy(v->data, shared, sizeof(double) * (size_t)n);
    }
    GLHostSharedRelease();
    return v;
}

/* hpc/display.c:85-100 */
void WriteMatCol(Mat x, const unsigned int col_num, const char* const filename)
{
    Vec v = GatherMatCol(x, col_num);
    if (!v) return;
    WriteVec(v, filename);
    VecDestroy(&v);
}

/* hpc/display.c:102-126: the column as a width x height grey image, values cast to bytes as OneColMat2pngbytes does
 * (hpc/utils.c:525; clamped to [0,255] first, the reference's cast of a negative double is undefined).  Eigenvector
 * entries are O(1/sqrt(n)), so that image is black; `-dump_scaled` maps [min, max] of the column to [0, 255] instead. */
void WritePngMatCol(Mat x, const unsigned int col_num, const unsigned int width, const unsigned int height, const char* const filename)
{
    Vec v = GatherMatCol(x, col_num);
    if (!v) return;
    if (v->n != width * height) {
        fprintf(stderr, "WritePngMatCol: column has %u rows, image %ux%u\n", v->n, width, height);
        VecDestroy(&v);
        return;
    }
    double lo = 0.0, scale = 1.0;
    if (g_opt.dump_scaled) {
        double hi = v->data[0];
        lo = v->data[0];
        for (unsigned int i = 1; i < v->n; ++i) {
            if (v->data[i] < lo) lo = v->data[i];
            if (v->data[i] > hi) hi = v->data[i];
        }
        scale = hi > lo ? 255.0 / (hi - lo) : 0.0;
    }
    png_bytep* img = (png_bytep*)malloc(sizeof(png_bytep) * height);
    for (unsigned int i = 0; i < height; ++i) {
        img[i] = (png_bytep)malloc(width);
        for (unsigned int j = 0; j < width; ++j) {
            double val = (v->data[(size_t)i * width + j] - lo) * scale;
            img[i][j] = (png_byte)(val < 0.0 ? 0.0 : (val > 255.0 ? 255.0 : val));
        }
    }
    write_png(filename, img, width, height);
    for (unsigned int i = 0; i < height; ++i) free(img[i]);
    free(img);
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
    GLHostSharedBegin();
    if (gl_filter(GLHostContext(), phi, Pi, g_opt.filter_gain, 0, NULL, shared[0]) != GL_OK) GLHostFatal("ComputeResultFromLaplacian");
    GLHostBandDone();
    png_bytep* out = NULL;
    if (GLHostRank() == 0) {
        GLHostWaitBands();
        out = (png_bytep*)malloc(sizeof(png_bytep) * height);
        for (unsigned int i = 0; i < height; ++i) {
            out[i] = (png_bytep)malloc(row_bytes);
            memcpy(out[i], shared[i], row_bytes);
        }
    }
    GLHostSharedRelease();
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
    GLHostSharedBegin();
    if (gl_full_result(GLHostContext(), Lapl, NULL, shared[0]) != GL_OK) GLHostFatal("ComputeResultFromEntireLaplacian");
    GLHostBandDone();
    png_bytep* out = NULL;
    if (GLHostRank() == 0) {
        GLHostWaitBands();
        out = (png_bytep*)malloc(sizeof(png_bytep) * height);
        for (unsigned int i = 0; i < height; ++i) {
            out[i] = (png_bytep)malloc(row_bytes);
            memcpy(out[i], shared[i], row_bytes);
        }
    }
    GLHostSharedRelease();
    return out;
}
