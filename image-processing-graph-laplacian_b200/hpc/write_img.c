/* PNG writer on zlib alone.  Replaces write_png of the reference, hpc/write_img.c:5-53: an 8-bit grey, non-interlaced
 * file from `height` row pointers; returns 0, or -1 with "Could not open file %s" on stderr. */
#include "write_img.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

static void put32(unsigned char* p, unsigned int v)
{
    p[0] = (unsigned char)(v >> 24); p[1] = (unsigned char)(v >> 16); p[2] = (unsigned char)(v >> 8); p[3] = (unsigned char)v;
}

static int chunk(FILE* f, const char* type, const unsigned char* body, unsigned int len)
{
    unsigned char head[8], tail[4];
    put32(head, len);
    memcpy(head + 4, type, 4);
    uLong c = crc32(0L, head + 4, 4);
    if (len) c = crc32(c, body, len);
    put32(tail, (unsigned int)c);
    return fwrite(head, 1, 8, f) == 8 && (!len || fwrite(body, 1, len, f) == len) && fwrite(tail, 1, 4, f) == 4 ? 0 : -1;
}

static int encode(const char* filename, png_bytep* img_bytes, unsigned int width, unsigned int height, int channels)
{
    FILE* f = fopen(filename, "wb");
    if (!f) {
        fprintf(stderr, "Could not open file %s\n", filename);
        return -1;
    }
    const size_t row_bytes = (size_t)width * (size_t)channels;
    const size_t raw_len = (size_t)height * (row_bytes + 1);
    unsigned char* raw = (unsigned char*)malloc(raw_len ? raw_len : 1);
    /* filter type 2 (Up) on every row but the first: smooth images compress well and decoding stays trivial */
    for (unsigned int y = 0; y < height; ++y) {
        unsigned char* dst = raw + (size_t)y * (row_bytes + 1);
        dst[0] = y ? 2 : 0;
        if (!y) memcpy(dst + 1, img_bytes[0], row_bytes);
        else for (size_t i = 0; i < row_bytes; ++i) dst[1 + i] = (unsigned char)(img_bytes[y][i] - img_bytes[y - 1][i]);
    }
    uLongf zlen = compressBound((uLong)raw_len);
    unsigned char* z = (unsigned char*)malloc(zlen);
    int rc = -1;
    if (raw && z && compress2(z, &zlen, raw, (uLong)raw_len, 6) == Z_OK) {
        static const unsigned char sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
        unsigned char ihdr[13];
        put32(ihdr, width);
        put32(ihdr + 4, height);
        ihdr[8] = 8;                                   /* bit depth */
        ihdr[9] = channels == 3 ? 2 : 0;               /* colour type: RGB or grey */
        ihdr[10] = ihdr[11] = ihdr[12] = 0;            /* deflate, adaptive filtering, no interlace */
        rc = fwrite(sig, 1, 8, f) == 8 ? 0 : -1;
        if (!rc) rc = chunk(f, "IHDR", ihdr, 13);
        if (!rc) rc = chunk(f, "IDAT", z, (unsigned int)zlen);
        if (!rc) rc = chunk(f, "IEND", NULL, 0);
    }
    free(raw);
    free(z);
    fclose(f);
    return rc;
}

int write_png(const char* const filename, png_bytep* img_bytes, const unsigned int width, const unsigned int height)
{
    return encode(filename, img_bytes, width, height, 1);
}

int write_png_rgb(const char* const filename, png_bytep* img_bytes, const unsigned int width, const unsigned int height)
{
    return encode(filename, img_bytes, width, height, 3);
}
