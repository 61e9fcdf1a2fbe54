/* PNG reader on zlib alone (this build has no libpng).
 *
 * Replaces read_png of the reference, hpc/read_img.c:9-65 (same name, arguments and return values: 0 on success, -1
 * with "Could not open file %s" on stderr when the file cannot be opened, -1 on a malformed file).  The reference
 * hands the caller `height` separately malloc'd rows of 8-bit grey samples; colour files are reduced to grey with
 * libpng's png_set_rgb_to_gray(png, 1, -1, -1) (hpc/read_img.c:47-50), whose default weights are the integer
 * coefficients 6968/23434/2366 over 32768 with truncation -- restated in rgb_to_gray() below.
 *
 * Deliberate differences, all on inputs the reference mis-reads: grey+alpha and RGBA files have their alpha channel
 * dropped (the reference leaves it interleaved in the row, hpc/read_img.c:52, so bear.png comes out scrambled);
 * palette files are expanded through their palette; 16-bit samples keep their high byte; 1/2/4-bit grey is scaled
 * to 0..255.  Adam7-interlaced files are decoded too.
 */
#include "read_img.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

static unsigned int be32(const unsigned char* p) { return ((unsigned int)p[0] << 24) | ((unsigned int)p[1] << 16) | ((unsigned int)p[2] << 8) | p[3]; }

static unsigned char rgb_to_gray(unsigned int r, unsigned int g, unsigned int b)
{
    if (r == g && g == b) return (unsigned char)r;
    return (unsigned char)((6968u * r + 23434u * g + 2366u * b) >> 15);
}

static int paeth(int a, int b, int c)
{
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

/* undo the per-scanline filters of one (sub)image in place; `bpp` = bytes per complete pixel (>= 1) */
static int unfilter(unsigned char* data, size_t rows, size_t row_bytes, unsigned int bpp)
{
    const size_t stride = row_bytes + 1;
    for (size_t y = 0; y < rows; ++y) {
        unsigned char* cur = data + y * stride + 1;
        const unsigned char* up = y ? data + (y - 1) * stride + 1 : NULL;
        switch (data[y * stride]) {
        case 0: break;
        case 1:
            for (size_t i = bpp; i < row_bytes; ++i) cur[i] = (unsigned char)(cur[i] + cur[i - bpp]);
            break;
        case 2:
            if (up) for (size_t i = 0; i < row_bytes; ++i) cur[i] = (unsigned char)(cur[i] + up[i]);
            break;
        case 3:
            for (size_t i = 0; i < row_bytes; ++i) {
                int a = i >= bpp ? cur[i - bpp] : 0, b = up ? up[i] : 0;
                cur[i] = (unsigned char)(cur[i] + ((a + b) >> 1));
            }
            break;
        case 4:
            for (size_t i = 0; i < row_bytes; ++i) {
                int a = i >= bpp ? cur[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= bpp) ? up[i - bpp] : 0;
                cur[i] = (unsigned char)(cur[i] + paeth(a, b, c));
            }
            break;
        default: return -1;
        }
    }
    return 0;
}

typedef struct {
    unsigned int width, height, depth, color_type, interlace, channels_in;
    unsigned char palette[256][3];
} PngHeader;

/* sample k (0-based, in units of `depth` bits) of a scanline, scaled to 8 bits for grey, raw for palette indices */
static unsigned int get_sample(const unsigned char* row, size_t k, unsigned int depth)
{
    if (depth == 8) return row[k];
    if (depth == 16) return row[2 * k];
    const unsigned int per = 8 / depth, shift = (unsigned int)((per - 1 - k % per) * depth);
    return (row[k / per] >> shift) & ((1u << depth) - 1u);
}

/* one decoded scanline -> out_channels (1 or 3) bytes per pixel at out[x * xstep * out_channels] */
static void emit_row(const PngHeader* h, const unsigned char* row, size_t w, unsigned char* out, size_t x0, size_t xstep, int out_channels)
{
    for (size_t i = 0; i < w; ++i) {
        unsigned int r, g, b;
        switch (h->color_type) {
        case 0:
        case 4: {
            unsigned int v = get_sample(row, i * h->channels_in, h->depth);
            if (h->depth < 8) v = v * 255u / ((1u << h->depth) - 1u);
            r = g = b = v;
            break;
        }
        case 3: {
            unsigned int idx = get_sample(row, i, h->depth);
            r = h->palette[idx][0]; g = h->palette[idx][1]; b = h->palette[idx][2];
            break;
        }
        default:
            r = get_sample(row, i * h->channels_in, h->depth);
            g = get_sample(row, i * h->channels_in + 1, h->depth);
            b = get_sample(row, i * h->channels_in + 2, h->depth);
        }
        unsigned char* o = out + (x0 + i * xstep) * (size_t)out_channels;
        if (out_channels == 1) o[0] = rgb_to_gray(r, g, b);
        else { o[0] = (unsigned char)r; o[1] = (unsigned char)g; o[2] = (unsigned char)b; }
    }
}

static int decode(const char* filename, png_bytep** row_pointers, int* width, int* height, int out_channels, int* file_is_colour)
{
    FILE* f = fopen(filename, "rb");
    if (!f) {
        fprintf(stderr, "Could not open file %s\n", filename);
        return -1;
    }
    fseek(f, 0, SEEK_END);
    long fsize = ftell(f);
    fseek(f, 0, SEEK_SET);
    unsigned char* file = fsize > 0 ? (unsigned char*)malloc((size_t)fsize) : NULL;
    if (!file || fread(file, 1, (size_t)fsize, f) != (size_t)fsize) {
        fclose(f);
        free(file);
        return -1;
    }
    fclose(f);
    static const unsigned char sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    int rc = -1;
    unsigned char *idat = NULL, *raw = NULL;
    size_t idat_len = 0;
    PngHeader h;
    memset(&h, 0, sizeof h);
    int have_ihdr = 0;
    if (fsize < 8 || memcmp(file, sig, 8)) goto done;
    idat = (unsigned char*)malloc((size_t)fsize);
    for (size_t pos = 8; pos + 12 <= (size_t)fsize;) {
        const unsigned int len = be32(file + pos);
        const unsigned char* type = file + pos + 4;
        const unsigned char* body = file + pos + 8;
        if (pos + 12 + (size_t)len > (size_t)fsize) goto done;
        if (be32(body + len) != (unsigned int)crc32(crc32(0L, type, 4), body, len)) goto done;
        if (!memcmp(type, "IHDR", 4) && len >= 13) {
            h.width = be32(body); h.height = be32(body + 4); h.depth = body[8]; h.color_type = body[9]; h.interlace = body[12];
            static const unsigned int chans[7] = {1, 0, 3, 1, 2, 0, 4};
            if (h.color_type > 6 || !chans[h.color_type] || !h.width || !h.height || body[10] || body[11] || h.interlace > 1) goto done;
            if (!(h.depth == 8 || h.depth == 16 || ((h.color_type == 0 || h.color_type == 3) && (h.depth == 1 || h.depth == 2 || h.depth == 4)))) goto done;
            if (h.color_type == 3 && h.depth == 16) goto done;
            h.channels_in = chans[h.color_type];
            have_ihdr = 1;
        } else if (!memcmp(type, "PLTE", 4)) {
            for (unsigned int i = 0; i < len / 3 && i < 256; ++i) memcpy(h.palette[i], body + 3 * i, 3);
        } else if (!memcmp(type, "IDAT", 4)) {
            memcpy(idat + idat_len, body, len);
            idat_len += len;
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || !idat_len) goto done;
    {
        const unsigned int bits = h.depth * h.channels_in;
        const unsigned int bpp = bits >= 8 ? bits / 8 : 1;
        /* sub-images: the whole picture, or the seven Adam7 passes */
        static const unsigned int ax0[7] = {0, 4, 0, 2, 0, 1, 0}, ay0[7] = {0, 0, 4, 0, 2, 0, 1};
        static const unsigned int adx[7] = {8, 8, 4, 4, 2, 2, 1}, ady[7] = {8, 8, 8, 4, 4, 2, 2};
        const int passes = h.interlace ? 7 : 1;
        size_t pw[7], ph[7], prb[7], total = 0;
        for (int p = 0; p < passes; ++p) {
            pw[p] = h.interlace ? (h.width + adx[p] - 1 - ax0[p]) / adx[p] : h.width;
            ph[p] = h.interlace ? (h.height + ady[p] - 1 - ay0[p]) / ady[p] : h.height;
            prb[p] = (pw[p] * bits + 7) / 8;
            if (pw[p] && ph[p]) total += ph[p] * (prb[p] + 1);
        }
        raw = (unsigned char*)malloc(total ? total : 1);
        uLongf got = (uLongf)total;
        if (!raw || uncompress(raw, &got, idat, (uLong)idat_len) != Z_OK || got != total) goto done;

        png_bytep* rows = (png_bytep*)malloc(sizeof(png_bytep) * h.height);
        for (unsigned int y = 0; y < h.height; ++y) rows[y] = (png_bytep)malloc((size_t)h.width * (size_t)out_channels);
        unsigned char* cursor = raw;
        int bad = 0;
        for (int p = 0; p < passes && !bad; ++p) {
            if (!pw[p] || !ph[p]) continue;
            if (unfilter(cursor, ph[p], prb[p], bpp)) { bad = 1; break; }
            for (size_t y = 0; y < ph[p]; ++y) {
                const size_t oy = h.interlace ? ay0[p] + y * ady[p] : y;
                emit_row(&h, cursor + y * (prb[p] + 1) + 1, pw[p], rows[oy], h.interlace ? ax0[p] : 0, h.interlace ? adx[p] : 1, out_channels);
            }
            cursor += ph[p] * (prb[p] + 1);
        }
        if (bad) {
            for (unsigned int y = 0; y < h.height; ++y) free(rows[y]);
            free(rows);
            goto done;
        }
        *row_pointers = rows;
        *width = (int)h.width;
        *height = (int)h.height;
        if (file_is_colour) *file_is_colour = (h.color_type == 2 || h.color_type == 6 || h.color_type == 3);
        rc = 0;
    }
done:
    free(file);
    free(idat);
    free(raw);
    return rc;
}

int read_png(const char* const filename, png_bytep** row_pointers, int* const width, int* const height)
{
    return decode(filename, row_pointers, width, height, 1, NULL);
}

int read_png_rgb(const char* const filename, png_bytep** row_pointers, int* const width, int* const height, int* const file_is_colour)
{
    return decode(filename, row_pointers, width, height, 3, file_is_colour);
}
