/* hpc/write_img.h:4 of the reference, with png_bytep from petsc_compat.h instead of <png.h>. */
#ifndef GLB200_WRITE_IMG_H
#define GLB200_WRITE_IMG_H
#include "petsc_compat.h"

/* 8-bit grey, non-interlaced; 0 on success, -1 on failure */
int write_png(const char* const filename, png_bytep* img_bytes, const unsigned int width, const unsigned int height);
/* interleaved 8-bit RGB rows (-color) */
int write_png_rgb(const char* const filename, png_bytep* img_bytes, const unsigned int width, const unsigned int height);
#endif
