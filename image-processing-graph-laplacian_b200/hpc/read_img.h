/* hpc/read_img.h:3 of the reference, with png_bytep from petsc_compat.h instead of <png.h>. */
#ifndef GLB200_READ_IMG_H
#define GLB200_READ_IMG_H
#include "petsc_compat.h"

/* 8-bit grey rows (colour files reduced with libpng's default rgb_to_gray weights); 0 on success, -1 on failure */
int read_png(const char* const filename, png_bytep** row_pointers, int* const width, int* const height);
/* same file as interleaved RGB rows of 3*width bytes (-color); *file_is_colour tells whether the file had colour */
int read_png_rgb(const char* const filename, png_bytep** row_pointers, int* const width, int* const height, int* const file_is_colour);
#endif
