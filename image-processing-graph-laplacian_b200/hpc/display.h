/* Entry points of the reference's hpc/display.h:6-14 that the approximation path uses. */
#ifndef GLB200_DISPLAY_H
#define GLB200_DISPLAY_H
#include "petsc_compat.h"
void WriteVec(Vec v, const char* const filename);
void WriteDiagMat(Mat x, const char* const filename);
png_bytep* ComputeResultFromLaplacian(const png_bytep* const img_bytes, Mat phi, Mat Pi, const unsigned int width, const unsigned int height);
png_bytep* ComputeResultFromEntireLaplacian(const png_bytep* const img_bytes, Mat Lapl, const unsigned int width, const unsigned int height);
#endif
