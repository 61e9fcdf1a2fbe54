/* Same entry points as the reference's hpc/affinity.h:5-7. */
#ifndef GLB200_AFFINITY_H
#define GLB200_AFFINITY_H
#include "petsc_compat.h"
void ComputeAffinityMatrices(Mat* K_A, Mat* K_B, const png_bytep* const img_bytes, const int width, const int height,
                             const unsigned int sample_size, const unsigned int* sample_indices);
void ComputeEntireAffinityMatrix(Mat* K, const png_bytep* const img_bytes, const int width, const int height);
#endif
