/* Same entry point as the reference's hpc/sampling.h:1; runs on the device (libglcuda gl_sampling_*). */
#ifndef GLB200_SAMPLING_H
#define GLB200_SAMPLING_H
void Sampling(const int width, const int height, unsigned int* const sample_size, unsigned int** const sample_indices);
#endif
