#include "sampling.h"

#include <stdlib.h>

#include "glhost.h"

/* Reference: hpc/sampling.c:30-33 (dispatcher) over UniformSampling :6-23; with `-sampling random` the
 * Python prototype's random_sample (python/sampling/random.py:8-16) seeded with `-seed`.  The indices are
 * produced by a device kernel and copied back because the reference hands the caller a malloc'd host array
 * (caller frees, hpc/sampling.c:14). */
void Sampling(const int width, const int height, unsigned int* const sample_size, unsigned int** const sample_indices)
{
    gl_ctx* ctx = GLHostContext();
    int r0, r1;
    if (gl_get_band(ctx, &r0, &r1) != GL_OK) {
        /* no image uploaded yet: only the geometry matters for sampling */
        if (gl_set_synthetic_image(ctx, width, height, 1, 0) != GL_OK) GLHostFatal("Sampling");
    }
    unsigned int actual = 0;
    int rc = g_opt.sampling_random ? gl_sampling_random(ctx, *sample_size, g_opt.seed, &actual)
                                   : gl_sampling_uniform(ctx, *sample_size, &actual);
    if (rc != GL_OK) GLHostFatal("Sampling");
    *sample_size = actual;
    *sample_indices = (unsigned int*)malloc(sizeof(unsigned int) * actual);
    if (gl_get_samples(ctx, *sample_indices, actual, NULL) != GL_OK) GLHostFatal("Sampling");
}
