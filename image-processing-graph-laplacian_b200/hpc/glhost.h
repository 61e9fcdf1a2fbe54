/* Host runtime of the drop-in: process-per-GPU SPMD start-up (replaces SlepcInitialize / MPI_Comm_rank / MPI_Comm_size,
 * hpc/image_processing.c:30-38), the -name value option table (replaces the PETSc options database) and the knobs
 * the reference hard-codes (hpc/affinity.c:117-121, hpc/display.c:73, hpc/image_processing.c:187,263). */
#ifndef GLB200_GLHOST_H
#define GLB200_GLHOST_H
#include "petsc_compat.h"

typedef struct GLHostOptions {
    int affinity_kind;       /* -affinity bilateral|photometric|spatial|nlm */
    double h_loc, h_val;     /* -h_loc, -h_val */
    int sampling_random;     /* -sampling uniform|random */
    unsigned seed;           /* -seed */
    unsigned sample_size;    /* -sample_size (0 = 1 % of the pixels) */
    double filter_gain;      /* -filter_gain (3.0) */
    double filter_pow;       /* -filter_pow: exponent MatPow really applies; unset = the reference's no-op */
    int filter_pow_set;
    int gram_schmidt;        /* -gram_schmidt */
    int dump_eigvecs;        /* -dump_eigvecs K: write the first K extrapolated eigenvectors (hpc/image_processing.c:255-260 writes 3) */
    int dump_scaled;         /* -dump_scaled: eigenvector PNGs use the column's [min, max] as [0, 255] */
    int inverse_iteration;   /* -inverse_iteration: the reference's inverse subspace iteration instead of the converged solver */
    int color;               /* -color: keep RGB, photometric term on the three channels */
    int ngpus;               /* -ngpus */
    int synthetic_w, synthetic_h; /* -synthetic WxH */
} GLHostOptions;

extern GLHostOptions g_opt;

/* options: "-name value" pairs and bare flags; unknown options are ignored, as PETSc does */
void OptionsInit(int argc, char** argv);
int OptionsGetString(const char* name, char* out, size_t len);
int OptionsGetInt(const char* name, int* out);
int OptionsGetScalar(const char* name, double* out);
int OptionsHasName(const char* name);

/* SPMD start-up: forks ngpus-1 children (one process per GPU), creates the context, joins NCCL */
int GLHostInit(int argc, char** argv, int* rank, int* size);
void GLHostFinalize(void);
/* output image rows shared by all ranks (anonymous shared mapping made before the fork) */
png_bytep* GLHostSharedImage(unsigned int width, unsigned int height);
/* hand-over of that mapping, usable any number of times per run (glshare.h): every rank brackets its write with
 * SharedBegin / BandDone, rank 0 reads between WaitBands and SharedRelease, the others call SharedRelease right away */
void GLHostSharedBegin(void);     /* every rank: wait until rank 0 has read the previous content */
void GLHostBandDone(void);        /* every rank: my part of the shared buffer is written */
void GLHostWaitBands(void);       /* rank 0: wait until every rank has reported */
void GLHostSharedRelease(void);   /* every rank: done with this use (rank 0: after reading) */
double GLHostWtime(void);
void GLHostPrintf(const char* fmt, ...);          /* rank-0 stdout, like PetscPrintf(PETSC_COMM_WORLD, ...) */
#endif
