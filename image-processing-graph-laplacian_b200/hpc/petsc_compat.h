/*
 * petsc_compat.h -- the few PETSc / libpng names the reference's hpc/ headers use, re-declared so that
 * code written against the hpc/ headers of David-Wobrock/image-processing-graph-laplacian compiles against this
 * build with no PETSc, SLEPc, MPI or libpng installed.
 *
 * Mat is an opaque handle to a device matrix owned by libglcuda.so (include/gl_cuda.h): the host never
 * dereferences it.  Vec is a small host vector (the reference only uses Vec arrays for p-length iterates
 * inside its inverse iteration, hpc/inverse_power_it.c:12-47).
 */
#ifndef GLB200_PETSC_COMPAT_H
#define GLB200_PETSC_COMPAT_H

#include <stddef.h>
#include "../../include/gl_cuda.h"

typedef gl_mat* Mat;
typedef struct GLVec_s { double* data; unsigned int n; }* Vec;
typedef double PetscScalar;
typedef double PetscReal;
typedef int PetscInt;
typedef int PetscMPIInt;
typedef int PetscBool;
typedef int PetscErrorCode;
#define PETSC_TRUE 1
#define PETSC_FALSE 0
#define PETSC_MAX_PATH_LEN 4096

typedef unsigned char png_byte;
typedef png_byte* png_bytep;

/* MatDestroy / VecDestroy as the reference calls them (hpc/image_processing.c:170,177,210-211,235-239) */
PetscErrorCode MatDestroy(Mat* m);
PetscErrorCode VecDestroy(Vec* v);
Vec VecCreateHost(unsigned int n);

/* host-side context shared by the hpc/ entry points (glhost.c) */
gl_ctx* GLHostContext(void);
int GLHostRank(void);
int GLHostSize(void);
void GLHostFatal(const char* where);              /* prints gl_last_error() and exit(1) */

#endif
