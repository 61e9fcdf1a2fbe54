/* Driver of the B200 build: a drop-in for the reference's hpc/image_processing.c (main :279-335,
 * ApproximationComputation :183-277) with the same stage order, the same stdout lines and the same six options
 * (-f, -num_eigvals, -no_approx, -use_slepc, -opti_gs, -inv_it_epsilon; hpc/README.md:21-29).  Underneath there is
 * no PETSc/SLEPc/MPI: every stage function below calls libglcuda.so (include/gl_cuda.h) and runs on the GPU.
 *
 * What differs from the reference at HEAD, on purpose:
 *   - the stages after the eigensolve (Nystroem, Permutation, MatPow, ComputeResultFromLaplacian) are live; the
 *     reference has them inside a comment and returns NULL (hpc/image_processing.c:237-276);
 *   - knobs the reference hard-codes are options (glhost.h): -sample_size, -sampling, -seed, -affinity, -h_loc,
 *     -h_val, -filter_gain, -filter_pow, -gram_schmidt, -color, -ngpus, -synthetic WxH, -o OUTPUT;
 *   - "processes" are one forked process per GPU (-ngpus), not MPI ranks.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include "affinity.h"
#include "display.h"
#include "eigendecomposition.h"
#include "glhost.h"
#include "gram_schmidt.h"
#include "inverse_power_it.h"
#include "laplacian.h"
#include "nystroem.h"
#include "read_img.h"
#include "sampling.h"
#include "utils.h"
#include "write_img.h"

/* wall clock after the device has drained, so that the per-stage lines mean what the reference's do */
static double StageClock(void)
{
    if (gl_ctx_sync(GLHostContext()) != GL_OK) GLHostFatal("gl_ctx_sync");
    return GLHostWtime();
}

static void GetFilePath(char* const filename)
{
    if (g_opt.synthetic_w > 0 && g_opt.synthetic_h > 0) {
        snprintf(filename, PETSC_MAX_PATH_LEN, "synthetic %dx%d", g_opt.synthetic_w, g_opt.synthetic_h);
        return;
    }
    if (!OptionsGetString("-f", filename, PETSC_MAX_PATH_LEN)) {
        if (GLHostRank() == 0) fprintf(stderr, "No filename found (option -f)\n");
        GLHostFinalize();
        exit(1);
    }
}

/* Every rank ends up with the whole image, as after the reference's ReadAndBcastImage (:45-76).  The ranks are
 * processes of one box, so each decodes the file itself instead of receiving `height` broadcasts. */
static int ReadImageOnEveryRank(const char* const filename, png_bytep** const img_bytes, int* const width, int* const height)
{
    if (g_opt.synthetic_w > 0 && g_opt.synthetic_h > 0) {
        const int w = g_opt.synthetic_w, h = g_opt.synthetic_h, ch = g_opt.color ? 3 : 1;
        gl_ctx* ctx = GLHostContext();
        if (gl_set_synthetic_image(ctx, w, h, ch, 1234) != GL_OK) GLHostFatal("gl_set_synthetic_image");
        unsigned char* flat = (unsigned char*)malloc((size_t)w * h * ch);
        if (gl_get_image(ctx, flat) != GL_OK) GLHostFatal("gl_get_image");
        *img_bytes = (png_bytep*)malloc(sizeof(png_bytep) * h);
        for (int i = 0; i < h; ++i) {
            (*img_bytes)[i] = (png_bytep)malloc((size_t)w * ch);
            memcpy((*img_bytes)[i], flat + (size_t)i * w * ch, (size_t)w * ch);
        }
        free(flat);
        *width = w;
        *height = h;
        return 0;
    }
    if (g_opt.color) return read_png_rgb(filename, img_bytes, width, height, NULL);
    return read_png(filename, img_bytes, width, height);
}

static PetscInt GetNumberEigenvalues(const unsigned int sample_size)
{
    int num_eigvals = 0;
    const int found = OptionsGetInt("-num_eigvals", &num_eigvals);
    if (!found || num_eigvals < 0 || num_eigvals >= (int)sample_size) {
        num_eigvals = (int)sample_size - 1;
        if (GLHostRank() == 0)
            fprintf(stderr, "Invalid or invalid number of eigenvalues found (option -num_eigvals), so using %d\n", num_eigvals);
    }
    return num_eigvals;
}

static PetscInt GetOptiGramSchmidt(void)
{
    int value = 1;
    if (!OptionsGetInt("-opti_gs", &value) || value < 1) value = 1;
    return value;
}

static PetscScalar GetInverseIterationEpsilon(void)
{
    double epsilon = 0.1;
    if (!OptionsGetScalar("-inv_it_epsilon", &epsilon)) epsilon = 0.1;
    return epsilon;
}

/* -no_approx (hpc/image_processing.c:155-181), same three stages and stdout lines; the N x N matrices are matrix-free */
static png_bytep* EntireComputation(const png_bytep* const img_bytes, const unsigned int width, const unsigned int height)
{
    Mat K = NULL, Lapl = NULL;
    double t0 = StageClock();
    GLHostPrintf("Computing entire affinity matrix... ");
    ComputeEntireAffinityMatrix(&K, img_bytes, width, height);
    GLHostPrintf("%fs\n", StageClock() - t0);
    t0 = StageClock();
    GLHostPrintf("Computing entire Laplacian matrix... ");
    ComputeEntireLaplacianMatrix(&Lapl, K);
    GLHostPrintf("%fs\n", StageClock() - t0);
    MatDestroy(&K);
    t0 = StageClock();
    GLHostPrintf("Computing output image... ");
    png_bytep* out = ComputeResultFromEntireLaplacian(img_bytes, Lapl, width, height);
    GLHostPrintf("%fs\n", GLHostWtime() - t0);
    MatDestroy(&Lapl);
    return out;
}

static png_bytep* ApproximationComputation(png_bytep* img_bytes, const unsigned int width, const unsigned int height)
{
    unsigned int p = g_opt.sample_size ? g_opt.sample_size : (unsigned int)(width * height * 0.01); /* 1 %, :187 */
    unsigned int* sample_indices = NULL; /* ascending */
    Sampling(width, height, &p, &sample_indices);
    GLHostPrintf("Sample size: %d\n", p);

    const PetscInt m = GetNumberEigenvalues(p);

    double t0 = StageClock();
    GLHostPrintf("Computing affinity matrices... ");
    Mat K_A, K_B;
    ComputeAffinityMatrices(&K_A, &K_B, img_bytes, width, height, p, sample_indices);
    GLHostPrintf("%fs\n", StageClock() - t0);

    t0 = StageClock();
    GLHostPrintf("Computing Laplacian matrices... ");
    Mat L_A, L_B;
    ComputeLaplacianMatrix(&L_A, &L_B, K_A, K_B);
    GLHostPrintf("%fs\n", StageClock() - t0);
    MatDestroy(&K_A);
    MatDestroy(&K_B);

    t0 = StageClock();
    Mat eigvals, eigvecs_A;
    GLHostPrintf("Computing %d smallest eigenvalues... ", m);
    if (OptionsHasName("-use_slepc")) {
        EigendecompositionSmallest(L_A, m, &eigvecs_A, &eigvals, NULL);
    } else {
        const PetscScalar epsilon = GetInverseIterationEpsilon();
        GLHostPrintf("(epsilon: %g) ", epsilon);
        InversePowerIteration(L_A, m, &eigvecs_A, &eigvals, GetOptiGramSchmidt(), epsilon);
    }
    GLHostPrintf("%fs\n", StageClock() - t0);
    WriteDiagMat(eigvals, "results/eigenvalues_laplacian.txt");
    MatDestroy(&L_A);

    Mat eigvals_inv = InverseDiagMat(eigvals);

    t0 = StageClock();
    GLHostPrintf("Computing Nystr\xc3\xb6m approximation... ");
    Mat eigvecs = Nystroem(L_B, eigvecs_A, eigvals_inv, width * height, p, m);
    GLHostPrintf("%fs\n", StageClock() - t0);
    MatDestroy(&eigvecs_A);
    MatDestroy(&eigvals_inv);
    MatDestroy(&L_B);

    Mat eigvecs_perm = Permutation(eigvecs, sample_indices, p);
    MatDestroy(&eigvecs);
    eigvecs = eigvecs_perm;

    if (g_opt.gram_schmidt) {
        t0 = StageClock();
        GLHostPrintf("Orthonormalising eigenvectors... ");
        OrthonormaliseMat(eigvecs, NULL);
        GLHostPrintf("%fs\n", StageClock() - t0);
    }

    Mat f_eigvals = MatPow(eigvals, 6);
    MatDestroy(&eigvals);

    t0 = StageClock();
    GLHostPrintf("Computing output image... ");
    png_bytep* output_img = ComputeResultFromLaplacian(img_bytes, eigvecs, f_eigvals, width, height);
    GLHostPrintf("%fs\n", GLHostWtime() - t0);
    MatDestroy(&eigvecs);
    MatDestroy(&f_eigvals);
    free(sample_indices);
    return output_img;
}

int main(int argc, char** argv)
{
    char filename[PETSC_MAX_PATH_LEN];
    char outname[PETSC_MAX_PATH_LEN];
    PetscMPIInt rank, size;

    if (GLHostInit(argc, argv, &rank, &size)) {
        fprintf(stderr, "could not start the ranks\n");
        return 1;
    }
    const double start_time = GLHostWtime();
    if (rank == 0) mkdir("results", 0777); /* the reference expects results/ to exist (it ships a .gitkeep there) */
    GLHostPrintf("Running with %d processes\n", size);
    GetFilePath(filename);
    if (!OptionsGetString("-o", outname, sizeof outname)) strcpy(outname, "results/output.png");

    int width = 0, height = 0;
    png_bytep *img_bytes = NULL, *output_img = NULL;
    if (ReadImageOnEveryRank(filename, &img_bytes, &width, &height)) {
        GLHostFinalize();
        return 1;
    }
    GLHostPrintf("Read image %s of size %dx%d => %d pixels\n", filename, width, height, width * height);

    if (OptionsHasName("-no_approx")) output_img = EntireComputation(img_bytes, width, height);
    else output_img = ApproximationComputation(img_bytes, width, height);

    if (rank == 0) {
        if (g_opt.color) write_png_rgb("results/input.png", img_bytes, width, height);
        else write_png("results/input.png", img_bytes, width, height);
        if (output_img) {
            if (g_opt.color) write_png_rgb(outname, output_img, width, height);
            else write_png(outname, output_img, width, height);
        }
    }

    const double total = GLHostWtime() - start_time;
    GLHostPrintf("Total computation time: %fs\n", total);
    if (rank == 0 && output_img && !OptionsHasName("-no_approx")) {
        float ms[GL_T_COUNT];
        if (gl_ctx_stage_ms(GLHostContext(), ms) == GL_OK) {
            const float dev = ms[GL_T_SAMPLING] + ms[GL_T_AFFINITY] + ms[GL_T_LAPLACIAN] + ms[GL_T_EIGEN] + ms[GL_T_NYSTROEM] +
                              (g_opt.gram_schmidt ? ms[GL_T_GRAM_SCHMIDT] : 0.f) + ms[GL_T_FILTER];
            GLHostPrintf("Device time per stage (ms): sampling %.3f affinity %.3f laplacian %.3f eigen %.3f nystroem %.3f gram_schmidt %.3f filter %.3f\n",
                         ms[GL_T_SAMPLING], ms[GL_T_AFFINITY], ms[GL_T_LAPLACIAN], ms[GL_T_EIGEN], ms[GL_T_NYSTROEM],
                         g_opt.gram_schmidt ? ms[GL_T_GRAM_SCHMIDT] : 0.f, ms[GL_T_FILTER]);
            if (dev > 0.f) GLHostPrintf("Device throughput: %.1f Mpixel/s on %d GPU(s)\n", (double)width * height / (dev * 1e-3) / 1e6, size);
        }
    }

    for (int i = 0; i < height; ++i) free(img_bytes[i]);
    free(img_bytes);
    if (rank == 0 && output_img) {
        for (int i = 0; i < height; ++i) free(output_img[i]);
        free(output_img);
    }
    GLHostFinalize();
    return 0;
}
