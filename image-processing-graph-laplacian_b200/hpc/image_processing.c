/* Driver of the B200 build: a drop-in for the reference's hpc/image_processing.c (main :279-335,
 * ApproximationComputation :183-277) with the same stage order, the same stdout lines and the same six options
 * (-f, -num_eigvals, -no_approx, -use_slepc, -opti_gs, -inv_it_epsilon; hpc/README.md:21-29).  Underneath there is
 * no PETSc/SLEPc/MPI: every stage function below calls libglcuda.so (include/gl_cuda.h) and runs on the GPU.
 *
 * What differs from the reference at HEAD, on purpose:
 *   - the stages after the eigensolve (Nystroem, Permutation, MatPow, ComputeResultFromLaplacian) are live; the
 *     reference has them inside a comment and returns NULL (hpc/image_processing.c:237-276);
 *   - knobs the reference hard-codes are options (glhost.h): -sample_size, -sampling, -seed, -affinity, -h_loc,
 *     -h_val, -filter_gain, -filter_pow, -gram_schmidt, -dump_eigvecs K, -dump_scaled, -color, -ngpus, -synthetic WxH, -o OUTPUT;
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

/* Everything the reference reads from the PETSc options database while it runs (hpc/image_processing.c:82-153),
 * parsed once up front.  Same defaults and the same stderr notes. */
typedef struct RunOptions {
    char input[PETSC_MAX_PATH_LEN];   /* -f (or "synthetic WxH") */
    char output[PETSC_MAX_PATH_LEN];  /* -o, default results/output.png */
    int num_eigvals;                  /* -num_eigvals, -1 when absent: resolved against the sample count later */
    int no_approx, use_slepc;         /* -no_approx, -use_slepc */
    int opti_gs;                      /* -opti_gs, values < 1 become 1 (:134-138) */
    double inv_it_epsilon;            /* -inv_it_epsilon, default 0.1 (:148-152) */
} RunOptions;

static void ParseRunOptions(RunOptions* o)
{
    memset(o, 0, sizeof *o);
    if (g_opt.synthetic_w > 0 && g_opt.synthetic_h > 0) {
        snprintf(o->input, sizeof o->input, "synthetic %dx%d", g_opt.synthetic_w, g_opt.synthetic_h);
    } else if (!OptionsGetString("-f", o->input, sizeof o->input)) {
        if (GLHostRank() == 0) fprintf(stderr, "No filename found (option -f)\n");   /* :88-92 */
        GLHostFinalize();
        exit(1);
    }
    if (!OptionsGetString("-o", o->output, sizeof o->output)) strcpy(o->output, "results/output.png");
    if (!OptionsGetInt("-num_eigvals", &o->num_eigvals)) o->num_eigvals = -1;
    o->no_approx = OptionsHasName("-no_approx");
    o->use_slepc = OptionsHasName("-use_slepc");
    if (!OptionsGetInt("-opti_gs", &o->opti_gs) || o->opti_gs < 1) o->opti_gs = 1;
    if (!OptionsGetScalar("-inv_it_epsilon", &o->inv_it_epsilon)) o->inv_it_epsilon = 0.1;
}

/* m = -num_eigvals, or sample_size - 1 with the reference's note on stderr when absent or out of range (:96-108) */
static PetscInt ResolveEigenpairCount(const RunOptions* o, const unsigned int sample_size)
{
    int m = o->num_eigvals;
    if (m < 0 || m >= (int)sample_size) {
        m = (int)sample_size - 1;
        if (GLHostRank() == 0)
            fprintf(stderr, "Invalid or invalid number of eigenvalues found (option -num_eigvals), so using %d\n", m);
    }
    return m;
}

/* Every rank ends up with the whole image, as after the reference's ReadAndBcastImage (:45-76).  The ranks are
 * processes of one box, so each decodes the file itself instead of receiving `height` broadcasts. */
static int LoadImageOnEveryRank(const RunOptions* o, png_bytep** const img_bytes, int* const width, int* const height)
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
    if (g_opt.color) return read_png_rgb(o->input, img_bytes, width, height, NULL);
    return read_png(o->input, img_bytes, width, height);
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

static png_bytep* ApproximationComputation(const RunOptions* o, png_bytep* img_bytes, const unsigned int width, const unsigned int height)
{
    unsigned int p = g_opt.sample_size ? g_opt.sample_size : (unsigned int)(width * height * 0.01); /* 1 %, :187 */
    unsigned int* sample_indices = NULL; /* ascending */
    Sampling(width, height, &p, &sample_indices);
    GLHostPrintf("Sample size: %d\n", p);

    const PetscInt m = ResolveEigenpairCount(o, p);

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
    if (o->use_slepc) {
        EigendecompositionSmallest(L_A, m, &eigvecs_A, &eigvals, NULL);
    } else {
        GLHostPrintf("(epsilon: %g) ", o->inv_it_epsilon);
        InversePowerIteration(L_A, m, &eigvecs_A, &eigvals, o->opti_gs, o->inv_it_epsilon);
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

    /* hpc/image_processing.c:255-260 (eigenvectors 0..2 as text and as images; here on request: at 4K a column is 8.3 M lines) */
    for (int k = 0; k < g_opt.dump_eigvecs && k < (int)m; ++k) {
        char path[96];
        snprintf(path, sizeof path, "results/eigenvector_%d_laplacian.txt", k);
        WriteMatCol(eigvecs, (unsigned int)k, path);
        snprintf(path, sizeof path, "results/eigenvector_%d_laplacian.png", k);
        WritePngMatCol(eigvecs, (unsigned int)k, width, height, path);
    }

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
    RunOptions opt;
    PetscMPIInt rank, size;

    if (GLHostInit(argc, argv, &rank, &size)) {
        fprintf(stderr, "could not start the ranks\n");
        return 1;
    }
    const double start_time = GLHostWtime();
    if (rank == 0) mkdir("results", 0777); /* the reference expects results/ to exist (it ships a .gitkeep there) */
    GLHostPrintf("Running with %d processes\n", size);
    ParseRunOptions(&opt);

    int width = 0, height = 0;
    png_bytep *img_bytes = NULL, *output_img = NULL;
    if (LoadImageOnEveryRank(&opt, &img_bytes, &width, &height)) {
        GLHostFinalize();
        return 1;
    }
    GLHostPrintf("Read image %s of size %dx%d => %d pixels\n", opt.input, width, height, width * height);

    output_img = opt.no_approx ? EntireComputation(img_bytes, width, height) : ApproximationComputation(&opt, img_bytes, width, height);

    if (rank == 0) {
        if (g_opt.color) write_png_rgb("results/input.png", img_bytes, width, height);
        else write_png("results/input.png", img_bytes, width, height);
        if (output_img) {
            if (g_opt.color) write_png_rgb(opt.output, output_img, width, height);
            else write_png(opt.output, output_img, width, height);
        }
    }

    const double total = GLHostWtime() - start_time;
    GLHostPrintf("Total computation time: %fs\n", total);
    if (rank == 0 && output_img && !opt.no_approx) {
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
