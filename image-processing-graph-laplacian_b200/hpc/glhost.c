#include "glhost.h"
#include "glshare.h"

#include <signal.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/prctl.h>
#include <sys/time.h>
#include <sys/wait.h>
#include <unistd.h>

GLHostOptions g_opt;

static int g_argc;
static char** g_argv;
static gl_ctx* g_ctx = NULL;
static int g_rank = 0, g_size = 1;
static pid_t g_children[64];
static pid_t g_parent = 0;   /* rank 0's pid, as seen by the forked ranks */
static GLShare g_share;                     /* hand-over of the shared mapping (glshare.h) */
static unsigned char* g_shared = NULL;
static size_t g_shared_bytes = 0;
static png_bytep* g_shared_rows = NULL;

/* ---------------------------------------------------------------- options */
static int find_opt(const char* name)
{
    for (int i = 1; i < g_argc; ++i)
        if (!strcmp(g_argv[i], name)) return i;
    return 0;
}
int OptionsHasName(const char* name) { return find_opt(name) != 0; }
int OptionsGetString(const char* name, char* out, size_t len)
{
    int i = find_opt(name);
    if (!i || i + 1 >= g_argc) return 0;
    strncpy(out, g_argv[i + 1], len - 1);
    out[len - 1] = 0;
    return 1;
}
int OptionsGetInt(const char* name, int* out)
{
    char buf[64];
    if (!OptionsGetString(name, buf, sizeof buf)) return 0;
    char* end;
    long v = strtol(buf, &end, 10);
    if (end == buf) return 0;
    *out = (int)v;
    return 1;
}
int OptionsGetScalar(const char* name, double* out)
{
    char buf[64];
    if (!OptionsGetString(name, buf, sizeof buf)) return 0;
    char* end;
    double v = strtod(buf, &end);
    if (end == buf) return 0;
    *out = v;
    return 1;
}

void OptionsInit(int argc, char** argv)
{
    g_argc = argc;
    g_argv = argv;
    memset(&g_opt, 0, sizeof g_opt);
    g_opt.affinity_kind = GL_BILATERAL; /* hpc/affinity.c:121 */
    g_opt.h_loc = 40.0;                 /* hpc/affinity.c:118 */
    g_opt.h_val = 30.0;                 /* hpc/affinity.c:117 */
    g_opt.filter_gain = 3.0;            /* hpc/display.c:73 */
    g_opt.filter_pow = 1.0;             /* MatPow leaves the values unchanged, hpc/utils.c:721 */
    g_opt.ngpus = 1;
    char buf[128];
    if (OptionsGetString("-affinity", buf, sizeof buf)) {
        if (!strcmp(buf, "photometric")) g_opt.affinity_kind = GL_PHOTOMETRIC;
        else if (!strcmp(buf, "spatial")) g_opt.affinity_kind = GL_SPATIAL;
        else if (!strcmp(buf, "nlm") || !strcmp(buf, "NLM")) {
            g_opt.affinity_kind = GL_NLM;    /* python/affinity_methods/NLM.py */
            g_opt.h_val = 3.0;               /* NLM.py:12 (overridden by -h_val) */
        }
        else if (strcmp(buf, "bilateral")) fprintf(stderr, "Unknown -affinity %s, using bilateral\n", buf);
    }
    if (OptionsGetString("-sampling", buf, sizeof buf)) g_opt.sampling_random = !strcmp(buf, "random");
    int iv;
    double dv;
    if (OptionsGetInt("-seed", &iv)) g_opt.seed = (unsigned)iv;
    if (OptionsGetInt("-sample_size", &iv) && iv > 0) g_opt.sample_size = (unsigned)iv;
    if (OptionsGetScalar("-h_loc", &dv) && dv > 0) g_opt.h_loc = dv;
    if (OptionsGetScalar("-h_val", &dv) && dv > 0) g_opt.h_val = dv;
    if (OptionsGetScalar("-filter_gain", &dv)) g_opt.filter_gain = dv;
    if (OptionsGetScalar("-filter_pow", &dv)) { g_opt.filter_pow = dv; g_opt.filter_pow_set = 1; }
    g_opt.gram_schmidt = OptionsHasName("-gram_schmidt");
    g_opt.color = OptionsHasName("-color");
    if (OptionsGetInt("-dump_eigvecs", &iv) && iv > 0) g_opt.dump_eigvecs = iv;
    g_opt.dump_scaled = OptionsHasName("-dump_scaled");
    g_opt.inverse_iteration = OptionsHasName("-inverse_iteration");
    if (OptionsGetInt("-ngpus", &iv) && iv >= 1 && iv <= 64) g_opt.ngpus = iv;
    if (OptionsGetString("-synthetic", buf, sizeof buf)) sscanf(buf, "%dx%d", &g_opt.synthetic_w, &g_opt.synthetic_h);
}

/* ---------------------------------------------------------------- runtime */
gl_ctx* GLHostContext(void) { return g_ctx; }
int GLHostRank(void) { return g_rank; }
int GLHostSize(void) { return g_size; }

/* rank 0 takes every other rank down with it (MPI_Abort semantics): they may be blocked inside a collective or a hand-over
 * that will never complete.  Async-signal-safe: kill and _exit only. */
static void kill_ranks_and_exit(int sig)
{
    (void)sig;
    for (int r = 1; r < g_size; ++r)
        if (g_children[r] > 0) kill(g_children[r], SIGTERM);
    _exit(1);
}

void GLHostFatal(const char* where)
{
    fprintf(stderr, "[rank %d] %s: %s\n", g_rank, where, gl_last_error());
    if (g_size > 1) {
        if (g_rank == 0) kill_ranks_and_exit(0);
        /* a failing rank > 0 tells rank 0, whose SIGTERM handler ends all the others (and PR_SET_PDEATHSIG ends whoever is left
         * when rank 0 is gone) */
        if (g_parent > 1) kill(g_parent, SIGTERM);
    }
    exit(1);
}

double GLHostWtime(void)
{
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return (double)tv.tv_sec + 1e-6 * (double)tv.tv_usec;
}

void GLHostPrintf(const char* fmt, ...)
{
    if (g_rank != 0) return;
    va_list ap;
    va_start(ap, fmt);
    vprintf(fmt, ap);
    va_end(ap);
    fflush(stdout);
}

int GLHostInit(int argc, char** argv, int* rank, int* size)
{
    OptionsInit(argc, argv);
    g_size = g_opt.ngpus;
    g_rank = 0;
    /* the NCCL id and the output image live in memory shared across the fork */
    unsigned char* id = (unsigned char*)mmap(NULL, 4096, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (id == MAP_FAILED) return 1;
    volatile int* id_ready = (volatile int*)(id + 256);
    g_share.done = (volatile int*)(id + 512);
    g_share.consumed = (volatile int*)(id + 768);
    g_shared_bytes = (size_t)1 << 31;  /* reserve 2 GiB of address space; pages are touched on demand */
    g_shared = (unsigned char*)mmap(NULL, g_shared_bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (g_shared == MAP_FAILED) return 1;
    /* fork BEFORE any CUDA call: one process per GPU, SPMD like the reference's MPI ranks */
    g_parent = getpid();
    for (int r = 1; r < g_size; ++r) {
        pid_t pid = fork();
        if (pid < 0) return 1;
        if (pid == 0) {
            g_rank = r;
            /* never outlive rank 0: the kernel delivers SIGTERM when the parent dies (checked once for a parent already gone) */
            prctl(PR_SET_PDEATHSIG, SIGTERM);
            if (getppid() != g_parent) _exit(1);
            break;
        }
        g_children[r] = pid;
    }
    if (g_rank == 0 && g_size > 1) signal(SIGTERM, kill_ranks_and_exit);
    g_share.parent = g_rank == 0 ? 0 : g_parent;
    g_share.rank = g_rank;
    g_share.size = g_size;
    g_share.gen = 0;
    if (gl_ctx_create(&g_ctx, g_rank, g_rank, g_size) != GL_OK) GLHostFatal("gl_ctx_create");
    if (g_size > 1) {
        if (g_rank == 0) {
            if (gl_comm_unique_id(id) != GL_OK) GLHostFatal("gl_comm_unique_id");
            __sync_synchronize();
            *id_ready = 1;
        } else {
            while (!*id_ready) usleep(1000);
            __sync_synchronize();
        }
        if (gl_comm_init(g_ctx, id) != GL_OK) GLHostFatal("gl_comm_init");
    }
    *rank = g_rank;
    *size = g_size;
    return 0;
}

png_bytep* GLHostSharedImage(unsigned int width, unsigned int height)
{
    static unsigned int capacity = 0;    /* the row table grows with the tallest view asked for (images, column dumps) */
    if ((size_t)width * height > g_shared_bytes) return NULL;
    if (height > capacity) {
        png_bytep* rows = (png_bytep*)realloc(g_shared_rows, sizeof(png_bytep) * height);
        if (!rows) return NULL;
        g_shared_rows = rows;
        capacity = height;
    }
    for (unsigned int i = 0; i < height; ++i) g_shared_rows[i] = g_shared + (size_t)i * width;
    return g_shared_rows;
}

/* The shared mapping is handed over with the protocol of glshare.h: Begin (wait until rank 0 has read the previous content),
 * write, BandDone, [rank 0: WaitBands, read], Release.  (Not a waitpid: the ranks still have to tear their NCCL communicator
 * down together in GLHostFinalize.) */
void GLHostSharedBegin(void) { GLShareBegin(&g_share); }

void GLHostBandDone(void) { GLShareDone(&g_share); }

static int a_rank_died(void* unused)
{
    (void)unused;
    for (int r = 1; r < g_size; ++r) {   /* a rank that died will never report: do not wait for ever */
        int st = 0;
        if (waitpid(g_children[r], &st, WNOHANG) == g_children[r]) {
            fprintf(stderr, "rank %d ended before writing its band\n", r);
            return 1;
        }
    }
    return 0;
}

void GLHostWaitBands(void)
{
    if (g_rank != 0) return;
    if (GLShareWait(&g_share, a_rank_died, NULL)) exit(1);
}

void GLHostSharedRelease(void) { GLShareRelease(&g_share); }

void GLHostFinalize(void)
{
    if (g_ctx) gl_ctx_destroy(g_ctx);   /* collective over the ranks of one box (ncclCommDestroy) */
    g_ctx = NULL;
    if (g_rank != 0) _exit(0);
    for (int r = 1; r < g_size; ++r) {
        int st = 0;
        if (g_children[r] > 0 && waitpid(g_children[r], &st, 0) == g_children[r] && (!WIFEXITED(st) || WEXITSTATUS(st) != 0))
            fprintf(stderr, "rank %d ended abnormally\n", r);
    }
}

PetscErrorCode MatDestroy(Mat* m)
{
    if (m && *m) {
        gl_mat_destroy(*m);
        *m = NULL;
    }
    return 0;
}

Vec VecCreateHost(unsigned int n)
{
    Vec v = (Vec)malloc(sizeof(*v));
    v->n = n;
    v->data = (double*)calloc(n, sizeof(double));
    return v;
}

PetscErrorCode VecDestroy(Vec* v)
{
    if (v && *v) {
        free((*v)->data);
        free(*v);
        *v = NULL;
    }
    return 0;
}
