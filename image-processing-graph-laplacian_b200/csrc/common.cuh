// Internal definitions shared by the libglcuda.so translation units.
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "../../include/gl_cuda.h"

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
void gl_set_error(const char* fmt, ...);

#define GL_CUDA_CHECK(expr)                                                                             \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            gl_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));    \
            return GL_ERR_CUDA;                                                                         \
        }                                                                                               \
    } while (0)

#define GL_CHECK(expr)                   \
    do {                                 \
        int _s = (expr);                 \
        if (_s != GL_OK) return _s;      \
    } while (0)

#define GL_REQUIRE(cond, ...)            \
    do {                                 \
        if (!(cond)) {                   \
            gl_set_error(__VA_ARGS__);   \
            return GL_ERR_ARG;           \
        }                                \
    } while (0)

// the same checks for use inside a `do { ... } while (0)` block that owns buffers: record the status and leave the block,
// so that the clean-up after it runs
#define GL_CUDA_BREAK(rc, expr)                                                                         \
    {                                                                                                   \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            gl_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));    \
            (rc) = GL_ERR_CUDA;                                                                         \
            break;                                                                                      \
        }                                                                                               \
    }
#define GL_BREAK(rc, expr)               \
    {                                    \
        (rc) = (expr);                   \
        if ((rc) != GL_OK) break;        \
    }

#define GL_LAUNCH_CHECK(ctx)                                                                            \
    do {                                                                                                \
        (ctx)->launches++;                                                                              \
        cudaError_t _e = cudaGetLastError();                                                            \
        if (_e != cudaSuccess) {                                                                        \
            gl_set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return GL_ERR_CUDA;                                                                         \
        }                                                                                               \
    } while (0)

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int64_t ceil_div(int64_t x, int64_t m) { return (x + m - 1) / m; }
// padded column count of Phi / rows of W^T: 64, 128 or a multiple of 256, so that (m_pad / 8) 16-byte column
// groups are either a power of two <= 32 or a whole number of warps (filter.cu), and GEMM N tiles are full.
static inline int gl_m_pad(int m) { return m <= 64 ? 64 : (m <= 128 ? 128 : (int)round_up(m, 256)); }

// ---------------------------------------------------------------------------------------------
// refcounted device buffers + matrices
// ---------------------------------------------------------------------------------------------
struct gl_buf {
    void* ptr = nullptr;
    size_t bytes = 0;
    int refs = 1;
    gl_ctx* owner = nullptr;
};

struct gl_mat {
    int kind = 0;
    int64_t rows = 0, cols = 0;  // logical
    int64_t local_rows = 0;      // band rows for KB/PHI
    int64_t ld = 0;              // elements
    int elem_bytes = 0;
    double scale = 1.0;
    gl_buf* buf = nullptr;       // main storage
    gl_buf* aux = nullptr;       // KB: fp64 row sums D[p] (summed over ranks)
    gl_buf* tiles = nullptr;     // KB: int4 per 512-pixel tile {first entry in `starts`, block count, storage offset, 0}
    gl_buf* starts = nullptr;    // KB: first internal sample slot of every stored block
    gl_buf* perm = nullptr;      // KB: internal sample slot -> index in the caller's sample list (u32 [p_pad + 64])
    // KB, patch layout (patch.cu): per 64 x 16 pixel patch a gathered sample list; null when the handle holds the blocked layout only
    gl_buf* pt_info = nullptr;   // int4 per patch {first slot block, blocks, samples in reach, 0}
    gl_buf* pt_slots = nullptr;  // u32 [pt_blocks][32]: sample index of every slot (0xffffffff: empty)
    gl_buf* pt_buf = nullptr;    // fp16 A tiles [pt_blocks * 8][128 pixels][32 slots]
    int64_t pt_blocks = 0;
    int64_t pt_ksteps = 0;       // 16-slot K steps over all M tiles (the MMA work the extrapolation issues: x 128 x m_pad x 16 x 2 flop)
    int pt_npatch = 0;
    int64_t total_blocks = 0;    // KB: stored [512 x kbs] blocks (see affinity.cu)
    int kbs = 64;                // KB: sample slots per block (64 or 32)
    gl_buf* dscale = nullptr;    // optional device double holding `scale` (L_B: -alpha), so no host sync is needed
    bool scale_on_host = true;   // false until the device value has been fetched
    int refs = 1;
    gl_ctx* ctx = nullptr;
    // KB / PHI bookkeeping
    int64_t q0 = 0;              // first raster pixel of the band
    int p = 0, p_pad = 0, m = 0, m_pad = 0;
    int channels = 0;            // KB: channels of the image-weighted sums T in aux; PHI: channels of proj
    gl_buf* proj = nullptr;      // PHI: c = Phi^T y, fp64 [m_pad][channels], from the T sums (nystroem) -- see filter.cu
    unsigned long long image_epoch = 0;  // image the sums / proj belong to
    unsigned long long sample_epoch = 0; // KB: the sample set it was computed from
    int aff_kind = 0;            // KB: the affinity that produced it (to rebuild K_A y_S in fp64)
    double aff_h_loc = 0, aff_h_val = 0;
    float phi_scale = 1.0f;      // PHI: stored value * phi_scale = logical (always 1; scale folded in the epilogue)
    // PHI not computed yet (gl_nystroem with option lazy_phi): the retained inputs it will be computed from
    gl_mat *def_LB = nullptr, *def_U = nullptr, *def_muinv = nullptr;
};

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
struct gl_nccl;  // comm.cu
#define GL_MARKS 128

struct gl_ctx {
    int device = 0, rank = 0, world = 1;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    long long launches = 0;
    bool total_started = false;

    // image (whole image on every rank)
    int width = 0, height = 0, channels = 0;
    int64_t n = 0;
    int row0 = 0, row1 = 0;  // band of image rows owned by this rank
    int64_t q0 = 0, q1 = 0;  // raster range of the band
    gl_buf* img = nullptr;   // u8 [n * channels]
    // gl_run's image upload runs on its own stream so that the stages that do not read pixels (sampling, the patch lists) overlap it;
    // gl_image_ready() makes ctx->stream wait for it, and is called before the first kernel of a run that reads ctx->img
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_h2d = nullptr, ev_prev = nullptr;
    bool h2d_pending = false;
    const uint8_t* host_pixels = nullptr;  // during a multi-GPU gl_run: the caller's host image (band + sample pixels are uploaded)

    unsigned long long image_epoch = 0;  // bumped by every gl_set_*image
    unsigned long long sample_epoch = 0; // bumped by every new sample set (gl_sampling_*, gl_set_samples)
    int filter_apply_impl = 0;  // 0 = warp-per-row kernel when the shape allows, 1 = always the generic kernel
    int projection_mode = 0;  // 0 = c from the affinity sums (default), 1 = always recompute c with a pass over Phi
    int fuse_filter = 1;      // gl_run: apply the filter inside the extrapolation GEMM's epilogue when possible
    int lazy_phi = 1;         // gl_nystroem returns a deferred Phi; gl_filter then runs extrapolation + filter as one pass
    int keep_phi = 0;         // fused filter (gl_run, gl_filter on a deferred Phi): 0 = Phi tiles are consumed in the GEMM epilogue and never
                              // written to HBM (nobody reads them; a later consumer recomputes), 1 = write Phi as well (the reference's data flow)

    // samples
    unsigned p = 0;
    int p_pad = 0;
    gl_buf* samples = nullptr;  // u32 [p_pad] (padding = 0xffffffff)
    std::vector<uint32_t> h_samples;  // host mirror of the sample indices (tile-table construction, affinity.cu)
    bool h_samples_valid = false;

    // cached K_B tile table (affinity.cu): rebuilt only when the geometry, the samples or the cutoff change
    gl_buf* tile_tab = nullptr;
    gl_buf* tile_starts = nullptr;
    gl_buf* tile_perm = nullptr;
    int tile_strips = 1;      // column strips of the cached layout
    int tile_kbs = 64;        // sample slots per block of the cached layout
    long long phi_limit_mb = 0;   // option phi_limit_mb: gl_run stores no Phi larger than this (0 = no limit but the memory)
    bool last_phi_stored = true;  // whether the last fused gl_run wrote Phi
    size_t phi_nomem_bytes = 0;   // smallest Phi whose allocation failed on this context (0: none yet): not tried again
    int kb_block = 64;        // option kb_block: 64 or 32 sample slots per stored K_B block
    int kb_strips = 0;        // option kb_strips: 0 = choose the strip count with the fewest blocks, n = force n strips
    std::vector<int4> h_tile_tab;
    std::vector<uint32_t> tab_samples;  // the samples and parameters the cached table was built from
    int64_t tab_key[6] = {-1, -1, -1, -1, -1, -1};
    int64_t tile_total_blocks = 0;
    int kb_cutoff = 1;        // option kb_cutoff: 1 = skip sample blocks whose K_B entries fp16 flushes to zero, 0 = dense
    int pt_dual = 1;          // option pt_dual: 1 = the patch extrapolation runs as two pipelines per SM when every patch is resident (one channel)
    int z8_direct = 1;        // option z8_direct: the fused patch path writes the u8 result straight into a pinned host destination (no D2H copy)
    int kb_layout = 1;        // option kb_layout: 1 = patch layout for the spatially decaying affinities (patch.cu), 0 = always blocked
    bool want_blocked = false;  // set by gl_run_resident around gl_affinity when the path will need Phi itself (no fused filter)

    // size-keyed cache of freed device blocks (no cudaMalloc in steady state)
    std::multimap<size_t, void*> free_blocks;
    size_t bytes_cached = 0, bytes_live = 0, bytes_peak = 0;

    // stage timers
    cudaEvent_t ev_begin[GL_T_COUNT] = {}, ev_end[GL_T_COUNT] = {};
    bool ev_valid[GL_T_COUNT] = {};
    cudaEvent_t marks[GL_MARKS] = {};
    gl_buf* flush_buf = nullptr;   // scratch of gl_ctx_flush_l2

    // pinned staging for small D2H/H2D
    void* pinned = nullptr;
    size_t pinned_bytes = 0;

    gl_nccl* comm = nullptr;

    // Deferred status (gl_run_resident): inside the one-call path the stages do not stop for the host; what they would have
    // reported lands in `dstat` on the device and is read back with the results, or at the next call (GL_DS_* slots)
    bool async_mode = false;
    gl_buf* dstat = nullptr;      // int [GL_DS_COUNT]
    int* hstat = nullptr;         // pinned mirror
    cudaEvent_t stat_ev = nullptr;
    bool stat_pending = false;
    // patch layout: blocks to allocate without asking the device first, learnt from the last run with the same key
    int64_t pt_cap_blocks = 0;
    int64_t pt_cap_key[5] = {-1, -1, -1, -1, -1};

    // options
    int gemm_impl = 0;        // 0 = tcgen05 (default), 1 = simple CUDA-core checker kernel
    int gram_impl = 0;        // orthonormalise: 0 = tcgen05 Gram when m_pad % 256 == 0, 1 = always the CUDA-core tiles
    int gram_lbo = 8192;      // leading byte offset of the MN-major operand descriptors (tuning/debug)
    int gemm_stages = 0;      // 0 = automatic, 3 | 4 = force that ring depth (tuning)
    int gemm_prefetch = 2;    // blocked A: L2-prefetch the A blocks of the tile this many iterations ahead (0 = off)
    int eig_largest = 0;      // 1: keep the m LARGEST eigenpairs (descending) instead of the smallest (ascending)
    int jacobi_max_sweeps = 40;
    int jacobi_inner = 1;     // inner 16 x 16 Jacobi sweeps per pair visit (0 = until converged, at most 12); 1 is enough: the outer sweeps repeat
    // largest relative off-diagonal of G^T G at convergence.  Phi is stored in fp16 (relative rounding 2^-11 = 4.9e-4): the
    // eigenvectors are converged to a tenth of that, so their error stays an order below the storage rounding (tighten with
    // option jacobi_tol; the eigenvalues themselves are refined by fp64 Rayleigh quotients and come out at ~1e-7)
    float jacobi_tol = 5e-5f;
    int verbose = 0;
};

enum { GL_DS_SAMPLING = 0, GL_DS_PT_OVERFLOW = 1, GL_DS_PT_BLOCKS = 2, GL_DS_JACOBI = 3, GL_DS_JACOBI_SWEEPS = 4, GL_DS_JACOBI_OFF = 5,
       GL_DS_PT_KSTEPS = 6 /* and 7: u64 */, GL_DS_PT_MAXNB = 8 /* most slot blocks of any patch */, GL_DS_COUNT = 16 };
int gl_status_begin(gl_ctx* ctx);                    // zero the device words (start of an asynchronous run)
int gl_status_flush(gl_ctx* ctx);                    // enqueue their copy to the host
int gl_status_check(gl_ctx* ctx, bool* pt_overflow); // wait for the copy and turn the words into a status / error message
// host mirror of the sample indices (fetched from the device on first use after a device-side sampling)
int gl_host_samples(gl_ctx* ctx);

int gl_alloc(gl_ctx* ctx, size_t bytes, gl_buf** out);
int gl_image_ready(gl_ctx* ctx);   // ctx->stream waits for a pending image upload (no-op otherwise)
void gl_buf_release(gl_buf* b);
gl_mat* gl_mat_new(gl_ctx* ctx, int kind);
int gl_ensure_pinned(gl_ctx* ctx, size_t bytes);

struct StageTimer {
    gl_ctx* ctx;
    int stage;
    StageTimer(gl_ctx* c, int s) : ctx(c), stage(s) { cudaEventRecord(ctx->ev_begin[s], ctx->stream); }
    ~StageTimer() {
        cudaEventRecord(ctx->ev_end[stage], ctx->stream);
        ctx->ev_valid[stage] = true;
    }
};

// collectives (comm.cu): in-place sum over ranks on ctx->stream; no-ops when world == 1
int gl_allreduce_f64(gl_ctx* ctx, double* dev, size_t count);
int gl_allreduce_f32(gl_ctx* ctx, float* dev, size_t count);
void gl_comm_destroy(gl_ctx* ctx);

// patch layout (patch.cu)
bool gl_patch_applicable(const gl_ctx* ctx, int kind);
int gl_patch_affinity(gl_ctx* ctx, int kind, double h_loc, double h_val, gl_mat* KB);
int gl_patch_download(gl_ctx* ctx, const gl_mat* KB, double scale, double* dst_dev);
bool gl_patch_nystroem_fits(int m, int C);
int gl_patch_nystroem_filter(gl_ctx* ctx, const gl_mat* L_B, const float* U, int ldU, int m, const double* mu_inv, const float* scales,
                             const float* w, int C, int clip_low, float* z, uint8_t* z8);
// computes the blocked storage of a K_B handle that holds the patch layout only (same image and samples required)
int gl_kb_require_blocked(gl_ctx* ctx, gl_mat* KB);

// stage implementations (one .cu each)
int gl_impl_sampling_uniform(gl_ctx* ctx, unsigned requested, unsigned* actual);
int gl_impl_sampling_random(gl_ctx* ctx, unsigned requested, uint32_t seed, unsigned* actual);
int gl_impl_synthetic(gl_ctx* ctx, uint32_t seed);
int gl_impl_affinity(gl_ctx* ctx, int kind, double h_loc, double h_val, gl_mat** K_A, gl_mat** K_B);
int gl_impl_laplacian(gl_ctx* ctx, gl_mat* K_A, gl_mat* K_B, gl_mat** L_A, gl_mat** L_B);
int gl_impl_eigensolve(gl_ctx* ctx, gl_mat* L_A, int m, gl_mat** eigvecs, gl_mat** eigvals, gl_mat** eigvals_inv);
int gl_impl_inverse_iteration(gl_ctx* ctx, gl_mat* L_A, int m, int opti_gs, double epsilon, int max_iterations, gl_mat** eigvecs,
                              gl_mat** eigvals, gl_mat** eigvals_inv, int* iterations_out, double* residual_out);
// small dense fp64 algebra (dense_small.cu): C = alpha op(A) op(B) + beta C, row-major
int gl_dgemm(gl_ctx* ctx, int M, int N, int K, double alpha, const double* A, int lda, int ta, const double* B, int ldb, int tb, double beta,
             double* C, int ldc);
int gl_chol_inverse_upper(gl_ctx* ctx, double* G, int m, int ld, double* T, int* status_dev);
// request to apply the filter inside the extrapolation GEMM (gl_nystroem_filter)
struct gl_fused_filter {
    gl_mat* f_eigvals = nullptr;
    double gain = 0.0;
    int clip_low = 0;
    float* z_f32 = nullptr;   // host destinations as in gl_filter (may be null)
    uint8_t* z_u8 = nullptr;
};
int gl_impl_nystroem(gl_ctx* ctx, gl_mat* L_B, gl_mat* phi_A, gl_mat* eigvals_inv, gl_mat** phi, const gl_fused_filter* ff = nullptr);
int gl_phi_defer(gl_ctx* ctx, gl_mat* L_B, gl_mat* phi_A, gl_mat* eigvals_inv, gl_mat** phi);
int gl_phi_materialise(gl_ctx* ctx, gl_mat* phi, const gl_fused_filter* ff = nullptr);
int gl_filter_weights_from_proj(gl_ctx* ctx, const double* proj, const double* f, double gain, int m, int m_pad, int C, float* w);
// parts > 0: z = y + the sum of `parts` row partials in zpart; parts == 0: z_dev / z8_dev already hold the band's filtered pixels (the
// patch kernel wrote them).  Either way the sample pixels' rows are patched and the result is copied to the host destinations.
// z8_direct: the u8 result was (and the sample rows are) written by the kernels straight into the caller's pinned host image (its band
// starts at this pointer): no device copy of it exists and nothing is copied afterwards
int gl_filter_fused_finish(gl_ctx* ctx, gl_mat* phi, const float* zpart, int parts, const float* w, const float* U, int ldU,
                           int clip_low, float* z_f32, uint8_t* z_u8, gl_buf* z_dev = nullptr, gl_buf* z8_dev = nullptr,
                           uint8_t* z8_direct = nullptr);
int gl_impl_orthonormalise(gl_ctx* ctx, gl_mat* phi, double* norms_out);
int gl_impl_filter(gl_ctx* ctx, gl_mat* phi, gl_mat* f_eigvals, double gain, int clip_low, float* z_f32, uint8_t* z_u8);
int gl_impl_diag_map(gl_ctx* ctx, gl_mat* d, int op, double arg, gl_mat** out);
int gl_impl_full_affinity(gl_ctx* ctx, int kind, double h_loc, double h_val, gl_mat** K);
int gl_impl_full_laplacian(gl_ctx* ctx, gl_mat* K, gl_mat** L);
int gl_impl_full_result(gl_ctx* ctx, gl_mat* L, float* z_f32, uint8_t* z_u8);
// the prototype's experimental blocks (proto.cu)
int gl_impl_sinkhorn(gl_ctx* ctx, gl_mat* phi, gl_mat* Pi, int iterations, gl_mat** W_A, gl_mat** W_ABt);
int gl_impl_smoothing_matrix(gl_ctx* ctx, gl_mat* phi, gl_mat* Pi, gl_mat** V, gl_mat** L);
int gl_impl_matrix_filter(gl_ctx* ctx, gl_mat* V, gl_mat* L, const double* coef, int ncoef, float* z_f32);
int gl_impl_orthogonalisation(gl_ctx* ctx, gl_mat* K_A, gl_mat* K_B, gl_mat** V, gl_mat** Pi);
// optional fusion of the filter application into the GEMM epilogue (nystroem_gemm.cu / filter.cu)
struct gl_gemm_fuse {
    const float* w = nullptr;   // [n_pad][C] filter weights gain * f(lambda) o c
    float* zpart = nullptr;     // [parts][rows][C] partial row dots, to be summed in the order of `parts`
    int C = 0;
    int parts = 0;              // out: number of partials per row the kernel writes
};
// A is either a dense [rows][k_pad] K-major matrix (a_tab == nullptr) or K_B's blocked storage with its tile table
int gl_gemm_kmajor(gl_ctx* ctx, const void* A, int ab_bf16, int64_t rows, int k_pad, const void* Bt, int n_pad,
                   const float* scales, const void* addend, void* D, const int4* a_tab = nullptr, int64_t a_total_blocks = 0,
                   gl_gemm_fuse* fuse = nullptr, const int* a_starts = nullptr, int a_kbs = 64);

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float fast_exp2(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#endif
