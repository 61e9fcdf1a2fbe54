// Context, allocator, matrix handles, image upload and the one-call pipeline of libglcuda.so.
#include "common.cuh"

static thread_local char g_err[512] = "";

void gl_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    if (getenv("GLB200_VERBOSE")) fprintf(stderr, "[libglcuda] %s\n", g_err);
}

extern "C" {

int gl_version(void) { return 100; }
const char* gl_last_error(void) { return g_err; }

void gl_default_params(gl_params* p)
{
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->affinity_kind = GL_BILATERAL;
    p->h_loc = 40.0;
    p->h_val = 30.0;
    p->sampling_random = 0;
    p->seed = 0;
    p->sample_size = 0;
    p->num_eigvals = -1;
    p->gain = 3.0;
    p->power = 1.0;
    p->gram_schmidt = 0;
    p->clip_low = 0;
}

int gl_device_count(int* count)
{
    GL_REQUIRE(count, "gl_device_count: null");
    int n = 0;
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        gl_set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        return GL_ERR_CUDA;
    }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, i) == cudaSuccess && prop.major == 10) ++ok;
    }
    *count = ok;
    return GL_OK;
}

int gl_memory_stats(gl_ctx* ctx, size_t* live, size_t* cached, size_t* peak, int reset_peak)
{
    GL_REQUIRE(ctx, "gl_memory_stats: null");
    if (live) *live = ctx->bytes_live;
    if (cached) *cached = ctx->bytes_cached;
    if (peak) *peak = ctx->bytes_peak;
    if (reset_peak) ctx->bytes_peak = ctx->bytes_live;
    return GL_OK;
}

int gl_kernel_launches(gl_ctx* ctx, long long* count)
{
    GL_REQUIRE(ctx && count, "gl_kernel_launches: null");
    *count = ctx->launches;
    return GL_OK;
}

// -------------------------------------------------------------------------------------------
int gl_ctx_create(gl_ctx** out, int device, int rank, int world)
{
    GL_REQUIRE(out, "gl_ctx_create: null out");
    GL_REQUIRE(world >= 1 && rank >= 0 && rank < world, "gl_ctx_create: bad rank %d / world %d", rank, world);
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        gl_set_error("no CUDA device (%s); libglcuda has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "count 0");
        return GL_ERR_CUDA;
    }
    GL_REQUIRE(device >= 0 && device < n, "gl_ctx_create: device %d out of range (0..%d)", device, n - 1);
    cudaDeviceProp prop;
    GL_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        gl_set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return GL_ERR_CUDA;
    }
    GL_CUDA_CHECK(cudaSetDevice(device));
    gl_ctx* ctx = new gl_ctx();
    ctx->device = device;
    ctx->rank = rank;
    ctx->world = world;
    ctx->sm_count = prop.multiProcessorCount;
    GL_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    GL_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    GL_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_h2d, cudaEventDisableTiming));
    GL_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_prev, cudaEventDisableTiming));
    for (int i = 0; i < GL_T_COUNT; ++i) {
        GL_CUDA_CHECK(cudaEventCreate(&ctx->ev_begin[i]));
        GL_CUDA_CHECK(cudaEventCreate(&ctx->ev_end[i]));
    }
    for (int i = 0; i < GL_MARKS; ++i) GL_CUDA_CHECK(cudaEventCreate(&ctx->marks[i]));
    GL_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->stat_ev, cudaEventDisableTiming));
    GL_CUDA_CHECK(cudaMallocHost((void**)&ctx->hstat, sizeof(int) * GL_DS_COUNT));
    memset(ctx->hstat, 0, sizeof(int) * GL_DS_COUNT);
    GL_CHECK(gl_alloc(ctx, sizeof(int) * GL_DS_COUNT, &ctx->dstat));
    GL_CUDA_CHECK(cudaMemsetAsync(ctx->dstat->ptr, 0, sizeof(int) * GL_DS_COUNT, ctx->stream));
    const char* v = getenv("GLB200_VERBOSE");
    ctx->verbose = v ? atoi(v) : 0;
    if (const char* g = getenv("GLB200_GEMM")) gl_ctx_set_option(ctx, "gemm", g);
    if (const char* g = getenv("GLB200_JACOBI_TOL")) gl_ctx_set_option(ctx, "jacobi_tol", g);
    if (const char* g = getenv("GLB200_KB_BLOCK")) gl_ctx_set_option(ctx, "kb_block", g);
    if (const char* g = getenv("GLB200_KB_LAYOUT")) gl_ctx_set_option(ctx, "kb_layout", g);
    *out = ctx;
    return GL_OK;
}

int gl_ctx_set_option(gl_ctx* ctx, const char* key, const char* value)
{
    GL_REQUIRE(ctx && key && value, "gl_ctx_set_option: null");
    if (!strcmp(key, "gemm")) {
        if (!strcmp(value, "tcgen05")) ctx->gemm_impl = 0;
        else if (!strcmp(value, "simple")) ctx->gemm_impl = 1;
        else GL_REQUIRE(false, "option gemm: want tcgen05|simple, got %s", value);
    } else if (!strcmp(key, "projection")) {
        if (!strcmp(value, "sums")) ctx->projection_mode = 0;
        else if (!strcmp(value, "recompute")) ctx->projection_mode = 1;
        else GL_REQUIRE(false, "option projection: want sums|recompute, got %s", value);
    } else if (!strcmp(key, "fuse_filter")) {
        ctx->fuse_filter = atoi(value) != 0;
    } else if (!strcmp(key, "lazy_phi")) {
        ctx->lazy_phi = atoi(value) != 0;
    } else if (!strcmp(key, "keep_phi")) {
        ctx->keep_phi = atoi(value) != 0;
    } else if (!strcmp(key, "filter_apply")) {
        if (!strcmp(value, "warp")) ctx->filter_apply_impl = 0;
        else if (!strcmp(value, "generic")) ctx->filter_apply_impl = 1;
        else GL_REQUIRE(false, "option filter_apply: want warp|generic, got %s", value);
    } else if (!strcmp(key, "gram")) {
        if (!strcmp(value, "tcgen05")) ctx->gram_impl = 0;
        else if (!strcmp(value, "simple")) ctx->gram_impl = 1;
        else GL_REQUIRE(false, "option gram: want tcgen05|simple, got %s", value);
    } else if (!strcmp(key, "gram_lbo")) {
        ctx->gram_lbo = atoi(value);
    } else if (!strcmp(key, "gemm_stages")) {
        ctx->gemm_stages = atoi(value);
        GL_REQUIRE(ctx->gemm_stages == 0 || ctx->gemm_stages == 3 || ctx->gemm_stages == 4, "option gemm_stages: want 0|3|4");
    } else if (!strcmp(key, "gemm_prefetch")) {
        ctx->gemm_prefetch = atoi(value);
        GL_REQUIRE(ctx->gemm_prefetch >= 0 && ctx->gemm_prefetch <= 16, "option gemm_prefetch: want 0..16");
    } else if (!strcmp(key, "kb_cutoff")) {
        ctx->kb_cutoff = atoi(value) != 0;
    } else if (!strcmp(key, "kb_layout")) {
        if (!strcmp(value, "patch")) ctx->kb_layout = 1;
        else if (!strcmp(value, "blocked")) ctx->kb_layout = 0;
        else GL_REQUIRE(false, "option kb_layout: want patch|blocked, got %s", value);
    } else if (!strcmp(key, "phi_limit_mb")) {
        ctx->phi_limit_mb = atoll(value);
        GL_REQUIRE(ctx->phi_limit_mb >= 0, "option phi_limit_mb: want >= 0 (0 = no limit)");
    } else if (!strcmp(key, "kb_block")) {
        ctx->kb_block = atoi(value);
        GL_REQUIRE(ctx->kb_block == 64 || ctx->kb_block == 32, "option kb_block: want 64|32");
    } else if (!strcmp(key, "kb_strips")) {
        ctx->kb_strips = atoi(value);
        GL_REQUIRE(ctx->kb_strips >= 0 && ctx->kb_strips <= 64, "option kb_strips: want 0..64");
    } else if (!strcmp(key, "z8_direct")) {
        ctx->z8_direct = atoi(value) != 0;
    } else if (!strcmp(key, "pt_dual")) {
        ctx->pt_dual = atoi(value) != 0;
    } else if (!strcmp(key, "eig_largest")) {
        ctx->eig_largest = atoi(value) != 0;   // EigendecompositionLargest (hpc/eigendecomposition.c:116-119)
    } else if (!strcmp(key, "jacobi_max_sweeps")) {
        ctx->jacobi_max_sweeps = atoi(value);
    } else if (!strcmp(key, "jacobi_inner")) {
        ctx->jacobi_inner = atoi(value);
    } else if (!strcmp(key, "jacobi_tol")) {
        ctx->jacobi_tol = (float)atof(value);
    } else if (!strcmp(key, "verbose")) {
        ctx->verbose = atoi(value);
    } else {
        GL_REQUIRE(false, "unknown option %s", key);
    }
    return GL_OK;
}

int gl_ctx_sync(gl_ctx* ctx)
{
    GL_REQUIRE(ctx, "gl_ctx_sync: null");
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    GL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return gl_status_check(ctx, nullptr);   // what an asynchronous gl_run_resident had to report, if anything
}

int gl_ctx_destroy(gl_ctx* ctx)
{
    if (!ctx) return GL_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    gl_comm_destroy(ctx);
    if (ctx->img) gl_buf_release(ctx->img);
    if (ctx->samples) gl_buf_release(ctx->samples);
    if (ctx->tile_tab) gl_buf_release(ctx->tile_tab);
    if (ctx->tile_starts) gl_buf_release(ctx->tile_starts);
    if (ctx->tile_perm) gl_buf_release(ctx->tile_perm);
    if (ctx->dstat) gl_buf_release(ctx->dstat);
    if (ctx->flush_buf) gl_buf_release(ctx->flush_buf);
    if (ctx->hstat) cudaFreeHost(ctx->hstat);
    if (ctx->stat_ev) cudaEventDestroy(ctx->stat_ev);
    for (auto& kv : ctx->free_blocks) cudaFree(kv.second);
    ctx->free_blocks.clear();
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    for (int i = 0; i < GL_T_COUNT; ++i) {
        cudaEventDestroy(ctx->ev_begin[i]);
        cudaEventDestroy(ctx->ev_end[i]);
    }
    for (int i = 0; i < GL_MARKS; ++i) cudaEventDestroy(ctx->marks[i]);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->ev_h2d) cudaEventDestroy(ctx->ev_h2d);
    if (ctx->ev_prev) cudaEventDestroy(ctx->ev_prev);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return GL_OK;
}

int gl_ctx_stage_ms(gl_ctx* ctx, float* ms)
{
    GL_REQUIRE(ctx && ms, "gl_ctx_stage_ms: null");
    GL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < GL_T_COUNT; ++i) {
        ms[i] = 0.f;
        if (ctx->ev_valid[i]) cudaEventElapsedTime(&ms[i], ctx->ev_begin[i], ctx->ev_end[i]);
    }
    return GL_OK;
}

int gl_ctx_mark(gl_ctx* ctx, int slot)
{
    GL_REQUIRE(ctx && slot >= 0 && slot < GL_MARKS, "gl_ctx_mark: bad slot");
    GL_CUDA_CHECK(cudaEventRecord(ctx->marks[slot], ctx->stream));
    return GL_OK;
}

int gl_ctx_mark_elapsed_ms(gl_ctx* ctx, int a, int b, float* ms)
{
    GL_REQUIRE(ctx && ms && a >= 0 && a < GL_MARKS && b >= 0 && b < GL_MARKS, "gl_ctx_mark_elapsed_ms: bad args");
    GL_CUDA_CHECK(cudaEventSynchronize(ctx->marks[b]));
    GL_CUDA_CHECK(cudaEventElapsedTime(ms, ctx->marks[a], ctx->marks[b]));
    return GL_OK;
}

// Writes `bytes` of scratch on the context stream (benchmarks: evicts the L2 between timed iterations; 0 = twice the L2 size).
int gl_ctx_flush_l2(gl_ctx* ctx, size_t bytes)
{
    GL_REQUIRE(ctx, "gl_ctx_flush_l2: null");
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (bytes == 0) {
        int l2 = 0;
        GL_CUDA_CHECK(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, ctx->device));
        bytes = (size_t)2 * (size_t)(l2 > 0 ? l2 : (128 << 20));
    }
    if (ctx->flush_buf && ctx->flush_buf->bytes < bytes) { gl_buf_release(ctx->flush_buf); ctx->flush_buf = nullptr; }
    if (!ctx->flush_buf) GL_CHECK(gl_alloc(ctx, bytes, &ctx->flush_buf));
    GL_CUDA_CHECK(cudaMemsetAsync(ctx->flush_buf->ptr, 0x5a, bytes, ctx->stream));
    return GL_OK;
}

int gl_host_alloc(void** p, size_t bytes)
{
    GL_REQUIRE(p, "gl_host_alloc: null");
    GL_CUDA_CHECK(cudaMallocHost(p, bytes));
    return GL_OK;
}
int gl_host_free(void* p)
{
    if (p) GL_CUDA_CHECK(cudaFreeHost(p));
    return GL_OK;
}

}  // extern "C"

// -------------------------------------------------------------------------------------------
// allocator: blocks go back to a size-keyed cache; a request is served by the smallest cached
// block within 12.5 % of the size, so a steady-state pipeline never calls cudaMalloc/cudaFree.
// -------------------------------------------------------------------------------------------
int gl_alloc(gl_ctx* ctx, size_t bytes, gl_buf** out)
{
    *out = nullptr;
    if (bytes == 0) bytes = 256;
    bytes = (size_t)round_up((int64_t)bytes, 512);
    void* p = nullptr;
    size_t got = bytes;
    auto it = ctx->free_blocks.lower_bound(bytes);
    if (it != ctx->free_blocks.end() && it->first <= bytes + bytes / 8 + 4096) {
        p = it->second;
        got = it->first;
        ctx->bytes_cached -= got;
        ctx->free_blocks.erase(it);
    } else {
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            // drop the cache and retry once
            (void)cudaGetLastError();
            cudaStreamSynchronize(ctx->stream);
            for (auto& kv : ctx->free_blocks) cudaFree(kv.second);
            ctx->free_blocks.clear();
            ctx->bytes_cached = 0;
            e = cudaMalloc(&p, bytes);
        }
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            gl_set_error("device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
            return GL_ERR_NOMEM;
        }
    }
    gl_buf* b = new gl_buf();
    b->ptr = p;
    b->bytes = got;
    b->owner = ctx;
    ctx->bytes_live += got;
    if (ctx->bytes_live > ctx->bytes_peak) ctx->bytes_peak = ctx->bytes_live;
    *out = b;
    return GL_OK;
}

void gl_buf_release(gl_buf* b)
{
    if (!b) return;
    if (--b->refs > 0) return;
    gl_ctx* ctx = b->owner;
    // stream-ordered reuse: every consumer of this library runs on ctx->stream
    ctx->free_blocks.emplace(b->bytes, b->ptr);
    ctx->bytes_cached += b->bytes;
    ctx->bytes_live -= b->bytes;
    delete b;
}

int gl_image_ready(gl_ctx* ctx)
{
    if (!ctx->h2d_pending) return GL_OK;
    ctx->h2d_pending = false;
    GL_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d, 0));
    return GL_OK;
}

int gl_ensure_pinned(gl_ctx* ctx, size_t bytes)
{
    if (ctx->pinned_bytes >= bytes) return GL_OK;
    if (ctx->pinned) {
        cudaStreamSynchronize(ctx->stream);
        cudaFreeHost(ctx->pinned);
        ctx->pinned = nullptr;
        ctx->pinned_bytes = 0;
    }
    size_t want = (size_t)round_up((int64_t)bytes, 1 << 16);
    GL_CUDA_CHECK(cudaMallocHost(&ctx->pinned, want));
    ctx->pinned_bytes = want;
    return GL_OK;
}

int gl_status_begin(gl_ctx* ctx)
{
    GL_CUDA_CHECK(cudaMemsetAsync(ctx->dstat->ptr, 0, sizeof(int) * GL_DS_COUNT, ctx->stream));
    return GL_OK;
}

int gl_status_flush(gl_ctx* ctx)
{
    GL_CUDA_CHECK(cudaMemcpyAsync(ctx->hstat, ctx->dstat->ptr, sizeof(int) * GL_DS_COUNT, cudaMemcpyDeviceToHost, ctx->stream));
    GL_CUDA_CHECK(cudaEventRecord(ctx->stat_ev, ctx->stream));
    ctx->stat_pending = true;
    return GL_OK;
}

int gl_status_check(gl_ctx* ctx, bool* pt_overflow)
{
    if (pt_overflow) *pt_overflow = false;
    if (!ctx->stat_pending) return GL_OK;
    GL_CUDA_CHECK(cudaEventSynchronize(ctx->stat_ev));
    ctx->stat_pending = false;
    const int* h = ctx->hstat;
    if (h[GL_DS_SAMPLING] != 0) {
        gl_set_error("random sampling kernel failed (status %d)", h[GL_DS_SAMPLING]);
        return h[GL_DS_SAMPLING] == 1 ? GL_ERR_UNSUPPORTED : GL_ERR_CUDA;
    }
    if (h[GL_DS_PT_OVERFLOW] != 0) {
        // the sample lists needed more blocks than were set aside: nothing was computed; forget the capacity (the next run asks
        // the device first).  The caller that can, runs the step again.
        ctx->pt_cap_blocks = 0;
        if (pt_overflow) { *pt_overflow = true; return GL_OK; }
        gl_set_error("K_B sample lists outgrew their storage (%d blocks): run again", h[GL_DS_PT_BLOCKS]);
        return GL_ERR_NOMEM;
    }
    if (h[GL_DS_JACOBI] != 0) {
        float off;
        memcpy(&off, &h[GL_DS_JACOBI_OFF], sizeof(float));
        gl_set_error("eigensolve: not converged after %d sweeps (off-orthogonality %.3g > %.3g)", h[GL_DS_JACOBI_SWEEPS], off, ctx->jacobi_tol);
        return GL_ERR_NOTCONVERGED;
    }
    return GL_OK;
}

int gl_host_samples(gl_ctx* ctx)
{
    if (ctx->h_samples_valid) return GL_OK;
    GL_REQUIRE(ctx->samples && ctx->p > 0, "no samples");
    ctx->h_samples.resize(ctx->p);
    GL_CUDA_CHECK(cudaMemcpyAsync(ctx->h_samples.data(), ctx->samples->ptr, sizeof(uint32_t) * ctx->p, cudaMemcpyDeviceToHost, ctx->stream));
    GL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->h_samples_valid = true;
    return GL_OK;
}

gl_mat* gl_mat_new(gl_ctx* ctx, int kind)
{
    gl_mat* m = new gl_mat();
    m->kind = kind;
    m->ctx = ctx;
    return m;
}

// -------------------------------------------------------------------------------------------
// image
// -------------------------------------------------------------------------------------------
static int set_image_geometry(gl_ctx* ctx, int width, int height, int channels)
{
    GL_REQUIRE(width > 1 && height > 1, "image %dx%d too small", width, height);
    GL_REQUIRE(channels == 1 || channels == 3, "channels must be 1 or 3, got %d", channels);
    GL_REQUIRE((int64_t)width * height < (int64_t)0x7fffffff, "image too large for 32-bit raster indices");
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    const int64_t n = (int64_t)width * height;
    if (!ctx->img || ctx->img->bytes < (size_t)(n * channels)) {
        if (ctx->img) gl_buf_release(ctx->img);
        ctx->img = nullptr;
        GL_CHECK(gl_alloc(ctx, (size_t)(n * channels), &ctx->img));
    }
    ctx->image_epoch++;
    ctx->width = width;
    ctx->height = height;
    ctx->channels = channels;
    ctx->n = n;
    // contiguous bands of whole image rows
    ctx->row0 = (int)((int64_t)height * ctx->rank / ctx->world);
    ctx->row1 = (int)((int64_t)height * (ctx->rank + 1) / ctx->world);
    ctx->q0 = (int64_t)ctx->row0 * width;
    ctx->q1 = (int64_t)ctx->row1 * width;
    return GL_OK;
}

extern "C" {

int gl_set_image(gl_ctx* ctx, const uint8_t* pixels, int width, int height, int channels)
{
    GL_REQUIRE(ctx && pixels, "gl_set_image: null");
    GL_CHECK(set_image_geometry(ctx, width, height, channels));
    StageTimer t(ctx, GL_T_H2D);
    GL_CUDA_CHECK(cudaMemcpyAsync(ctx->img->ptr, pixels, (size_t)(ctx->n * channels), cudaMemcpyHostToDevice, ctx->stream));
    return GL_OK;
}

int gl_set_image_rows(gl_ctx* ctx, const uint8_t* const* rows, int width, int height)
{
    GL_REQUIRE(ctx && rows, "gl_set_image_rows: null");
    GL_CHECK(set_image_geometry(ctx, width, height, 1));
    GL_CHECK(gl_ensure_pinned(ctx, (size_t)ctx->n));
    // the pinned block may still be the source of an earlier async copy
    GL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r < height; ++r) memcpy((uint8_t*)ctx->pinned + (size_t)r * width, rows[r], (size_t)width);
    StageTimer t(ctx, GL_T_H2D);
    GL_CUDA_CHECK(cudaMemcpyAsync(ctx->img->ptr, ctx->pinned, (size_t)ctx->n, cudaMemcpyHostToDevice, ctx->stream));
    return GL_OK;
}

int gl_set_synthetic_image(gl_ctx* ctx, int width, int height, int channels, uint32_t seed)
{
    GL_REQUIRE(ctx, "gl_set_synthetic_image: null");
    GL_CHECK(set_image_geometry(ctx, width, height, channels));
    return gl_impl_synthetic(ctx, seed);
}

int gl_get_image(gl_ctx* ctx, uint8_t* out)
{
    GL_REQUIRE(ctx && out && ctx->img && ctx->n > 0, "gl_get_image: no image");
    GL_CUDA_CHECK(cudaMemcpyAsync(out, ctx->img->ptr, (size_t)(ctx->n * ctx->channels), cudaMemcpyDeviceToHost, ctx->stream));
    GL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return GL_OK;
}

int gl_get_band(gl_ctx* ctx, int* row0, int* row1)
{
    GL_REQUIRE(ctx && row0 && row1 && ctx->n > 0, "gl_get_band: no image");
    *row0 = ctx->row0;
    *row1 = ctx->row1;
    return GL_OK;
}

// -------------------------------------------------------------------------------------------
// samples
// -------------------------------------------------------------------------------------------
int gl_sampling_uniform(gl_ctx* ctx, unsigned requested, unsigned* actual)
{
    GL_REQUIRE(ctx && ctx->n > 0, "gl_sampling_uniform: set an image first");
    GL_REQUIRE(requested >= 1, "gl_sampling_uniform: requested must be >= 1");
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    StageTimer t(ctx, GL_T_SAMPLING);
    return gl_impl_sampling_uniform(ctx, requested, actual);
}

int gl_sampling_random(gl_ctx* ctx, unsigned requested, uint32_t seed, unsigned* actual)
{
    GL_REQUIRE(ctx && ctx->n > 0, "gl_sampling_random: set an image first");
    GL_REQUIRE(requested >= 1 && (int64_t)requested <= ctx->n, "gl_sampling_random: requested out of range");
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    StageTimer t(ctx, GL_T_SAMPLING);
    return gl_impl_sampling_random(ctx, requested, seed, actual);
}

int gl_set_samples(gl_ctx* ctx, const uint32_t* indices, unsigned count)
{
    GL_REQUIRE(ctx && indices && ctx->n > 0, "gl_set_samples: set an image first");
    GL_REQUIRE(count >= 2, "gl_set_samples: need at least 2 samples");
    for (unsigned i = 0; i < count; ++i) {
        GL_REQUIRE((int64_t)indices[i] < ctx->n, "gl_set_samples: index %u out of range", indices[i]);
        GL_REQUIRE(i == 0 || indices[i] > indices[i - 1], "gl_set_samples: indices must be strictly ascending");
    }
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    const int p_pad = (int)round_up(count, 64);
    if (ctx->samples) gl_buf_release(ctx->samples);
    ctx->samples = nullptr;
    GL_CHECK(gl_alloc(ctx, sizeof(uint32_t) * p_pad, &ctx->samples));
    GL_CHECK(gl_ensure_pinned(ctx, sizeof(uint32_t) * p_pad));
    GL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    uint32_t* h = (uint32_t*)ctx->pinned;
    memcpy(h, indices, sizeof(uint32_t) * count);
    for (int i = count; i < p_pad; ++i) h[i] = 0xffffffffu;
    GL_CUDA_CHECK(cudaMemcpyAsync(ctx->samples->ptr, h, sizeof(uint32_t) * p_pad, cudaMemcpyHostToDevice, ctx->stream));
    GL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->p = count;
    ctx->p_pad = p_pad;
    ctx->sample_epoch++;
    ctx->h_samples.assign(indices, indices + count);
    ctx->h_samples_valid = true;
    return GL_OK;
}

int gl_get_samples(gl_ctx* ctx, uint32_t* out, unsigned cap, unsigned* count)
{
    GL_REQUIRE(ctx && ctx->samples && ctx->p > 0, "gl_get_samples: no samples");
    if (count) *count = ctx->p;
    if (out) {
        unsigned k = cap < ctx->p ? cap : ctx->p;
        GL_CUDA_CHECK(cudaMemcpyAsync(out, ctx->samples->ptr, sizeof(uint32_t) * k, cudaMemcpyDeviceToHost, ctx->stream));
        GL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    return GL_OK;
}

// -------------------------------------------------------------------------------------------
// stage entry points (argument checks; the work is in the per-stage files)
// -------------------------------------------------------------------------------------------
int gl_affinity(gl_ctx* ctx, int kind, double h_loc, double h_val, gl_mat** K_A, gl_mat** K_B)
{
    GL_REQUIRE(ctx && K_A && K_B, "gl_affinity: null");
    GL_REQUIRE(ctx->n > 0 && ctx->p >= 2, "gl_affinity: need an image and samples first");
    GL_REQUIRE(kind >= GL_BILATERAL && kind <= GL_NLM, "gl_affinity: bad kind %d", kind);
    GL_REQUIRE(h_loc > 0 && h_val > 0, "gl_affinity: bandwidths must be positive");
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    StageTimer t(ctx, GL_T_AFFINITY);
    return gl_impl_affinity(ctx, kind, h_loc, h_val, K_A, K_B);
}

int gl_laplacian(gl_ctx* ctx, gl_mat* K_A, gl_mat* K_B, gl_mat** L_A, gl_mat** L_B)
{
    GL_REQUIRE(ctx && K_A && K_B && L_A && L_B, "gl_laplacian: null");
    GL_REQUIRE(K_A->kind == GL_MAT_KA && K_B->kind == GL_MAT_KB, "gl_laplacian: wrong handle kinds");
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    StageTimer t(ctx, GL_T_LAPLACIAN);
    return gl_impl_laplacian(ctx, K_A, K_B, L_A, L_B);
}

int gl_eigensolve(gl_ctx* ctx, gl_mat* L_A, int m, gl_mat** eigvecs, gl_mat** eigvals, gl_mat** eigvals_inv)
{
    GL_REQUIRE(ctx && L_A, "gl_eigensolve: null");
    GL_REQUIRE(L_A->kind == GL_MAT_KA && L_A->rows == L_A->cols, "gl_eigensolve: want a square p x p matrix");
    // m in [1, p]; the reference's driver never asks for more than p - 1 (hpc/image_processing.c:102-106, applied by the
    // callers), its Python prototype keeps all p (python/image_processing.py:287): both are served
    if (m < 0 || m > L_A->rows) m = (int)L_A->rows - 1;
    GL_REQUIRE(m >= 1, "gl_eigensolve: m must be >= 1");
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    StageTimer t(ctx, GL_T_EIGEN);
    return gl_impl_eigensolve(ctx, L_A, m, eigvecs, eigvals, eigvals_inv);
}

int gl_inverse_iteration(gl_ctx* ctx, gl_mat* L_A, int m, int opti_gs, double epsilon, int max_iterations, gl_mat** eigvecs, gl_mat** eigvals,
                         gl_mat** eigvals_inv, int* iterations_out, double* residual_out)
{
    GL_REQUIRE(ctx && L_A, "gl_inverse_iteration: null");
    GL_REQUIRE(L_A->kind == GL_MAT_KA && L_A->rows == L_A->cols, "gl_inverse_iteration: want a square p x p matrix");
    if (m < 0 || m > L_A->rows) m = (int)L_A->rows - 1;       // GetNumberEigenvalues, hpc/image_processing.c:96-108
    GL_REQUIRE(m >= 1, "gl_inverse_iteration: m must be >= 1");
    GL_REQUIRE(epsilon > 0.0 && max_iterations >= 1, "gl_inverse_iteration: epsilon must be positive, max_iterations >= 1");
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    StageTimer t(ctx, GL_T_EIGEN);
    return gl_impl_inverse_iteration(ctx, L_A, m, opti_gs, epsilon, max_iterations, eigvecs, eigvals, eigvals_inv, iterations_out,
                                     residual_out);
}

int gl_nystroem(gl_ctx* ctx, gl_mat* L_B, gl_mat* phi_A, gl_mat* eigvals_inv, gl_mat** phi)
{
    GL_REQUIRE(ctx && L_B && phi_A && eigvals_inv && phi, "gl_nystroem: null");
    GL_REQUIRE(L_B->kind == GL_MAT_KB && phi_A->kind == GL_MAT_EIGVEC && eigvals_inv->kind == GL_MAT_DIAG,
               "gl_nystroem: wrong handle kinds");
    GL_REQUIRE(phi_A->rows == L_B->p && eigvals_inv->rows == phi_A->cols, "gl_nystroem: shape mismatch");
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (!L_B->dscale) {
        // a bare K_B (the prototype's nystroem(K_A, K_B), python/image_processing.py:69-88): the factor is 1 instead of -alpha
        GL_CHECK(gl_alloc(ctx, 2 * sizeof(double), &L_B->dscale));
        GL_CHECK(gl_ensure_pinned(ctx, 2 * sizeof(double)));
        ((double*)ctx->pinned)[0] = L_B->scale;
        ((double*)ctx->pinned)[1] = -L_B->scale;
        GL_CUDA_CHECK(cudaMemcpyAsync(L_B->dscale->ptr, ctx->pinned, 2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        GL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    StageTimer t(ctx, GL_T_NYSTROEM);
    // Deferred by default: the handle is complete as far as the caller can tell, the matrix itself is computed when it is
    // first needed.  The reference's driver calls Nystroem and then ComputeResultFromLaplacian (hpc/image_processing.c:
    // 249-268): deferring lets gl_filter run both as ONE pass over Phi although they arrive as two calls.
    if (ctx->lazy_phi && ctx->gemm_impl == 0) return gl_phi_defer(ctx, L_B, phi_A, eigvals_inv, phi);
    return gl_impl_nystroem(ctx, L_B, phi_A, eigvals_inv, phi);
}

int gl_nystroem_filter(gl_ctx* ctx, gl_mat* L_B, gl_mat* phi_A, gl_mat* eigvals_inv, gl_mat* f_eigvals, double gain, int clip_low,
                       gl_mat** phi, float* z_f32, uint8_t* z_u8)
{
    GL_REQUIRE(ctx && L_B && phi_A && eigvals_inv && f_eigvals, "gl_nystroem_filter: null");
    GL_REQUIRE(L_B->kind == GL_MAT_KB && phi_A->kind == GL_MAT_EIGVEC && eigvals_inv->kind == GL_MAT_DIAG && f_eigvals->kind == GL_MAT_DIAG,
               "gl_nystroem_filter: wrong handle kinds");
    GL_REQUIRE(phi_A->rows == L_B->p && eigvals_inv->rows == phi_A->cols && f_eigvals->rows == phi_A->cols,
               "gl_nystroem_filter: shape mismatch");
    GL_REQUIRE(L_B->q0 == ctx->q0 && L_B->local_rows == ctx->q1 - ctx->q0 && L_B->image_epoch == ctx->image_epoch,
               "gl_nystroem_filter: L_B does not belong to the current image");
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    gl_fused_filter ff;
    ff.f_eigvals = f_eigvals;
    ff.gain = gain;
    ff.clip_low = clip_low;
    ff.z_f32 = z_f32;
    ff.z_u8 = z_u8;
    StageTimer t(ctx, GL_T_NYSTROEM);
    return gl_impl_nystroem(ctx, L_B, phi_A, eigvals_inv, phi, &ff);
}

int gl_orthonormalise(gl_ctx* ctx, gl_mat* phi, double* norms_out)
{
    GL_REQUIRE(ctx && phi && phi->kind == GL_MAT_PHI, "gl_orthonormalise: want a Phi handle");
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (phi->def_LB) {
        StageTimer tn(ctx, GL_T_NYSTROEM);
        GL_CHECK(gl_phi_materialise(ctx, phi));
    }
    StageTimer t(ctx, GL_T_GRAM_SCHMIDT);
    return gl_impl_orthonormalise(ctx, phi, norms_out);
}

int gl_filter(gl_ctx* ctx, gl_mat* phi, gl_mat* f_eigvals, double gain, int clip_low, float* z_f32, uint8_t* z_u8)
{
    GL_REQUIRE(ctx && phi && f_eigvals, "gl_filter: null");
    GL_REQUIRE(phi->kind == GL_MAT_PHI && f_eigvals->kind == GL_MAT_DIAG, "gl_filter: wrong handle kinds");
    GL_REQUIRE(f_eigvals->rows == phi->m, "gl_filter: %lld eigenvalues for %d columns", (long long)f_eigvals->rows, phi->m);
    GL_REQUIRE(phi->q0 == ctx->q0 && phi->local_rows == ctx->q1 - ctx->q0, "gl_filter: Phi does not belong to the current image");
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (phi->def_LB) {
        // Phi was deferred by gl_nystroem: compute it now -- with the filter riding on the GEMM's epilogue when that is possible
        const gl_mat* LB = phi->def_LB;
        const bool fuse = ctx->fuse_filter && ctx->projection_mode == 0 && ctx->gemm_impl == 0 && LB->aux &&
                          LB->channels == ctx->channels && LB->image_epoch == ctx->image_epoch;
        gl_fused_filter ff;
        ff.f_eigvals = f_eigvals;
        ff.gain = gain;
        ff.clip_low = clip_low;
        ff.z_f32 = z_f32;
        ff.z_u8 = z_u8;
        {
            StageTimer t(ctx, GL_T_NYSTROEM);
            if (fuse) ctx->ev_valid[GL_T_FILTER] = false;
            GL_CHECK(gl_phi_materialise(ctx, phi, fuse ? &ff : nullptr));
        }
        if (fuse) return GL_OK;
    }
    return gl_impl_filter(ctx, phi, f_eigvals, gain, clip_low, z_f32, z_u8);
}

// ---- the prototype's experimental blocks (proto.cu) ----------------------------------------------------------
int gl_sinkhorn(gl_ctx* ctx, gl_mat* phi, gl_mat* Pi, int iterations, gl_mat** W_A, gl_mat** W_ABt)
{
    GL_REQUIRE(ctx && phi && Pi, "gl_sinkhorn: null");
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    return gl_impl_sinkhorn(ctx, phi, Pi, iterations, W_A, W_ABt);
}

int gl_orthogonalisation(gl_ctx* ctx, gl_mat* K_A, gl_mat* K_B, gl_mat** V, gl_mat** Pi)
{
    GL_REQUIRE(ctx && K_A && K_B, "gl_orthogonalisation: null");
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    return gl_impl_orthogonalisation(ctx, K_A, K_B, V, Pi);
}

int gl_smoothing_matrix(gl_ctx* ctx, gl_mat* phi, gl_mat* Pi, gl_mat** V, gl_mat** L)
{
    GL_REQUIRE(ctx && phi && Pi, "gl_smoothing_matrix: null");
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    return gl_impl_smoothing_matrix(ctx, phi, Pi, V, L);
}

int gl_matrix_filter(gl_ctx* ctx, gl_mat* V, gl_mat* L, const double* coef, int ncoef, float* z_f32)
{
    GL_REQUIRE(ctx && V && L, "gl_matrix_filter: null");
    GL_REQUIRE(V->q0 == ctx->q0 && V->local_rows == ctx->q1 - ctx->q0, "gl_matrix_filter: V does not belong to the current image");
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    StageTimer t(ctx, GL_T_FILTER);
    return gl_impl_matrix_filter(ctx, V, L, coef, ncoef, z_f32);
}

int gl_full_affinity(gl_ctx* ctx, int kind, double h_loc, double h_val, gl_mat** K)
{
    GL_REQUIRE(ctx && K, "gl_full_affinity: null");
    GL_REQUIRE(ctx->n > 0, "gl_full_affinity: set an image first");
    GL_REQUIRE(kind >= GL_BILATERAL && kind <= GL_SPATIAL, "gl_full_affinity: bad kind %d (the matrix-free full mode has the three kinds of hpc/affinity.c)", kind);
    GL_REQUIRE(h_loc > 0 && h_val > 0, "gl_full_affinity: bandwidths must be positive");
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    StageTimer t(ctx, GL_T_AFFINITY);
    return gl_impl_full_affinity(ctx, kind, h_loc, h_val, K);
}

int gl_full_laplacian(gl_ctx* ctx, gl_mat* K, gl_mat** L)
{
    GL_REQUIRE(ctx && K && L && K->kind == GL_MAT_FULL && !K->dscale, "gl_full_laplacian: want the handle gl_full_affinity returned");
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    StageTimer t(ctx, GL_T_LAPLACIAN);
    return gl_impl_full_laplacian(ctx, K, L);
}

int gl_full_result(gl_ctx* ctx, gl_mat* L, float* z_f32, uint8_t* z_u8)
{
    GL_REQUIRE(ctx && L && L->kind == GL_MAT_FULL && L->dscale, "gl_full_result: want the handle gl_full_laplacian returned");
    GL_REQUIRE(L->image_epoch == ctx->image_epoch && L->q0 == ctx->q0, "gl_full_result: the matrix does not belong to the current image");
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    StageTimer t(ctx, GL_T_FILTER);
    return gl_impl_full_result(ctx, L, z_f32, z_u8);
}

int gl_diag_inverse(gl_ctx* ctx, gl_mat* d, gl_mat** out)
{
    GL_REQUIRE(ctx && d && out && d->kind == GL_MAT_DIAG, "gl_diag_inverse: want a diagonal handle");
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    return gl_impl_diag_map(ctx, d, 0, 0.0, out);
}

int gl_diag_pow(gl_ctx* ctx, gl_mat* d, double power, gl_mat** out)
{
    GL_REQUIRE(ctx && d && out && d->kind == GL_MAT_DIAG, "gl_diag_pow: want a diagonal handle");
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    return gl_impl_diag_map(ctx, d, 1, power, out);
}

// -------------------------------------------------------------------------------------------
// matrices
// -------------------------------------------------------------------------------------------
int gl_mat_info_get(const gl_mat* m, gl_mat_info* info)
{
    GL_REQUIRE(m && info, "gl_mat_info_get: null");
    info->kind = m->kind;
    info->rows = m->rows;
    info->cols = m->cols;
    info->local_rows = m->local_rows;
    info->ld = m->ld;
    info->elem_bytes = m->elem_bytes;
    if (!m->scale_on_host && m->dscale) {
        gl_mat* mm = const_cast<gl_mat*>(m);
        GL_CUDA_CHECK(cudaSetDevice(m->ctx->device));
        GL_CUDA_CHECK(cudaMemcpyAsync(&mm->scale, m->dscale->ptr, sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream));
        GL_CUDA_CHECK(cudaStreamSynchronize(m->ctx->stream));
        mm->scale_on_host = true;
    }
    info->scale = m->scale;
    info->stored_blocks = m->total_blocks;
    info->layout = m->kind == GL_MAT_KB && m->pt_buf ? 1 : 0;
    // (pixel, sample slot) pairs held, padding included: the affinity kernel evaluates and the extrapolation multiplies these
    info->stored_pairs = m->kind != GL_MAT_KB ? 0 : (m->pt_buf ? m->pt_blocks * 8 * 128 * 32 : m->total_blocks * 512 * (int64_t)m->kbs);
    // pairs the extrapolation multiplies (x 2 m_pad flop each): the patch kernel skips the empty upper half of a last block
    info->mma_pairs = m->kind != GL_MAT_KB ? 0 : (m->pt_buf ? m->pt_ksteps * 128 * 16 : info->stored_pairs);
    return GL_OK;
}

int gl_mat_retain(gl_mat* m)
{
    GL_REQUIRE(m, "gl_mat_retain: null");
    m->refs++;
    return GL_OK;
}

int gl_mat_destroy(gl_mat* m)
{
    if (!m) return GL_OK;
    if (--m->refs > 0) return GL_OK;
    if (m->buf) gl_buf_release(m->buf);
    if (m->aux) gl_buf_release(m->aux);
    if (m->tiles) gl_buf_release(m->tiles);
    if (m->starts) gl_buf_release(m->starts);
    if (m->perm) gl_buf_release(m->perm);
    if (m->pt_info) gl_buf_release(m->pt_info);
    if (m->pt_slots) gl_buf_release(m->pt_slots);
    if (m->pt_buf) gl_buf_release(m->pt_buf);
    if (m->dscale) gl_buf_release(m->dscale);
    if (m->proj) gl_buf_release(m->proj);
    if (m->def_LB) gl_mat_destroy(m->def_LB);
    if (m->def_U) gl_mat_destroy(m->def_U);
    if (m->def_muinv) gl_mat_destroy(m->def_muinv);
    delete m;
    return GL_OK;
}

}  // extern "C"

// conversion kernels for gl_mat_download / gl_mat_upload
__global__ void k_half_to_f64(const __half* __restrict__ src, int64_t ld, int64_t rows, int64_t cols, double scale,
                              double* __restrict__ dst)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    int64_t r = i / cols, c = i % cols;
    dst[i] = scale * (double)__half2float(src[r * ld + c]);
}
// K_B from its blocked storage ([block][512 pixels][64 sample slots], affinity.cu) to dense fp64 rows x cols in the caller's
// sample order; sample slots that no stored block of the tile covers (entries below the fp16 flush-to-zero cutoff) read as 0
__global__ void k_kb_blocked_to_f64(const __half* __restrict__ src, const int4* __restrict__ tab, const int* __restrict__ starts,
                                    const uint32_t* __restrict__ perm, int kbs, int p_int, int64_t rows, int64_t cols, double scale,
                                    double* __restrict__ dst)
{
    // one thread per (pixel row, internal slot): the slot's sample decides the output column
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * p_int) return;
    const int64_t r = i / p_int;
    const int slot = (int)(i % p_int);
    const uint32_t c = perm[slot];
    if (c == 0xffffffffu || (int64_t)c >= cols) return;
    const int4 tl = tab[r >> 9];
    double v = 0.0;
    for (int k = 0; k < tl.y; ++k) {
        const int sb = starts[tl.x + k];
        if (slot >= sb && slot < sb + kbs) {
            v = (double)__half2float(src[(((size_t)tl.z + k) * 512 + (size_t)(r & 511)) * kbs + (slot - sb)]);
            break;
        }
    }
    dst[r * cols + c] = scale * v;
}
__global__ void k_colmajor_f32_to_f64(const float* __restrict__ src, int64_t ld, int64_t rows, int64_t cols,
                                      double* __restrict__ dst)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    int64_t r = i / cols, c = i % cols;
    dst[i] = (double)src[c * ld + r];
}
__global__ void k_f64_to_half_padded(const double* __restrict__ src, int64_t rows, int64_t cols, int64_t ld, __half* __restrict__ dst)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * ld) return;
    int64_t r = i / ld, c = i % ld;
    dst[i] = __float2half_rn(c < cols ? (float)src[r * cols + c] : 0.f);
}
__global__ void k_f64_to_colmajor_f32(const double* __restrict__ src, int64_t rows, int64_t cols, int64_t ld,
                                      float* __restrict__ dst)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    int64_t r = i / cols, c = i % cols;
    dst[c * ld + r] = (float)src[i];
}
// dst[r][c] = scale * src(r, col0 + c) for a strip of columns; layout: 0 fp16 row-major (Phi), 1 fp32 column-major (eigenvectors),
// 2 fp64 row-major (K_A / L_A)
__global__ void k_col_strip_to_f64(const void* __restrict__ src, int layout, int64_t ld, int64_t rows, int col0, int ncols, double scale,
                                   double* __restrict__ dst)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * ncols) return;
    const int64_t r = i / ncols, c = col0 + i % ncols;
    double v;
    if (layout == 0) v = (double)__half2float(((const __half*)src)[r * ld + c]);
    else if (layout == 1) v = (double)((const float*)src)[c * ld + r];
    else v = ((const double*)src)[r * ld + c];
    dst[i] = scale * v;
}
__global__ void k_scale_f64(const double* __restrict__ src, int64_t count, double scale, double* __restrict__ dst)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) dst[i] = scale * src[i];
}

extern "C" {

int gl_mat_download(gl_ctx* ctx, const gl_mat* m, double* out, size_t cap)
{
    GL_REQUIRE(ctx && m && out, "gl_mat_download: null");
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (m->kind == GL_MAT_PHI && m->def_LB) GL_CHECK(gl_phi_materialise(ctx, const_cast<gl_mat*>(m)));
    int64_t rows = m->rows, cols = m->cols;
    if (m->kind == GL_MAT_KB) { rows = m->local_rows; cols = m->p; }
    if (m->kind == GL_MAT_PHI) { rows = m->local_rows; cols = m->m; }
    if (m->kind == GL_MAT_DIAG) cols = 1;
    const int64_t count = rows * cols;
    GL_REQUIRE((size_t)count <= cap, "gl_mat_download: need room for %lld doubles, got %zu", (long long)count, cap);
    if (!m->scale_on_host) {
        gl_mat_info tmp_info;
        GL_CHECK(gl_mat_info_get(m, &tmp_info));
    }
    gl_buf* tmp = nullptr;
    GL_CHECK(gl_alloc(ctx, sizeof(double) * (size_t)count, &tmp));
    const int T = 256;
    const unsigned blocks = (unsigned)ceil_div(count, T);
    double* d = (double*)tmp->ptr;
    switch (m->kind) {
    case GL_MAT_KA:
    case GL_MAT_DIAG:
        k_scale_f64<<<blocks, T, 0, ctx->stream>>>((const double*)m->buf->ptr, count, m->scale, d);
        break;
    case GL_MAT_KB: {
        if (m->pt_buf) {   // patch layout (patch.cu)
            const int rcp = gl_patch_download(ctx, m, m->scale, d);
            if (rcp != GL_OK) { gl_buf_release(tmp); return rcp; }
            ctx->launches--;   // (counted again by the check below)
            break;
        }
        const int p_int = m->p_pad + 64;
        k_kb_blocked_to_f64<<<(unsigned)ceil_div(rows * p_int, T), T, 0, ctx->stream>>>(
            (const __half*)m->buf->ptr, (const int4*)m->tiles->ptr, (const int*)m->starts->ptr, (const uint32_t*)m->perm->ptr, m->kbs, p_int,
            rows, cols, m->scale, d);
        break;
    }
    case GL_MAT_PHI:
        k_half_to_f64<<<blocks, T, 0, ctx->stream>>>((const __half*)m->buf->ptr, m->ld, rows, cols, m->scale, d);
        break;
    case GL_MAT_EIGVEC:
        k_colmajor_f32_to_f64<<<blocks, T, 0, ctx->stream>>>((const float*)m->buf->ptr, m->ld, rows, cols, d);
        break;
    default:
        gl_buf_release(tmp);
        GL_REQUIRE(false, "gl_mat_download: bad kind %d", m->kind);
    }
    GL_LAUNCH_CHECK(ctx);
    GL_CUDA_CHECK(cudaMemcpyAsync(out, d, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
    GL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    gl_buf_release(tmp);
    return GL_OK;
}

int gl_mat_download_cols(gl_ctx* ctx, const gl_mat* m, int col0, int ncols, double* out, size_t cap)
{
    GL_REQUIRE(ctx && m && out, "gl_mat_download_cols: null");
    GL_REQUIRE(m->kind == GL_MAT_PHI || m->kind == GL_MAT_EIGVEC || m->kind == GL_MAT_KA,
               "gl_mat_download_cols: want a Phi, eigenvector or p x p matrix, got kind %d", m->kind);
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (m->kind == GL_MAT_PHI && m->def_LB) GL_CHECK(gl_phi_materialise(ctx, const_cast<gl_mat*>(m)));
    const int64_t rows = m->kind == GL_MAT_PHI ? m->local_rows : m->rows;
    const int64_t cols = m->kind == GL_MAT_PHI ? m->m : m->cols;
    GL_REQUIRE(col0 >= 0 && ncols >= 1 && (int64_t)col0 + ncols <= cols, "gl_mat_download_cols: columns [%d, %d) of %lld", col0,
               col0 + ncols, (long long)cols);
    const int64_t count = rows * ncols;
    GL_REQUIRE((size_t)count <= cap, "gl_mat_download_cols: need room for %lld doubles, got %zu", (long long)count, cap);
    if (!m->scale_on_host) {
        gl_mat_info tmp_info;
        GL_CHECK(gl_mat_info_get(m, &tmp_info));
    }
    gl_buf* tmp = nullptr;
    GL_CHECK(gl_alloc(ctx, sizeof(double) * (size_t)count, &tmp));
    const int layout = m->kind == GL_MAT_PHI ? 0 : (m->kind == GL_MAT_EIGVEC ? 1 : 2);
    k_col_strip_to_f64<<<(unsigned)ceil_div(count, 256), 256, 0, ctx->stream>>>(m->buf->ptr, layout, m->ld, rows, col0, ncols,
                                                                               layout == 1 ? 1.0 : m->scale, (double*)tmp->ptr);
    GL_LAUNCH_CHECK(ctx);
    GL_CUDA_CHECK(cudaMemcpyAsync(out, tmp->ptr, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
    GL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    gl_buf_release(tmp);
    return GL_OK;
}

int gl_mat_rowsums(gl_ctx* ctx, const gl_mat* K_B, double* out, size_t cap)
{
    GL_REQUIRE(ctx && K_B && out && K_B->kind == GL_MAT_KB && K_B->aux, "gl_mat_rowsums: want a K_B handle");
    GL_REQUIRE(cap >= (size_t)K_B->p, "gl_mat_rowsums: need room for %d doubles", K_B->p);
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    GL_CUDA_CHECK(cudaMemcpyAsync(out, K_B->aux->ptr, sizeof(double) * K_B->p, cudaMemcpyDeviceToHost, ctx->stream));
    GL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return GL_OK;
}

int gl_mat_upload(gl_ctx* ctx, int kind, const double* data, int64_t rows, int64_t cols, gl_mat** out)
{
    GL_REQUIRE(ctx && data && out, "gl_mat_upload: null");
    GL_REQUIRE(rows > 0 && cols > 0, "gl_mat_upload: empty");
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    const int64_t count = rows * cols;
    gl_buf* stage = nullptr;
    GL_CHECK(gl_alloc(ctx, sizeof(double) * (size_t)count, &stage));
    GL_CUDA_CHECK(cudaMemcpyAsync(stage->ptr, data, sizeof(double) * (size_t)count, cudaMemcpyHostToDevice, ctx->stream));
    GL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // `data` may be pageable and freed by the caller
    gl_mat* m = gl_mat_new(ctx, kind);
    m->rows = rows;
    m->cols = cols;
    m->local_rows = rows;
    if (kind == GL_MAT_KA) {
        GL_REQUIRE(rows == cols, "gl_mat_upload: K_A/L_A must be square");
        m->buf = stage;
        m->ld = cols;
        m->elem_bytes = 8;
    } else if (kind == GL_MAT_DIAG) {
        GL_REQUIRE(cols == 1, "gl_mat_upload: diagonal is rows x 1");
        m->buf = stage;
        m->ld = 1;
        m->elem_bytes = 8;
    } else if (kind == GL_MAT_EIGVEC) {
        m->ld = round_up(rows, 4);
        m->elem_bytes = 4;
        int rc = gl_alloc(ctx, sizeof(float) * (size_t)(m->ld * cols), &m->buf);
        if (rc != GL_OK) { gl_buf_release(stage); delete m; return rc; }
        cudaMemsetAsync(m->buf->ptr, 0, sizeof(float) * (size_t)(m->ld * cols), ctx->stream);
        k_f64_to_colmajor_f32<<<(unsigned)ceil_div(count, 256), 256, 0, ctx->stream>>>((const double*)stage->ptr, rows, cols,
                                                                                      m->ld, (float*)m->buf->ptr);
        GL_LAUNCH_CHECK(ctx);
        gl_buf_release(stage);
    } else if (kind == GL_MAT_PHI) {
        // this rank's band rows x cols of a Phi in raster order (tests, host-built bases for gl_filter / the prototype blocks)
        const int64_t band = ctx->q1 - ctx->q0;
        if (rows != band || !ctx->samples) {
            gl_buf_release(stage);
            delete m;
            GL_REQUIRE(false, "gl_mat_upload: a Phi needs the %lld rows of this rank's band (got %lld) and a sample set", (long long)band,
                       (long long)rows);
        }
        const int m_pad = gl_m_pad((int)cols);
        m->rows = ctx->n;
        m->ld = m_pad;
        m->elem_bytes = 2;
        m->p = (int)ctx->p; m->p_pad = ctx->p_pad; m->m = (int)cols; m->m_pad = m_pad; m->q0 = ctx->q0;
        int rc = gl_alloc(ctx, sizeof(__half) * (size_t)rows * m_pad, &m->buf);
        if (rc != GL_OK) { gl_buf_release(stage); delete m; return rc; }
        k_f64_to_half_padded<<<(unsigned)ceil_div(rows * m_pad, 256), 256, 0, ctx->stream>>>((const double*)stage->ptr, rows, cols, m_pad,
                                                                                           (__half*)m->buf->ptr);
        GL_LAUNCH_CHECK(ctx);
        gl_buf_release(stage);
    } else {
        gl_buf_release(stage);
        delete m;
        GL_REQUIRE(false, "gl_mat_upload: kind %d cannot be uploaded", kind);
    }
    *out = m;
    return GL_OK;
}

// -------------------------------------------------------------------------------------------
// whole path: the stage order of hpc/image_processing.c:183-275 (restored block)
// -------------------------------------------------------------------------------------------
}  // extern "C"

__global__ void k_scatter_sample_pixels(uint8_t* __restrict__ img, const uint32_t* __restrict__ samples, const uint8_t* __restrict__ vals,
                                        int p, int C)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p * C) return;
    img[(size_t)samples[i / C] * C + (i % C)] = vals[i];
}

__global__ void k_gather_sample_pixels(const uint8_t* __restrict__ img, const uint32_t* __restrict__ samples, int p, int C, int64_t q0, int64_t q1,
                                       float* __restrict__ vals)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p * C) return;
    const int64_t q = samples[i / C];
    vals[i] = (q >= q0 && q < q1) ? (float)img[(size_t)q * C + (i % C)] : 0.f;
}
__global__ void k_scatter_sample_pixels_f32(uint8_t* __restrict__ img, const uint32_t* __restrict__ samples, const float* __restrict__ vals, int p, int C)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p * C) return;
    img[(size_t)samples[i / C] * C + (i % C)] = (uint8_t)vals[i];
}

// multi-GPU gl_run: every rank has uploaded its own band of rows only; the p sampled pixels lie in all bands.  Each rank contributes the
// values of the samples it holds (zero for the others) to ONE small allreduce and writes the sums into its image -- no host in between
// (the earlier version fetched the indices, gathered the values from the caller's host image and uploaded them: two round trips).
static int exchange_sample_pixels(gl_ctx* ctx)
{
    const int p = (int)ctx->p, C = ctx->channels;
    gl_buf* vals = nullptr;
    GL_CHECK(gl_alloc(ctx, sizeof(float) * (size_t)p * C, &vals));
    const unsigned blocks = (unsigned)ceil_div(p * C, 256);
    k_gather_sample_pixels<<<blocks, 256, 0, ctx->stream>>>((const uint8_t*)ctx->img->ptr, (const uint32_t*)ctx->samples->ptr, p, C, ctx->q0, ctx->q1,
                                                            (float*)vals->ptr);
    ctx->launches++;
    int rc = gl_allreduce_f32(ctx, (float*)vals->ptr, (size_t)p * C);
    if (rc == GL_OK) {
        k_scatter_sample_pixels_f32<<<blocks, 256, 0, ctx->stream>>>((uint8_t*)ctx->img->ptr, (const uint32_t*)ctx->samples->ptr, (const float*)vals->ptr, p, C);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) { gl_set_error("gl_run: exchanging the sample pixels failed"); rc = GL_ERR_CUDA; }
    }
    gl_buf_release(vals);
    return rc;
}

// (the same through the host, kept for a context whose communicator is not up: the values of the p sampled pixels, gathered from the
// caller's host image, placed into ctx->img)
static int upload_sample_pixels(gl_ctx* ctx)
{
    const int p = (int)ctx->p, C = ctx->channels;
    GL_CHECK(gl_host_samples(ctx));
    GL_REQUIRE(ctx->h_samples_valid && (int)ctx->h_samples.size() == p, "gl_run: no host copy of the sample indices");
    gl_buf* vals = nullptr;
    GL_CHECK(gl_alloc(ctx, (size_t)p * C, &vals));
    int rc = gl_ensure_pinned(ctx, (size_t)p * C);
    if (rc == GL_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = GL_ERR_CUDA;   // the pinned block may be in flight
    if (rc == GL_OK) {
        uint8_t* h = (uint8_t*)ctx->pinned;
        for (int i = 0; i < p; ++i)
            for (int ch = 0; ch < C; ++ch) h[(size_t)i * C + ch] = ctx->host_pixels[(size_t)ctx->h_samples[i] * C + ch];
        StageTimer t(ctx, GL_T_H2D);
        if (cudaMemcpyAsync(vals->ptr, h, (size_t)p * C, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = GL_ERR_CUDA;
    }
    if (rc == GL_OK) {
        k_scatter_sample_pixels<<<(unsigned)ceil_div(p * C, 256), 256, 0, ctx->stream>>>((uint8_t*)ctx->img->ptr, (const uint32_t*)ctx->samples->ptr,
                                                                                         (const uint8_t*)vals->ptr, p, C);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) rc = GL_ERR_CUDA;
    }
    if (rc == GL_ERR_CUDA) gl_set_error("gl_run: uploading the sample pixels failed");
    gl_buf_release(vals);
    return rc;
}

extern "C" {

}  // extern "C"

// One pass of the whole path.  The stages run in "asynchronous mode": none of them stops for the host (sampling status, size of
// the K_B sample lists, convergence report of the eigensolve go to the deferred status block), so the host enqueues the ~30
// launches of a step ahead of the device.
static int run_resident_once(gl_ctx* ctx, const gl_params* prm, float* z_f32, uint8_t* z_u8, unsigned* p_out, int* m_out,
                             double* eigvals_out, size_t eigvals_cap)
{

    unsigned requested = prm->sample_size ? prm->sample_size : (unsigned)((double)ctx->n * 0.01);  // image_processing.c:187
    unsigned p = 0;
    if (prm->sampling_random) GL_CHECK(gl_sampling_random(ctx, requested, prm->seed, &p));
    else GL_CHECK(gl_sampling_uniform(ctx, requested, &p));
    if (ctx->host_pixels) {
        GL_CHECK(gl_image_ready(ctx));      // the band upload (copy stream) before anything touches ctx->img
        GL_CHECK(ctx->comm ? exchange_sample_pixels(ctx) : upload_sample_pixels(ctx));
    }

    gl_mat *K_A = nullptr, *K_B = nullptr, *L_A = nullptr, *L_B = nullptr;
    gl_mat *U = nullptr, *mu = nullptr, *mu_inv = nullptr, *phi = nullptr, *f_mu = nullptr;
    int rc = GL_OK;
    do {
        // the patch layout of K_B serves the fused extrapolation + filter only; a run that needs Phi itself (orthonormalisation,
        // stages apart, Phi kept) asks for the blocked layout right away instead of computing it on demand later
        const bool will_fuse = ctx->fuse_filter && !prm->gram_schmidt && ctx->projection_mode == 0 && ctx->gemm_impl == 0 && !ctx->keep_phi;
        ctx->want_blocked = !will_fuse;
        rc = gl_affinity(ctx, prm->affinity_kind, prm->h_loc, prm->h_val, &K_A, &K_B);
        ctx->want_blocked = false;
        if (rc != GL_OK) break;
        if ((rc = gl_laplacian(ctx, K_A, K_B, &L_A, &L_B)) != GL_OK) break;
        gl_mat_destroy(K_A); K_A = nullptr;            // image_processing.c:210-211
        gl_mat_destroy(K_B); K_B = nullptr;
        // GetNumberEigenvalues (hpc/image_processing.c:96-108): absent, negative or >= p means p - 1
        const int m_req = (prm->num_eigvals < 0 || prm->num_eigvals >= (int)p) ? (int)p - 1 : prm->num_eigvals;
        if ((rc = gl_eigensolve(ctx, L_A, m_req, &U, &mu, &mu_inv)) != GL_OK) break;
        gl_mat_destroy(L_A); L_A = nullptr;
        if ((rc = gl_diag_pow(ctx, mu, prm->power, &f_mu)) != GL_OK) break;  // MatPow(eigvals, .) as intended
        // without an orthonormalisation in between, the filter can ride on the extrapolation GEMM's epilogue: Phi is
        // still written, but never read back (option fuse_filter=0 runs the two stages apart, as the reference does)
        const bool fused = ctx->fuse_filter && !prm->gram_schmidt && ctx->projection_mode == 0 && ctx->gemm_impl == 0;
        if (fused) {
            ctx->ev_valid[GL_T_FILTER] = false;
            // Phi is only a temporary of this one-call path (no argument returns it): its tiles are consumed in the epilogue
            // and never written to HBM, unless option keep_phi=1 asks for the reference's data flow
            // A Phi that does not fit (config 5 on one GPU: 67 M pixels x 2048 columns = 275 GB) is not stored either:
            // the fused pass needs K_B and the row partials only.  phi_limit_mb forces that path (tests).
            bool keep = ctx->keep_phi;
            const size_t phi_bytes = (size_t)L_B->local_rows * (size_t)gl_m_pad((int)U->cols) * 2;
            if (keep && ctx->phi_limit_mb > 0 && phi_bytes > (size_t)ctx->phi_limit_mb << 20) keep = false;
            if (keep && ctx->phi_nomem_bytes && phi_bytes >= ctx->phi_nomem_bytes) keep = false;
            rc = gl_nystroem_filter(ctx, L_B, U, mu_inv, f_mu, prm->gain, prm->clip_low, keep ? &phi : nullptr, z_f32, z_u8);
            if (rc == GL_ERR_NOMEM && keep) {
                if (ctx->verbose) fprintf(stderr, "[glb200] Phi (%.1f GB) does not fit: consumed in the GEMM epilogue, not stored\n", phi_bytes / 1e9);
                keep = false;
                ctx->phi_nomem_bytes = phi_bytes;
                rc = gl_nystroem_filter(ctx, L_B, U, mu_inv, f_mu, prm->gain, prm->clip_low, nullptr, z_f32, z_u8);
            }
            ctx->last_phi_stored = keep;
            if (rc != GL_OK) break;
        } else if ((rc = gl_nystroem(ctx, L_B, U, mu_inv, &phi)) != GL_OK) break;
        gl_mat_destroy(L_B); L_B = nullptr;
        gl_mat_destroy(U); U = nullptr;
        gl_mat_destroy(mu_inv); mu_inv = nullptr;
        if (!fused) {
            if (prm->gram_schmidt && (rc = gl_orthonormalise(ctx, phi, nullptr)) != GL_OK) break;
            if ((rc = gl_filter(ctx, phi, f_mu, prm->gain, prm->clip_low, z_f32, z_u8)) != GL_OK) break;
        }
        if (p_out) *p_out = p;
        if (m_out) *m_out = (int)mu->rows;
        // uniform sampling may return up to ~4x the requested count (hpc/sampling.c:8-13), so the caller's buffer is checked
        if (eigvals_out) rc = gl_mat_download(ctx, mu, eigvals_out, eigvals_cap);
    } while (0);
    gl_mat_destroy(K_A); gl_mat_destroy(K_B); gl_mat_destroy(L_A); gl_mat_destroy(L_B);
    gl_mat_destroy(U); gl_mat_destroy(mu); gl_mat_destroy(mu_inv); gl_mat_destroy(phi); gl_mat_destroy(f_mu);
    return rc;
}

extern "C" {

int gl_run_resident(gl_ctx* ctx, const gl_params* prm, float* z_f32, uint8_t* z_u8, unsigned* p_out, int* m_out,
                    double* eigvals_out, size_t eigvals_cap)
{
    GL_REQUIRE(ctx && prm, "gl_run_resident: null");
    GL_REQUIRE(ctx->n > 0, "gl_run_resident: no image on the device");
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (!ctx->total_started) cudaEventRecord(ctx->ev_begin[GL_T_TOTAL], ctx->stream);
    ctx->total_started = false;
    // what the previous asynchronous run had to report (normally already there: this does not wait for the device)
    bool overflow = false;
    GL_CHECK(gl_status_check(ctx, &overflow));
    int rc = GL_OK;
    for (int attempt = 0; attempt < 2; ++attempt) {
        GL_CHECK(gl_status_begin(ctx));
        ctx->async_mode = true;
        rc = run_resident_once(ctx, prm, z_f32, z_u8, p_out, m_out, eigvals_out, eigvals_cap);
        ctx->async_mode = false;
        if (rc != GL_OK) break;
        GL_CHECK(gl_status_flush(ctx));
        if (!(z_f32 || z_u8 || eigvals_out)) break;     // nothing comes back to the host: the status is read at the next call
        // results were copied to the host: they are valid only if the deferred status says so
        rc = gl_status_check(ctx, &overflow);
        if (rc != GL_OK || !overflow) break;
        // (the K_B sample lists outgrew the storage set aside from the previous run: once more, asking the device first)
    }
    cudaEventRecord(ctx->ev_end[GL_T_TOTAL], ctx->stream);
    ctx->ev_valid[GL_T_TOTAL] = true;
    if (rc == GL_OK && (z_f32 || z_u8)) GL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return rc;
}

int gl_run(gl_ctx* ctx, const uint8_t* pixels, int width, int height, int channels, const gl_params* prm, float* z_f32,
           uint8_t* z_u8, unsigned* p_out, int* m_out, double* eigvals_out, size_t eigvals_cap)
{
    GL_REQUIRE(ctx && pixels, "gl_run: null");
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaEventRecord(ctx->ev_begin[GL_T_TOTAL], ctx->stream);
    if (ctx->world == 1 || prm->affinity_kind == GL_NLM) {
        // (NLM reads 7x7 patches around every band pixel and every sample, wherever it lies: every rank takes the whole image)
        // The upload goes on the copy stream, behind whatever the context stream still has in flight on the old image; the stages wait
        // for it only where they first read pixels (gl_image_ready), so sampling and the patch lists run under it.
        GL_CHECK(set_image_geometry(ctx, width, height, channels));
        GL_CUDA_CHECK(cudaEventRecord(ctx->ev_prev, ctx->stream));
        GL_CUDA_CHECK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_prev, 0));
        GL_CUDA_CHECK(cudaEventRecord(ctx->ev_begin[GL_T_H2D], ctx->copy_stream));
        GL_CUDA_CHECK(cudaMemcpyAsync(ctx->img->ptr, pixels, (size_t)(ctx->n * channels), cudaMemcpyHostToDevice, ctx->copy_stream));
        GL_CUDA_CHECK(cudaEventRecord(ctx->ev_end[GL_T_H2D], ctx->copy_stream));
        ctx->ev_valid[GL_T_H2D] = true;
        GL_CUDA_CHECK(cudaEventRecord(ctx->ev_h2d, ctx->copy_stream));
        ctx->h2d_pending = true;
    } else {
        // a rank of a multi-GPU run needs its own band of rows and the sampled pixels, nothing else: upload the band now, the
        // p sample values right after the sampling stage (gl_run_resident); the rest of ctx->img is not meaningful
        GL_CHECK(set_image_geometry(ctx, width, height, channels));
        GL_CUDA_CHECK(cudaEventRecord(ctx->ev_prev, ctx->stream));
        GL_CUDA_CHECK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_prev, 0));
        GL_CUDA_CHECK(cudaEventRecord(ctx->ev_begin[GL_T_H2D], ctx->copy_stream));
        GL_CUDA_CHECK(cudaMemcpyAsync((uint8_t*)ctx->img->ptr + (size_t)ctx->q0 * channels, pixels + (size_t)ctx->q0 * channels,
                                      (size_t)(ctx->q1 - ctx->q0) * channels, cudaMemcpyHostToDevice, ctx->copy_stream));
        GL_CUDA_CHECK(cudaEventRecord(ctx->ev_end[GL_T_H2D], ctx->copy_stream));
        ctx->ev_valid[GL_T_H2D] = true;
        GL_CUDA_CHECK(cudaEventRecord(ctx->ev_h2d, ctx->copy_stream));
        ctx->h2d_pending = true;
        ctx->host_pixels = pixels;
    }
    ctx->total_started = true;
    const int rc = gl_run_resident(ctx, prm, z_f32, z_u8, p_out, m_out, eigvals_out, eigvals_cap);
    gl_image_ready(ctx);     // (a run that failed before its first pixel read: the upload still orders before later work)
    ctx->host_pixels = nullptr;
    return rc;
}

}  // extern "C"
