// a-1: pixel sampling on the device, bit-exact with the reference.
//   uniform grid  <- UniformSampling, hpc/sampling.c:6-23 (= python/sampling/spatially_uniform.py:9-24)
//   random        <- random_sample, python/sampling/random.py:8-16 after np.random.seed(seed):
//                    MT19937 words, masked rejection to [0,n), first p distinct values, sorted.
// Also the deterministic synthetic-image generator used by the benchmarks (integer-only, equal to
// oracle_np.synthetic_image bit for bit).
#include <cmath>

#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// uniform grid
// ---------------------------------------------------------------------------------------------
__global__ void k_uniform_grid(uint32_t* __restrict__ out, unsigned count, unsigned p_pad, unsigned cols, unsigned xy0,
                               unsigned dist, unsigned width)
{
    unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p_pad) return;
    if (c < count) {
        unsigned i = xy0 + (c / cols) * dist, j = xy0 + (c % cols) * dist;
        out[c] = width * i + j;
    } else {
        out[c] = 0xffffffffu;
    }
}

static int set_sample_buffer(gl_ctx* ctx, unsigned count)
{
    const int p_pad = (int)round_up(count, 64);
    if (ctx->samples) gl_buf_release(ctx->samples);
    ctx->samples = nullptr;
    GL_CHECK(gl_alloc(ctx, sizeof(uint32_t) * p_pad, &ctx->samples));
    ctx->p = count;
    ctx->p_pad = p_pad;
    ctx->h_samples_valid = false;
    ctx->sample_epoch++;
    return GL_OK;
}

int gl_impl_sampling_uniform(gl_ctx* ctx, unsigned requested, unsigned* actual)
{
    const int width = ctx->width, height = ctx->height;
    GL_REQUIRE((int64_t)requested <= ctx->n, "sampling: requested %u > pixels", requested);
    // same arithmetic as sampling.c:8-11 (integer division before the double sqrt)
    const unsigned dist = (unsigned)std::sqrt((double)((width * height) / (int)requested));
    GL_REQUIRE(dist >= 1, "sampling: sample distance 0");
    const unsigned xy0 = dist / 2;
    const unsigned rows = (unsigned)std::ceil((height - 1 - (int)xy0) / (double)dist);
    const unsigned cols = (unsigned)std::ceil((width - 1 - (int)xy0) / (double)dist);
    const unsigned count = rows * cols;
    GL_REQUIRE(count >= 2, "sampling: grid has %u samples (need >= 2)", count);
    GL_CHECK(set_sample_buffer(ctx, count));
    k_uniform_grid<<<(unsigned)ceil_div(ctx->p_pad, 256), 256, 0, ctx->stream>>>((uint32_t*)ctx->samples->ptr, count,
                                                                                 (unsigned)ctx->p_pad, cols, xy0, dist,
                                                                                 (unsigned)width);
    GL_LAUNCH_CHECK(ctx);
    // host mirror (same formula; needed by the K_B tile table, affinity.cu)
    ctx->h_samples.resize(count);
    for (unsigned c = 0; c < count; ++c) ctx->h_samples[c] = (unsigned)width * (xy0 + (c / cols) * dist) + xy0 + (c % cols) * dist;
    ctx->h_samples_valid = true;
    if (actual) *actual = count;
    return GL_OK;
}

// ---------------------------------------------------------------------------------------------
// random: one CTA.  MT19937 twist in its three dependency-free phases, masked rejection with an
// order-preserving compaction, then "first p distinct in stream order" by a bitonic sort of
// (value, position) keys, and a final ascending compaction.
// ---------------------------------------------------------------------------------------------
#define RS_THREADS 1024
#define RS_MAXCAND 8192

__device__ __forceinline__ uint32_t mt_mix(uint32_t cur, uint32_t nxt, uint32_t far)
{
    uint32_t y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
    return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

// exclusive block scan of one int per thread; returns the exclusive prefix, *total = block sum
__device__ int block_scan_excl(int v, int* total, int* warp_sums /* [33] smem */)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
        int s = warp_sums[lane];  // RS_THREADS/32 == 32 warps
        int si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += t;
        }
        warp_sums[lane] = si - s;
        if (lane == 31) warp_sums[32] = si;
    }
    __syncthreads();
    const int base = warp_sums[w];
    *total = warp_sums[32];
    __syncthreads();
    return base + incl - v;
}

__global__ void __launch_bounds__(RS_THREADS, 1)
k_random_sampling(uint32_t seed, uint32_t n, uint32_t p, uint32_t p_pad, uint32_t* __restrict__ out, int* __restrict__ status)
{
    extern __shared__ unsigned char rs_smem[];
    unsigned long long* keys = (unsigned long long*)rs_smem;             // [RS_MAXCAND]
    uint32_t* cand = (uint32_t*)(keys + RS_MAXCAND);                      // [RS_MAXCAND] accepted draws in stream order
    unsigned short* first_at = (unsigned short*)(cand + RS_MAXCAND);      // [RS_MAXCAND] 1 if first occurrence (by position)
    unsigned short* rank_at = first_at + RS_MAXCAND;                      // [RS_MAXCAND] inclusive count of firsts up to pos
    __shared__ uint32_t mt[624];
    __shared__ int warp_sums[33];
    __shared__ int s_accepted, s_target, s_done;

    const int tid = threadIdx.x;
    const uint32_t rng = n - 1;
    uint32_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;

    if (tid == 0) {
        uint32_t x = seed;
        mt[0] = x;
        for (int i = 1; i < 624; ++i) {
            x = 1812433253u * (x ^ (x >> 30)) + (uint32_t)i;
            mt[i] = x;
        }
        s_accepted = 0;
        s_target = (int)p;
        s_done = 0;
    }
    __syncthreads();

    for (int round = 0; round < 64; ++round) {
        // ---- draw until s_target accepted candidates are held ----
        while (s_accepted < s_target) {
            // twist, phase A [0,227), B [227,454), C [454,623), then word 623
            uint32_t nv = 0;
            if (tid < 227) nv = mt_mix(mt[tid], mt[tid + 1], mt[tid + 397]);
            __syncthreads();
            if (tid < 227) mt[tid] = nv;
            __syncthreads();
            if (tid >= 227 && tid < 454) nv = mt_mix(mt[tid], mt[tid + 1], mt[tid - 227]);
            __syncthreads();
            if (tid >= 227 && tid < 454) mt[tid] = nv;
            __syncthreads();
            if (tid >= 454 && tid < 623) nv = mt_mix(mt[tid], mt[tid + 1], mt[tid - 227]);
            __syncthreads();
            if (tid >= 454 && tid < 623) mt[tid] = nv;
            __syncthreads();
            if (tid == 623) mt[623] = mt_mix(mt[623], mt[0], mt[396]);
            __syncthreads();
            // temper + masked rejection (numpy legacy randint: word & mask, reject > n-1)
            uint32_t v = 0;
            int ok = 0;
            if (tid < 624) {
                uint32_t y = mt[tid];
                y ^= y >> 11;
                y ^= (y << 7) & 0x9d2c5680u;
                y ^= (y << 15) & 0xefc60000u;
                y ^= y >> 18;
                v = y & mask;
                ok = v <= rng;
            }
            int total;
            const int base = s_accepted;
            const int off = block_scan_excl(ok, &total, warp_sums);
            if (base + total > RS_MAXCAND) {
                if (tid == 0) *status = 1;  // more candidates than this kernel holds
                return;
            }
            if (ok) cand[base + off] = v;
            __syncthreads();
            if (tid == 0) s_accepted = base + total;
            __syncthreads();
        }
        // numpy looks at exactly s_target draws in this round (python/sampling/random.py:9-12: p values, then p - len(unique) more, ...);
        // what the twist produced beyond them waits for the next round -- and the sort is over 1 024 keys instead of 2 048 at p = 1000
        const int A = min(s_accepted, s_target);
        int N = 2;
        while (N < A) N <<= 1;
        // ---- sort (value, position) ----
        for (int i = tid; i < N; i += RS_THREADS)
            keys[i] = i < A ? (((unsigned long long)cand[i] << 32) | (unsigned)i) : ~0ull;
        __syncthreads();
        for (int k = 2; k <= N; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < N; i += RS_THREADS) {
                    int ixj = i ^ j;
                    if (ixj > i) {
                        unsigned long long a = keys[i], b = keys[ixj];
                        bool up = (i & k) == 0;
                        if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
                    }
                }
                __syncthreads();
            }
        // ---- first occurrences, by position ----
        for (int i = tid; i < A; i += RS_THREADS) {
            unsigned long long kk = keys[i];
            bool first = (i == 0) || ((uint32_t)(keys[i - 1] >> 32) != (uint32_t)(kk >> 32));
            first_at[(uint32_t)kk] = first ? 1 : 0;
        }
        __syncthreads();
        // inclusive prefix over positions: each thread owns a contiguous chunk
        const int chunk = (A + RS_THREADS - 1) / RS_THREADS;
        const int lo = tid * chunk, hi = min(lo + chunk, A);
        int local = 0;
        for (int i = lo; i < hi; ++i) local += first_at[i];
        int distinct;
        int run = block_scan_excl(local, &distinct, warp_sums);
        for (int i = lo; i < hi; ++i) {
            run += first_at[i];
            rank_at[i] = (unsigned short)run;
        }
        __syncthreads();
        if (distinct >= (int)p) {
            // ---- ascending compaction of the kept keys ----
            const int chunk2 = (A + RS_THREADS - 1) / RS_THREADS;
            const int lo2 = tid * chunk2, hi2 = min(lo2 + chunk2, A);
            int cnt = 0;
            for (int i = lo2; i < hi2; ++i) {
                uint32_t pos = (uint32_t)keys[i];
                cnt += (first_at[pos] && rank_at[pos] <= p) ? 1 : 0;
            }
            int tot;
            int o = block_scan_excl(cnt, &tot, warp_sums);
            for (int i = lo2; i < hi2; ++i) {
                uint32_t pos = (uint32_t)keys[i];
                if (first_at[pos] && rank_at[pos] <= p) out[o++] = (uint32_t)(keys[i] >> 32);
            }
            for (uint32_t i = p + tid; i < p_pad; i += RS_THREADS) out[i] = 0xffffffffu;
            if (tid == 0) { *status = (tot == (int)p) ? 0 : 2; s_done = 1; }
            return;
        }
        if (tid == 0) s_target = A + ((int)p - distinct);  // top up, python/sampling/random.py:11-12
        __syncthreads();
    }
    if (tid == 0) *status = 3;
}

int gl_impl_sampling_random(gl_ctx* ctx, unsigned requested, uint32_t seed, unsigned* actual)
{
    GL_REQUIRE(requested >= 2, "random sampling: need >= 2 samples");
    if (requested > RS_MAXCAND - 1024) {
        gl_set_error("random sampling on device holds at most %d samples (asked %u)", RS_MAXCAND - 1024, requested);
        return GL_ERR_UNSUPPORTED;
    }
    GL_CHECK(set_sample_buffer(ctx, requested));
    const size_t smem = RS_MAXCAND * (8 + 4 + 2 + 2);
    GL_CUDA_CHECK(cudaFuncSetAttribute(k_random_sampling, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (ctx->async_mode) {
        // inside gl_run_resident: the kernel's status word goes to the deferred status block, nobody waits for it here; the host
        // mirror of the indices is fetched only if somebody asks (gl_host_samples)
        int* st = (int*)ctx->dstat->ptr + GL_DS_SAMPLING;
        GL_CUDA_CHECK(cudaMemsetAsync(st, 0xff, sizeof(int), ctx->stream));
        k_random_sampling<<<1, RS_THREADS, smem, ctx->stream>>>(seed, (uint32_t)ctx->n, requested, (uint32_t)ctx->p_pad,
                                                               (uint32_t*)ctx->samples->ptr, st);
        GL_LAUNCH_CHECK(ctx);
        if (actual) *actual = requested;
        return GL_OK;
    }
    gl_buf* st = nullptr;
    GL_CHECK(gl_alloc(ctx, sizeof(int), &st));
    GL_CUDA_CHECK(cudaMemsetAsync(st->ptr, 0xff, sizeof(int), ctx->stream));
    k_random_sampling<<<1, RS_THREADS, smem, ctx->stream>>>(seed, (uint32_t)ctx->n, requested, (uint32_t)ctx->p_pad,
                                                           (uint32_t*)ctx->samples->ptr, (int*)st->ptr);
    GL_LAUNCH_CHECK(ctx);
    // the status word and the indices come back in the same synchronisation (host mirror for the K_B tile table)
    GL_CHECK(gl_ensure_pinned(ctx, 64 + sizeof(uint32_t) * requested));
    GL_CUDA_CHECK(cudaMemcpyAsync(ctx->pinned, st->ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    GL_CUDA_CHECK(cudaMemcpyAsync((char*)ctx->pinned + 64, ctx->samples->ptr, sizeof(uint32_t) * requested, cudaMemcpyDeviceToHost,
                                  ctx->stream));
    GL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    gl_buf_release(st);
    const int status = *(int*)ctx->pinned;
    ctx->h_samples_valid = false;
    if (status == 0) {
        const uint32_t* hs = (const uint32_t*)((char*)ctx->pinned + 64);
        ctx->h_samples.assign(hs, hs + requested);
        ctx->h_samples_valid = true;
    }
    if (status != 0) {
        gl_set_error("random sampling kernel failed (status %d)", status);
        return status == 1 ? GL_ERR_UNSUPPORTED : GL_ERR_CUDA;
    }
    if (actual) *actual = requested;
    return GL_OK;
}

// ---------------------------------------------------------------------------------------------
// synthetic image
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

__global__ void k_synthetic(uint8_t* __restrict__ img, int width, int height, int channels, uint32_t salt)
{
    const int per[3][2] = {{97, 131}, {113, 89}, {71, 149}};
    const int64_t total = (int64_t)width * height * channels;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int ch = (int)(idx % channels);
        const int64_t q = idx / channels;
        const int r = (int)(q / width), c = (int)(q % width);
        const int pr = per[ch][0], pc = per[ch][1];
        const int tr = 64 - abs(((r % pr) * 256) / pr - 128);
        const int tc = 64 - abs(((c % pc) * 256) / pc - 128);
        const int smooth = (80 * tr * tc + 4096 * 80) / 4096 - 80;
        const int edges = 24 * (((r >> 6) + (c >> 6)) & 1);
        const int noise = (int)(hash32((uint32_t)idx + salt) % 33u) - 16;
        const int v = 128 + smooth + edges + noise;
        img[idx] = (uint8_t)min(max(v, 0), 255);
    }
}

int gl_impl_synthetic(gl_ctx* ctx, uint32_t seed)
{
    k_synthetic<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((uint8_t*)ctx->img->ptr, ctx->width, ctx->height, ctx->channels,
                                                            seed * 0x9e3779b1u);
    GL_LAUNCH_CHECK(ctx);
    return GL_OK;
}
