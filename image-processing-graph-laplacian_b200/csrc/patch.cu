// The PATCH layout of K_B and the kernels built on it: the default path of the spatially decaying affinities (bilateral, spatial).
//
// A sample further than r_c = h_loc sqrt(25 ln 2) from a pixel has exp(-d^2/h_loc^2) < 2^-25, which the fp16 K_B flushes to zero
// (SURVEY H1).  At 4K / p = 1000 / h_loc = 40 a pixel has ~12 samples within reach out of 1000, so K_B is 99 % structural zeros.
// The blocked layout (affinity.cu) covers them with contiguous runs of a sorted sample order and still stores ~106 slots per
// pixel; here every PATCH of 64 x 16 pixels gets its own GATHERED list of the samples within r_c of the patch rectangle
// (ascending sample index, padded to blocks of 32 slots: ~14 samples -> one block), and
//
//   K_B   is stored as A tiles [128 pixels = 2 patch rows][32 slots] fp16 (8 KB, the K-major A operand of one MMA tile),
//         tile (first_block(patch) * 8 + mt * nb + b) for M tile mt and slot block b of the patch, loaded by one SWIZZLE_64B tensor-map
//         copy (a pre-swizzled tile fetched by a plain 8 KB bulk copy was measured 10 % slower, profiles/r02_patch_timeline.md);
//   W     = -alpha U Lambda^-1 is kept SAMPLE-major [p_pad][m_pad] fp16, so that the W rows of a patch's samples are gathered
//         straight into the MN-major B operand [32 slots][256 columns] in shared memory, once per (patch, N tile), and stay
//         there for the patch's eight M tiles;
//   Phi   tile = A . B on tcgen05 (128 x 256 x 16, fp32 accumulators in tensor memory, 1 or 2 K steps), consumed by the fused
//         filter epilogue (row dots with the weights g f(lambda) o c) and never written to HBM.
//
// Replaces, for these affinities: ComputeAffinityMatrices' K_B part (hpc/affinity.c:196-262), Nystroem (hpc/nystroem.c:5-69),
// Permutation (hpc/utils.c:134-173) and ComputeResultFromLaplacian's two products (hpc/display.c:64-73).  Whoever needs the
// matrices themselves (Phi download, orthonormalisation, K_B without cutoff) goes through the blocked path (affinity.cu,
// nystroem_gemm.cu), which stays the checker of this one.
//
// The lists are built on the device from the device-resident sample indices (no host planning).
#include <cmath>

#include "tc_common.cuh"

namespace pt {

constexpr int G = 8;              // M tiles per patch
constexpr int PW = 64, PR = 16;   // patch: 64 columns x 16 rows; M tile = 2 rows of it
constexpr int SLOTS = 32;         // sample slots per block
constexpr int A_TILE_BYTES = 128 * SLOTS * 2;     // 8 KB
constexpr int B_BLOCK_BYTES = SLOTS * 256 * 2;    // 16 KB: 4 chunks of [32 K rows][64 columns = 128 B]
constexpr int B_CHUNK_BYTES = SLOTS * 128;        // 4 KB
constexpr int SA = 6;             // A ring stages
constexpr int SBT = 8;            // B units: [32 slots][256 columns of one N tile], 16 KB each
constexpr int EPI_WARPS = 8;       // epilogue warps per group: 4 TMEM lane quarters x 2 column shares
constexpr int THREADS = 128 + 32 * EPI_WARPS;

struct Geom {
    int width, row0, band_rows;   // image width, first image row of the band, rows in the band
    int pcols, prows, npatch;
    float rc2;                    // squared reach (pixels^2): samples further than this from the patch rectangle are dropped
};

// squared distance from sample (sr, sc) to the patch rectangle (image coordinates)
__device__ __forceinline__ float patch_dist2(const Geom& g, int patch, int sr, int sc)
{
    const int py = patch / g.pcols, px = patch - py * g.pcols;
    const int r_lo = g.row0 + py * PR, r_hi = min(g.row0 + g.band_rows, r_lo + PR) - 1;
    const int c_lo = px * PW, c_hi = min(g.width, c_lo + PW) - 1;
    const int dy = max(max(r_lo - sr, sr - r_hi), 0), dx = max(max(c_lo - sc, sc - c_hi), 0);
    return (float)dy * (float)dy + (float)dx * (float)dx;
}

// The samples are in ascending raster order, i.e. sorted by image row: the ones that can reach a patch (rows within ceil(r_c) of its
// rows) are one contiguous run [lo, hi) of the list, found by two binary searches -- at 4K that is 15 % of the samples.
__device__ __forceinline__ int sample_lower_bound(const uint32_t* __restrict__ samples, int p, long long q)
{
    int lo = 0, hi = p;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((long long)samples[mid] < q) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ void sample_window(const Geom& g, int patch, const uint32_t* __restrict__ samples, int p, int& lo, int& hi)
{
    const int py = patch / g.pcols;
    const int r_lo = g.row0 + py * PR, r_hi = min(g.row0 + g.band_rows, r_lo + PR) - 1;
    const int R = (int)ceilf(sqrtf(g.rc2)) + 1;
    lo = sample_lower_bound(samples, p, (long long)max(0, r_lo - R) * g.width);
    hi = sample_lower_bound(samples, p, (long long)(r_hi + R + 1) * g.width);
}

// ---------------------------------------------------------------------------------------------
// lists: count -> scan -> fill (one warp per patch; ascending sample index, deterministic)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_patch_count(Geom g, const uint32_t* __restrict__ samples, int p, int4* __restrict__ pinfo)
{
    const int patch = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (patch >= g.npatch) return;
    int cnt = 0, lo, hi;
    sample_window(g, patch, samples, p, lo, hi);
    for (int i0 = lo; i0 < hi; i0 += 32) {
        const int i = i0 + lane;
        bool in = false;
        if (i < hi) {
            const uint32_t q = samples[i];
            in = patch_dist2(g, patch, (int)(q / (uint32_t)g.width), (int)(q % (uint32_t)g.width)) <= g.rc2;
        }
        cnt += __popc(__ballot_sync(0xffffffffu, in));
    }
    if (lane == 0) pinfo[patch] = make_int4(0, max(1, (cnt + SLOTS - 1) / SLOTS), cnt, 0);
}

// exclusive scan of the block counts (one CTA); total[0] = blocks in all, total[1] = 16-slot K steps the extrapolation will issue
// (per M tile of a patch: 2 per full block, 1 for a last block with at most 16 samples)
__global__ void __launch_bounds__(1024) k_patch_scan(Geom g, int npatch, int4* __restrict__ pinfo, int* __restrict__ total, long long cap_blocks,
                                                     int* __restrict__ dstat)
{
    // Warp w owns the contiguous run of `per` x 32 patches [w per 32, (w + 1) per 32); in every iteration its lanes read 32 consecutive
    // records (512 contiguous bytes: with one thread per run of `per` records every load touched its own sector, and the single SM
    // that runs this kernel spent 20 us on them).  Pass 1: block counts per warp; scan of the 32 warp totals; pass 2: the records again,
    // warp-scanned, offsets written.
    __shared__ int wsum[32], wbase[33];
    __shared__ unsigned long long wks[32];
    __shared__ int wmax[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = (npatch + 1023) / 1024;
    const int a = warp * per * 32, b = min(npatch, a + per * 32);
    int s = 0, nbmax = 0;
    unsigned long long ks = 0;
    for (int i = a + lane; i < b; i += 32) {
        const int4 pi = pinfo[i];
        s += pi.y;
        nbmax = max(nbmax, pi.y);
        const int mtc = min(G, (g.band_rows - (i / g.pcols) * PR + 1) >> 1);
        const int last = pi.z - SLOTS * (pi.y - 1);
        ks += (unsigned long long)mtc * (unsigned long long)(2 * (pi.y - 1) + (last > 16 ? 2 : 1));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        ks += __shfl_xor_sync(0xffffffffu, ks, o);
        nbmax = max(nbmax, __shfl_xor_sync(0xffffffffu, nbmax, o));
    }
    if (lane == 0) { wsum[warp] = s; wks[warp] = ks; wmax[warp] = nbmax; }
    __syncthreads();
    if (warp == 0) {
        const int v = wsum[lane];
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        wbase[lane] = incl - v;
        if (lane == 31) wbase[32] = incl;
    }
    __syncthreads();
    int run = wbase[warp];
    for (int i0 = a; i0 < b; i0 += 32) {
        const int i = i0 + lane;
        const int nb = i < b ? pinfo[i].y : 0;
        int incl = nb;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (i < b) pinfo[i].x = run + incl - nb;
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (threadIdx.x == 0) {
        unsigned long long ksum = 0ull;
        int max_nb = 0;
        for (int w = 0; w < 32; ++w) { ksum += wks[w]; max_nb = max(max_nb, wmax[w]); }
        const int blocks = wbase[32];
        total[0] = blocks;
        *(unsigned long long*)(total + 2) = ksum;
        // storage set aside without asking (cap_blocks > 0): too small -> every kernel of the patch path returns at once and the
        // host, which reads the status block with the results, runs the step again
        dstat[GL_DS_PT_BLOCKS] = blocks;
        dstat[GL_DS_PT_MAXNB] = max_nb;
        *(unsigned long long*)(dstat + GL_DS_PT_KSTEPS) = ksum;
        dstat[GL_DS_PT_OVERFLOW] = (cap_blocks > 0 && (long long)blocks > cap_blocks) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256) k_patch_fill(Geom g, const uint32_t* __restrict__ samples, int p, const int4* __restrict__ pinfo,
                                                    uint32_t* __restrict__ slots, const int* __restrict__ dstat)
{
    if (dstat[GL_DS_PT_OVERFLOW]) return;
    const int patch = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (patch >= g.npatch) return;
    const int4 pi = pinfo[patch];
    uint32_t* out = slots + (size_t)pi.x * SLOTS;
    int pos = 0, lo, hi;
    sample_window(g, patch, samples, p, lo, hi);
    for (int i0 = lo; i0 < hi; i0 += 32) {
        const int i = i0 + lane;
        bool in = false;
        if (i < hi) {
            const uint32_t q = samples[i];
            in = patch_dist2(g, patch, (int)(q / (uint32_t)g.width), (int)(q % (uint32_t)g.width)) <= g.rc2;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        if (in) out[pos + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)i;
        pos += __popc(bal);
    }
    for (int j = pos + lane; j < pi.y * SLOTS; j += 32) out[j] = 0xffffffffu;
}

// sample features in the caller's order, SoA with stride p_pad: [0] row, [1] col, [2..2+C) values
__global__ void k_patch_sample_features(const uint8_t* __restrict__ img, const uint32_t* __restrict__ samples, int p, int p_pad, int width,
                                        int channels, float* __restrict__ sf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p_pad) return;
    if (i < p) {
        const uint32_t q = samples[i];
        sf[i] = (float)(q / (uint32_t)width);
        sf[p_pad + i] = (float)(q % (uint32_t)width);
        for (int ch = 0; ch < channels; ++ch) sf[(2 + ch) * p_pad + i] = (float)img[(size_t)q * channels + ch];
    } else {
        for (int k = 0; k < 2 + channels; ++k) sf[k * p_pad + i] = 1e18f;
    }
}

// ---------------------------------------------------------------------------------------------
// affinity: K_B tiles + the row sums D and image-weighted sums T (from the fp32 values, before the fp16 rounding)
// ---------------------------------------------------------------------------------------------
// Slot group tx (eight consecutive slots of the block) is uniform per WARP, so that a group without a sample costs nothing: at
// config 4 a patch has 13.8 samples in reach on average, and with the slot groups spread over the lanes (round 2's first cut) half
// of every warp's pairs were empty slots.  A block with at most 16 samples ("narrow": its slots 16..31 are never multiplied, the
// extrapolation takes one K step) is split as warp -> (group w & 1, column half (w >> 1) & 1, rows 8 (w >> 2) .. + 8); a fuller one
// as warp -> (group w & 3, column half w >> 2, all 16 rows).  A lane is one pixel column: the column part of the exponent is the
// same for all of the thread's pixels and leaves the pair loop.  A pixel's 8 values leave as one 16-byte store into the A-tile
// position the MMA will read (the four groups of a pixel row fill its 64 bytes between them; L2 merges the sectors).

// Warp-wide sums of 8 values per lane with 9 shuffles (the halves a lane does not keep travel): on return lane L holds the sum over
// the warp of v[(L >> 2) & 7].  Fixed tree: deterministic.
__device__ __forceinline__ float warp_sum8(const float (&v)[8], int lane)
{
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
    float a[4], b[2];
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] = (h16 ? v[4 + j] : v[j]) + __shfl_xor_sync(0xffffffffu, h16 ? v[j] : v[4 + j], 16);
#pragma unroll
    for (int j = 0; j < 2; ++j) b[j] = (h8 ? a[2 + j] : a[j]) + __shfl_xor_sync(0xffffffffu, h8 ? a[j] : a[2 + j], 8);
    float c = (h4 ? b[1] : b[0]) + __shfl_xor_sync(0xffffffffu, h4 ? b[0] : b[1], 4);
    c += __shfl_xor_sync(0xffffffffu, c, 2);
    c += __shfl_xor_sync(0xffffffffu, c, 1);
    return c;
}

template <int KIND, int C>
__global__ void __launch_bounds__(256, C == 1 ? 3 : 2)
k_patch_affinity(Geom g, const uint8_t* __restrict__ img, const float* __restrict__ sf, int p_pad, float a2, float b2,
                 const int4* __restrict__ pinfo, const uint32_t* __restrict__ slots, __half* __restrict__ KB,
                 float* __restrict__ partial /* [gridDim.x][1 + C][p_pad] */, const int* __restrict__ dstat)
{
    if (dstat[GL_DS_PT_OVERFLOW]) return;
    extern __shared__ float pa_smem[];
    constexpr int NS = 1 + C;
    float* cta_sum = pa_smem;                  // [NS][p_pad]
    float* ws = cta_sum + NS * p_pad;          // 2 x [NS][8 warps][8 slots of the warp's group] (room for [NS][8][SLOTS])
    float* px = ws + NS * 8 * SLOTS;           // 2 x [C][PW * PR] pixel values
    __shared__ uint32_t sid[3][SLOTS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < NS * p_pad; i += 256) cta_sum[i] = 0.f;

    // The global loads a patch starts with (its list header, its first block of sample indices, its pixels) are issued a patch ahead
    // and land in registers under the previous patch's arithmetic (ncu: a quarter of this kernel's stall samples sat on the pixel
    // load, another fifth on its three barriers per block).  Shared buffers rotate (sid x 3, ws x 2, px x 2), which leaves ONE barrier
    // per block: what a block's arithmetic reads was written before the previous barrier, what its closing sums read is not
    // rewritten before the next one.
    const int stride = gridDim.x;
    int patch = blockIdx.x;
    if (patch >= g.npatch) return;
    // (the bytes stay untouched in their registers until store_pixels: anything computed on them here would wait for the loads)
    auto load_pixels = [&](int pt_, uint8_t (&r)[C][4]) {
        const int py = pt_ / g.pcols, pxi = pt_ - py * g.pcols;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = tid + 256 * j;
            const int rr = py * PR + (i >> 6), c = pxi * PW + (i & 63);
            const bool in = rr < g.band_rows && c < g.width;
            const size_t q = (size_t)(g.row0 + (in ? rr : 0)) * g.width + (in ? c : 0);
#pragma unroll
            for (int ch = 0; ch < C; ++ch) r[ch][j] = img[q * C + ch];
        }
    };
    auto store_pixels = [&](float* dst, const uint8_t (&r)[C][4]) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int ch = 0; ch < C; ++ch) dst[ch * PW * PR + tid + 256 * j] = (float)r[ch][j];
    };
    // list headers and sample indices travel global -> shared without a register (cp.async): header of local patch k in spi[k % 3],
    // fetched two patches ahead; a block's indices fetched a block ahead; both waited for right before the block's barrier
    __shared__ int4 spi[3];
    auto async_copy = [](void* dst, const void* src, int bytes) {
        if (bytes == 16) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(dst)), "l"(src) : "memory");
        else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(tc::smem_u32(dst)), "l"(src) : "memory");
    };
    {
        uint8_t r0[C][4];
        load_pixels(patch, r0);
        const int4 pi0 = pinfo[patch];
        if (tid == 0) {
            spi[0] = pi0;
            spi[1] = patch + stride < g.npatch ? pinfo[patch + stride] : make_int4(0, 1, 0, 0);
        }
        if (tid < SLOTS) sid[0][tid] = slots[(size_t)pi0.x * SLOTS + tid];
        store_pixels(px, r0);
    }
    __syncthreads();
    int cur_s = 0, cur_w = 0, cur_p = 0, kp = 0;   // kp = local patch counter mod 3

    for (; patch < g.npatch; patch += stride) {
        const int py = patch / g.pcols, pxi = patch - py * g.pcols;
        const int r_base = py * PR, c_base = pxi * PW;     // band-local row, column
        const int nxt = patch + stride;
        const int kp1 = kp == 2 ? 0 : kp + 1, kp2 = kp1 == 2 ? 0 : kp1 + 1;
        const int4 pi = spi[kp];
        const int pin_x = spi[kp1].x;      // (meaningful only while nxt < npatch)
        if (tid == 0 && nxt + stride < g.npatch) async_copy(&spi[kp2], pinfo + (nxt + stride), 16);
        uint8_t pxr[C][4];
        if (nxt < g.npatch) load_pixels(nxt, pxr);
        __half* kb_patch = KB + ((size_t)pi.x * G) * 128 * SLOTS;
        const float* pxc = px + cur_p * (C * PW * PR);
        for (int b = 0; b < pi.y; ++b) {
            const uint32_t* sidc = sid[cur_s];
            float* wsc = ws + cur_w * (NS * 8 * 8);
            const bool last_b = b + 1 == pi.y;
            const int nxt_s = cur_s == 2 ? 0 : cur_s + 1;
            // what the next block (of this patch, or the first one of the next patch) reads
            if (tid < SLOTS && (!last_b || nxt < g.npatch))
                async_copy(&sid[nxt_s][tid], slots + ((size_t)(last_b ? pin_x : pi.x + b + 1) * SLOTS + tid), 4);
            const int cnt = min(SLOTS, pi.z - SLOTS * b);    // samples of this block: slots 0 .. cnt - 1 (k_patch_fill fills from 0 upwards)
            const bool narrow = cnt <= 16;
            const int tx = narrow ? (warp & 1) : (warp & 3);
            const int ty = (narrow ? ((warp >> 1) & 1) : (warp >> 2)) * 32 + lane;
            const int i0 = narrow ? (warp >> 2) * 8 : 0, i1 = narrow ? i0 + 8 : PR;
            const float pc = (float)(c_base + ty);
            const bool col_ok = c_base + ty < g.width;
            if (tx * 8 >= cnt) {
                // a group without a sample inside a K step that is multiplied: zeros, no arithmetic
                if (KB) {
                    for (int i = i0; i < i1; ++i) {
                        const int mt = i >> 1;
                        if (r_base + 2 * mt >= g.band_rows) break;
                        *(uint4*)&kb_patch[((size_t)(mt * pi.y + b) * 128 + (size_t)((i & 1) * PW + ty)) * SLOTS + tx * 8] = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
            } else {
                float sr[8], sv[C][8], cterm[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t s = sidc[tx * 8 + k];
                    const bool empty = s == 0xffffffffu;   // an empty slot gets features at 1e18: its exponent is hugely negative, K = 0
                    const int si = empty ? 0 : (int)s;
                    sr[k] = empty ? 1e18f : sf[si];
                    const float sc = empty ? 1e18f : sf[p_pad + si];
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) sv[ch][k] = empty ? 1e18f : sf[(2 + ch) * p_pad + si];
                    cterm[k] = 0.f;
                    if (KIND != GL_PHOTOMETRIC) {
                        const float dc = pc - sc;
                        cterm[k] = dc * dc * a2;   // (an empty slot: -inf -> K = 0)
                    }
                }
                float acc[8], tacc[C][8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    acc[k] = 0.f;
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) tacc[ch][k] = 0.f;
                }
#pragma unroll 2
                for (int i = i0; i < i1; ++i) {
                    const int mt = i >> 1;
                    if (r_base + 2 * mt >= g.band_rows) break;      // M tiles wholly below the band are neither stored nor multiplied
                    // a pixel outside the image / band gets a row coordinate at 1e18: its spatial exponent is hugely negative and K = 0
                    // without a select per pair (both kinds served here have the spatial term; 1e36 * a2 stays finite)
                    const bool ok = col_ok && r_base + i < g.band_rows;
                    const float pr = ok ? (float)(g.row0 + r_base + i) : 1e18f;
                    float pv[C];
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) pv[ch] = pxc[ch * PW * PR + i * PW + ty];
                    float kv[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        float x = cterm[k];
                        if (KIND != GL_PHOTOMETRIC) {
                            const float dr = pr - sr[k];
                            x = fmaf(dr * dr, a2, x);
                        }
                        if (KIND != GL_SPATIAL) {
                            float d = pv[0] - sv[0][k];
                            float t = d * d;
#pragma unroll
                            for (int ch = 1; ch < C; ++ch) {
                                d = pv[ch] - sv[ch][k];
                                t = fmaf(d, d, t);
                            }
                            x = fmaf(t, b2, x);
                        }
                        kv[k] = fast_exp2(x);
                        acc[k] += kv[k];
#pragma unroll
                        for (int ch = 0; ch < C; ++ch) tacc[ch][k] = fmaf(kv[k], pv[ch], tacc[ch][k]);
                    }
                    __half2 h0 = __floats2half2_rn(kv[0], kv[1]), h1 = __floats2half2_rn(kv[2], kv[3]);
                    __half2 h2 = __floats2half2_rn(kv[4], kv[5]), h3 = __floats2half2_rn(kv[6], kv[7]);
                    uint4 pk;
                    pk.x = *(uint32_t*)&h0; pk.y = *(uint32_t*)&h1; pk.z = *(uint32_t*)&h2; pk.w = *(uint32_t*)&h3;
                    if (KB) *(uint4*)&kb_patch[((size_t)(mt * pi.y + b) * 128 + (size_t)((i & 1) * PW + ty)) * SLOTS + tx * 8] = pk;
                }
                // sums: over the warp's 32 columns (and its rows), then over the warps of the group through shared memory, in a fixed
                // order (deterministic)
                const float s0 = warp_sum8(acc, lane);
                if ((lane & 3) == 0) wsc[warp * 8 + ((lane >> 2) & 7)] = s0;
#pragma unroll
                for (int ch = 0; ch < C; ++ch) {
                    const float s1 = warp_sum8(tacc[ch], lane);
                    if ((lane & 3) == 0) wsc[((1 + ch) * 8 + warp) * 8 + ((lane >> 2) & 7)] = s1;
                }
            }
            if (last_b && nxt < g.npatch) store_pixels(px + (cur_p ^ 1) * (C * PW * PR), pxr);
            if (tid < SLOTS) asm volatile("cp.async.wait_all;" ::: "memory");
            __syncthreads();
            if (tid < NS * SLOTS) {
                const int which = tid / SLOTS, sl = tid % SLOTS;
                const uint32_t s = sidc[sl];
                if (s != 0xffffffffu) {
                    // (a non-empty slot's group is active in every warp that owns it: narrow -> warps grp, grp + 2, ..; else grp, grp + 4)
                    float sum = 0.f;
                    for (int wi = sl >> 3; wi < 8; wi += narrow ? 2 : 4) sum += wsc[(which * 8 + wi) * 8 + (sl & 7)];
                    cta_sum[which * p_pad + s] += sum;
                }
            }
            cur_s = nxt_s;
            cur_w ^= 1;
        }
        cur_p ^= 1;
        kp = kp1;
    }
    __syncthreads();
    for (int i = tid; i < NS * p_pad; i += 256) partial[(size_t)blockIdx.x * NS * p_pad + i] = cta_sum[i];
}

// DT[which][j] = sum over CTAs in fixed order (fp64); one warp per output
__global__ void k_patch_reduce(const float* __restrict__ partial, int nblocks, int p, int p_pad, int ns, double* __restrict__ DT)
{
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= ns * p_pad) return;
    const int which = i / p_pad, j = i - which * p_pad;
    if (j >= p) return;
    double acc = 0.0;
    for (int b = lane; b < nblocks; b += 32) acc += (double)partial[(size_t)b * ns * p_pad + i];
    acc = warp_sum(acc);
    if (lane == 0) DT[(size_t)which * p_pad + j] = acc;
}

// ---------------------------------------------------------------------------------------------
// W rows: W[s][j] = -alpha U[s][j] / mu_j * 2^e (fp16), sample-major [p_pad][m_pad]; 32 x 32 transpose tiles of the column-major U
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_w_rows(const float* __restrict__ U, int ld, int p, int m, int p_pad, int m_pad,
                                               const double* __restrict__ mu_inv, const double* __restrict__ neg_alpha,
                                               const float* __restrict__ scales, __half* __restrict__ W)
{
    __shared__ float t[32][33];
    const int s0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, tyy = threadIdx.x >> 5;
    for (int jj = tyy; jj < 32; jj += 8) {
        const int j = j0 + jj, s = s0 + tx;
        float v = 0.f;
        if (j < m && s < p) v = U[(size_t)j * ld + s] * ((float)(neg_alpha[0] * mu_inv[j]) * scales[0]);
        t[jj][tx] = v;
    }
    __syncthreads();
    for (int ss = tyy; ss < 32; ss += 8) {
        const int s = s0 + ss, j = j0 + tx;
        if (s < p_pad && j < m_pad) W[(size_t)s * m_pad + j] = __float2half_rn(t[tx][ss]);
    }
}

// ---------------------------------------------------------------------------------------------
// extrapolation + fused filter
// ---------------------------------------------------------------------------------------------
// contiguous global -> shared bulk copy (1-D TMA), completion on an mbarrier
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}

// Experiment switches of the kernel below are compiled in only with -DGLB200_PT_DEBUG (make PT_DEBUG=1): in the serial issuer loop
// even predicated-off instrumentation costs ~10 % of the kernel.
#ifdef GLB200_PT_DEBUG
#define PT_DBG(bit) (dbg & (bit))
#else
#define PT_DBG(bit) false
#endif

// One persistent CTA per SM works through whole patches (b, b + gridDim.x, ...).  For a patch: its W rows for ALL N tiles are
// gathered once (one 16 KB unit per N tile and slot block), each A tile is loaded once per M tile and multiplied against every N
// tile, and the epilogue carries a pixel's dot product across the N tiles in registers -- so the filtered pixel z = y + Phi_row . w
// is written by this kernel (no partial sums in HBM, no summation pass) and an A tile crosses L2 -> shared memory once instead of
// once per N tile.  Roles (12 warps):
//   warp 0      producer: one SWIZZLE_64B tensor-map copy (8 KB) per A tile into a ring of SA slots
//   warps 1, 3  MMA issuers (one thread each), issuer w owns accumulator w = every other (M tile, N tile).  These threads are the
//               serial part of the pipeline: every operation they execute costs 50-150 cycles (tools/probe_issue.cu), and with K
//               loops of one or two steps nothing hides that (profiles/r02_patch_timeline.md).  Two issuers only when the N tiles
//               come in pairs and every patch's operands stay resident; otherwise issuer 0 works alone.
//   warp 2      allocates the tensor memory, then gathers the W rows (cp.async, MN-major SWIZZLE_128B image) into the B units
//   warps 4-11  epilogue: warp w drains TMEM lanes 32 (w % 4) .. +31 (= 32 pixels of one patch row) and half of the accumulator's
//               columns -- all of them requested back to back, one wait, accumulator handed back, THEN the arithmetic; after the last N
//               tile the two column halves of a pixel meet in shared memory and the filtered pixel is stored.
template <int FC, int BN>
__global__ void __launch_bounds__(THREADS, 1)
k_patch_nystroem(const __grid_constant__ CUtensorMap map_a, Geom g, const int4* __restrict__ pinfo, const uint32_t* __restrict__ slots,
                 const __half* __restrict__ W, int m_pad, int n_tiles, const float* __restrict__ scales,
                 const float* __restrict__ fuse_w /* [m_pad][FC] */, const uint8_t* __restrict__ img /* band rows, [pixel][FC] */,
                 int clip_low, float* __restrict__ z /* [band pixels][FC] */, uint8_t* __restrict__ z8 /* or null */,
                 int* __restrict__ err, long long* __restrict__ prof /* debug timeline of CTA 0, or null */,
                 const int* __restrict__ dstat, int dbg /* experiments: bit 0 = no epilogue arithmetic, bit 1 = no stores */,
                 int skip_if_resident /* the dual-pipeline kernel was launched for the all-resident case: leave it to that one */)
{
    using namespace tc;
    if (dstat[GL_DS_PT_OVERFLOW]) return;
    if (skip_if_resident && dstat[GL_DS_PT_MAXNB] * 2 <= SA && dstat[GL_DS_PT_MAXNB] * n_tiles <= SBT) return;
    extern __shared__ uint8_t pn_smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)pn_smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem_a + SA * A_TILE_BYTES;
    uint64_t* bars = (uint64_t*)(smem_b + SBT * B_BLOCK_BYTES);
    const uint32_t bar_afull = smem_u32(bars), bar_aempty = smem_u32(bars + SA);
    const uint32_t bar_bfull = smem_u32(bars + 2 * SA), bar_bempty = smem_u32(bars + 2 * SA + SBT);
    const uint32_t bar_tfull = smem_u32(bars + 2 * SA + 2 * SBT), bar_tempty = smem_u32(bars + 2 * SA + 2 * SBT + 2);
    static_assert(2 * SA + 2 * SBT + 4 + 1 <= 64, "barrier block is 512 bytes");
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * SA + 2 * SBT + 4);
    float* xch = (float*)(bars + 64);            // [4 quarters][32 lanes][FC]: column share 1 hands its part of a pixel's dot to share 0
    float* w_s = xch + 4 * 32 * FC;              // [m_pad][FC] filter weights, times the GEMM's output scale

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int first_patch = blockIdx.x, patch_step = gridDim.x;
    const int max_nb = dstat[GL_DS_PT_MAXNB];
    // every patch keeps its operands resident (A tiles of an M tile in half the ring, W units of all N tiles in the B units)?
    const bool all_resident = max_nb * 2 <= SA && max_nb * n_tiles <= SBT;
    const bool two = all_resident && (n_tiles & 1) == 0;

    if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    if (warp == 1 && lane == 0) {
        // an A slot is released by TWO arrivals (its M tile is multiplied by both issuers; a lone issuer arrives twice)
        for (int s = 0; s < SA; ++s) { mbar_init(bar_afull + 8 * s, 1); mbar_init(bar_aempty + 8 * s, 2); }
        for (int s = 0; s < SBT; ++s) { mbar_init(bar_bfull + 8 * s, 1); mbar_init(bar_bempty + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {
        const float sc = scales[1];
        for (int i = threadIdx.x; i < m_pad * FC; i += THREADS) w_s[i] = __ldg(fuse_w + i) * sc;
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // per patch: nb slot blocks, mtc M tiles; resident = operands loaded once per M tile (A) / per patch (W units)
    if (warp == 0) {
        // ===== A producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int patch = first_patch; patch < g.npatch; patch += patch_step) {
                const int4 pi = pinfo[patch];
                const int py = patch / g.pcols;
                const int mtc = min(G, (g.band_rows - py * PR + 1) >> 1);
                const int tile0 = pi.x * G;
                const bool resident = pi.y * 2 <= SA && pi.y * n_tiles <= SBT;
                const int reps = resident ? 1 : n_tiles;          // not resident: the A tiles stream past once per N tile
                for (int mt = 0; mt < mtc; ++mt)
                    for (int rep = 0; rep < reps; ++rep)
                        for (int b = 0; b < pi.y; ++b) {
                            mbar_wait(bar_aempty + 8 * stage, phase ^ 1, err, 1);
                            mbar_expect_tx(bar_afull + 8 * stage, (uint32_t)A_TILE_BYTES);
                            tma_load_2d(smem_u32(smem_a + stage * A_TILE_BYTES), &map_a, bar_afull + 8 * stage, 0, (tile0 + mt * pi.y + b) * 128);
                            if (++stage == SA) { stage = 0; phase ^= 1; }
                        }
            }
        }
    } else if (warp == 1 || warp == 3) {
        // ===== MMA issuers =====
        if (lane == 0 && (two || warp == 1)) {
            const int me = warp == 1 ? 0 : 1;
#ifdef GLB200_PT_DEBUG
            long long pa[6] = {0, 0, 0, 0, 0, 0}, pt0 = 0, pt1;   // phase clocks of this issuer: tempty wait, operand waits, MMA issue, commits, tiles
#define PT_T(k) do { if (prof) { pt1 = clock64(); pa[k] += pt1 - pt0; pt0 = pt1; } } while (0)
            if (prof) pt0 = clock64();
#else
#define PT_T(k) do { } while (0)
#endif
            int stage = 0, bpos = 0, it = 0;
            uint32_t phase = 0, buses = 0;   // buses bit s: parity of the number of fills of B unit slot s consumed so far
            const uint32_t idesc = make_idesc(BLOCK_M, BN, 0) | (1u << 16);   // B is MN-major
            if (two) {
                // ---- the common case, written out lean: every patch resident, N tiles in pairs.  Issuer `me` owns accumulator `me`
                // and the N tiles of its parity; this thread is the serial resource of the SM (one scalar instruction every ~5 cycles),
                // so nothing is recomputed per tile that can be kept per M tile or per patch, and it never walks the other issuer's
                // tiles.  B unit u (in the order the gather warp fills them) lives in slot u & 7, fill u >> 3.
                static_assert(SBT == 8, "unit slot arithmetic");
                uint32_t own = 0, useq = 0;
                int4 pi_next = first_patch < g.npatch ? pinfo[first_patch] : make_int4(0, 1, 0, 0);
                for (int patch = first_patch; patch < g.npatch; patch += patch_step) {
                    const int4 pi = pi_next;
                    if (patch + patch_step < g.npatch) pi_next = pinfo[patch + patch_step];    // in flight while this patch is multiplied
                    const int py = patch / g.pcols;
                    const int mtc = min(G, (g.band_rows - py * PR + 1) >> 1);
                    const int nb = pi.y;
                    const int kk_last = (pi.z - SLOTS * (nb - 1)) > 16 ? 2 : 1;
                    for (int mt = 0; mt < mtc; ++mt) {
                        uint64_t da[SA / 2];
                        uint32_t abar[SA / 2];
#pragma unroll
                        for (int b = 0; b < SA / 2; ++b) {
                            if (b >= nb) break;
                            int as = stage + b;
                            uint32_t aph = phase;
                            if (as >= SA) { as -= SA; aph ^= 1u; }
                            mbar_wait(bar_afull + 8 * as, aph, err, 3);
                            da[b] = make_smem_desc_k<32>(smem_u32(smem_a + as * A_TILE_BYTES));
                            abar[b] = bar_aempty + 8 * as;
                        }
                        for (int nt = me; nt < n_tiles; nt += 2) {
                            const uint32_t u0 = useq + (uint32_t)(nt * nb);
                            if (mt == 0) {
#pragma unroll
                                for (int b = 0; b < SA / 2; ++b)
                                    if (b < nb) mbar_wait(bar_bfull + 8 * ((u0 + b) & 7u), ((u0 + b) >> 3) & 1u, err, 6);
                            }
                            PT_T(1);
                            mbar_wait(bar_tempty + 8 * me, (own & 1u) ^ 1u, err, 2);
                            tcgen05_fence_after();
                            PT_T(0);
                            const uint32_t d_tmem = tmem_base + (uint32_t)(me * 256);
#pragma unroll
                            for (int b = 0; b < SA / 2; ++b) {
                                if (b >= nb) break;
                                const uint64_t db = make_smem_desc_mn(smem_u32(smem_b + ((u0 + b) & 7u) * B_BLOCK_BYTES), (uint32_t)B_CHUNK_BYTES);
                                const int kk = b == nb - 1 ? kk_last : 2;
                                for (int k = 0; k < kk; ++k)
                                    umma_f16(d_tmem, da[b] + (uint64_t)(2 * k), db + (uint64_t)(128 * k), idesc, (uint32_t)((b | k) != 0));
                            }
                            PT_T(2);
                            umma_commit(bar_tfull + 8 * me);
                            ++own;
                            if (mt == mtc - 1) {    // the W units of this N tile go back after the patch's last M tile
#pragma unroll
                                for (int b = 0; b < SA / 2; ++b)
                                    if (b < nb) umma_commit(bar_bempty + 8 * ((u0 + b) & 7u));
                            }
                            PT_T(3);
#ifdef GLB200_PT_DEBUG
                            pa[5] += 1;
#endif
                        }
#pragma unroll
                        for (int b = 0; b < SA / 2; ++b)
                            if (b < nb) umma_commit(abar[b]);      // (the other issuer's arrival completes the release)
                        stage += nb;
                        if (stage >= SA) { stage -= SA; phase ^= 1u; }
                        PT_T(4);
                    }
                    useq += (uint32_t)(nb * n_tiles);
                }
            } else
            for (int patch = first_patch; patch < g.npatch; patch += patch_step) {
                const int4 pi = pinfo[patch];
                const int py = patch / g.pcols;
                const int mtc = min(G, (g.band_rows - py * PR + 1) >> 1);
                const int nb = pi.y;
                const bool resident = nb * 2 <= SA && nb * n_tiles <= SBT;
                const int b0 = bpos;
                uint32_t b_seen = 0;                   // resident W units this issuer has already waited for in this patch
                for (int mt = 0; mt < mtc; ++mt) {
                    const int stage_m = stage;         // first A slot of this M tile (resident)
                    const uint32_t phase_m = phase;
                    bool a_seen = false;
                    for (int nt = 0; nt < n_tiles; ++nt, ++it) {
                        const int acc = it & 1;
                        const bool mine = !two || acc == me;
                        if (mine) {
                            PT_T(4);   // everything between two own tiles that is not one of the phases below (loop, other issuer's tiles)
                            mbar_wait(bar_tempty + 8 * acc, (uint32_t)(((it >> 1) & 1) ^ 1), err, 2);
                            tcgen05_fence_after();
                            PT_T(0);
                        }
                        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
                        for (int b = 0; b < nb; ++b) {
                            int as, bs;
                            uint32_t aph;
                            if (resident) {
                                as = stage_m + b; aph = phase_m;
                                if (as >= SA) { as -= SA; aph ^= 1u; }
                                bs = b0 + nt * nb + b;
                                if (bs >= SBT) bs -= SBT;
                            } else {
                                as = stage; aph = phase;
                                bs = bpos;
                                if (++bpos == SBT) bpos = 0;
                            }
                            if (mine) {
                                if (!resident || !a_seen) mbar_wait(bar_afull + 8 * as, aph, err, 3);
                                if (!resident || !((b_seen >> (nt * nb + b)) & 1u)) mbar_wait(bar_bfull + 8 * bs, (buses >> bs) & 1u, err, 6);
                                tcgen05_fence_after();
                                PT_T(1);
                                const int kk = (pi.z - SLOTS * b) > 16 ? 2 : 1;     // a last block with at most 16 samples: one K step
                                const uint64_t da = make_smem_desc_k<32>(smem_u32(smem_a + as * A_TILE_BYTES));
                                const uint64_t db = make_smem_desc_mn(smem_u32(smem_b + bs * B_BLOCK_BYTES), (uint32_t)B_CHUNK_BYTES);
                                for (int k = 0; k < kk; ++k)
                                    umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(128 * k), idesc, (uint32_t)((b | k) != 0));
                                PT_T(2);
                                if (!resident) {
                                    umma_commit(bar_aempty + 8 * as);
                                    umma_commit(bar_aempty + 8 * as);
                                    umma_commit(bar_bempty + 8 * bs);
                                } else {
                                    b_seen |= 1u << (nt * nb + b);
                                    // the W unit goes back after the patch's last M tile (every use of it was this issuer's: in order)
                                    if (mt == mtc - 1) umma_commit(bar_bempty + 8 * bs);
                                    // the A slot after this issuer's last N tile of the M tile; the other issuer's arrival completes it
                                    const int my_last_nt = !two ? n_tiles - 1 : (((it - nt + n_tiles - 1) & 1) == me ? n_tiles - 1 : n_tiles - 2);
                                    if (nt == my_last_nt) {
                                        umma_commit(bar_aempty + 8 * as);
                                        if (!two) umma_commit(bar_aempty + 8 * as);
                                    }
                                }
                            }
                            if (!resident || mt == mtc - 1) buses ^= 1u << bs;
                            if (!resident) { if (++stage == SA) { stage = 0; phase ^= 1; } }
                        }
                        if (mine) {
                            a_seen = true;
                            umma_commit(bar_tfull + 8 * acc);
                            PT_T(3);
#ifdef GLB200_PT_DEBUG
                            pa[5] += 1;
#endif
                        }
                    }
                    if (resident) { stage += nb; if (stage >= SA) { stage -= SA; phase ^= 1; } }
                }
                if (resident) { bpos = b0 + nb * n_tiles; if (bpos >= SBT) bpos -= SBT; }
            }
#ifdef GLB200_PT_DEBUG
            if (prof && blockIdx.x == 0) for (int k = 0; k < 6; ++k) prof[me * 8 + k] = pa[k];
#endif
#undef PT_T
        }
    } else if (warp == 2) {
        // ===== B gather: rows of W for the slots of a block and the columns of an N tile, in the MN-major SWIZZLE_128B layout =====
        const int u = lane;                      // 16-byte unit of the 512-byte row part of an N tile
        int bpos = 0;
        uint32_t bloads = 0;
        for (int patch = first_patch; patch < g.npatch; patch += patch_step) {
            const int4 pi = pinfo[patch];
            const int py = patch / g.pcols;
            const int mtc = min(G, (g.band_rows - py * PR + 1) >> 1);
            const bool resident = pi.y * 2 <= SA && pi.y * n_tiles <= SBT;
            const int reps = resident ? 1 : mtc;
            for (int rep = 0; rep < reps; ++rep)
                for (int nt = 0; nt < n_tiles; ++nt)
                    for (int b = 0; b < pi.y; ++b) {
                        mbar_wait(bar_bempty + 8 * bpos, ((bloads >> bpos) & 1u) ^ 1u, err, 5);
                        bloads ^= 1u << bpos;
                        const uint32_t dst_s = smem_u32(smem_b + bpos * B_BLOCK_BYTES);
                        const uint32_t my_slot = __ldg(slots + (size_t)(pi.x + b) * SLOTS + lane);
                        // thirty-two 16-byte cp.async per thread, all in flight at once (no registers hold the data); an empty slot
                        // and the chunks beyond a narrow N tile are zero-filled (source size 0)
#pragma unroll 8
                        for (int k = 0; k < SLOTS; ++k) {
                            const uint32_t s = __shfl_sync(0xffffffffu, my_slot, k);
                            const bool live = s != 0xffffffffu && u * 8 < BN;
                            const __half* src = W + (size_t)(live ? s : 0) * m_pad + (size_t)nt * BN + (live ? u * 8 : 0);
                            const uint32_t d = dst_s + (uint32_t)((u >> 3) * B_CHUNK_BYTES + k * 128 + (((u & 7) ^ (k & 7)) << 4));
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(live ? 16 : 0) : "memory");
                        }
                        asm volatile("cp.async.wait_all;" ::: "memory");
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_bfull + 8 * bpos);
                        if (++bpos == SBT) bpos = 0;
                    }
        }
    } else {
        // ===== epilogue =====
        const int wq = warp & 3, share = (warp - 4) >> 2;
        constexpr int COLS = BN >= 128 ? BN / 2 : BN;          // columns per share (BN = 64: share 0 takes them all)
        const bool active = BN >= 128 || share == 0;
        constexpr int NCH = COLS / 32;
        constexpr bool AT_ONCE = FC == 1;
        const uint32_t wbase = smem_u32(w_s) + (uint32_t)((BN >= 128 ? share * COLS : 0) * FC * 4);
        auto mul32 = [&](const uint32_t (&v)[32], uint32_t wv, float (&dot)[FC][8]) {
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
                float wr[8 * FC];
#pragma unroll
                for (int q = 0; q < 2 * FC; ++q) lds_f4(wv + (uint32_t)((g8 * 2 * FC + q) * 16), &wr[4 * q]);
                if (FC == 1) {
#pragma unroll
                    for (int i = 0; i < 8; i += 2)
                        ffma2(dot[0][i], dot[0][i + 1], __uint_as_float(v[8 * g8 + i]), __uint_as_float(v[8 * g8 + i + 1]), wr[i], wr[i + 1]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float val = __uint_as_float(v[8 * g8 + i]);
#pragma unroll
                        for (int q = 0; q < FC; ++q) dot[q][i & 3] = fmaf(val, wr[i * FC + q], dot[q][i & 3]);
                    }
                }
            }
        };
        float* my_x = xch + (wq * 32 + lane) * FC;
#ifdef GLB200_PT_DEBUG
        long long ea[6] = {0, 0, 0, 0, 0, 0}, et0 = 0, et1;   // tfull wait, tcgen05.ld (issue + wait), hand-back, arithmetic, end of M tile, tiles
#define PE_T(k) do { if (prof) { et1 = clock64(); ea[k] += et1 - et0; et0 = et1; } } while (0)
        if (prof) et0 = clock64();
#else
#define PE_T(k) do { } while (0)
#endif
        int it = 0;
        for (int patch = first_patch; patch < g.npatch; patch += patch_step) {
            const int py = patch / g.pcols, pxi = patch - py * g.pcols;
            const int mtc = min(G, (g.band_rows - py * PR + 1) >> 1);
            for (int mt = 0; mt < mtc; ++mt) {
                float dot[FC][8];
#pragma unroll
                for (int q = 0; q < FC; ++q)
#pragma unroll
                    for (int i = 0; i < 8; ++i) dot[q][i] = 0.f;
                // this thread's pixel: tile pixel wq * 32 + lane -> patch row 2 mt + (wq >> 1), column (wq & 1) * 32 + lane; its input
                // value is requested now, long before it is needed (an L2 round trip at the end of the M tile would be on the chain)
                const int r = py * PR + 2 * mt + (wq >> 1), c = pxi * PW + (wq & 1) * 32 + lane;
                const bool px_ok = share == 0 && r < g.band_rows && c < g.width;
                const size_t o = ((size_t)r * g.width + c) * FC;
                float yv[FC];
#pragma unroll
                for (int q = 0; q < FC; ++q) yv[q] = px_ok ? (float)__ldg(img + o + q) : 0.f;
                for (int nt = 0; nt < n_tiles; ++nt, ++it) {
                    const int acc = it & 1;
                    mbar_wait(bar_tfull + 8 * acc, (uint32_t)((it >> 1) & 1), err, 4);
                    tcgen05_fence_after();
                    PE_T(0);
                    const uint32_t t_row = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * 256 + (BN >= 128 ? share * COLS : 0));
                    const uint32_t wt = wbase + (uint32_t)(nt * BN * FC * 4);
                    if (AT_ONCE) {
                        uint32_t v[NCH][32];
                        if (active) {
#pragma unroll
                            for (int k = 0; k < NCH; ++k) tmem_ld_32x32b_x32(t_row + (uint32_t)(32 * k), v[k]);
                        }
                        tmem_ld_wait();
                        PE_T(1);
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
                        PE_T(2);
                        if (active && !PT_DBG(1)) {
#pragma unroll
                            for (int k = 0; k < NCH; ++k) mul32(v[k], wt + (uint32_t)(32 * k * FC * 4), dot);
                        }
                        PE_T(3);
#ifdef GLB200_PT_DEBUG
                        ea[5] += 1;
#endif
                    } else {
                        uint32_t v[2][32];
                        if (active) tmem_ld_32x32b_x32(t_row, v[0]);
#pragma unroll
                        for (int k = 0; k < NCH; ++k) {
                            tmem_ld_wait();
                            if (k + 1 < NCH) {
                                if (active) tmem_ld_32x32b_x32(t_row + (uint32_t)(32 * (k + 1)), v[(k + 1) & 1]);
                            } else {
                                // every tcgen05.ld of this accumulator has completed: hand it back before the last chunk's arithmetic
                                tcgen05_fence_before();
                                __syncwarp();
                                if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
                            }
                            if (active && !PT_DBG(1)) mul32(v[k & 1], wt + (uint32_t)(32 * k * FC * 4), dot);
                        }
                    }
                }
                // the pixel's dot over all columns: share 1 hands its part to share 0 (same quarter, same lane), which stores
                // z = min(y + dot, 255) (AboveXSetY(z, 255, 255), display.c:76) -- fixed summation order, bit-reproducible
                float part[FC];
#pragma unroll
                for (int q = 0; q < FC; ++q)
                    part[q] = ((dot[q][0] + dot[q][1]) + (dot[q][2] + dot[q][3])) + ((dot[q][4] + dot[q][5]) + (dot[q][6] + dot[q][7]));
                if (share == 1) {
#pragma unroll
                    for (int q = 0; q < FC; ++q) my_x[q] = part[q];
                }
                asm volatile("bar.sync %0, 64;" ::"r"(1 + wq) : "memory");
                if (px_ok && !PT_DBG(2)) {
#pragma unroll
                    for (int q = 0; q < FC; ++q) {
                        float val = fminf(yv[q] + (part[q] + my_x[q]), 255.f);
                        if (clip_low) val = fmaxf(val, 0.f);
                        z[o + q] = val;
                        if (z8) z8[o + q] = (uint8_t)fminf(fmaxf(val, 0.f), 255.f);
                    }
                }
                PE_T(4);
            }
        }
#ifdef GLB200_PT_DEBUG
        if (prof && blockIdx.x == 0 && lane == 0 && (warp == 4 || warp == 8)) for (int k = 0; k < 6; ++k) prof[16 + (warp == 8) * 8 + k] = ea[k];
#endif
#undef PE_T
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// The same extrapolation + filter as TWO independent pipelines per SM (every patch resident; the kernel above stays for the rest).
// Pipeline g = {issuer g, accumulator g, epilogue group g of eight warps} takes every other M tile of the CTA's patches with all its N
// tiles: a pixel's dot product never leaves its group (no rendezvous between the groups), and while group g multiplies the tile it
// has just read, its accumulator is already being refilled -- by its own issuer, whose chain (wake-up, MMA, commit, wake-up of the
// epilogue) no longer has to fit into the other accumulator's epilogue.  What the two pipelines share: the A ring (a slot belongs to
// one M tile, hence to one issuer), the W units of the patch (released when both issuers are through with the patch) and the SM's
// FMA and shared-memory pipes -- the bound that is left (profiles/r02_patch_timeline.md).
// 20 warps: 0 producer, 1 and 3 issuers, 2 tensor-memory allocation + W gather, 4-11 epilogue group 0, 12-19 epilogue group 1.
// ---------------------------------------------------------------------------------------------
constexpr int THREADS2 = 128 + 2 * 32 * EPI_WARPS;

template <int FC, int BN>
__global__ void __launch_bounds__(THREADS2, 1)
k_patch_nystroem_dual(const __grid_constant__ CUtensorMap map_a, Geom g, const int4* __restrict__ pinfo, const uint32_t* __restrict__ slots,
                      const __half* __restrict__ W, int m_pad, int n_tiles, const float* __restrict__ scales,
                      const float* __restrict__ fuse_w /* [m_pad][FC] */, const uint8_t* __restrict__ img /* band rows, [pixel][FC] */,
                      int clip_low, float* __restrict__ z /* [band pixels][FC] */, uint8_t* __restrict__ z8 /* or null */,
                      int* __restrict__ err, const int* __restrict__ dstat, long long* __restrict__ prof /* PT_DEBUG: phase clocks of CTA 0 */)
{
    using namespace tc;
    static_assert(SBT == 8, "unit slot arithmetic");
    if (dstat[GL_DS_PT_OVERFLOW]) return;
    const int max_nb = dstat[GL_DS_PT_MAXNB];
    if (!(max_nb * 2 <= SA && max_nb * n_tiles <= SBT)) return;      // not every patch resident: the general kernel runs instead
    extern __shared__ uint8_t pn_smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)pn_smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem_a + SA * A_TILE_BYTES;
    uint64_t* bars = (uint64_t*)(smem_b + SBT * B_BLOCK_BYTES);
    const uint32_t bar_afull = smem_u32(bars), bar_aempty = smem_u32(bars + SA);
    const uint32_t bar_bfull = smem_u32(bars + 2 * SA), bar_bempty = smem_u32(bars + 2 * SA + SBT);
    const uint32_t bar_tfull = smem_u32(bars + 2 * SA + 2 * SBT), bar_tempty = smem_u32(bars + 2 * SA + 2 * SBT + 2);
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * SA + 2 * SBT + 4);
    float* xch = (float*)(bars + 64);            // [2 slots][2 groups][4 quarters][32 lanes][FC]
    float* w_s = xch + 2 * 2 * 4 * 32 * FC;          // [m_pad][FC] filter weights, times the GEMM's output scale

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int first_patch = blockIdx.x, patch_step = gridDim.x;

    if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < SA; ++s) { mbar_init(bar_afull + 8 * s, 1); mbar_init(bar_aempty + 8 * s, 1); }      // a slot belongs to one issuer
        for (int s = 0; s < SBT; ++s) { mbar_init(bar_bfull + 8 * s, 1); mbar_init(bar_bempty + 8 * s, 2); }     // a W unit to both
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {
        const float sc = scales[1];
        for (int i = threadIdx.x; i < m_pad * FC; i += THREADS2) w_s[i] = __ldg(fuse_w + i) * sc;
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== A producer: the tiles of every M tile, in order =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int patch = first_patch; patch < g.npatch; patch += patch_step) {
                const int4 pi = pinfo[patch];
                const int py = patch / g.pcols;
                const int mtc = min(G, (g.band_rows - py * PR + 1) >> 1);
                const int tile0 = pi.x * G;
                for (int mt = 0; mt < mtc; ++mt)
                    for (int b = 0; b < pi.y; ++b) {
                        mbar_wait(bar_aempty + 8 * stage, phase ^ 1, err, 1);
                        mbar_expect_tx(bar_afull + 8 * stage, (uint32_t)A_TILE_BYTES);
                        tma_load_2d(smem_u32(smem_a + stage * A_TILE_BYTES), &map_a, bar_afull + 8 * stage, 0, (tile0 + mt * pi.y + b) * 128);
                        if (++stage == SA) { stage = 0; phase ^= 1; }
                    }
            }
        }
    } else if (warp == 1 || warp == 3) {
        // ===== MMA issuers: issuer `me` owns accumulator `me` and the M tiles of its parity =====
        if (lane == 0) {
            const int me = warp == 1 ? 0 : 1;
            const uint32_t idesc = make_idesc(BLOCK_M, BN, 0) | (1u << 16);   // B is MN-major
            const uint32_t d_tmem = tmem_base + (uint32_t)(me * 256);
            uint32_t own = 0, useq = 0, mseq = 0;
            int stage = 0;
            uint32_t phase = 0;
#ifdef GLB200_PT_DEBUG
            long long pa[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pt0 = 0, pt1;
#define PT_T(k) do { if (prof) { pt1 = clock64(); pa[k] += pt1 - pt0; pt0 = pt1; } } while (0)
            if (prof) pt0 = clock64();
#else
#define PT_T(k) do { } while (0)
#endif
            int4 pi_next = first_patch < g.npatch ? pinfo[first_patch] : make_int4(0, 1, 0, 0);
            for (int patch = first_patch; patch < g.npatch; patch += patch_step) {
                const int4 pi = pi_next;
                if (patch + patch_step < g.npatch) pi_next = pinfo[patch + patch_step];
                const int py = patch / g.pcols;
                const int mtc = min(G, (g.band_rows - py * PR + 1) >> 1);
                const int nb = pi.y;
                const int kk_last = (pi.z - SLOTS * (nb - 1)) > 16 ? 2 : 1;
                bool touched = false;
                for (int mt = 0; mt < mtc; ++mt, ++mseq) {
                    if ((mseq & 1u) == (uint32_t)me) {
                        uint64_t da[SA / 2];
                        uint32_t abar[SA / 2];
                        PT_T(7);
#pragma unroll
                        for (int b = 0; b < SA / 2; ++b) {
                            if (b >= nb) break;
                            int as = stage + b;
                            uint32_t aph = phase;
                            if (as >= SA) { as -= SA; aph ^= 1u; }
                            mbar_wait(bar_afull + 8 * as, aph, err, 3);
                            da[b] = make_smem_desc_k<32>(smem_u32(smem_a + as * A_TILE_BYTES));
                            abar[b] = bar_aempty + 8 * as;
                        }
                        PT_T(6);
                        for (int nt = 0; nt < n_tiles; ++nt) {
                            const uint32_t u0 = useq + (uint32_t)(nt * nb);
                            uint64_t db[SA / 2];
#pragma unroll
                            for (int b = 0; b < SA / 2; ++b) {
                                if (b >= nb) break;
                                if (!touched) mbar_wait(bar_bfull + 8 * ((u0 + b) & 7u), ((u0 + b) >> 3) & 1u, err, 6);
                                db[b] = make_smem_desc_mn(smem_u32(smem_b + ((u0 + b) & 7u) * B_BLOCK_BYTES), (uint32_t)B_CHUNK_BYTES);
                            }
                            PT_T(1);
                            mbar_wait(bar_tempty + 8 * me, (own & 1u) ^ 1u, err, 2);
                            tcgen05_fence_after();
                            PT_T(0);
#pragma unroll
                            for (int b = 0; b < SA / 2; ++b) {
                                if (b >= nb) break;
                                const int kk = b == nb - 1 ? kk_last : 2;
                                for (int k = 0; k < kk; ++k)
                                    umma_f16(d_tmem, da[b] + (uint64_t)(2 * k), db[b] + (uint64_t)(128 * k), idesc, (uint32_t)((b | k) != 0));
                            }
                            PT_T(2);
                            umma_commit(bar_tfull + 8 * me);
                            ++own;
                            PT_T(3);
#ifdef GLB200_PT_DEBUG
                            pa[5] += 1;
#endif
                        }
                        touched = true;
#pragma unroll
                        for (int b = 0; b < SA / 2; ++b)
                            if (b < nb) umma_commit(abar[b]);
                    }
                    stage += nb;
                    if (stage >= SA) { stage -= SA; phase ^= 1u; }
                }
                // The W units go back when BOTH issuers are through with the patch.  tcgen05.commit arrives once this thread's MMAs are
                // complete; an issuer that had no M tile in the patch first makes sure the unit was filled (its arrival must land in
                // this fill's phase, not in the previous one's), then arrives plainly.
                const uint32_t nu = (uint32_t)(nb * n_tiles);
                for (uint32_t u = useq; u < useq + nu; ++u) {
                    if (touched) umma_commit(bar_bempty + 8 * (u & 7u));
                    else {
                        mbar_wait(bar_bfull + 8 * (u & 7u), (u >> 3) & 1u, err, 6);
                        mbar_arrive(bar_bempty + 8 * (u & 7u));
                    }
                }
                useq += nu;
                PT_T(4);
            }
#ifdef GLB200_PT_DEBUG
            if (prof && blockIdx.x == 0) for (int q = 0; q < 8; ++q) prof[me * 8 + q] = pa[q];
#endif
#undef PT_T
        }
    } else if (warp == 2) {
        // ===== B gather (as in the kernel above, resident case) =====
        const int u = lane;
        int bpos = 0;
        uint32_t bloads = 0;
        for (int patch = first_patch; patch < g.npatch; patch += patch_step) {
            const int4 pi = pinfo[patch];
            for (int nt = 0; nt < n_tiles; ++nt)
                for (int b = 0; b < pi.y; ++b) {
                    mbar_wait(bar_bempty + 8 * bpos, ((bloads >> bpos) & 1u) ^ 1u, err, 5);
                    bloads ^= 1u << bpos;
                    const uint32_t dst_s = smem_u32(smem_b + bpos * B_BLOCK_BYTES);
                    const uint32_t my_slot = __ldg(slots + (size_t)(pi.x + b) * SLOTS + lane);
#pragma unroll 8
                    for (int k = 0; k < SLOTS; ++k) {
                        const uint32_t s = __shfl_sync(0xffffffffu, my_slot, k);
                        const bool live = s != 0xffffffffu && u * 8 < BN;
                        const __half* src = W + (size_t)(live ? s : 0) * m_pad + (size_t)nt * BN + (live ? u * 8 : 0);
                        const uint32_t d = dst_s + (uint32_t)((u >> 3) * B_CHUNK_BYTES + k * 128 + (((u & 7) ^ (k & 7)) << 4));
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(live ? 16 : 0) : "memory");
                    }
                    asm volatile("cp.async.wait_all;" ::: "memory");
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_bfull + 8 * bpos);
                    if (++bpos == SBT) bpos = 0;
                }
        }
    } else {
        // ===== epilogue groups =====
        const int grp = (warp - 4) >> 3, wq = warp & 3, share = ((warp - 4) >> 2) & 1;
        constexpr int COLS = BN >= 128 ? BN / 2 : BN;          // columns per share (BN = 64: share 0 takes them all)
        const bool active = BN >= 128 || share == 0;
        constexpr int NCH = COLS / 32;
        const uint32_t wbase = smem_u32(w_s) + (uint32_t)((BN >= 128 ? share * COLS : 0) * FC * 4);
        auto mul32 = [&](const uint32_t (&v)[32], uint32_t wv, float (&dot)[FC][8]) {
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
                float wr[8 * FC];
#pragma unroll
                for (int q = 0; q < 2 * FC; ++q) lds_f4(wv + (uint32_t)((g8 * 2 * FC + q) * 16), &wr[4 * q]);
                if (FC == 1) {
#pragma unroll
                    for (int i = 0; i < 8; i += 2)
                        ffma2(dot[0][i], dot[0][i + 1], __uint_as_float(v[8 * g8 + i]), __uint_as_float(v[8 * g8 + i + 1]), wr[i], wr[i + 1]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float val = __uint_as_float(v[8 * g8 + i]);
#pragma unroll
                        for (int q = 0; q < FC; ++q) dot[q][i & 3] = fmaf(val, wr[i * FC + q], dot[q][i & 3]);
                    }
                }
            }
        };
        float* my_x = xch + ((grp * 4 + wq) * 32 + lane) * FC;
        uint32_t own = 0, mseq = 0, mslot = 0;
#ifdef GLB200_PT_DEBUG
        long long ea[6] = {0, 0, 0, 0, 0, 0}, et0 = 0, et1;
#define PE_T(k) do { if (prof) { et1 = clock64(); ea[k] += et1 - et0; et0 = et1; } } while (0)
        if (prof) et0 = clock64();
#else
#define PE_T(k) do { } while (0)
#endif
        for (int patch = first_patch; patch < g.npatch; patch += patch_step) {
            const int py = patch / g.pcols, pxi = patch - py * g.pcols;
            const int mtc = min(G, (g.band_rows - py * PR + 1) >> 1);
            for (int mt = 0; mt < mtc; ++mt, ++mseq) {
                if ((mseq & 1u) != (uint32_t)grp) continue;
                float dot[FC][8];
#pragma unroll
                for (int q = 0; q < FC; ++q)
#pragma unroll
                    for (int i = 0; i < 8; ++i) dot[q][i] = 0.f;
                const int r = py * PR + 2 * mt + (wq >> 1), c = pxi * PW + (wq & 1) * 32 + lane;
                const bool px_ok = share == 0 && r < g.band_rows && c < g.width;
                const size_t o = ((size_t)r * g.width + c) * FC;
                // the pixel's input value: requested now as a raw byte and converted only when z is formed -- a conversion right here
                // would wait for the load (an L2 / HBM round trip of ~1 000 cycles at the start of every M tile: it showed up as 415
                // cycles per tile in the phase clocks)
                uint32_t yraw[FC];
#pragma unroll
                for (int q = 0; q < FC; ++q) {
                    yraw[q] = 0u;
                    if (px_ok) asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(yraw[q]) : "l"(img + o + q));
                }
                for (int nt = 0; nt < n_tiles; ++nt, ++own) {
                    PE_T(4);
                    mbar_wait(bar_tfull + 8 * grp, own & 1u, err, 4);
                    tcgen05_fence_after();
                    PE_T(0);
                    const uint32_t t_row = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(grp * 256 + (BN >= 128 ? share * COLS : 0));
                    const uint32_t wt = wbase + (uint32_t)(nt * BN * FC * 4);
                    uint32_t v[2][32];
                    if (active) tmem_ld_32x32b_x32(t_row, v[0]);
#pragma unroll
                    for (int k = 0; k < NCH; ++k) {
                        tmem_ld_wait();
                        if (k + 1 < NCH) {
                            if (active) tmem_ld_32x32b_x32(t_row + (uint32_t)(32 * (k + 1)), v[(k + 1) & 1]);
                        } else {
                            // every tcgen05.ld of this accumulator has completed: hand it back before the last chunk's arithmetic
                            PE_T(1);
                            tcgen05_fence_before();
                            if (lane == 0) mbar_arrive(bar_tempty + 8 * grp);
                            PE_T(2);
                        }
                        if (active) mul32(v[k & 1], wt + (uint32_t)(32 * k * FC * 4), dot);
                    }
                    PE_T(3);
#ifdef GLB200_PT_DEBUG
                    ea[5] += 1;
#endif
                }
                float part[FC];
#pragma unroll
                for (int q = 0; q < FC; ++q)
                    part[q] = ((dot[q][0] + dot[q][1]) + (dot[q][2] + dot[q][3])) + ((dot[q][4] + dot[q][5]) + (dot[q][6] + dot[q][7]));
                // share 1 hands its part to share 0 through one of two alternating slots: share 1 writes a slot again two M tiles later,
                // i.e. after the next barrier, which share 0 reaches only after it has read this one
                float* xs = my_x + (mslot & 1u) * (2 * 4 * 32 * FC);
                ++mslot;
                if (share == 1) {
#pragma unroll
                    for (int q = 0; q < FC; ++q) xs[q] = part[q];
                }
                asm volatile("bar.sync %0, 64;" ::"r"(1 + grp * 4 + wq) : "memory");
                if (px_ok) {
#pragma unroll
                    for (int q = 0; q < FC; ++q) {
                        float val = fminf((float)(yraw[q] & 0xffu) + (part[q] + xs[q]), 255.f);
                        if (clip_low) val = fmaxf(val, 0.f);
                        z[o + q] = val;
                        if (z8) z8[o + q] = (uint8_t)fminf(fmaxf(val, 0.f), 255.f);
                    }
                }
            }
        }
#ifdef GLB200_PT_DEBUG
        if (prof && blockIdx.x == 0 && lane == 0 && (warp == 4 || warp == 12)) for (int q = 0; q < 6; ++q) prof[16 + (warp == 12) * 8 + q] = ea[q];
#endif
#undef PE_T
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

constexpr int nystroem_smem(int fc, int m_pad) { return SA * A_TILE_BYTES + SBT * B_BLOCK_BYTES + 512 + 4 * 4 * 32 * fc * 4 + m_pad * fc * 4 + 1024; }

// K_B from the patch layout to dense fp64 [band pixels][p] in the caller's sample order (dst zeroed beforehand)
__global__ void k_patch_to_f64(Geom g, const int4* __restrict__ pinfo, const uint32_t* __restrict__ slots, const __half* __restrict__ KB, int p,
                               double scale, double* __restrict__ dst)
{
    const int patch = blockIdx.x;
    const int4 pi = pinfo[patch];
    const int py = patch / g.pcols, pxi = patch - py * g.pcols;
    for (int e = threadIdx.x; e < PW * PR * pi.y * SLOTS; e += blockDim.x) {
        const int sl = e % (pi.y * SLOTS), pix = e / (pi.y * SLOTS);
        const int r = py * PR + (pix >> 6), c = pxi * PW + (pix & 63);
        if (r >= g.band_rows || c >= g.width) continue;
        const uint32_t s = slots[(size_t)pi.x * SLOTS + sl];
        if (s == 0xffffffffu) continue;
        const int mt = (pix >> 6) >> 1, b = sl / SLOTS;
        const size_t tile = (size_t)pi.x * G + (size_t)mt * pi.y + b;
        const __half v = KB[(tile * 128 + (size_t)(((pix >> 6) & 1) * PW + (pix & 63))) * SLOTS + (sl % SLOTS)];
        dst[((size_t)r * g.width + c) * p + s] = scale * (double)__half2float(v);
    }
}

}  // namespace pt

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static pt::Geom patch_geom(const gl_ctx* ctx, double h_loc)
{
    pt::Geom g;
    g.width = ctx->width;
    g.row0 = ctx->row0;
    g.band_rows = ctx->row1 - ctx->row0;
    g.pcols = (int)ceil_div(ctx->width, pt::PW);
    g.prows = (int)ceil_div(g.band_rows, pt::PR);
    g.npatch = g.pcols * g.prows;
    // exp(-d^2 / h_loc^2) < 2^-25  <=>  d^2 > 25 ln 2 h_loc^2 ; one pixel of slack on the radius
    const double rc = h_loc * std::sqrt(25.0 * 0.6931471805599453) + 1.0;
    g.rc2 = rc < 3e4 ? (float)(rc * rc) : 9e8f;
    return g;
}

// whether the patch path serves this affinity on this context
bool gl_patch_applicable(const gl_ctx* ctx, int kind)
{
    return ctx->kb_layout == 1 && ctx->kb_cutoff && (kind == GL_BILATERAL || kind == GL_SPATIAL) && ctx->gemm_impl == 0;
}

template <int KIND, int C>
static int launch_patch_affinity(gl_ctx* ctx, const pt::Geom& g, double h_loc, double h_val, const float* sf, gl_mat* KB, float* partial, int grid)
{
    const int p_pad = ctx->p_pad;
    const size_t smem = sizeof(float) * ((size_t)(1 + C) * p_pad + (size_t)(1 + C) * 8 * pt::SLOTS + (size_t)2 * C * pt::PW * pt::PR);
    GL_CUDA_CHECK(cudaFuncSetAttribute(pt::k_patch_affinity<KIND, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const float log2e = 1.4426950408889634f;
    StageTimer kt(ctx, GL_T_K_AFFINITY_B);
    pt::k_patch_affinity<KIND, C><<<grid, 256, smem, ctx->stream>>>(g, (const uint8_t*)ctx->img->ptr, sf, p_pad, (float)(-log2e / (h_loc * h_loc)),
                                                                   (float)(-log2e / (h_val * h_val)), (const int4*)KB->pt_info->ptr,
                                                                   (const uint32_t*)KB->pt_slots->ptr,
                                                                   // (measurement switch: the sums-only pass a fused affinity + extrapolation
                                                                   // kernel would still need before the eigensolve; z is garbage with it)
                                                                   getenv("GLB200_PT_SUMS_ONLY") ? nullptr : (__half*)KB->pt_buf->ptr, partial,
                                                                   (const int*)ctx->dstat->ptr);
    GL_LAUNCH_CHECK(ctx);
    return GL_OK;
}

// Fills the patch-layout part of the K_B handle (lists, tiles) and the band-partial sums D / T into KB->aux (not yet reduced over
// ranks: the caller adds K_A y and runs the allreduce, as for the blocked layout).
int gl_patch_affinity(gl_ctx* ctx, int kind, double h_loc, double h_val, gl_mat* KB)
{
    const int p = (int)ctx->p, p_pad = ctx->p_pad, C = ctx->channels;
    const pt::Geom g = patch_geom(ctx, h_loc);
    {
        const size_t smem = sizeof(float) * ((size_t)(1 + C) * p_pad + (size_t)(1 + C) * 8 * pt::SLOTS + (size_t)2 * C * pt::PW * pt::PR);
        if (smem > 200 * 1024) {
            gl_set_error("affinity: p = %d samples need %zu bytes of shared memory per CTA", p, smem);
            return GL_ERR_UNSUPPORTED;
        }
    }
    gl_buf *total = nullptr, *sf = nullptr, *partial = nullptr;
    int rc = GL_OK;
    do {
        GL_BREAK(rc, gl_alloc(ctx, sizeof(int4) * (size_t)g.npatch, &KB->pt_info));
        GL_BREAK(rc, gl_alloc(ctx, sizeof(int) * 4, &total));
        GL_BREAK(rc, gl_ensure_pinned(ctx, 64));
        const unsigned wgrid = (unsigned)ceil_div(g.npatch, 8);
        pt::k_patch_count<<<wgrid, 256, 0, ctx->stream>>>(g, (const uint32_t*)ctx->samples->ptr, p, (int4*)KB->pt_info->ptr);
        ctx->launches++;
        // Storage is sized by the number of slot blocks, which only the device knows.  A stand-alone gl_affinity reads it back (one
        // small round trip); inside gl_run_resident the count of the last run with the same geometry, sample count and reach, plus a
        // quarter, is set aside without asking -- the scan kernel checks it and raises the overflow word if it was not enough.
        const int64_t key[5] = {g.width, g.band_rows, g.row0, p, (int64_t)g.rc2};
        const bool have_cap = ctx->async_mode && ctx->pt_cap_blocks > 0 && !memcmp(key, ctx->pt_cap_key, sizeof(key));
        pt::k_patch_scan<<<1, 1024, 0, ctx->stream>>>(g, g.npatch, (int4*)KB->pt_info->ptr, (int*)total->ptr, have_cap ? ctx->pt_cap_blocks : 0,
                                                      (int*)ctx->dstat->ptr);
        ctx->launches++;
        int64_t blocks = ctx->pt_cap_blocks;
        if (!have_cap) {
            GL_CUDA_BREAK(rc, cudaMemcpyAsync(ctx->pinned, total->ptr, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            GL_CUDA_BREAK(rc, cudaStreamSynchronize(ctx->stream));
            blocks = *(const int*)ctx->pinned;
            KB->pt_ksteps = (int64_t) * (const unsigned long long*)((const int*)ctx->pinned + 2);
            if (blocks <= 0 || (blocks + blocks / 4 + 64) * pt::G * 128 >= 0x7fffffffll) {
                gl_set_error("affinity: %lld sample blocks do not fit 32-bit tile indices", (long long)blocks);
                rc = GL_ERR_UNSUPPORTED;
                break;
            }
            ctx->pt_cap_blocks = blocks + blocks / 4 + 64;
            memcpy(ctx->pt_cap_key, key, sizeof(key));
        }
        KB->pt_blocks = blocks;          // (with have_cap: the capacity; the exact count is in the status block)
        KB->pt_npatch = g.npatch;
        GL_BREAK(rc, gl_alloc(ctx, sizeof(uint32_t) * (size_t)blocks * pt::SLOTS, &KB->pt_slots));
        GL_BREAK(rc, gl_alloc(ctx, (size_t)blocks * pt::G * pt::A_TILE_BYTES, &KB->pt_buf));
        pt::k_patch_fill<<<wgrid, 256, 0, ctx->stream>>>(g, (const uint32_t*)ctx->samples->ptr, p, (const int4*)KB->pt_info->ptr,
                                                        (uint32_t*)KB->pt_slots->ptr, (const int*)ctx->dstat->ptr);
        ctx->launches++;
        GL_BREAK(rc, gl_alloc(ctx, sizeof(float) * (size_t)(2 + C) * p_pad, &sf));
        GL_BREAK(rc, gl_image_ready(ctx));      // from here on pixels are read
        pt::k_patch_sample_features<<<(unsigned)ceil_div(p_pad, 256), 256, 0, ctx->stream>>>((const uint8_t*)ctx->img->ptr,
                                                                                            (const uint32_t*)ctx->samples->ptr, p, p_pad,
                                                                                            ctx->width, C, (float*)sf->ptr);
        ctx->launches++;
        int grid = ctx->sm_count * (C == 1 ? 3 : 2);
        if (grid > g.npatch) grid = g.npatch;
        GL_BREAK(rc, gl_alloc(ctx, sizeof(float) * (size_t)grid * (1 + C) * p_pad, &partial));
#define PT_CASE(K, CC) \
    if (kind == K && C == CC) rc = launch_patch_affinity<K, CC>(ctx, g, h_loc, h_val, (const float*)sf->ptr, KB, (float*)partial->ptr, grid);
        PT_CASE(GL_BILATERAL, 1) else PT_CASE(GL_BILATERAL, 3) else PT_CASE(GL_SPATIAL, 1) else PT_CASE(GL_SPATIAL, 3)
        else { gl_set_error("affinity(patch): kind %d with %d channels is not served", kind, C); rc = GL_ERR_UNSUPPORTED; }
#undef PT_CASE
        if (rc != GL_OK) break;
        GL_CUDA_BREAK(rc, cudaMemsetAsync(KB->aux->ptr, 0, sizeof(double) * (size_t)(1 + 2 * C) * p_pad, ctx->stream));
        pt::k_patch_reduce<<<(unsigned)ceil_div((1 + C) * p_pad, 8), 256, 0, ctx->stream>>>((const float*)partial->ptr, grid, p, p_pad, 1 + C,
                                                                                           (double*)KB->aux->ptr);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) { gl_set_error("affinity(patch): kernel launch failed"); rc = GL_ERR_CUDA; }
    } while (0);
    if (total) gl_buf_release(total);
    if (sf) gl_buf_release(sf);
    if (partial) gl_buf_release(partial);
    return rc;
}

int gl_patch_download(gl_ctx* ctx, const gl_mat* KB, double scale, double* dst_dev)
{
    const pt::Geom g = patch_geom(ctx, KB->aff_h_loc);
    GL_REQUIRE(g.npatch == KB->pt_npatch && KB->q0 == ctx->q0, "K_B download: the handle does not belong to the current image geometry");
    GL_CUDA_CHECK(cudaMemsetAsync(dst_dev, 0, sizeof(double) * (size_t)KB->local_rows * KB->p, ctx->stream));
    pt::k_patch_to_f64<<<g.npatch, 256, 0, ctx->stream>>>(g, (const int4*)KB->pt_info->ptr, (const uint32_t*)KB->pt_slots->ptr,
                                                         (const __half*)KB->pt_buf->ptr, KB->p, scale, dst_dev);
    GL_LAUNCH_CHECK(ctx);
    return GL_OK;
}

// whether the fused patch kernel serves this spectrum width (its filter weights for all N tiles live in shared memory)
bool gl_patch_nystroem_fits(int m, int C) { return pt::nystroem_smem(C, gl_m_pad(m)) <= tc::SMEM_LIMIT; }

// extrapolation + fused filter over the patch layout: z (and z8) of the band's pixels are written by the kernel; the sample pixels'
// rows are patched afterwards by the caller (gl_filter_fused_finish with parts = 0)
int gl_patch_nystroem_filter(gl_ctx* ctx, const gl_mat* L_B, const float* U, int ldU, int m, const double* mu_inv, const float* scales,
                             const float* w, int C, int clip_low, float* z, uint8_t* z8)
{
    const int p = L_B->p, p_pad = L_B->p_pad, m_pad = gl_m_pad(m);
    const pt::Geom g = patch_geom(ctx, L_B->aff_h_loc);
    GL_REQUIRE(g.npatch == L_B->pt_npatch && L_B->q0 == ctx->q0, "nystroem(patch): K_B does not belong to the current image geometry");
    GL_REQUIRE(C == 1 || C == 3, "nystroem(patch): 1 or 3 channels");
    const int BN = m_pad < 256 ? m_pad : 256;
    const int n_tiles = m_pad / BN;
    const int SM = pt::nystroem_smem(C, m_pad);
    GL_REQUIRE(SM <= tc::SMEM_LIMIT, "nystroem(patch): %d eigenpairs x %d channels of filter weights do not fit in shared memory", m, C);
    gl_buf *Wr = nullptr, *err = nullptr, *prof = nullptr;
    const bool want_prof = getenv("GLB200_PT_PROF") != nullptr;   // debug (make PT_DEBUG=1): timeline of CTA 0's first tiles on stderr
    const int dbg = getenv("GLB200_PT_DBG") ? atoi(getenv("GLB200_PT_DBG")) : 0;
    int rc = GL_OK;
    do {
        if (want_prof) {
            GL_BREAK(rc, gl_alloc(ctx, sizeof(long long) * 64 * 9, &prof));
            GL_CUDA_BREAK(rc, cudaMemsetAsync(prof->ptr, 0, sizeof(long long) * 64 * 9, ctx->stream));
        }
        GL_BREAK(rc, gl_alloc(ctx, sizeof(__half) * (size_t)p_pad * m_pad, &Wr));
        GL_BREAK(rc, gl_alloc(ctx, sizeof(int) * 4, &err));
        GL_CUDA_BREAK(rc, cudaMemsetAsync(err->ptr, 0, sizeof(int) * 4, ctx->stream));
        dim3 wg((unsigned)ceil_div(p_pad, 32), (unsigned)ceil_div(m_pad, 32));
        pt::k_w_rows<<<wg, 256, 0, ctx->stream>>>(U, ldU, p, m, p_pad, m_pad, mu_inv, (const double*)L_B->dscale->ptr, scales, (__half*)Wr->ptr);
        ctx->launches++;
        int grid = ctx->sm_count;
        if (grid > g.npatch) grid = g.npatch;
        CUtensorMap map_a;
        GL_BREAK(rc, make_map_2d(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, L_B->pt_buf->ptr, (uint64_t)L_B->pt_blocks * pt::G * 128, pt::SLOTS,
                                 pt::SLOTS, pt::SLOTS, 128, CU_TENSOR_MAP_SWIZZLE_64B));
        const uint8_t* y = (const uint8_t*)ctx->img->ptr + (size_t)L_B->q0 * C;
        StageTimer kt(ctx, GL_T_K_GEMM);
        // One channel: two pipelines per SM when every patch is resident (decided on the device: the general kernel, launched right
        // after, returns at once in that case and does all the work otherwise).
        const bool dual = C == 1 && ctx->pt_dual && dbg == 0;
        if (dual) {
#define PT_LAUNCH2(BNN)                                                                                                                \
    do {                                                                                                                               \
        GL_CUDA_BREAK(rc, cudaFuncSetAttribute(pt::k_patch_nystroem_dual<1, BNN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM));   \
        pt::k_patch_nystroem_dual<1, BNN><<<grid, pt::THREADS2, SM, ctx->stream>>>(map_a, g, (const int4*)L_B->pt_info->ptr,           \
                                                                                   (const uint32_t*)L_B->pt_slots->ptr, (const __half*)Wr->ptr, \
                                                                                   m_pad, n_tiles, scales, w, y, clip_low, z, z8,      \
                                                                                   (int*)err->ptr, (const int*)ctx->dstat->ptr,        \
                                                                                   prof ? (long long*)prof->ptr : nullptr);            \
    } while (0)
            if (BN == 256) PT_LAUNCH2(256);
            else if (BN == 128) PT_LAUNCH2(128);
            else PT_LAUNCH2(64);
#undef PT_LAUNCH2
            if (rc != GL_OK) break;
            ctx->launches++;
        }
#define PT_LAUNCH(FC, BNN)                                                                                                          \
    do {                                                                                                                            \
        GL_CUDA_BREAK(rc, cudaFuncSetAttribute(pt::k_patch_nystroem<FC, BNN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM));    \
        pt::k_patch_nystroem<FC, BNN><<<grid, pt::THREADS, SM, ctx->stream>>>(map_a, g, (const int4*)L_B->pt_info->ptr,             \
                                                                              (const uint32_t*)L_B->pt_slots->ptr, (const __half*)Wr->ptr, \
                                                                              m_pad, n_tiles, scales, w, y, clip_low, z, z8,        \
                                                                              (int*)err->ptr, prof ? (long long*)prof->ptr : nullptr, \
                                                                              (const int*)ctx->dstat->ptr, dbg, dual ? 1 : 0);      \
    } while (0)
        if (C == 1 && BN == 256) PT_LAUNCH(1, 256);
        else if (C == 1 && BN == 128) PT_LAUNCH(1, 128);
        else if (C == 1) PT_LAUNCH(1, 64);
        else if (BN == 256) PT_LAUNCH(3, 256);
        else if (BN == 128) PT_LAUNCH(3, 128);
        else PT_LAUNCH(3, 64);
#undef PT_LAUNCH
        if (rc != GL_OK) break;
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) { gl_set_error("nystroem(patch): kernel launch failed"); rc = GL_ERR_CUDA; }
    } while (0);
    if (prof) {
        long long h[32];
        cudaMemcpyAsync(h, prof->ptr, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        for (int i = 0; i < 2; ++i)
            if (h[i * 8 + 5] > 0)
                fprintf(stderr, "[pt prof] issuer %d: %lld tiles; cycles per own tile: accumulator free %lld | operands %lld | mma issue %lld | commit %lld | rest %lld | (dual: A wait %lld, before it %lld)\n", i,
                        h[i * 8 + 5], h[i * 8 + 0] / h[i * 8 + 5], h[i * 8 + 1] / h[i * 8 + 5], h[i * 8 + 2] / h[i * 8 + 5], h[i * 8 + 3] / h[i * 8 + 5],
                        h[i * 8 + 4] / h[i * 8 + 5], h[i * 8 + 6] / h[i * 8 + 5], h[i * 8 + 7] / h[i * 8 + 5]);
        for (int i = 0; i < 2; ++i)
            if (h[16 + i * 8 + 5] > 0)
                fprintf(stderr, "[pt prof] epilogue warp %d: %lld tiles; cycles per tile: accumulator full %lld | loads (+ 3/4 of the arithmetic, dual kernel) %lld | hand back %lld | arithmetic (rest) %lld | end of M tile %lld\n",
                        4 + 4 * i, h[16 + i * 8 + 5], h[16 + i * 8 + 0] / h[16 + i * 8 + 5], h[16 + i * 8 + 1] / h[16 + i * 8 + 5],
                        h[16 + i * 8 + 2] / h[16 + i * 8 + 5], h[16 + i * 8 + 3] / h[16 + i * 8 + 5], h[16 + i * 8 + 4] / h[16 + i * 8 + 5]);
        gl_buf_release(prof);
    }
    if (Wr) gl_buf_release(Wr);
    if (err) gl_buf_release(err);
    return rc;
}
