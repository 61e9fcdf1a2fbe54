// a-2: affinity blocks K_A (p x p, fp64) and K_B (band pixels x p_pad, fp16, pixel-major) with the
// row sums D = K_A.1 + K_B.1 taken from the fp32 kernel values before rounding, and the
// image-weighted row sums T = [K_A K_B] y (one per channel) from the same fp32 values: the filter's
// projection c = Phi^T y = U^T y_S + W^T (K_B y_B) is then a p-sized product instead of a pass over Phi,
// and -- more important -- it is free of the fp16 rounding of W, which the heavy cancellation in Phi^T y
// (|c| ~ 1 against |y| ~ 1e4) would amplify to a ~1e-2 error of z - y (measured).
// Replaces ComputeAffinityMatrices / ComputeDistance / ComputeBilateralFilter,
// hpc/affinity.c:129-262,115-122,59-113 (photometric :8-17, spatial :19-57).
//
// Layout choices (B200-first, not the reference's):
//   * K_B is stored TRANSPOSED (one row per pixel, samples contiguous): it is the K-major "A" operand
//     of the extrapolation GEMM (nystroem_gemm.cu) and each pixel row is one 16-byte-vector store target.
//   * every band pixel gets a row, sample pixels included: their rows of Phi are overwritten with the
//     eigenvectors afterwards, which removes the reference's Permutation pass (hpc/utils.c:134-173), and
//     the row sum over ALL pixels is exactly rowsum(K_A) + rowsum(K_B) (hpc/laplacian.c:18-20).
//   * coordinates and grey values are integers: differences are exact in fp32, the two bandwidth
//     factors are applied to the exact squared distances, one ex2 per pair (the reference takes two exps
//     and multiplies, hpc/affinity.c:99,107,110 -- equal up to rounding).
//   * K_B is stored in BLOCKS of [512 pixels][64 sample slots] fp16 (a pixel row of a block = 128 contiguous bytes), and
//     only blocks that can hold a non-zero are stored.  The kernels use an internal sample order (column strip, then
//     raster index); for a 512-pixel tile the samples within reach R = floor(h_loc sqrt(25 ln 2)) + 1 in rows AND columns
//     are a few contiguous runs of that order, everything else has exp(-d^2/h_loc^2) < 2^-25, which fp16 storage flushes
//     to zero anyway.  The layout {tile table, block starts, permutation} is built on the host and cached
//     (kb_layout_for below).  Photometric affinity has no cutoff: every block is stored and the layout degenerates to
//     a dense blocked matrix.  At 4K / p=1000 / h_loc=40 this keeps ~10 % of the blocks: the kernel evaluations, the K_B
//     bytes and the extrapolation GEMM's K loop all shrink by that factor.
#include <algorithm>
#include <cmath>

#include "common.cuh"

#define AFF_THREADS 256
#define AFF_TP 512               // pixels per tile
#define AFF_PPT (AFF_TP / 32)    // pixels per thread per 64-sample chunk

// sample features in INTERNAL sample order (perm[i] = index of the i-th internal sample in the caller's ascending list),
// SoA with stride p_int: [0] row, [1] col, [2..2+C) values; slots without a sample carry 1e18 so that K == 0
__global__ void k_sample_features(const uint8_t* __restrict__ img, const uint32_t* __restrict__ samples, const uint32_t* __restrict__ perm,
                                  int p_int, int width, int channels, float* __restrict__ sf)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p_int) return;
    const uint32_t j = perm[i];
    if (j != 0xffffffffu) {
        uint32_t q = samples[j];
        sf[i] = (float)(q / width);
        sf[p_int + i] = (float)(q % width);
        for (int ch = 0; ch < channels; ++ch) sf[(2 + ch) * p_int + i] = (float)img[(size_t)q * channels + ch];
    } else {
        for (int k = 0; k < 2 + channels; ++k) sf[k * p_int + i] = 1e18f;
    }
}

// K_A in fp64 exactly as the reference writes it: exp(-d2/h_loc^2) * exp(-dv2/h_val^2)
template <int KIND, int C>
__global__ void k_affinity_A(const uint8_t* __restrict__ img, const uint32_t* __restrict__ samples, int p, int width,
                             double inv_hl2, double inv_hv2, double* __restrict__ KA)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int i = blockIdx.y;
    if (j >= p) return;
    uint32_t a = samples[i], b = samples[j];
    double k = 1.0;
    if (KIND != GL_PHOTOMETRIC) {
        double dr = (double)(a / width) - (double)(b / width), dc = (double)(a % width) - (double)(b % width);
        k *= exp(-(dr * dr + dc * dc) * inv_hl2);
    }
    if (KIND != GL_SPATIAL) {
        double d2 = 0.0;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
            double dv = (double)img[(size_t)a * C + ch] - (double)img[(size_t)b * C + ch];
            d2 += dv * dv;
        }
        k *= exp(-d2 * inv_hv2);
    }
    KA[(size_t)i * p + j] = k;
}

// K_B tile kernel.  KBS = sample slots per block (64 or 32).  Thread (tx = tid % (KBS/8), ty = tid / (KBS/8)): 8 consecutive
// slots of the current block x pixels ty, ty + PXL, ... of the tile.  A warp stores 32/(KBS/8) pixel rows of KBS*2 contiguous bytes.
template <int KIND, int C, int KBS>
// three CTAs per SM for grey images (80 registers), two for colour (the third would spill)
#define AFF_CTAS_PER_SM(C) ((C) == 1 ? 3 : 2)
__global__ void __launch_bounds__(AFF_THREADS, AFF_CTAS_PER_SM(C))
k_affinity_B(const uint8_t* __restrict__ img, const float* __restrict__ sf, int p_pad /* = p_int: internal sample slots */, int width,
             int64_t q0, int64_t q1, float a2, float b2,  // -log2(e)/h_loc^2, -log2(e)/h_val^2
             const int4* __restrict__ tab /* per tile: first entry of its block list, block count, storage offset */,
             const int* __restrict__ starts /* first internal sample of every stored block */,
             __half* __restrict__ KB /* [block][512][64] */, float* __restrict__ partial /* [gridDim.x][1 + C][p_int] */)
{
    extern __shared__ float aff_smem[];
    constexpr int NS = 1 + C;                  // sums per sample: D and T[ch]
    float* cta_sum = aff_smem;                 // [NS][p_pad]
    constexpr int TXN = KBS / 8;               // threads across the slots of a block
    constexpr int PXL = AFF_THREADS / TXN;     // pixel lanes
    constexpr int PPT = AFF_TP / PXL;          // pixels per thread per block
    float* ws = cta_sum + NS * p_pad;          // [2][NS][8 warps][KBS]
    float* px = ws + 2 * NS * 8 * KBS;         // [(2 + C)][AFF_TP] pixel features
    const int tid = threadIdx.x, tx = tid % TXN, ty = tid / TXN, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < NS * p_pad; i += AFF_THREADS) cta_sum[i] = 0.f;

    const int64_t n_band = q1 - q0;
    const int64_t tiles = (n_band + AFF_TP - 1) / AFF_TP;
    int flip = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t base = q0 + tile * AFF_TP;
        __syncthreads();  // previous tile's readers of px are done
        for (int i = tid; i < AFF_TP; i += AFF_THREADS) {
            int64_t q = base + i;
            bool in = q < q1;
            int64_t qq = in ? q : q1 - 1;
            px[i] = (float)(qq / width);
            px[AFF_TP + i] = (float)(qq % width);
#pragma unroll
            for (int ch = 0; ch < C; ++ch) px[(2 + ch) * AFF_TP + i] = (float)img[(size_t)qq * C + ch];
        }
        __syncthreads();
        const int4 tl = tab[tile];
        const bool single_row = base / width == (min(base + AFF_TP, q1) - 1) / width;
        for (int ci = 0; ci < tl.y; ++ci) {
            const int sb = starts[tl.x + ci];          // multiple of 8: the float4 loads below stay aligned
            const int s0 = sb + (tx << 3);
            __half* kb_blk = KB + ((size_t)(tl.z + ci) * AFF_TP) * KBS + (tx << 3);
            float sr[8], sc[8], sv[C][8];
            if (KIND != GL_PHOTOMETRIC) {
                *(float4*)&sr[0] = *(const float4*)&sf[s0];
                *(float4*)&sr[4] = *(const float4*)&sf[s0 + 4];
                *(float4*)&sc[0] = *(const float4*)&sf[p_pad + s0];
                *(float4*)&sc[4] = *(const float4*)&sf[p_pad + s0 + 4];
            }
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
                *(float4*)&sv[ch][0] = *(const float4*)&sf[(2 + ch) * p_pad + s0];
                *(float4*)&sv[ch][4] = *(const float4*)&sf[(2 + ch) * p_pad + s0 + 4];
            }
            float acc[8], tacc[C][8];
            // the row part of the exponent, a2 * dr^2: when the tile lies in one image row (the usual case on wide images) it is
            // the same for all of the thread's pixels and is taken out of the pair loop; otherwise it is refreshed per pixel
            float rterm[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                acc[k] = 0.f;
#pragma unroll
                for (int ch = 0; ch < C; ++ch) tacc[ch][k] = 0.f;
                rterm[k] = 0.f;
                if (KIND != GL_PHOTOMETRIC) {
                    const float dr = px[0] - sr[k];
                    rterm[k] = dr * dr * a2;
                }
            }
#pragma unroll 2
            for (int i = 0; i < PPT; ++i) {
                const int pi = ty + i * PXL;
                const int64_t q = base + pi;
                const float pc = px[AFF_TP + pi];
                float pv[C];
#pragma unroll
                for (int ch = 0; ch < C; ++ch) pv[ch] = px[(2 + ch) * AFF_TP + pi];
                if (KIND != GL_PHOTOMETRIC && !single_row) {
                    const float pr = px[pi];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float dr = pr - sr[k];
                        rterm[k] = dr * dr * a2;
                    }
                }
                float kv[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float x = rterm[k];
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
                    if (KIND != GL_PHOTOMETRIC) {
                        const float dc = pc - sc[k];
                        x = fmaf(dc * dc, a2, x);
                    }
                    kv[k] = fast_exp2(x);
                }
                if (q < q1) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        acc[k] += kv[k];
#pragma unroll
                        for (int ch = 0; ch < C; ++ch) tacc[ch][k] = fmaf(kv[k], pv[ch], tacc[ch][k]);
                    }
                    __half2 h0 = __floats2half2_rn(kv[0], kv[1]), h1 = __floats2half2_rn(kv[2], kv[3]);
                    __half2 h2 = __floats2half2_rn(kv[4], kv[5]), h3 = __floats2half2_rn(kv[6], kv[7]);
                    uint4 pk;
                    pk.x = *(uint32_t*)&h0; pk.y = *(uint32_t*)&h1; pk.z = *(uint32_t*)&h2; pk.w = *(uint32_t*)&h3;
                    *(uint4*)&kb_blk[(size_t)pi * KBS] = pk;
                }
            }
            // row sums: fixed-order reduction (deterministic): lanes sharing tx, then the 8 warps
#pragma unroll
            for (int k = 0; k < 8; ++k) {
#pragma unroll
                for (int o = TXN; o < 32; o <<= 1) {
                    acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) tacc[ch][k] += __shfl_xor_sync(0xffffffffu, tacc[ch][k], o);
                }
            }
            float* w = ws + flip * NS * 8 * KBS;
            if (lane < TXN) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    w[warp * KBS + (lane << 3) + k] = acc[k];
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) w[(1 + ch) * 8 * KBS + warp * KBS + (lane << 3) + k] = tacc[ch][k];
                }
            }
            __syncthreads();
            for (int i = tid; i < NS * KBS; i += AFF_THREADS) {
                const int which = i / KBS, sidx = i % KBS;
                float sum = 0.f;
#pragma unroll
                for (int wi = 0; wi < 8; ++wi) sum += w[which * 8 * KBS + wi * KBS + sidx];
                cta_sum[which * p_pad + sb + sidx] += sum;
            }
            flip ^= 1;
        }
    }
    __syncthreads();
    for (int i = tid; i < NS * p_pad; i += AFF_THREADS) partial[(size_t)blockIdx.x * NS * p_pad + i] = cta_sum[i];
}

// DT[which][j] = sum over CTAs (fixed order, fp64) for the sample j = perm[i] of internal slot i; which 0 = D, 1.. = T[ch].
// DT is in the caller's sample order (stride p_pad) and must be zeroed beforehand (its padding stays 0).
__global__ void k_reduce_partials(const float* __restrict__ partial, int nblocks, int p_int, int ns, const uint32_t* __restrict__ perm,
                                  int p_pad, double* __restrict__ DT)
{
    // one warp per output: lane l adds the CTA partials l, l + 32, ... in order, then a fixed shuffle tree (deterministic)
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= ns * p_int) return;
    const int which = i / p_int, slot = i - which * p_int;
    const uint32_t j = perm[slot];
    if (j == 0xffffffffu) return;
    double acc = 0.0;
    for (int b = lane; b < nblocks; b += 32) acc += (double)partial[(size_t)b * ns * p_int + i];
    acc = warp_sum(acc);
    if (lane == 0) DT[(size_t)which * p_pad + j] = acc;
}

// Layout of the stored K_B blocks (see the header).  Samples get an INTERNAL order: by column strip (S strips of equal
// width), then by raster index, so that inside a strip they are sorted by image row.  For a 512-pixel tile the samples that
// can matter are, per strip within R columns of the tile, a contiguous run of that strip's rows within R rows of the tile;
// each run is covered by 64-sample blocks starting at a multiple of 8 (unaligned to 64 on purpose: a run of <= 57 samples
// is one block).  Blocks of a tile are ascending and never overlap, so no (pixel, sample) pair is stored twice.
//   tile table : int4 {first entry in `starts`, block count, storage offset in blocks, 0} per tile
//   starts     : first internal sample slot of every stored block
//   perm       : internal slot -> index in the caller's ascending sample list (0xffffffff: empty slot), p_int = p_pad + 64
// S is chosen on the host as the candidate with the fewest stored blocks; S = 1 with the row cutoff alone when the image is
// narrow, and every block (aligned, S = 1) when there is no spatial term.  Built from the host mirror of the indices and
// cached on (geometry, samples, R).
struct KbLayout {
    std::vector<int4> tab;
    std::vector<int> starts;
    std::vector<uint32_t> perm;
    int64_t total = 0;
};

struct KbGeom {
    int W;            // image width
    int64_t q0, q1;   // raster range of the band
    int p_pad;        // sample count rounded up to 64
    int kbs;          // sample slots per block: 64 or 32
};

static void kb_layout_for(const KbGeom& geo, const std::vector<uint32_t>& samples, bool cut, int64_t R, int S, bool count_only, KbLayout* out,
                          int64_t tile_stride = 1 /* count_only: visit every tile_stride-th tile */)
{
    const int p = (int)samples.size(), W = geo.W, p_pad = geo.p_pad, p_int = p_pad + 64;
    const int64_t n_band = geo.q1 - geo.q0;
    const int64_t tiles = (n_band + AFF_TP - 1) / AFF_TP;
    // internal order: (strip, raster index); the input is ascending, so a stable bucket pass does it
    std::vector<int> strip_of(p), strip_begin(S + 1, 0);
    for (int i = 0; i < p; ++i) {
        strip_of[i] = (int)((int64_t)(samples[i] % (uint32_t)W) * S / W);
        strip_begin[strip_of[i] + 1]++;
    }
    for (int s = 0; s < S; ++s) strip_begin[s + 1] += strip_begin[s];
    std::vector<int> row_int(p);        // image row of the sample in internal slot i
    std::vector<uint32_t> perm(p_int, 0xffffffffu);
    {
        std::vector<int> fill(strip_begin.begin(), strip_begin.end() - 1);
        for (int i = 0; i < p; ++i) {
            const int slot = fill[strip_of[i]]++;
            perm[slot] = (uint32_t)i;
            row_int[slot] = (int)(samples[i] / (uint32_t)W);
        }
    }
    out->total = 0;
    if (!count_only) {
        out->tab.assign((size_t)tiles, make_int4(0, 0, 0, 0));
        out->starts.clear();
        out->starts.reserve((size_t)tiles * 2);
        out->perm = perm;
    }
    int64_t run_ra = -1, run_rb = -1;
    std::vector<int64_t> run_lo(S, 0), run_hi(S, 0);
    for (int64_t t = 0; t < tiles; t += tile_stride) {
        const int64_t first_entry = out->total;
        int64_t prev_end = 0;
        int cnt = 0;
        auto emit = [&](int64_t lo, int64_t hi) {   // cover internal slots [lo, hi) with blocks
            int64_t start = std::max<int64_t>(lo & ~(int64_t)7, prev_end);
            while (start < hi) {
                if (!count_only) out->starts.push_back((int)start);
                ++cnt;
                start += geo.kbs;
                prev_end = start;
            }
        };
        if (!cut) {
            emit(0, p_pad);
        } else {
            const int64_t qa = geo.q0 + t * AFF_TP, qb = std::min(geo.q1, qa + AFF_TP) - 1;
            const int64_t ra = qa / W, rb = qb / W;
            // column intervals the tile's pixels occupy: one or two row segments, or whole rows
            int64_t seg[2][2];
            int nseg = 1;
            if (ra == rb) { seg[0][0] = qa % W; seg[0][1] = qb % W; }
            else if (rb == ra + 1 && qb % W < qa % W) { seg[0][0] = qa % W; seg[0][1] = W - 1; seg[1][0] = 0; seg[1][1] = qb % W; nseg = 2; }
            else { seg[0][0] = 0; seg[0][1] = W - 1; }
            if (ra != run_ra || rb != run_rb) {   // tiles of the same image rows share the per-strip runs
                run_ra = ra;
                run_rb = rb;
                for (int s = 0; s < S; ++s) {
                    const int* rb_ = row_int.data() + strip_begin[s];
                    const int* re_ = row_int.data() + strip_begin[s + 1];
                    run_lo[s] = strip_begin[s] + (std::lower_bound(rb_, re_, (int)std::max<int64_t>(ra - R, -1)) - rb_);
                    run_hi[s] = strip_begin[s] + (std::upper_bound(rb_, re_, (int)std::min<int64_t>(rb + R, 0x7fffffff)) - rb_);
                }
            }
            for (int s = 0; s < S; ++s) {
                if (run_hi[s] <= run_lo[s]) continue;
                const int64_t c_lo = (int64_t)s * W / S, c_hi = (int64_t)(s + 1) * W / S - 1;   // columns of this strip (approx. bounds)
                bool need = false;
                for (int g = 0; g < nseg; ++g) need = need || (seg[g][0] - R <= c_hi + 1 && seg[g][1] + R >= c_lo - 1);
                if (need) emit(run_lo[s], run_hi[s]);
            }
            if (cnt == 0) emit(0, 1);   // keep one block so that the GEMM writes zeros
        }
        if (!count_only) out->tab[(size_t)t] = make_int4((int)first_entry, cnt, (int)first_entry, 0);
        out->total += cnt;
    }
}

// reach of the spatial term in pixels: |d| > h_loc sqrt(25 ln 2)  =>  exp(-d^2/h_loc^2) < 2^-25  =>  the fp16 value is 0;
// one more pixel for rounding slack
static int64_t kb_reach(double h_loc)
{
    const double rr = std::floor(h_loc * std::sqrt(25.0 * 0.6931471805599453)) + 1.0;
    return rr < 1e9 ? (int64_t)rr : (int64_t)1e9;
}

// picks the strip count (forced_S > 0: that one) and builds the layout; returns the strip count
static int kb_choose_and_build(const KbGeom& geo, const std::vector<uint32_t>& samples, bool cut, int64_t R, int forced_S, KbLayout* lay)
{
    int best_S = 1;
    if (cut) {
        if (forced_S > 0) {
            best_S = forced_S;
        } else {
            int64_t best = -1;
            for (int S : {1, 2, 3, 4, 5, 6, 8, 10, 12, 16}) {
                if (S > 1 && (geo.W / S < 64 || 2 * R * S > 4 * (int64_t)geo.W)) continue;   // strips much narrower than the reach cannot help
                kb_layout_for(geo, samples, true, R, S, true, lay, 7);   // a sample of the tiles is enough to rank the candidates
                if (best < 0 || lay->total < best) { best = lay->total; best_S = S; }
            }
        }
    }
    kb_layout_for(geo, samples, cut, R, best_S, false, lay);
    return best_S;
}

// Host-only view of the layout (no GPU, no context): what gl_affinity would store for these samples.  Used by the CPU tests.
extern "C" int gl_kb_layout_host(int width, int64_t q0, int64_t q1, const uint32_t* samples, unsigned p, double h_loc, int cutoff,
                                 int strips, int block_slots, int* strips_out, int64_t* n_tiles, int64_t* n_blocks, int32_t* tile_first,
                                 int32_t* tile_count, int32_t* starts, uint32_t* perm)
{
    GL_REQUIRE(samples && p >= 1 && width > 0 && q1 > q0 && h_loc > 0, "gl_kb_layout_host: bad arguments");
    for (unsigned i = 1; i < p; ++i) GL_REQUIRE(samples[i] > samples[i - 1], "gl_kb_layout_host: samples must be strictly ascending");
    const KbGeom geo = {width, q0, q1, (int)round_up(p, 64), block_slots == 32 ? 32 : 64};
    std::vector<uint32_t> sv(samples, samples + p);
    KbLayout lay;
    const int S = kb_choose_and_build(geo, sv, cutoff != 0, kb_reach(h_loc), strips, &lay);
    if (strips_out) *strips_out = S;
    if (n_tiles) *n_tiles = (int64_t)lay.tab.size();
    if (n_blocks) *n_blocks = lay.total;
    if (tile_first && tile_count)
        for (size_t t = 0; t < lay.tab.size(); ++t) { tile_first[t] = lay.tab[t].x; tile_count[t] = lay.tab[t].y; }
    if (starts) memcpy(starts, lay.starts.data(), sizeof(int) * lay.starts.size());
    if (perm) memcpy(perm, lay.perm.data(), sizeof(uint32_t) * lay.perm.size());
    return GL_OK;
}

static int build_tile_table(gl_ctx* ctx, int kind, double h_loc)
{
    const int p = (int)ctx->p, W = ctx->width;
    GL_CHECK(gl_host_samples(ctx));
    const bool cut = ctx->kb_cutoff && kind != GL_PHOTOMETRIC && kind != GL_NLM;   // no spatial term, no cutoff
    const int64_t R = kb_reach(h_loc);
    const int kbs = ctx->kb_block == 32 ? 32 : 64;
    const int64_t key[6] = {W, ctx->q0, ctx->q1, p, R, (cut ? 1 : 0) + 2 * ctx->kb_strips + 1000 * kbs};
    if (ctx->tile_tab && !memcmp(key, ctx->tab_key, sizeof(key)) && ctx->tab_samples.size() == (size_t)p &&
        !memcmp(ctx->tab_samples.data(), ctx->h_samples.data(), sizeof(uint32_t) * p))
        return GL_OK;  // same geometry, samples and cutoff as last time: the cached layout stands
    KbLayout lay;
    const KbGeom geo = {W, ctx->q0, ctx->q1, ctx->p_pad, kbs};
    // the strip count depends on the geometry and the sample density, not on where exactly the samples fell: it is chosen
    // once per (geometry, p, reach) and reused when only the sample positions change (a new random draw costs one build)
    int forced = ctx->kb_strips;
    if (forced == 0 && ctx->tile_tab && !memcmp(key, ctx->tab_key, sizeof(key))) forced = ctx->tile_strips;
    const int best_S = kb_choose_and_build(geo, ctx->h_samples, cut, R, forced, &lay);
    GL_REQUIRE(lay.total < 0x7fffffff / 512, "affinity: K_B has too many blocks for 32-bit tile coordinates");
    // upload: table, block starts, permutation (one staging pass through the pinned block each)
    gl_buf** dst[3] = {&ctx->tile_tab, &ctx->tile_starts, &ctx->tile_perm};
    const void* src[3] = {lay.tab.data(), lay.starts.data(), lay.perm.data()};
    const size_t bytes[3] = {sizeof(int4) * lay.tab.size(), sizeof(int) * lay.starts.size(), sizeof(uint32_t) * lay.perm.size()};
    GL_CHECK(gl_ensure_pinned(ctx, bytes[0] + bytes[1] + bytes[2] + 64));
    GL_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // the pinned block may still feed an earlier copy
    size_t off = 0;
    for (int k = 0; k < 3; ++k) {
        if (*dst[k]) gl_buf_release(*dst[k]);
        *dst[k] = nullptr;
        GL_CHECK(gl_alloc(ctx, bytes[k], dst[k]));
        memcpy((char*)ctx->pinned + off, src[k], bytes[k]);
        GL_CUDA_CHECK(cudaMemcpyAsync((*dst[k])->ptr, (char*)ctx->pinned + off, bytes[k], cudaMemcpyHostToDevice, ctx->stream));
        off += (bytes[k] + 15) & ~(size_t)15;
    }
    ctx->tile_total_blocks = lay.total;
    ctx->tile_strips = best_S;
    ctx->tile_kbs = kbs;
    memcpy(ctx->tab_key, key, sizeof(key));
    ctx->tab_samples = ctx->h_samples;
    return GL_OK;
}

template <int KIND, int C, int KBS>
static int launch_affinity_b(gl_ctx* ctx, double h_loc, double h_val, const float* sf, const int4* tab, const int* starts, __half* KB,
                             float* partial, int grid)
{
    const int p_pad = ctx->p_pad + 64;   // the K_B kernel works on the internal sample slots (p_int)
    const size_t smem = sizeof(float) * ((size_t)(1 + C) * p_pad + 2 * (1 + C) * 8 * KBS + (size_t)(2 + C) * AFF_TP);
    GL_CUDA_CHECK(cudaFuncSetAttribute(k_affinity_B<KIND, C, KBS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const float log2e = 1.4426950408889634f;
    StageTimer kt(ctx, GL_T_K_AFFINITY_B);
    k_affinity_B<KIND, C, KBS><<<grid, AFF_THREADS, smem, ctx->stream>>>(
        (const uint8_t*)ctx->img->ptr, sf, p_pad, ctx->width, ctx->q0, ctx->q1, (float)(-log2e / (h_loc * h_loc)),
        (float)(-log2e / (h_val * h_val)), tab, starts, KB, partial);
    GL_LAUNCH_CHECK(ctx);
    return GL_OK;
}

template <int KIND, int C>
static int launch_affinity_a(gl_ctx* ctx, double h_loc, double h_val, double* KA)
{
    const int p = (int)ctx->p;
    dim3 ga((unsigned)ceil_div(p, 128), (unsigned)p);
    k_affinity_A<KIND, C><<<ga, 128, 0, ctx->stream>>>((const uint8_t*)ctx->img->ptr, (const uint32_t*)ctx->samples->ptr, p,
                                                      ctx->width, 1.0 / (h_loc * h_loc), 1.0 / (h_val * h_val), KA);
    GL_LAUNCH_CHECK(ctx);
    return GL_OK;
}

template <int KIND, int C>
static int launch_affinity_kb(gl_ctx* ctx, double h_loc, double h_val, const float* sf, const int4* tab, const int* starts,
                              __half* KB, float* partial, int grid)
{
    if (ctx->tile_kbs == 32) return launch_affinity_b<KIND, C, 32>(ctx, h_loc, h_val, sf, tab, starts, KB, partial, grid);
    return launch_affinity_b<KIND, C, 64>(ctx, h_loc, h_val, sf, tab, starts, KB, partial, grid);
}

// kay[ch][i] = (K_A y_S)[i][ch] in fp64: the sample pixels' share of T = [K_A K_B] y, which the filter's projection takes out
// again (nystroem_gemm.cu: the sample rows of Phi are Phi_A, not the extrapolation).  One block per sample row.
__global__ void k_ka_times_y(const double* __restrict__ KA, const uint8_t* __restrict__ img, const uint32_t* __restrict__ samples, int p,
                             int p_pad, int C, double* __restrict__ kay)
{
    const int i = blockIdx.x;
    __shared__ double red[3][32];
    double acc[3] = {0.0, 0.0, 0.0};
    for (int j = threadIdx.x; j < p; j += blockDim.x) {
        const double k = KA[(size_t)i * p + j];
        const size_t b = samples[j];
        for (int ch = 0; ch < C; ++ch) acc[ch] += k * (double)img[b * C + ch];
    }
    for (int ch = 0; ch < C; ++ch) {
        double v = warp_sum(acc[ch]);
        if ((threadIdx.x & 31) == 0) red[ch][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < C) {
        double v = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[threadIdx.x][w];
        kay[(size_t)threadIdx.x * p_pad + i] = v;
    }
}

// affinity_nlm.cu
size_t gl_nlm_smem_bytes(int p_int);
int gl_nlm_affinity_launch(gl_ctx* ctx, double h, int p_int, double* KA, const int4* tab, const int* starts, const uint32_t* perm, __half* KB,
                           float* partial, int grid);

// K_A (fp64, p x p) for the non-patch kinds
static int affinity_a(gl_ctx* ctx, int kind, double h_loc, double h_val, double* KA)
{
    const int C = ctx->channels;
#define AFF_A(K, CC) \
    if (kind == K && C == CC) return launch_affinity_a<K, CC>(ctx, h_loc, h_val, KA);
    AFF_A(GL_BILATERAL, 1) AFF_A(GL_BILATERAL, 3) AFF_A(GL_PHOTOMETRIC, 1) AFF_A(GL_PHOTOMETRIC, 3) AFF_A(GL_SPATIAL, 1) AFF_A(GL_SPATIAL, 3)
#undef AFF_A
    gl_set_error("affinity: kind %d with %d channels is not supported", kind, C);
    return GL_ERR_UNSUPPORTED;
}

// The blocked storage of K_B (layout + tiles) into the handle; `DT` (may be null) receives the band-partial sums D / T in the caller's
// sample order ([1 + C][p_pad] doubles, zeroed here).  NLM also writes K_A (it shares the patch staging), the other kinds do not.
static int affinity_blocked_fill(gl_ctx* ctx, int kind, double h_loc, double h_val, gl_mat* KB, double* KA_for_nlm, double* DT)
{
    const int p_pad = ctx->p_pad, C = ctx->channels;
    const int p_int = p_pad + 64;   // internal sample slots (a block may start at any multiple of 8 below p)
    gl_buf *sf = nullptr, *partial = nullptr;
    const int grid = ctx->sm_count * (kind == GL_NLM ? 1 : AFF_CTAS_PER_SM(C));
    int rc = GL_OK;
    do {
        if ((rc = build_tile_table(ctx, kind, h_loc)) != GL_OK) break;
        KB->tiles = ctx->tile_tab;      // shared with the context's cache (a new layout is a new set of buffers)
        KB->tiles->refs++;
        KB->starts = ctx->tile_starts;
        KB->starts->refs++;
        KB->perm = ctx->tile_perm;
        KB->perm->refs++;
        KB->total_blocks = ctx->tile_total_blocks;
        KB->kbs = ctx->tile_kbs;
        KB->ld = KB->kbs;
        if ((rc = gl_alloc(ctx, sizeof(__half) * (size_t)KB->total_blocks * AFF_TP * KB->kbs, &KB->buf)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(float) * (size_t)(2 + C) * p_int, &sf)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(float) * (size_t)grid * (1 + C) * p_int, &partial)) != GL_OK) break;
        if (kind != GL_NLM) {
            k_sample_features<<<(unsigned)ceil_div(p_int, 256), 256, 0, ctx->stream>>>(
                (const uint8_t*)ctx->img->ptr, (const uint32_t*)ctx->samples->ptr, (const uint32_t*)KB->perm->ptr, p_int, ctx->width, C,
                (float*)sf->ptr);
            ctx->launches++;
        }
#define AFF_CASE(K, CC)                                                                                              \
    if (kind == K && C == CC)                                                                                        \
        rc = launch_affinity_kb<K, CC>(ctx, h_loc, h_val, (const float*)sf->ptr, (const int4*)KB->tiles->ptr,        \
                                       (const int*)KB->starts->ptr, (__half*)KB->buf->ptr, (float*)partial->ptr, grid);
        AFF_CASE(GL_BILATERAL, 1) else AFF_CASE(GL_BILATERAL, 3) else AFF_CASE(GL_PHOTOMETRIC, 1)
        else AFF_CASE(GL_PHOTOMETRIC, 3) else AFF_CASE(GL_SPATIAL, 1) else AFF_CASE(GL_SPATIAL, 3)
        else if (kind == GL_NLM) {
            gl_buf* ka_tmp = nullptr;
            if (!KA_for_nlm) {      // (a lazily rebuilt NLM K_B: K_A is recomputed into scratch)
                if ((rc = gl_alloc(ctx, sizeof(double) * (size_t)ctx->p * ctx->p, &ka_tmp)) != GL_OK) break;
                KA_for_nlm = (double*)ka_tmp->ptr;
            }
            rc = gl_nlm_affinity_launch(ctx, h_val, p_int, KA_for_nlm, (const int4*)KB->tiles->ptr, (const int*)KB->starts->ptr,
                                        (const uint32_t*)KB->perm->ptr, (__half*)KB->buf->ptr, (float*)partial->ptr, grid);
            if (ka_tmp) gl_buf_release(ka_tmp);
        }
        else { gl_set_error("affinity: kind %d with %d channels is not supported", kind, C); rc = GL_ERR_UNSUPPORTED; }
#undef AFF_CASE
        if (rc != GL_OK) break;
        if (DT) {
            GL_CUDA_BREAK(rc, cudaMemsetAsync(DT, 0, sizeof(double) * (size_t)(1 + C) * p_pad, ctx->stream));
            k_reduce_partials<<<(unsigned)ceil_div((1 + C) * p_int, 8), 256, 0, ctx->stream>>>(
                (const float*)partial->ptr, grid, p_int, 1 + C, (const uint32_t*)KB->perm->ptr, p_pad, DT);
            ctx->launches++;
            if (cudaGetLastError() != cudaSuccess) { gl_set_error("affinity: kernel launch failed"); rc = GL_ERR_CUDA; }
        }
    } while (0);
    if (sf) gl_buf_release(sf);
    if (partial) gl_buf_release(partial);
    return rc;
}

int gl_kb_require_blocked(gl_ctx* ctx, gl_mat* KB)
{
    if (KB->buf) return GL_OK;
    GL_REQUIRE(KB->pt_buf, "K_B handle without storage");
    GL_REQUIRE(KB->image_epoch == ctx->image_epoch && KB->sample_epoch == ctx->sample_epoch && KB->q0 == ctx->q0 && KB->p == (int)ctx->p,
               "K_B: the blocked storage is computed on demand from the image and the samples, which have changed since gl_affinity");
    return affinity_blocked_fill(ctx, KB->aff_kind, KB->aff_h_loc, KB->aff_h_val, KB, nullptr, nullptr);
}

int gl_impl_affinity(gl_ctx* ctx, int kind, double h_loc, double h_val, gl_mat** K_A_out, gl_mat** K_B_out)
{
    const int p = (int)ctx->p, p_pad = ctx->p_pad, C = ctx->channels;
    const int p_int = p_pad + 64;
    const int64_t n_band = ctx->q1 - ctx->q0;
    const bool patch = gl_patch_applicable(ctx, kind) && !ctx->want_blocked;
    if (!patch) {
        const size_t smem = kind == GL_NLM ? gl_nlm_smem_bytes(p_int)
                                           : sizeof(float) * ((size_t)(1 + C) * p_int + 2 * (1 + C) * 8 * 64 + (size_t)(2 + C) * AFF_TP);
        if (smem > 227 * 1024) {
            gl_set_error("affinity: p = %d samples need %zu bytes of shared memory per CTA (limit 227 KB)", p, smem);
            return GL_ERR_UNSUPPORTED;
        }
    }

    gl_mat* KA = gl_mat_new(ctx, GL_MAT_KA);
    gl_mat* KB = gl_mat_new(ctx, GL_MAT_KB);
    int rc = GL_OK;
    do {
        KA->rows = KA->cols = KA->local_rows = p;
        KA->ld = p;
        KA->elem_bytes = 8;
        if ((rc = gl_alloc(ctx, sizeof(double) * (size_t)p * p, &KA->buf)) != GL_OK) break;
        KB->rows = p;                   // logical K_B: p x (n - p); stored transposed for the whole band
        KB->cols = ctx->n - p;
        KB->local_rows = n_band;
        KB->ld = patch ? 32 : (ctx->kb_block == 32 ? 32 : 64);   // sample slots per stored row of a block / tile
        KB->elem_bytes = 2;
        KB->p = p;
        KB->p_pad = p_pad;
        KB->q0 = ctx->q0;
        KB->aff_kind = kind;
        KB->aff_h_loc = h_loc;
        KB->aff_h_val = h_val;
        if ((rc = gl_alloc(ctx, sizeof(double) * (size_t)(1 + 2 * C) * p_pad, &KB->aux)) != GL_OK) break;  // [D | T[ch] | (K_A y_S)[ch]]
        // (patch path: the sample lists are built first -- they need no pixels, so gl_run's image upload runs under them --, K_A after)
        if (!patch) GL_BREAK(rc, gl_image_ready(ctx));
        if (!patch && kind != GL_NLM && (rc = affinity_a(ctx, kind, h_loc, h_val, (double*)KA->buf->ptr)) != GL_OK) break;
        if (patch) {
            rc = gl_patch_affinity(ctx, kind, h_loc, h_val, KB);
            if (rc == GL_OK) rc = gl_image_ready(ctx);
            if (rc == GL_OK) rc = affinity_a(ctx, kind, h_loc, h_val, (double*)KA->buf->ptr);
        } else {
            GL_CUDA_BREAK(rc, cudaMemsetAsync(KB->aux->ptr, 0, sizeof(double) * (size_t)(1 + 2 * C) * p_pad, ctx->stream));
            rc = affinity_blocked_fill(ctx, kind, h_loc, h_val, KB, (double*)KA->buf->ptr, (double*)KB->aux->ptr);
        }
        if (rc != GL_OK) break;
        k_ka_times_y<<<p, 128, 0, ctx->stream>>>((const double*)KA->buf->ptr, (const uint8_t*)ctx->img->ptr, (const uint32_t*)ctx->samples->ptr,
                                                 p, p_pad, C, (double*)KB->aux->ptr + (size_t)(1 + C) * p_pad);
        GL_LAUNCH_CHECK(ctx);
        // ONE allreduce of the band-partial sums: D (p doubles) and T (C x p doubles)
        if ((rc = gl_allreduce_f64(ctx, (double*)KB->aux->ptr, (size_t)(1 + C) * p_pad)) != GL_OK) break;
        KB->channels = C;
        KB->image_epoch = ctx->image_epoch;
        KB->sample_epoch = ctx->sample_epoch;
    } while (0);
    if (rc != GL_OK) {
        gl_mat_destroy(KA);
        gl_mat_destroy(KB);
        return rc;
    }
    *K_A_out = KA;
    *K_B_out = KB;
    return GL_OK;
}
