// a-6 (+ a-7): Nystroem extrapolation Phi = [Phi_A ; L_B^T . Phi_A . Lambda^-1] as a tcgen05 GEMM.
// Replaces Nystroem (hpc/nystroem.c:5-69: MatMatMult + MatTransposeMatMult + two row-block copies),
// InverseDiagMat's use there (hpc/utils.c:559-586) and Permutation (hpc/utils.c:134-173).
//
//   Phi[q, j] = sum_s K_B[q, s] * W[s, j],   W = -alpha * U * diag(1/mu)        (p x m)
//
//   * A operand  = K_B band in its blocked storage ([block][512 pixels][64 sample slots] fp16, K-major; affinity.cu):
//                  M = 128 pixels per tile, the K loop walks the tile's stored blocks only
//   * B operand  = W^T, [m_pad x (p_pad + 64)] fp16, K-major, columns in K_B's internal sample order, scaled by one power
//                  of two so that it sits at the top of the fp16 range (W itself is ~1e-4 and would be subnormal);
//                  unscaled in the epilogue; a block's 64 rows of W start at the block's first sample slot
//   * D          = fp32 accumulators in TMEM (2 x 256 columns, double buffered), written as fp16 Phi (|Phi| <~ 1:
//                  fp16's 11-bit mantissa beats bf16's 8 and the filter sums cancel heavily)
//   * TMA (SWIZZLE_128B) feeds a shared-memory ring; one elected thread issues tcgen05.mma; the epilogue warps drain
//     TMEM with tcgen05.ld, convert, stage swizzled rows in shared memory and TMA-store them while the next tile's MMAs
//     run.  Long K loops: 4 stages, 4 epilogue warps; short K loops (blocked K_B): 3 stages, 8 epilogue warps, L2 prefetch.
//   * optionally (gl_nystroem_filter) the epilogue also accumulates each row's product with the filter weights from the fp32
//     accumulators, so that the filter needs no second pass over Phi -- and, with no Phi requested, nothing is stored.
//   * every pixel's row is written at its raster position, so there is no permutation pass; the p sample
//     rows are then overwritten with Phi_A (nystroem.c:25-34).
// A plain CUDA-core kernel (option gemm=simple) computes the same thing for cross-checking in tests.
#include "tc_common.cuh"

// ---------------------------------------------------------------------------------------------
// W^T build
// ---------------------------------------------------------------------------------------------
__global__ void k_w_colmax(const float* __restrict__ U, int ld, int p, int m, const double* __restrict__ mu_inv,
                           const double* __restrict__ neg_alpha, float* __restrict__ colmax)
{
    const int j = blockIdx.x;
    float mx = 0.f;
    for (int s = threadIdx.x; s < p; s += blockDim.x) mx = fmaxf(mx, fabsf(U[(size_t)j * ld + s]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    __shared__ float sh[32];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, sh[w]);
        colmax[j] = mx * (float)fabs(neg_alpha[0] * mu_inv[j]);
    }
}

// scales[0] = 2^e (applied to W), scales[1] = 2^-e (applied in the GEMM epilogue); max |W| * 2^e in [2^13, 2^14)
__global__ void k_w_scale(const float* __restrict__ colmax, int m, float* __restrict__ scales)
{
    __shared__ float sh[32];
    float mx = 0.f;
    for (int j = threadIdx.x; j < m; j += blockDim.x) mx = fmaxf(mx, colmax[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, sh[w]);
        int e = 0;
        if (mx > 0.f && isfinite(mx)) {
            int ex;
            frexpf(mx, &ex);  // mx = f * 2^ex, f in [0.5, 1)
            e = 14 - ex;
        }
        e = max(-100, min(100, e));
        scales[0] = ldexpf(1.f, e);
        scales[1] = ldexpf(1.f, -e);
    }
}

// W^T in K_B's INTERNAL sample order: column `slot` of row j holds W[perm[slot]][j] (0 for empty slots); k_dim = p_pad + 64
__global__ void k_w_write(const float* __restrict__ U, int ld, int m, int k_dim, const uint32_t* __restrict__ perm,
                          const double* __restrict__ mu_inv, const double* __restrict__ neg_alpha,
                          const float* __restrict__ scales, __half* __restrict__ Wt)
{
    const int j = blockIdx.x;  // row of W^T
    const float f = j < m ? (float)(neg_alpha[0] * mu_inv[j]) * scales[0] : 0.f;
    for (int s = threadIdx.x; s < k_dim; s += blockDim.x) {
        const uint32_t i = perm[s];
        float v = (j < m && i != 0xffffffffu) ? U[(size_t)j * ld + i] * f : 0.f;
        Wt[(size_t)j * k_dim + s] = __float2half_rn(v);
    }
}

// Phi rows of the sample pixels <- Phi_A (nystroem.c:25-34); band-local
__global__ void k_phi_sample_rows(const float* __restrict__ U, int ld, int p, int m, const uint32_t* __restrict__ samples,
                                  int64_t q0, int64_t q1, int m_pad, __half* __restrict__ phi)
{
    const int i = blockIdx.x;
    const int64_t q = samples[i];
    if (q < q0 || q >= q1) return;
    for (int j = threadIdx.x; j < m; j += blockDim.x)
        phi[(size_t)(q - q0) * m_pad + j] = __float2half_rn(U[(size_t)j * ld + i]);
}

// c = Phi^T y from the affinity stage's image-weighted sums, all in fp64 (one block per eigenvector j):
//   c[j][ch] = sum_i U[i][j] y[s_i][ch]  +  sum_i W[i][j] (T[ch][i] - (K_A y_S)[i][ch]),   W = -alpha U / mu
// T = [K_A K_B] y over ALL pixels (affinity.cu); removing K_A y_S leaves K_B y_B because the sample pixels' rows of
// Phi are Phi_A, not the extrapolation (nystroem.c:25-34).  kay[ch][i] = (K_A y_S)[i][ch] was stored behind D and T by the
// affinity stage (affinity.cu: k_ka_times_y).
__global__ void k_proj_from_sums(const float* __restrict__ U, int ld, int p, int p_pad, int m, const double* __restrict__ mu_inv,
                                 const double* __restrict__ neg_alpha, const double* __restrict__ DT /* [1+C][p_pad] */,
                                 const double* __restrict__ kay /* [C][p_pad] */, const uint8_t* __restrict__ img,
                                 const uint32_t* __restrict__ samples, int C, double* __restrict__ proj /* [m_pad][C] */)
{
    const int j = blockIdx.x;
    __shared__ double red[3][32];
    double acc[3] = {0.0, 0.0, 0.0};
    if (j < m) {
        const double f = neg_alpha[0] * mu_inv[j];
        for (int i = threadIdx.x; i < p; i += blockDim.x) {
            const double u = (double)U[(size_t)j * ld + i];
            const uint32_t q = samples[i];
            for (int ch = 0; ch < C; ++ch)
                acc[ch] += u * ((double)img[(size_t)q * C + ch] + f * (DT[(size_t)(1 + ch) * p_pad + i] - kay[(size_t)ch * p_pad + i]));
        }
    }
    for (int ch = 0; ch < C; ++ch) {
        double v = warp_sum(acc[ch]);
        if ((threadIdx.x & 31) == 0) red[ch][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < C) {
        double v = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[threadIdx.x][w];
        proj[(size_t)j * C + threadIdx.x] = j < m ? v : 0.0;
    }
}

// ---------------------------------------------------------------------------------------------
// checker GEMM on CUDA cores (tests / debugging only; selected with option gemm=simple)
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__global__ void __launch_bounds__(256) k_gemm_simple(const T* __restrict__ A, const T* __restrict__ Bt, int64_t M, int N, int K,
                                                     const float* __restrict__ scales, const __half* __restrict__ addend,
                                                     __half* __restrict__ D)
{
    __shared__ float As[32][33], Bs[32][33];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 2 x 2 outputs each
    const int64_t m0 = (int64_t)blockIdx.y * 32;
    const int n0 = blockIdx.x * 32;
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    for (int k0 = 0; k0 < K; k0 += 32) {
        for (int i = threadIdx.x; i < 32 * 32; i += 256) {
            const int r = i >> 5, c = i & 31;
            As[r][c] = (m0 + r < M) ? to_f32<T>(A[(size_t)(m0 + r) * K + k0 + c]) : 0.f;
            Bs[r][c] = (n0 + r < N) ? to_f32<T>(Bt[(size_t)(n0 + r) * K + k0 + c]) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            const float a0 = As[ty][k], a1 = As[ty + 16][k], b0 = Bs[tx][k], b1 = Bs[tx + 16][k];
            acc[0][0] = fmaf(a0, b0, acc[0][0]);
            acc[0][1] = fmaf(a0, b1, acc[0][1]);
            acc[1][0] = fmaf(a1, b0, acc[1][0]);
            acc[1][1] = fmaf(a1, b1, acc[1][1]);
        }
        __syncthreads();
    }
    const float sc = scales[1];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int64_t r = m0 + ty + 16 * a;
            const int c = n0 + tx + 16 * b;
            if (r < M && c < N) {
                float v = acc[a][b] * sc;
                if (addend) v += __half2float(addend[(size_t)r * N + c]);
                D[(size_t)r * N + c] = __float2half_rn(v);
            }
        }
}

// element (pixel row r, internal sample slot k) of K_B in its blocked storage; slots no stored block of the tile covers are zero
__device__ __forceinline__ float kb_blocked_at(const __half* __restrict__ A, const int4* __restrict__ tab, const int* __restrict__ starts,
                                               int kbs, int64_t r, int k)
{
    const int4 tl = tab[r >> 9];
    for (int b = 0; b < tl.y; ++b) {
        const int sb = starts[tl.x + b];
        if (k >= sb && k < sb + kbs) return __half2float(A[(((size_t)tl.z + b) * 512 + (size_t)(r & 511)) * kbs + (k - sb)]);
    }
    return 0.f;
}

__global__ void __launch_bounds__(256) k_gemm_simple_blocked(const __half* __restrict__ A, const int4* __restrict__ tab,
                                                             const int* __restrict__ starts, int kbs, const __half* __restrict__ Bt,
                                                             int64_t M, int N, int K,
                                                             const float* __restrict__ scales, __half* __restrict__ D)
{
    __shared__ float As[32][33], Bs[32][33];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t m0 = (int64_t)blockIdx.y * 32;
    const int n0 = blockIdx.x * 32;
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    for (int k0 = 0; k0 < K; k0 += 32) {
        for (int i = threadIdx.x; i < 32 * 32; i += 256) {
            const int r = i >> 5, c = i & 31;
            As[r][c] = (m0 + r < M) ? kb_blocked_at(A, tab, starts, kbs, m0 + r, k0 + c) : 0.f;
            Bs[r][c] = (n0 + r < N) ? __half2float(Bt[(size_t)(n0 + r) * K + k0 + c]) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            const float a0 = As[ty][k], a1 = As[ty + 16][k], b0 = Bs[tx][k], b1 = Bs[tx + 16][k];
            acc[0][0] = fmaf(a0, b0, acc[0][0]);
            acc[0][1] = fmaf(a0, b1, acc[0][1]);
            acc[1][0] = fmaf(a1, b0, acc[1][0]);
            acc[1][1] = fmaf(a1, b1, acc[1][1]);
        }
        __syncthreads();
    }
    const float sc = scales[1];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int64_t r = m0 + ty + 16 * a;
            const int c = n0 + tx + 16 * b;
            if (r < M && c < N) D[(size_t)r * N + c] = __float2half_rn(acc[a][b] * sc);
        }
}

// ---------------------------------------------------------------------------------------------
// tcgen05 GEMM
// ---------------------------------------------------------------------------------------------
namespace tc {

constexpr int MAX_STAGES = 4;
constexpr int MAX_THREADS = 384;   // warps 0-3: producer, MMA issuer, TMEM allocator, idle; then 4 or 8 epilogue warps
constexpr int C_SLAB_BYTES = 32 * 128;                     // 32 rows x 64 bf16
// two shapes of the shared-memory budget (227 KB): a deep operand ring for long K loops (dense A, the epilogue has
// slack: one store slab per epilogue warp), or a shorter ring with double-buffered store slabs for the short K loops of
// the blocked K_B, where the epilogue is the critical path
template <int STAGES, int CBUFS, int EPI_WARPS, int BK = 64>
constexpr int smem_bytes()
{
    return STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) * BK / 64 + EPI_WARPS * CBUFS * C_SLAB_BYTES + 256 /*barriers*/ + 1024 /*alignment*/;
}

// One persistent CTA per SM.  warp 0: TMA producer, warp 1: MMA issuer, warp 2: TMEM allocator,
// warps 4..: epilogue, EPI_WARPS / 4 per TMEM lane quarter (warp w drains lanes 32*(w%4) .. +31 and its share of the
// accumulator's columns).
// FC > 0 fuses the filter application into the epilogue (FC = image channels): besides storing the fp16 tile, every
// epilogue thread accumulates dot(row of D, w[:, ch]) over its columns from the fp32 accumulators and writes one partial
// per (N tile, column share) to zpart[part][row][ch]; filter.cu sums the partials in fixed order (deterministic).
// BK = elements of K per shared-memory stage: 64 (SWIZZLE_128B operands) or 32 (SWIZZLE_64B; K_B stored in 32-slot blocks).
// ST = false (only with FC > 0): D is consumed by the fused filter and never written -- no packing, no store slabs, no TMA stores.
template <int STAGES, int CBUFS, int EPI_WARPS, int FC, int BK, bool ST>
__global__ void __launch_bounds__(128 + 32 * EPI_WARPS, 1)
k_gemm_tcgen05(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_d, int m_tiles, int n_tiles, int k_blocks, int n_total, int block_n,
               int ab_bf16, const float* __restrict__ scales, const __half* __restrict__ addend, int64_t m_rows,
               const int4* __restrict__ a_tab /* K_B tile table, or null for a dense A */,
               const int* __restrict__ a_starts /* K_B: first W row (sample slot) of every stored block */, int prefetch_tiles,
               const float* __restrict__ fuse_w /* [n_total][FC] */, float* __restrict__ zpart /* [parts][m_rows][FC] */,
               int* __restrict__ err)
{
    static_assert(ST || FC > 0, "a GEMM that neither stores nor filters has no output");
    constexpr bool store_d = ST;
    extern __shared__ uint8_t gemm_smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)gemm_smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    constexpr int A_STAGE = A_STAGE_BYTES * BK / 64, B_STAGE = B_STAGE_BYTES * BK / 64;
    uint8_t* smem_b = smem_a + STAGES * A_STAGE;
    uint8_t* smem_c = smem_b + STAGES * B_STAGE;
    uint64_t* bars = (uint64_t*)(smem_c + EPI_WARPS * CBUFS * C_SLAB_BYTES);
    // bars: [0,S) full (A of the step, and B when the step loads one), [S,2S) A slot empty, [2S,3S) B slot empty,
    // [3S,3S+2) tmem_full, [3S+2,3S+4) tmem_empty, then the TMEM base address
    static_assert(3 * STAGES + 5 <= 32, "barrier block is 256 bytes");
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + STAGES), bar_bempty = smem_u32(bars + 2 * STAGES);
    const uint32_t bar_tfull = smem_u32(bars + 3 * STAGES), bar_tempty = smem_u32(bars + 3 * STAGES + 2);
    uint32_t* tmem_slot = (uint32_t*)(bars + 3 * STAGES + 4);
    float* w_s = (float*)(bars + 32);   // after the 256-byte barrier block (FC > 0 only): per epilogue warp, the filter
                                        // weights of the columns it drains in the current tile

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Work is handed out in groups: the (up to) four 128-row M tiles of one 512-pixel tile of K_B times one N tile.  They share
    // their list of K blocks, so the B operand (the W rows of those blocks) is loaded ONCE per group and stays in its slots
    // while the four A tiles stream past it, which halves the L2 -> shared-memory traffic of a tile -- the bound of this kernel
    // once Phi is not stored.  A group with more blocks than B slots, and the dense A (groups of one M tile), stream B too.
    // With Phi stored the kernel sits on the HBM write wall instead, and handing out single M tiles (the four N tiles of a row
    // written by neighbouring CTAs at the same time) writes faster: groups of one.
    const int GROUP = (a_tab && !store_d) ? 4 : 1;
    const int TSH = a_tab ? (GROUP == 4 ? 0 : 2) : 0;     // group index -> index of its 512-pixel tile in a_tab
    const int total_groups = ((m_tiles + GROUP - 1) / GROUP) * n_tiles;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_d) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
            mbar_init(bar_bempty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, EPI_WARPS);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t a_tx = (uint32_t)A_STAGE, b_tx = (uint32_t)(block_n * BK * 2);

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;           // A ring
            uint32_t phase = 0;
            int bpos = 0;            // B ring: slot of the next B load
            uint32_t bloads = 0;     // bit s: parity of the number of loads into B slot s so far
            for (int grp = blockIdx.x; grp < total_groups; grp += gridDim.x) {
                const int T = grp / n_tiles, nt = grp % n_tiles;
                // dense A: K block kb of rows T*128..; blocked A (K_B): stored block (offset + kb) of the 512-pixel tile T, rows
                // i * 128.. inside it, multiplying the 64 rows of W that start at the block's first sample slot
                int kb_count = k_blocks, cnt = 1, a_row0 = T * BLOCK_M, a_row_step = 0, a_col_step = BK;
                const int* my_starts = nullptr;
                if (a_tab) {
                    const int4 tl = a_tab[T >> TSH];
                    my_starts = a_starts + tl.x;
                    kb_count = tl.y;
                    cnt = GROUP == 4 ? min(4, m_tiles - 4 * T) : 1;
                    a_row0 = tl.z * 512 + (GROUP == 4 ? 0 : (T & 3) * BLOCK_M);
                    a_row_step = 512;
                    a_col_step = 0;
                }
                const bool resident = cnt > 1 && kb_count <= STAGES;
                if (a_tab && prefetch_tiles > 0) {
                    const int fg = grp + prefetch_tiles * (int)gridDim.x;
                    if (fg < total_groups) {
                        const int fT = fg / n_tiles;
                        const int4 fl = a_tab[fT >> TSH];
                        const int fcnt = GROUP == 4 ? min(4, m_tiles - 4 * fT) : 1;
                        const int frow = GROUP == 4 ? 0 : (fT & 3) * BLOCK_M;
                        for (int kb = 0; kb < fl.y; ++kb)
                            for (int i = 0; i < fcnt; ++i) tma_prefetch_2d(&map_a, 0, (fl.z + kb) * 512 + frow + i * BLOCK_M);
                    }
                }
                for (int i = 0; i < cnt; ++i) {
                    for (int kb = 0; kb < kb_count; ++kb) {
                        const bool load_b = !resident || i == 0;
                        mbar_wait(bar_empty + 8 * stage, phase ^ 1, err, 1);
                        int bslot = stage;     // groups of one: B travels with A in the same slot, one barrier pair
                        if (load_b && GROUP == 4) {
                            bslot = bpos;
                            mbar_wait(bar_bempty + 8 * bslot, ((bloads >> bslot) & 1u) ^ 1u, err, 5);
                            bloads ^= 1u << bslot;
                            if (++bpos == STAGES) bpos = 0;
                        }
                        mbar_expect_tx(bar_full + 8 * stage, load_b ? a_tx + b_tx : a_tx);
                        tma_load_2d(smem_u32(smem_a + stage * A_STAGE), &map_a, bar_full + 8 * stage, kb * a_col_step,
                                    a_row0 + i * BLOCK_M + kb * a_row_step);
                        if (load_b)
                            tma_load_2d(smem_u32(smem_b + bslot * B_STAGE), &map_b, bar_full + 8 * stage,
                                        my_starts ? my_starts[kb] : kb * BK, nt * block_n);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int bpos = 0;
            int it = 0;
            for (int grp = blockIdx.x; grp < total_groups; grp += gridDim.x) {
                const int T = grp / n_tiles, nt = grp % n_tiles;
                int kb_count = k_blocks, cnt = 1;
                if (a_tab) {
                    kb_count = a_tab[T >> TSH].y;
                    cnt = GROUP == 4 ? min(4, m_tiles - 4 * T) : 1;
                }
                const bool resident = cnt > 1 && kb_count <= STAGES;
                const int n_size = min(block_n, n_total - nt * block_n);
                const uint32_t idesc = make_idesc(BLOCK_M, n_size, (uint32_t)ab_bf16);
                const int b0 = bpos;
                for (int i = 0; i < cnt; ++i, ++it) {
                    const int acc = it & 1;
                    const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
                    mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1, err, 2);
                    tcgen05_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * MAX_BLOCK_N);
                    for (int kb = 0; kb < kb_count; ++kb) {
                        int bslot = stage;
                        if (resident) {
                            bslot = b0 + kb;
                            if (bslot >= STAGES) bslot -= STAGES;
                        } else if (GROUP == 4) {
                            bslot = bpos;
                            if (++bpos == STAGES) bpos = 0;
                        }
                        mbar_wait(bar_full + 8 * stage, phase, err, 3);
                        tcgen05_fence_after();
                        const uint64_t da = make_smem_desc_k<BK>(smem_u32(smem_a + stage * A_STAGE));
                        const uint64_t db = make_smem_desc_k<BK>(smem_u32(smem_b + bslot * B_STAGE));
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            // advance 32 bytes (16 fp16) along K inside the swizzle row: +2 in the >>4 address field
                            umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
                        }
                        umma_commit(bar_empty + 8 * stage);   // frees the A slot when these MMAs retire
                        if (GROUP == 4 && (!resident || i == cnt - 1)) umma_commit(bar_bempty + 8 * bslot);   // and the B slot after its last reader
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(bar_tfull + 8 * acc);  // accumulator complete
                }
                if (resident) {
                    bpos = b0 + kb_count;
                    if (bpos >= STAGES) bpos -= STAGES;
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue =====
        const int wq = warp & 3;          // TMEM lane quarter
        const int ch = (warp - 4) >> 2;   // which share of the accumulator's columns
        constexpr int SHARES = EPI_WARPS / 4;
        uint8_t* slab0 = smem_c + (warp - 4) * CBUFS * C_SLAB_BYTES;
        const float sc = scales[1];
        int it = 0;
        int buf = 0;
        for (int grp = blockIdx.x; grp < total_groups; grp += gridDim.x)
        for (int gi = 0, gcnt = min(GROUP, m_tiles - GROUP * (grp / n_tiles)); gi < gcnt; ++gi, ++it) {
            const int mt = GROUP * (grp / n_tiles) + gi, nt = grp % n_tiles;
            const int n_size = min(block_n, n_total - nt * block_n);
            // n_size is a multiple of 128, or 64 (then only share 0 works)
            const int c_lo = ch * (n_size / SHARES), c_hi = c_lo + n_size / SHARES;
            const int acc = it & 1;
            const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
            const bool split = n_size >= 64 * SHARES;
            const int c_base = split ? c_lo : 0;
            float* w_mine = w_s + (size_t)(warp - 4) * (MAX_BLOCK_N / SHARES) * (FC > 0 ? FC : 1);
            if (FC > 0 && gi == 0) {
                // this warp's slice of the weights for this group's N tile, fetched while the MMAs are still running
                const int ncols = split ? n_size / SHARES : (ch == 0 ? n_size : 0);
                const float* src = fuse_w + (size_t)(nt * block_n + c_base) * FC;
                for (int i = lane; i < ncols * FC; i += 32) w_mine[i] = __ldg(src + i) * sc;   // the GEMM's output scale folded in
                __syncwarp();
            }
            mbar_wait(bar_tfull + 8 * acc, acc_phase, err, 4);
            tcgen05_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * MAX_BLOCK_N);
            float dot[FC > 0 ? FC : 1][8];   // eight accumulation chains per channel (column mod 8), advanced two at a time by FFMA2
#pragma unroll
            for (int q = 0; q < (FC > 0 ? FC : 1); ++q)
#pragma unroll
                for (int u = 0; u < 8; ++u) dot[q][u] = 0.f;
            const int c_end = split ? c_hi : (ch == 0 ? n_size : 0);
            for (int c0 = (split ? c_lo : 0); c0 < c_end; c0 += 64) {
                uint8_t* slab = slab0 + buf * C_SLAB_BYTES;
                // both halves of the 64-column group in flight before the wait
                uint32_t v[2][32];
                tmem_ld_32x32b_x32(t_row + (uint32_t)c0, v[0]);
                tmem_ld_32x32b_x32(t_row + (uint32_t)(c0 + 32), v[1]);
                // the TMA store that last read this slab must be done with it
                if (store_d && lane == 0) tma_store_wait_read<CBUFS - 1>();
                tmem_ld_wait();
                __syncwarp();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t pk[16];
                    if (ST && addend != nullptr) {
                        // D = addend + scale * acc (orthonormalise.cu: Phi + Phi E); one 64-byte run of this thread's row
                        const int64_t row = (int64_t)mt * BLOCK_M + wq * 32 + lane;
                        const bool ok = row < m_rows;
                        const uint4* src = (const uint4*)(addend + (size_t)(ok ? row : 0) * n_total + nt * block_n + c0 + 32 * h);
#pragma unroll
                        for (int i4 = 0; i4 < 4; ++i4) {
                            const uint4 a = ok ? src[i4] : make_uint4(0, 0, 0, 0);
                            const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int i = 4 * i4 + j;
                                const float2 ad = unpack_h2(aw[j]);
                                pk[i] = pack_h2(fmaf(__uint_as_float(v[h][2 * i]), sc, ad.x), fmaf(__uint_as_float(v[h][2 * i + 1]), sc, ad.y));
                            }
                        }
                    } else {
                        if (store_d) {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                pk[i] = pack_h2(__uint_as_float(v[h][2 * i]) * sc, __uint_as_float(v[h][2 * i + 1]) * sc);
                        }
                        if (FC > 0) {
                            // 32 columns starting at nt * block_n + c0 + 32 h; weights (already times the output scale) read as
                            // broadcast 16-byte shared-memory loads: one FFMA per element and channel, nothing else
                            const uint32_t wv = smem_u32(w_mine) + (uint32_t)((c0 - c_base + 32 * h) * FC * 4);
#pragma unroll
                            for (int g8 = 0; g8 < 4; ++g8) {   // 8 columns = 8 * FC floats = 2 * FC float4
                                float wr[8 * (FC > 0 ? FC : 1)];
#pragma unroll
                                for (int q = 0; q < 2 * FC; ++q) lds_f4(wv + (uint32_t)((g8 * 2 * FC + q) * 16), &wr[4 * q]);
                                if (FC == 1) {
                                    // one channel: accumulator and weight pairs are register pairs already, two columns per FFMA2
#pragma unroll
                                    for (int i = 0; i < 8; i += 2)
                                        ffma2(dot[0][i], dot[0][i + 1], __uint_as_float(v[h][8 * g8 + i]),
                                              __uint_as_float(v[h][8 * g8 + i + 1]), wr[i], wr[i + 1]);
                                } else {
                                    // three channels: the weights of a column pair are not adjacent, pairing them costs more moves than
                                    // the FFMA2 saves (measured); four chains per channel
#pragma unroll
                                    for (int i = 0; i < 8; ++i) {
                                        const float val = __uint_as_float(v[h][8 * g8 + i]);
#pragma unroll
                                        for (int q = 0; q < FC; ++q) dot[q][i & 3] = fmaf(val, wr[i * FC + q], dot[q][i & 3]);
                                    }
                                }
                            }
                        }
                    }
                    // row `lane` of the slab, 16-byte chunks 4h .. 4h+3, XOR-swizzled like SWIZZLE_128B
                    if (store_d) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int chunk = (4 * h + c) ^ (lane & 7);
                            *(uint4*)(slab + lane * 128 + chunk * 16) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
                        }
                    }
                }
                if (store_d) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&map_d, smem_u32(slab), nt * block_n + c0, mt * BLOCK_M + wq * 32);
                        tma_store_commit();
                    }
                    if (CBUFS > 1) buf ^= 1;
                }
            }
            // all tcgen05.ld of this accumulator have completed (wait::ld above): hand it back to the MMA warp
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
            if (FC > 0) {
                const int64_t row = (int64_t)mt * BLOCK_M + wq * 32 + lane;
                if (row < m_rows) {   // a share that had no columns (narrow N tile) writes its zero
                    const int part = nt * SHARES + ch;
#pragma unroll
                    for (int q = 0; q < FC; ++q)
                        zpart[((size_t)part * m_rows + row) * FC + q] =
                            ((dot[q][0] + dot[q][1]) + (dot[q][2] + dot[q][3])) + ((dot[q][4] + dot[q][5]) + (dot[q][6] + dot[q][7]));
                }
            }
        }
        if (lane == 0) tma_store_wait_all<0>();
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// D[rows][n_pad] (fp16) = scales[1] * A[rows][k_pad] . Bt[n_pad][k_pad]^T (+ addend), A and Bt 16-bit K-major
// (ab_bf16: 0 = fp16, 1 = bf16).  k_pad % 64 == 0; n_pad is 64, 128 or a multiple of 256 (gl_m_pad).
int gl_gemm_kmajor(gl_ctx* ctx, const void* A, int ab_bf16, int64_t rows, int k_pad, const void* Bt, int n_pad,
                   const float* scales, const void* addend, void* D, const int4* a_tab, int64_t a_total_blocks, gl_gemm_fuse* fuse,
                   const int* a_starts, int a_kbs)
{
    GL_REQUIRE(a_kbs == 64 || (a_kbs == 32 && a_tab), "gemm: 32-slot blocks only exist for the blocked operand");
    GL_REQUIRE(!a_tab == !a_starts, "gemm: the blocked operand needs both its tile table and its block starts");
    GL_REQUIRE(k_pad % 64 == 0 && n_pad % 64 == 0, "gemm: k_pad %d / n_pad %d must be multiples of 64", k_pad, n_pad);
    GL_REQUIRE(!(fuse && ctx->gemm_impl == 1), "gemm: the CUDA-core checker has no fused filter");
    if (ctx->gemm_impl == 1) {
        dim3 grid((unsigned)ceil_div(n_pad, 32), (unsigned)ceil_div(rows, 32));
        GL_REQUIRE(ceil_div(rows, 32) < 2147483647ll, "gemm(simple): band too large");
        if (a_tab) {
            GL_REQUIRE(!ab_bf16 && !addend, "gemm(simple): the blocked operand is fp16 K_B without addend");
            k_gemm_simple_blocked<<<grid, 256, 0, ctx->stream>>>((const __half*)A, a_tab, a_starts, a_kbs, (const __half*)Bt, rows, n_pad,
                                                                 k_pad, scales, (__half*)D);
        } else if (ab_bf16)
            k_gemm_simple<__nv_bfloat16><<<grid, 256, 0, ctx->stream>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)Bt, rows, n_pad,
                                                                         k_pad, scales, (const __half*)addend,
                                                                         (__half*)D);
        else
            k_gemm_simple<__half><<<grid, 256, 0, ctx->stream>>>((const __half*)A, (const __half*)Bt, rows, n_pad, k_pad, scales,
                                                                  (const __half*)addend, (__half*)D);
        GL_LAUNCH_CHECK(ctx);
        return GL_OK;
    }
    const int block_n = n_pad < tc::MAX_BLOCK_N ? n_pad : tc::MAX_BLOCK_N;
    GL_REQUIRE(n_pad % block_n == 0, "gemm: n_pad %d is not a multiple of the N tile %d", n_pad, block_n);
    const CUtensorMapDataType dt = ab_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUtensorMap map_a, map_b, map_d;
    const CUtensorMapSwizzle sw = a_kbs == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
    if (a_tab)  // K_B's blocked storage seen as one tall [blocks * 512][slots per block] matrix
        GL_CHECK(make_map_2d(&map_a, dt, A, (uint64_t)a_total_blocks * 512, (uint64_t)a_kbs, (uint64_t)a_kbs, (uint32_t)a_kbs, tc::BLOCK_M, sw));
    else
        GL_CHECK(make_map_2d(&map_a, dt, A, (uint64_t)rows, (uint64_t)k_pad, (uint64_t)k_pad, tc::BLOCK_K, tc::BLOCK_M));
    GL_CHECK(make_map_2d(&map_b, dt, Bt, (uint64_t)n_pad, (uint64_t)k_pad, (uint64_t)k_pad, (uint32_t)a_kbs, (uint32_t)block_n, sw));
    // (no D: the store map is never used; it is encoded over A's storage only to have a valid descriptor)
    GL_CHECK(make_map_2d(&map_d, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, D ? D : A, (uint64_t)(D ? rows : 32), (uint64_t)(D ? n_pad : 64),
                         (uint64_t)(D ? n_pad : 64), 64, 32));
    const int m_tiles = (int)ceil_div(rows, tc::BLOCK_M);
    const int n_tiles = n_pad / block_n;
    const int k_blocks = k_pad / tc::BLOCK_K;
    gl_buf* err = nullptr;
    GL_CHECK(gl_alloc(ctx, sizeof(int) * 4, &err));
    cudaMemsetAsync(err->ptr, 0, sizeof(int) * 4, ctx->stream);
    int grid = ctx->sm_count;
    const int64_t groups = (int64_t)ceil_div(m_tiles, (a_tab && D == nullptr) ? 4 : 1) * n_tiles;   // the kernel's unit of work
    if ((int64_t)grid > groups) grid = (int)groups;
    // short K loops (blocked K_B with few blocks per tile): the epilogue is the critical path
    const bool short_k = a_tab != nullptr && a_total_blocks * 4 * a_kbs < (int64_t)m_tiles * 8 * 64;   // < 512 K elements per M tile on average
    const int pf = short_k ? ctx->gemm_prefetch : 0;   // with long K loops the ring already covers the latency; prefetch only adds L2 churn
    const bool deep = !(ctx->gemm_stages == 3 || (ctx->gemm_stages == 0 && short_k));
    const int FCH = fuse ? fuse->C : 0;
    const int w_bytes = FCH * 4 * tc::MAX_BLOCK_N * (int)sizeof(float);   // per epilogue warp: its columns of one tile
    if (fuse) {
        GL_REQUIRE(!addend && (FCH == 1 || FCH == 3), "gemm: fused filter wants 1 or 3 channels and no addend");
        GL_REQUIRE(ctx->gemm_impl == 0, "gemm: the CUDA-core checker has no fused filter");
        GL_REQUIRE((deep ? tc::smem_bytes<4, 1, 4>() : tc::smem_bytes<3, 2, 8>()) + w_bytes <= tc::SMEM_LIMIT,
                   "gemm: no shared memory left for the filter weights");
        fuse->parts = n_tiles * (deep ? 1 : 2);
    }
    const float* fw = fuse ? fuse->w : nullptr;
    float* zp = fuse ? fuse->zpart : nullptr;
    const int store_d = D != nullptr;
    GL_REQUIRE(store_d || fuse, "gemm: no output requested");
    StageTimer kt(ctx, GL_T_K_GEMM);
#define GEMM_LAUNCH_ST(S, CB, EW, FC, BK, ST)                                                                                   \
    do {                                                                                                                       \
        const int SM = tc::smem_bytes<S, CB, EW, BK>() + w_bytes;                                                              \
        GL_CUDA_CHECK(cudaFuncSetAttribute(tc::k_gemm_tcgen05<S, CB, EW, FC, BK, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM)); \
        tc::k_gemm_tcgen05<S, CB, EW, FC, BK, ST><<<grid, 128 + 32 * EW, SM, ctx->stream>>>(                                    \
            map_a, map_b, map_d, m_tiles, n_tiles, k_blocks, n_pad, block_n, ab_bf16, scales, (const __half*)addend, rows, a_tab, \
            a_starts, pf, fw, zp, (int*)err->ptr);                                                                             \
    } while (0)
    // 32-slot blocks: stages are half as large, so the rings are twice as deep
#define GEMM_LAUNCH(S, CB, EW, FC, ST)                                \
    do {                                                              \
        if (a_kbs == 32) GEMM_LAUNCH_ST(2 * S, CB, EW, FC, 32, ST);   \
        else GEMM_LAUNCH_ST(S, CB, EW, FC, 64, ST);                   \
    } while (0)
    if (!deep) {
        if (FCH == 1 && store_d) GEMM_LAUNCH(3, 2, 8, 1, true);
        else if (FCH == 1) GEMM_LAUNCH(3, 2, 8, 1, false);
        else if (FCH == 3 && store_d) GEMM_LAUNCH(3, 2, 8, 3, true);
        else if (FCH == 3) GEMM_LAUNCH(3, 2, 8, 3, false);
        else GEMM_LAUNCH(3, 2, 8, 0, true);
    } else {
        // the deep ring leaves room for the weights only with single store slabs (the epilogue has slack here)
        if (FCH == 1 && store_d) GEMM_LAUNCH(4, 1, 4, 1, true);
        else if (FCH == 1) GEMM_LAUNCH(4, 1, 4, 1, false);
        else if (FCH == 3 && store_d) GEMM_LAUNCH(4, 1, 4, 3, true);
        else if (FCH == 3) GEMM_LAUNCH(4, 1, 4, 3, false);
        else GEMM_LAUNCH(4, 2, 4, 0, true);
    }
#undef GEMM_LAUNCH
#undef GEMM_LAUNCH_ST
    gl_buf_release(err);
    GL_LAUNCH_CHECK(ctx);
    return GL_OK;
}

// shape and bookkeeping of the Phi that L_B, phi_A produce (no storage yet)
void gl_phi_describe(gl_ctx* ctx, gl_mat* phi, const gl_mat* L_B, const gl_mat* phi_A)
{
    const int m = (int)phi_A->cols;
    phi->rows = ctx->n;
    phi->cols = m;
    phi->local_rows = L_B->local_rows;
    phi->ld = gl_m_pad(m);
    phi->elem_bytes = 2;
    phi->p = L_B->p;
    phi->p_pad = L_B->p_pad;
    phi->m = m;
    phi->m_pad = gl_m_pad(m);
    phi->q0 = L_B->q0;
}

// Computes Phi into the handle `phi` (described by gl_phi_describe); keep_phi = false: only the fused filter output `ff`.
int gl_impl_nystroem_into(gl_ctx* ctx, gl_mat* L_B, gl_mat* phi_A, gl_mat* eigvals_inv, gl_mat* phi, bool keep_phi, const gl_fused_filter* ff)
{
    const int p = L_B->p, p_pad = L_B->p_pad;
    const int m = (int)phi_A->cols;
    const int m_pad = gl_m_pad(m);
    const int64_t rows = L_B->local_rows;
    GL_REQUIRE(L_B->dscale, "nystroem: expected L_B (from gl_laplacian), got a bare K_B");
    GL_REQUIRE(rows > 0, "nystroem: empty band");
    GL_REQUIRE(keep_phi || ff, "nystroem: nothing to return");
    // the sample rows of Phi, the projection and the filter's sample rows read ctx->samples: they must still be the samples
    // K_B (and the eigenvectors) were computed from -- a deferred Phi may be computed long after gl_nystroem returned
    GL_REQUIRE(L_B->sample_epoch == ctx->sample_epoch, "nystroem: the samples changed since this K_B was computed");
    // K_B in its patch layout (patch.cu) feeds the fused extrapolation + filter directly; everything that needs Phi itself goes
    // through the blocked storage, computed on demand when the handle does not hold it yet
    const bool use_patch = L_B->pt_buf != nullptr && ff != nullptr && !keep_phi && ctx->gemm_impl == 0 &&
                           gl_patch_nystroem_fits(m, ctx->channels);
    if (!use_patch) GL_CHECK(gl_kb_require_blocked(ctx, L_B));
    GL_REQUIRE(use_patch || (L_B->tiles && L_B->starts && L_B->perm), "nystroem: K_B handle without its block layout");

    gl_buf *Wt = nullptr, *colmax = nullptr, *scales = nullptr;
    int rc = GL_OK;
    do {
        if (keep_phi && (rc = gl_alloc(ctx, sizeof(__half) * (size_t)rows * m_pad, &phi->buf)) != GL_OK) break;
        const int k_dim = p_pad + 64;   // K_B's internal sample slots
        if (!use_patch && (rc = gl_alloc(ctx, sizeof(__half) * (size_t)m_pad * k_dim, &Wt)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(float) * (size_t)m_pad, &colmax)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(float) * 4, &scales)) != GL_OK) break;

        const float* U = (const float*)phi_A->buf->ptr;
        const double* mu_inv = (const double*)eigvals_inv->buf->ptr;
        const double* neg_alpha = (const double*)L_B->dscale->ptr;
        k_w_colmax<<<m, 256, 0, ctx->stream>>>(U, (int)phi_A->ld, p, m, mu_inv, neg_alpha, (float*)colmax->ptr);
        GL_LAUNCH_CHECK(ctx);
        k_w_scale<<<1, 1024, 0, ctx->stream>>>((const float*)colmax->ptr, m, (float*)scales->ptr);
        GL_LAUNCH_CHECK(ctx);
        if (!use_patch) {
            k_w_write<<<m_pad, 256, 0, ctx->stream>>>(U, (int)phi_A->ld, m, k_dim, (const uint32_t*)L_B->perm->ptr, mu_inv, neg_alpha,
                                                      (const float*)scales->ptr, (__half*)Wt->ptr);
            GL_LAUNCH_CHECK(ctx);
        }

        // projection c = Phi^T y from the affinity sums (valid while the image and the samples are unchanged)
        if (L_B->aux && L_B->channels == ctx->channels && L_B->image_epoch == ctx->image_epoch) {
            const int C = ctx->channels;
            if (phi->proj) { gl_buf_release(phi->proj); phi->proj = nullptr; }
            if ((rc = gl_alloc(ctx, sizeof(double) * (size_t)m_pad * C, &phi->proj)) != GL_OK) break;
            const double* kay = (const double*)L_B->aux->ptr + (size_t)(1 + C) * p_pad;   // (K_A y_S), behind D and T
            k_proj_from_sums<<<m_pad, 128, 0, ctx->stream>>>(U, (int)phi_A->ld, p, p_pad, m, mu_inv, neg_alpha,
                                                            (const double*)L_B->aux->ptr, kay, (const uint8_t*)ctx->img->ptr,
                                                            (const uint32_t*)ctx->samples->ptr, C, (double*)phi->proj->ptr);
            ctx->launches++;
            phi->channels = C;
            phi->image_epoch = ctx->image_epoch;
        }
        gl_gemm_fuse fuse;
        gl_buf *wbuf = nullptr, *zpart = nullptr;
        const bool do_fuse = ff != nullptr;
        if (do_fuse) {
            // filter weights before the GEMM: w = gain * f(lambda) o c with c from the affinity sums
            const int C = ctx->channels;
            if (!phi->proj) { gl_set_error("nystroem: fused filter needs the affinity sums of the current image"); rc = GL_ERR_ARG; break; }
            if ((rc = gl_alloc(ctx, sizeof(float) * (size_t)m_pad * C, &wbuf)) != GL_OK) break;
            const int parts_max = 2 * (m_pad / (m_pad < 256 ? m_pad : 256));
            if (!use_patch && (rc = gl_alloc(ctx, sizeof(float) * (size_t)parts_max * rows * C, &zpart)) != GL_OK) { gl_buf_release(wbuf); break; }
            if ((rc = gl_filter_weights_from_proj(ctx, (const double*)phi->proj->ptr, (const double*)ff->f_eigvals->buf->ptr, ff->gain, m, m_pad,
                                                  C, (float*)wbuf->ptr)) != GL_OK) { gl_buf_release(wbuf); gl_buf_release(zpart); break; }
            fuse.w = (const float*)wbuf->ptr;
            fuse.zpart = zpart ? (float*)zpart->ptr : nullptr;
            fuse.C = C;
        }
        if (use_patch) {
            // the patch kernel writes the band's filtered pixels itself; the finish only patches the sample pixels' rows
            const int C = ctx->channels;
            gl_buf *zd = nullptr, *z8d = nullptr;
            // A u8 destination in pinned host memory is written by the kernel itself, pixel by pixel as the patches finish (32 contiguous
            // bytes per warp store over PCIe): the 8 MB device-to-host copy that followed the kernel disappears under it.
            uint8_t* z8_direct = nullptr;
            // (one channel only: three interleaved byte stores per pixel turn into partial-sector writes over PCIe -- config 5 on 8 GPUs
            // fell from 8 107 to 2 556 Mpixel/s end to end with them)
            if (ff->z_u8 && ctx->z8_direct && C == 1) {
                cudaPointerAttributes at;
                if (cudaPointerGetAttributes(&at, ff->z_u8) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
                    z8_direct = (uint8_t*)at.devicePointer + (size_t)L_B->q0 * C;
                else
                    (void)cudaGetLastError();
            }
            if ((rc = gl_alloc(ctx, sizeof(float) * (size_t)rows * C, &zd)) == GL_OK && ff->z_u8 && !z8_direct) rc = gl_alloc(ctx, (size_t)rows * C, &z8d);
            if (rc == GL_OK)
                rc = gl_patch_nystroem_filter(ctx, L_B, U, (int)phi_A->ld, m, mu_inv, (const float*)scales->ptr, (const float*)wbuf->ptr, C,
                                              ff->clip_low, (float*)zd->ptr, z8_direct ? z8_direct : (z8d ? (uint8_t*)z8d->ptr : nullptr));
            if (rc == GL_OK)
                rc = gl_filter_fused_finish(ctx, phi, nullptr, 0, (const float*)wbuf->ptr, U, (int)phi_A->ld, ff->clip_low, ff->z_f32, ff->z_u8,
                                            zd, z8d, z8_direct);
            else { if (zd) gl_buf_release(zd); if (z8d) gl_buf_release(z8d); }
        } else {
            rc = gl_gemm_kmajor(ctx, L_B->buf->ptr, 0, rows, k_dim, Wt->ptr, m_pad, (const float*)scales->ptr, nullptr,
                                keep_phi ? phi->buf->ptr : nullptr, (const int4*)L_B->tiles->ptr, L_B->total_blocks, do_fuse ? &fuse : nullptr,
                                (const int*)L_B->starts->ptr, L_B->kbs);
            if (rc == GL_OK && do_fuse)
                rc = gl_filter_fused_finish(ctx, phi, (const float*)zpart->ptr, fuse.parts, (const float*)wbuf->ptr, U, (int)phi_A->ld,
                                            ff->clip_low, ff->z_f32, ff->z_u8);
        }
        if (wbuf) gl_buf_release(wbuf);
        if (zpart) gl_buf_release(zpart);
        if (rc != GL_OK) break;

        if (keep_phi) {
            k_phi_sample_rows<<<p, 128, 0, ctx->stream>>>(U, (int)phi_A->ld, p, m, (const uint32_t*)ctx->samples->ptr, phi->q0,
                                                          phi->q0 + rows, m_pad, (__half*)phi->buf->ptr);
            GL_LAUNCH_CHECK(ctx);
        }
    } while (0);
    if (Wt) gl_buf_release(Wt);
    if (colmax) gl_buf_release(colmax);
    if (scales) gl_buf_release(scales);
    return rc;
}

int gl_impl_nystroem(gl_ctx* ctx, gl_mat* L_B, gl_mat* phi_A, gl_mat* eigvals_inv, gl_mat** phi_out, const gl_fused_filter* ff)
{
    gl_mat* phi = gl_mat_new(ctx, GL_MAT_PHI);
    gl_phi_describe(ctx, phi, L_B, phi_A);
    const int rc = gl_impl_nystroem_into(ctx, L_B, phi_A, eigvals_inv, phi, phi_out != nullptr, ff);
    if (rc != GL_OK || !phi_out) {
        gl_mat_destroy(phi);
        return rc;
    }
    *phi_out = phi;
    return GL_OK;
}

// Nystroem() deferred: the handle only remembers its inputs (retained); gl_phi_materialise computes it when somebody needs
// the matrix -- and when that somebody is the filter, the two stages run as one pass (gl_filter, api.cu)
int gl_phi_defer(gl_ctx* ctx, gl_mat* L_B, gl_mat* phi_A, gl_mat* eigvals_inv, gl_mat** phi_out)
{
    GL_REQUIRE(L_B->dscale, "nystroem: expected L_B (from gl_laplacian), got a bare K_B");
    GL_REQUIRE(L_B->local_rows > 0, "nystroem: empty band");
    gl_mat* phi = gl_mat_new(ctx, GL_MAT_PHI);
    gl_phi_describe(ctx, phi, L_B, phi_A);
    phi->def_LB = L_B;
    phi->def_U = phi_A;
    phi->def_muinv = eigvals_inv;
    L_B->refs++;
    phi_A->refs++;
    eigvals_inv->refs++;
    *phi_out = phi;
    return GL_OK;
}

int gl_phi_materialise(gl_ctx* ctx, gl_mat* phi, const gl_fused_filter* ff)
{
    if (!phi->def_LB) return GL_OK;
    gl_mat *LB = phi->def_LB, *U = phi->def_U, *mi = phi->def_muinv;
    // with the filter riding along, Phi is consumed in the epilogue and stays deferred: nobody has asked for the matrix itself
    // yet, and whoever does later (download, column dumps, orthonormalisation) makes this call again without a filter.
    // Option keep_phi=1 stores it in the same pass -- unless it does not fit or exceeds option phi_limit_mb.
    bool keep = ff ? ctx->keep_phi != 0 : true;
    const size_t bytes = (size_t)phi->local_rows * (size_t)phi->m_pad * 2;
    if (ff && ctx->phi_limit_mb > 0 && bytes > (size_t)ctx->phi_limit_mb << 20) keep = false;
    if (ff && ctx->phi_nomem_bytes && bytes >= ctx->phi_nomem_bytes) keep = false;
    int rc = gl_impl_nystroem_into(ctx, LB, U, mi, phi, keep, ff);
    if (rc == GL_ERR_NOMEM && keep && ff) {
        keep = false;
        ctx->phi_nomem_bytes = bytes;
        rc = gl_impl_nystroem_into(ctx, LB, U, mi, phi, false, ff);
    }
    if (rc != GL_OK) return rc;
    if (!keep) return GL_OK;
    phi->def_LB = phi->def_U = phi->def_muinv = nullptr;
    gl_mat_destroy(LB);
    gl_mat_destroy(U);
    gl_mat_destroy(mi);
    return GL_OK;
}
