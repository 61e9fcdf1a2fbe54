// a-8: orthonormalisation of the columns of Phi (n x m).
// Replaces OrthonormaliseVecs / Projection, hpc/gram_schmidt.c:11-64: classical Gram-Schmidt,
//   u_k = v_k - sum_{j<k} <v_k,u_j>/<u_j,u_j> u_j,  u_k /= |u_k|,  norms[k] = |u_k| before normalising,
// which the reference runs as m^2 VecDot calls (two allreduces each).  In exact arithmetic that is the QR
// factorisation Phi = Q R with R upper triangular and a positive diagonal, norms = diag(R).  Computed here as
// a blocked CholeskyQR, which needs ONE reduction over pixels (and over ranks) instead of m^2:
//   1. G = Phi^T Phi        pixel-parallel partial dot products (warp/CTA tiles), fixed-order slab reduction,
//                           one m_pad^2 allreduce;
//   2. R = chol(G), T = R^-1   m x m, fp64, on device (replicated, deterministic);
//   3. Q = Phi + Phi (T - I)   tensor-core GEMM (nystroem_gemm.cu) with the identity part added in the
//                           epilogue, so that only the small correction E = T - I is rounded to fp16.
// Phi is nearly orthonormal on entry (|Phi^T Phi - I|_F ~ 1e-3..1e-2), so G is well
// conditioned and one CholeskyQR pass is stable; the result is orthonormal up to the fp16 storage of Q.
#include "tc_common.cuh"

// ---- 1. Gram matrix: CTA = 64 x 64 tile of G over a slab of rows; 256 threads, 4 x 4 outputs each ----------
__global__ void __launch_bounds__(256) k_gram_tile(const __half* __restrict__ phi, int64_t rows, int m_pad, int slabs,
                                                   float* __restrict__ partial /* [slabs][m_pad][m_pad] */)
{
    __shared__ float As[32][64 + 4], Bs[32][64 + 4];
    const int ti = blockIdx.x, tj = blockIdx.y, slab = blockIdx.z;
    if (tj < ti) return;  // upper triangle only; mirrored by k_gram_reduce
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t per = (rows + slabs - 1) / slabs;
    const int64_t r_begin = per * slab, r_end = min(rows, r_begin + per);
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    for (int64_t r0 = r_begin; r0 < r_end; r0 += 32) {
        // 32 rows x 64 columns of each operand: 256 threads x one 16-byte load (8 bf16) each
        {
            const int rr = threadIdx.x >> 3, cg = threadIdx.x & 7;
            const int64_t r = r0 + rr;
            uint4 va = make_uint4(0, 0, 0, 0), vb = make_uint4(0, 0, 0, 0);
            if (r < r_end) {
                va = *(const uint4*)(phi + (size_t)r * m_pad + ti * 64 + cg * 8);
                vb = *(const uint4*)(phi + (size_t)r * m_pad + tj * 64 + cg * 8);
            }
            const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 fa = __half22float2(*(const __half2*)&wa[k]), fb = __half22float2(*(const __half2*)&wb[k]);
                As[rr][cg * 8 + 2 * k] = fa.x;
                As[rr][cg * 8 + 2 * k + 1] = fa.y;
                Bs[rr][cg * 8 + 2 * k] = fb.x;
                Bs[rr][cg * 8 + 2 * k + 1] = fb.y;
            }
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            const float4 a = *(const float4*)&As[k][ty * 4];
            const float4 b = *(const float4*)&Bs[k][tx * 4];
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(av[x], bv[y], acc[x][y]);
        }
        __syncthreads();
    }
    float* out = partial + (size_t)slab * m_pad * m_pad;
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) out[(size_t)(ti * 64 + ty * 4 + x) * m_pad + tj * 64 + tx * 4 + y] = acc[x][y];
}

// ---- 1b. the same Gram matrix on the tensor cores (m_pad a multiple of 256) -----------------------------------------
// C = Phi^T Phi with BOTH operands read straight from the row-major Phi: element (k, i) of a 64-row x 64-column TMA box is
// A^T's (i, k), i.e. the operands are "MN-major" for tcgen05 (M/N contiguous, K strided), which the shared-memory
// descriptor expresses (tc::make_smem_desc_mn) -- no transposed copy of Phi is ever made.
//   work item = (K slab of `kb_per_item` 64-row blocks, 128 x 256 tile of the upper triangle of C); fp32 accumulation in
//   TMEM over at most kb_per_item * 64 rows, partial tiles written to partial[slab] and summed in fp64 by k_gram_reduce;
//   items are ordered slab-major so that the tiles of one slab run at the same time and share its rows through L2.
// warp 0: TMA producer (6 boxes per stage: 2 for the 128 A columns, 4 for the 256 B columns), warp 1: MMA issuer,
// warp 2: TMEM allocator, warps 4-7: epilogue (tcgen05.ld -> fp32 stores).
namespace gram {
constexpr int STAGES = 4;
constexpr int CHUNK_BYTES = 64 * 128;                        // one TMA box: 64 K rows x 64 columns fp16
constexpr int STAGE_BYTES = 6 * CHUNK_BYTES;                 // 48 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;
constexpr int THREADS = 256;

__global__ void __launch_bounds__(THREADS, 1)
k_gram_tcgen05(const __grid_constant__ CUtensorMap map_phi, int64_t rows, int m_pad, int kb_per_item, int n_slabs,
               float* __restrict__ partial /* [n_slabs][m_pad][m_pad] */, uint32_t lbo_bytes, int* __restrict__ err)
{
    using namespace tc;
    extern __shared__ uint8_t gram_smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)gram_smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + STAGES * STAGE_BYTES);
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + STAGES);
    const uint32_t bar_tfull = smem_u32(bars + 2 * STAGES), bar_tempty = smem_u32(bars + 2 * STAGES + 2);
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STAGES + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int MT = m_pad / 128, NT = m_pad / 256;
    // upper-triangle tiles: (ti, tj) with tj >= ti / 2; enumerated row by row
    int ntiles = 0;
    for (int ti = 0; ti < MT; ++ti) ntiles += NT - ti / 2;
    const int total = n_slabs * ntiles;
    const int64_t kb_total = (rows + 63) / 64;

    if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_phi) : "memory");
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto decode = [&](int item, int& slab, int& ti, int& tj) {
        slab = item / ntiles;
        int t = item - slab * ntiles;
        ti = 0;
        while (t >= NT - ti / 2) { t -= NT - ti / 2; ++ti; }
        tj = ti / 2 + t;
    };

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < total; item += gridDim.x) {
                int slab, ti, tj;
                decode(item, slab, ti, tj);
                const int64_t kb0 = (int64_t)slab * kb_per_item, kb1 = min(kb_total, kb0 + kb_per_item);
                for (int64_t kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1, err, 1);
                    mbar_expect_tx(bar_full + 8 * stage, STAGE_BYTES);
                    const uint32_t dst = smem_u32(smem + stage * STAGE_BYTES);
                    const int r = (int)(kb * 64);
#pragma unroll
                    for (int c = 0; c < 2; ++c) tma_load_2d(dst + c * CHUNK_BYTES, &map_phi, bar_full + 8 * stage, ti * 128 + c * 64, r);
#pragma unroll
                    for (int c = 0; c < 4; ++c) tma_load_2d(dst + (2 + c) * CHUNK_BYTES, &map_phi, bar_full + 8 * stage, tj * 256 + c * 64, r);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            const uint32_t idesc = make_idesc_mn(128, 256, 0);
            for (int item = blockIdx.x; item < total; item += gridDim.x, ++it) {
                int slab, ti, tj;
                decode(item, slab, ti, tj);
                const int64_t kb0 = (int64_t)slab * kb_per_item, kb1 = min(kb_total, kb0 + kb_per_item);
                const int acc = it & 1;
                mbar_wait(bar_tempty + 8 * acc, (uint32_t)(((it >> 1) & 1) ^ 1), err, 2);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
                for (int64_t kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(bar_full + 8 * stage, phase, err, 3);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                    const uint64_t da = make_smem_desc_mn(sa, lbo_bytes);
                    const uint64_t db = make_smem_desc_mn(sa + 2 * CHUNK_BYTES, lbo_bytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k)  // 16 K rows = two 8-row groups = 2048 bytes further
                        umma_f16(d_tmem, da + (uint64_t)(128 * k), db + (uint64_t)(128 * k), idesc, (uint32_t)((kb > kb0) || k != 0));
                    umma_commit(bar_empty + 8 * stage);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(bar_tfull + 8 * acc);
            }
        }
    } else if (warp >= 4) {
        const int wq = warp & 3;
        int it = 0;
        for (int item = blockIdx.x; item < total; item += gridDim.x, ++it) {
            int slab, ti, tj;
            decode(item, slab, ti, tj);
            const int acc = it & 1;
            mbar_wait(bar_tfull + 8 * acc, (uint32_t)((it >> 1) & 1), err, 4);
            tcgen05_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * 256);
            float* out = partial + ((size_t)slab * m_pad + ti * 128 + wq * 32 + lane) * m_pad + tj * 256;
            for (int c0 = 0; c0 < 256; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(t_row + (uint32_t)c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    *(float4*)(out + c0 + 4 * i) = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                                __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}
}  // namespace gram

// G[i][j] (fp64, full symmetric) = sum over slabs in fixed order; lower triangle mirrored
__global__ void k_gram_reduce(const float* __restrict__ partial, int slabs, int m_pad, double* __restrict__ G)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= m_pad) return;
    const int a = min(i, j) , b = max(i, j);
    // tile (a/64, b/64) was computed iff a/64 <= b/64, always true for a <= b
    double s = 0.0;
    for (int k = 0; k < slabs; ++k) s += (double)partial[(size_t)k * m_pad * m_pad + (size_t)a * m_pad + b];
    G[(size_t)i * m_pad + j] = s;
}

// ---- 2. R = chol(G) (upper, G = R^T R) and T = R^-1, fp64, blocked by 32 and spread over the GPU ---------------------
// Right-looking blocked Cholesky: per 32-column block k  (a) one CTA factors the diagonal block in shared memory and
// also inverts it (T_kk = R_kk^-1),  (b) the block row right of it becomes R_kj = T_kk^T G_kj,  (c) the trailing
// upper triangle is updated G_ij -= R_ki^T R_kj.  Rows/columns >= m (padding) behave like the identity.
#define CB 32

__global__ void __launch_bounds__(CB* CB) k_chol_diag(double* __restrict__ G, int ld, int m, int kb, double* __restrict__ Tdiag,
                                                      int* __restrict__ status)
{
    __shared__ double S[CB][CB + 1], Ti[CB][CB + 1];
    const int tx = threadIdx.x % CB, ty = threadIdx.x / CB;
    const int gi = kb * CB + ty, gj = kb * CB + tx;
    S[ty][tx] = (gi < m && gj < m) ? G[(size_t)gi * ld + gj] : (gi == gj ? 1.0 : 0.0);
    __syncthreads();
    for (int k = 0; k < CB; ++k) {
        if (tx == k && ty == k) {
            const double d = S[k][k];
            if (!(d > 0.0)) {
                if (*status == 0) *status = kb * CB + k + 1;
                S[k][k] = 1.0;
            } else {
                S[k][k] = sqrt(d);
            }
        }
        __syncthreads();
        if (ty == k && tx > k) S[k][tx] /= S[k][k];
        __syncthreads();
        if (ty > k && tx >= ty) S[ty][tx] -= S[k][ty] * S[k][tx];
        __syncthreads();
    }
    if (gi < m && gj < m && tx >= ty) G[(size_t)gi * ld + gj] = S[ty][tx];
    // T_kk = R_kk^-1: column tx by back substitution (one thread per column)
    Ti[ty][tx] = 0.0;
    __syncthreads();
    if (ty == 0) {
        const int j = tx;
        Ti[j][j] = 1.0 / S[j][j];
        for (int i = j - 1; i >= 0; --i) {
            double acc = 0.0;
            for (int k = i + 1; k <= j; ++k) acc += S[i][k] * Ti[k][j];
            Ti[i][j] = -acc / S[i][i];
        }
    }
    __syncthreads();
    Tdiag[((size_t)kb * CB + ty) * CB + tx] = Ti[ty][tx];
}

// R_kj = T_kk^T G_kj for the block columns j > k (one CTA each)
__global__ void __launch_bounds__(CB* CB) k_chol_row(double* __restrict__ G, int ld, int m, int kb, const double* __restrict__ Tdiag)
{
    __shared__ double Tk[CB][CB + 1], Gb[CB][CB + 1];
    const int tx = threadIdx.x % CB, ty = threadIdx.x / CB;
    const int jb = kb + 1 + blockIdx.x;
    const int gi = kb * CB + ty, gj = jb * CB + tx;
    Tk[ty][tx] = Tdiag[((size_t)kb * CB + ty) * CB + tx];
    Gb[ty][tx] = (gi < m && gj < m) ? G[(size_t)gi * ld + gj] : 0.0;
    __syncthreads();
    double acc = 0.0;
#pragma unroll 8
    for (int t = 0; t < CB; ++t) acc = fma(Tk[t][ty], Gb[t][tx], acc);
    if (gi < m && gj < m) G[(size_t)gi * ld + gj] = acc;
}

// G_ij -= R_ki^T R_kj for k < i <= j (blockIdx.x enumerates the upper triangle of the trailing blocks)
__global__ void __launch_bounds__(CB* CB) k_chol_trail(double* __restrict__ G, int ld, int m, int kb, int nrem)
{
    __shared__ double Ri[CB][CB + 1], Rj[CB][CB + 1];
    const int tx = threadIdx.x % CB, ty = threadIdx.x / CB;
    int bi = 0, rem = blockIdx.x;
    while (rem >= nrem - bi) { rem -= nrem - bi; ++bi; }
    const int ib = kb + 1 + bi, jb = ib + rem;
    const int r = kb * CB + ty;
    Ri[ty][tx] = (r < m && ib * CB + tx < m) ? G[(size_t)r * ld + ib * CB + tx] : 0.0;
    Rj[ty][tx] = (r < m && jb * CB + tx < m) ? G[(size_t)r * ld + jb * CB + tx] : 0.0;
    __syncthreads();
    double acc = 0.0;
#pragma unroll 8
    for (int t = 0; t < CB; ++t) acc = fma(Ri[t][ty], Rj[t][tx], acc);
    const int gi = ib * CB + ty, gj = jb * CB + tx;
    if (gi < m && gj < m) G[(size_t)gi * ld + gj] -= acc;
}

// T = R^-1 (upper): one CTA per block column j, block rows from the diagonal upwards:
//   T_jj = R_jj^-1,   T_ij = -T_ii sum_{k=i+1..j} R_ik T_kj.   T (row-major [m][ld]) must be zeroed beforehand.
__global__ void __launch_bounds__(CB* CB) k_upper_inverse_blocked(const double* __restrict__ R, int m, int ld, const double* __restrict__ Tdiag,
                                                                  double* __restrict__ T)
{
    __shared__ double A[CB][CB + 1], B[CB][CB + 1];
    const int tx = threadIdx.x % CB, ty = threadIdx.x / CB;
    const int jb = blockIdx.x;
    {
        const int gi = jb * CB + ty, gj = jb * CB + tx;
        if (gi < m && gj < m) T[(size_t)gi * ld + gj] = Tdiag[((size_t)jb * CB + ty) * CB + tx];
    }
    __syncthreads();
    for (int ib = jb - 1; ib >= 0; --ib) {
        double acc = 0.0;
        for (int kb = ib + 1; kb <= jb; ++kb) {
            const int ar = ib * CB + ty, ac = kb * CB + tx;     // R_ik
            const int br = kb * CB + ty, bc = jb * CB + tx;     // T_kj
            A[ty][tx] = (ar < m && ac < m) ? R[(size_t)ar * ld + ac] : 0.0;
            B[ty][tx] = (br < m && bc < m) ? T[(size_t)br * ld + bc] : 0.0;
            __syncthreads();
#pragma unroll 8
            for (int t = 0; t < CB; ++t) acc = fma(A[ty][t], B[t][tx], acc);
            __syncthreads();
        }
        // T_ij = -T_ii acc
        A[ty][tx] = Tdiag[((size_t)ib * CB + ty) * CB + tx];
        B[ty][tx] = acc;
        __syncthreads();
        double v = 0.0;
#pragma unroll 8
        for (int t = 0; t < CB; ++t) v = fma(A[ty][t], B[t][tx], v);
        const int gi = ib * CB + ty, gj = jb * CB + tx;
        if (gi < m && gj < m) T[(size_t)gi * ld + gj] = -v;
        __syncthreads();   // T_ij is read back (as T_kj) in the next block row; also protects A/B
    }
}

// Et[j][k] = fp16(2^12 * (T[k][j] - delta_kj)) for k, j < m; 0 in the padding (K-major B operand of Q = Phi + Phi E).
// E is ~1e-3: the 2^12 keeps it in fp16's normal range; the GEMM epilogue multiplies by 2^-12.
#define ET_SCALE_LOG2 12
__global__ void k_build_et(const double* __restrict__ T, int m, int ld, int m_pad, __half* __restrict__ Et,
                           const double* __restrict__ R, double* __restrict__ norms)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (k >= m_pad) return;
    double v = 0.0;
    if (k < m && j < m) v = T[(size_t)k * ld + j] - (k == j ? 1.0 : 0.0);
    v = ldexp(v, ET_SCALE_LOG2);
    v = fmin(fmax(v, -60000.0), 60000.0);
    Et[(size_t)j * m_pad + k] = __float2half_rn((float)v);
    if (norms && k == j && j < m) norms[j] = R[(size_t)j * ld + j];
}

// Q = Phi T  =>  Q^T y = T^T (Phi^T y): out[j][ch] = sum_{k<=j} T[k][j] c[k][ch]
__global__ void k_proj_times_t(const double* __restrict__ T, int m, int ld, int m_pad, int C, const double* __restrict__ c,
                               double* __restrict__ out)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m_pad) return;
    for (int ch = 0; ch < C; ++ch) {
        double s = 0.0;
        if (j < m)
            for (int k = 0; k <= j; ++k) s += T[(size_t)k * ld + j] * c[(size_t)k * C + ch];
        out[(size_t)j * C + ch] = s;
    }
}

__global__ void k_et_scales(float* s)
{
    s[0] = ldexpf(1.f, ET_SCALE_LOG2);
    s[1] = ldexpf(1.f, -ET_SCALE_LOG2);
}

// R = chol(G) in place (upper triangle of the m x m fp64 matrix G with leading dimension ld, G = R^T R) and T = R^-1 (upper,
// same shape, zeroed here).  *status_dev (zeroed by the caller) receives the first non-positive pivot + 1.  Used by the
// orthonormalisation (CholeskyQR) and by the inverse subspace iteration (A^-1 = T T^T, dense_small.cu).
int gl_chol_inverse_upper(gl_ctx* ctx, double* G, int m, int ld, double* T, int* status_dev)
{
    const int nblk = (int)ceil_div(m, CB);
    gl_buf* Tdiag = nullptr;
    GL_CHECK(gl_alloc(ctx, sizeof(double) * (size_t)nblk * CB * CB, &Tdiag));
    for (int kb = 0; kb < nblk; ++kb) {
        k_chol_diag<<<1, CB * CB, 0, ctx->stream>>>(G, ld, m, kb, (double*)Tdiag->ptr, status_dev);
        ctx->launches++;
        const int nrem = nblk - kb - 1;
        if (nrem > 0) {
            k_chol_row<<<nrem, CB * CB, 0, ctx->stream>>>(G, ld, m, kb, (const double*)Tdiag->ptr);
            k_chol_trail<<<nrem * (nrem + 1) / 2, CB * CB, 0, ctx->stream>>>(G, ld, m, kb, nrem);
            ctx->launches += 2;
        }
    }
    cudaMemsetAsync(T, 0, sizeof(double) * (size_t)m * ld, ctx->stream);
    k_upper_inverse_blocked<<<nblk, CB * CB, 0, ctx->stream>>>(G, m, ld, (const double*)Tdiag->ptr, T);
    ctx->launches++;
    gl_buf_release(Tdiag);
    if (cudaGetLastError() != cudaSuccess) { gl_set_error("Cholesky: kernel launch failed"); return GL_ERR_CUDA; }
    return GL_OK;
}

int gl_impl_orthonormalise(gl_ctx* ctx, gl_mat* phi, double* norms_out)
{
    const int m = phi->m, m_pad = phi->m_pad;
    const int64_t rows = phi->local_rows;
    GL_REQUIRE(m_pad % 64 == 0, "orthonormalise: m_pad %d", m_pad);
    const int tiles = m_pad / 64;
    int slabs = (int)((4 * (int64_t)ctx->sm_count) / ((int64_t)tiles * (tiles + 1) / 2) + 1);
    if (slabs > 64) slabs = 64;
    if ((int64_t)slabs * 32 > rows) slabs = (int)ceil_div(rows, 32);
    if (slabs < 1) slabs = 1;
    // tensor-core Gram for wide Phi: K slabs of at most 1024 blocks of 64 rows (fp32 chains of <= 65 536 terms)
    const bool tc_gram = ctx->gram_impl == 0 && m_pad % 256 == 0;
    const int kb_per_item = 1024;
    if (tc_gram) slabs = (int)ceil_div(ceil_div(rows, 64), kb_per_item);

    gl_buf *partial = nullptr, *G = nullptr, *T = nullptr, *Et = nullptr, *Q = nullptr, *st = nullptr, *norms = nullptr, *sc = nullptr;
    int rc = GL_OK;
    do {
        if ((rc = gl_alloc(ctx, sizeof(float) * (size_t)slabs * m_pad * m_pad, &partial)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(double) * (size_t)m_pad * m_pad, &G)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(double) * (size_t)m_pad * m_pad, &T)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(__half) * (size_t)m_pad * m_pad, &Et)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(__half) * (size_t)rows * m_pad, &Q)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(int) * 4, &st)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(double) * (size_t)m_pad, &norms)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(float) * 4, &sc)) != GL_OK) break;
        GL_CUDA_BREAK(rc, cudaMemsetAsync(st->ptr, 0, sizeof(int) * 4, ctx->stream));

        if (tc_gram) {
            CUtensorMap map_phi;
            if ((rc = make_map_2d(&map_phi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, phi->buf->ptr, (uint64_t)rows, (uint64_t)m_pad, (uint64_t)m_pad, 64,
                                  64)) != GL_OK) break;
            gl_buf* gerr = nullptr;
            if ((rc = gl_alloc(ctx, sizeof(int) * 4, &gerr)) != GL_OK) break;
            cudaMemsetAsync(gerr->ptr, 0, sizeof(int) * 4, ctx->stream);
            GL_CUDA_BREAK(rc, cudaFuncSetAttribute(gram::k_gram_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, gram::SMEM_BYTES));
            int ntiles = 0;
            for (int ti = 0; ti < m_pad / 128; ++ti) ntiles += m_pad / 256 - ti / 2;
            int grid = ctx->sm_count;
            if (grid > slabs * ntiles) grid = slabs * ntiles;
            gram::k_gram_tcgen05<<<grid, gram::THREADS, gram::SMEM_BYTES, ctx->stream>>>(map_phi, rows, m_pad, kb_per_item, slabs,
                                                                                         (float*)partial->ptr, (uint32_t)ctx->gram_lbo,
                                                                                         (int*)gerr->ptr);
            gl_buf_release(gerr);
        } else {
            dim3 gg((unsigned)tiles, (unsigned)tiles, (unsigned)slabs);
            k_gram_tile<<<gg, 256, 0, ctx->stream>>>((const __half*)phi->buf->ptr, rows, m_pad, slabs, (float*)partial->ptr);
        }
        GL_LAUNCH_CHECK(ctx);
        dim3 gr((unsigned)ceil_div(m_pad, 128), (unsigned)m_pad);
        k_gram_reduce<<<gr, 128, 0, ctx->stream>>>((const float*)partial->ptr, slabs, m_pad, (double*)G->ptr);
        GL_LAUNCH_CHECK(ctx);
        if ((rc = gl_allreduce_f64(ctx, (double*)G->ptr, (size_t)m_pad * m_pad)) != GL_OK) break;

        if ((rc = gl_chol_inverse_upper(ctx, (double*)G->ptr, m, m_pad, (double*)T->ptr, (int*)st->ptr)) != GL_OK) break;
        dim3 ge((unsigned)ceil_div(m_pad, 128), (unsigned)m_pad);
        k_build_et<<<ge, 128, 0, ctx->stream>>>((const double*)T->ptr, m, m_pad, m_pad, (__half*)Et->ptr,
                                                (const double*)G->ptr, (double*)norms->ptr);
        GL_LAUNCH_CHECK(ctx);
        k_et_scales<<<1, 1, 0, ctx->stream>>>((float*)sc->ptr);
        GL_LAUNCH_CHECK(ctx);
        if ((rc = gl_gemm_kmajor(ctx, phi->buf->ptr, 0, rows, m_pad, Et->ptr, m_pad, (const float*)sc->ptr, phi->buf->ptr,
                                 Q->ptr)) != GL_OK) break;

        GL_BREAK(rc, gl_ensure_pinned(ctx, sizeof(double) * (size_t)m_pad + 64));
        GL_CUDA_BREAK(rc, cudaMemcpyAsync(ctx->pinned, st->ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        GL_CUDA_BREAK(rc, cudaMemcpyAsync((char*)ctx->pinned + 64, norms->ptr, sizeof(double) * (size_t)m, cudaMemcpyDeviceToHost,
                                      ctx->stream));
        GL_CUDA_BREAK(rc, cudaStreamSynchronize(ctx->stream));
        const int status = *(int*)ctx->pinned;
        if (status != 0) {
            gl_set_error("orthonormalise: Gram matrix not positive definite at column %d (Phi rank deficient)", status - 1);
            rc = GL_ERR_NOTCONVERGED;
            break;
        }
        if (norms_out) memcpy(norms_out, (char*)ctx->pinned + 64, sizeof(double) * (size_t)m);
        if (phi->proj) {
            gl_buf* np_ = nullptr;
            if ((rc = gl_alloc(ctx, sizeof(double) * (size_t)m_pad * phi->channels, &np_)) != GL_OK) break;
            k_proj_times_t<<<(unsigned)ceil_div(m_pad, 128), 128, 0, ctx->stream>>>((const double*)T->ptr, m, m_pad, m_pad, phi->channels,
                                                                                    (const double*)phi->proj->ptr, (double*)np_->ptr);
            ctx->launches++;
            gl_buf_release(phi->proj);
            phi->proj = np_;
        }
        // Phi <- Q (swap storage; the handle keeps its identity)
        gl_buf* old = phi->buf;
        phi->buf = Q;
        Q = old;
    } while (0);
    if (partial) gl_buf_release(partial);
    if (G) gl_buf_release(G);
    if (T) gl_buf_release(T);
    if (Et) gl_buf_release(Et);
    if (Q) gl_buf_release(Q);
    if (st) gl_buf_release(st);
    if (norms) gl_buf_release(norms);
    if (sc) gl_buf_release(sc);
    return rc;
}
