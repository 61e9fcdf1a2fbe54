// a-4 / a-5: the p x p symmetric eigensolve as a hand-written device block-Jacobi kernel.
// Replaces EigendecompositionSmallest (SLEPc Krylov-Schur, hpc/eigendecomposition.c:12-124) and
// InversePowerIteration (hpc/inverse_power_it.c:86-252): both ask for the m eigenpairs of L_A nearest
// zero; this solver returns the CONVERGED pairs, ascending.
//
// Method: one-sided (Hestenes) block Jacobi on G = L_A.  Right rotations G <- G J make the columns of G
// mutually orthogonal; since L_A is symmetric positive definite, at convergence G = U diag(lambda), so
// lambda_j = |g_j| and u_j = g_j / |g_j| -- no separate eigenvector accumulation.
//   * columns are grouped in panels of 8; G is stored panel-major ([panel][row][8] fp32) so a pair of
//     panels is two contiguous streams;
//   * one CTA owns a panel pair per step: loads it to shared memory (p x 16), forms the 16 x 16 Gram
//     block, diagonalises it with a two-sided cyclic Jacobi in shared memory, applies the 16 x 16
//     rotation to the panel and streams it back;
//   * a round-robin tournament gives (panels - 1) steps per sweep with panels/2 independent pairs per
//     step; steps are separated by a cooperative grid barrier, the whole solve is ONE kernel launch;
//   * no atomics on floating-point data and fixed reduction orders: the result is bit-reproducible, so
//     every GPU of a multi-GPU run can solve redundantly and hold identical eigenvectors.
#include <cooperative_groups.h>

#include <algorithm>

#include "common.cuh"

namespace cg = cooperative_groups;

#define JB 8           // columns per panel
#define JP (2 * JB)    // columns per panel pair
#define J_THREADS 256
#define J_INNER_MAX 12

// Row-chunk occupancy of the panels.  L_A of this pipeline is banded (couplings between spatially close samples only; in raster order
// of the samples that is a band of ~15 % of the rows), and a Jacobi rotation of two panels only mixes their rows: a panel's rows that are
// zero in both stay zero.  Every panel carries a 32-bit mask of the row chunks (cr = ceil(p / 32) rows each) that may hold a non-zero;
// the Gram screen, the pair Gram blocks and the rotations skip the rest.  Entries below 1e-20 (the diagonal is ~1) are flushed to zero
// when G is set up, so that "zero" is exact.
__device__ __forceinline__ int jacobi_chunk_rows(int p) { return (p + 31) / 32; }

__global__ void k_jacobi_init(const double* __restrict__ A, int p, int nb, float* __restrict__ G, unsigned* __restrict__ pmask)
{
    // G[panel][r][c] = A[r][panel*8 + c]; zero columns beyond p
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)nb * p * JB;
    if (idx >= total) return;
    int c = (int)(idx % JB);
    int64_t t = idx / JB;
    int r = (int)(t % p);
    int panel = (int)(t / p);
    int col = panel * JB + c;
    float v = col < p ? (float)A[(size_t)r * p + col] : 0.f;
    if (fabsf(v) < 1e-20f) v = 0.f;
    G[idx] = v;
    // one atomic per warp and (panel, chunk) instead of one per non-zero element: a warp covers 4 consecutive rows of one panel
    const unsigned bit = v != 0.f ? 1u << (r / jacobi_chunk_rows(p)) : 0u;
    const unsigned peers = __match_any_sync(__activemask(), panel);
    const unsigned bits = __reduce_or_sync(peers, bit);
    if (bits && (threadIdx.x & 31) == __ffs(peers) - 1) atomicOr(&pmask[panel], bits);
}

// round-robin tournament over nb (even) panels: step in [0, nb-1), pair in [0, nb/2)
__device__ __forceinline__ void tournament(int step, int pair, int nb, int& a, int& b)
{
    const int mth = nb - 1;
    if (pair == 0) {
        a = mth;
        b = step % mth;
    } else {
        a = (step + pair) % mth;
        b = (step - pair + mth) % mth;
    }
    if (a > b) { int t = a; a = b; b = t; }
}

// One panel pair (I, J): load the two panels, form the 16 x 16 Gram block, diagonalise it, rotate the panels.
// `pc` = rows of the pair that fit in shared memory at once: the whole panels when p <= pc (loaded once, rotated in place),
// otherwise the Gram block is accumulated chunk by chunk and the chunks are fetched a second time for the rotation.
__device__ __forceinline__ void jacobi_pair(float* __restrict__ G, int p, int pc, int I, int J, float tol, int inner_max, float* P, float* red,
                                            double* Bm, double* Qm, double* cs, int* role, int* pq, float* s_off, unsigned* s_cta_off,
                                            unsigned* __restrict__ pmask, bool allow_cross, long long* __restrict__ pprof = nullptr)
{
    const int tid = threadIdx.x, lane = tid & 31;
    const int bi = tid >> 4, bj = tid & 15;  // this thread's element of the 16 x 16 block
#define PPROF(k) do { if (pprof && tid == 0) pprof[k] = clock64(); } while (0)
    PPROF(0);
    float* GI = G + (size_t)I * p * JB;
    float* GJ = G + (size_t)J * p * JB;
    const int nchunks = (p + pc - 1) / pc;
    // Only the row chunks in which one of the two panels may be non-zero take part (single-pass case: the whole pair in shared
    // memory); they are held compactly: compact row k of the pair = row (k % cr) of the (k / cr)-th occupied chunk.
    const int cr = jacobi_chunk_rows(p);
    const unsigned occ = nchunks == 1 ? (__ldcg(&pmask[I]) | __ldcg(&pmask[J])) : 0xffffffffu;
    const int n_occ = nchunks == 1 ? __popc(occ) * cr : p;
    // (the k-th occupied chunk comes from a small table: __fns is a software loop, and sat in every row of the loads and the apply)
    __shared__ int s_list[32];
    if (nchunks == 1 && tid < 32 && ((occ >> tid) & 1u)) s_list[__popc(occ & ((1u << tid) - 1u))] = tid;
    __syncthreads();
    auto actual_row = [&](int k) -> int {
        if (nchunks != 1) return k;
        const int ch = k / cr;
        return s_list[ch] * cr + (k - ch * cr);
    };
    // rows [r0, r0 + n) of the panel pair -> P (L2 loads: the data was written by other SMs)
    auto load_chunk = [&](int r0, int n) {
        for (int idx = tid; idx < 2 * n; idx += J_THREADS) {
            const int r = idx >> 1, h = idx & 1;
            const int ra = actual_row(r0 + r);
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            if (ra < p) {
                a = __ldcg((const float4*)(GI + (size_t)ra * JB + h * 4));
                b = __ldcg((const float4*)(GJ + (size_t)ra * JB + h * 4));
            }
            *(float4*)(P + (size_t)r * JP + h * 4) = a;
            *(float4*)(P + (size_t)r * JP + JB + h * 4) = b;
        }
    };
    __shared__ unsigned s_occ[2];
    if (tid == 0) { *s_cta_off = 0u; s_occ[0] = 0u; s_occ[1] = 0u; }
    // ---- Gram block: 16 row groups x (4 x 4 register tiles), accumulated over the chunks ----
    {
        const int rg = tid >> 4, ti = (tid >> 2) & 3, tj = tid & 3;
        float acc[4][4];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[x][y] = 0.f;
        for (int c = 0; c < nchunks; ++c) {
            const int r0 = c * pc, n = nchunks == 1 ? n_occ : min(pc, p - r0);
            if (c > 0) __syncthreads();   // the previous chunk's readers are done
            load_chunk(r0, n);
            __syncthreads();
            for (int r = rg; r < n; r += 16) {
                const float4 a = *(const float4*)(P + (size_t)r * JP + 4 * ti);
                const float4 b = *(const float4*)(P + (size_t)r * JP + 4 * tj);
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(av[x], bv[y], acc[x][y]);
            }
        }
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) red[rg * 256 + (4 * ti + x) * 16 + 4 * tj + y] = acc[x][y];
    }
    __syncthreads();
    PPROF(1);
    {
        double s = 0.0;
#pragma unroll
        for (int rg = 0; rg < 16; ++rg) s += (double)red[rg * 256 + tid];
        Bm[bi * 17 + bj] = s;
        Qm[bi * 17 + bj] = (bi == bj) ? 1.0 : 0.0;
    }
    __syncthreads();
    // ---- how far from orthogonal is this pair now (rotations made earlier in the sweep may have changed it) ----
    {
        float rel = 0.f;
        if (bi < bj) {
            const double dii = Bm[bi * 17 + bi], djj = Bm[bj * 17 + bj];
            if (dii > 0.0 && djj > 0.0) rel = (float)(fabs(Bm[bi * 17 + bj]) / sqrt(dii * djj));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rel = fmaxf(rel, __shfl_xor_sync(0xffffffffu, rel, o));
        if (lane == 0) atomicMax(s_cta_off, __float_as_uint(rel));
    }
    __syncthreads();
    const float pair_off = __uint_as_float(*s_cta_off);
    // ---- diagonalise the 16 x 16 block: cyclic two-sided Jacobi, Q accumulates the rotations ----
    PPROF(2);
    if (pair_off > 0.25f * tol) {
        // One warp diagonalises the block with warp-level barriers only: with all 256 threads on it, every one of the 15 steps of a
        // sweep cost three CTA barriers around a few dozen fp64 operations per thread (1 060 cycles per step, 60 % of a pair's time).
        // Lane (k, sub) = (lane >> 2, lane & 3) works for pair k of the step: it computes the pair's rotation itself (the four lanes of
        // a pair redundantly, from the same three entries), then B <- B J and Q <- Q J on rows 4 sub .. 4 sub + 3 of the pair's two
        // columns, then B <- J^T B on columns 4 sub .. 4 sub + 3 of the pair's two rows -- in place: within a half step no two lanes
        // touch the same entry.
        if (tid < 32) {
            const int k = lane >> 2, sub = lane & 3;
            // A panel's own 8 x 8 blocks are diagonalised when its HOME pair (2k, 2k+1) is visited; every other visit rotates only
            // the 64 cross pairs (8 steps of 8 disjoint pairs (k, 8 + (k + st) mod 8) instead of the 15 of the full cyclic sweep).
            // What a cross rotation spills into the panels' own blocks is second order and is swept up by the next home visit:
            // same number of outer sweeps on every fixture and config 4 (tools/emulate_jacobi_cross.py), 47 % less of the
            // serial fp64 rotation chain that bounds a pair.
            // Near convergence (allow_cross = false: the sweep started below 4 tol) every visit runs the full sweep again: at p = 4 200
            // the cross-only sweeps stalled a hair above the tolerance (5.1e-5), fed by what they leave in the panels' own blocks.
            const bool cross_only = allow_cross && !(J == I + 1 && (I & 1) == 0);
            const int nst = cross_only ? JB : JP - 1;
            for (int isw = 0; isw < inner_max; ++isw) {
                float sw_off = 0.f;
                for (int st = 0; st < nst; ++st) {
                    int a, b;
                    if (cross_only) { a = k; b = JB + ((k + st) & (JB - 1)); }
                    else tournament(st, k, JP, a, b);
                    const double app = Bm[a * 17 + a], aqq = Bm[b * 17 + b], apq = Bm[a * 17 + b];
                    double c = 1.0, sn = 0.0;
                    if (app > 0.0 && aqq > 0.0) {
                        // The rotation ANGLE only has to be about right (fp32: fast rsqrt / divide instead of chains of fp64 sqrt and
                        // divide); the rotation itself must be orthogonal to fp64 accuracy, so c = (1 + t^2)^-1/2 is polished with two
                        // Newton steps in fp64 and s = t c.
                        const float rel = fabsf((float)apq) * rsqrtf((float)app * (float)aqq);
                        if (rel > 1e-12f) {
                            const float tau = __fdividef((float)(aqq - app), 2.f * (float)apq);
                            const float tf = copysignf(1.f, tau) / (fabsf(tau) + sqrtf(fmaf(tau, tau, 1.f)));
                            const double t = isfinite(tf) ? (double)tf : 0.0;
                            const double x = fma(t, t, 1.0);
                            double y = (double)rsqrtf((float)x);
                            y = y * fma(-0.5 * x, y * y, 1.5);
                            y = y * fma(-0.5 * x, y * y, 1.5);
                            c = y;
                            sn = t * y;
                        }
                        sw_off = fmaxf(sw_off, rel);
                    }
                    __syncwarp();      // every lane has read its pair's entries before the columns are overwritten
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        const int r = 4 * sub + rr;
                        const double ba = Bm[r * 17 + a], bb = Bm[r * 17 + b], qa = Qm[r * 17 + a], qb = Qm[r * 17 + b];
                        Bm[r * 17 + a] = c * ba - sn * bb;
                        Bm[r * 17 + b] = sn * ba + c * bb;
                        Qm[r * 17 + a] = c * qa - sn * qb;
                        Qm[r * 17 + b] = sn * qa + c * qb;
                    }
                    __syncwarp();
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int j = 4 * sub + jj;
                        const double xa = Bm[a * 17 + j], xb = Bm[b * 17 + j];
                        Bm[a * 17 + j] = c * xa - sn * xb;
                        Bm[b * 17 + j] = sn * xa + c * xb;
                    }
                    __syncwarp();
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sw_off = fmaxf(sw_off, __shfl_xor_sync(0xffffffffu, sw_off, o));
                if (sw_off <= 0.1f * tol) break;
            }
        }
        __syncthreads();
        PPROF(3);
        // ---- apply: [G_I G_J] <- [G_I G_J] Q, streamed straight back to global ----
        const int cgp = tid & 3;
        float q[JP][4];
#pragma unroll
        for (int k = 0; k < JP; ++k)
#pragma unroll
            for (int x = 0; x < 4; ++x) q[k][x] = (float)Qm[k * 17 + 4 * cgp + x];
        float* dst = (cgp < 2) ? (GI + cgp * 4) : (GJ + (cgp - 2) * 4);
        for (int c = 0; c < nchunks; ++c) {
            const int r0 = c * pc, n = nchunks == 1 ? n_occ : min(pc, p - r0);
            if (nchunks > 1) {            // the single chunk is still in P from the Gram pass
                __syncthreads();
                load_chunk(r0, n);
                __syncthreads();
            }
            for (int r = tid >> 2; r < n; r += J_THREADS / 4) {
                const int ra = actual_row(r0 + r);
                if (ra >= p) continue;
                const float4 v0 = *(const float4*)(P + (size_t)r * JP);
                const float4 v1 = *(const float4*)(P + (size_t)r * JP + 4);
                const float4 v2 = *(const float4*)(P + (size_t)r * JP + 8);
                const float4 v3 = *(const float4*)(P + (size_t)r * JP + 12);
                const float pv[JP] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w,
                                      v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
                float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < JP; ++k)
#pragma unroll
                    for (int x = 0; x < 4; ++x) o[x] = fmaf(pv[k], q[k][x], o[x]);
                // what a small-angle rotation spills into rows that were zero is second order: flushed below 1e-12 (the tolerance
                // of the solve is 5e-5), so that the occupancy masks keep following the band instead of filling up within a sweep
                bool nz = false;
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    if (fabsf(o[x]) < 1e-12f) o[x] = 0.f;
                    nz = nz || o[x] != 0.f;
                }
                __stcg((float4*)(dst + (size_t)ra * JB), make_float4(o[0], o[1], o[2], o[3]));
                if (nz && nchunks == 1) atomicOr(cgp < 2 ? &s_occ[0] : &s_occ[1], 1u << (ra / cr));
            }
        }
        if (nchunks == 1) {
            __syncthreads();
            if (tid == 0) { pmask[I] = s_occ[0]; pmask[J] = s_occ[1]; }
        }
    }
    __syncthreads();
    PPROF(4);
#undef PPROF
}

// The whole solve in ONE cooperative launch.  Per sweep:
//   A. screen: C = G^T G (fp32, 64 x 64 tiles over all CTAs);
//   B. per panel pair the largest relative off-diagonal of its 8 x 8 block of C (the pair (2k, 2k+1) also carries the
//      within-panel blocks of its two panels); global maximum = the convergence measure; pairs above 0.2 tol are ACTIVE;
//   C. rotations over the active pairs only.  L_A of this pipeline is nearly diagonal with couplings between spatially
//      close samples only (at 4K / p = 1000: 600 of 7875 panel pairs, all with |I - J| <= 8), so the pairs are visited
//      by distance class: step (d, par) = all pairs (I, I + d) with floor(I / d) of parity par -- mutually disjoint,
//      and steps without an active pair cost nothing (no barrier).  When more than half of the pairs are active
//      (photometric affinity) the round-robin tournament is used instead: (panels - 1) steps of panels / 2 pairs.
__global__ void __launch_bounds__(J_THREADS, 1)
k_jacobi(float* __restrict__ G, int p, int pc /* panel rows held in shared memory */, int nb, int max_sweeps, float tol, int inner_max,
         float* __restrict__ C /* [cp][cp], cp = nb * 8 */,
         float* __restrict__ pair_rel /* [nb][nb] */, int* __restrict__ step_cnt /* [max_sweeps][2 * nb] */,
         unsigned* __restrict__ sweep_off, int* __restrict__ sweeps_done, long long* __restrict__ jprof /* debug: phase clocks of CTA 0, or null */,
         unsigned* __restrict__ pmask /* [nb] occupied row chunks of every panel */)
{
    cg::grid_group grid = cg::this_grid();
    int jp = 0;
#define JPROF() do { if (jprof && blockIdx.x == 0 && threadIdx.x == 0 && jp < 60) jprof[jp++] = clock64(); } while (0)
    JPROF();
    extern __shared__ __align__(16) float jsm[];
    float* P = jsm;                          // [pc][16]
    float* red = P + (size_t)pc * JP;        // [16][256]
    double* Bm = (double*)(red + 16 * 256);  // [16][17]  Gram block, fp64 from here on: the rotations that are
    double* Qm = Bm + JP * 17;               // [16][17]  accumulated into Q must stay orthogonal to ~1e-16, or the
    double* cs = Qm + JP * 17;               // [8][2]    column norms (= eigenvalues) drift by ~1e-6 per update
    int* role = (int*)(cs + 16);             // [16]  (pair << 1) | is_second
    int* pq = role + JP;                     // [8][2]
    __shared__ float s_off;
    __shared__ unsigned s_cta_off;
    __shared__ short s_cls[2 * 1024];        // non-empty (distance, parity) classes of the sweep: nb <= 1024 panels (p <= 8192)
    __shared__ int s_ncls, s_wcnt[J_THREADS / 32];

    const int tid = threadIdx.x;
    const int cp = nb * JB;
    const int T = (cp + 63) >> 6;            // 64-column tiles of C
    const int ntile = T * (T + 1) / 2;
    const int npairs_all = nb * (nb - 1) / 2;

    int sweep = 0;
    float prev_off = 3.0e38f;
    for (; sweep < max_sweeps; ++sweep) {
        int* cnt = step_cnt + (size_t)sweep * 2 * nb;   // [0] = active pairs in total, [2 d + par] = per step
        // ---- A. C = G^T G, upper tiles ----
        {
            float* As = jsm;                 // [32][68]
            float* Bs = jsm + 32 * 68;       // [32][68]
            const int tx = tid & 15, ty = tid >> 4;
            for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
                int ti = 0, rem = tile;
                while (rem >= T - ti) { rem -= T - ti; ++ti; }
                const int tj = ti + rem;
                float acc[4][4];
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[x][y] = 0.f;
                const int lp = tid >> 5, lr = tid & 31;          // panel within the tile, row within the chunk
                const int pa = ti * 8 + lp, pb = tj * 8 + lp;
                // rows of the two panels this thread stages; the NEXT chunk's loads are in flight while the current one is multiplied
                // (an L2 round trip per 32-row chunk, unhidden, made this screen 3x slower than its arithmetic)
                auto fetch = [&](int r0, float4& a0, float4& a1, float4& b0, float4& b1) {
                    const int r = r0 + lr;
                    a0 = make_float4(0.f, 0.f, 0.f, 0.f); a1 = a0; b0 = a0; b1 = a0;
                    if (r < p) {
                        if (pa < nb) {
                            a0 = __ldcg((const float4*)(G + ((size_t)pa * p + r) * JB));
                            a1 = __ldcg((const float4*)(G + ((size_t)pa * p + r) * JB + 4));
                        }
                        if (pb < nb) {
                            b0 = __ldcg((const float4*)(G + ((size_t)pb * p + r) * JB));
                            b1 = __ldcg((const float4*)(G + ((size_t)pb * p + r) * JB + 4));
                        }
                    }
                };
                // row chunks in which BOTH tiles may be non-zero (the others contribute nothing to C)
                unsigned mi = 0u, mj = 0u;
                for (int q = 0; q < 8; ++q) {
                    if (ti * 8 + q < nb) mi |= __ldcg(&pmask[ti * 8 + q]);
                    if (tj * 8 + q < nb) mj |= __ldcg(&pmask[tj * 8 + q]);
                }
                const unsigned both = mi & mj;
                const int cr = jacobi_chunk_rows(p);
                const int nsub = (cr + 31) / 32;                 // 32-row staging steps per row chunk
                const int steps = __popc(both) * nsub;
                // cursor of the step being FETCHED (one ahead of the step being multiplied): lowest remaining chunk, sub-step in it
                unsigned rest = both;
                int sub = 0, nrow_next = 0;
                float4 a0, a1, b0, b1;
                auto fetch_next = [&]() {
                    nrow_next = min(32, cr - sub * 32);
                    fetch((__ffs(rest) - 1) * cr + sub * 32, a0, a1, b0, b1);
                    if (++sub == nsub) { sub = 0; rest &= rest - 1u; }
                };
                if (steps > 0) fetch_next();
                for (int st = 0; st < steps; ++st) {
                    const int nrow = nrow_next;
                    __syncthreads();
                    const bool live = lr < nrow;
                    *(float4*)&As[lr * 68 + lp * 8] = live ? a0 : make_float4(0.f, 0.f, 0.f, 0.f);
                    *(float4*)&As[lr * 68 + lp * 8 + 4] = live ? a1 : make_float4(0.f, 0.f, 0.f, 0.f);
                    *(float4*)&Bs[lr * 68 + lp * 8] = live ? b0 : make_float4(0.f, 0.f, 0.f, 0.f);
                    *(float4*)&Bs[lr * 68 + lp * 8 + 4] = live ? b1 : make_float4(0.f, 0.f, 0.f, 0.f);
                    __syncthreads();
                    if (st + 1 < steps) fetch_next();
#pragma unroll 8
                    for (int k = 0; k < 32; ++k) {
                        const float4 a = *(const float4*)&As[k * 68 + ty * 4];
                        const float4 b = *(const float4*)&Bs[k * 68 + tx * 4];
                        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                        for (int x = 0; x < 4; ++x)
#pragma unroll
                            for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(av[x], bv[y], acc[x][y]);
                    }
                }
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const int row = ti * 64 + ty * 4 + x, col = tj * 64 + tx * 4;
                    if (row < cp && col < cp) __stcg((float4*)(C + (size_t)row * cp + col), make_float4(acc[x][0], acc[x][1], acc[x][2], acc[x][3]));
                }
                __syncthreads();
            }
        }
        grid.sync();
        JPROF();
        // ---- B. panel-pair screen ----
        {
            float cta_max = 0.f;
            for (int idx = blockIdx.x * J_THREADS + tid; idx < nb * nb; idx += gridDim.x * J_THREADS) {
                const int I = idx / nb, J = idx - I * nb;
                if (J <= I) continue;
                float rel = 0.f;
                float di[JB], dj[JB];
#pragma unroll
                for (int a = 0; a < JB; ++a) {
                    di[a] = __ldcg(C + (size_t)(I * JB + a) * cp + I * JB + a);
                    dj[a] = __ldcg(C + (size_t)(J * JB + a) * cp + J * JB + a);
                }
#pragma unroll
                for (int a = 0; a < JB; ++a)
#pragma unroll
                    for (int b = 0; b < JB; ++b) {
                        const float c = fabsf(__ldcg(C + (size_t)(I * JB + a) * cp + J * JB + b));
                        const float d = di[a] * dj[b];
                        if (d > 0.f) rel = fmaxf(rel, c * rsqrtf(d));
                    }
                if (J == (I ^ 1)) {  // this pair also answers for the blocks inside its two panels
#pragma unroll
                    for (int a = 0; a < JB; ++a)
#pragma unroll
                        for (int b = a + 1; b < JB; ++b) {
                            const float c1 = fabsf(__ldcg(C + (size_t)(I * JB + a) * cp + I * JB + b)), d1 = di[a] * di[b];
                            const float c2 = fabsf(__ldcg(C + (size_t)(J * JB + a) * cp + J * JB + b)), d2 = dj[a] * dj[b];
                            if (d1 > 0.f) rel = fmaxf(rel, c1 * rsqrtf(d1));
                            if (d2 > 0.f) rel = fmaxf(rel, c2 * rsqrtf(d2));
                        }
                }
                pair_rel[idx] = rel;
                cta_max = fmaxf(cta_max, rel);
                if (rel > 0.2f * tol) {
                    const int d = J - I, par = (I / d) & 1;
                    atomicAdd(&cnt[2 * d + par], 1);
                    atomicAdd(&cnt[0], 1);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cta_max = fmaxf(cta_max, __shfl_xor_sync(0xffffffffu, cta_max, o));
            if ((tid & 31) == 0 && cta_max > 0.f) atomicMax(&sweep_off[sweep], __float_as_uint(cta_max));
        }
        grid.sync();
        JPROF();
        const float off = __uint_as_float(__ldcg(&sweep_off[sweep]));
        if (off <= tol) break;
        // The fp32 floor: a column that belongs to a small eigenvalue carries direction noise of about 6e-8 lambda_max / lambda, and
        // when that sits above the tolerance the sweeps stop improving (350 x 300 pixels, p = 4 200: 5.1e-5 sweep after sweep against
        // tol = 5e-5).  Jacobi converges quadratically at the end, so two screens in a row within 30 % of each other are the floor:
        // accepted up to 4 tol (still below half of fp16's rounding of Phi) and reported as such (sweeps_done[1]).
        if (sweep >= 2 && off <= 4.f * tol && off > 0.7f * prev_off) {
            if (blockIdx.x == 0 && tid == 0) sweeps_done[1] = 1;
            break;
        }
        prev_off = off;
        // ---- C. rotations ----
        const int active = __ldcg(&cnt[0]);
        if (2 * active > npairs_all) {
            for (int step = 0; step < nb - 1; ++step) {
                for (int pair = blockIdx.x; pair < (nb >> 1); pair += gridDim.x) {
                    int I, J;
                    tournament(step, pair, nb, I, J);
                    if (__ldcg(&pair_rel[I * nb + J]) > 0.2f * tol)
                        jacobi_pair(G, p, pc, I, J, tol, inner_max, P, red, Bm, Qm, cs, role, pq, &s_off, &s_cta_off, pmask, off > 4.f * tol);
                }
                grid.sync();
            }
        } else {
            // the non-empty classes, gathered once (polling cnt[] class by class cost an L2 round trip for each of the 2 (nb - 1) classes,
            // ~110 000 cycles per sweep at p = 1000, most of them empty); every CTA builds the same list in the same order
            __syncthreads();
            if (tid == 0) s_ncls = 0;
            __syncthreads();
            for (int base = 2; base < 2 * nb; base += J_THREADS) {
                const int k = base + tid;
                const bool live = k < 2 * nb && __ldcg(&cnt[k]) != 0;
                const unsigned bal = __ballot_sync(0xffffffffu, live);
                if ((tid & 31) == 0) s_wcnt[tid >> 5] = __popc(bal);
                __syncthreads();
                int before = s_ncls;
                for (int w = 0; w < (tid >> 5); ++w) before += s_wcnt[w];
                if (live) s_cls[before + __popc(bal & ((1u << (tid & 31)) - 1u))] = (short)k;
                __syncthreads();
                if (tid == 0) { int t = 0; for (int w = 0; w < J_THREADS / 32; ++w) t += s_wcnt[w]; s_ncls += t; }
                __syncthreads();
            }
            const int ncls = s_ncls;
            for (int ci = 0; ci < ncls; ++ci) {
                {
                    const int d = s_cls[ci] >> 1, par = s_cls[ci] & 1;
                    const int ncand = ((nb + 2 * d - 1) / (2 * d)) * d;
                    for (int c = blockIdx.x; c < ncand; c += gridDim.x) {
                        const int I = (c / d) * 2 * d + par * d + (c % d), J = I + d;
                        if (J < nb && __ldcg(&pair_rel[I * nb + J]) > 0.2f * tol)
                            jacobi_pair(G, p, pc, I, J, tol, inner_max, P, red, Bm, Qm, cs, role, pq, &s_off, &s_cta_off, pmask, off > 4.f * tol,
                                        (jprof && blockIdx.x == 0) ? jprof + 40 : nullptr);
                    }
                    grid.sync();
                    JPROF();
                }
            }
        }
    }
    JPROF();
    if (jprof && blockIdx.x == 0 && tid == 0) jprof[63] = jp;
    if (blockIdx.x == 0 && tid == 0) *sweeps_done = sweep;
}

// lambda_c = |g_c| (fp64 accumulation), one warp per original column
__global__ void k_jacobi_norms(const float* __restrict__ G, int p, double* __restrict__ lam)
{
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= p) return;
    const float* col = G + (size_t)(c / JB) * p * JB + (c % JB);
    double acc = 0.0;
    for (int r = threadIdx.x & 31; r < p; r += 32) {
        double v = (double)col[(size_t)r * JB];
        acc += v * v;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) lam[c] = sqrt(acc);
}

// Rayleigh quotients in fp64 against the ORIGINAL matrix: mu_c = (g_c^T A g_c) / (g_c^T g_c).  The fp32 sweeps leave
// eigenvector errors of ~1e-6..1e-5; the quotient's error is their square, so eigenvalues come out at fp64-level
// accuracy whatever the conditioning of A.  CTA = 64 rows x 64 columns of V = A G, 256 threads x (4 x 4), K loop of 16;
// V is never stored: each CTA emits partial numerators for its 64 columns (fixed-order reduction afterwards).
__global__ void __launch_bounds__(256) k_rayleigh_partial(const double* __restrict__ A, const float* __restrict__ G, int p,
                                                          double* __restrict__ part /* [row_tiles][cols_pad] */, int cols_pad,
                                                          int cols /* columns that exist in G */,
                                                          const unsigned char* __restrict__ nz /* [row_tiles][ceil(p/16)] or null */,
                                                          const unsigned* __restrict__ pmask /* occupied row chunks of G's panels, or null */,
                                                          int nb)
{
    __shared__ double As[16][64 + 1], Us[16][64 + 1];
    __shared__ double red[16][64];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int j0 = blockIdx.x * 64, i0 = blockIdx.y * 64;
    // row chunks in which any of this tile's 64 columns of G may be non-zero: rows outside contribute neither to V = A G (as the k
    // index) nor to the numerator sum_i G[i][c] V[i][c] (as the row index) -- a tile whose 64 rows miss them all has nothing to add
    unsigned gm = 0xffffffffu;
    const int cr = jacobi_chunk_rows(p);
    if (pmask) {
        gm = 0u;
        for (int q = 0; q < 8; ++q) if (blockIdx.x * 8 + q < nb) gm |= pmask[blockIdx.x * 8 + q];
        const int ca = i0 / cr, cb = min(i0 + 63, p - 1) / cr;
        unsigned rows_mask = 0u;
        for (int c = ca; c <= cb; ++c) rows_mask |= 1u << c;
        if (!(gm & rows_mask)) {
            if (threadIdx.x < 64) part[(size_t)blockIdx.y * cols_pad + j0 + threadIdx.x] = 0.0;
            return;
        }
    }
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    const int nz_ld = (p + 15) >> 4;
    for (int k0 = 0; k0 < p; k0 += 16) {
        if (nz && !nz[(size_t)blockIdx.y * nz_ld + (k0 >> 4)]) continue;   // this chunk of A is all zeros (block-uniform)
        if (!((gm >> (k0 / cr)) & 1u) && !((gm >> (min(k0 + 15, p - 1) / cr)) & 1u)) continue;   // these rows of G are zero in all 64 columns
        for (int idx = threadIdx.x; idx < 16 * 64; idx += 256) {
            const int kk = idx & 15, ii = idx >> 4;          // A[i0+ii][k0+kk], 16 contiguous doubles per row
            const int i = i0 + ii, k = k0 + kk;
            As[kk][ii] = (i < p && k < p) ? A[(size_t)i * p + k] : 0.0;
            const int jj = idx & 63, k2 = idx >> 6;          // G column j0+jj, row k0+k2
            const int c = j0 + jj, kr = k0 + k2;
            Us[k2][jj] = (kr < p && c < cols) ? (double)G[(size_t)(c / JB) * p * JB + (size_t)kr * JB + (c % JB)] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            double av[4], uv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) av[a] = As[kk][ty * 4 + a];
#pragma unroll
            for (int b = 0; b < 4; ++b) uv[b] = Us[kk][tx * 4 + b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], uv[b], acc[a][b]);
        }
        __syncthreads();
    }
    // numerator partials: sum_i G[i][c] * V[i][c] over this CTA's 64 rows
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const int c = j0 + tx * 4 + b;
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int i = i0 + ty * 4 + a;
            if (i < p && c < cols) s += (double)G[(size_t)(c / JB) * p * JB + (size_t)i * JB + (c % JB)] * acc[a][b];
        }
        red[ty][tx * 4 + b] = s;
    }
    __syncthreads();
    if (threadIdx.x < 64) {
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < 16; ++r) s += red[r][threadIdx.x];
        part[(size_t)blockIdx.y * cols_pad + j0 + threadIdx.x] = s;
    }
}

// The same quotients when the occupancy masks of G's panels are valid (p < 3072), one CTA per column: g_c is non-zero only in the
// occupied row chunks of its panel (~15 % of the rows at 4K), so (A g_c)_i is needed for those rows i only and sums over those k
// only -- p x occ x occ products instead of p^3 / 3.  Also writes the column norm (k_jacobi_norms) -- one launch instead of three.
// Fixed summation order: every rank of a multi-GPU run gets identical bits.
__global__ void __launch_bounds__(256) k_rayleigh_cols(const double* __restrict__ A, const float* __restrict__ G, int p,
                                                       const unsigned* __restrict__ pmask, double* __restrict__ lam, double* __restrict__ mu)
{
    extern __shared__ double gs[];   // [p] the column, zero outside its occupied chunks
    __shared__ double red[8];
    __shared__ double s_n2;
    const int c = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* col = G + (size_t)(c / JB) * p * JB + (c % JB);
    const unsigned occ = pmask[c / JB];
    const int cr = jacobi_chunk_rows(p);
    double nn = 0.0;
    for (int r = tid; r < p; r += 256) {
        const double v = ((occ >> (r / cr)) & 1u) ? (double)col[(size_t)r * JB] : 0.0;
        gs[r] = v;
        nn += v * v;
    }
    nn = warp_sum(nn);
    if (lane == 0) red[warp] = nn;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        s_n2 = t;
    }
    __syncthreads();
    double num = 0.0;
    for (int ch = 0; ch < 32; ++ch) {
        if (!((occ >> ch) & 1u)) continue;
        const int i1 = min(p, (ch + 1) * cr);
        for (int i = ch * cr + warp; i < i1; i += 8) {
            const double gi = gs[i];
            if (gi == 0.0) continue;                       // (uniform over the warp)
            // every lane keeps its own partial of g_i A_ik g_k over the rows of its warp; ONE warp sum at the end (a sum per row
            // put ten dependent shuffles between the loads of consecutive rows)
            const double* Ai = A + (size_t)i * p;
            for (int ch2 = 0; ch2 < 32; ++ch2) {
                if (!((occ >> ch2) & 1u)) continue;
                const int k1 = min(p, (ch2 + 1) * cr);
                for (int k = ch2 * cr + lane; k < k1; k += 32) num = fma(gi * Ai[k], gs[k], num);
            }
        }
    }
    num = warp_sum(num);
    __syncthreads();
    if (lane == 0) red[warp] = num;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        lam[c] = sqrt(s_n2);
        mu[c] = t / s_n2;
    }
}

// mu_c = sum_tiles(part) / lam_c^2 ; also keeps lam (the norm) for the normalisation
__global__ void k_rayleigh_finish(const double* __restrict__ part, int row_tiles, int cols_pad, int p,
                                  const double* __restrict__ lam, double* __restrict__ mu)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p) return;
    double s = 0.0;
    for (int t = 0; t < row_tiles; ++t) s += part[(size_t)t * cols_pad + c];
    const double l = lam[c];
    mu[c] = s / (l * l);
}

// ascending order of the eigenvalues: one CTA, bitonic sort of (lambda, index)
__global__ void __launch_bounds__(1024, 1) k_jacobi_sort(const double* __restrict__ lam, int p, int N, int* __restrict__ order)
{
    extern __shared__ unsigned long long skeys[];
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        unsigned long long k = ~0ull;
        if (i < p) {
            // order-preserving key for signed doubles (gl_eigensolve accepts any symmetric matrix, also an indefinite one): flip all
            // bits of a negative value, set the sign bit of a non-negative one; keep 40 bits of the value, 24 of the index
            unsigned long long bits = (unsigned long long)__double_as_longlong(lam[i]);
            bits = (bits >> 63) ? ~bits : (bits | 0x8000000000000000ull);
            k = (bits & ~0xffffffull) | (unsigned)i;
        }
        skeys[i] = k;
    }
    __syncthreads();
    for (int k = 2; k <= N; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < N; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    unsigned long long a = skeys[i], b = skeys[ixj];
                    bool up = (i & k) == 0;
                    if ((a > b) == up) { skeys[i] = b; skeys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < p; i += blockDim.x) order[i] = (int)(skeys[i] & 0xffffffull);
}

// U[:, j] = g_order[j] / lambda, column-major fp32 with leading dimension ld; mu / 1/mu in fp64
__global__ void k_jacobi_extract(const float* __restrict__ G, const double* __restrict__ lam, const double* __restrict__ ray,
                                 const int* __restrict__ order, int p, int m, int ld, int largest, float* __restrict__ U,
                                 double* __restrict__ mu, double* __restrict__ mu_inv)
{
    const int j = blockIdx.x;
    if (j >= m) return;
    const int c = order[largest ? p - 1 - j : j];
    const float inv = (float)(1.0 / lam[c]);
    const double l = ray[c];
    const float* col = G + (size_t)(c / JB) * p * JB + (c % JB);
    for (int r = threadIdx.x; r < ld; r += blockDim.x) U[(size_t)j * ld + r] = r < p ? col[(size_t)r * JB] * inv : 0.f;
    if (threadIdx.x == 0) {
        mu[j] = l;
        if (mu_inv) mu_inv[j] = 1.0 / l;
    }
}

// deferred convergence report (gl_run_resident): the words the host would have read, turned into the status block on the device
__global__ void k_jacobi_status(const unsigned* __restrict__ ctl, int max_sweeps, float tol, int* __restrict__ dstat)
{
    const int sweeps = (int)ctl[max_sweeps + 1];
    const unsigned last_bits = sweeps < max_sweeps ? ctl[sweeps] : ctl[max_sweeps - 1];
    const float last = __uint_as_float(last_bits);
    dstat[GL_DS_JACOBI_SWEEPS] = sweeps;
    dstat[GL_DS_JACOBI_OFF] = (int)last_bits;
    const bool at_floor = ctl[max_sweeps + 2] != 0u;     // stopped at the fp32 floor, within 4 tol (k_jacobi)
    dstat[GL_DS_JACOBI] = (sweeps >= max_sweeps || (last > tol && !at_floor)) ? 1 : 0;
}

int gl_impl_eigensolve(gl_ctx* ctx, gl_mat* L_A, int m, gl_mat** eigvecs, gl_mat** eigvals, gl_mat** eigvals_inv)
{
    const int p = (int)L_A->rows;
    const int nb = (int)(round_up(p, JP) / JB);  // even number of panels
    // panel rows per shared-memory chunk: the whole pair (its 32 row chunks of ceil(p / 32) rows) up to p = 3072, larger p: two passes
    // over the pair (see jacobi_pair)
    const int pc = p < 3072 ? 32 * (int)ceil_div(p, 32) : 3072;
    const size_t smem = sizeof(float) * ((size_t)pc * JP + 16 * 256) + sizeof(double) * (2 * JP * 17 + 16) + sizeof(int) * (JP + 2 * JB);
    GL_REQUIRE(p <= 8192, "eigensolve: p = %d is beyond what this build sorts and screens (8192)", p);
    gl_buf *G = nullptr, *lam = nullptr, *order = nullptr, *ctl = nullptr, *ray = nullptr, *part = nullptr;
    gl_buf *Cg = nullptr, *prel = nullptr, *scnt = nullptr, *pmaskb = nullptr;
    const int cols_pad = nb * JB;
    const int row_tiles = (int)ceil_div(p, 64);
    gl_mat *U = nullptr, *mu = nullptr, *mui = nullptr;
    int rc = GL_OK;
    const int max_sweeps = ctx->jacobi_max_sweeps;
    do {
        if ((rc = gl_alloc(ctx, sizeof(float) * (size_t)nb * p * JB, &G)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(double) * (size_t)p, &lam)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(int) * (size_t)p, &order)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(double) * (size_t)p, &ray)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(double) * (size_t)row_tiles * (cols_pad + 64), &part)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(unsigned) * (size_t)(max_sweeps + 4), &ctl)) != GL_OK) break;
        GL_CUDA_BREAK(rc, cudaMemsetAsync(ctl->ptr, 0, sizeof(unsigned) * (size_t)(max_sweeps + 4), ctx->stream));
        if ((rc = gl_alloc(ctx, sizeof(float) * (size_t)cols_pad * cols_pad, &Cg)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(float) * (size_t)nb * nb, &prel)) != GL_OK) break;
        if ((rc = gl_alloc(ctx, sizeof(int) * (size_t)max_sweeps * 2 * nb, &scnt)) != GL_OK) break;
        GL_CUDA_BREAK(rc, cudaMemsetAsync(scnt->ptr, 0, sizeof(int) * (size_t)max_sweeps * 2 * nb, ctx->stream));
        const int64_t total = (int64_t)nb * p * JB;
        if ((rc = gl_alloc(ctx, sizeof(unsigned) * (size_t)nb, &pmaskb)) != GL_OK) break;
        GL_CUDA_BREAK(rc, cudaMemsetAsync(pmaskb->ptr, 0, sizeof(unsigned) * (size_t)nb, ctx->stream));
        k_jacobi_init<<<(unsigned)ceil_div(total, 256), 256, 0, ctx->stream>>>((const double*)L_A->buf->ptr, p, nb,
                                                                               (float*)G->ptr, (unsigned*)pmaskb->ptr);
        GL_LAUNCH_CHECK(ctx);

        GL_CUDA_BREAK(rc, cudaFuncSetAttribute(k_jacobi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        GL_CUDA_BREAK(rc, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_jacobi, J_THREADS, smem));
        GL_REQUIRE(per_sm >= 1, "eigensolve: kernel does not fit on an SM");
        const int ctiles = (int)ceil_div(cols_pad, 64);
        int grid = std::max(nb / 2, ctiles * (ctiles + 1) / 2);   // rotation pairs per step / tiles of the Gram screen
        if (grid > ctx->sm_count) grid = ctx->sm_count;
        if (grid > per_sm * ctx->sm_count) grid = per_sm * ctx->sm_count;
        float* Gp = (float*)G->ptr;
        int p_ = p, nb_ = nb, ms = max_sweeps;
        float tol = ctx->jacobi_tol;
        unsigned* off = (unsigned*)ctl->ptr;                       // [max_sweeps + 1] screen results, then the sweep count
        int* done = (int*)((unsigned*)ctl->ptr + max_sweeps + 1);
        float* Cp = (float*)Cg->ptr;
        float* prp = (float*)prel->ptr;
        int* scp = (int*)scnt->ptr;
        int inner = ctx->jacobi_inner > 0 ? ctx->jacobi_inner : J_INNER_MAX;
        int pc_ = pc;
        gl_buf* jprof = nullptr;
        long long* jpp = nullptr;
        if (getenv("GLB200_JACOBI_PROF")) {
            GL_BREAK(rc, gl_alloc(ctx, sizeof(long long) * 64, &jprof));
            GL_CUDA_BREAK(rc, cudaMemsetAsync(jprof->ptr, 0, sizeof(long long) * 64, ctx->stream));
            jpp = (long long*)jprof->ptr;
        }
        unsigned* pmp = (unsigned*)pmaskb->ptr;
        void* args[] = {&Gp, &p_, &pc_, &nb_, &ms, &tol, &inner, &Cp, &prp, &scp, &off, &done, &jpp, &pmp};
        {
            StageTimer kt(ctx, GL_T_K_JACOBI);
            GL_CUDA_BREAK(rc, cudaLaunchCooperativeKernel((void*)k_jacobi, dim3(grid), dim3(J_THREADS), args, smem, ctx->stream));
        }
        ctx->launches++;
        if (jprof) {
            long long h[64];
            cudaMemcpyAsync(h, jprof->ptr, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
            cudaStreamSynchronize(ctx->stream);
            fprintf(stderr, "[jacobi prof] %lld marks, cycles between:", h[63]);
            for (int i = 1; i < (int)h[63]; ++i) fprintf(stderr, " %lld", h[i] - h[i - 1]);
            fprintf(stderr, "  total %lld | last pair of CTA 0: gram %lld reduce+off %lld inner %lld apply %lld\n", h[h[63] - 1] - h[0], h[41] - h[40],
                    h[42] - h[41], h[43] - h[42], h[44] - h[43]);
            gl_buf_release(jprof);
        }

        if (p < 3072) {   // the panels' occupancy masks are maintained: one launch, only the occupied rows of every column
            k_rayleigh_cols<<<p, 256, sizeof(double) * (size_t)p, ctx->stream>>>((const double*)L_A->buf->ptr, (const float*)G->ptr, p,
                                                                               (const unsigned*)pmaskb->ptr, (double*)lam->ptr, (double*)ray->ptr);
            GL_LAUNCH_CHECK(ctx);
        } else {
        k_jacobi_norms<<<(unsigned)ceil_div(p, 8), 256, 0, ctx->stream>>>((const float*)G->ptr, p, (double*)lam->ptr);
        GL_LAUNCH_CHECK(ctx);
        {
            dim3 gr((unsigned)ceil_div(cols_pad, 64), (unsigned)row_tiles);
            k_rayleigh_partial<<<gr, 256, 0, ctx->stream>>>((const double*)L_A->buf->ptr, (const float*)G->ptr, p, (double*)part->ptr,
                                                            cols_pad + 64, cols_pad,
                                                            L_A->aux ? (const unsigned char*)L_A->aux->ptr : nullptr,
                                                            p < 3072 ? (const unsigned*)pmaskb->ptr : nullptr, nb);
            GL_LAUNCH_CHECK(ctx);
            k_rayleigh_finish<<<(unsigned)ceil_div(p, 128), 128, 0, ctx->stream>>>((const double*)part->ptr, row_tiles, cols_pad + 64, p,
                                                                                   (const double*)lam->ptr, (double*)ray->ptr);
            GL_LAUNCH_CHECK(ctx);
        }
        }
        int N = 2;
        while (N < p) N <<= 1;
        GL_REQUIRE(N <= 8192, "eigensolve: p too large for the single-CTA sort");
        GL_CUDA_BREAK(rc, cudaFuncSetAttribute(k_jacobi_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8));
        k_jacobi_sort<<<1, 1024, (size_t)N * 8, ctx->stream>>>((const double*)ray->ptr, p, N, (int*)order->ptr);
        GL_LAUNCH_CHECK(ctx);

        U = gl_mat_new(ctx, GL_MAT_EIGVEC);
        U->rows = U->local_rows = p;
        U->cols = m;
        U->ld = round_up(p, 64);
        U->elem_bytes = 4;
        if ((rc = gl_alloc(ctx, sizeof(float) * (size_t)U->ld * m, &U->buf)) != GL_OK) break;
        mu = gl_mat_new(ctx, GL_MAT_DIAG);
        mu->rows = mu->local_rows = m;
        mu->cols = m;
        mu->ld = 1;
        mu->elem_bytes = 8;
        if ((rc = gl_alloc(ctx, sizeof(double) * (size_t)m, &mu->buf)) != GL_OK) break;
        mui = gl_mat_new(ctx, GL_MAT_DIAG);
        *mui = *mu;
        mui->buf = nullptr;
        if ((rc = gl_alloc(ctx, sizeof(double) * (size_t)m, &mui->buf)) != GL_OK) break;
        k_jacobi_extract<<<m, 256, 0, ctx->stream>>>((const float*)G->ptr, (const double*)lam->ptr, (const double*)ray->ptr,
                                                     (const int*)order->ptr, p, m, (int)U->ld, ctx->eig_largest, (float*)U->buf->ptr,
                                                     (double*)mu->buf->ptr, (double*)mui->buf->ptr);
        GL_LAUNCH_CHECK(ctx);

        if (ctx->async_mode) {   // the report goes to the deferred status block; nobody waits here
            k_jacobi_status<<<1, 1, 0, ctx->stream>>>((const unsigned*)ctl->ptr, max_sweeps, ctx->jacobi_tol, (int*)ctx->dstat->ptr);
            GL_LAUNCH_CHECK(ctx);
            break;
        }
        // convergence report (one small D2H; the solve itself never synchronises with the host)
        GL_BREAK(rc, gl_ensure_pinned(ctx, sizeof(unsigned) * (size_t)(max_sweeps + 4)));
        GL_CUDA_BREAK(rc, cudaMemcpyAsync(ctx->pinned, ctl->ptr, sizeof(unsigned) * (size_t)(max_sweeps + 4), cudaMemcpyDeviceToHost,
                                      ctx->stream));
        GL_CUDA_BREAK(rc, cudaStreamSynchronize(ctx->stream));
        const unsigned* h = (const unsigned*)ctx->pinned;
        const int sweeps = (int)h[max_sweeps + 1];   // rotation sweeps done; screen number `sweeps` ended the loop
        float last = 1.f;
        if (sweeps < max_sweeps) memcpy(&last, &h[sweeps], sizeof(float));
        else memcpy(&last, &h[max_sweeps - 1], sizeof(float));
        const bool at_floor = h[max_sweeps + 2] != 0u;
        if (ctx->verbose)
            fprintf(stderr, "[libglcuda] jacobi: p=%d panels=%d grid=%d sweeps=%d last off=%.3g%s\n", p, nb, grid, sweeps, last,
                    at_floor ? " (fp32 floor, accepted within 4 tol)" : "");
        if (sweeps >= max_sweeps || (last > ctx->jacobi_tol && !at_floor)) {
            gl_set_error("eigensolve: not converged after %d sweeps (off-orthogonality %.3g > %.3g)", sweeps, last, ctx->jacobi_tol);
            rc = GL_ERR_NOTCONVERGED;
            break;
        }
    } while (0);
    if (G) gl_buf_release(G);
    if (lam) gl_buf_release(lam);
    if (order) gl_buf_release(order);
    if (ctl) gl_buf_release(ctl);
    if (Cg) gl_buf_release(Cg);
    if (prel) gl_buf_release(prel);
    if (scnt) gl_buf_release(scnt);
    if (pmaskb) gl_buf_release(pmaskb);
    if (ray) gl_buf_release(ray);
    if (part) gl_buf_release(part);
    if (rc != GL_OK) {
        gl_mat_destroy(U);
        gl_mat_destroy(mu);
        gl_mat_destroy(mui);
        return rc;
    }
    if (eigvecs) *eigvecs = U; else gl_mat_destroy(U);
    if (eigvals) *eigvals = mu; else gl_mat_destroy(mu);
    if (eigvals_inv) *eigvals_inv = mui; else gl_mat_destroy(mui);
    return GL_OK;
}
