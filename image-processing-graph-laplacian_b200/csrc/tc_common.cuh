// tcgen05 / TMA / mbarrier building blocks shared by the tensor-core kernels (nystroem_gemm.cu, orthonormalise.cu):
// inline PTX wrappers, shared-memory and instruction descriptors, and the host-side tensor-map encoder.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;       // 64 fp16 = 128 bytes = one SWIZZLE_128B row
constexpr int MAX_BLOCK_N = 256;
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;       // 16 KB
constexpr int B_STAGE_BYTES = MAX_BLOCK_N * BLOCK_K * 2;   // 32 KB
constexpr int SMEM_LIMIT = 227 * 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// try_wait with a suspend-time hint (nanoseconds): the warp sleeps in hardware until the phase completes or the hint runs out,
// instead of coming back at once to poll again.  Measured on the patch kernel (ncu, profiles/r02_patch_polling.md): without the hint
// the waiting warps of a 20-warp CTA executed ~2 000 spin instructions per tile -- more than the epilogue's arithmetic -- and
// their mbarrier polls took 40 % of the shared-memory pipe.
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(ns)
        : "memory");
    return ok != 0;
}
// Bounded wait for long waits (epilogue, producers): sleeps on the barrier; traps like mbar_wait when a pipeline bug hangs it
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, int* err, int code)
{
    if (mbar_try_wait(bar, parity)) return;
    int spins = 0;
    long long t0 = 0;
    while (!mbar_try_wait_hint(bar, parity, 2000u)) {
        if ((++spins & 255) == 0) {
            const long long t = clock64();
            if (t0 == 0) t0 = t;
            if (t - t0 > 4000000000ll) {
                if (err) atomicExch(err, code);
                __threadfence_system();
                asm volatile("trap;");
            }
        }
    }
}

// Bounded wait: a pipeline bug must end in a trap (a CUDA error the host reports), never in a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int code)
{
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) {  // ~2 s
            if (err) atomicExch(err, code);
            __threadfence_system();
            asm volatile("trap;");
        }
    }
}

// two fp32 FMAs in one instruction (FFMA2, sm_100): (d0, d1) += (a0, a1) * (b0, b1), each lane rounded as a plain fma
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1)
{
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%0, %1};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
        : "+f"(d0), "+f"(d1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

// 16-byte load from shared memory by its 32-bit shared address (LDS.128; through a generic pointer the compiler emits LD.E)
__device__ __forceinline__ void lds_f4(uint32_t addr, float* dst)
{
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(dst[0]), "=f"(dst[1]), "=f"(dst[2]), "=f"(dst[3]) : "r"(addr));
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// L2 prefetch of a tile (no shared-memory destination): used to pull the A blocks of tiles a few iterations ahead out
// of HBM early, which buys prefetch depth the shared-memory ring cannot hold
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1)
{
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0),
                 "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all()
{
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (rows of 128 bytes, 8-row groups 1024 bytes apart)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);  // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell), bits [46,48)
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B, bits [61,64)
    return d;
}
// kind::f16 instruction descriptor: (fp16 x fp16 | bf16 x bf16) -> fp32, both operands K-major
__device__ __forceinline__ uint32_t make_idesc(int M, int N, uint32_t ab_format /* 0 = f16, 1 = bf16 */)
{
    uint32_t d = 0;
    d |= 1u << 4;                     // D format: f32
    d |= ab_format << 7;              // A format
    d |= ab_format << 10;             // B format
    d |= (uint32_t)(N >> 3) << 17;    // N / 8
    d |= (uint32_t)(M >> 4) << 24;    // M / 16
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x64(uint32_t taddr, uint32_t (&r)[64])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi)
{
    __half2 v = __floats2half2_rn(lo, hi);
    return *(uint32_t*)&v;
}
__device__ __forceinline__ float2 unpack_h2(uint32_t w)
{
    return __half22float2(*(const __half2*)&w);
}

// The same for rows of BK 16-bit elements: BK = 64 -> SWIZZLE_128B (above), BK = 32 -> SWIZZLE_64B (rows of 64 bytes, 8-row groups
// 512 bytes apart, layout type 4)
template <int BK>
__device__ __forceinline__ uint64_t make_smem_desc_k(uint32_t saddr)
{
    if (BK == 64) return make_smem_desc(saddr);
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;           // stride byte offset: 8 rows x 64 bytes
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;                    // SWIZZLE_64B
    return d;
}

// MN-major ("transposed") SWIZZLE_128B operand: the tile sits in shared memory as K rows of 64 M/N-contiguous elements
// (128 bytes), 8-row groups 1024 bytes apart (stride byte offset), and successive 64-element M/N chunks `chunk_bytes`
// apart (leading byte offset) -- the canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr, uint32_t chunk_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
    d |= (uint64_t)((chunk_bytes >> 4) & 0x3fffu) << 16;   // leading byte offset: next 64-element M/N chunk
    d |= (uint64_t)(1024 >> 4) << 32;                       // stride byte offset: next group of 8 K rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                                 // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor with both operands MN-major (bits 15 and 16)
__device__ __forceinline__ uint32_t make_idesc_mn(int M, int N, uint32_t ab_format)
{
    return make_idesc(M, N, ab_format) | (1u << 15) | (1u << 16);
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------
// host side: tensor maps
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline PFN_encodeTiled get_encode()
{
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// 2-D row-major 16-bit tensor [rows][cols] (cols contiguous), box = box_cols x box_rows, SWIZZLE_128B
static inline int make_map_2d(CUtensorMap* map, CUtensorMapDataType dt, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                       uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        gl_set_error("cuTensorMapEncodeTiled is not available from the driver");
        return GL_ERR_CUDA;
    }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld_elems * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        gl_set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu box=%ux%u", (int)r, (unsigned long long)rows,
                     (unsigned long long)cols, (unsigned long long)ld_elems, box_cols, box_rows);
        return GL_ERR_CUDA;
    }
    return GL_OK;
}

