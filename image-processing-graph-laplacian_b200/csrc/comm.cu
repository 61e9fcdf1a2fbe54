// NCCL plumbing for the three small reductions of the path: row sums D (p doubles),
// the optional Gram blocks of the orthogonalisation, and c = Phi^T y (m*C floats).
// NCCL is bound at run time with dlopen so that a single-GPU build/run has no NCCL dependency
// and the library loads on a machine without it.  Replaces the reference's MPI allreduces behind
// VecSum/VecDot/MatNorm (hpc/utils.c:382, hpc/gram_schmidt.c:14-15,59).
#include <dlfcn.h>

#include "common.cuh"

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7, ncclFloat64 = 8 };  // ncclDataType_t
enum { ncclSum = 0 };                        // ncclRedOp_t

struct gl_nccl {
    void* lib = nullptr;
    ncclComm_t comm = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static void* open_nccl()
{
    const char* names[] = {getenv("GLB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n || !*n) continue;
        // RTLD_NOLOAD first: reuse the copy a host framework (e.g. torch) already mapped
        void* h = dlopen(n, RTLD_NOW | RTLD_NOLOAD);
        if (!h) h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) return h;
    }
    return nullptr;
}

static int load_nccl(gl_nccl* c)
{
    if (c->lib) return GL_OK;
    c->lib = open_nccl();
    if (!c->lib) {
        gl_set_error("NCCL not found (set GLB200_NCCL_LIB): %s", dlerror());
        return GL_ERR_NCCL;
    }
    *(void**)&c->GetUniqueId = dlsym(c->lib, "ncclGetUniqueId");
    *(void**)&c->CommInitRank = dlsym(c->lib, "ncclCommInitRank");
    *(void**)&c->CommDestroy = dlsym(c->lib, "ncclCommDestroy");
    *(void**)&c->AllReduce = dlsym(c->lib, "ncclAllReduce");
    *(void**)&c->GetErrorString = dlsym(c->lib, "ncclGetErrorString");
    if (!c->GetUniqueId || !c->CommInitRank || !c->CommDestroy || !c->AllReduce) {
        gl_set_error("NCCL library lacks required symbols");
        return GL_ERR_NCCL;
    }
    return GL_OK;
}

#define GL_NCCL_CHECK(c, expr)                                                                       \
    do {                                                                                             \
        ncclResult_t _r = (expr);                                                                    \
        if (_r != 0) {                                                                               \
            gl_set_error("%s failed: %s", #expr, (c)->GetErrorString ? (c)->GetErrorString(_r) : "?"); \
            return GL_ERR_NCCL;                                                                      \
        }                                                                                            \
    } while (0)

extern "C" int gl_comm_unique_id(void* id128)
{
    GL_REQUIRE(id128, "gl_comm_unique_id: null");
    gl_nccl tmp;
    GL_CHECK(load_nccl(&tmp));
    ncclUniqueId id;
    GL_NCCL_CHECK(&tmp, tmp.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return GL_OK;
}

extern "C" int gl_comm_init(gl_ctx* ctx, const void* id128)
{
    GL_REQUIRE(ctx && id128, "gl_comm_init: null");
    if (ctx->world == 1) return GL_OK;
    GL_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (!ctx->comm) ctx->comm = new gl_nccl();
    GL_CHECK(load_nccl(ctx->comm));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    GL_NCCL_CHECK(ctx->comm, ctx->comm->CommInitRank(&ctx->comm->comm, ctx->world, id, ctx->rank));
    return GL_OK;
}

void gl_comm_destroy(gl_ctx* ctx)
{
    if (!ctx->comm) return;
    if (ctx->comm->comm && ctx->comm->CommDestroy) ctx->comm->CommDestroy(ctx->comm->comm);
    delete ctx->comm;
    ctx->comm = nullptr;
}

static int allreduce(gl_ctx* ctx, void* dev, size_t count, int dtype)
{
    if (ctx->world == 1) return GL_OK;
    if (!ctx->comm || !ctx->comm->comm) {
        gl_set_error("world=%d but gl_comm_init was not called", ctx->world);
        return GL_ERR_NCCL;
    }
    GL_NCCL_CHECK(ctx->comm, ctx->comm->AllReduce(dev, dev, count, dtype, ncclSum, ctx->comm->comm, ctx->stream));
    return GL_OK;
}

int gl_allreduce_f64(gl_ctx* ctx, double* dev, size_t count) { return allreduce(ctx, dev, count, ncclFloat64); }
int gl_allreduce_f32(gl_ctx* ctx, float* dev, size_t count) { return allreduce(ctx, dev, count, ncclFloat32); }
