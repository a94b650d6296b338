// Multi-GPU exchange step of the hot path: cone blocks / row panels of
// G are sharded over ranks (one process per GPU); the partial Schur matrices are summed with ONE
// ncclAllReduce per iteration, G'z partial n-vectors with one small allreduce per pass, and
// q-vectors are re-replicated by summing zero-padded local slices.
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 that torch ships / the system one), so the
// library has no link-time dependency on it and single-GPU users never load it.
#include "common.cuh"

#include <dlfcn.h>
#include <cstdlib>
#include <cstring>

namespace {

typedef struct { char internal[128]; } nccl_uid;
typedef int (*fn_get_uid)(nccl_uid*);
typedef int (*fn_init_rank)(void**, int, nccl_uid, int);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_allgather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*fn_destroy)(void*);
typedef const char* (*fn_errstr)(int);

struct NcclApi {
    void* handle = nullptr;
    fn_get_uid get_uid = nullptr;
    fn_init_rank init_rank = nullptr;
    fn_allreduce allreduce = nullptr;
    fn_allgather allgather = nullptr;
    fn_destroy destroy = nullptr;
    fn_errstr errstr = nullptr;
} g_nccl;

constexpr int NCCL_UINT8 = 1, NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_MIN = 3;

void load_nccl() {
    if (g_nccl.handle) return;
    const char* env = getenv("HYP_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        if (!nm || !*nm) continue;
        g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) throw HypError{"NCCL not found: set HYP_NCCL_LIB to the path of libnccl.so.2"};
    g_nccl.get_uid = (fn_get_uid)dlsym(g_nccl.handle, "ncclGetUniqueId");
    g_nccl.init_rank = (fn_init_rank)dlsym(g_nccl.handle, "ncclCommInitRank");
    g_nccl.allreduce = (fn_allreduce)dlsym(g_nccl.handle, "ncclAllReduce");
    g_nccl.allgather = (fn_allgather)dlsym(g_nccl.handle, "ncclAllGather");
    g_nccl.destroy = (fn_destroy)dlsym(g_nccl.handle, "ncclCommDestroy");
    g_nccl.errstr = (fn_errstr)dlsym(g_nccl.handle, "ncclGetErrorString");
    if (!g_nccl.get_uid || !g_nccl.init_rank || !g_nccl.allreduce || !g_nccl.destroy)
        throw HypError{"NCCL library lacks the expected symbols"};
}

void nccl_check(int rc, const char* what) {
    if (rc != 0) {
        std::string msg = std::string(what) + " failed: " +
                          (g_nccl.errstr ? g_nccl.errstr(rc) : "nccl error");
        throw HypError{msg};
    }
}

}  // namespace

void hyp_allreduce_sum(hyp_ctx* ctx, double* buf, int64_t count) {
    if (ctx->nranks <= 1 || count <= 0) return;
    TimeScope ts(ctx, T_ALLREDUCE);
    nccl_check(g_nccl.allreduce(buf, buf, (size_t)count, NCCL_FLOAT64, NCCL_SUM, ctx->nccl_comm, ctx->stream),
               "ncclAllReduce(sum)");
}

void hyp_allreduce_min_u8(hyp_ctx* ctx, uint8_t* buf, int64_t count) {
    if (ctx->nranks <= 1 || count <= 0) return;
    TimeScope ts(ctx, T_ALLREDUCE);
    nccl_check(g_nccl.allreduce(buf, buf, (size_t)count, NCCL_UINT8, NCCL_MIN, ctx->nccl_comm, ctx->stream),
               "ncclAllReduce(min)");
}

void hyp_allgather_inplace(hyp_ctx* ctx, double* buf, int64_t count) {
    if (ctx->nranks <= 1 || count <= 0) return;
    if (!g_nccl.allgather) throw HypError{"NCCL library lacks ncclAllGather"};
    TimeScope ts(ctx, T_ALLREDUCE);
    nccl_check(g_nccl.allgather(buf + (int64_t)ctx->rank * count, buf, (size_t)count, NCCL_FLOAT64, ctx->nccl_comm,
                                ctx->stream),
               "ncclAllGather");
}

void hyp_replicate_q(hyp_ctx* ctx, double* v) {
    if (!hyp_row_sharded(ctx)) return;
    hyp_zero_outside(ctx, v);
    hyp_allreduce_sum(ctx, v, ctx->q);
}

void hyp_comm_destroy(hyp_ctx* ctx) {
    if (ctx->nccl_comm && g_nccl.destroy) g_nccl.destroy(ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->nranks = 1;
    ctx->rank = 0;
}

extern "C" int hyp_comm_unique_id(char* id128) {
    try {
        load_nccl();
        nccl_uid id;
        memset(&id, 0, sizeof(id));
        nccl_check(g_nccl.get_uid(&id), "ncclGetUniqueId");
        memcpy(id128, id.internal, 128);
        return 0;
    } catch (HypError& e) {
        fprintf(stderr, "hyp_comm_unique_id: %s\n", e.msg.c_str());
        return -1;
    }
}

extern "C" int hyp_comm_init(hyp_ctx* ctx, int rank, int nranks, const char* id128) {
    if (!ctx) return -1;
    try {
        if (nranks < 1 || rank < 0 || rank >= nranks) throw HypError{"hyp_comm_init: bad rank / nranks"};
        CUDA_TRY(cudaSetDevice(ctx->device));
        hyp_comm_destroy(ctx);
        if (nranks == 1) return 0;
        load_nccl();
        nccl_uid id;
        memcpy(id.internal, id128, 128);
        void* comm = nullptr;
        nccl_check(g_nccl.init_rank(&comm, nranks, id, rank), "ncclCommInitRank");
        ctx->nccl_comm = comm;
        ctx->rank = rank;
        ctx->nranks = nranks;
        return 0;
    } catch (HypError& e) {
        ctx->last_error = e.msg;
        return -1;
    }
}
