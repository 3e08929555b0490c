// b2_comm.cu -- NCCL point-to-point plumbing for the z-slab decomposition.
// Replaces BoundaryCommunicator.exchange_domains (fbpic/boundaries/boundary_communicator.py:674-707:
// mpi4py Isend/Irecv through pinned host buffers) with device-to-device ncclSend/ncclRecv between
// z-neighbours over NVLink; slabs of `ng` rows of a [Nz][Nr] array are contiguous, so no pack kernel.
#include "b2_common.cuh"
#include <nccl.h>
#include <cstring>

#define B2_NCCL(call)                                                              \
    do {                                                                           \
        ncclResult_t _r = (call);                                                  \
        if (_r != ncclSuccess) return b2_fail((int)_r, ncclGetErrorString(_r), __FILE__, __LINE__); \
    } while (0)

extern "C" {

int b2_nccl_unique_id(void *id128) {
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    B2_NCCL(ncclGetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return 0;
}

int b2_nccl_init(b2_ctx *ctx, const void *id128, int rank, int size) {
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    B2_CUDA(cudaSetDevice(ctx->device));
    ncclComm_t comm;
    B2_NCCL(ncclCommInitRank(&comm, size, id, rank));
    ctx->nccl_comm = (void *)comm;
    ctx->nccl_rank = rank;
    ctx->nccl_size = size;
    return 0;
}

int b2_nccl_destroy(b2_ctx *ctx) {
    if (ctx->nccl_comm) {
        B2_NCCL(ncclCommDestroy((ncclComm_t)ctx->nccl_comm));
        ctx->nccl_comm = nullptr;
    }
    return 0;
}

int b2_nccl_group_start(void) { B2_NCCL(ncclGroupStart()); return 0; }
int b2_nccl_group_end(void) { B2_NCCL(ncclGroupEnd()); return 0; }

// group start/end bracketed by profiler events on the context stream (slot "comm")
static B2Prof *g_comm_prof = nullptr;
int b2_comm_begin(b2_ctx *ctx) {
    if (g_comm_prof) { delete g_comm_prof; g_comm_prof = nullptr; }
    g_comm_prof = new B2Prof(B2P_COMM, ctx->stream);
    B2_NCCL(ncclGroupStart());
    return 0;
}
int b2_comm_end(b2_ctx *ctx) {
    (void)ctx;
    ncclResult_t r = ncclGroupEnd();
    if (g_comm_prof) { delete g_comm_prof; g_comm_prof = nullptr; }
    if (r != ncclSuccess) return b2_fail((int)r, ncclGetErrorString(r), __FILE__, __LINE__);
    return 0;
}

int b2_nccl_send(b2_ctx *ctx, const void *buf, size_t nbytes, int peer, void *stream) {
    B2_NCCL(ncclSend(buf, nbytes, ncclChar, peer, (ncclComm_t)ctx->nccl_comm, b2_stream_of(ctx, stream)));
    g_b2_launches.fetch_add(1);
    return 0;
}
int b2_nccl_recv(b2_ctx *ctx, void *buf, size_t nbytes, int peer, void *stream) {
    B2_NCCL(ncclRecv(buf, nbytes, ncclChar, peer, (ncclComm_t)ctx->nccl_comm, b2_stream_of(ctx, stream)));
    return 0;
}
int b2_nccl_allreduce_max_f64(b2_ctx *ctx, double *buf, size_t count, void *stream) {
    B2_NCCL(ncclAllReduce(buf, buf, count, ncclDouble, ncclMax, (ncclComm_t)ctx->nccl_comm, b2_stream_of(ctx, stream)));
    g_b2_launches.fetch_add(1);
    return 0;
}

}  // extern "C"
