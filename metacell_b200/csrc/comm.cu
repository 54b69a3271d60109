// comm.cu -- NCCL inside the library (SURVEY.md section 8b "Threading", 8e): one mcb_comm per rank = the rank's
// context + an ncclComm_t.  The same per-rank code serves both deployments:
//   * one process driving P devices (the R caller: mcb_multi, multi.cu) -- ncclCommInitAll, one host thread per device
//   * one process per device (torchrun bench harness)                   -- ncclCommInitRank with a shared unique id
// NCCL is resolved with dlopen("libnccl.so.2") at the first multi-GPU call, so libmcb200.so has no link-time
// dependency on it (single-GPU users need no NCCL at all) and, inside a Python process that already imported torch,
// the NCCL torch loaded is reused instead of a second copy.  Only the stable 2.x entry points are used.
// Every collective is enqueued on the context's stream: stream order is the only synchronisation.
#include "../../include/mcb200.h"
#include "internal.h"
#include "comm.h"
#include <dlfcn.h>
#include <cstring>
#include <mutex>

namespace mcb {

#define NCCL_FN(ret, name, args) ret(*name) args = nullptr;
struct NcclApi {
    void* handle = nullptr;
    NCCL_FN(ncclResult_t, GetVersion, (int*))
    NCCL_FN(ncclResult_t, GetUniqueId, (ncclUniqueId*))
    NCCL_FN(ncclResult_t, CommInitRank, (ncclComm_t*, int, ncclUniqueId, int))
    NCCL_FN(ncclResult_t, CommInitAll, (ncclComm_t*, int, const int*))
    NCCL_FN(ncclResult_t, CommDestroy, (ncclComm_t))
    NCCL_FN(const char*, GetErrorString, (ncclResult_t))
    NCCL_FN(ncclResult_t, AllGather, (const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t))
    NCCL_FN(ncclResult_t, Broadcast, (const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t))
    NCCL_FN(ncclResult_t, Reduce, (const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t))
    NCCL_FN(ncclResult_t, Send, (const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t))
    NCCL_FN(ncclResult_t, Recv, (void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t))
    NCCL_FN(ncclResult_t, GroupStart, ())
    NCCL_FN(ncclResult_t, GroupEnd, ())
    int version = 0;
};
#undef NCCL_FN

static NcclApi g_nccl;
static std::once_flag g_nccl_once;
static std::string g_nccl_err;

static void load_nccl()
{
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) { g_nccl_err = std::string("MC-ERR: multi-GPU calls need NCCL (libnccl.so.2 not found: ") + dlerror() + ")"; return; }
#define LOAD(field, sym)                                                                          \
    *(void**)(&g_nccl.field) = dlsym(g_nccl.handle, sym);                                         \
    if (!g_nccl.field) { g_nccl_err = std::string("MC-ERR: NCCL symbol missing: ") + sym; return; }
    LOAD(GetVersion, "ncclGetVersion") LOAD(GetUniqueId, "ncclGetUniqueId") LOAD(CommInitRank, "ncclCommInitRank")
    LOAD(CommInitAll, "ncclCommInitAll") LOAD(CommDestroy, "ncclCommDestroy") LOAD(GetErrorString, "ncclGetErrorString")
    LOAD(AllGather, "ncclAllGather") LOAD(Broadcast, "ncclBroadcast") LOAD(Reduce, "ncclReduce") LOAD(Send, "ncclSend")
    LOAD(Recv, "ncclRecv") LOAD(GroupStart, "ncclGroupStart") LOAD(GroupEnd, "ncclGroupEnd")
#undef LOAD
    g_nccl.GetVersion(&g_nccl.version);
}

static NcclApi& nccl()
{
    std::call_once(g_nccl_once, load_nccl);
    if (!g_nccl_err.empty()) throw Error(2, g_nccl_err);
    return g_nccl;
}

#define MCB_NCCL(expr)                                                                           \
    do {                                                                                         \
        ncclResult_t _r = (expr);                                                                \
        if (_r != ncclSuccess) {                                                                 \
            char _b[512];                                                                        \
            snprintf(_b, sizeof(_b), "NCCL error at %s:%d: %s", __FILE__, __LINE__, nccl().GetErrorString(_r)); \
            throw ::mcb::Error(2, _b);                                                           \
        }                                                                                        \
    } while (0)

int nccl_version() { return nccl().version; }

void comm_init_all(int ndev, const int* devices, void** comms_out)
{
    MCB_NCCL(nccl().CommInitAll((ncclComm_t*)comms_out, ndev, devices));
}
void comm_destroy_raw(void* comm) { if (comm) nccl().CommDestroy((ncclComm_t)comm); }

// ---- collectives on the context's stream (bytes; every buffer is device memory of this rank) -----------------
void comm_allgather_inplace(mcb_comm* c, void* buf, size_t bytes_per_rank)
{
    if (c->nranks == 1 || bytes_per_rank == 0) return;
    MCB_NCCL(nccl().AllGather((const char*)buf + (size_t)c->rank * bytes_per_rank, buf, bytes_per_rank, ncclInt8,
                              (ncclComm_t)c->comm, c->ctx->stream));
}
// blocks of different sizes: rank r owns [offs[r], offs[r] + counts[r]) bytes of buf, in place.  Every rank sends its
// block straight to every peer (NVSwitch: all pairs at full bandwidth) -- grouped broadcasts ran at half this rate on 8
// ranks.  `stream` = 0: the context's stream.  Blocks should start on 16-byte boundaries (see comm_alltoallv's callers).
void comm_allgatherv_inplace(mcb_comm* c, void* buf, const size_t* offs, const size_t* counts, cudaStream_t stream)
{
    if (c->nranks == 1) return;
    if (!stream) stream = c->ctx->stream;
    const char* mine = (const char*)buf + offs[c->rank];
    MCB_NCCL(nccl().GroupStart());
    for (int r = 0; r < c->nranks; ++r) {
        if (r == c->rank) continue;
        if (counts[c->rank]) MCB_NCCL(nccl().Send(mine, counts[c->rank], ncclInt8, r, (ncclComm_t)c->comm, stream));
        if (counts[r]) MCB_NCCL(nccl().Recv((char*)buf + offs[r], counts[r], ncclInt8, r, (ncclComm_t)c->comm, stream));
    }
    MCB_NCCL(nccl().GroupEnd());
}
// send_cnt[r] bytes at send + send_off[r] go to rank r; recv_cnt[r] bytes from rank r land at recv + recv_off[r]
void comm_alltoallv(mcb_comm* c, const void* send, const size_t* send_off, const size_t* send_cnt, void* recv,
                    const size_t* recv_off, const size_t* recv_cnt)
{
    if (send_cnt[c->rank])      // own block: a device copy, not a network transfer
        MCB_CUDA(cudaMemcpyAsync((char*)recv + recv_off[c->rank], (const char*)send + send_off[c->rank], send_cnt[c->rank],
                                 cudaMemcpyDeviceToDevice, c->ctx->stream));
    if (c->nranks == 1) return;
    MCB_NCCL(nccl().GroupStart());
    for (int r = 0; r < c->nranks; ++r) {
        if (r == c->rank) continue;
        if (send_cnt[r]) MCB_NCCL(nccl().Send((const char*)send + send_off[r], send_cnt[r], ncclInt8, r, (ncclComm_t)c->comm, c->ctx->stream));
        if (recv_cnt[r]) MCB_NCCL(nccl().Recv((char*)recv + recv_off[r], recv_cnt[r], ncclInt8, r, (ncclComm_t)c->comm, c->ctx->stream));
    }
    MCB_NCCL(nccl().GroupEnd());
}
// every rank's block to `root` (recv_off / recv_cnt only read on root); root's own block is copied on the device
void comm_gatherv(mcb_comm* c, const void* send, size_t send_cnt, void* recv, const size_t* recv_off, const size_t* recv_cnt,
                  int root)
{
    if (c->rank == root && send_cnt && (const char*)send != (char*)recv + recv_off[root])
        MCB_CUDA(cudaMemcpyAsync((char*)recv + recv_off[root], send, send_cnt, cudaMemcpyDeviceToDevice, c->ctx->stream));
    if (c->nranks == 1) return;
    MCB_NCCL(nccl().GroupStart());
    if (c->rank == root) {
        for (int r = 0; r < c->nranks; ++r)
            if (r != root && recv_cnt[r]) MCB_NCCL(nccl().Recv((char*)recv + recv_off[r], recv_cnt[r], ncclInt8, r, (ncclComm_t)c->comm, c->ctx->stream));
    } else if (send_cnt) {
        MCB_NCCL(nccl().Send(send, send_cnt, ncclInt8, root, (ncclComm_t)c->comm, c->ctx->stream));
    }
    MCB_NCCL(nccl().GroupEnd());
}
void comm_reduce_sum_u32(mcb_comm* c, void* buf, size_t count, int root)
{
    if (c->nranks == 1 || count == 0) return;
    MCB_NCCL(nccl().Reduce(buf, buf, count, ncclUint32, ncclSum, root, (ncclComm_t)c->comm, c->ctx->stream));
}
void comm_broadcast_on(mcb_comm* c, void* buf, size_t bytes, int root, cudaStream_t stream)
{
    if (c->nranks == 1 || bytes == 0) return;
    MCB_NCCL(nccl().Broadcast(buf, buf, bytes, ncclInt8, root, (ncclComm_t)c->comm, stream));
}
void comm_broadcast(mcb_comm* c, void* buf, size_t bytes, int root)
{
    if (c->nranks == 1 || bytes == 0) return;
    MCB_NCCL(nccl().Broadcast(buf, buf, bytes, ncclInt8, root, (ncclComm_t)c->comm, c->ctx->stream));
}

// one int64 per rank, gathered to every rank's host (one small collective + one stream sync)
std::vector<int64_t> comm_allgather_i64(mcb_comm* c, int64_t mine)
{
    std::vector<int64_t> out((size_t)c->nranks, mine);
    if (c->nranks == 1) return out;
    mcb_ctx* ctx = c->ctx;
    DevBuf<int64_t> d(ctx, (size_t)c->nranks);
    MCB_CUDA(cudaMemcpyAsync(d.p + c->rank, &mine, sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    comm_allgather_inplace(c, d.p, sizeof(int64_t));
    MCB_CUDA(cudaMemcpyAsync(out.data(), d.p, sizeof(int64_t) * out.size(), cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    return out;
}
// `n` int64 per rank that are already ON THE DEVICE (dev_mine), all ranks' blocks back on the host: [nranks][n].  One host
// synchronisation for the device -> host hop and the exchange together (the host-in variant below costs the caller a
// second one for its own device -> host copy).
std::vector<int64_t> comm_allgather_dev_i64v(mcb_comm* c, const int64_t* dev_mine, int n)
{
    std::vector<int64_t> out((size_t)c->nranks * n);
    mcb_ctx* ctx = c->ctx;
    if (c->nranks == 1) {
        MCB_CUDA(cudaMemcpyAsync(out.data(), dev_mine, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        MCB_CUDA(cudaStreamSynchronize(ctx->stream));
        return out;
    }
    DevBuf<int64_t> d(ctx, out.size());
    MCB_CUDA(cudaMemcpyAsync(d.p + (size_t)c->rank * n, dev_mine, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    comm_allgather_inplace(c, d.p, sizeof(int64_t) * (size_t)n);
    MCB_CUDA(cudaMemcpyAsync(out.data(), d.p, sizeof(int64_t) * out.size(), cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    return out;
}
// `n` int64 per rank (host in), all ranks' blocks back on the host: [nranks][n]
std::vector<int64_t> comm_allgather_i64v(mcb_comm* c, const int64_t* mine, int n)
{
    std::vector<int64_t> out((size_t)c->nranks * n);
    if (c->nranks == 1) { for (int q = 0; q < n; ++q) out[q] = mine[q]; return out; }
    mcb_ctx* ctx = c->ctx;
    DevBuf<int64_t> d(ctx, out.size());
    MCB_CUDA(cudaMemcpyAsync(d.p + (size_t)c->rank * n, mine, sizeof(int64_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    comm_allgather_inplace(c, d.p, sizeof(int64_t) * (size_t)n);
    MCB_CUDA(cudaMemcpyAsync(out.data(), d.p, sizeof(int64_t) * out.size(), cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    return out;
}

}  // namespace mcb

using namespace mcb;

// ---------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------
#define MCB_API_BEGIN try {
#define MCB_API_END                                                                            \
    return 0;                                                                                  \
    } catch (const mcb::Error& e) { mcb::set_error(e.what()); return e.code; }                 \
    catch (const std::bad_alloc&) { mcb::set_error("host out of memory"); return 3; }          \
    catch (const std::exception& e) { mcb::set_error(e.what()); return 4; }                    \
    catch (...) { mcb::set_error("unknown error"); return 5; }

extern "C" int mcb_comm_unique_id(uint8_t* id128)
{
    MCB_API_BEGIN
    MCB_CHECK(id128 != nullptr, "mcb_comm_unique_id: null output");
    static_assert(sizeof(ncclUniqueId) == 128, "unique id size");
    ncclUniqueId id;
    MCB_NCCL(nccl().GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    MCB_API_END
}
extern "C" int mcb_comm_create(mcb_ctx* ctx, int32_t nranks, int32_t rank, const uint8_t* id128, mcb_comm** out)
{
    MCB_API_BEGIN
    MCB_CHECK(ctx && out, "mcb_comm_create: null argument");
    MCB_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, "mcb_comm_create: bad rank / nranks");
    MCB_CUDA(cudaSetDevice(ctx->device));
    std::unique_ptr<mcb_comm> c(new mcb_comm());
    c->ref.set(ctx); c->ctx = ctx; c->rank = rank; c->nranks = nranks;
    if (nranks > 1) {
        MCB_CHECK(id128 != nullptr, "mcb_comm_create: null unique id");
        ncclUniqueId id;
        memcpy(&id, id128, sizeof(id));
        ncclComm_t nc = nullptr;
        MCB_NCCL(nccl().CommInitRank(&nc, nranks, id, rank));
        c->comm = nc; c->owns = true;
    }
    *out = c.release();
    MCB_API_END
}
extern "C" void mcb_comm_destroy(mcb_comm* c)
{
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    if (c->comm && c->owns) comm_destroy_raw(c->comm);
    delete c;
}
extern "C" int mcb_comm_rank(const mcb_comm* c, int32_t* rank, int32_t* nranks)
{
    MCB_API_BEGIN
    MCB_CHECK(c, "null argument");
    if (rank) *rank = c->rank;
    if (nranks) *nranks = c->nranks;
    MCB_API_END
}
