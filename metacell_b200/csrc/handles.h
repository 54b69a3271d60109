// handles.h -- the structures behind the opaque handles of include/mcb200.h (mcb_csc, mcb_cgraph, mcb_coclust) and the
// rank-level stage functions shared by api.cu (single device), the rank-collective entry points and multi.cu (one
// process driving several devices).  Not part of the ABI.
#pragma once
#include "internal.h"
#include "comm.h"
#include <cuda_fp16.h>

struct mcb_csc {
    mcb::CtxRef ref;
    mcb_ctx* ctx = nullptr;
    int32_t ngenes = 0;
    int64_t ncells = 0, nnz = 0;
    const int64_t* p = nullptr; const int32_t* i = nullptr; const uint32_t* x = nullptr;   // device views
    mcb::DevBuf<int64_t> own_p; mcb::DevBuf<int32_t> own_i; mcb::DevBuf<uint32_t> own_x; mcb::DevBuf<uint8_t> kept;
};

struct mcb_cgraph {
    mcb::CtxRef ref;
    mcb_ctx* ctx = nullptr;
    mcb_comm* comm = nullptr;      // borrowed: set for rank-collective builds (exchanges run inside the stage functions)
    int64_t ncells = 0; int32_t G = 0, Gpad = 0;
    int64_t M = 0, Mpad = 0;
    mcb::DevBuf<float> Zf;              // fp32 plane (CUDA-core kernels; single-rank builds only)
    mcb::DevBuf<__half> Zh, Zl;         // fp16 hi / lo planes of z * 2^12 (tensor-core kernels)
    mcb::DevBuf<int32_t> cell_of_row;
    cudaEvent_t planes_ready = nullptr;  // multi-rank: the lo plane's all-gather runs on the side stream under the sampling pass
    bool planes_pending = false;
    void wait_planes();                  // orders the context's stream behind that all-gather (idempotent)
    ~mcb_cgraph();
    // knn
    int32_t K = 0, k_expand = 0, Kc = 0, kc_out = 0, rank = 0, nranks = 1;
    int64_t rows_per_rank = 0, row0 = 0, rows = 0;
    mcb::DevBuf<int32_t> knn_col; mcb::DevBuf<float> knn_cor;
    int64_t n_fallback = 0, n_candidates = 0; int32_t cap = 0, n_sample = 0;
    double expect_per_row = 0.0; int exchange = 0;
    mcb::DevBuf<float> thr_all;                               // filter thresholds, [nranks * rows_per_rank]
    mcb::DevBuf<uint64_t> cand; mcb::DevBuf<uint32_t> cnt;    // per-row candidate lists of the owned rows
    mcb::DevBuf<uint4> pool_records; mcb::DevBuf<uint32_t> pool_fill, pool_ctl; uint32_t pool_chunks = 0;
    mcb::DevBuf<uint4> send, recv; std::vector<int64_t> send_counts, recv_counts;   // symmetric multi-GPU record exchange
    // balance
    int32_t H = 0, rank_bits = 0, k_beta = 0;
    mcb::DevBuf<uint32_t> sym; mcb::DevBuf<uint64_t> thr_in;
    mcb::DevBuf<uint64_t> mut; mcb::DevBuf<uint32_t> n_mut;   // blocked join: compact per-row lists of mutual keys (replace sym)
    mcb::DevBuf<int32_t> out_deg, out_dst, out_s;
    int64_t n_edges = 0;
    mcb::DevBuf<int32_t> e_src, e_dst; mcb::DevBuf<double> e_w;
};

struct mcb_coclust {
    mcb::CtxRef ref;
    mcb_ctx* ctx = nullptr;
    int32_t N = 0, n_local = 0, ncl_cap = 0;
    int mode = 0;                       // 0 = method "full" (cnt is N x N), 1 = method "edges" (cnt is one counter per edge, out-CSR order)
    int64_t n_edges = 0;
    mcb::DevBuf<int32_t> e_src, e_dst;  // method "edges": the edge behind every counter
    mcb::DevBuf<uint32_t> cnt; mcb::DevBuf<int32_t> nsamp;
    mcb::DevBuf<int32_t> mc, sweeps;
    int64_t n_pairs = -1;
    mcb::DevBuf<int32_t> p1, p2, pc;
};

namespace mcb {

// ---- rank-level stages (api.cu).  `comm` may be null (single rank).  All work is enqueued on ctx->stream. ------------
// host CSC columns [0, ncells) with ABSOLUTE offsets in p (p[0] = first entry of the range) -> device matrix
mcb_csc* csc_upload_impl(mcb_ctx* ctx, int32_t ngenes, int64_t ncells, const int64_t* p, const int32_t* i, const double* x);
// A1 + A2 on a cell shard: depth from the totals of ALL cells (gathered over comm), Philox keys by global cell index
mcb_csc* csc_downsample_auto(mcb_ctx* ctx, mcb_comm* comm, const mcb_csc* m, int64_t cell_base, uint64_t seed, double* depth_out);
// A3 + A4: feature planes of all valid cells on every rank (local rows built here, the rest all-gathered)
mcb_cgraph* cgraph_features_csc(mcb_ctx* ctx, mcb_comm* comm, const mcb_csc* local, int64_t cell_base, const int32_t* feat_rows,
                                int32_t G, double k_nonz);
// A5 .. A7 for this rank's rows, exchanges inside when g->comm is set
void cgraph_knn(mcb_cgraph* g, int32_t K, int32_t k_expand, int32_t rank, int32_t nranks);
void cgraph_mutual(mcb_cgraph* g, int32_t k_beta);
int64_t cgraph_prune(mcb_cgraph* g);
// this rank's edges into host arrays (at the start of each array)
void cgraph_edges_to_host(mcb_cgraph* g, int32_t* host_src, int32_t* host_dst, double* host_w);

// one rank's share of mcell_add_cgraph_from_mat_bknn's body: the host matrix is the WHOLE dgCMatrix (p absolute), the rank
// takes the cells [cell0, cell1)
struct BknnJob {
    int32_t ngenes; int64_t ncells, cell0, cell1; const int64_t* p; const int32_t* i; const double* x;
    const int32_t* feat_rows; int32_t G, K, k_expand, k_beta; double k_nonz; int dsamp; uint64_t seed;
};
mcb_cgraph* bknn_rank(mcb_ctx* ctx, mcb_comm* comm, const BknnJob& job);
mcb_cgraph* bknn_rank_uploaded(mcb_ctx* ctx, mcb_comm* comm, const BknnJob& job, const mcb_csc* local);

}  // namespace mcb
