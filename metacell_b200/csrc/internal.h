// internal.h -- host-side structures behind the opaque handles of include/mcb200.h and the
// launcher prototypes of the CUDA stages.  Not part of the ABI.
#pragma once
#include "common.cuh"
#include <algorithm>
#include <atomic>
#include <memory>
#include <vector>
#include <string>
#include <map>
#include <mutex>

namespace mcb { struct Staging; }

struct mcb_ctx {
    std::atomic<int> refs{1};   // the creator + every live handle; destroyed when it drops to 0
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t side_stream = nullptr;   // second stream for an exchange that overlaps compute (created on first use)
    cudaStream_t side();
    // size-keyed cache of device blocks: a step repeats the same allocation sizes, so after the first
    // step no allocation reaches the driver (driver allocation calls were measured to stall for tens of
    // ms under a busy host).  Reuse is safe because every user of a block is ordered on `stream`.
    std::multimap<size_t, void*> free_blocks;
    size_t cached_bytes = 0, live_bytes = 0;
    std::mutex alloc_mu;
    void* dev_alloc(size_t bytes);
    void dev_free(void* p, size_t bytes);
    void trim();
    // options
    int gemm_impl = 0;       // 0 tcgen05, 1 simt
    int profile = 0;
    int keep_self = 0;
    int thresh_strict = 1;
    int cap_override = 0;    // test hook: shrink the candidate capacity so rows take the exact-row path
    int sample_impl = -1;    // kNN threshold pass: -1 / 0 = fused histogram kernel when the shape allows (corsample.cu), 1 = stored fp16 block +
                             // threshold kernel, 2 = fused kernel on e4m3 operands (measured dead end, see corsample.cu)
    int sample_factor = 128; // kNN threshold pass: sampled columns expected above a row's true cut (more = tighter filter, longer pass)
    int join_impl = -1;      // balancing join: 0 = routed records, 1 = direct probing (single rank only), 2 = blocked lists (single rank
                             // only), -1 = auto: blocked on one rank, routed across ranks (no list all-gather, no replicated tables)
    int host_threads = 0;    // staging threads per context (0 = min(16, cores / 2), env MCB_HOST_THREADS)
    mcb::Staging* staging = nullptr;     // pinned ring + copy stream + host threads, created on first use (staging.cu)
    // cached sample of column indices for the kNN threshold pass, keyed by (M, S): the sample only steers the filter
    int64_t samp_M = 0, samp_S = 0; void* samp_dev = nullptr; size_t samp_bytes = 0;
    // stage timers
    struct Stage { const char* name; cudaEvent_t a, b; int64_t launches; };
    std::vector<Stage> stages;
    struct Agg { const char* name; float ms; int64_t launches; };
    std::vector<Agg> agg;
};

namespace mcb {

// stream-ordered device buffer from the context's pool (memory stays cached in the pool)
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    mcb_ctx* ctx = nullptr;
    DevBuf() {}
    DevBuf(mcb_ctx* c, size_t count) { alloc(c, count); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept { p = o.p; n = o.n; ctx = o.ctx; o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept { release(); p = o.p; n = o.n; ctx = o.ctx; o.p = nullptr; o.n = 0; return *this; }
    void alloc(mcb_ctx* c, size_t count)
    {
        release();
        ctx = c; n = count;
        if (count == 0) { p = nullptr; return; }
        p = (T*)c->dev_alloc(count * sizeof(T));
    }
    void zero() { if (p) MCB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), ctx->stream)); }
    void release()
    {
        if (p) ctx->dev_free(p, n * sizeof(T));
        p = nullptr; n = 0;
    }
    ~DevBuf() { release(); }
    size_t bytes() const { return n * sizeof(T); }
};

// cudaFuncSetAttribute applies to the CURRENT device only, and one process can hold contexts on several devices
// (mcb_ctx_create(device), mcb_multi): the opt-in to large dynamic shared memory is tracked per (kernel, device).
struct SmemOptIn { std::mutex mu; size_t done[64] = {}; };
template <typename F>
inline void ensure_dyn_smem(SmemOptIn& st, mcb_ctx* ctx, F* kernel, size_t smem)
{
    std::lock_guard<std::mutex> lk(st.mu);
    size_t& d = st.done[ctx->device & 63];
    if (smem > d) {
        MCB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        d = smem;
    }
}

// declared FIRST in every handle struct so it is destroyed LAST (after the DevBufs that use the ctx)
void ctx_retain(mcb_ctx* c);
void ctx_release(mcb_ctx* c);
struct CtxRef {
    mcb_ctx* c = nullptr;
    CtxRef() {}
    CtxRef(const CtxRef&) = delete;
    CtxRef& operator=(const CtxRef&) = delete;
    void set(mcb_ctx* ctx) { c = ctx; ctx_retain(ctx); }
    ~CtxRef() { if (c) ctx_release(c); }
};

struct StageTimer {
    mcb_ctx* ctx; int idx = -1;
    StageTimer(mcb_ctx* c, const char* name, int64_t launches = 1);
    ~StageTimer();
};

// api.cu: library-allocated host results (released by the caller with mcb_free).  Freed blocks are kept in a small
// process-wide cache, so a caller that repeats a call gets pages that are already mapped (first-touch page faults of a
// fresh 150 MB result cost more than its PCIe transfer).
void* host_alloc(size_t bytes);

// staging.cu: pinned-ring transfers of the large host objects (see the header comment there)
Staging* ctx_staging(mcb_ctx* ctx);
void ctx_staging_destroy(mcb_ctx* ctx);
void staged_upload_csc(mcb_ctx* ctx, int32_t ngenes, int64_t ncells, const int64_t* p, const int32_t* i, const double* x,
                       int64_t* d_p, int32_t* d_i, uint32_t* d_x);
void staged_download_edges(mcb_ctx* ctx, int64_t rows, int32_t K, int64_t ne, const int32_t* d_out_deg, const int32_t* d_src_cell,
                           const int32_t* d_e_dst, int32_t* host_src, int32_t* host_dst, double* host_w);
void staged_copy_h2d(mcb_ctx* ctx, void* dev, const void* host, size_t bytes);      // pageable host bytes -> device, via the pinned ring
void staged_copy_d2h(mcb_ctx* ctx, void* host, const void* dev, size_t bytes);      // device -> host bytes via the ring; synchronous
void launch_csc_validate(mcb_ctx* ctx, int32_t ngenes, int64_t ncells, int64_t nnz, const int64_t* p, const int32_t* ridx, int32_t* flags);

// ---- launchers (each enqueues on ctx->stream; no host sync unless stated) -------------------

// downsample.cu
void launch_convert_x(mcb_ctx* ctx, const double* d_xd, uint32_t* d_x, int64_t nnz);
void launch_col_sums(mcb_ctx* ctx, int64_t ncells, const int64_t* p, const uint32_t* x, int64_t* tot);
void launch_downsample(mcb_ctx* ctx, int64_t ncells, const int64_t* p, const uint32_t* x, double n, uint64_t seed,
                       int add_non_dsamp, uint32_t* out_x, uint8_t* kept, int64_t cell_base = 0);
void launch_x_to_double(mcb_ctx* ctx, const uint32_t* d_x, double* d_xd, int64_t nnz);

// features.cu
void launch_feat_stats(mcb_ctx* ctx, int64_t ncells, const int64_t* p, const int32_t* ridx, const uint32_t* x,
                       const int32_t* feat_map, int32_t G, double k_nonz, double* mu, double* inv, int32_t* valid);
void launch_exclusive_scan_i32(mcb_ctx* ctx, const int32_t* in, int32_t* out, int64_t n, int32_t* total);
void launch_feat_rows(mcb_ctx* ctx, int64_t ncells, const int64_t* p, const int32_t* ridx, const uint32_t* x,
                      const int32_t* feat_map, int32_t G, int32_t Gpad, double k_nonz, const double* mu,
                      const double* inv, const int32_t* valid, const int32_t* rowpos, float* Zf, void* Zh, void* Zl,
                      int32_t* cell_of_row, int64_t cell_base = 0);
void launch_dense_stats(mcb_ctx* ctx, int64_t ncells, const double* x, int32_t G, double* mu, double* inv, int32_t* valid);
void launch_dense_rows(mcb_ctx* ctx, int64_t ncells, const double* x, int32_t G, int32_t Gpad, const double* mu,
                       const double* inv, const int32_t* valid, const int32_t* rowpos, float* Zf, void* Zh, void* Zl,
                       int32_t* cell_of_row);
void launch_gather_rows(mcb_ctx* ctx, const void* Z, int64_t row_bytes, const int32_t* rows, int32_t nrows, void* out);

// corknn_simt.cu : C[i][j] = sum_g A[i][g] * B[j][g] on the CUDA cores, fp32 inputs, fp64 accumulate
void launch_simt_store(mcb_ctx* ctx, const float* A, int64_t a_rows, const float* B, int64_t b_rows, int32_t Gpad,
                       float* C, int64_t ldc);
void launch_simt_filter(mcb_ctx* ctx, const float* A, int64_t a_row0, int64_t a_rows, const float* B, int64_t b_rows,
                        int32_t Gpad, const float* thr, uint64_t* cand, uint32_t* cnt, int32_t cap);

// corknn_tc.cu : the tcgen05 / TMEM / TMA kernels on the fp16 hi/lo planes (void* = __half*)
void tc_init();   // resolves cuTensorMapEncodeTiled
void launch_tc_store(mcb_ctx* ctx, const void* Ah, int64_t a_rows_alloc, const void* Bh, int64_t b_rows_alloc,
                     int32_t Gpad, float* C, int64_t ldc, int64_t a_rows, int64_t b_rows);
void launch_tc_store3(mcb_ctx* ctx, const void* Zh, const void* Zl, int64_t rows_alloc, int64_t a_rows,
                      int64_t b_rows, int32_t Gpad, float* C, int64_t ldc);
// candidate records (uint4 = key lo, key hi, local row, 0) appended by the filter epilogue in chunks
struct CandPool {
    void* records; uint32_t* cursor; uint32_t* chunk_fill; uint32_t n_chunks; uint32_t* overflow;
};
uint32_t tc_pool_chunk_records();
uint32_t tc_pool_min_chunks(mcb_ctx* ctx);
void launch_tc_store16(mcb_ctx* ctx, const void* Ahi, int64_t a_rows_alloc, const void* Bhi, int64_t b_rows_alloc,
                       int32_t Gpad, void* C16, int64_t ldc16, int64_t a_rows, int64_t b_rows);
void launch_tc_store3_ab(mcb_ctx* ctx, const void* Ahi, const void* Alo, int64_t a_rows_alloc, const void* Bhi, const void* Blo,
                         int64_t b_rows_alloc, int64_t a_rows, int64_t b_rows, int32_t Gpad, float* C, int64_t ldc);
void launch_tc_filter(mcb_ctx* ctx, const void* Zh, const void* Zl, int64_t rows_alloc, int64_t a_row0,
                      int64_t a_rows, int64_t b_rows, int32_t Gpad, const float* thr, const CandPool& pool, int32_t part = 0,
                      int32_t nparts = 1, int32_t G = 0 /* valid columns (0 = Gpad): trailing all-zero k-steps are skipped */);
// corsample.cu: fused sampling pass (e4m3 operands, per-row histograms in the epilogue)
void launch_z8_rows(mcb_ctx* ctx, const void* Zh, int32_t Gpad, const int32_t* rows, int64_t row0, int64_t nrows, int32_t G8pad, void* out);
void launch_sample_thresholds_fused(mcb_ctx* ctx, int f8, const void* A, int64_t a_rows, const void* B, int64_t Spad, int32_t row_bytes,
                                    int32_t S, int32_t n_tiles1, int32_t r_sel, int32_t c_hi, int32_t c_lo, float eps, float* thr,
                                    uint32_t* n_escaped, int32_t valid_bytes = 0 /* row bytes that hold data (0 = all) */);
// select.cu: symmetric multi-GPU record exchange helpers
void launch_count_records_by_dest(mcb_ctx* ctx, const CandPool& pool, int64_t rows_per_rank, int32_t nranks, uint64_t* counts);
void launch_partition_records(mcb_ctx* ctx, const CandPool& pool, int64_t rows_per_rank, int32_t nranks, const uint64_t* counts,
                              uint64_t* cursor, void* send);
void launch_scatter_flat(mcb_ctx* ctx, const void* records, int64_t n, int64_t row0, uint64_t* cand, uint32_t* cnt, int32_t cap);
// select.cu: bin the pool's records into the per-row candidate lists (cnt counts every record, also past cap)
void launch_scatter_candidates(mcb_ctx* ctx, const CandPool& pool, uint64_t* cand, uint32_t* cnt, int32_t cap);

// select.cu
void launch_sample_threshold(mcb_ctx* ctx, const float* C, int64_t ldc, int64_t rows, int32_t S, int32_t r,
                             float eps, float* thr);
void launch_sample_threshold16(mcb_ctx* ctx, const void* C16, int64_t ldc, int64_t rows, int32_t S, int32_t r,
                               float eps, float* thr);
void launch_topk_rows(mcb_ctx* ctx, const uint64_t* cand, const uint32_t* cnt, int32_t cap, int64_t row0,
                      int64_t rows, int32_t kc_out, int32_t Kc, int32_t* knn_col, float* knn_cor, int32_t* fail_flag);
void launch_compact_flags(mcb_ctx* ctx, const int32_t* flag, int64_t n, int32_t* list, int32_t* n_out);
void launch_sum_u32(mcb_ctx* ctx, const uint32_t* v, int64_t n, uint64_t* out);
void launch_fill_u32(mcb_ctx* ctx, void* v, int64_t n, uint32_t value);
void launch_topk_big(mcb_ctx* ctx, const float* C, int64_t ldc, int64_t M, const int32_t* rows, int32_t nrows,
                     int64_t row0, int32_t kc_out, int32_t Kc, const float* thr_rows /* [own rows] or null */,
                     int32_t* knn_col, float* knn_cor, bool with_self = true);

// gstats.cu: per-gene statistics (scm_gene_stat, reference R/gstats.r:65-205)
struct GeneRaw {            // exact per-gene sums, finished on the host
    unsigned long long tot, sumsq_lo, sumsq_hi, up, sx_s, sxx_lo, sxx_hi, sxt_lo, sxt_hi;
    unsigned long long dtot, dsumsq_lo, dsumsq_hi;
    double nmean_sum;
    uint32_t on, d_on, top[3], pad;
};
static_assert(sizeof(GeneRaw) == 128, "GeneRaw layout");
void launch_gene_bucket(mcb_ctx* ctx, int32_t ngenes, int64_t ncells, int64_t nnz, const int64_t* p, const int32_t* gi,
                        const uint32_t* x, const uint32_t* dsx, uint32_t* cnt, int64_t* start, uint32_t* tx, int32_t* tc,
                        uint32_t* tds);
struct alignas(16) CellInfo { double scale; long long tot; };      // K_std_n / total; total, or -1 - total for cells outside the sz_cor subset
void launch_gene_stats(mcb_ctx* ctx, int32_t ngenes, const int64_t* start, const uint32_t* tx, const int32_t* tc,
                       const uint32_t* tds, const CellInfo* cinfo, int64_t niche_r, void* out_raw);

void launch_mc_tapply(mcb_ctx* ctx, int64_t ncells, int32_t ngenes, const int64_t* p, const int32_t* gi, const uint32_t* x,
                      const int32_t* mc_of_cell, int32_t n_mc, const int64_t* mc_size, unsigned long long* logsum, uint32_t* on,
                      unsigned long long* mc_tot, double* geomean, double* frac_on);

// coclust.cu: device-side CSR build of the host edge table + input validation
void launch_build_cover_csr(mcb_ctx* ctx, int32_t N, int64_t ne, const int32_t* src, const int32_t* dst, const double* w,
                            int32_t* odeg, int32_t* ideg, int32_t* optr, int32_t* iptr, int32_t* odst, uint32_t* ow,
                            int32_t* isrc, uint32_t* iw, void* oe, void* ie, void* nd, uint16_t* ipos, bool sort_out, int32_t* scratch4);
void launch_sample_check(mcb_ctx* ctx, int64_t total, int32_t n_s, const int32_t* samples, int32_t N, int32_t* err);
void launch_iota_i32(mcb_ctx* ctx, int32_t* v, int32_t n);

// balance.cu
void launch_build_hash(mcb_ctx* ctx, const int32_t* knn_col, int64_t M, int32_t Kc, int32_t H, int32_t rank_bits,
                       uint32_t* table);
// knn_col of launch_mutual / launch_prune / launch_route_records: the rank's OWN rows ([rows][Kc], row 0 = global row row0)
void launch_mutual(mcb_ctx* ctx, const int32_t* knn_col, const uint32_t* table, int64_t row0, int64_t rows,
                   int32_t Kc, int32_t H, int32_t rank_bits, int64_t smax, int strict, int keep_self, int32_t kin,
                   uint32_t* sym, uint64_t* thr_in, int have_scores, const uint64_t* mut = nullptr, const uint32_t* n_mut = nullptr);
// blocked join (balance.cu, single rank): lists bucketed by the block of rows of their target, probed block-major, mutual
// keys appended to compact per-row lists (mut [rows][Kc] u64 (s << 32 | j), n_mut [rows])
struct BlockedShape { int32_t blk_shift, nb; };
BlockedShape blocked_shape(int64_t M, int32_t H);
void launch_prepare_blocked(mcb_ctx* ctx, const int32_t* knn_col, int64_t M, int32_t Kc, int32_t H, int32_t rank_bits, const BlockedShape& bs,
                            int keep_self, uint32_t* table, uint32_t* blist, uint16_t* off);
void launch_probe_blocked(mcb_ctx* ctx, const uint32_t* blist, const uint16_t* off, int64_t M, int32_t Kc, const BlockedShape& bs,
                          const uint32_t* table, int32_t H, int32_t rank_bits, int64_t smax, int strict, uint64_t* mut, uint32_t* n_mut);
// routed join (balance.cu): entries bucketed by (owner rank, block of rows) of their target, then probed block by block
struct RouteShape { int32_t blk_shift, nbr, nbuckets, grid; };
RouteShape route_shape(mcb_ctx* ctx, int64_t rows_per_rank, int32_t nranks, int32_t H, int32_t rank_bits);
void launch_route_records(mcb_ctx* ctx, const int32_t* knn_col_own, int64_t row0, int64_t rows, int32_t Kc, int64_t rows_per_rank,
                          const RouteShape& rs, int32_t rank_bits, int keep_self, uint32_t* counts, int64_t* bucket_total,
                          int64_t* bucket_padded, int64_t* bucket_start, uint64_t* records);
void launch_probe(mcb_ctx* ctx, const uint64_t* records, int64_t n_records, const int64_t* seg_prefix, const int64_t* seg_src,
                  const int32_t* seg_block, int32_t nseg, int32_t blk_shift, const uint32_t* table_own, int32_t H, int32_t rank_bits,
                  int32_t Kc, int64_t smax, int strict, uint32_t* sym_own);
void launch_prune(mcb_ctx* ctx, const int32_t* knn_col, const uint32_t* sym, const uint64_t* thr_in, int64_t row0,
                  int64_t rows, int32_t Kc, int32_t K, int64_t smax, int32_t* out_deg, int32_t* out_dst, int32_t* out_s,
                  const uint64_t* mut = nullptr, const uint32_t* n_mut = nullptr);
void launch_widen_i32(mcb_ctx* ctx, const int32_t* in, int64_t* out, int64_t n);
void launch_emit_edges(mcb_ctx* ctx, const int32_t* out_deg, const int32_t* out_dst, const int64_t* offs,
                       const int32_t* cell_of_row, int64_t row0, int64_t rows, int32_t K, int32_t* src, int32_t* dst,
                       double* w);

// coclust.cu
struct CoverGraph {
    int32_t N; const int64_t* optr; const int32_t* odst; const uint32_t* ow;
    const int64_t* iptr; const int32_t* isrc; const uint32_t* iw;
};
// optr32 / iptr32: 32-bit CSR offsets; nd / oe / ie: packed node descriptors (int4) and edges (int2) of the same CSR;
// max_out / max_in: largest degrees
void launch_graph_cover(mcb_ctx* ctx, const CoverGraph& g, const int32_t* optr32, const int32_t* iptr32, const void* nd,
                        const void* oe, const void* ie, int32_t max_out, int32_t max_in, int32_t n_local, const int32_t* resamp_ids, int32_t n_s,
                        const int32_t* samples, const uint64_t* seeds, int32_t min_size, double cooling,
                        int32_t burn_in, int32_t max_sweeps, int32_t ncl_cap, int32_t* mc /* [n_local][N] */,
                        int32_t* f_ws /* [n_local][N] */, int32_t* sweeps, int32_t* status, int32_t knn = 0,
                        const uint16_t* ipos = nullptr, uint16_t* ocut_ws /* [n_local][N] when knn > 0 */ = nullptr);
void launch_cooccur(mcb_ctx* ctx, int32_t N, int32_t n_local, const int32_t* mc, int32_t ncl_cap, uint32_t* cnt,
                    int32_t* nsamp, int32_t* ws_order /* [n_local][N] */, int32_t* ws_start /* [n_local][ncl_cap+1] */);
void launch_cooccur_edges(mcb_ctx* ctx, int32_t N, int32_t n_local, const int32_t* mc, const int32_t* optr32, const int32_t* odst,
                          uint32_t* cnt_e, int32_t* nsamp);
void launch_expand_src(mcb_ctx* ctx, int32_t N, const int32_t* optr32, int32_t* esrc);
void launch_count_pairs(mcb_ctx* ctx, int32_t N, const uint32_t* cnt, int64_t* row_counts);
void launch_extract_pairs(mcb_ctx* ctx, int32_t N, const uint32_t* cnt, const int64_t* row_offs, int32_t* n1,
                          int32_t* n2, int32_t* c);
void launch_exclusive_scan_i64(mcb_ctx* ctx, const int64_t* in, int64_t* out, int64_t n, int64_t* total);

}  // namespace mcb
