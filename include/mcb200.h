/*
 * mcb200.h -- C ABI of libmcb200.so: the B200 (sm_100a) implementation of MetaCell's balanced
 * cell-graph + resampled co-clustering hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, int status codes,
 * no C++ / torch / R types.  The R `.Call` glue in r/ and the Python host mirror in metacell_b200/
 * bind exactly these symbols.  Citations are to the reference tree (tanaylab/metacell, R package):
 * the arithmetic these entry points replace lives in tgstat, which metacell reaches through
 *     R/utils.r:119   tgs_cor_knn(x, y = x, knn = K * k_expand)           -> mcb_cgraph_knn
 *     R/utils.r:120   tgs_graph(x_knn, knn, k_expand, k_beta)             -> mcb_cgraph_mutual / _prune
 *     R/utils.r:30-69 scm_downsamp(umis, n)                               -> mcb_csc_downsample
 *     R/scmat.r:399-407 scm_which_downsamp_n(mat)                         -> mcb_downsamp_depth
 *     R/gset.r:151-214 gset_get_feat_mat + R/cgraph_knn.r:13-14 log2(1+7u) -> mcb_cgraph_create
 *     R/cgraph_knn.r:10-25 mcell_add_cgraph_from_mat_bknn                 -> mcb_cgraph_bknn (one shot)
 *     R/atlas.r:164-166   tgs_cor_knn(x, y, knn), y != x                     -> mcb_cor_knn_xy
 *     R/gstats.r:65-205   scm_gene_stat per-gene statistics                  -> mcb_gene_stats
 *     R/mc.r:164-241      mc_compute_fp / e_gc / cov_gc (tgs_matrix_tapply)  -> mcb_mc_gene_tapply
 *     R/coclust_graph_cov.r:41 tgs_graph_cover_resample(method = "full")  -> mcb_coclust_*
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; mcb_last_error() then holds a
 *     message (thread local).  No exception or CUDA error crosses this boundary.
 *   - all indices are 0-based (the R glue adds 1); cell / gene counts are int64 / int32.
 *   - "host" pointers are ordinary (ideally pinned) memory owned by the caller, never retained.
 *     "dev" pointers are CUDA device memory on the context's device.
 *   - handles own their device memory; destroy them with the matching *_destroy / *_free.
 *   - the library never falls back to the CPU: without a usable sm_100 device every call fails.
 */
#ifndef MCB200_H
#define MCB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mcb_ctx     mcb_ctx;      /* device + stream + memory pool                    */
typedef struct mcb_csc     mcb_csc;      /* device-resident genes x cells CSC count matrix   */
typedef struct mcb_cgraph  mcb_cgraph;   /* one balanced-graph build (features .. edges)     */
typedef struct mcb_coclust mcb_coclust;  /* one resampled co-clustering run                  */
typedef struct mcb_comm    mcb_comm;     /* one rank of a multi-GPU communicator (context + NCCL)    */
typedef struct mcb_multi   mcb_multi;    /* one process driving several devices (contexts + NCCL + threads) */

int         mcb_version(void);
const char* mcb_last_error(void);
int64_t     mcb_launch_count(void);       /* kernels this library has launched in this process */

/* ---- context ------------------------------------------------------------------------------ */
int   mcb_ctx_create(int device, mcb_ctx** out);
void  mcb_ctx_destroy(mcb_ctx* ctx);
int   mcb_ctx_sync(mcb_ctx* ctx);
void* mcb_ctx_stream(mcb_ctx* ctx);                   /* cudaStream_t all kernels are launched on */
/* options: "gemm_impl" 0 = tcgen05 CTA-pair kernel (default) | 2 = tcgen05 1-SM kernel | 1 = CUDA-core validation kernel;
 *          "profile"   1 = record per-stage CUDA-event timings (mcb_ctx_timings);
 *          "keep_self" 1 = keep self edges in the balanced graph (default 0);
 *          "thresh_strict" 1 = s < K*K*k_expand (default), 0 = s <= ...;
 *          "trim"      release the context's cache of device blocks back to the driver;
 *          "cap_override" n > 0: test hook, caps the per-row candidate capacity so rows take the
 *                      exact-row path;
 *          "host_threads" staging threads of this context (0 = min(16, cores); cores shared between the processes of a node)              */
int   mcb_ctx_set_option(mcb_ctx* ctx, const char* key, int64_t value);
/* stage timings of the calls since the last mcb_ctx_timings_reset (needs option profile=1).
 * Writes up to cap entries; *n receives the number available.  names[i] are static strings.  */
int   mcb_ctx_timings(mcb_ctx* ctx, int cap, const char** names, float* ms, int64_t* launches, int* n);
int   mcb_ctx_timings_reset(mcb_ctx* ctx);

/* ---- A1: scm_which_downsamp_n (R/scmat.r:399-407); pure host arithmetic ------------------- */
int mcb_downsamp_depth(const int64_t* col_sums, int64_t ncells, double* depth);

/* ---- CSC matrix (the dgCMatrix slots of tgScMat@mat: @p, @i, @x, @Dim) --------------------- */
int  mcb_csc_upload(mcb_ctx* ctx, int32_t ngenes, int64_t ncells, const int64_t* p, const int32_t* i,
                    const double* x, mcb_csc** out);                       /* host -> device      */
/* the cell range [cell0, cell0 + ncells) of a larger host matrix: p points at p_full[cell0] and keeps its ABSOLUTE
 * offsets, i and x are the full arrays.  What each rank of a multi-GPU build uploads.
 * Both upload entry points stage through pinned memory with host threads (ordinary pageable R vectors reach PCIe speed),
 * narrow the values on the way (counts to 16 bits with escapes, row indices to 16 bits when ngenes <= 65536), and
 * validate the matrix: x integral, >= 0 and < 2^32; p non-decreasing; 0 <= i < ngenes, strictly ascending per column. */
int  mcb_csc_upload_range(mcb_ctx* ctx, int32_t ngenes, int64_t ncells, const int64_t* p, const int32_t* i,
                          const double* x, mcb_csc** out);
int  mcb_csc_wrap_device(mcb_ctx* ctx, int32_t ngenes, int64_t ncells, const int64_t* dev_p,
                         const int32_t* dev_i, const uint32_t* dev_x, mcb_csc** out); /* no copy  */
void mcb_csc_free(mcb_csc* m);
int  mcb_csc_col_sums(mcb_ctx* ctx, const mcb_csc* m, int64_t* host_tot);  /* colSums(mat@mat)    */
/* A2: scm_downsamp(umis, n) (R/utils.r:30-69).  Cells with colSum < n are dropped (kept = 0, all
 * counts 0) unless add_non_dsamp != 0, in which case they keep their raw counts (R/gset.r:179-187).
 * Every kept cell gets a uniformly random floor(n)-subset of its UMIs (Philox keys).
 * The result shares @p/@i with the input; only the counts differ.                              */
int  mcb_csc_downsample(mcb_ctx* ctx, const mcb_csc* m, double n, uint64_t seed, int add_non_dsamp,
                        mcb_csc** out);
int  mcb_csc_download(mcb_ctx* ctx, const mcb_csc* m, double* host_x, uint8_t* host_kept);

/* ---- A3..A7: balanced K-nn cell graph ------------------------------------------------------- */
/* features: rows feat_rows[0..G) of the matrix (gene-set order, R/gset.r:192-197), log2(1+k_nonz*u)
 * (R/cgraph_knn.r:13-14), per-cell standardisation.  Zero-variance cells are excluded.          */
int  mcb_cgraph_create(mcb_ctx* ctx, const mcb_csc* m, const int32_t* feat_rows, int32_t G, double k_nonz,
                       mcb_cgraph** out);
/* tgs_cor_graph(x = feat, ...) (R/utils.r:115): x is the dense genes x cells matrix the caller
 * already transformed, column-major as in R (one contiguous vector of G doubles per cell).      */
int  mcb_cgraph_create_dense(mcb_ctx* ctx, const double* x, int32_t G, int64_t ncells, mcb_cgraph** out);
void mcb_cgraph_destroy(mcb_cgraph* g);
int  mcb_cgraph_num_valid(const mcb_cgraph* g, int64_t* M);
int  mcb_cgraph_valid_cells(mcb_cgraph* g, int32_t* host_cell_of_row /* [M] */);
int  mcb_cgraph_features(mcb_cgraph* g, float* host_z /* [M][G] */);      /* test / debug export */
/* A5: tgs_cor_knn(x, y = x, knn = K*k_expand).  Self is rank 1 (R/atlas.r:475-476).  rank / nranks: the contiguous
 * split of the M valid cells whose rows this call ranks; a handle made by mcb_cgraph_create is a single-rank build
 * (rank 0 of 1 -- or a row block against all columns, for the kNN lists alone); row-sharded graphs go through
 * mcb_cgraph_build_sharded / mcb_cgraph_bknn_multi, which run the exchanges inside the library.              */
int  mcb_cgraph_knn(mcb_cgraph* g, int32_t K, int32_t k_expand, int32_t rank, int32_t nranks);
int  mcb_cgraph_download_knn(mcb_cgraph* g, int32_t* host_col /* [rows][Kc] */, float* host_cor);
int  mcb_cgraph_knn_stats(mcb_cgraph* g, int64_t* n_fallback_rows, int64_t* n_candidates, int32_t* cap,
                          int32_t* n_sample_cols);
/* A6 step 1: mutual rank products + per-target in-degree cut (K*k_beta) for the owned rows.
 * context option "join_impl": 0 = routed, L2-blocked join | 1 = direct probing join (single rank) | -1 = auto (default:
 * direct on one rank, routed across ranks); both give identical results                                        */
int  mcb_cgraph_mutual(mcb_cgraph* g, int32_t k_beta);
/* A6 steps 2-4: out-degree cut (K) and weights 1 - (place-1)/K for the owned rows.             */
int  mcb_cgraph_prune(mcb_cgraph* g, int64_t* n_edges);
/* edges of the owned rows as (mc1, mc2, w) (R/cgraph_knn.r:22), original 0-based cell ids,
 * ordered by (source, place).                                                                  */
int  mcb_cgraph_edges(mcb_cgraph* g, int32_t* host_src, int32_t* host_dst, double* host_w);
/* the same table left on the device; pointers stay valid until the handle is destroyed                    */
int  mcb_cgraph_edges_device(mcb_cgraph* g, int64_t* n_edges, const int32_t** dev_src, const int32_t** dev_dst,
                             const double** dev_w);
/* test hook: dense correlation block C[a_rows][b_rows] of the first rows of Z through one kernel:
 * impl 0 = tcgen05 3-pass split-TF32, 2 = tcgen05 1-pass TF32, 1 = SIMT fp32.  Call before _mutual. */
int  mcb_cgraph_debug_cor(mcb_cgraph* g, int impl, int64_t a_rows, int64_t b_rows, float* host_c);

/* scm_gene_stat's per-gene statistics (reference R/gstats.r:65-205, up to but excluding the rolling-median
 * trends of :153-190, which are O(genes) and stay with the caller).  m: raw matrix; ds: the result of
 * mcb_csc_downsample(m, n, seed, add_non_dsamp = 0), or NULL (the ds_* columns are then NaN); sub_cells / n_sub:
 * the cells sz_cor is computed on (R draws 50,000 when there are more, :143-146; host-supplied index set, 0-based,
 * distinct), NULL = all cells.  out is [12][ngenes] doubles, unrounded, one row per statistic in this order:
 * tot, var, niche_stat, n_mean, is_on_count, sz_cor, ds_top1, ds_top2, ds_top3, ds_mean, ds_var, ds_is_on_count
 * (definitions: R/gstats.r:83-87 quant_mean with k_reg = 10, :113-127).  All genes are reported; the caller
 * applies f_g = tot > 10 (:95).                                                                     */
int  mcb_gene_stats(mcb_ctx* ctx, const mcb_csc* m, const mcb_csc* ds, double niche_quantile, double k_std_n,
                    const int32_t* sub_cells, int64_t n_sub, double* out);

/* tgs_matrix_tapply(us, mc@mc, f) for the two reductions mc_compute_fp / mc_compute_e_gc / mc_compute_cov_gc use
 * (reference R/mc.r:185-186, :222, :240): geomean[k][g] = exp(mean(log(1 + u[g, cells of k]))) - 1 and
 * frac_on[k][g] = mean(u[g, cells of k] > 0), both [n_mc][ngenes]; mc_size[k] = cells in metacell k and
 * mc_meansize[k] = mean cell total (tapply(colSums(us), mc@mc, mean), R/mc.r:194).  mc_of_cell: 0-based metacell
 * per cell, -1 = outside the cover.  The gene filter rowSums(us) > 10 and the normalisations stay with the caller. */
int  mcb_mc_gene_tapply(mcb_ctx* ctx, const mcb_csc* m, const int32_t* mc_of_cell, int32_t n_mc, double* geomean,
                        double* frac_on, int64_t* mc_size, double* mc_meansize);

/* tgs_cor_knn(x, y, knn) with y != x (reference call sites R/atlas.r:164-166,277,475-478,584): for every
 * cell of x (nx cells, one contiguous vector of G doubles each) the knn cells of y with the highest
 * Pearson correlation, by descending correlation (ties: lower y index).  host_col / host_cor are
 * [nx][knn]; entries beyond the number of non-constant y cells, and all entries of a constant x cell,
 * are col = -1 / cor = NaN (NA in R).  No self match is forced.  knn <= 4096.                      */
int  mcb_cor_knn_xy(mcb_ctx* ctx, const double* x, int64_t nx, const double* y, int64_t ny, int32_t G,
                    int32_t knn, int32_t* host_col, float* host_cor);

/* one shot, single GPU, host in / host out: what mcell_add_cgraph_from_mat_bknn's body becomes.
 * dsamp != 0: depth rule + downsample + add back small cells first (R/cgraph_knn.r:12).
 * Outputs are library-allocated host arrays; release each with mcb_free.                       */
int  mcb_cgraph_bknn(mcb_ctx* ctx, int32_t ngenes, int64_t ncells, const int64_t* p, const int32_t* i,
                     const double* x, const int32_t* feat_rows, int32_t G, int32_t K, int32_t k_expand,
                     int32_t k_beta, double k_nonz, int dsamp, uint64_t seed, int64_t* n_edges,
                     int32_t** src, int32_t** dst, double** w, int64_t* n_valid_cells);
void mcb_free(void* host_ptr);

/* ---- multi-GPU.  NCCL is used inside the library (resolved with dlopen at the first call):
 * correlation tiles and rows shard over ranks; exchanges = all-gather of the feature planes, all-gather of the filter
 * thresholds, all-to-all of candidate records, all-gather of the kNN lists and of the in-degree cuts; co-clustering
 * resamples stride over ranks with one reduce(sum) of the counters.  Two deployments share the same rank-level code:
 *
 * (1) ONE PROCESS, several devices -- what the single R process calls (R/cgraph_knn.r:10-25,
 *     R/coclust_graph_cov.r:13-58).  A session owns a context + communicator + host thread per device; every device
 *     uploads its own cell range straight from the caller's dgCMatrix slots and writes its own edge rows straight
 *     into the result table.  devices == NULL: devices 0 .. ndev-1.                                                */
int      mcb_multi_create(const int32_t* devices, int32_t ndev, mcb_multi** out);
void     mcb_multi_destroy(mcb_multi* m);
int      mcb_multi_size(const mcb_multi* m);
mcb_ctx* mcb_multi_ctx(mcb_multi* m, int32_t rank);                       /* borrowed: timings / options of one device */
int      mcb_multi_set_option(mcb_multi* m, const char* key, int64_t value);
/* arguments and outputs exactly as mcb_cgraph_bknn */
int  mcb_cgraph_bknn_multi(mcb_multi* m, int32_t ngenes, int64_t ncells, const int64_t* p, const int32_t* i,
                           const double* x, const int32_t* feat_rows, int32_t G, int32_t K, int32_t k_expand,
                           int32_t k_beta, double k_nonz, int dsamp, uint64_t seed, int64_t* n_edges,
                           int32_t** src, int32_t** dst, double** w, int64_t* n_valid_cells);
/* tgs_graph_cover_resample(edges, knn, min_mc_size, cooling, burn_in, p_resamp, n_resamp, method = "full")
 * (R/coclust_graph_cov.r:41) on all devices; outputs library-allocated (mcb_free): co_cluster rows + samples[N]     */
int  mcb_coclust_multi(mcb_multi* m, int32_t N, int64_t n_edges, const int32_t* src, const int32_t* dst, const double* w,
                       int32_t n_resamp, int32_t n_s, const int32_t* samples, const uint64_t* seeds, int32_t min_mc_size,
                       double cooling, int32_t burn_in, int32_t max_sweeps, int32_t knn, int64_t* n_pairs, int32_t** node1,
                       int32_t** node2, int32_t** cnt, int32_t** samples_per_node);
/* (2) ONE PROCESS PER DEVICE (e.g. under torchrun / mpirun): rank 0 makes a unique id, the launcher's own channel
 *     carries its 128 bytes to the other ranks, every rank creates its communicator; then the rank-collective calls.  */
int  mcb_comm_unique_id(uint8_t* id128);
int  mcb_comm_create(mcb_ctx* ctx, int32_t nranks, int32_t rank, const uint8_t* id128, mcb_comm** out);
void mcb_comm_destroy(mcb_comm* c);
int  mcb_comm_rank(const mcb_comm* c, int32_t* rank, int32_t* nranks);
/* collective.  local = this rank's cells [cell_base, cell_base + ncells_local) (ascending contiguous ranges in rank
 * order).  A1..A7; afterwards the handle holds the edges of this rank's share of the valid rows.                   */
int  mcb_cgraph_build_sharded(mcb_comm* comm, const mcb_csc* local, int64_t cell_base, const int32_t* feat_rows, int32_t G,
                              double k_nonz, int32_t K, int32_t k_expand, int32_t k_beta, int dsamp, uint64_t seed,
                              mcb_cgraph** out, int64_t* n_edges_local);
/* collective.  All ranks' edges to root's host in source order; outputs library-allocated on root, NULL elsewhere   */
int  mcb_cgraph_gather_edges(mcb_comm* comm, mcb_cgraph* g, int32_t root, int64_t* n_edges_total, int32_t** src,
                             int32_t** dst, double** w);
/* collective.  reduce(sum) of cnt[N][N] and samples[N] into root                                                    */
int  mcb_coclust_reduce(mcb_comm* comm, mcb_coclust* c, int32_t root);

/* ---- A8: tgs_graph_cover_resample(..., method = "full") ------------------------------------ */
/* Graph = directed weighted edges (0-based), every (src, dst) pair listed once as in a tgCellGraph
 * (MC-ERR otherwise).  samples: [n_resamp][n_s] node ids and one uint64
 * seed per resample, both host-supplied.  Runs the resamples r with r % nranks == rank and
 * accumulates cnt[a][b] (a < b) and nsamp[a] on the device.                                     */
int  mcb_coclust_run(mcb_ctx* ctx, int32_t N, int64_t n_edges, const int32_t* src, const int32_t* dst,
                     const double* w, int32_t n_resamp, int32_t n_s, const int32_t* samples,
                     const uint64_t* seeds, int32_t min_mc_size, double cooling, int32_t burn_in,
                     int32_t max_sweeps, int32_t rank, int32_t nranks, mcb_coclust** out);
/* the same with tgs_graph_cover_resample's `knn` argument (R/coclust_graph_cov.r:39-41: K = round(nrow(edges)/ncells), the
 * graph's AVERAGE out-degree): the most out edges a node may use inside one resample, 0 = no limit.  Among a sampled node's
 * out edges whose target is sampled too, the knn of highest weight (ties: lower target index) exist in that resample; the
 * others are dropped for it, in both directions (oracle/oracle.py truncate_edges states the rule).                  */
int  mcb_coclust_run2(mcb_ctx* ctx, int32_t N, int64_t n_edges, const int32_t* src, const int32_t* dst,
                      const double* w, int32_t n_resamp, int32_t n_s, const int32_t* samples,
                      const uint64_t* seeds, int32_t min_mc_size, double cooling, int32_t burn_in,
                      int32_t max_sweeps, int32_t knn, int32_t rank, int32_t nranks, mcb_coclust** out);
/* method = "edges" ([RECALL] tgstat's tgs_graph_cover_resample offers method = "hash" | "full" | "edges"; metacell calls
 * "full", R/coclust_graph_cov.r:41): the same covers, but co-occurrence is counted only for pairs that are edges of the
 * graph -- O(edges) memory instead of the N x N matrix, which cannot exist at 10^6 cells (4 TB).  Read the result with
 * mcb_coclust_edge_counts: one row per edge, ordered by (source, weight descending, target ascending); cnt may be 0.   */
int  mcb_coclust_run_edges(mcb_ctx* ctx, int32_t N, int64_t n_edges, const int32_t* src, const int32_t* dst,
                           const double* w, int32_t n_resamp, int32_t n_s, const int32_t* samples,
                           const uint64_t* seeds, int32_t min_mc_size, double cooling, int32_t burn_in,
                           int32_t max_sweeps, int32_t knn, int32_t rank, int32_t nranks, mcb_coclust** out);
int  mcb_coclust_edge_counts(mcb_coclust* c, int32_t* host_node1 /* [n_edges] */, int32_t* host_node2, int32_t* host_cnt,
                             int32_t* host_samples /* [N] */);
void mcb_coclust_destroy(mcb_coclust* c);
/* tgs_graph_cover(edges, min_mc_size, cooling, burn_in) (R/mc_graphcov.r:27, R/mc_from_coclust.r:247): one cover over
 * all N nodes.  host_cluster[v] = 0 for outliers (R/mc_graphcov.r:28), otherwise a cluster id >= 1.            */
int  mcb_graph_cover(mcb_ctx* ctx, int32_t N, int64_t n_edges, const int32_t* src, const int32_t* dst, const double* w,
                     int32_t min_mc_size, double cooling, int32_t burn_in, int32_t max_sweeps, uint64_t seed,
                     int32_t* host_cluster /* [N] */, int32_t* n_sweeps);
/* device buffers for the NCCL reduce: uint32 cnt[N][N] (upper triangle) and int32 nsamp[N]     */
int  mcb_coclust_buffers(mcb_coclust* c, void** dev_cnt, void** dev_nsamp);
/* per-resample cluster assignment of this rank's resamples (debug / parity): mc[r][v] in
 * {-2 not sampled, -1 outlier, >= 0 cluster}, and sweeps[r]                                    */
int  mcb_coclust_assignments(mcb_coclust* c, int32_t* host_mc /* [n_local][N] */, int32_t* host_sweeps,
                             int32_t* n_local);
/* co_cluster rows with cnt > 0 (node1 < node2, sorted) + samples[N] (R/coclust.r:39-41)        */
int  mcb_coclust_extract(mcb_coclust* c, int64_t* n_pairs);
int  mcb_coclust_pairs(mcb_coclust* c, int32_t* host_node1, int32_t* host_node2, int32_t* host_cnt,
                       int32_t* host_samples);

#ifdef __cplusplus
}
#endif
#endif /* MCB200_H */
