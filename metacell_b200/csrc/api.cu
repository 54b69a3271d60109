// api.cu -- the extern "C" surface of include/mcb200.h: handles, orchestration of the CUDA stages,
// host<->device staging.  No arithmetic of the path happens on the host; host code here only sizes
// buffers, derives scalar parameters and moves bytes.
#include "../../include/mcb200.h"
#include "internal.h"
#include "handles.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include <cuda_fp16.h>

using namespace mcb;

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_err;
void mcb::set_error(const char* msg) { g_err = msg ? msg : ""; }

#define MCB_API_BEGIN try {
#define MCB_API_END                                                                            \
    return 0;                                                                                  \
    } catch (const mcb::Error& e) { mcb::set_error(e.what()); return e.code; }                 \
    catch (const std::bad_alloc&) { mcb::set_error("host out of memory"); return 3; }          \
    catch (const std::exception& e) { mcb::set_error(e.what()); return 4; }                    \
    catch (...) { mcb::set_error("unknown error"); return 5; }

std::atomic<long long> mcb::g_launches{0};
extern "C" int mcb_version(void) { return 100; }
extern "C" int64_t mcb_launch_count(void) { return (int64_t)mcb::g_launches.load(); }
extern "C" const char* mcb_last_error(void) { return g_err.c_str(); }

// ---------------------------------------------------------------------------------------------
// stage timers
// ---------------------------------------------------------------------------------------------
mcb::StageTimer::StageTimer(mcb_ctx* c, const char* name, int64_t launches) : ctx(c)
{
    if (!c->profile) return;
    mcb_ctx::Stage s; s.name = name; s.launches = launches;
    cudaEventCreate(&s.a); cudaEventCreate(&s.b);
    cudaEventRecord(s.a, c->stream);
    idx = (int)c->stages.size();
    c->stages.push_back(s);
}
mcb::StageTimer::~StageTimer()
{
    if (idx >= 0) cudaEventRecord(ctx->stages[idx].b, ctx->stream);
}
static void resolve_timers(mcb_ctx* ctx)
{
    for (auto& s : ctx->stages) {
        cudaEventSynchronize(s.b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, s.a, s.b);
        bool found = false;
        for (auto& a : ctx->agg) if (strcmp(a.name, s.name) == 0) { a.ms += ms; a.launches += s.launches; found = true; break; }
        if (!found) ctx->agg.push_back({s.name, ms, s.launches});
        cudaEventDestroy(s.a); cudaEventDestroy(s.b);
    }
    ctx->stages.clear();
}

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
static size_t block_size(size_t bytes) { return (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1); }   // 2 MB granules

void* mcb_ctx::dev_alloc(size_t bytes)
{
    const size_t sz = block_size(bytes);
    {
        std::lock_guard<std::mutex> lk(alloc_mu);
        auto it = free_blocks.find(sz);
        if (it != free_blocks.end()) {
            void* p = it->second;
            free_blocks.erase(it);
            cached_bytes -= sz; live_bytes += sz;
            return p;
        }
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, sz);
    if (e == cudaErrorMemoryAllocation) {        // give the cache back and retry once
        cudaGetLastError();
        MCB_CUDA(cudaStreamSynchronize(stream));
        trim();
        e = cudaMalloc(&p, sz);
    }
    if (e != cudaSuccess) {
        char b[200];
        snprintf(b, sizeof(b), "MC-ERR: device allocation of %zu MB failed: %s", sz >> 20, cudaGetErrorString(e));
        cudaGetLastError();
        throw mcb::Error(2, b);
    }
    std::lock_guard<std::mutex> lk(alloc_mu);
    live_bytes += sz;
    return p;
}
void mcb_ctx::dev_free(void* p, size_t bytes)
{
    const size_t sz = block_size(bytes);
    std::lock_guard<std::mutex> lk(alloc_mu);
    free_blocks.emplace(sz, p);
    cached_bytes += sz; live_bytes -= sz;
}
void mcb_ctx::trim()
{
    std::lock_guard<std::mutex> lk(alloc_mu);
    for (auto& kv : free_blocks) cudaFree(kv.second);
    free_blocks.clear();
    cached_bytes = 0;
}

extern "C" int mcb_ctx_create(int device, mcb_ctx** out)
{
    MCB_API_BEGIN
    MCB_CHECK(out != nullptr, "mcb_ctx_create: null output");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw Error(2, "MC-ERR: no CUDA device available; libmcb200 has no CPU fallback");
    MCB_CHECK(device >= 0 && device < ndev, "mcb_ctx_create: bad device index");
    MCB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MCB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        char b[200];
        snprintf(b, sizeof(b), "MC-ERR: libmcb200 is built for sm_100a (B200); device %d is sm_%d%d", device, prop.major, prop.minor);
        throw Error(2, b);
    }
    mcb_ctx* c = new mcb_ctx();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    MCB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    *out = c;
    MCB_API_END
}
cudaStream_t mcb_ctx::side()
{
    if (!side_stream) MCB_CUDA(cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking));
    return side_stream;
}
void mcb_cgraph::wait_planes()
{
    if (!planes_pending) return;
    planes_pending = false;
    MCB_CUDA(cudaStreamWaitEvent(ctx->stream, planes_ready, 0));
}
mcb_cgraph::~mcb_cgraph()
{
    // the buffers go back to the stream-ordered pool: the context's stream must be behind the side stream's writes
    if (planes_pending) cudaStreamWaitEvent(ctx->stream, planes_ready, 0);
    if (planes_ready) cudaEventDestroy(planes_ready);
}
void mcb::ctx_retain(mcb_ctx* c) { c->refs.fetch_add(1); }
void mcb::ctx_release(mcb_ctx* ctx)
{
    if (ctx->refs.fetch_sub(1) != 1) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    resolve_timers(ctx);
    mcb::ctx_staging_destroy(ctx);
    if (ctx->samp_dev) { cudaFree(ctx->samp_dev); ctx->samp_dev = nullptr; }
    ctx->trim();
    if (ctx->side_stream) { cudaStreamSynchronize(ctx->side_stream); cudaStreamDestroy(ctx->side_stream); }
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}
// the context stays alive until the last handle created from it is gone
extern "C" void mcb_ctx_destroy(mcb_ctx* ctx)
{
    if (ctx) mcb::ctx_release(ctx);
}
extern "C" int mcb_ctx_sync(mcb_ctx* ctx)
{
    MCB_API_BEGIN
    MCB_CUDA(cudaSetDevice(ctx->device));
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    MCB_API_END
}
extern "C" void* mcb_ctx_stream(mcb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" int mcb_ctx_set_option(mcb_ctx* ctx, const char* key, int64_t value)
{
    MCB_API_BEGIN
    MCB_CHECK(ctx && key, "mcb_ctx_set_option: null argument");
    if (!strcmp(key, "gemm_impl")) ctx->gemm_impl = (int)value;
    else if (!strcmp(key, "profile")) ctx->profile = (int)value;
    else if (!strcmp(key, "keep_self")) ctx->keep_self = (int)value;
    else if (!strcmp(key, "thresh_strict")) ctx->thresh_strict = (int)value;
    else if (!strcmp(key, "cap_override")) ctx->cap_override = (int)value;
    else if (!strcmp(key, "join_impl")) ctx->join_impl = (int)value;
    else if (!strcmp(key, "sample_impl")) { MCB_CHECK(value >= -1 && value <= 2, "sample_impl must be -1, 0, 1 or 2"); ctx->sample_impl = (int)value; }
    else if (!strcmp(key, "sample_factor")) { MCB_CHECK(value >= 16 && value <= 1024, "sample_factor must be between 16 and 1024"); ctx->sample_factor = (int)value; }
    else if (!strcmp(key, "host_threads")) {
        MCB_CHECK(value >= 0 && value <= 64, "host_threads must be between 0 and 64");
        MCB_CUDA(cudaStreamSynchronize(ctx->stream));
        mcb::ctx_staging_destroy(ctx);                 // rebuilt with the new thread count on next use
        ctx->host_threads = (int)value;
    }
    else if (!strcmp(key, "trim")) { MCB_CUDA(cudaStreamSynchronize(ctx->stream)); ctx->trim(); }
    else throw Error(1, std::string("unknown option: ") + key);
    MCB_API_END
}
extern "C" int mcb_ctx_timings(mcb_ctx* ctx, int cap, const char** names, float* ms, int64_t* launches, int* n)
{
    MCB_API_BEGIN
    MCB_CUDA(cudaSetDevice(ctx->device));
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    resolve_timers(ctx);
    *n = (int)ctx->agg.size();
    for (int i = 0; i < *n && i < cap; ++i) { names[i] = ctx->agg[i].name; ms[i] = ctx->agg[i].ms; launches[i] = ctx->agg[i].launches; }
    MCB_API_END
}
extern "C" int mcb_ctx_timings_reset(mcb_ctx* ctx)
{
    MCB_API_BEGIN
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    resolve_timers(ctx);
    ctx->agg.clear();
    MCB_API_END
}
// ---- host result blocks: malloc + a bounded cache of freed blocks ------------------------------------------------
namespace {
struct HostPool {
    std::mutex mu;
    std::map<void*, size_t> live;                 // blocks handed out by host_alloc
    std::multimap<size_t, void*> idle;            // freed blocks, by size
    size_t idle_bytes = 0;
    static constexpr size_t CAP = (size_t)4 << 30;
    ~HostPool() { for (auto& kv : idle) free(kv.second); }
};
HostPool g_host_pool;
}  // namespace
void* mcb::host_alloc(size_t bytes)
{
    const size_t sz = (std::max<size_t>(bytes, 1) + 4095) & ~(size_t)4095;
    std::lock_guard<std::mutex> lk(g_host_pool.mu);
    auto it = g_host_pool.idle.lower_bound(sz);
    void* p = nullptr;
    size_t got = sz;
    if (it != g_host_pool.idle.end() && it->first <= sz + sz / 4) {       // close enough in size: reuse (pages already mapped)
        p = it->second; got = it->first;
        g_host_pool.idle_bytes -= got;
        g_host_pool.idle.erase(it);
    } else if (posix_memalign(&p, 4096, sz) != 0) p = nullptr;
    if (!p) throw std::bad_alloc();
    g_host_pool.live[p] = got;
    return p;
}
extern "C" void mcb_free(void* p)
{
    if (!p) return;
    std::unique_lock<std::mutex> lk(g_host_pool.mu);
    auto it = g_host_pool.live.find(p);
    if (it == g_host_pool.live.end()) { lk.unlock(); free(p); return; }
    const size_t sz = it->second;
    g_host_pool.live.erase(it);
    if (g_host_pool.idle_bytes + sz <= HostPool::CAP) { g_host_pool.idle.emplace(sz, p); g_host_pool.idle_bytes += sz; return; }
    lk.unlock();
    free(p);
}

// ---------------------------------------------------------------------------------------------
// A1: scm_which_downsamp_n (reference R/scmat.r:399-407): R quantile type 7 + round-half-even
// ---------------------------------------------------------------------------------------------
static double r_quantile7(const std::vector<double>& s, double prob)
{
    const int64_t n = (int64_t)s.size();
    const double h = (double)(n - 1) * prob;
    const int64_t lo = (int64_t)std::floor(h);
    const int64_t hi = lo + 1 < n ? lo + 1 : n - 1;
    return s[lo] + (h - (double)lo) * (s[hi] - s[lo]);
}
extern "C" int mcb_downsamp_depth(const int64_t* col_sums, int64_t ncells, double* depth)
{
    MCB_API_BEGIN
    MCB_CHECK(depth != nullptr, "mcb_downsamp_depth: null output");
    if (ncells <= 0) { *depth = 0.0; return 0; }
    std::vector<double> s((size_t)ncells);
    for (int64_t i = 0; i < ncells; ++i) s[i] = (double)col_sums[i];
    std::sort(s.begin(), s.end());
    const double q50 = std::nearbyint(r_quantile7(s, 0.5));
    const double q05 = std::nearbyint(r_quantile7(s, 0.05));
    const double m = q05 > 750.0 ? q05 : 750.0;
    *depth = q50 < m ? q50 : m;
    MCB_API_END
}

// ---------------------------------------------------------------------------------------------
// CSC matrix
// ---------------------------------------------------------------------------------------------
// Host CSC -> device (pinned-ring staging with narrowing, staging.cu) + structure validation on the device.
// p holds absolute offsets into i / x (p[0] is the first entry of this column range).
mcb_csc* mcb::csc_upload_impl(mcb_ctx* ctx, int32_t ngenes, int64_t ncells, const int64_t* p, const int32_t* i, const double* x)
{
    MCB_CHECK(ngenes >= 0 && ncells >= 0, "mcb_csc_upload: negative dimension");
    MCB_CUDA(cudaSetDevice(ctx->device));
    const int64_t nnz = p[ncells] - p[0];
    MCB_CHECK(p[0] >= 0 && nnz >= 0, "MC-ERR: mcb_csc_upload: malformed column pointer");
    MCB_CHECK(nnz == 0 || (i && x), "mcb_csc_upload: null index/value array");
    std::unique_ptr<mcb_csc> m(new mcb_csc());
    m->ref.set(ctx); m->ctx = ctx; m->ngenes = ngenes; m->ncells = ncells; m->nnz = nnz;
    m->own_p.alloc(ctx, (size_t)ncells + 1);
    m->own_i.alloc(ctx, (size_t)nnz);
    m->own_x.alloc(ctx, (size_t)nnz);
    DevBuf<int32_t> flags(ctx, 1);
    flags.zero();
    {
        StageTimer t(ctx, "h2d_csc");
        staged_upload_csc(ctx, ngenes, ncells, p, i, x, m->own_p.p, m->own_i.p, m->own_x.p);
    }
    // a malformed matrix would send the feature / statistics kernels out of bounds: check the structure once, here
    launch_csc_validate(ctx, ngenes, ncells, nnz, m->own_p.p, m->own_i.p, flags.p);
    int32_t hf = 0;
    MCB_CUDA(cudaMemcpyAsync(&hf, flags.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    MCB_CHECK(!(hf & 1), "MC-ERR: mcb_csc_upload: column pointers must be non-decreasing and end at nnz");
    MCB_CHECK(!(hf & 2), "MC-ERR: mcb_csc_upload: gene row index out of range (is ngenes smaller than max(i) + 1?)");
    MCB_CHECK(!(hf & 4), "MC-ERR: mcb_csc_upload: row indices must be strictly ascending within a column");
    m->p = m->own_p.p; m->i = m->own_i.p; m->x = m->own_x.p;
    return m.release();
}

extern "C" int mcb_csc_upload(mcb_ctx* ctx, int32_t ngenes, int64_t ncells, const int64_t* p, const int32_t* i,
                              const double* x, mcb_csc** out)
{
    MCB_API_BEGIN
    MCB_CHECK(ctx && p && out, "mcb_csc_upload: null argument");
    MCB_CHECK(ncells < 0 || p[0] == 0, "MC-ERR: mcb_csc_upload: malformed column pointer (p[0] must be 0)");
    *out = csc_upload_impl(ctx, ngenes, ncells, p, i, x);
    MCB_API_END
}
// the cell range [cell0, cell0 + ncells) of a larger host matrix: p points at p_full[cell0] (absolute offsets), i and x
// are the full arrays.  What each rank of a multi-GPU build uploads.
extern "C" int mcb_csc_upload_range(mcb_ctx* ctx, int32_t ngenes, int64_t ncells, const int64_t* p, const int32_t* i,
                                    const double* x, mcb_csc** out)
{
    MCB_API_BEGIN
    MCB_CHECK(ctx && p && out, "mcb_csc_upload_range: null argument");
    *out = csc_upload_impl(ctx, ngenes, ncells, p, i, x);
    MCB_API_END
}
extern "C" int mcb_csc_wrap_device(mcb_ctx* ctx, int32_t ngenes, int64_t ncells, const int64_t* dev_p,
                                   const int32_t* dev_i, const uint32_t* dev_x, mcb_csc** out)
{
    MCB_API_BEGIN
    MCB_CHECK(ctx && dev_p && out, "mcb_csc_wrap_device: null argument");
    MCB_CUDA(cudaSetDevice(ctx->device));
    std::unique_ptr<mcb_csc> m(new mcb_csc());
    m->ref.set(ctx); m->ctx = ctx; m->ngenes = ngenes; m->ncells = ncells;
    int64_t nnz = 0;
    MCB_CUDA(cudaMemcpy(&nnz, dev_p + ncells, sizeof(int64_t), cudaMemcpyDeviceToHost));
    m->nnz = nnz; m->p = dev_p; m->i = dev_i; m->x = dev_x;
    *out = m.release();
    MCB_API_END
}
extern "C" void mcb_csc_free(mcb_csc* m) { delete m; }

extern "C" int mcb_csc_col_sums(mcb_ctx* ctx, const mcb_csc* m, int64_t* host_tot)
{
    MCB_API_BEGIN
    MCB_CHECK(ctx && m && host_tot, "mcb_csc_col_sums: null argument");
    MCB_CUDA(cudaSetDevice(ctx->device));
    DevBuf<int64_t> tot(ctx, (size_t)m->ncells);
    { StageTimer t(ctx, "col_sums"); launch_col_sums(ctx, m->ncells, m->p, m->x, tot.p); }
    MCB_CUDA(cudaMemcpyAsync(host_tot, tot.p, sizeof(int64_t) * (size_t)m->ncells, cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    MCB_API_END
}
static mcb_csc* csc_downsample_impl(mcb_ctx* ctx, const mcb_csc* m, double n, uint64_t seed, int add_non_dsamp, int64_t cell_base)
{
    MCB_CHECK(n >= 1.0, "MC-ERR: downsample depth must be >= 1");
    MCB_CUDA(cudaSetDevice(ctx->device));
    std::unique_ptr<mcb_csc> d(new mcb_csc());
    d->ref.set(ctx); d->ctx = ctx; d->ngenes = m->ngenes; d->ncells = m->ncells; d->nnz = m->nnz;
    d->p = m->p; d->i = m->i;                      // shares structure with the input (caller keeps it alive)
    d->own_x.alloc(ctx, (size_t)m->nnz);
    d->kept.alloc(ctx, (size_t)m->ncells);
    { StageTimer t(ctx, "downsample"); launch_downsample(ctx, m->ncells, m->p, m->x, n, seed, add_non_dsamp, d->own_x.p, d->kept.p, cell_base); }
    d->x = d->own_x.p;
    return d.release();
}
extern "C" int mcb_csc_downsample(mcb_ctx* ctx, const mcb_csc* m, double n, uint64_t seed, int add_non_dsamp,
                                  mcb_csc** out)
{
    MCB_API_BEGIN
    MCB_CHECK(ctx && m && out, "mcb_csc_downsample: null argument");
    *out = csc_downsample_impl(ctx, m, n, seed, add_non_dsamp, 0);
    MCB_API_END
}
// gset_get_feat_mat(downsamp = T, add_non_dsamp = T) on a cell shard (R/gset.r:173-187): the depth rule
// (R/scmat.r:399-407) needs the totals of ALL cells, so the shards' column sums are gathered over the communicator and
// every rank derives the same depth; the Philox keys name cells by their index in the whole matrix.
mcb_csc* mcb::csc_downsample_auto(mcb_ctx* ctx, mcb_comm* comm, const mcb_csc* m, int64_t cell_base, uint64_t seed, double* depth_out)
{
    MCB_CUDA(cudaSetDevice(ctx->device));
    const bool multi = comm && comm->nranks > 1;
    int64_t n_all = m->ncells;
    std::vector<int64_t> range;
    if (multi) {
        const int64_t mine[2] = {cell_base, m->ncells};
        range = comm_allgather_i64v(comm, mine, 2);
        n_all = 0;
        for (int r = 0; r < comm->nranks; ++r) n_all = std::max(n_all, range[2 * r] + range[2 * r + 1]);
    }
    DevBuf<int64_t> tot(ctx, (size_t)std::max<int64_t>(n_all, 1));
    { StageTimer t(ctx, "col_sums"); launch_col_sums(ctx, m->ncells, m->p, m->x, tot.p + (multi ? cell_base : 0)); }
    if (multi) {
        std::vector<size_t> offs((size_t)comm->nranks), cnts((size_t)comm->nranks);
        for (int r = 0; r < comm->nranks; ++r) { offs[r] = (size_t)range[2 * r] * 8; cnts[r] = (size_t)range[2 * r + 1] * 8; }
        comm_allgatherv_inplace(comm, tot.p, offs.data(), cnts.data());
    }
    std::vector<int64_t> htot((size_t)n_all);
    if (n_all) MCB_CUDA(cudaMemcpyAsync(htot.data(), tot.p, sizeof(int64_t) * (size_t)n_all, cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    double depth = 0.0;
    if (mcb_downsamp_depth(htot.data(), n_all, &depth)) throw Error(1, mcb_last_error());
    if (depth_out) *depth_out = depth;
    return csc_downsample_impl(ctx, m, depth, seed, /*add_non_dsamp=*/1, multi ? cell_base : 0);
}
extern "C" int mcb_csc_download(mcb_ctx* ctx, const mcb_csc* m, double* host_x, uint8_t* host_kept)
{
    MCB_API_BEGIN
    MCB_CHECK(ctx && m, "mcb_csc_download: null argument");
    MCB_CUDA(cudaSetDevice(ctx->device));
    if (host_x && m->nnz) {
        DevBuf<double> xd(ctx, (size_t)m->nnz);
        launch_x_to_double(ctx, m->x, xd.p, m->nnz);
        MCB_CUDA(cudaMemcpyAsync(host_x, xd.p, sizeof(double) * (size_t)m->nnz, cudaMemcpyDeviceToHost, ctx->stream));
        MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (host_kept) {
        if (m->kept.p) MCB_CUDA(cudaMemcpyAsync(host_kept, m->kept.p, (size_t)m->ncells, cudaMemcpyDeviceToHost, ctx->stream));
        else memset(host_kept, 1, (size_t)m->ncells);
    }
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    MCB_API_END
}

// ---------------------------------------------------------------------------------------------
// per-gene statistics (scm_gene_stat, reference R/gstats.r:65-205)
// ---------------------------------------------------------------------------------------------
static long double u128_ld(unsigned long long lo, unsigned long long hi) { return (long double)hi * 18446744073709551616.0L + (long double)lo; }

extern "C" int mcb_gene_stats(mcb_ctx* ctx, const mcb_csc* m, const mcb_csc* ds, double niche_quantile, double k_std_n,
                              const int32_t* sub_cells, int64_t n_sub, double* out)
{
    MCB_API_BEGIN
    MCB_CHECK(ctx && m && out, "mcb_gene_stats: null argument");
    MCB_CHECK(niche_quantile >= 0.0 && niche_quantile <= 1.0, "niche_quantile must be between 0 and 1");      // R/gstats.r:78-80
    MCB_CHECK(!ds || (ds->ncells == m->ncells && ds->nnz == m->nnz && ds->ngenes == m->ngenes),
              "mcb_gene_stats: the downsampled matrix must share the raw matrix's structure");
    MCB_CHECK(m->nnz < ((int64_t)1 << 32), "mcb_gene_stats: more than 2^32 stored entries");
    MCB_CUDA(cudaSetDevice(ctx->device));
    const int32_t ng = m->ngenes;
    const int64_t n = m->ncells;
    const double nan = std::nan("");
    for (int64_t q = 0; q < (int64_t)12 * ng; ++q) out[q] = nan;
    if (ng == 0 || n == 0) return 0;
    // cell totals over all genes (R/gstats.r:102 colSums(mat)) and the per-cell scale K_std_n / total
    DevBuf<int64_t> ctot(ctx, (size_t)n);
    { StageTimer t(ctx, "col_sums"); launch_col_sums(ctx, n, m->p, m->x, ctot.p); }
    std::vector<int64_t> htot((size_t)n);
    MCB_CUDA(cudaMemcpyAsync(htot.data(), ctot.p, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<uint8_t> hkept;
    if (ds) {
        hkept.resize((size_t)n, 1);
        if (ds->kept.p) MCB_CUDA(cudaMemcpyAsync(hkept.data(), ds->kept.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<double> hscale((size_t)n);
    for (int64_t c = 0; c < n; ++c) hscale[c] = htot[c] > 0 ? k_std_n / (double)htot[c] : 0.0;
    std::vector<uint8_t> hmask;
    int64_t ns = n;
    if (sub_cells) {
        hmask.assign((size_t)n, 0);
        for (int64_t q = 0; q < n_sub; ++q) {
            MCB_CHECK(sub_cells[q] >= 0 && sub_cells[q] < n, "mcb_gene_stats: sz_cor cell index out of range");
            MCB_CHECK(!hmask[sub_cells[q]], "mcb_gene_stats: duplicate sz_cor cell index");
            hmask[sub_cells[q]] = 1;
        }
        ns = n_sub;
    }
    long double T = 0.0L, TT = 0.0L;                                  // sum and sum of squares of the totals over the sz_cor cells
    for (int64_t c = 0; c < n; ++c)
        if (!sub_cells || hmask[c]) { T += (long double)htot[c]; TT += (long double)htot[c] * (long double)htot[c]; }
    int64_t n_ds = 0;
    if (ds) for (int64_t c = 0; c < n; ++c) n_ds += hkept[c] ? 1 : 0;
    std::vector<CellInfo> hinfo((size_t)n);
    for (int64_t c = 0; c < n; ++c) { hinfo[c].scale = hscale[c]; hinfo[c].tot = (!sub_cells || hmask[c]) ? (long long)htot[c] : -1 - (long long)htot[c]; }
    DevBuf<CellInfo> cinfo(ctx, (size_t)n);
    MCB_CUDA(cudaMemcpyAsync(cinfo.p, hinfo.data(), sizeof(CellInfo) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    // gene-major buckets
    DevBuf<uint32_t> cnt(ctx, (size_t)ng), tx(ctx, (size_t)std::max<int64_t>(m->nnz, 1)), tds(ctx, ds ? (size_t)std::max<int64_t>(m->nnz, 1) : 0);
    DevBuf<int32_t> tc(ctx, (size_t)std::max<int64_t>(m->nnz, 1));
    DevBuf<int64_t> start(ctx, (size_t)ng + 1);
    {
        StageTimer t(ctx, "gene_bucket", 3);
        launch_gene_bucket(ctx, ng, n, m->nnz, m->p, m->i, m->x, ds ? ds->x : nullptr, cnt.p, start.p, tx.p, tc.p, ds ? tds.p : nullptr);
    }
    const long long niche_r = (long long)std::nearbyint((double)n * niche_quantile);      // R round(): half to even
    DevBuf<GeneRaw> raw(ctx, (size_t)ng);
    {
        StageTimer t(ctx, "gene_stats");
        launch_gene_stats(ctx, ng, start.p, tx.p, tc.p, ds ? tds.p : nullptr, cinfo.p, niche_r, raw.p);
    }
    std::vector<GeneRaw> h((size_t)ng);
    MCB_CUDA(cudaMemcpyAsync(h.data(), raw.p, sizeof(GeneRaw) * (size_t)ng, cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    double* o_tot = out; double* o_var = out + ng; double* o_niche = out + 2 * (int64_t)ng; double* o_nmean = out + 3 * (int64_t)ng;
    double* o_on = out + 4 * (int64_t)ng; double* o_szcor = out + 5 * (int64_t)ng; double* o_t1 = out + 6 * (int64_t)ng;
    double* o_t2 = out + 7 * (int64_t)ng; double* o_t3 = out + 8 * (int64_t)ng; double* o_dmean = out + 9 * (int64_t)ng;
    double* o_dvar = out + 10 * (int64_t)ng; double* o_don = out + 11 * (int64_t)ng;
    const long double nl = (long double)n, nsl = (long double)ns, ndl = (long double)n_ds;
    const long double vart = ns > 1 ? TT - T * T / nsl : 0.0L;
    for (int32_t g = 0; g < ng; ++g) {
        const GeneRaw& r = h[g];
        const long double tot = (long double)r.tot;
        o_tot[g] = (double)r.tot;
        o_var[g] = n > 1 ? (double)((u128_ld(r.sumsq_lo, r.sumsq_hi) - tot * tot / nl) / (nl - 1.0L)) : nan;
        o_niche[g] = (double)((10.0L + (long double)r.up) / (10.0L + tot));
        o_nmean[g] = r.nmean_sum / (double)n;
        o_on[g] = (double)r.on;
        {
            const long double sx = (long double)r.sx_s;
            const long double varx = u128_ld(r.sxx_lo, r.sxx_hi) - sx * sx / nsl;
            const long double cov = u128_ld(r.sxt_lo, r.sxt_hi) - sx * T / nsl;
            o_szcor[g] = (ns > 1 && varx > 0.0L && vart > 0.0L) ? (double)(cov / sqrtl(varx * vart)) : nan;
        }
        if (ds) {
            o_t1[g] = n_ds >= 1 ? (double)r.top[0] : nan;
            o_t2[g] = n_ds >= 2 ? (double)r.top[1] : nan;
            o_t3[g] = n_ds >= 3 ? (double)r.top[2] : nan;
            const long double dt = (long double)r.dtot;
            o_dmean[g] = n_ds >= 1 ? (double)(dt / ndl) : nan;
            o_dvar[g] = n_ds > 1 ? (double)((u128_ld(r.dsumsq_lo, r.dsumsq_hi) - dt * dt / ndl) / (ndl - 1.0L)) : nan;
            o_don[g] = (double)r.d_on;
        }
    }
    MCB_API_END
}

// tgs_matrix_tapply(us, mc@mc, f) for the two functions metacell uses (reference R/mc.r:185-186, :222, :240):
// geomean[k][g] = exp(mean(log(1 + u[g, cells of k]))) - 1 and frac_on[k][g] = mean(u > 0); plus, per metacell, the
// number of cells and the mean cell total (tapply(colSums(us), mc@mc, mean), R/mc.r:194).  mc_of_cell: 0-based
// metacell of every cell, -1 for cells outside the cover (outliers).  Outputs are [n_mc][ngenes] (row per metacell).
extern "C" int mcb_mc_gene_tapply(mcb_ctx* ctx, const mcb_csc* m, const int32_t* mc_of_cell, int32_t n_mc, double* geomean,
                                  double* frac_on, int64_t* mc_size, double* mc_meansize)
{
    MCB_API_BEGIN
    MCB_CHECK(ctx && m && mc_of_cell && geomean && frac_on && mc_size && mc_meansize, "mcb_mc_gene_tapply: null argument");
    MCB_CHECK(n_mc >= 1, "MC-ERR: no metacells to summarise");
    MCB_CUDA(cudaSetDevice(ctx->device));
    const int64_t n = m->ncells, cells = (int64_t)n_mc * m->ngenes;
    std::vector<int64_t> hsize((size_t)n_mc, 0);
    for (int64_t c = 0; c < n; ++c) {
        MCB_CHECK(mc_of_cell[c] >= -1 && mc_of_cell[c] < n_mc, "MC-ERR: metacell id out of range");
        if (mc_of_cell[c] >= 0) ++hsize[mc_of_cell[c]];
    }
    DevBuf<int32_t> dmc(ctx, (size_t)std::max<int64_t>(n, 1));
    DevBuf<int64_t> dsize(ctx, (size_t)n_mc);
    DevBuf<unsigned long long> logsum(ctx, (size_t)std::max<int64_t>(cells, 1)), mtot(ctx, (size_t)n_mc);
    DevBuf<uint32_t> on(ctx, (size_t)std::max<int64_t>(cells, 1));
    DevBuf<double> dgeo(ctx, (size_t)std::max<int64_t>(cells, 1)), dfrac(ctx, (size_t)std::max<int64_t>(cells, 1));
    MCB_CUDA(cudaMemcpyAsync(dmc.p, mc_of_cell, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    MCB_CUDA(cudaMemcpyAsync(dsize.p, hsize.data(), sizeof(int64_t) * (size_t)n_mc, cudaMemcpyHostToDevice, ctx->stream));
    {
        StageTimer t(ctx, "mc_tapply", 2);
        launch_mc_tapply(ctx, n, m->ngenes, m->p, m->i, m->x, dmc.p, n_mc, dsize.p, logsum.p, on.p, mtot.p, dgeo.p, dfrac.p);
    }
    std::vector<unsigned long long> htot((size_t)n_mc);
    MCB_CUDA(cudaMemcpyAsync(geomean, dgeo.p, sizeof(double) * (size_t)cells, cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaMemcpyAsync(frac_on, dfrac.p, sizeof(double) * (size_t)cells, cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaMemcpyAsync(htot.data(), mtot.p, sizeof(unsigned long long) * (size_t)n_mc, cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int32_t k = 0; k < n_mc; ++k) {
        mc_size[k] = hsize[k];
        mc_meansize[k] = hsize[k] > 0 ? (double)htot[k] / (double)hsize[k] : std::nan("");
    }
    MCB_API_END
}

// ---------------------------------------------------------------------------------------------
// balanced K-nn cell graph
// ---------------------------------------------------------------------------------------------

static void alloc_planes(mcb_cgraph* g, int64_t M, bool with_f32)
{
    mcb_ctx* ctx = g->ctx;
    const size_t n = (size_t)g->Mpad * g->Gpad;
    if (with_f32) g->Zf.alloc(ctx, n);
    g->Zh.alloc(ctx, n); g->Zl.alloc(ctx, n);
    g->cell_of_row.alloc(ctx, (size_t)std::max<int64_t>(M, 1));
    if (g->Mpad > M) {     // zero the padding rows
        const size_t off = (size_t)M * g->Gpad, cnt = (size_t)(g->Mpad - M) * g->Gpad;
        if (with_f32) MCB_CUDA(cudaMemsetAsync(g->Zf.p + off, 0, cnt * sizeof(float), ctx->stream));
        MCB_CUDA(cudaMemsetAsync(g->Zh.p + off, 0, cnt * sizeof(__half), ctx->stream));
        MCB_CUDA(cudaMemsetAsync(g->Zl.p + off, 0, cnt * sizeof(__half), ctx->stream));
    }
}

// A3 + A4.  `m` holds this rank's cells (all cells when comm is null); cell_base = index of its first cell in the whole
// matrix.  Every rank builds the rows of its own valid cells directly at their place in the full planes and the rest
// arrives by an in-place all-gather of variable-size blocks -- the fp16 planes only: the fp32 plane
// serves the CUDA-core validation kernels and exists on single-rank builds alone.  Ranks must hold ascending,
// contiguous cell ranges in rank order, so that valid-row order == cell order.
mcb_cgraph* mcb::cgraph_features_csc(mcb_ctx* ctx, mcb_comm* comm, const mcb_csc* m, int64_t cell_base, const int32_t* feat_rows,
                                     int32_t G, double k_nonz)
{
    MCB_CHECK(G >= 2, "MC-ERR: need at least 2 feature genes");
    MCB_CUDA(cudaSetDevice(ctx->device));
    std::vector<int32_t> fmap((size_t)m->ngenes, -1);
    for (int32_t g = 0; g < G; ++g) {
        MCB_CHECK(feat_rows[g] >= 0 && feat_rows[g] < m->ngenes, "MC-ERR: feature gene row out of range");
        MCB_CHECK(fmap[feat_rows[g]] < 0, "MC-ERR: duplicate feature gene row");
        fmap[feat_rows[g]] = g;
    }
    const bool multi = comm && comm->nranks > 1;
    std::unique_ptr<mcb_cgraph> g(new mcb_cgraph());
    g->ref.set(ctx); g->ctx = ctx; g->comm = multi ? comm : nullptr; g->ncells = m->ncells; g->G = G; g->Gpad = (int32_t)round_up(G, 64);
    const size_t nc1 = (size_t)std::max<int64_t>(m->ncells, 1);
    DevBuf<int32_t> d_fmap(ctx, std::max<size_t>(fmap.size(), 1));
    MCB_CUDA(cudaMemcpyAsync(d_fmap.p, fmap.data(), fmap.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    DevBuf<double> mu(ctx, nc1), inv(ctx, nc1);
    DevBuf<int32_t> valid(ctx, nc1), rowpos(ctx, nc1), total(ctx, 1);
    total.zero();
    {
        StageTimer t(ctx, "feat_stats", 2);
        launch_feat_stats(ctx, m->ncells, m->p, m->i, m->x, d_fmap.p, G, k_nonz, mu.p, inv.p, valid.p);
        if (m->ncells) launch_exclusive_scan_i32(ctx, valid.p, rowpos.p, m->ncells, total.p);
    }
    std::vector<int64_t> counts(1, 0);
    if (multi) {            // every rank's count of valid cells: widened and all-gathered on the device, one host synchronisation
        DevBuf<int64_t> tot64(ctx, 1);
        launch_widen_i32(ctx, total.p, tot64.p, 1);
        counts = comm_allgather_dev_i64v(comm, tot64.p, 1);
    } else {
        int32_t Mloc = 0;
        MCB_CUDA(cudaMemcpyAsync(&Mloc, total.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        MCB_CUDA(cudaStreamSynchronize(ctx->stream));
        counts[0] = Mloc;
    }
    int64_t M = 0, before = 0;
    for (size_t r = 0; r < counts.size(); ++r) { if (multi && (int)r < comm->rank) before += counts[r]; M += counts[r]; }
    g->M = M; g->Mpad = round_up(std::max<int64_t>(M, 1), 256);
    alloc_planes(g.get(), M, !multi);
    {
        StageTimer t(ctx, "feat_rows");
        const size_t off = (size_t)before * g->Gpad;
        launch_feat_rows(ctx, m->ncells, m->p, m->i, m->x, d_fmap.p, G, g->Gpad, k_nonz, mu.p, inv.p, valid.p, rowpos.p,
                         multi ? nullptr : g->Zf.p, g->Zh.p + off, g->Zl.p + off, g->cell_of_row.p + before, multi ? cell_base : 0);
    }
    if (multi) {
        // the hi plane (all the sampling pass needs) on the context's stream; the lo plane and the row -> cell map on the
        // side stream, under the sampling pass (cgraph_knn waits for them before the filter pass)
        std::vector<size_t> offs(counts.size()), cnts(counts.size());
        size_t run = 0;
        for (size_t r = 0; r < counts.size(); ++r) { offs[r] = run * g->Gpad * sizeof(__half); cnts[r] = (size_t)counts[r] * g->Gpad * sizeof(__half); run += (size_t)counts[r]; }
        cudaStream_t side = ctx->side();
        if (!g->planes_ready) MCB_CUDA(cudaEventCreateWithFlags(&g->planes_ready, cudaEventDisableTiming));
        MCB_CUDA(cudaEventRecord(g->planes_ready, ctx->stream));                 // feat_rows done
        MCB_CUDA(cudaStreamWaitEvent(side, g->planes_ready, 0));
        { StageTimer t(ctx, "allgather_features", 0); comm_allgatherv_inplace(comm, g->Zh.p, offs.data(), cnts.data()); }
        comm_allgatherv_inplace(comm, g->Zl.p, offs.data(), cnts.data(), side);
        // row -> cell map: 4-byte items, so a rank's block may start off a 16-byte boundary; it is small, broadcast per rank
        for (size_t r = 0, at = 0; r < counts.size(); at += (size_t)counts[r], ++r)
            if (counts[r]) comm_broadcast_on(comm, g->cell_of_row.p + at, (size_t)counts[r] * sizeof(int32_t), (int)r, side);
        MCB_CUDA(cudaEventRecord(g->planes_ready, side));
        g->planes_pending = true;
    }
    return g.release();          // temporaries are released in stream order (the block cache reuses on this stream only)
}

extern "C" int mcb_cgraph_create(mcb_ctx* ctx, const mcb_csc* m, const int32_t* feat_rows, int32_t G, double k_nonz,
                                 mcb_cgraph** out)
{
    MCB_API_BEGIN
    MCB_CHECK(ctx && m && feat_rows && out, "mcb_cgraph_create: null argument");
    *out = cgraph_features_csc(ctx, nullptr, m, 0, feat_rows, G, k_nonz);
    MCB_API_END
}
// tgs_cor_graph(x = feat, ...) entry (reference R/utils.r:115): x is the dense genes x cells matrix
// the caller already transformed (column-major, one contiguous vector of G doubles per cell)
extern "C" int mcb_cgraph_create_dense(mcb_ctx* ctx, const double* x, int32_t G, int64_t ncells, mcb_cgraph** out)
{
    MCB_API_BEGIN
    MCB_CHECK(ctx && x && out, "mcb_cgraph_create_dense: null argument");
    MCB_CHECK(G >= 2 && ncells >= 1, "MC-ERR: need at least 2 feature genes and 1 cell");
    MCB_CUDA(cudaSetDevice(ctx->device));
    std::unique_ptr<mcb_cgraph> g(new mcb_cgraph());
    g->ref.set(ctx); g->ctx = ctx; g->ncells = ncells; g->G = G; g->Gpad = (int32_t)round_up(G, 64);
    DevBuf<double> dx(ctx, (size_t)ncells * G);
    {
        StageTimer t(ctx, "h2d_dense");
        MCB_CUDA(cudaMemcpyAsync(dx.p, x, sizeof(double) * (size_t)ncells * G, cudaMemcpyHostToDevice, ctx->stream));
    }
    DevBuf<double> mu(ctx, (size_t)ncells), inv(ctx, (size_t)ncells);
    DevBuf<int32_t> valid(ctx, (size_t)ncells), rowpos(ctx, (size_t)ncells), total(ctx, 1);
    {
        StageTimer t(ctx, "feat_stats", 2);
        launch_dense_stats(ctx, ncells, dx.p, G, mu.p, inv.p, valid.p);
        launch_exclusive_scan_i32(ctx, valid.p, rowpos.p, ncells, total.p);
    }
    int32_t M = 0;
    MCB_CUDA(cudaMemcpyAsync(&M, total.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    g->M = M; g->Mpad = round_up(std::max<int64_t>(M, 1), 256);
    alloc_planes(g.get(), M, true);
    {
        StageTimer t(ctx, "feat_rows");
        launch_dense_rows(ctx, ncells, dx.p, G, g->Gpad, mu.p, inv.p, valid.p, rowpos.p, g->Zf.p, g->Zh.p, g->Zl.p, g->cell_of_row.p);
    }
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = g.release();
    MCB_API_END
}
extern "C" void mcb_cgraph_destroy(mcb_cgraph* g)
{
    if (!g) return;
    cudaSetDevice(g->ctx->device);
    delete g;
}
extern "C" int mcb_cgraph_num_valid(const mcb_cgraph* g, int64_t* M)
{
    MCB_API_BEGIN
    MCB_CHECK(g && M, "null argument");
    *M = g->M;
    MCB_API_END
}
extern "C" int mcb_cgraph_valid_cells(mcb_cgraph* g, int32_t* host)
{
    MCB_API_BEGIN
    MCB_CHECK(g && host, "null argument");
    MCB_CUDA(cudaSetDevice(g->ctx->device));
    g->wait_planes();
    if (g->M) MCB_CUDA(cudaMemcpyAsync(host, g->cell_of_row.p, sizeof(int32_t) * (size_t)g->M, cudaMemcpyDeviceToHost, g->ctx->stream));
    MCB_CUDA(cudaStreamSynchronize(g->ctx->stream));
    MCB_API_END
}
extern "C" int mcb_cgraph_features(mcb_cgraph* g, float* host_z)
{
    MCB_API_BEGIN
    MCB_CHECK(g && host_z, "null argument");
    mcb_ctx* ctx = g->ctx;
    MCB_CUDA(cudaSetDevice(ctx->device));
    std::vector<float> z((size_t)g->M * g->Gpad);
    if (g->M) MCB_CUDA(cudaMemcpyAsync(z.data(), g->Zf.p, z.size() * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int64_t r = 0; r < g->M; ++r) memcpy(host_z + r * g->G, z.data() + (size_t)r * g->Gpad, sizeof(float) * (size_t)g->G);
    MCB_API_END
}

// threshold rule of the sampling pass: smallest sample rank r whose expected full-row count stays
// above kc_out with z sigma of margin (hypergeometric variance), plus the matching capacity
static void sample_rule(int64_t M, int64_t S, int32_t kc_out, int32_t& r_out, int32_t& cap_out)
{
    const double z = 4.0;
    const double fpc = S >= M ? 0.0 : std::sqrt(1.0 - (double)S / (double)M);
    const double scale = (double)M / (double)S;
    int32_t r = -1;
    for (int64_t q = 1; q <= S; ++q) {
        const double sd = std::sqrt((double)q * (1.0 - (double)q / (double)S)) * fpc;
        if (((double)q - z * sd) * scale >= (double)kc_out + 2.0) { r = (int32_t)q; break; }
    }
    if (r < 0) { r_out = (int32_t)S + 1; cap_out = (int32_t)round_up(M, 64); return; }    // keep everything
    const double sd = std::sqrt((double)r * (1.0 - (double)r / (double)S)) * fpc;
    const double upper = ((double)r + z * sd) * scale;
    int64_t cap = round_up((int64_t)(upper * 1.05) + 64, 64);       // slack for eps and the threshold's bin resolution
    cap = std::min<int64_t>(cap, round_up(M, 64));
    r_out = r; cap_out = (int32_t)cap;
}

// ---- the kNN stage in three steps (the multi-GPU plumbing exchanges data between them) -------------------
//   knn_begin  : sizes, sampling pass, per-row thresholds of the owned rows (block of thr_all)
//   knn_filter : full 3-pass contraction.  exchange = 0: the rank's own rows (symmetric on one rank, rectangular
//                otherwise) into the pool;  exchange = 1 (multi-GPU): the rank's share of the SYMMETRIC tile list,
//                records for rows of every rank, partitioned by owner into g->send
//   knn_finish : bin records (pool, or the received buffer) into candidate lists, ranked lists, exact rows
static void knn_begin(mcb_cgraph* g, int32_t K, int32_t k_expand, int32_t rank, int32_t nranks)
{
    MCB_CHECK(K >= 1 && k_expand >= 1, "MC-ERR: K and k_expand must be >= 1");
    MCB_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
    MCB_CHECK((int64_t)K * k_expand <= 4096, "MC-ERR: K * k_expand above 4096 is not supported");
    MCB_CHECK(g->M >= 2, "MC-ERR: fewer than 2 cells with non-constant features");
    MCB_CHECK(g->Zh.p, "mcb_cgraph_knn: the feature planes were released (call before mcb_cgraph_mutual)");
    MCB_CHECK(g->ctx->gemm_impl != 1 || g->Zf.p, "gemm_impl = 1 (CUDA-core validation kernels) is a single-rank path");
    mcb_ctx* ctx = g->ctx;
    const int64_t M = g->M;
    const int32_t Gpad = g->Gpad;
    g->K = K; g->k_expand = k_expand; g->Kc = K * k_expand; g->kc_out = (int32_t)std::min<int64_t>(g->Kc, M);
    g->rank = rank; g->nranks = nranks;
    g->rows_per_rank = ceil_div(M, nranks);
    g->row0 = std::min<int64_t>(M, (int64_t)rank * g->rows_per_rank);
    g->rows = std::min<int64_t>(M, g->row0 + g->rows_per_rank) - g->row0;
    const int64_t rows = g->rows, row0 = g->row0;
    const int32_t Kc = g->Kc, kc_out = g->kc_out;
    g->knn_col.alloc(ctx, (size_t)std::max<int64_t>(rows, 1) * Kc);              // the rank's own rows
    g->knn_cor.alloc(ctx, (size_t)std::max<int64_t>(rows, 1) * Kc);
    g->thr_all.alloc(ctx, (size_t)nranks * g->rows_per_rank);
    // rows past M (padding of the last rank's block) never pass: +inf
    launch_fill_u32(ctx, g->thr_all.p, (int64_t)nranks * g->rows_per_rank, 0x7f800000u);
    // ---- sampling pass: per-row filter threshold from a random column sample -----------------
    int64_t S = std::min<int64_t>(262144, round_up(ceil_div((int64_t)ctx->sample_factor * M, kc_out), 256));     // ~sample_factor sampled columns above the true cut
    if (S >= M) S = M;
    const int64_t Spad = round_up(S, 256);
    int32_t r_sel = 0, cap = 0;
    // fused path (corsample.cu): e4m3 operands, per-row histograms in the epilogue, nothing stored.  The first
    // ~5/32 of the sampled columns only bracket each row's quantile; the rank is located among the rest (S_fine).
    bool fused = ctx->gemm_impl == 0 && ctx->sample_impl != 1 && S >= 4096 && S < M;
    int32_t n_tiles1 = 0, c_hi = 0, c_lo = 0;
    int64_t S_fine = S;
    if (fused) {
        const int32_t n_tiles = (int32_t)(Spad / 256);
        n_tiles1 = std::max<int32_t>(2, std::min<int32_t>(n_tiles / 2, (int32_t)((S * 5 / 32 + 128) / 256)));
        S_fine = S - (int64_t)n_tiles1 * 256;
        sample_rule(M, S_fine, kc_out, r_sel, cap);
        if (r_sel > S_fine) fused = false;
        else {
            // bracket ranks in the coarse sample: Poisson tails of the coarse count above the row's fine-sample quantile,
            // each side below ~1e-7 (an escaped row costs a batch of the exact path)
            const double m1 = (double)r_sel / (double)S_fine * (double)n_tiles1 * 256.0;
            const double m_lo = 0.8 * m1, m_hi = 1.25 * m1;
            double term = std::exp(-m_lo), cdf = term;
            c_hi = -1;
            for (int c = 0; cdf <= 1e-7 && c < 100000; ++c) { c_hi = c; term *= m_lo / (double)(c + 1); cdf += term; }
            term = std::exp(-m_hi); cdf = term;               // P(X <= c)
            int c = 0;
            while (1.0 - cdf > 1e-7 && c < 100000) { ++c; term *= m_hi / (double)c; cdf += term; }
            c_lo = c + 1;
            if (c_hi < 0) c_hi = 0;
            if (c_lo <= c_hi) c_lo = c_hi + 1;
        }
    }
    if (!fused) { S_fine = S; sample_rule(M, S, kc_out, r_sel, cap); }
    MCB_CHECK(cap <= 4096, "MC-ERR: K * k_expand is too large relative to the number of cells for this build");
    if (ctx->cap_override > 0) cap = std::min(cap, ctx->cap_override);
    g->cap = cap; g->n_sample = (int32_t)S;
    g->expect_per_row = r_sel > S_fine ? (double)M : (double)r_sel * (double)M / (double)S_fine;
    if (rows == 0) return;
    // the sampled columns (fixed seed: the sample only steers the filter, never the result) are cached on the device
    // per (M, S): a repeated build of the same shape draws nothing on the host
    if (ctx->samp_M != M || ctx->samp_S != S || !ctx->samp_dev) {
        std::vector<int32_t> samp((size_t)Spad, -1);
        if (S >= M) { for (int64_t q = 0; q < M; ++q) samp[q] = (int32_t)q; }
        else {   // partial Fisher-Yates
            std::vector<int32_t> perm((size_t)M);
            for (int64_t q = 0; q < M; ++q) perm[q] = (int32_t)q;
            std::mt19937_64 gen(0x6d63623230300aull);
            for (int64_t q = 0; q < S; ++q) {
                const int64_t j = q + (int64_t)(gen() % (uint64_t)(M - q));
                std::swap(perm[q], perm[j]);
                samp[q] = perm[q];
            }
        }
        MCB_CUDA(cudaStreamSynchronize(ctx->stream));            // nothing queued may still read the old sample
        if (ctx->samp_bytes < sizeof(int32_t) * (size_t)Spad) {
            if (ctx->samp_dev) MCB_CUDA(cudaFree(ctx->samp_dev));
            ctx->samp_dev = nullptr; ctx->samp_bytes = 0; ctx->samp_M = ctx->samp_S = 0;
            MCB_CUDA(cudaMalloc(&ctx->samp_dev, sizeof(int32_t) * (size_t)Spad));
            ctx->samp_bytes = sizeof(int32_t) * (size_t)Spad;
        }
        MCB_CUDA(cudaMemcpyAsync(ctx->samp_dev, samp.data(), sizeof(int32_t) * (size_t)Spad, cudaMemcpyHostToDevice, ctx->stream));
        MCB_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->samp_M = M; ctx->samp_S = S;
    }
    struct { int32_t* p; } d_samp{(int32_t*)ctx->samp_dev};
    float* thr = g->thr_all.p + row0;
    const float eps = 1.0f / 2048.0f;     // covers the 1-pass error of the sampled correlations (hi plane only: ~1e-5 rms, < 5e-4 worst case)
    if (fused && ctx->sample_impl == 2) {            // e4m3 operands: measured dead end, kept for the record (corsample.cu)
        const int32_t G8pad = (int32_t)round_up(Gpad, 128);
        DevBuf<uint8_t> A8(ctx, (size_t)rows * G8pad), B8(ctx, (size_t)Spad * G8pad);
        {
            StageTimer t(ctx, "sample_gather", 2);
            launch_z8_rows(ctx, g->Zh.p, Gpad, nullptr, row0, rows, G8pad, A8.p);
            launch_z8_rows(ctx, g->Zh.p, Gpad, d_samp.p, 0, Spad, G8pad, B8.p);
        }
        StageTimer t(ctx, "cor_sample_hist");
        launch_sample_thresholds_fused(ctx, 1, A8.p, rows, B8.p, Spad, G8pad, (int32_t)S, n_tiles1, r_sel, c_hi, c_lo, eps, thr, nullptr);
        return;                 // temporaries are released in stream order
    }
    if (fused) {
        DevBuf<__half> Zs(ctx, (size_t)Spad * Gpad);
        {
            StageTimer t(ctx, "sample_gather");
            launch_gather_rows(ctx, g->Zh.p, (int64_t)Gpad * sizeof(__half), d_samp.p, (int32_t)Spad, Zs.p);
        }
        StageTimer t(ctx, "cor_sample_hist");
        launch_sample_thresholds_fused(ctx, 0, g->Zh.p + (size_t)row0 * Gpad, rows, Zs.p, Spad, Gpad * (int32_t)sizeof(__half), (int32_t)S,
                                       n_tiles1, r_sel, c_hi, c_lo, eps, thr, nullptr, g->G * (int32_t)sizeof(__half));
        return;
    }
    DevBuf<__half> Zs(ctx, ctx->gemm_impl != 1 ? (size_t)Spad * Gpad : 0);
    DevBuf<float> Zsf(ctx, ctx->gemm_impl != 1 ? 0 : (size_t)Spad * Gpad);
    {
        StageTimer t(ctx, "sample_gather");
        if (ctx->gemm_impl != 1) launch_gather_rows(ctx, g->Zh.p, (int64_t)Gpad * sizeof(__half), d_samp.p, (int32_t)Spad, Zs.p);
        else launch_gather_rows(ctx, g->Zf.p, (int64_t)Gpad * sizeof(float), d_samp.p, (int32_t)Spad, Zsf.p);
    }
    int64_t chunk = std::max<int64_t>(128, ((int64_t)2 << 30) / (Spad * (ctx->gemm_impl != 1 ? 2 : 4)) / 128 * 128);
    chunk = std::min(chunk, round_up(rows, 128));
    const bool tc = ctx->gemm_impl != 1;              // tensor-core path keeps the sampled correlations in fp16
    DevBuf<float> Cs(ctx, tc ? 0 : (size_t)chunk * Spad);
    DevBuf<__half> Cs16(ctx, tc ? (size_t)chunk * Spad : 0);
    for (int64_t c0 = 0; c0 < rows; c0 += chunk) {
        const int64_t nr = std::min(chunk, rows - c0);
        {
            StageTimer t(ctx, "cor_sample_gemm");
            if (ctx->gemm_impl != 1)
                launch_tc_store16(ctx, g->Zh.p + (size_t)(row0 + c0) * Gpad, g->Mpad - (row0 + c0), Zs.p, Spad, Gpad,
                                  Cs16.p, Spad, nr, S);
            else
                launch_simt_store(ctx, g->Zf.p + (size_t)(row0 + c0) * Gpad, nr, Zsf.p, S, Gpad, Cs.p, Spad);
        }
        {
            StageTimer t(ctx, "sample_threshold");
            if (tc) launch_sample_threshold16(ctx, Cs16.p, Spad, nr, (int32_t)S, r_sel, eps, thr + c0);
            else launch_sample_threshold(ctx, Cs.p, Spad, nr, (int32_t)S, r_sel, eps, thr + c0);
        }
    }
    // temporaries are released in stream order: no host synchronisation here
}

static void knn_filter(mcb_cgraph* g, int exchange)
{
    mcb_ctx* ctx = g->ctx;
    g->wait_planes();              // multi-rank: the lo plane arrives on the side stream
    const int64_t M = g->M, rows = g->rows, row0 = g->row0;
    const int32_t Gpad = g->Gpad, cap = g->cap, nranks = g->nranks;
    g->exchange = exchange;
    g->cand.alloc(ctx, (size_t)std::max<int64_t>(rows, 1) * cap);
    g->cnt.alloc(ctx, (size_t)std::max<int64_t>(rows, 1));
    g->cnt.zero();
    g->send_counts.assign((size_t)nranks, 0);
    if (ctx->gemm_impl == 1) {      // CUDA-core validation path: straight into the candidate lists, own rows only
        MCB_CHECK(!exchange, "the symmetric multi-GPU exchange needs the tensor-core kernels");
        g->pool_records.release();
        if (rows) {
            StageTimer t(ctx, "cor_filter_gemm");
            launch_simt_filter(ctx, g->Zf.p + (size_t)row0 * Gpad, row0, rows, g->Zf.p, M, Gpad, g->thr_all.p + row0, g->cand.p, g->cnt.p, cap);
        }
        return;
    }
    // the epilogue appends records to pool chunks; expected records ~ (rows served) x (count at the sampled threshold)
    const uint32_t crec = tc_pool_chunk_records();
    double expect = (exchange ? (double)M / nranks : (double)rows) * g->expect_per_row;
    for (int attempt = 0;; ++attempt) {
        const uint32_t n_chunks = (uint32_t)std::min<double>(4.0e6, expect * 1.3 / crec + tc_pool_min_chunks(ctx));
        g->pool_records.alloc(ctx, (size_t)n_chunks * crec);
        g->pool_fill.alloc(ctx, (size_t)n_chunks);
        g->pool_ctl.alloc(ctx, 2);
        g->pool_fill.zero(); g->pool_ctl.zero();
        g->pool_chunks = n_chunks;
        CandPool pool{g->pool_records.p, g->pool_ctl.p, g->pool_fill.p, n_chunks, g->pool_ctl.p + 1};
        {
            StageTimer t(ctx, "cor_filter_gemm");
            if (exchange) launch_tc_filter(ctx, g->Zh.p, g->Zl.p, g->Mpad, 0, M, M, Gpad, g->thr_all.p, pool, g->rank, nranks, g->G);
            else if (rows) launch_tc_filter(ctx, g->Zh.p, g->Zl.p, g->Mpad, row0, rows, M, Gpad, g->thr_all.p + row0, pool, 0, 1, g->G);
        }
        if (!exchange) {
            uint32_t hctl[2] = {0, 0};
            MCB_CUDA(cudaMemcpyAsync(hctl, g->pool_ctl.p, sizeof(hctl), cudaMemcpyDeviceToHost, ctx->stream));
            MCB_CUDA(cudaStreamSynchronize(ctx->stream));
            if (!hctl[1]) break;
        } else {
            // records per owner rank + this rank's overflow flag, all-gathered on the device: ONE host synchronisation tells
            // every rank what it sends, what it receives, and whether anyone's pool ran out (then all repeat the pass together)
            DevBuf<uint64_t> cnt_ov(ctx, (size_t)nranks + 1), cur(ctx, (size_t)nranks);
            cnt_ov.zero(); cur.zero();
            std::vector<int64_t> all;
            {
                StageTimer t(ctx, "records_partition", 1);
                launch_count_records_by_dest(ctx, pool, g->rows_per_rank, nranks, cnt_ov.p);
                MCB_CUDA(cudaMemcpyAsync(cnt_ov.p + nranks, g->pool_ctl.p + 1, sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
                all = comm_allgather_dev_i64v(g->comm, reinterpret_cast<const int64_t*>(cnt_ov.p), nranks + 1);
            }
            bool overflow = false;
            for (int r = 0; r < nranks; ++r) overflow |= all[(size_t)r * (nranks + 1) + nranks] != 0;
            if (!overflow) {
                g->recv_counts.assign((size_t)nranks, 0);
                int64_t tot = 0;
                for (int r = 0; r < nranks; ++r) {
                    g->send_counts[r] = all[(size_t)g->rank * (nranks + 1) + r];
                    g->recv_counts[r] = all[(size_t)r * (nranks + 1) + g->rank];
                    tot += g->send_counts[r];
                }
                StageTimer t(ctx, "records_partition", 1);
                g->send.alloc(ctx, (size_t)std::max<int64_t>(tot, 1));
                launch_partition_records(ctx, pool, g->rows_per_rank, nranks, cnt_ov.p, cur.p, g->send.p);
                g->pool_records.release();              // stream ordered
                break;
            }
        }
        MCB_CHECK(attempt < 3, "MC-ERR: candidate pool overflow (correlation structure far from the sampled estimate)");
        expect *= 2.0;          // the pool ran out: repeat the pass with twice the room
    }
}

static void knn_finish(mcb_cgraph* g, int64_t n_recv)
{
    mcb_ctx* ctx = g->ctx;
    const int64_t M = g->M, rows = g->rows, row0 = g->row0;
    const int32_t Gpad = g->Gpad, cap = g->cap, Kc = g->Kc, kc_out = g->kc_out;
    if (rows == 0) return;
    if (ctx->gemm_impl != 1) {
        StageTimer t(ctx, "cand_scatter");
        if (g->exchange) launch_scatter_flat(ctx, g->recv.p, n_recv, row0, g->cand.p, g->cnt.p, cap);
        else {
            CandPool pool{g->pool_records.p, g->pool_ctl.p, g->pool_fill.p, g->pool_chunks, g->pool_ctl.p + 1};
            launch_scatter_candidates(ctx, pool, g->cand.p, g->cnt.p, cap);
        }
    }
    // ---- ranked lists -------------------------------------------------------------------------
    DevBuf<int32_t> fail(ctx, (size_t)rows), fail_list(ctx, (size_t)rows), n_fail_d(ctx, 1);
    int32_t* own_col = g->knn_col.p;
    {
        StageTimer t(ctx, "topk_rows");
        launch_topk_rows(ctx, g->cand.p, g->cnt.p, cap, row0, rows, kc_out, Kc, own_col, g->knn_cor.p, fail.p);
    }
    launch_compact_flags(ctx, fail.p, rows, fail_list.p, n_fail_d.p);
    DevBuf<uint64_t> cand_total(ctx, 1);
    launch_sum_u32(ctx, g->cnt.p, rows, cand_total.p);              // statistics for DESIGN / bench: how selective the filter was
    int32_t n_fail = 0;
    uint64_t h_total = 0;
    MCB_CUDA(cudaMemcpyAsync(&n_fail, n_fail_d.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaMemcpyAsync(&h_total, cand_total.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    g->n_candidates = (int64_t)h_total;
    g->n_fallback = n_fail;
    if (n_fail > 0) {
        // rows whose candidate count missed [kc_out-1, cap]: their full correlation rows through the 3-pass
        // tensor-core kernel in STORE mode (same arithmetic as the filter kernel), then an exact global select;
        // with "gemm_impl" = 1 the CUDA-core kernel plays that role
        StageTimer t(ctx, "knn_exact_rows");
        const int32_t B = 128;
        const int64_t ldcf = round_up(M, 4);            // 16-byte aligned rows for the vector stores
        DevBuf<float> Cf(ctx, (size_t)B * ldcf);
        DevBuf<__half> Ah(ctx, ctx->gemm_impl != 1 ? (size_t)B * Gpad : 0), Al(ctx, ctx->gemm_impl != 1 ? (size_t)B * Gpad : 0);
        DevBuf<float> Af(ctx, ctx->gemm_impl != 1 ? 0 : (size_t)B * Gpad);
        DevBuf<int32_t> glist(ctx, (size_t)B);
        std::vector<int32_t> hl((size_t)n_fail), gl((size_t)B);
        MCB_CUDA(cudaMemcpyAsync(hl.data(), fail_list.p, sizeof(int32_t) * (size_t)n_fail, cudaMemcpyDeviceToHost, ctx->stream));
        MCB_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int32_t b0 = 0; b0 < n_fail; b0 += B) {
            const int32_t nb = std::min(B, n_fail - b0);
            for (int32_t q = 0; q < B; ++q) gl[q] = q < nb ? (int32_t)(row0 + hl[b0 + q]) : -1;      // -1 -> zero row
            MCB_CUDA(cudaMemcpyAsync(glist.p, gl.data(), sizeof(int32_t) * (size_t)B, cudaMemcpyHostToDevice, ctx->stream));
            if (ctx->gemm_impl != 1) {
                launch_gather_rows(ctx, g->Zh.p, (int64_t)Gpad * sizeof(__half), glist.p, B, Ah.p);
                launch_gather_rows(ctx, g->Zl.p, (int64_t)Gpad * sizeof(__half), glist.p, B, Al.p);
                launch_tc_store3_ab(ctx, Ah.p, Al.p, B, g->Zh.p, g->Zl.p, g->Mpad, nb, M, Gpad, Cf.p, ldcf);
            } else {
                launch_gather_rows(ctx, g->Zf.p, (int64_t)Gpad * sizeof(float), glist.p, B, Af.p);
                launch_simt_store(ctx, Af.p, nb, g->Zf.p, M, Gpad, Cf.p, ldcf);
            }
            launch_topk_big(ctx, Cf.p, ldcf, M, fail_list.p + b0, nb, row0, kc_out, Kc, g->thr_all.p + row0, own_col, g->knn_cor.p);
            MCB_CUDA(cudaStreamSynchronize(ctx->stream));      // gl is reused
        }
    }
    g->cand.release(); g->cnt.release(); g->pool_records.release(); g->pool_fill.release(); g->send.release(); g->recv.release();
}

// A5 for this rank's rows.  With a communicator (rank-collective build) the exchanges of the symmetric schedule run
// here, on the context's stream: all-gather of the filter thresholds, all-to-all of the candidate records (sizes
// through one small all-gather), all-gather of the ranked kNN columns.
void mcb::cgraph_knn(mcb_cgraph* g, int32_t K, int32_t k_expand, int32_t rank, int32_t nranks)
{
    MCB_CUDA(cudaSetDevice(g->ctx->device));
    mcb_comm* comm = g->comm;
    if (comm) MCB_CHECK(comm->rank == rank && comm->nranks == nranks, "mcb_cgraph_knn: rank / nranks differ from the communicator's");
    knn_begin(g, K, k_expand, rank, nranks);
    if (!comm) {
        knn_filter(g, 0);
        knn_finish(g, 0);
        return;
    }
    mcb_ctx* ctx = g->ctx;
    { StageTimer t(ctx, "allgather_thresholds", 0); comm_allgather_inplace(comm, g->thr_all.p, (size_t)g->rows_per_rank * sizeof(float)); }
    knn_filter(g, 1);
    std::vector<size_t> soff((size_t)nranks), scnt((size_t)nranks), roff((size_t)nranks), rcnt((size_t)nranks);
    size_t srun = 0, rrun = 0;
    for (int r = 0; r < nranks; ++r) {
        soff[r] = srun * sizeof(uint4); scnt[r] = (size_t)g->send_counts[r] * sizeof(uint4); srun += (size_t)g->send_counts[r];
        const size_t from_r = (size_t)g->recv_counts[r];            // exchanged in knn_filter
        roff[r] = rrun * sizeof(uint4); rcnt[r] = from_r * sizeof(uint4); rrun += from_r;
    }
    g->recv.alloc(ctx, std::max<size_t>(rrun, 1));
    { StageTimer t(ctx, "alltoall_records", 0); comm_alltoallv(comm, g->send.p, soff.data(), scnt.data(), g->recv.p, roff.data(), rcnt.data()); }
    knn_finish(g, (int64_t)rrun);
    // the ranked lists stay with their owner: the balancing join routes records, not lists (cgraph_mutual)
}

extern "C" int mcb_cgraph_knn(mcb_cgraph* g, int32_t K, int32_t k_expand, int32_t rank, int32_t nranks)
{
    MCB_API_BEGIN
    MCB_CHECK(g, "null argument");
    cgraph_knn(g, K, k_expand, rank, nranks);
    MCB_API_END
}
extern "C" int mcb_cgraph_download_knn(mcb_cgraph* g, int32_t* host_col, float* host_cor)
{
    MCB_API_BEGIN
    MCB_CHECK(g && g->knn_col.p, "mcb_cgraph_download_knn: run mcb_cgraph_knn first");
    mcb_ctx* ctx = g->ctx;
    MCB_CUDA(cudaSetDevice(ctx->device));
    const size_t n = (size_t)g->rows * g->Kc;
    if (host_col && n) MCB_CUDA(cudaMemcpyAsync(host_col, g->knn_col.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (host_cor && n) MCB_CUDA(cudaMemcpyAsync(host_cor, g->knn_cor.p, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    MCB_API_END
}
extern "C" int mcb_cgraph_knn_stats(mcb_cgraph* g, int64_t* n_fallback_rows, int64_t* n_candidates, int32_t* cap,
                                    int32_t* n_sample_cols)
{
    MCB_API_BEGIN
    MCB_CHECK(g, "null argument");
    if (n_fallback_rows) *n_fallback_rows = g->n_fallback;
    if (n_candidates) *n_candidates = g->n_candidates;
    if (cap) *cap = g->cap;
    if (n_sample_cols) *n_sample_cols = g->n_sample;
    MCB_API_END
}

static int ilog2_ceil(int64_t v) { int l = 0; while (((int64_t)1 << l) < v) ++l; return l; }

// A6 step 1 for the owned rows: mutual scores s(i,j) = rank(i,j) * rank(j,i) and the per-target in-degree cut.
// Routed join (balance.cu): every list entry becomes a record bucketed by (owner rank, row block) of its target; with a
// communicator the buckets of other ranks travel in one all-to-all (the kNN lists themselves never leave their owner and
// a rank builds rank tables for its own rows only); the probe pass then walks this rank's blocks one by one.  The
// in-degree cuts (8 bytes per row) are all-gathered at the end.  Option "join_impl": 1 = direct join (every entry probes
// the table of its target row wherever it lies in memory; single rank only), 0 = routed join, 2 = blocked-list join
// (single rank: the lists are bucketed by target block while the tables are built, probed block-major, mutual keys kept
// in compact lists), -1 (default) = blocked on a single rank, routed across ranks.  All three are bit-identical
// (tests/test_gpu_cgraph.py::test_routed_join_equals_direct_join).
void mcb::cgraph_mutual(mcb_cgraph* g, int32_t k_beta)
{
    MCB_CHECK(g->knn_col.p, "mcb_cgraph_mutual: run mcb_cgraph_knn first");
    MCB_CHECK(k_beta >= 1, "MC-ERR: k_beta must be >= 1");
    mcb_ctx* ctx = g->ctx;
    MCB_CUDA(cudaSetDevice(ctx->device));
    mcb_comm* comm = g->comm;
    MCB_CHECK(comm || g->nranks == 1, "mcb_cgraph_mutual: a row-sharded build needs a communicator (mcb_cgraph_build_sharded)");
    const int64_t M = g->M, rows = g->rows, row0 = g->row0;
    const int32_t Kc = g->Kc, P = g->nranks;
    g->k_beta = k_beta;
    g->rank_bits = std::max(1, ilog2_ceil(Kc));
    int32_t H = 2; while (H < 2 * Kc) H <<= 1;
    g->H = H;
    MCB_CHECK(M < ((int64_t)1 << (32 - g->rank_bits)) - 1, "MC-ERR: too many cells for the packed rank table at this K * k_expand");
    const int64_t smax = (int64_t)g->K * g->K * g->k_expand;
    // drop the feature planes: they are not needed after the kNN stage and the tables are large
    g->Zf.release(); g->Zh.release(); g->Zl.release();
    const size_t rows1 = (size_t)std::max<int64_t>(rows, 1);
    g->thr_in.alloc(ctx, (size_t)P * g->rows_per_rank);
    DevBuf<uint32_t> table(ctx, rows1 * H);
    const int join = P > 1 ? 0 : (ctx->join_impl < 0 ? 2 : ctx->join_impl);
    g->mut.release(); g->n_mut.release(); g->sym.release();
    if (join == 2) {
        // blocked join: lists bucketed by target block in the same pass that builds the tables, block-major probe into
        // compact lists of mutual keys, in-degree cut on the compact lists
        const BlockedShape bs = blocked_shape(rows, H);
        DevBuf<uint32_t> blist(ctx, rows1 * Kc);
        DevBuf<uint16_t> off(ctx, rows1 * (size_t)(bs.nb + 1));
        g->mut.alloc(ctx, rows1 * Kc);
        g->n_mut.alloc(ctx, rows1);
        g->n_mut.zero();
        {
            StageTimer t(ctx, "balance_hash");
            launch_prepare_blocked(ctx, g->knn_col.p, rows, Kc, H, g->rank_bits, bs, ctx->keep_self, table.p, blist.p, off.p);
        }
        {
            StageTimer t(ctx, "balance_probe");
            launch_probe_blocked(ctx, blist.p, off.p, rows, Kc, bs, table.p, H, g->rank_bits, smax, ctx->thresh_strict, g->mut.p, g->n_mut.p);
        }
        StageTimer t(ctx, "balance_cut");
        launch_mutual(ctx, g->knn_col.p, table.p, row0, rows, Kc, H, g->rank_bits, smax, ctx->thresh_strict, ctx->keep_self, g->K * k_beta,
                      nullptr, g->thr_in.p, /*have_scores=*/2, g->mut.p, g->n_mut.p);
        return;
    }
    g->sym.alloc(ctx, rows1 * Kc);
    {
        StageTimer t(ctx, "balance_hash");
        launch_build_hash(ctx, g->knn_col.p, rows, Kc, H, g->rank_bits, table.p);        // own rows only
    }
    if (join == 1) {
        StageTimer t(ctx, "balance_mutual");
        launch_mutual(ctx, g->knn_col.p, table.p, row0, rows, Kc, H, g->rank_bits, smax, ctx->thresh_strict, ctx->keep_self,
                      g->K * k_beta, g->sym.p, g->thr_in.p, /*have_scores=*/0);
        return;
    }
    const RouteShape rs = route_shape(ctx, g->rows_per_rank, P, H, g->rank_bits);
    DevBuf<uint32_t> counts(ctx, (size_t)rs.grid * rs.nbuckets);
    DevBuf<int64_t> btotal(ctx, (size_t)rs.nbuckets), bpad(ctx, (size_t)rs.nbuckets), bstart(ctx, (size_t)rs.nbuckets + 1);
    DevBuf<uint64_t> records(ctx, rows1 * Kc + (size_t)rs.nbuckets);          // every bucket padded to an even number of records
    {
        StageTimer t(ctx, "balance_route", 5);
        launch_route_records(ctx, g->knn_col.p, row0, rows, Kc, g->rows_per_rank, rs, g->rank_bits, ctx->keep_self, counts.p, btotal.p,
                             bpad.p, bstart.p, records.p);
    }
    g->sym.zero();
    const int32_t nbr = rs.nbr;
    if (!comm) {
        // one rank: the buckets are the blocks, already in visiting order
        DevBuf<int32_t> seg_block(ctx, (size_t)nbr);
        launch_iota_i32(ctx, seg_block.p, nbr);
        StageTimer t(ctx, "balance_probe");
        launch_probe(ctx, records.p, (int64_t)rows * Kc + rs.nbuckets, bstart.p, bstart.p, seg_block.p, nbr, rs.blk_shift, table.p, H, g->rank_bits, Kc,
                     smax, ctx->thresh_strict, g->sym.p);
        // (n_records is an upper bound: the prefix ends at the true count, positions past it fall in no segment)
    } else {
        // bucket totals of every rank -> who sends how much to whom, and where each (sender, block) piece lands here
        const std::vector<int64_t> all = comm_allgather_dev_i64v(comm, btotal.p, rs.nbuckets);         // [sender][bucket], one host sync
        const int me = comm->rank;
        const int64_t* mine = all.data() + (size_t)me * rs.nbuckets;
        std::vector<size_t> soff((size_t)P), scnt((size_t)P), roff((size_t)P), rcnt((size_t)P);
        size_t srun = 0, rrun = 0;
        auto even = [](int64_t c) { return (size_t)((c + 1) & ~(int64_t)1); };            // bucket strides (launch_route_records)
        for (int r = 0; r < P; ++r) {
            size_t to_r = 0, from_r = 0;
            for (int b = 0; b < nbr; ++b) { to_r += even(mine[(size_t)r * nbr + b]); from_r += even(all[(size_t)r * rs.nbuckets + (size_t)me * nbr + b]); }
            soff[r] = srun * 8; scnt[r] = to_r * 8; srun += to_r;
            roff[r] = rrun * 8; rcnt[r] = from_r * 8; rrun += from_r;
        }
        DevBuf<uint64_t> recv(ctx, std::max<size_t>(rrun, 1));
        { StageTimer t(ctx, "alltoall_join_records", 0); comm_alltoallv(comm, records.p, soff.data(), scnt.data(), recv.p, roff.data(), rcnt.data()); }
        records.release();
        // visiting order: block-major, the P senders' pieces of a block one after the other
        const int32_t nseg = nbr * P;
        std::vector<int64_t> h_prefix((size_t)nseg + 1), h_src((size_t)nseg);
        std::vector<int32_t> h_blk((size_t)nseg);
        std::vector<size_t> piece((size_t)P);
        for (int r = 0; r < P; ++r) piece[r] = roff[r] / 8;
        int64_t run = 0;
        // pieces of sender r are laid out by ascending block inside its contribution: walk the blocks and advance each sender's cursor
        for (int b = 0; b < nbr; ++b)
            for (int r = 0; r < P; ++r) {
                const int q = b * P + r;
                const int64_t c = all[(size_t)r * rs.nbuckets + (size_t)me * nbr + b];
                h_prefix[q] = run; h_src[q] = (int64_t)piece[r]; h_blk[q] = b;
                run += c; piece[r] += even(c);
            }
        h_prefix[nseg] = run;
        DevBuf<int64_t> seg_prefix(ctx, (size_t)nseg + 1), seg_src(ctx, (size_t)nseg);
        DevBuf<int32_t> seg_block(ctx, (size_t)nseg);
        MCB_CUDA(cudaMemcpyAsync(seg_prefix.p, h_prefix.data(), sizeof(int64_t) * h_prefix.size(), cudaMemcpyHostToDevice, ctx->stream));
        MCB_CUDA(cudaMemcpyAsync(seg_src.p, h_src.data(), sizeof(int64_t) * h_src.size(), cudaMemcpyHostToDevice, ctx->stream));
        MCB_CUDA(cudaMemcpyAsync(seg_block.p, h_blk.data(), sizeof(int32_t) * h_blk.size(), cudaMemcpyHostToDevice, ctx->stream));
        {
            StageTimer t(ctx, "balance_probe");
            launch_probe(ctx, recv.p, run, seg_prefix.p, seg_src.p, seg_block.p, nseg, rs.blk_shift, table.p, H, g->rank_bits, Kc, smax,
                         ctx->thresh_strict, g->sym.p);
        }
        // (the copies above come from pageable host vectors: cudaMemcpyAsync stages such a source before it returns)
    }
    {
        StageTimer t(ctx, "balance_cut");
        launch_mutual(ctx, g->knn_col.p, table.p, row0, rows, Kc, H, g->rank_bits, smax, ctx->thresh_strict, ctx->keep_self, g->K * k_beta,
                      g->sym.p, g->thr_in.p, /*have_scores=*/1);
    }
    if (comm) { StageTimer t(ctx, "allgather_cuts", 0); comm_allgather_inplace(comm, g->thr_in.p, (size_t)g->rows_per_rank * sizeof(uint64_t)); }
    // temporaries are released in stream order
}
extern "C" int mcb_cgraph_mutual(mcb_cgraph* g, int32_t k_beta)
{
    MCB_API_BEGIN
    MCB_CHECK(g, "null argument");
    cgraph_mutual(g, k_beta);
    MCB_API_END
}
int64_t mcb::cgraph_prune(mcb_cgraph* g)
{
    MCB_CHECK(g->sym.p || g->mut.p, "mcb_cgraph_prune: run mcb_cgraph_mutual first");
    mcb_ctx* ctx = g->ctx;
    MCB_CUDA(cudaSetDevice(ctx->device));
    const int64_t rows = g->rows;
    const int32_t K = g->K;
    const size_t rk = (size_t)std::max<int64_t>(rows, 1) * K;
    g->out_deg.alloc(ctx, (size_t)std::max<int64_t>(rows, 1));
    g->out_dst.alloc(ctx, rk);
    g->out_s.alloc(ctx, rk);
    {
        StageTimer t(ctx, "balance_prune");
        launch_prune(ctx, g->knn_col.p, g->sym.p, g->thr_in.p, g->row0, rows, g->Kc, K, (int64_t)g->K * g->K * g->k_expand, g->out_deg.p, g->out_dst.p, g->out_s.p,
                     g->mut.p, g->n_mut.p);
    }
    DevBuf<int64_t> deg64(ctx, (size_t)std::max<int64_t>(rows, 1)), offs(ctx, (size_t)std::max<int64_t>(rows, 1)), total(ctx, 1);
    int64_t ne = 0;
    if (rows) {
        StageTimer t(ctx, "edge_scan", 2);
        launch_widen_i32(ctx, g->out_deg.p, deg64.p, rows);
        launch_exclusive_scan_i64(ctx, deg64.p, offs.p, rows, total.p);
        MCB_CUDA(cudaMemcpyAsync(&ne, total.p, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    g->n_edges = ne;
    g->e_src.alloc(ctx, (size_t)std::max<int64_t>(ne, 1));
    g->e_dst.alloc(ctx, (size_t)std::max<int64_t>(ne, 1));
    g->e_w.alloc(ctx, (size_t)std::max<int64_t>(ne, 1));
    if (rows) {
        StageTimer t(ctx, "edge_emit");
        launch_emit_edges(ctx, g->out_deg.p, g->out_dst.p, offs.p, g->cell_of_row.p, g->row0, rows, K, g->e_src.p, g->e_dst.p, g->e_w.p);
    }
    return ne;
}
extern "C" int mcb_cgraph_prune(mcb_cgraph* g, int64_t* n_edges)
{
    MCB_API_BEGIN
    MCB_CHECK(g, "null argument");
    const int64_t ne = cgraph_prune(g);
    MCB_CUDA(cudaStreamSynchronize(g->ctx->stream));
    if (n_edges) *n_edges = ne;
    MCB_API_END
}
// The edge table of the owned rows -> host.  The device sends the targets and the out-degrees; the host threads of the
// staging ring rebuild mc1 and w = 1 - place/K while the targets are in flight (staging.cu).
void mcb::cgraph_edges_to_host(mcb_cgraph* g, int32_t* host_src, int32_t* host_dst, double* host_w)
{
    MCB_CHECK(g->e_src.p, "mcb_cgraph_edges: run mcb_cgraph_prune first");
    mcb_ctx* ctx = g->ctx;
    MCB_CUDA(cudaSetDevice(ctx->device));
    StageTimer t(ctx, "d2h_edges", 0);
    staged_download_edges(ctx, g->rows, g->K, g->n_edges, g->out_deg.p, g->cell_of_row.p + g->row0, g->e_dst.p, host_src, host_dst, host_w);
}
extern "C" int mcb_cgraph_edges(mcb_cgraph* g, int32_t* host_src, int32_t* host_dst, double* host_w)
{
    MCB_API_BEGIN
    MCB_CHECK(g, "null argument");
    cgraph_edges_to_host(g, host_src, host_dst, host_w);
    MCB_API_END
}
// the same table left on the device (e.g. for a caller that keeps working there): pointers valid until the handle dies
extern "C" int mcb_cgraph_edges_device(mcb_cgraph* g, int64_t* n_edges, const int32_t** dev_src, const int32_t** dev_dst, const double** dev_w)
{
    MCB_API_BEGIN
    MCB_CHECK(g && g->e_src.p, "mcb_cgraph_edges_device: run mcb_cgraph_prune first");
    if (n_edges) *n_edges = g->n_edges;
    if (dev_src) *dev_src = g->e_src.p;
    if (dev_dst) *dev_dst = g->e_dst.p;
    if (dev_w) *dev_w = g->e_w.p;
    MCB_API_END
}

// test hook: dense correlation block through the tensor-core (impl 0, 3-pass; impl 2, 1-pass) or
// SIMT (impl 1) kernel: C[a_rows][b_rows] = Z[0:a_rows] . Z[0:b_rows]^T
extern "C" int mcb_cgraph_debug_cor(mcb_cgraph* g, int impl, int64_t a_rows, int64_t b_rows, float* host_c)
{
    MCB_API_BEGIN
    MCB_CHECK(g && g->Zf.p && host_c, "mcb_cgraph_debug_cor: features not available");
    MCB_CHECK(a_rows <= g->M && b_rows <= g->M, "mcb_cgraph_debug_cor: block exceeds matrix");
    mcb_ctx* ctx = g->ctx;
    MCB_CUDA(cudaSetDevice(ctx->device));
    const int64_t ldc = round_up(b_rows, 4);
    DevBuf<float> C(ctx, (size_t)a_rows * ldc);
    C.zero();
    if (impl == 0) launch_tc_store3(ctx, g->Zh.p, g->Zl.p, g->Mpad, a_rows, b_rows, g->Gpad, C.p, ldc);
    else if (impl == 2) launch_tc_store(ctx, g->Zh.p, g->Mpad, g->Zh.p, g->Mpad, g->Gpad, C.p, ldc, a_rows, b_rows);
    else launch_simt_store(ctx, g->Zf.p, a_rows, g->Zf.p, b_rows, g->Gpad, C.p, ldc);
    std::vector<float> h((size_t)a_rows * ldc);
    MCB_CUDA(cudaMemcpyAsync(h.data(), C.p, h.size() * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int64_t r = 0; r < a_rows; ++r) memcpy(host_c + r * b_rows, h.data() + r * ldc, sizeof(float) * (size_t)b_rows);
    MCB_API_END
}

// tgs_cor_knn(x, y, knn) with y != x (reference call sites R/atlas.r:164-166,277,475-478,584):
// for every cell of x the knn cells of y with the highest Pearson correlation, ranked.  Both inputs are
// dense, one contiguous vector of G doubles per cell.  The full correlation rows go through the 3-pass
// tensor-core kernel in STORE mode in row chunks and an exact per-row select; there is no self match.
// Cells with a constant feature vector have NA correlations in R: a constant x cell gets col = -1 / cor = NaN,
// a constant y cell is never listed.
extern "C" int mcb_cor_knn_xy(mcb_ctx* ctx, const double* x, int64_t nx, const double* y, int64_t ny, int32_t G,
                              int32_t knn, int32_t* host_col, float* host_cor)
{
    MCB_API_BEGIN
    MCB_CHECK(ctx && x && y && host_col && host_cor, "mcb_cor_knn_xy: null argument");
    MCB_CHECK(knn >= 1, "MC-ERR: knn must be >= 1");
    MCB_CHECK(knn <= 4096, "MC-ERR: knn above 4096 is not supported");
    mcb_cgraph *gx = nullptr, *gy = nullptr;
    struct Guard { mcb_cgraph*& a; mcb_cgraph*& b; ~Guard() { mcb_cgraph_destroy(a); mcb_cgraph_destroy(b); } } guard{gx, gy};
    if (mcb_cgraph_create_dense(ctx, x, G, nx, &gx) != 0) throw std::runtime_error(mcb_last_error());
    if (mcb_cgraph_create_dense(ctx, y, G, ny, &gy) != 0) throw std::runtime_error(mcb_last_error());
    const int64_t Mx = gx->M, My = gy->M;
    const int32_t Gpad = gx->Gpad;
    const int32_t kk = (int32_t)std::min<int64_t>(knn, My);
    for (int64_t q = 0; q < nx * (int64_t)knn; ++q) { host_col[q] = -1; host_cor[q] = std::nanf(""); }
    if (Mx == 0 || My == 0) return 0;
    const int64_t ldcf = round_up(My, 4);
    int64_t B = std::min<int64_t>(4096, std::max<int64_t>(128, ((int64_t)1 << 28) / ldcf / 128 * 128));
    B = std::min<int64_t>(B, round_up(Mx, 128));
    DevBuf<float> Cf(ctx, (size_t)B * ldcf);
    DevBuf<int32_t> rows(ctx, (size_t)B), col(ctx, (size_t)Mx * knn);
    DevBuf<float> cor(ctx, (size_t)Mx * knn);
    std::vector<int32_t> iota((size_t)B);
    for (int64_t b0 = 0; b0 < Mx; b0 += B) {
        const int64_t nb = std::min(B, Mx - b0);
        for (int64_t q = 0; q < nb; ++q) iota[q] = (int32_t)(b0 + q);
        MCB_CUDA(cudaMemcpyAsync(rows.p, iota.data(), sizeof(int32_t) * (size_t)nb, cudaMemcpyHostToDevice, ctx->stream));
        {
            StageTimer t(ctx, "cor_xy_gemm");
            launch_tc_store3_ab(ctx, gx->Zh.p + (size_t)b0 * Gpad, gx->Zl.p + (size_t)b0 * Gpad, gx->Mpad - b0, gy->Zh.p, gy->Zl.p,
                                gy->Mpad, nb, My, Gpad, Cf.p, ldcf);
        }
        {
            StageTimer t(ctx, "topk_xy_rows");
            launch_topk_big(ctx, Cf.p, ldcf, My, rows.p, (int32_t)nb, 0, kk, knn, nullptr, col.p, cor.p, false);
        }
        MCB_CUDA(cudaStreamSynchronize(ctx->stream));      // iota is reused
    }
    std::vector<int32_t> hcol((size_t)Mx * knn), cx((size_t)Mx), cy((size_t)My);
    std::vector<float> hcor((size_t)Mx * knn);
    MCB_CUDA(cudaMemcpyAsync(hcol.data(), col.p, sizeof(int32_t) * hcol.size(), cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaMemcpyAsync(hcor.data(), cor.p, sizeof(float) * hcor.size(), cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaMemcpyAsync(cx.data(), gx->cell_of_row.p, sizeof(int32_t) * cx.size(), cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaMemcpyAsync(cy.data(), gy->cell_of_row.p, sizeof(int32_t) * cy.size(), cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int64_t r = 0; r < Mx; ++r) {
        int32_t* oc = host_col + (int64_t)cx[r] * knn;
        float* ov = host_cor + (int64_t)cx[r] * knn;
        for (int32_t t = 0; t < kk; ++t) { oc[t] = cy[hcol[r * knn + t]]; ov[t] = hcor[r * knn + t]; }
    }
    MCB_API_END
}

// ---------------------------------------------------------------------------------------------------------------
// the whole body of mcell_add_cgraph_from_mat_bknn (R/cgraph_knn.r:10-25) for ONE rank: upload of the rank's cell range,
// [depth rule + downsample + add-back], features, kNN, balancing.  comm == null: the only rank.
// ---------------------------------------------------------------------------------------------------------------
mcb_cgraph* mcb::bknn_rank(mcb_ctx* ctx, mcb_comm* comm, const BknnJob& j)
{
    MCB_CHECK(j.p && j.feat_rows, "mcb_cgraph_bknn: null argument");
    MCB_CHECK(j.cell0 >= 0 && j.cell0 <= j.cell1 && j.cell1 <= j.ncells, "mcb_cgraph_bknn: bad cell range");
    std::unique_ptr<mcb_csc> m(csc_upload_impl(ctx, j.ngenes, j.cell1 - j.cell0, j.p + j.cell0, j.i, j.x));
    return bknn_rank_uploaded(ctx, comm, j, m.get());
}
// the same with the rank's cell range already on the device (`m` is only read; the caller frees it after the return)
mcb_cgraph* mcb::bknn_rank_uploaded(mcb_ctx* ctx, mcb_comm* comm, const BknnJob& j, const mcb_csc* m)
{
    const int rank = comm ? comm->rank : 0, nranks = comm ? comm->nranks : 1;
    std::unique_ptr<mcb_csc> d;
    const mcb_csc* use = m;
    if (j.dsamp) { d.reset(csc_downsample_auto(ctx, comm, m, j.cell0, j.seed, nullptr)); use = d.get(); }
    std::unique_ptr<mcb_cgraph, void (*)(mcb_cgraph*)> g(cgraph_features_csc(ctx, comm, use, j.cell0, j.feat_rows, j.G, j.k_nonz),
                                                          [](mcb_cgraph* q) { mcb_cgraph_destroy(q); });
    cgraph_knn(g.get(), j.K, j.k_expand, rank, nranks);
    cgraph_mutual(g.get(), j.k_beta);
    cgraph_prune(g.get());
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));       // the CSC handles die here
    return g.release();
}

// one shot, single GPU, host in / host out: what mcell_add_cgraph_from_mat_bknn's body becomes
extern "C" int mcb_cgraph_bknn(mcb_ctx* ctx, int32_t ngenes, int64_t ncells, const int64_t* p, const int32_t* i,
                               const double* x, const int32_t* feat_rows, int32_t G, int32_t K, int32_t k_expand,
                               int32_t k_beta, double k_nonz, int dsamp, uint64_t seed, int64_t* n_edges,
                               int32_t** src, int32_t** dst, double** w, int64_t* n_valid_cells)
{
    mcb_cgraph* g = nullptr;
    int32_t* hs = nullptr; int32_t* hd = nullptr; double* hw = nullptr;
    int rc = 0;
    try {
        MCB_CHECK(ctx && n_edges && src && dst && w, "mcb_cgraph_bknn: null argument");
        MCB_CHECK(p && p[0] == 0, "MC-ERR: mcb_cgraph_bknn: malformed column pointer");
        BknnJob j{ngenes, ncells, 0, ncells, p, i, x, feat_rows, G, K, k_expand, k_beta, k_nonz, dsamp, seed};
        g = bknn_rank(ctx, nullptr, j);
        const int64_t ne = g->n_edges;
        hs = (int32_t*)host_alloc(sizeof(int32_t) * (size_t)std::max<int64_t>(ne, 1));
        hd = (int32_t*)host_alloc(sizeof(int32_t) * (size_t)std::max<int64_t>(ne, 1));
        hw = (double*)host_alloc(sizeof(double) * (size_t)std::max<int64_t>(ne, 1));
        cgraph_edges_to_host(g, hs, hd, hw);
        *n_edges = ne; *src = hs; *dst = hd; *w = hw;
        if (n_valid_cells) *n_valid_cells = g->M;
        hs = nullptr; hd = nullptr; hw = nullptr;
    }
    catch (const mcb::Error& e) { mcb::set_error(e.what()); rc = e.code; }
    catch (const std::bad_alloc&) { mcb::set_error("host out of memory"); rc = 3; }
    catch (const std::exception& e) { mcb::set_error(e.what()); rc = 4; }
    catch (...) { mcb::set_error("unknown error"); rc = 5; }
    mcb_free(hs); mcb_free(hd); mcb_free(hw);
    mcb_cgraph_destroy(g);
    return rc;
}

// ---- rank-collective form: every rank of the communicator calls these with its own cell range ---------------------
// `local` = this rank's cells [cell_base, cell_base + local->ncells) of the whole matrix (ranks hold ascending contiguous
// ranges in rank order).  Runs A1..A7; on return this rank holds the edges of its share of the valid rows.
extern "C" int mcb_cgraph_build_sharded(mcb_comm* comm, const mcb_csc* local, int64_t cell_base, const int32_t* feat_rows, int32_t G,
                                        double k_nonz, int32_t K, int32_t k_expand, int32_t k_beta, int dsamp, uint64_t seed,
                                        mcb_cgraph** out, int64_t* n_edges_local)
{
    MCB_API_BEGIN
    MCB_CHECK(comm && local && feat_rows && out, "mcb_cgraph_build_sharded: null argument");
    MCB_CHECK(local->ctx == comm->ctx, "mcb_cgraph_build_sharded: matrix and communicator live on different contexts");
    mcb_ctx* ctx = comm->ctx;
    std::unique_ptr<mcb_csc> d;
    const mcb_csc* use = local;
    if (dsamp) { d.reset(csc_downsample_auto(ctx, comm, local, cell_base, seed, nullptr)); use = d.get(); }
    std::unique_ptr<mcb_cgraph, void (*)(mcb_cgraph*)> g(cgraph_features_csc(ctx, comm, use, cell_base, feat_rows, G, k_nonz),
                                                          [](mcb_cgraph* q) { mcb_cgraph_destroy(q); });
    cgraph_knn(g.get(), K, k_expand, comm->rank, comm->nranks);
    cgraph_mutual(g.get(), k_beta);
    const int64_t ne = cgraph_prune(g.get());
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (n_edges_local) *n_edges_local = ne;
    *out = g.release();
    MCB_API_END
}
// every rank's edges on `root`'s host, in rank (= source) order.  Collective.  The compact form (targets + out-degrees +
// source ids) travels over NVLink and PCIe; root's staging threads rebuild (mc1, mc2, w).  Outputs are library-allocated
// on root (mcb_free) and null elsewhere; *n_edges_total is set on every rank.
extern "C" int mcb_cgraph_gather_edges(mcb_comm* comm, mcb_cgraph* g, int32_t root, int64_t* n_edges_total, int32_t** src,
                                       int32_t** dst, double** w)
{
    int32_t* hs = nullptr; int32_t* hd = nullptr; double* hw = nullptr;
    int rc = 0;
    try {
        MCB_CHECK(comm && g && g->e_src.p, "mcb_cgraph_gather_edges: run the build first");
        MCB_CHECK(root >= 0 && root < comm->nranks, "mcb_cgraph_gather_edges: bad root");
        mcb_ctx* ctx = g->ctx;
        MCB_CUDA(cudaSetDevice(ctx->device));
        const int P = comm->nranks;
        const int64_t mine[2] = {g->n_edges, g->rows};
        const std::vector<int64_t> all = comm_allgather_i64v(comm, mine, 2);
        int64_t ne = 0, nr = 0;
        std::vector<size_t> eoff((size_t)P), ecnt((size_t)P), roff((size_t)P), rcnt((size_t)P);
        for (int r = 0; r < P; ++r) {
            eoff[r] = (size_t)ne * 4; ecnt[r] = (size_t)all[2 * r] * 4; ne += all[2 * r];
            roff[r] = (size_t)nr * 4; rcnt[r] = (size_t)all[2 * r + 1] * 4; nr += all[2 * r + 1];
        }
        const bool is_root = comm->rank == root;
        DevBuf<int32_t> dst_all(ctx, is_root ? (size_t)std::max<int64_t>(ne, 1) : 0), deg_all(ctx, is_root ? (size_t)std::max<int64_t>(nr, 1) : 0),
            cell_all(ctx, is_root ? (size_t)std::max<int64_t>(nr, 1) : 0);
        {
            StageTimer t(ctx, "gather_edges", 0);
            comm_gatherv(comm, g->e_dst.p, (size_t)g->n_edges * 4, dst_all.p, eoff.data(), ecnt.data(), root);
            comm_gatherv(comm, g->out_deg.p, (size_t)g->rows * 4, deg_all.p, roff.data(), rcnt.data(), root);
            comm_gatherv(comm, g->cell_of_row.p + g->row0, (size_t)g->rows * 4, cell_all.p, roff.data(), rcnt.data(), root);
        }
        if (n_edges_total) *n_edges_total = ne;
        if (is_root) {
            MCB_CHECK(src && dst && w, "mcb_cgraph_gather_edges: null output on root");
            hs = (int32_t*)host_alloc(sizeof(int32_t) * (size_t)std::max<int64_t>(ne, 1));
            hd = (int32_t*)host_alloc(sizeof(int32_t) * (size_t)std::max<int64_t>(ne, 1));
            hw = (double*)host_alloc(sizeof(double) * (size_t)std::max<int64_t>(ne, 1));
                StageTimer t(ctx, "d2h_edges", 0);
            staged_download_edges(ctx, nr, g->K, ne, deg_all.p, cell_all.p, dst_all.p, hs, hd, hw);
            *src = hs; *dst = hd; *w = hw;
            hs = nullptr; hd = nullptr; hw = nullptr;
        } else {
            if (src) *src = nullptr;
            if (dst) *dst = nullptr;
            if (w) *w = nullptr;
        }
        MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    catch (const mcb::Error& e) { mcb::set_error(e.what()); rc = e.code; }
    catch (const std::bad_alloc&) { mcb::set_error("host out of memory"); rc = 3; }
    catch (const std::exception& e) { mcb::set_error(e.what()); rc = 4; }
    catch (...) { mcb::set_error("unknown error"); rc = 5; }
    mcb_free(hs); mcb_free(hd); mcb_free(hw);
    return rc;
}

// ---------------------------------------------------------------------------------------------
// resampled co-clustering
// ---------------------------------------------------------------------------------------------
// the host's edge table -> device CSR (out and in, 16-bit fixed-point weights, packed node / edge descriptors) + endpoint
// validation; shared by the resampled co-clustering and the single cover
struct CoverCsr {
    DevBuf<int32_t> src, dst; DevBuf<double> w;
    DevBuf<int32_t> odeg, ideg, optr32, iptr32; DevBuf<int64_t> optr, iptr;
    DevBuf<int32_t> odst, isrc, scratch; DevBuf<uint32_t> ow, iw; DevBuf<int2> oe, ie; DevBuf<int4> nd;
    DevBuf<uint16_t> ipos;          // position of every in edge inside its source's out segment (edge truncation, `knn`)
    int32_t max_out = 0, max_in = 0;
};
static void cover_csr_build(mcb_ctx* ctx, int32_t N, int64_t n_edges, const int32_t* src, const int32_t* dst, const double* w, bool sort_out,
                            CoverCsr& c)
{
    const size_t ne1 = (size_t)std::max<int64_t>(n_edges, 1);
    c.src.alloc(ctx, ne1); c.dst.alloc(ctx, ne1); c.w.alloc(ctx, ne1);
    c.odeg.alloc(ctx, (size_t)N + 1); c.ideg.alloc(ctx, (size_t)N + 1); c.optr32.alloc(ctx, (size_t)N + 1); c.iptr32.alloc(ctx, (size_t)N + 1);
    c.optr.alloc(ctx, (size_t)N + 1); c.iptr.alloc(ctx, (size_t)N + 1);
    c.odst.alloc(ctx, ne1); c.isrc.alloc(ctx, ne1); c.scratch.alloc(ctx, 4);
    c.ow.alloc(ctx, ne1); c.iw.alloc(ctx, ne1); c.oe.alloc(ctx, ne1); c.ie.alloc(ctx, ne1); c.nd.alloc(ctx, (size_t)N); c.ipos.alloc(ctx, ne1);
    {
        StageTimer t(ctx, "h2d_graph");
        if (n_edges) {
            staged_copy_h2d(ctx, c.src.p, src, (size_t)n_edges * 4);
            staged_copy_h2d(ctx, c.dst.p, dst, (size_t)n_edges * 4);
            staged_copy_h2d(ctx, c.w.p, w, (size_t)n_edges * 8);
        }
    }
    {
        StageTimer t(ctx, "graph_csr", 8);
        c.odeg.zero(); c.ideg.zero(); c.scratch.zero();
        launch_build_cover_csr(ctx, N, n_edges, c.src.p, c.dst.p, c.w.p, c.odeg.p, c.ideg.p, c.optr32.p, c.iptr32.p, c.odst.p, c.ow.p,
                               c.isrc.p, c.iw.p, c.oe.p, c.ie.p, c.nd.p, c.ipos.p, sort_out, c.scratch.p);
        launch_widen_i32(ctx, c.optr32.p, c.optr.p, (int64_t)N + 1);
        launch_widen_i32(ctx, c.iptr32.p, c.iptr.p, (int64_t)N + 1);
    }
}

// n_resamp covers on host-supplied node samples.  with_counts: accumulate the N x N co-occurrence counters
// (tgs_graph_cover_resample, method = "full"); without: assignments only (tgs_graph_cover).
static mcb_coclust* coclust_run_impl(mcb_ctx* ctx, int32_t N, int64_t n_edges, const int32_t* src, const int32_t* dst, const double* w,
                                     int32_t n_resamp, int32_t n_s, const int32_t* samples, const uint64_t* seeds, int32_t min_mc_size,
                                     double cooling, int32_t burn_in, int32_t max_sweeps, int32_t knn, int32_t rank, int32_t nranks,
                                     int count_mode /* 0 none (single cover), 1 method "full", 2 method "edges" */)
{
    const bool with_counts = count_mode == 1;
    MCB_CHECK(N >= 1 && n_edges >= 0 && n_resamp >= 0 && n_s >= 1 && n_s <= N, "MC-ERR: bad co-clustering dimensions");
    MCB_CHECK(n_edges == 0 || (src && dst && w), "mcb_coclust_run: null edge arrays");
    MCB_CHECK(n_resamp == 0 || seeds, "mcb_coclust_run: null seed array");
    MCB_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
    MCB_CHECK(knn >= 0, "MC-ERR: knn must be >= 0 (0 = no per-node edge limit)");
    MCB_CHECK(!with_counts || (int64_t)N * N < ((int64_t)1 << 36), "MC-ERR: method=\"full\" needs an N x N counter matrix; N is too large");
    MCB_CUDA(cudaSetDevice(ctx->device));
    MCB_CHECK(n_edges < ((int64_t)1 << 31), "MC-ERR: too many edges");
    std::vector<int32_t> ids;
    for (int32_t r = rank; r < n_resamp; r += nranks) ids.push_back(r);
    std::unique_ptr<mcb_coclust> c(new mcb_coclust());
    c->ref.set(ctx); c->ctx = ctx; c->N = N; c->n_local = (int32_t)ids.size();
    c->ncl_cap = n_s / (min_mc_size > 0 ? min_mc_size + 1 : 1) + 2;
    const int32_t nl = c->n_local;

    CoverCsr csr;
    cover_csr_build(ctx, N, n_edges, src, dst, w, /*sort_out=*/knn > 0 || count_mode == 2, csr);
    DevBuf<int32_t> d_samples(ctx, (size_t)std::max<int64_t>((int64_t)n_resamp * n_s, 1)), d_ids(ctx, (size_t)std::max(nl, 1));
    DevBuf<uint64_t> d_seeds(ctx, (size_t)std::max(n_resamp, 1));
    if (n_resamp) {
        if (samples) staged_copy_h2d(ctx, d_samples.p, samples, (size_t)n_resamp * n_s * 4);
        else { MCB_CHECK(n_s == N && n_resamp == 1, "mcb_coclust_run: null sample array"); launch_iota_i32(ctx, d_samples.p, N); }   // all nodes
        MCB_CUDA(cudaMemcpyAsync(d_seeds.p, seeds, (size_t)n_resamp * 8, cudaMemcpyHostToDevice, ctx->stream));
        launch_sample_check(ctx, (int64_t)n_resamp * n_s, n_s, d_samples.p, N, csr.scratch.p + 1);
    }
    if (nl) MCB_CUDA(cudaMemcpyAsync(d_ids.p, ids.data(), (size_t)nl * 4, cudaMemcpyHostToDevice, ctx->stream));
    int32_t hs[4] = {0, 0, 0, 0};
    MCB_CUDA(cudaMemcpyAsync(hs, csr.scratch.p, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    MCB_CHECK(!(hs[1] & 1), "MC-ERR: edge endpoint out of range");
    MCB_CHECK(!(hs[1] & 2), "MC-ERR: sampled node out of range");
    MCB_CHECK(!(hs[1] & 4), "MC-ERR: each resample index set must be strictly ascending");
    MCB_CHECK(!(hs[1] & 16), "MC-ERR: the edge table lists a (node1, node2) pair more than once");
    MCB_CHECK(!(hs[1] & 8), "MC-ERR: tgs_graph_cover_resample: knn and method = \"edges\" order every node's out edges, which supports out-degrees up to 1024");
    const int32_t max_out = hs[2], max_in = hs[3];
    // `knn` = the most out edges a node may use inside one resample (R/coclust_graph_cov.r:39-41 passes the graph's AVERAGE
    // out-degree, so well-connected nodes lose their weakest sampled edges).  It can only bind below the largest out-degree.
    const int32_t knn_eff = (knn > 0 && knn < max_out) ? knn : 0;
    MCB_CHECK(knn_eff == 0 || max_out < 65536, "MC-ERR: out-degree too large for per-node edge truncation");
    DevBuf<uint16_t> ocut(ctx, knn_eff ? (size_t)std::max(nl, 1) * N : 0);
    if (with_counts) {
        c->cnt.alloc(ctx, (size_t)N * N); c->cnt.zero();
        c->nsamp.alloc(ctx, (size_t)N); c->nsamp.zero();
    }
    c->mc.alloc(ctx, (size_t)std::max(nl, 1) * N);
    c->sweeps.alloc(ctx, (size_t)std::max(nl, 1));
    DevBuf<int32_t> f_ws(ctx, (size_t)std::max(nl, 1) * N), status(ctx, (size_t)std::max(nl, 1));
    CoverGraph cg{N, csr.optr.p, csr.odst.p, csr.ow.p, csr.iptr.p, csr.isrc.p, csr.iw.p};
    {
        StageTimer t(ctx, "graph_cover");
        launch_graph_cover(ctx, cg, csr.optr32.p, csr.iptr32.p, csr.nd.p, csr.oe.p, csr.ie.p, max_out, max_in, nl, d_ids.p, n_s, d_samples.p, d_seeds.p,
                           min_mc_size, cooling, burn_in, max_sweeps, c->ncl_cap, c->mc.p, f_ws.p, c->sweeps.p, status.p, knn_eff, csr.ipos.p, ocut.p);
    }
    if (with_counts) {
        DevBuf<int32_t> ws_order(ctx, (size_t)std::max(nl, 1) * N), ws_start(ctx, (size_t)std::max(nl, 1) * (c->ncl_cap + 1));
        StageTimer t(ctx, "cooccur");
        launch_cooccur(ctx, N, nl, c->mc.p, c->ncl_cap, c->cnt.p, c->nsamp.p, ws_order.p, ws_start.p);
    }
    if (count_mode == 2) {
        // one counter per edge (out-CSR order) + the edge behind it; O(edges) memory
        MCB_CHECK(!(hs[1] & 8), "MC-ERR: method = \"edges\" supports out-degrees up to 1024");
        c->mode = 1; c->n_edges = n_edges;
        const size_t ne1 = (size_t)std::max<int64_t>(n_edges, 1);
        c->cnt.alloc(ctx, ne1); c->cnt.zero();
        c->nsamp.alloc(ctx, (size_t)N); c->nsamp.zero();
        c->e_src.alloc(ctx, ne1); c->e_dst.alloc(ctx, ne1);
        StageTimer t(ctx, "cooccur", 2);
        launch_expand_src(ctx, N, csr.optr32.p, c->e_src.p);
        if (n_edges) MCB_CUDA(cudaMemcpyAsync(c->e_dst.p, csr.odst.p, (size_t)n_edges * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        launch_cooccur_edges(ctx, N, nl, c->mc.p, csr.optr32.p, csr.odst.p, c->cnt.p, c->nsamp.p);
    }
    std::vector<int32_t> hst((size_t)std::max(nl, 1), 0);
    if (nl) MCB_CUDA(cudaMemcpyAsync(hst.data(), status.p, (size_t)nl * 4, cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int32_t q = 0; q < nl; ++q) MCB_CHECK(hst[q] == 0, "MC-ERR: graph cover ran out of cluster slots");
    return c.release();
}

extern "C" int mcb_coclust_run2(mcb_ctx* ctx, int32_t N, int64_t n_edges, const int32_t* src, const int32_t* dst,
                                const double* w, int32_t n_resamp, int32_t n_s, const int32_t* samples,
                                const uint64_t* seeds, int32_t min_mc_size, double cooling, int32_t burn_in,
                                int32_t max_sweeps, int32_t knn, int32_t rank, int32_t nranks, mcb_coclust** out)
{
    MCB_API_BEGIN
    MCB_CHECK(ctx && out, "mcb_coclust_run: null argument");
    MCB_CHECK(n_resamp == 0 || samples, "mcb_coclust_run: null sample arrays");
    *out = coclust_run_impl(ctx, N, n_edges, src, dst, w, n_resamp, n_s, samples, seeds, min_mc_size, cooling, burn_in, max_sweeps,
                            knn, rank, nranks, 1);
    MCB_API_END
}
// method = "edges" of tgs_graph_cover_resample: the same covers, co-occurrence counted for graph edges only
extern "C" int mcb_coclust_run_edges(mcb_ctx* ctx, int32_t N, int64_t n_edges, const int32_t* src, const int32_t* dst,
                                     const double* w, int32_t n_resamp, int32_t n_s, const int32_t* samples,
                                     const uint64_t* seeds, int32_t min_mc_size, double cooling, int32_t burn_in,
                                     int32_t max_sweeps, int32_t knn, int32_t rank, int32_t nranks, mcb_coclust** out)
{
    MCB_API_BEGIN
    MCB_CHECK(ctx && out, "mcb_coclust_run_edges: null argument");
    MCB_CHECK(n_resamp == 0 || samples, "mcb_coclust_run_edges: null sample arrays");
    *out = coclust_run_impl(ctx, N, n_edges, src, dst, w, n_resamp, n_s, samples, seeds, min_mc_size, cooling, burn_in, max_sweeps,
                            knn, rank, nranks, 2);
    MCB_API_END
}
// one row per edge of the graph, ordered by (source, weight descending, target): (node1, node2, cnt) + samples[N]
extern "C" int mcb_coclust_edge_counts(mcb_coclust* c, int32_t* host_node1, int32_t* host_node2, int32_t* host_cnt,
                                       int32_t* host_samples)
{
    MCB_API_BEGIN
    MCB_CHECK(c && c->mode == 1, "mcb_coclust_edge_counts: not a method = \"edges\" run");
    mcb_ctx* ctx = c->ctx;
    MCB_CUDA(cudaSetDevice(ctx->device));
    const size_t ne = (size_t)c->n_edges;
    if (ne) {
        if (host_node1) staged_copy_d2h(ctx, host_node1, c->e_src.p, ne * 4);
        if (host_node2) staged_copy_d2h(ctx, host_node2, c->e_dst.p, ne * 4);
        if (host_cnt) staged_copy_d2h(ctx, host_cnt, c->cnt.p, ne * 4);
    }
    if (host_samples) MCB_CUDA(cudaMemcpyAsync(host_samples, c->nsamp.p, (size_t)c->N * 4, cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    MCB_API_END
}
extern "C" int mcb_coclust_run(mcb_ctx* ctx, int32_t N, int64_t n_edges, const int32_t* src, const int32_t* dst,
                               const double* w, int32_t n_resamp, int32_t n_s, const int32_t* samples,
                               const uint64_t* seeds, int32_t min_mc_size, double cooling, int32_t burn_in,
                               int32_t max_sweeps, int32_t rank, int32_t nranks, mcb_coclust** out)
{
    return mcb_coclust_run2(ctx, N, n_edges, src, dst, w, n_resamp, n_s, samples, seeds, min_mc_size, cooling, burn_in, max_sweeps,
                            /*knn=*/0, rank, nranks, out);
}
// NCCL reduce(sum) of the co-occurrence counters and of samples[N] into `root`; collective
extern "C" int mcb_coclust_reduce(mcb_comm* comm, mcb_coclust* c, int32_t root)
{
    MCB_API_BEGIN
    MCB_CHECK(comm && c && c->cnt.p, "mcb_coclust_reduce: null argument");
    MCB_CHECK(root >= 0 && root < comm->nranks, "mcb_coclust_reduce: bad root");
    mcb_ctx* ctx = c->ctx;
    MCB_CUDA(cudaSetDevice(ctx->device));
    {
        StageTimer t(ctx, "reduce_counts", 0);
        comm_reduce_sum_u32(comm, c->cnt.p, c->mode == 1 ? (size_t)c->n_edges : (size_t)c->N * c->N, root);
        comm_reduce_sum_u32(comm, c->nsamp.p, (size_t)c->N, root);
    }
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    MCB_API_END
}
// tgs_graph_cover(edges, min_mc_size, cooling, burn_in) (reference call sites R/mc_graphcov.r:27,
// R/mc_from_coclust.r:247): ONE cover over all nodes, no co-occurrence counters (so no N x N memory: the cover of a
// million-node graph is fine).  cluster[v] = 0 for outliers (R/mc_graphcov.r:28), 1..n for covered nodes (ids are the
// cover's seeding order, not compacted).
extern "C" int mcb_graph_cover(mcb_ctx* ctx, int32_t N, int64_t n_edges, const int32_t* src, const int32_t* dst,
                               const double* w, int32_t min_mc_size, double cooling, int32_t burn_in, int32_t max_sweeps,
                               uint64_t seed, int32_t* host_cluster, int32_t* n_sweeps)
{
    mcb_coclust* c = nullptr;
    int rc = 0;
    try {
        MCB_CHECK(ctx && host_cluster && N >= 1, "mcb_graph_cover: bad argument");
        c = coclust_run_impl(ctx, N, n_edges, src, dst, w, 1, N, /*samples = all nodes*/ nullptr, &seed, min_mc_size, cooling, burn_in,
                             max_sweeps, 0, 0, 1, 0);
    }
    catch (const mcb::Error& e) { mcb::set_error(e.what()); rc = e.code; }
    catch (const std::bad_alloc&) { mcb::set_error("host out of memory"); rc = 3; }
    catch (const std::exception& e) { mcb::set_error(e.what()); rc = 4; }
    if (rc) return rc;
    int32_t nl = 0, sw = 0;
    rc = mcb_coclust_assignments(c, host_cluster, &sw, &nl);
    mcb_coclust_destroy(c);
    if (rc) return rc;
    for (int32_t v = 0; v < N; ++v) host_cluster[v] = host_cluster[v] < 0 ? 0 : host_cluster[v] + 1;
    if (n_sweeps) *n_sweeps = sw;
    return 0;
}

extern "C" void mcb_coclust_destroy(mcb_coclust* c)
{
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    delete c;
}
extern "C" int mcb_coclust_buffers(mcb_coclust* c, void** dev_cnt, void** dev_nsamp)
{
    MCB_API_BEGIN
    MCB_CHECK(c && c->cnt.p, "mcb_coclust_buffers: no counters on this handle");
    if (dev_cnt) *dev_cnt = c->cnt.p;
    if (dev_nsamp) *dev_nsamp = c->nsamp.p;
    MCB_API_END
}
extern "C" int mcb_coclust_assignments(mcb_coclust* c, int32_t* host_mc, int32_t* host_sweeps, int32_t* n_local)
{
    MCB_API_BEGIN
    MCB_CHECK(c, "null argument");
    mcb_ctx* ctx = c->ctx;
    MCB_CUDA(cudaSetDevice(ctx->device));
    if (n_local) *n_local = c->n_local;
    if (host_mc && c->n_local) MCB_CUDA(cudaMemcpyAsync(host_mc, c->mc.p, (size_t)c->n_local * c->N * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (host_sweeps && c->n_local) MCB_CUDA(cudaMemcpyAsync(host_sweeps, c->sweeps.p, (size_t)c->n_local * 4, cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    MCB_API_END
}
extern "C" int mcb_coclust_extract(mcb_coclust* c, int64_t* n_pairs)
{
    MCB_API_BEGIN
    MCB_CHECK(c && n_pairs && c->cnt.p, "null argument");
    MCB_CHECK(c->mode == 0, "mcb_coclust_extract: a method = \"edges\" run is read with mcb_coclust_edge_counts");
    mcb_ctx* ctx = c->ctx;
    MCB_CUDA(cudaSetDevice(ctx->device));
    const int32_t N = c->N;
    DevBuf<int64_t> rc(ctx, (size_t)N), offs(ctx, (size_t)N), total(ctx, 1);
    {
        StageTimer t(ctx, "coclust_extract", 3);
        launch_count_pairs(ctx, N, c->cnt.p, rc.p);
        launch_exclusive_scan_i64(ctx, rc.p, offs.p, N, total.p);
        int64_t np = 0;
        MCB_CUDA(cudaMemcpyAsync(&np, total.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        MCB_CUDA(cudaStreamSynchronize(ctx->stream));
        c->n_pairs = np;
        c->p1.alloc(ctx, (size_t)std::max<int64_t>(np, 1));
        c->p2.alloc(ctx, (size_t)std::max<int64_t>(np, 1));
        c->pc.alloc(ctx, (size_t)std::max<int64_t>(np, 1));
        launch_extract_pairs(ctx, N, c->cnt.p, offs.p, c->p1.p, c->p2.p, c->pc.p);
    }
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    *n_pairs = c->n_pairs;
    MCB_API_END
}
extern "C" int mcb_coclust_pairs(mcb_coclust* c, int32_t* h1, int32_t* h2, int32_t* hc, int32_t* host_samples)
{
    MCB_API_BEGIN
    MCB_CHECK(c && c->n_pairs >= 0, "mcb_coclust_pairs: run mcb_coclust_extract first");
    mcb_ctx* ctx = c->ctx;
    MCB_CUDA(cudaSetDevice(ctx->device));
    const size_t np = (size_t)c->n_pairs;
    if (np) {           // tens of MB at C3: through the pinned ring with the host threads
        StageTimer t(ctx, "d2h_pairs", 0);
        if (h1) staged_copy_d2h(ctx, h1, c->p1.p, np * 4);
        if (h2) staged_copy_d2h(ctx, h2, c->p2.p, np * 4);
        if (hc) staged_copy_d2h(ctx, hc, c->pc.p, np * 4);
    }
    if (host_samples) MCB_CUDA(cudaMemcpyAsync(host_samples, c->nsamp.p, (size_t)c->N * 4, cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    MCB_API_END
}
