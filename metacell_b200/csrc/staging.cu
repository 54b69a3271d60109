// staging.cu -- host <-> device transfers of the two large objects that cross the C ABI: the dgCMatrix slots going in
// (R/scmat.r tgScMat@mat: @p, @i, @x) and the (mc1, mc2, w) edge table coming out (R/cgraph_knn.r:22).
//
// R hands over ordinary pageable vectors.  A cudaMemcpyAsync straight from pageable memory is staged by the driver
// on one thread at a fraction of PCIe speed, so the library stages itself: a ring of pinned slots per context, a few
// host threads that fill (or drain) slots while the copy engine moves the previous ones, one cudaMemcpyAsync per slot
// on a dedicated copy stream.  While a slot is filled the data is also NARROWED so fewer bytes cross PCIe:
//   @x  double -> uint16 (counts >= 65535 are escaped into a short patch list; non-integral, negative or >= 2^32 values
//       are rejected here, on the way through)                                         8 B -> 2 B per stored entry
//   @i  int32  -> uint16 when the matrix has <= 65536 gene rows                          4 B -> 2 B per stored entry
//   edges: the device sends the targets (int32 per edge) and the out-degree (int32 per row); the host threads rebuild
//       mc1 (runs of the source id) and w = 1 - place/K (a function of the position in the run)   16 B -> ~4 B per edge
// A small kernel widens the narrowed arrays on the device (HBM-bound, ~0.1 ms at C2).
#include "internal.h"
#include "narrow.h"
#include <condition_variable>
#include <cstring>
#include <functional>
#include <thread>

namespace mcb {

// ---------------------------------------------------------------------------------------------------------------
// a fixed pool of host threads: parallel_for(n, fn) runs fn(0..n-1), the caller participates
// ---------------------------------------------------------------------------------------------------------------
class HostWorkers {
public:
    explicit HostWorkers(int nthreads)
    {
        for (int t = 0; t < nthreads - 1; ++t) threads_.emplace_back([this] { loop(); });
    }
    ~HostWorkers()
    {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; ++gen_; }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    int size() const { return (int)threads_.size() + 1; }
    void parallel_for(int n, const std::function<void(int)>& fn)
    {
        if (n <= 0) return;
        if (n == 1 || threads_.empty()) { for (int q = 0; q < n; ++q) fn(q); return; }
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn; n_ = n; next_.store(0); done_ = 0; err_.clear(); ++gen_;
        }
        cv_.notify_all();
        run_items();
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [this] { return done_ == n_; });
        fn_ = nullptr;
        if (!err_.empty()) throw Error(1, err_);
    }

private:
    void run_items()
    {
        for (;;) {
            const int q = next_.fetch_add(1);
            if (q >= n_) break;
            try { (*fn_)(q); }
            catch (const std::exception& e) { std::lock_guard<std::mutex> lk(mu_); if (err_.empty()) err_ = e.what(); }
            std::lock_guard<std::mutex> lk(mu_);
            if (++done_ == n_) cv_done_.notify_all();
        }
    }
    void loop()
    {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
            }
            run_items();
        }
    }
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_, cv_done_;
    const std::function<void(int)>* fn_ = nullptr;
    int n_ = 0, done_ = 0;
    std::atomic<int> next_{0};
    uint64_t gen_ = 0;
    bool stop_ = false;
    std::string err_;
};

// ---------------------------------------------------------------------------------------------------------------
// per-context pinned ring + copy stream
// ---------------------------------------------------------------------------------------------------------------
struct Staging {
    static constexpr size_t SLOT = (size_t)4 << 20;        // bytes per pinned slot
    int nslots = 0, wave = 0, wave_max = 0;                   // two pinned slots per thread; `wave` threads stage uploads, `wave_max` downloads
    std::vector<char*> slot;
    std::vector<cudaEvent_t> ev;
    cudaStream_t copy = nullptr;
    cudaEvent_t done = nullptr;
    std::unique_ptr<HostWorkers> workers;
    ~Staging()
    {
        for (auto p : slot) if (p) cudaFreeHost(p);
        for (auto e : ev) if (e) cudaEventDestroy(e);
        if (done) cudaEventDestroy(done);
        if (copy) cudaStreamDestroy(copy);
    }
};

// threads for a phase in which every process of the node stages at the same time (uploads): the cores are shared.
// `exclusive`: a phase only one process runs (the root's downloads after a gather or a reduce): all cores.
static int default_host_threads(bool exclusive = false)
{
    if (const char* e = getenv("MCB_HOST_THREADS")) { const int v = atoi(e); if (v >= 1) return std::min(v, 64); }
    unsigned hw = std::thread::hardware_concurrency();
    if (!hw) hw = 4;
    if (exclusive) return (int)std::max(2u, std::min(16u, hw));
    // one process per device on the same host (torchrun, mpirun): share the cores, twice oversubscribed -- the ranks do not
    // all stage at the same moment (rank 0 alone downloads the gathered edge table).  Measured on 16 cores and 8 ranks:
    // end to end 18.7 ms with 2 threads per rank, 16.1 with 4, 15.8 with 8
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) { const int v = atoi(e); if (v > 1) hw = std::max(2u, 2u * hw / (unsigned)v); }
    return (int)std::max(2u, std::min(16u, hw));
}

Staging* ctx_staging(mcb_ctx* ctx)
{
    std::lock_guard<std::mutex> lk(ctx->alloc_mu);
    if (ctx->staging) return ctx->staging;
    std::unique_ptr<Staging> s(new Staging());
    const int nt = ctx->host_threads > 0 ? ctx->host_threads : default_host_threads();
    const int nmax = ctx->host_threads > 0 ? ctx->host_threads : std::max(nt, default_host_threads(true));
    s->workers.reset(new HostWorkers(nmax));
    s->wave = nt;
    s->wave_max = nmax;
    s->nslots = 2 * nmax;
    s->slot.assign((size_t)s->nslots, nullptr);
    s->ev.assign((size_t)s->nslots, nullptr);
    for (int q = 0; q < s->nslots; ++q) {
        MCB_CUDA(cudaHostAlloc((void**)&s->slot[q], Staging::SLOT, cudaHostAllocDefault));
        MCB_CUDA(cudaEventCreateWithFlags(&s->ev[q], cudaEventDisableTiming));
    }
    MCB_CUDA(cudaStreamCreateWithFlags(&s->copy, cudaStreamNonBlocking));
    MCB_CUDA(cudaEventCreateWithFlags(&s->done, cudaEventDisableTiming));
    ctx->staging = s.release();
    return ctx->staging;
}
void ctx_staging_destroy(mcb_ctx* ctx)
{
    delete ctx->staging;
    ctx->staging = nullptr;
}

// Generic staged host -> device pipeline.  The job is cut into chunks handed out dynamically; every host thread owns
// two pinned slots and alternates between them: while the copy engine drains one, the thread fills the other.
// work(chunk, slot_ptr, copy_stream) fills the slot and enqueues its own cudaMemcpyAsync calls on the copy stream.
// On return the context's stream has been made to wait for every copy.
static void staged_h2d(mcb_ctx* ctx, int64_t nchunks, const std::function<void(int64_t, char*, cudaStream_t)>& work)
{
    Staging* st = ctx_staging(ctx);
    // the destinations come from the context's block cache, whose reuse is ordered on ctx->stream only: the copy
    // stream must not run ahead of work already queued there
    MCB_CUDA(cudaEventRecord(st->done, ctx->stream));
    MCB_CUDA(cudaStreamWaitEvent(st->copy, st->done, 0));
    std::atomic<int64_t> next{0};
    const int nt = (int)std::min<int64_t>(st->wave, nchunks);
    st->workers->parallel_for(nt, [&](int t) {
        MCB_CUDA(cudaSetDevice(ctx->device));
        int flip = 0;
        for (;;) {
            const int64_t c = next.fetch_add(1);
            if (c >= nchunks) break;
            const int s = 2 * t + flip;
            flip ^= 1;
            MCB_CUDA(cudaEventSynchronize(st->ev[s]));           // the slot's previous copy has left
            work(c, st->slot[s], st->copy);
            MCB_CUDA(cudaEventRecord(st->ev[s], st->copy));
        }
    });
    MCB_CUDA(cudaEventRecord(st->done, st->copy));
    MCB_CUDA(cudaStreamWaitEvent(ctx->stream, st->done, 0));
}

// chunk of a plain copy: small enough that every staging thread gets a few (a 10 MB array cut into 4 MB slots kept
// three threads busy), at most one slot
static size_t copy_chunk(int nthreads, size_t bytes)
{
    size_t c = bytes / (size_t)(4 * std::max(nthreads, 1));
    c = std::max<size_t>(c, (size_t)256 << 10);
    c = std::min<size_t>(c, Staging::SLOT);
    return (c + 4095) & ~(size_t)4095;
}
// plain bytes, host -> device, through the pinned ring (pageable sources: the edge table and the resample index sets
// of the co-clustering call).  Small copies go straight through cudaMemcpyAsync.
void staged_copy_h2d(mcb_ctx* ctx, void* dev, const void* host, size_t bytes)
{
    if (!bytes) return;
    if (bytes < ((size_t)1 << 20)) { MCB_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream)); return; }
    const size_t CH = copy_chunk(ctx_staging(ctx)->wave, bytes);
    const int64_t nchunks = (int64_t)ceil_div((int64_t)bytes, (int64_t)CH);
    staged_h2d(ctx, nchunks, [&](int64_t c, char* slot, cudaStream_t copy) {
        const size_t a = (size_t)c * CH, n = std::min(CH, bytes - a);
        memcpy(slot, (const char*)host + a, n);
        MCB_CUDA(cudaMemcpyAsync((char*)dev + a, slot, n, cudaMemcpyHostToDevice, copy));
    });
}
// plain bytes, device -> host, through the pinned ring; returns when everything is in `host`.  The device data must have
// been produced on ctx->stream.
void staged_copy_d2h(mcb_ctx* ctx, void* host, const void* dev, size_t bytes)
{
    if (!bytes) return;
    if (bytes < ((size_t)1 << 20)) {
        MCB_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        MCB_CUDA(cudaStreamSynchronize(ctx->stream));
        return;
    }
    Staging* st = ctx_staging(ctx);
    MCB_CUDA(cudaEventRecord(st->done, ctx->stream));
    MCB_CUDA(cudaStreamWaitEvent(st->copy, st->done, 0));          // the copy stream starts after the producer kernels
    const size_t CH = copy_chunk(st->wave_max, bytes);
    const int64_t nchunks = (int64_t)ceil_div((int64_t)bytes, (int64_t)CH);
    std::atomic<int64_t> next{0};
    const int nt = (int)std::min<int64_t>(st->wave_max, nchunks);
    st->workers->parallel_for(nt, [&](int t) {
        MCB_CUDA(cudaSetDevice(ctx->device));
        int64_t pend[2] = {-1, -1};
        auto issue = [&](int sl) {
            const int64_t c = next.fetch_add(1);
            if (c >= nchunks) { pend[sl] = -1; return; }
            const size_t a = (size_t)c * CH, n = std::min(CH, bytes - a);
            MCB_CUDA(cudaMemcpyAsync(st->slot[2 * t + sl], (const char*)dev + a, n, cudaMemcpyDeviceToHost, st->copy));
            MCB_CUDA(cudaEventRecord(st->ev[2 * t + sl], st->copy));
            pend[sl] = c;
        };
        issue(0);
        issue(1);
        for (int sl = 0; pend[0] >= 0 || pend[1] >= 0; sl ^= 1) {
            if (pend[sl] < 0) continue;
            const size_t a = (size_t)pend[sl] * CH, n = std::min(CH, bytes - a);
            MCB_CUDA(cudaEventSynchronize(st->ev[2 * t + sl]));
            memcpy((char*)host + a, st->slot[2 * t + sl], n);
            issue(sl);
        }
    });
}

// ---------------------------------------------------------------------------------------------------------------
// device-side widening / validation kernels
// ---------------------------------------------------------------------------------------------------------------
__global__ void widen_u16_kernel(const uint16_t* __restrict__ in, uint32_t* __restrict__ out, int64_t n)
{
    // 8 values per thread: one 16-byte load, two 16-byte stores
    const int64_t nv = n >> 3;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += (int64_t)gridDim.x * blockDim.x) {
        const uint4 a = reinterpret_cast<const uint4*>(in)[v];
        uint4 lo, hi;
        lo.x = a.x & 0xffffu; lo.y = a.x >> 16; lo.z = a.y & 0xffffu; lo.w = a.y >> 16;
        hi.x = a.z & 0xffffu; hi.y = a.z >> 16; hi.z = a.w & 0xffffu; hi.w = a.w >> 16;
        reinterpret_cast<uint4*>(out)[2 * v] = lo;
        reinterpret_cast<uint4*>(out)[2 * v + 1] = hi;
    }
    for (int64_t e = (nv << 3) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) out[e] = in[e];
}
__global__ void patch_u32_kernel(const int64_t* __restrict__ pos, const uint32_t* __restrict__ val, int64_t n, uint32_t* __restrict__ out)
{
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) out[pos[q]] = val[q];
}
__global__ void rebase_p_kernel(int64_t* __restrict__ p, int64_t n, int64_t base)
{
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) p[q] -= base;
}
// structure checks of a CSC matrix (one warp per column): p non-decreasing and inside [0, nnz]; row indices inside
// [0, ngenes) and strictly ascending within a column (so they are unique).  flags: 1 = p, 2 = row range, 4 = row order
__global__ void csc_validate_kernel(int32_t ngenes, int64_t ncells, int64_t nnz, const int64_t* __restrict__ p,
                                    const int32_t* __restrict__ ridx, int32_t* __restrict__ flags)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    int bad = 0;
    for (int64_t c = warp; c < ncells; c += nwarps) {
        const int64_t e0 = p[c], e1 = p[c + 1];
        if (e0 < 0 || e1 < e0 || e1 > nnz) { bad |= 1; continue; }
        for (int64_t e = e0 + lane; e < e1; e += 32) {
            const int32_t g = ridx[e];
            if (g < 0 || g >= ngenes) bad |= 2;
            if (e > e0 && ridx[e - 1] >= g) bad |= 4;
        }
    }
    if (bad) atomicOr(flags, bad);
}

void launch_csc_validate(mcb_ctx* ctx, int32_t ngenes, int64_t ncells, int64_t nnz, const int64_t* p, const int32_t* ridx, int32_t* flags)
{
    if (ncells == 0) return;
    int grid = (int)std::min<int64_t>(ceil_div(ncells, 8), (int64_t)ctx->num_sms * 16);
    csc_validate_kernel<<<grid, 256, 0, ctx->stream>>>(ngenes, ncells, nnz, p, ridx, flags);
    MCB_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------------------
// CSC upload: columns [0, ncells) whose entries live at i[base .. base + nnz), x[base .. base + nnz); p holds ABSOLUTE
// offsets (p[0] == base), so a rank can upload its cell range of a matrix shared by every rank of the process.
// On return the copies are enqueued and ordered before anything later on ctx->stream; the host arrays may be reused
// (everything has left them).  d_p is rebased to start at 0.
// ---------------------------------------------------------------------------------------------------------------
void staged_upload_csc(mcb_ctx* ctx, int32_t ngenes, int64_t ncells, const int64_t* p, const int32_t* i, const double* x,
                       int64_t* d_p, int32_t* d_i, uint32_t* d_x)
{
    const int64_t base = p[0], nnz = p[ncells] - base;
    Staging* st = ctx_staging(ctx);
    // ---- @p: as is (8 B per cell), rebased on the device
    {
        const int64_t n = ncells + 1;
        MCB_CUDA(cudaEventRecord(st->done, ctx->stream));
        MCB_CUDA(cudaStreamWaitEvent(st->copy, st->done, 0));
        MCB_CUDA(cudaMemcpyAsync(d_p, p, (size_t)n * 8, cudaMemcpyHostToDevice, st->copy));    // small: the driver's own staging
        MCB_CUDA(cudaEventRecord(st->done, st->copy));
        MCB_CUDA(cudaStreamWaitEvent(ctx->stream, st->done, 0));
        if (base) {
            rebase_p_kernel<<<(int)std::min<int64_t>(ceil_div(n, 256), 1024), 256, 0, ctx->stream>>>(d_p, n, base);
            MCB_LAUNCH_CHECK();
        }
    }
    if (nnz == 0) return;
    // ---- @i and @x in one pass over the stored entries.  A slot holds, for E entries, the narrowed counts (2 B each)
    // followed by the narrowed (2 B) or plain (4 B) row indices.
    const bool rows16 = ngenes <= 65536;
    const int64_t per = (int64_t)(Staging::SLOT / (rows16 ? 4 : 6)) & ~(int64_t)63;
    const int64_t nchunks = ceil_div(nnz, per);
    DevBuf<uint16_t> x16(ctx, (size_t)nnz), i16(ctx, rows16 ? (size_t)nnz : 0);
    std::mutex esc_mu;
    std::vector<int64_t> esc_pos;
    std::vector<uint32_t> esc_val;
    staged_h2d(ctx, nchunks, [&](int64_t c, char* slot, cudaStream_t cs) {
        const int64_t a = c * per, b = std::min(nnz, a + per), n = b - a;
        uint16_t* ox = (uint16_t*)slot;
        char* oi = slot + (size_t)per * 2;
        const double* sx = x + base + a;
        const int rc = narrow_counts_u16(sx, (size_t)n, ox);
        if (rc == 2) throw Error(1, "MC-ERR: mcb_csc_upload: counts must be non-negative integers below 2^32");
        if (rc == 1) {                                       // rare: counts >= 65535 travel in a patch list
            std::lock_guard<std::mutex> lk(esc_mu);
            for (int64_t q = 0; q < n; ++q)
                if (ox[q] == 65535u) { esc_pos.push_back(a + q); esc_val.push_back((uint32_t)sx[q]); }
        }
        MCB_CUDA(cudaMemcpyAsync(x16.p + a, ox, (size_t)n * 2, cudaMemcpyHostToDevice, cs));
        if (rows16) {
            if (narrow_rows_u16(i + base + a, (size_t)n, (uint16_t*)oi)) throw Error(1, "MC-ERR: mcb_csc_upload: gene row index out of range");
            MCB_CUDA(cudaMemcpyAsync(i16.p + a, oi, (size_t)n * 2, cudaMemcpyHostToDevice, cs));
        } else {
            memcpy(oi, i + base + a, (size_t)n * 4);
            MCB_CUDA(cudaMemcpyAsync(d_i + a, oi, (size_t)n * 4, cudaMemcpyHostToDevice, cs));
        }
    });
    const int wgrid = (int)std::min<int64_t>(ceil_div(nnz, 2048), (int64_t)ctx->num_sms * 8);
    widen_u16_kernel<<<wgrid, 256, 0, ctx->stream>>>(x16.p, d_x, nnz);
    MCB_LAUNCH_CHECK();
    if (rows16) { widen_u16_kernel<<<wgrid, 256, 0, ctx->stream>>>(i16.p, (uint32_t*)d_i, nnz); MCB_LAUNCH_CHECK(); }
    if (!esc_pos.empty()) {
        const int64_t ne = (int64_t)esc_pos.size();
        DevBuf<int64_t> dp(ctx, (size_t)ne);
        DevBuf<uint32_t> dv(ctx, (size_t)ne);
        MCB_CUDA(cudaMemcpyAsync(dp.p, esc_pos.data(), (size_t)ne * 8, cudaMemcpyHostToDevice, ctx->stream));
        MCB_CUDA(cudaMemcpyAsync(dv.p, esc_val.data(), (size_t)ne * 4, cudaMemcpyHostToDevice, ctx->stream));
        patch_u32_kernel<<<(int)std::min<int64_t>(ceil_div(ne, 256), 1024), 256, 0, ctx->stream>>>(dp.p, dv.p, ne, d_x);
        MCB_LAUNCH_CHECK();
        MCB_CUDA(cudaStreamSynchronize(ctx->stream));                         // esc_* are pageable host vectors
    }
}

// ---------------------------------------------------------------------------------------------------------------
// edge table download: device (out_deg[rows], e_dst[ne] already in cell ids, cell ids of the rows) -> host
// (mc1, mc2, w) at [edge_base, edge_base + ne) of the caller's arrays.  w = 1 - place/K (R/cgraph_knn.r:46 rank
// fractions; the same expression as emit_edges_kernel, so the doubles are bit-identical).
// ---------------------------------------------------------------------------------------------------------------
void staged_download_edges(mcb_ctx* ctx, int64_t rows, int32_t K, int64_t ne, const int32_t* d_out_deg, const int32_t* d_src_cell,
                           const int32_t* d_e_dst, int32_t* host_src, int32_t* host_dst, double* host_w)
{
    if (rows == 0) return;
    Staging* st = ctx_staging(ctx);
    // degrees + source cell ids first (small), then the targets in slot-sized chunks
    std::vector<int32_t> deg((size_t)rows), cell((size_t)rows);
    MCB_CUDA(cudaMemcpyAsync(deg.data(), d_out_deg, (size_t)rows * 4, cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaMemcpyAsync(cell.data(), d_src_cell, (size_t)rows * 4, cudaMemcpyDeviceToHost, ctx->stream));
    MCB_CUDA(cudaEventRecord(st->done, ctx->stream));
    MCB_CUDA(cudaStreamWaitEvent(st->copy, st->done, 0));          // the copy stream starts after the producer kernels
    MCB_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<int64_t> offs((size_t)rows + 1, 0);
    for (int64_t r = 0; r < rows; ++r) offs[r + 1] = offs[r] + deg[r];
    MCB_CHECK(offs[rows] == ne, "edge table: degree sum does not match the edge count");
    const int64_t per = (int64_t)(Staging::SLOT / sizeof(int32_t));
    const int64_t nchunks = (host_dst && ne) ? ceil_div(ne, per) : 0;
    const int nt = st->wave_max;
    std::atomic<int64_t> next{0};
    st->workers->parallel_for(nt, [&](int t) {
        MCB_CUDA(cudaSetDevice(ctx->device));
        // every thread keeps up to two target chunks in flight (its two slots) ...
        int64_t pend[2] = {-1, -1};
        auto issue = [&](int sl) {
            const int64_t c = next.fetch_add(1);
            if (c >= nchunks) { pend[sl] = -1; return; }
            const int64_t a = c * per, b = std::min(ne, a + per);
            MCB_CUDA(cudaMemcpyAsync(st->slot[2 * t + sl], d_e_dst + a, (size_t)(b - a) * 4, cudaMemcpyDeviceToHost, st->copy));
            MCB_CUDA(cudaEventRecord(st->ev[2 * t + sl], st->copy));
            pend[sl] = c;
        };
        issue(0);
        issue(1);
        // ... and meanwhile rebuilds mc1 and w for its share of the rows: they depend on the degrees only
        {
            const int64_t r0 = rows * t / nt, r1 = rows * (t + 1) / nt;
            for (int64_t r = r0; r < r1; ++r) {
                const int64_t o = offs[r];
                const int32_t c = cell[r];
                const int32_t d = deg[r];
                if (host_src) for (int32_t q = 0; q < d; ++q) host_src[o + q] = c;
                if (host_w) for (int32_t q = 0; q < d; ++q) host_w[o + q] = 1.0 - (double)q / (double)K;
            }
        }
        for (int sl = 0; pend[0] >= 0 || pend[1] >= 0; sl ^= 1) {
            if (pend[sl] < 0) continue;
            const int64_t a = pend[sl] * per, b = std::min(ne, a + per);
            MCB_CUDA(cudaEventSynchronize(st->ev[2 * t + sl]));
            memcpy(host_dst + a, st->slot[2 * t + sl], (size_t)(b - a) * 4);
            issue(sl);
        }
    });
    // every recorded slot event was synchronised above: the ring is free for the next user
}

}  // namespace mcb
