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
    int nslots = 0, wave = 0;                                 // ring = 2 waves of `wave` slots
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

static int default_host_threads()
{
    if (const char* e = getenv("MCB_HOST_THREADS")) { const int v = atoi(e); if (v >= 1) return std::min(v, 64); }
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)std::max(2u, std::min(16u, hw ? hw / 2 : 4u));
}

Staging* ctx_staging(mcb_ctx* ctx)
{
    std::lock_guard<std::mutex> lk(ctx->alloc_mu);
    if (ctx->staging) return ctx->staging;
    std::unique_ptr<Staging> s(new Staging());
    const int nt = ctx->host_threads > 0 ? ctx->host_threads : default_host_threads();
    s->workers.reset(new HostWorkers(nt));
    s->wave = nt;
    s->nslots = 2 * nt;
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

// Generic staged host -> device pipeline.  The job is cut into chunks; fill(chunk, slot_ptr) writes the chunk's
// narrowed bytes into a pinned slot and returns their count; dst(chunk) is where they go on the device.  Waves of
// `wave` chunks are filled in parallel while the copy engine drains the previous wave.  The copies run on the
// staging stream; on return the context's stream has been made to wait for them.
static void staged_h2d(mcb_ctx* ctx, int64_t nchunks, const std::function<size_t(int64_t, char*)>& fill,
                       const std::function<char*(int64_t)>& dst)
{
    Staging* st = ctx_staging(ctx);
    // the destinations come from the context's block cache, whose reuse is ordered on ctx->stream only: the copy
    // stream must not run ahead of work already queued there
    MCB_CUDA(cudaEventRecord(st->done, ctx->stream));
    MCB_CUDA(cudaStreamWaitEvent(st->copy, st->done, 0));
    std::vector<size_t> bytes((size_t)st->wave);
    int half = 0;
    for (int64_t c0 = 0; c0 < nchunks; c0 += st->wave, half ^= 1) {
        const int n = (int)std::min<int64_t>(st->wave, nchunks - c0);
        const int s0 = half * st->wave;
        st->workers->parallel_for(n, [&](int q) {
            MCB_CUDA(cudaEventSynchronize(st->ev[s0 + q]));           // the slot's previous copy has left
            bytes[q] = fill(c0 + q, st->slot[s0 + q]);
        });
        for (int q = 0; q < n; ++q) {
            if (bytes[q]) MCB_CUDA(cudaMemcpyAsync(dst(c0 + q), st->slot[s0 + q], bytes[q], cudaMemcpyHostToDevice, st->copy));
            MCB_CUDA(cudaEventRecord(st->ev[s0 + q], st->copy));
        }
    }
    MCB_CUDA(cudaEventRecord(st->done, st->copy));
    MCB_CUDA(cudaStreamWaitEvent(ctx->stream, st->done, 0));
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
    // ---- @p: as is (8 B per cell), rebased on the device
    {
        const int64_t per = (int64_t)(Staging::SLOT / sizeof(int64_t)), n = ncells + 1;
        staged_h2d(ctx, ceil_div(n, per),
                   [&](int64_t c, char* s) { const int64_t a = c * per, b = std::min(n, a + per); memcpy(s, p + a, (size_t)(b - a) * 8); return (size_t)(b - a) * 8; },
                   [&](int64_t c) { return (char*)(d_p + c * per); });
        if (base) {
            rebase_p_kernel<<<(int)std::min<int64_t>(ceil_div(n, 256), 1024), 256, 0, ctx->stream>>>(d_p, n, base);
            MCB_LAUNCH_CHECK();
        }
    }
    if (nnz == 0) return;
    const int64_t per = (int64_t)(Staging::SLOT / sizeof(uint16_t));        // entries per chunk (2 B each)
    const int64_t nchunks = ceil_div(nnz, per);
    // ---- @i
    if (ngenes <= 65536) {
        DevBuf<uint16_t> n16(ctx, (size_t)nnz);
        staged_h2d(ctx, nchunks,
                   [&](int64_t c, char* s) {
                       const int64_t a = c * per, b = std::min(nnz, a + per);
                       const int32_t* src = i + base + a;
                       uint16_t* o = (uint16_t*)s;
                       int bad = 0;
                       for (int64_t q = 0; q < b - a; ++q) { const int32_t v = src[q]; bad |= (v < 0) | (v > 65535); o[q] = (uint16_t)v; }
                       if (bad) throw Error(1, "MC-ERR: mcb_csc_upload: gene row index out of range");
                       return (size_t)(b - a) * 2;
                   },
                   [&](int64_t c) { return (char*)(n16.p + c * per); });
        widen_u16_kernel<<<(int)std::min<int64_t>(ceil_div(nnz, 2048), (int64_t)ctx->num_sms * 8), 256, 0, ctx->stream>>>(n16.p, (uint32_t*)d_i, nnz);
        MCB_LAUNCH_CHECK();
    } else {
        const int64_t per4 = (int64_t)(Staging::SLOT / sizeof(int32_t));
        staged_h2d(ctx, ceil_div(nnz, per4),
                   [&](int64_t c, char* s) { const int64_t a = c * per4, b = std::min(nnz, a + per4); memcpy(s, i + base + a, (size_t)(b - a) * 4); return (size_t)(b - a) * 4; },
                   [&](int64_t c) { return (char*)(d_i + c * per4); });
    }
    // ---- @x: double -> uint16 with escapes, validated on the way
    {
        DevBuf<uint16_t> n16(ctx, (size_t)nnz);
        std::mutex esc_mu;
        std::vector<int64_t> esc_pos;
        std::vector<uint32_t> esc_val;
        staged_h2d(ctx, nchunks,
                   [&](int64_t c, char* s) {
                       const int64_t a = c * per, b = std::min(nnz, a + per);
                       const double* src = x + base + a;
                       uint16_t* o = (uint16_t*)s;
                       // fast pass: branch-free narrowing; `big` collects every bit seen, `ok` the round-trip test
                       uint64_t big = 0;
                       bool ok = true;
                       for (int64_t q = 0; q < b - a; ++q) {
                           const double v = src[q];
                           const int64_t t = (int64_t)v;                           // out-of-range / NaN -> INT64_MIN
                           ok &= ((double)t == v);
                           big |= (uint64_t)t;
                           o[q] = (uint16_t)((uint64_t)t < 65535u ? t : 65535);
                       }
                       if (!ok || big >= 4294967296ull)                           // non-integral, NaN, negative or >= 2^32
                           throw Error(1, "MC-ERR: mcb_csc_upload: counts must be non-negative integers below 2^32");
                       if (big >= 65535u) {                                        // rare: escape the large counts
                           for (int64_t q = 0; q < b - a; ++q) {
                               if (o[q] != 65535u) continue;
                               std::lock_guard<std::mutex> lk(esc_mu);
                               esc_pos.push_back(a + q); esc_val.push_back((uint32_t)(int64_t)src[q]);
                           }
                       }
                       return (size_t)(b - a) * 2;
                   },
                   [&](int64_t c) { return (char*)(n16.p + c * per); });
        widen_u16_kernel<<<(int)std::min<int64_t>(ceil_div(nnz, 2048), (int64_t)ctx->num_sms * 8), 256, 0, ctx->stream>>>(n16.p, d_x, nnz);
        MCB_LAUNCH_CHECK();
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
    // mc1 and w depend on the degrees only: rebuilt by the host threads while the targets are in flight
    const int nt = st->workers->size();
    auto fill_src_w = [&](int part) {
        const int64_t r0 = rows * part / nt, r1 = rows * (part + 1) / nt;
        for (int64_t r = r0; r < r1; ++r) {
            const int64_t o = offs[r];
            const int32_t c = cell[r];
            for (int32_t q = 0; q < deg[r]; ++q) {
                if (host_src) host_src[o + q] = c;
                if (host_w) host_w[o + q] = 1.0 - (double)q / (double)K;
            }
        }
    };
    if (!host_dst || ne == 0) { st->workers->parallel_for(nt, fill_src_w); return; }
    const int64_t per = (int64_t)(Staging::SLOT / sizeof(int32_t));
    const int64_t nchunks = ceil_div(ne, per);
    // issue the D2H copies in waves; drain wave w (memcpy into the caller's array) while wave w+1 is in flight
    auto issue = [&](int64_t c0, int half) {
        const int n = (int)std::min<int64_t>(st->wave, nchunks - c0);
        for (int q = 0; q < n; ++q) {
            const int64_t a = (c0 + q) * per, b = std::min(ne, a + per);
            MCB_CUDA(cudaMemcpyAsync(st->slot[half * st->wave + q], d_e_dst + a, (size_t)(b - a) * 4, cudaMemcpyDeviceToHost, st->copy));
            MCB_CUDA(cudaEventRecord(st->ev[half * st->wave + q], st->copy));
        }
    };
    issue(0, 0);
    bool src_done = false;
    int half = 0;
    for (int64_t c0 = 0; c0 < nchunks; c0 += st->wave, half ^= 1) {
        if (c0 + st->wave < nchunks) issue(c0 + st->wave, half ^ 1);
        if (!src_done) { st->workers->parallel_for(nt, fill_src_w); src_done = true; }
        const int n = (int)std::min<int64_t>(st->wave, nchunks - c0);
        st->workers->parallel_for(n, [&](int q) {
            const int64_t a = (c0 + q) * per, b = std::min(ne, a + per);
            MCB_CUDA(cudaEventSynchronize(st->ev[half * st->wave + q]));
            memcpy(host_dst + a, st->slot[half * st->wave + q], (size_t)(b - a) * 4);
        });
    }
    // leave the ring's events signalled for the next user (they are: every recorded event was synchronised above)
}

}  // namespace mcb
