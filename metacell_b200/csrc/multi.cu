// multi.cu -- ONE process driving P devices (SURVEY.md section 8b "Threading"): what the single R process of
// mcell_add_cgraph_from_mat_bknn (R/cgraph_knn.r:10-25) and mcell_coclust_from_graph_resamp
// (R/coclust_graph_cov.r:13-58) calls when the box has several B200s.  A mcb_multi owns one context, one NCCL
// communicator (ncclCommInitAll) and one host thread per device; a call runs the same rank-level code as the
// rank-collective entry points (api.cu) on every thread and joins before it returns.  Because all ranks share the
// caller's address space, every rank uploads its own cell range straight from the caller's dgCMatrix slots and
// downloads its own edge rows straight into the result table: P PCIe links work in parallel and nothing is funnelled
// through rank 0.
#include "../../include/mcb200.h"
#include "internal.h"
#include "handles.h"
#include <condition_variable>
#include <functional>
#include <thread>

using namespace mcb;

struct mcb_multi {
    int n = 0;
    std::vector<int> devices;
    std::vector<mcb_ctx*> ctx;
    std::vector<mcb_comm*> comm;
    // persistent workers: run(fn) executes fn(rank) on every worker and waits
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv, cv_done;
    const std::function<void(int)>* job = nullptr;
    uint64_t gen = 0;
    int done = 0;
    bool stop = false;
    std::vector<std::string> err;
    std::vector<int> err_code;
    // host barrier between the workers of one call (error agreement before a collective phase)
    int bar_count = 0; uint64_t bar_gen = 0; bool bar_fail = false;
};

static void worker_loop(mcb_multi* m, int r)
{
    cudaSetDevice(m->devices[r]);
    uint64_t seen = 0;
    for (;;) {
        const std::function<void(int)>* fn = nullptr;
        {
            std::unique_lock<std::mutex> lk(m->mu);
            m->cv.wait(lk, [&] { return m->gen != seen; });
            seen = m->gen;
            if (m->stop) return;
            fn = m->job;
        }
        int code = 0;
        std::string msg;
        try { (*fn)(r); }
        catch (const mcb::Error& e) { code = e.code; msg = e.what(); }
        catch (const std::bad_alloc&) { code = 3; msg = "host out of memory"; }
        catch (const std::exception& e) { code = 4; msg = e.what(); }
        catch (...) { code = 5; msg = "unknown error"; }
        std::lock_guard<std::mutex> lk(m->mu);
        m->err[r] = msg; m->err_code[r] = code;
        if (++m->done == m->n) m->cv_done.notify_all();
    }
}

// runs fn(rank) on every worker; throws the first rank's error
static void multi_run(mcb_multi* m, const std::function<void(int)>& fn)
{
    {
        std::lock_guard<std::mutex> lk(m->mu);
        m->job = &fn; m->done = 0; m->bar_count = 0; m->bar_fail = false; ++m->gen;
        for (int r = 0; r < m->n; ++r) { m->err[r].clear(); m->err_code[r] = 0; }
    }
    m->cv.notify_all();
    std::unique_lock<std::mutex> lk(m->mu);
    m->cv_done.wait(lk, [&] { return m->done == m->n; });
    m->job = nullptr;
    int first = -1;
    for (int r = 0; r < m->n; ++r)          // prefer a root-cause message over the "aborted with the others" echo
        if (m->err_code[r] && (first < 0 || (m->err_code[first] == 6 && m->err_code[r] != 6))) first = r;
    if (first >= 0) {
        char b[64];
        snprintf(b, sizeof(b), " (device %d)", m->devices[first]);
        throw Error(m->err_code[first], m->err[first] + b);
    }
}

// Error agreement: every worker calls it with its own status before entering a collective phase; if any worker failed
// they all leave (a rank that skipped a collective would hang the others inside NCCL).
static void multi_agree(mcb_multi* m, bool ok)
{
    std::unique_lock<std::mutex> lk(m->mu);
    if (!ok) m->bar_fail = true;
    const uint64_t g = m->bar_gen;
    if (++m->bar_count == m->n) { m->bar_count = 0; ++m->bar_gen; m->cv_done.notify_all(); }
    else m->cv_done.wait(lk, [&] { return m->bar_gen != g; });
    if (m->bar_fail && ok) throw Error(6, "aborted: another device failed");
}

#define MCB_API_BEGIN try {
#define MCB_API_END                                                                            \
    return 0;                                                                                  \
    } catch (const mcb::Error& e) { mcb::set_error(e.what()); return e.code; }                 \
    catch (const std::bad_alloc&) { mcb::set_error("host out of memory"); return 3; }          \
    catch (const std::exception& e) { mcb::set_error(e.what()); return 4; }                    \
    catch (...) { mcb::set_error("unknown error"); return 5; }

extern "C" void mcb_multi_destroy(mcb_multi* m)
{
    if (!m) return;
    {
        std::lock_guard<std::mutex> lk(m->mu);
        m->stop = true; ++m->gen;
    }
    m->cv.notify_all();
    for (auto& t : m->threads) if (t.joinable()) t.join();
    for (int r = 0; r < (int)m->comm.size(); ++r) {
        if (!m->comm[r]) continue;
        void* raw = m->comm[r]->comm;
        m->comm[r]->comm = nullptr;
        mcb_comm_destroy(m->comm[r]);
        if (raw) comm_destroy_raw(raw);
    }
    for (auto c : m->ctx) if (c) mcb_ctx_destroy(c);
    delete m;
}

extern "C" int mcb_multi_create(const int32_t* devices, int32_t ndev, mcb_multi** out)
{
    mcb_multi* m = nullptr;
    try {
        MCB_CHECK(out && ndev >= 1 && ndev <= 64, "mcb_multi_create: bad argument");
        m = new mcb_multi();
        m->n = ndev;
        for (int r = 0; r < ndev; ++r) m->devices.push_back(devices ? devices[r] : r);
        for (int r = 0; r < ndev; ++r)
            for (int q = 0; q < r; ++q) MCB_CHECK(m->devices[q] != m->devices[r], "mcb_multi_create: duplicate device");
        m->ctx.assign((size_t)ndev, nullptr);
        m->comm.assign((size_t)ndev, nullptr);
        m->err.assign((size_t)ndev, std::string());
        m->err_code.assign((size_t)ndev, 0);
        for (int r = 0; r < ndev; ++r) {
            if (mcb_ctx_create(m->devices[r], &m->ctx[r])) throw Error(2, mcb_last_error());
            if (!getenv("MCB_HOST_THREADS")) {                        // the devices of a session share the host's cores
                const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
                m->ctx[r]->host_threads = (int)std::max(2u, std::min(16u, hw / (unsigned)ndev));
            }
        }
        std::vector<void*> raw((size_t)ndev, nullptr);
        if (ndev > 1) comm_init_all(ndev, m->devices.data(), raw.data());
        for (int r = 0; r < ndev; ++r) {
            std::unique_ptr<mcb_comm> c(new mcb_comm());
            c->ref.set(m->ctx[r]); c->ctx = m->ctx[r]; c->rank = r; c->nranks = ndev; c->comm = raw[r]; c->owns = false;
            m->comm[r] = c.release();
        }
        for (int r = 0; r < ndev; ++r) m->threads.emplace_back(worker_loop, m, r);
        *out = m;
        return 0;
    }
    catch (const mcb::Error& e) { mcb::set_error(e.what()); mcb_multi_destroy(m); return e.code; }
    catch (const std::exception& e) { mcb::set_error(e.what()); mcb_multi_destroy(m); return 4; }
}
extern "C" int mcb_multi_size(const mcb_multi* m) { return m ? m->n : 0; }
extern "C" int mcb_multi_set_option(mcb_multi* m, const char* key, int64_t value)
{
    MCB_API_BEGIN
    MCB_CHECK(m && key, "mcb_multi_set_option: null argument");
    for (int r = 0; r < m->n; ++r)
        if (mcb_ctx_set_option(m->ctx[r], key, value)) throw Error(1, mcb_last_error());
    MCB_API_END
}
extern "C" mcb_ctx* mcb_multi_ctx(mcb_multi* m, int32_t rank) { return (m && rank >= 0 && rank < m->n) ? m->ctx[rank] : nullptr; }

// contiguous split of the cells by stored entries (so every rank uploads and transforms about the same number of bytes)
static std::vector<int64_t> split_cells(const int64_t* p, int64_t ncells, int n)
{
    std::vector<int64_t> cut((size_t)n + 1, ncells);
    cut[0] = 0;
    const int64_t nnz = p[ncells];
    int64_t c = 0;
    for (int r = 1; r < n; ++r) {
        const int64_t target = nnz / n * r;
        while (c < ncells && p[c] < target) ++c;
        cut[r] = c;
    }
    return cut;
}

// mcell_add_cgraph_from_mat_bknn's body on all devices of the session; same arguments and outputs as mcb_cgraph_bknn
extern "C" int mcb_cgraph_bknn_multi(mcb_multi* m, int32_t ngenes, int64_t ncells, const int64_t* p, const int32_t* i,
                                     const double* x, const int32_t* feat_rows, int32_t G, int32_t K, int32_t k_expand,
                                     int32_t k_beta, double k_nonz, int dsamp, uint64_t seed, int64_t* n_edges,
                                     int32_t** src, int32_t** dst, double** w, int64_t* n_valid_cells)
{
    int32_t* hs = nullptr; int32_t* hd = nullptr; double* hw = nullptr;
    int rc = 0;
    try {
        MCB_CHECK(m && n_edges && src && dst && w, "mcb_cgraph_bknn_multi: null argument");
        MCB_CHECK(p && feat_rows && ncells >= 0 && p[0] == 0, "MC-ERR: mcb_cgraph_bknn_multi: malformed matrix");
        const int P = m->n;
        const std::vector<int64_t> cut = split_cells(p, ncells, P);
        std::vector<mcb_cgraph*> g((size_t)P, nullptr);
        std::vector<int64_t> ne((size_t)P, 0);
        struct Cleanup { std::vector<mcb_cgraph*>& g; ~Cleanup() { for (auto q : g) mcb_cgraph_destroy(q); } } cleanup{g};
        std::mutex alloc_mu;
        int64_t total = 0;
        multi_run(m, [&](int r) {
            MCB_CUDA(cudaSetDevice(m->devices[r]));
            BknnJob j{ngenes, ncells, cut[r], cut[r + 1], p, i, x, feat_rows, G, K, k_expand, k_beta, k_nonz, dsamp, seed};
            // phase 1 (local: argument checks, upload, validation) under error agreement, then the collective build
            std::unique_ptr<mcb_csc> up;
            bool ok = true; Error first(0, "");
            try { up.reset(csc_upload_impl(m->ctx[r], ngenes, cut[r + 1] - cut[r], p + cut[r], i, x)); }
            catch (const Error& e) { ok = false; first = e; }
            multi_agree(m, ok);
            if (!ok) throw first;
            try { g[r] = bknn_rank_uploaded(m->ctx[r], m->comm[r], j, up.get()); ne[r] = g[r]->n_edges; }
            catch (const Error& e) { ok = false; first = e; }
            // every rank writes its rows into the shared result table: offsets from the edge counts of the lower ranks
            multi_agree(m, ok);
            if (!ok) throw first;
            {
                std::lock_guard<std::mutex> lk(alloc_mu);
                if (!hs) {
                    total = 0;
                    for (int q = 0; q < P; ++q) total += ne[q];
                    hs = (int32_t*)host_alloc(sizeof(int32_t) * (size_t)std::max<int64_t>(total, 1));
                    hd = (int32_t*)host_alloc(sizeof(int32_t) * (size_t)std::max<int64_t>(total, 1));
                    hw = (double*)host_alloc(sizeof(double) * (size_t)std::max<int64_t>(total, 1));
                }
            }
            int64_t off = 0;
            for (int q = 0; q < r; ++q) off += ne[q];
            cgraph_edges_to_host(g[r], hs + off, hd + off, hw + off);
        });
        *n_edges = total; *src = hs; *dst = hd; *w = hw;
        if (n_valid_cells) *n_valid_cells = g[0]->M;
        hs = nullptr; hd = nullptr; hw = nullptr;
    }
    catch (const mcb::Error& e) { mcb::set_error(e.what()); rc = e.code; }
    catch (const std::bad_alloc&) { mcb::set_error("host out of memory"); rc = 3; }
    catch (const std::exception& e) { mcb::set_error(e.what()); rc = 4; }
    catch (...) { mcb::set_error("unknown error"); rc = 5; }
    mcb_free(hs); mcb_free(hd); mcb_free(hw);
    return rc;
}

// mcell_coclust_from_graph_resamp's tgs_graph_cover_resample call (R/coclust_graph_cov.r:41) on all devices: resamples
// strided over the devices, graph replicated, counters reduced to device 0 (NCCL reduce), pairs extracted there.
// Outputs: library-allocated (mcb_free) node1 / node2 / cnt [n_pairs] and samples [N].
extern "C" int mcb_coclust_multi(mcb_multi* m, int32_t N, int64_t n_edges, const int32_t* src, const int32_t* dst, const double* w,
                                 int32_t n_resamp, int32_t n_s, const int32_t* samples, const uint64_t* seeds, int32_t min_mc_size,
                                 double cooling, int32_t burn_in, int32_t max_sweeps, int32_t knn, int64_t* n_pairs, int32_t** node1,
                                 int32_t** node2, int32_t** cnt, int32_t** samples_per_node)
{
    int32_t *h1 = nullptr, *h2 = nullptr, *hc = nullptr, *hn = nullptr;
    int rc = 0;
    try {
        MCB_CHECK(m && n_pairs && node1 && node2 && cnt && samples_per_node, "mcb_coclust_multi: null argument");
        const int P = m->n;
        std::vector<mcb_coclust*> c((size_t)P, nullptr);
        struct Cleanup { std::vector<mcb_coclust*>& c; ~Cleanup() { for (auto q : c) mcb_coclust_destroy(q); } } cleanup{c};
        multi_run(m, [&](int r) {
            MCB_CUDA(cudaSetDevice(m->devices[r]));
            bool ok = true; Error first(0, "");
            if (mcb_coclust_run2(m->ctx[r], N, n_edges, src, dst, w, n_resamp, n_s, samples, seeds, min_mc_size, cooling, burn_in,
                                 max_sweeps, knn, r, P, &c[r])) { ok = false; first = Error(1, mcb_last_error()); }
            multi_agree(m, ok);
            if (!ok) throw first;
            if (mcb_coclust_reduce(m->comm[r], c[r], 0)) throw Error(2, mcb_last_error());
            if (r == 0) {
                int64_t np = 0;
                if (mcb_coclust_extract(c[0], &np)) throw Error(2, mcb_last_error());
                h1 = (int32_t*)host_alloc(sizeof(int32_t) * (size_t)std::max<int64_t>(np, 1));
                h2 = (int32_t*)host_alloc(sizeof(int32_t) * (size_t)std::max<int64_t>(np, 1));
                hc = (int32_t*)host_alloc(sizeof(int32_t) * (size_t)std::max<int64_t>(np, 1));
                hn = (int32_t*)host_alloc(sizeof(int32_t) * (size_t)std::max<int32_t>(N, 1));
                if (mcb_coclust_pairs(c[0], h1, h2, hc, hn)) throw Error(2, mcb_last_error());
                *n_pairs = np;
            }
        });
        *node1 = h1; *node2 = h2; *cnt = hc; *samples_per_node = hn;
        h1 = h2 = hc = hn = nullptr;
    }
    catch (const mcb::Error& e) { mcb::set_error(e.what()); rc = e.code; }
    catch (const std::bad_alloc&) { mcb::set_error("host out of memory"); rc = 3; }
    catch (const std::exception& e) { mcb::set_error(e.what()); rc = 4; }
    catch (...) { mcb::set_error("unknown error"); rc = 5; }
    mcb_free(h1); mcb_free(h2); mcb_free(hc); mcb_free(hn);
    return rc;
}
