// coclust.cu -- tgs_graph_cover_resample(..., method = "full") (reference call site
// R/coclust_graph_cov.r:41): n_resamp independent graph covers on host-supplied node samples, then
// integer co-occurrence counts cnt[a][b] (a < b) and per-node sample counts (fields co_cluster /
// samples of the result, R/coclust.r:39-41, R/coclust_graph_cov.r:55).
//
// The cover algorithm is the one oracle/mc_oracle.c orc_graph_cover states (MetaCell paper Methods,
// SURVEY.md App. A7), reproduced operation for operation so the assignment is bit-identical:
// fixed-point integer edge weights, f^3-weighted seeding driven by Philox draws, Gauss-Seidel
// reassignment sweeps in a keyed Feistel permutation with fp64 scores w_out*w_in/size^2 and a
// cooling factor on the current cluster.
//
// Mapping: one CTA per resample (resamples are the parallel axis: 500-1000 >= 148 SMs); inside a
// resample the nodes are visited sequentially and the CTA's threads split the <= K + k_beta*K edges
// of the visited node.  Cluster accumulators live in shared memory; mc[] / f[] are per-resample
// global arrays that stay L1/L2 resident.  Latency-bound, neither roofline binds.
#include "internal.h"
#include <cstdlib>

namespace mcb {

constexpr int CC_THREADS = 128;
constexpr int CC_WARPS = CC_THREADS / 32;

__device__ __forceinline__ uint64_t cover_rng_u64(uint32_t seed_lo, uint32_t seed_hi, uint32_t stream, uint32_t a, uint32_t b)
{
    uint32_t o[4];
    philox4x32_10(a, b, stream, TAG_COVER, seed_lo, seed_hi, o);
    return (uint64_t)o[0] | ((uint64_t)o[1] << 32);
}
__device__ __forceinline__ uint32_t feistel_f(uint32_t x, uint32_t k)
{
    x ^= k; x *= 0x9E3779B1u; x ^= x >> 15; x *= 0x85EBCA77u; x ^= x >> 13; return x;
}
__device__ __forceinline__ uint32_t perm_index(uint32_t t, uint32_t n, uint32_t hb, const uint32_t rk[4])
{
    const uint32_t mask = (1u << hb) - 1u;
    uint32_t x = t;
    do {
        uint32_t L = x >> hb, R = x & mask;
#pragma unroll
        for (int r = 0; r < 4; ++r) { const uint32_t nl = R; R = L ^ (feistel_f(R, rk[r]) & mask); L = nl; }
        x = (L << hb) | R;
    } while (x >= n);
    return x;
}

struct Best { double sc; int k; };
__device__ __forceinline__ bool better(double sc, int k, double bsc, int bk)
{
    return bk < 0 || sc > bsc || (sc == bsc && k < bk);
}

__global__ void __launch_bounds__(CC_THREADS)
graph_cover_kernel(CoverGraph g, int32_t n_local, const int32_t* __restrict__ resamp_ids, int32_t n_s,
                   const int32_t* __restrict__ samples, const uint64_t* __restrict__ seeds, int32_t min_size,
                   double cooling, int32_t burn_in, int32_t max_sweeps, int32_t ncl_cap, int32_t* __restrict__ mc_all,
                   int32_t* __restrict__ f_all, int32_t* __restrict__ sweeps_out, int32_t* __restrict__ status)
{
    extern __shared__ uint32_t dsm[];
    int32_t* size = (int32_t*)dsm;                  // [ncl_cap]
    uint32_t* wout = dsm + ncl_cap;                 // [ncl_cap]
    uint32_t* win = dsm + 2 * ncl_cap;              // [ncl_cap]
    __shared__ unsigned long long s_part[CC_THREADS];
    __shared__ double s_bsc[CC_WARPS];
    __shared__ int s_bk[CC_WARPS];
    __shared__ int s_pick, s_flag;
    __shared__ unsigned long long s_total;

    const int N = g.N;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int lr = blockIdx.x; lr < n_local; lr += gridDim.x) {
        const int rid = resamp_ids[lr];
        const int32_t* S = samples + (int64_t)rid * n_s;
        const uint64_t seed = seeds[rid];
        const uint32_t seed_lo = (uint32_t)seed, seed_hi = (uint32_t)(seed >> 32);
        int32_t* mc = mc_all + (int64_t)lr * N;
        int32_t* f = f_all + (int64_t)lr * N;

        for (int v = tid; v < N; v += CC_THREADS) { mc[v] = -2; f[v] = 0; }
        for (int k = tid; k < ncl_cap; k += CC_THREADS) { size[k] = 0; wout[k] = 0; win[k] = 0; }
        __syncthreads();
        for (int q = tid; q < n_s; q += CC_THREADS) mc[S[q]] = -1;
        __syncthreads();
        for (int q = tid; q < n_s; q += CC_THREADS) {
            const int v = S[q];
            int c = 0;
            for (int64_t e = g.optr[v]; e < g.optr[v + 1]; ++e) c += (mc[g.odst[e]] == -1);
            f[v] = c;
        }
        __syncthreads();

        // ---- seeding -------------------------------------------------------------------------
        int ncl = 0;
        int err = 0;
        const int chunk = (n_s + CC_THREADS - 1) / CC_THREADS;
        const int q0 = tid * chunk, q1 = min(n_s, q0 + chunk);
        for (uint32_t iter = 0;; ++iter) {
            unsigned long long part = 0;
            for (int q = q0; q < q1; ++q) {
                const int v = S[q];
                const int fv = f[v];
                if (mc[v] == -1 && fv >= min_size && fv > 0) part += (unsigned long long)fv * fv * fv;
            }
            s_part[tid] = part;
            __syncthreads();
            if (tid == 0) {
                unsigned long long tot = 0;
                for (int q = 0; q < CC_THREADS; ++q) tot += s_part[q];
                s_total = tot;
                s_pick = -1;
            }
            __syncthreads();
            const unsigned long long total = s_total;
            if (total == 0) break;
            if (ncl >= ncl_cap) { err = 3; break; }
            const unsigned long long r = cover_rng_u64(seed_lo, seed_hi, STREAM_SEED, iter, 0) % total;
            unsigned long long excl = 0;
            for (int q = 0; q < tid; ++q) excl += s_part[q];
            if (r >= excl && r < excl + part) {
                unsigned long long cum = excl;
                for (int q = q0; q < q1; ++q) {
                    const int v = S[q];
                    const int fv = f[v];
                    if (mc[v] == -1 && fv >= min_size && fv > 0) {
                        cum += (unsigned long long)fv * fv * fv;
                        if (cum > r) { s_pick = v; break; }
                    }
                }
            }
            __syncthreads();
            const int pick = s_pick;
            const int k = ncl++;
            if (tid == 0) { mc[pick] = k; s_flag = 1; }
            __syncthreads();
            // claim the uncovered sampled out-neighbours (edges are unique per (src, dst))
            const int64_t oe0 = g.optr[pick], oe1 = g.optr[pick + 1];
            for (int64_t e = oe0 + tid; e < oe1; e += CC_THREADS) {
                const int u = g.odst[e];
                if (mc[u] == -1) { mc[u] = k; atomicAdd(&s_flag, 1); }
            }
            __syncthreads();
            if (tid == 0) size[k] = s_flag;
            // every newly covered node lowers f of its in-neighbours: one warp per covered node
            for (int64_t e = oe0 - 1 + wid; e < oe1; e += CC_WARPS) {
                const int u = (e < oe0) ? pick : g.odst[e];
                // u was claimed in this iteration iff mc[u] == k; the seed itself is handled at e = oe0-1.
                // A self edge pick -> pick must not be counted twice.
                if (mc[u] != k || (e >= oe0 && u == pick)) continue;
                for (int64_t e2 = g.iptr[u] + lane; e2 < g.iptr[u + 1]; e2 += 32) atomicSub(&f[g.isrc[e2]], 1);
            }
            __syncthreads();
        }
        if (err) {
            if (tid == 0) { status[lr] = err; sweeps_out[lr] = 0; }
            __syncthreads();
            continue;
        }

        // ---- reassignment sweeps --------------------------------------------------------------
        uint32_t hb = 1;
        while ((1ull << (2 * hb)) < (unsigned long long)n_s) ++hb;
        double factor = 1.0;
        int sweep = 0;
        for (; sweep < max_sweeps; ++sweep) {
            if (sweep > burn_in) factor = factor * cooling;
            uint32_t rk[4];
            {
                const uint64_t w0 = cover_rng_u64(seed_lo, seed_hi, STREAM_PERM, (uint32_t)sweep, 0u);
                const uint64_t w1 = cover_rng_u64(seed_lo, seed_hi, STREAM_PERM, (uint32_t)sweep, 1u);
                rk[0] = (uint32_t)w0; rk[1] = (uint32_t)(w0 >> 32); rk[2] = (uint32_t)w1; rk[3] = (uint32_t)(w1 >> 32);
            }
            int moved = 0;                  // tracked identically by every thread
            int pend_v = -1, pend_k = -1;   // assignment made in the previous step (visible after the next barrier)
            for (int t = 0; t < n_s; ++t) {
                const int v = S[perm_index((uint32_t)t, (uint32_t)n_s, hb, rk)];
                const int cur = (v == pend_v) ? pend_k : mc[v];
                const int64_t oe0 = g.optr[v], oe1 = g.optr[v + 1];
                const int64_t ie0 = g.iptr[v], ie1 = g.iptr[v + 1];
                // [A] accumulate association towards every neighbouring cluster
                for (int64_t e = oe0 + tid; e < oe1; e += CC_THREADS) {
                    const int u = g.odst[e];
                    const int k = (u == pend_v) ? pend_k : mc[u];
                    if (k >= 0) atomicAdd(&wout[k], g.ow[e]);
                }
                for (int64_t e = ie0 + tid; e < ie1; e += CC_THREADS) {
                    const int u = g.isrc[e];
                    const int k = (u == pend_v) ? pend_k : mc[u];
                    if (k >= 0) atomicAdd(&win[k], g.iw[e]);
                }
                __syncthreads();
                // [B] score the candidate clusters (those reached by an out edge) and take the argmax
                double bsc = 0.0; int bk = -1;
                for (int64_t e = oe0 + tid; e < oe1; e += CC_THREADS) {
                    const int k = mc[g.odst[e]];
                    if (k < 0) continue;
                    const uint32_t wo = wout[k], wi = win[k];
                    if (wo == 0 || wi == 0) continue;
                    const double sz = (double)size[k];
                    double sc = (double)((unsigned long long)wo * (unsigned long long)wi) / (sz * sz);
                    if (k == cur) sc = sc * factor;
                    if (better(sc, k, bsc, bk)) { bsc = sc; bk = k; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double osc = __shfl_xor_sync(0xffffffffu, bsc, o);
                    const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
                    if (ok >= 0 && better(osc, ok, bsc, bk)) { bsc = osc; bk = ok; }
                }
                if (lane == 0) { s_bsc[wid] = bsc; s_bk[wid] = bk; }
                __syncthreads();
                bsc = 0.0; bk = -1;
#pragma unroll
                for (int w = 0; w < CC_WARPS; ++w) {
                    const double osc = s_bsc[w]; const int ok = s_bk[w];
                    if (ok >= 0 && better(osc, ok, bsc, bk)) { bsc = osc; bk = ok; }
                }
                // [C] clear the touched accumulators
                for (int64_t e = oe0 + tid; e < oe1; e += CC_THREADS) { const int k = mc[g.odst[e]]; if (k >= 0) wout[k] = 0; }
                for (int64_t e = ie0 + tid; e < ie1; e += CC_THREADS) { const int k = mc[g.isrc[e]]; if (k >= 0) win[k] = 0; }
                __syncthreads();
                // [D] apply the move; other threads see it through (pend_v, pend_k) until the next barrier
                pend_v = -1;
                if (bk >= 0 && bk != cur) {
                    if (tid == 0) {
                        if (cur >= 0) size[cur]--;
                        size[bk]++;
                        mc[v] = bk;
                    }
                    pend_v = v; pend_k = bk;
                    ++moved;
                }
            }
            __syncthreads();
            if (moved == 0) { ++sweep; break; }
        }
        if (tid == 0) { sweeps_out[lr] = sweep; status[lr] = 0; }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Fast variant used whenever the per-resample state fits in shared memory: mc[] (int16) and the sweep's
// visit order live in shared memory, the visited node's edges sit in registers (one out edge and up to
// IN_SLOTS in edges per thread; graphs with larger degrees use graph_cover_kernel), and the CSR offsets / edges of the
// next two nodes are prefetched while the current node is processed, so no global-memory latency is left
// on the per-node critical path (the first version spent ~8,000 cycles per node visit in dependent L2
// loads).  Cluster accumulators are double buffered by visit parity, which removes one of the three
// block barriers per visit.  Same arithmetic as graph_cover_kernel / orc_graph_cover: bit-identical.
// ---------------------------------------------------------------------------------------------
struct FastGraph {      // CoverGraph with 32-bit CSR offsets (n_edges < 2^31), fewer integer instructions per visit
    int32_t N; const int32_t* optr; const int32_t* odst; const uint32_t* ow;
    const int32_t* iptr; const int32_t* isrc; const uint32_t* iw;
};
template <int IN_SLOTS> struct NodeOffs { int v, o0, o1, i0, i1; };
template <int IN_SLOTS> struct NodeEdges { int v; int uo; uint32_t wo; int ui[IN_SLOTS]; uint32_t wi[IN_SLOTS]; };

template <int IN_SLOTS>
__device__ __forceinline__ NodeOffs<IN_SLOTS> load_offs(const FastGraph& g, const int32_t* vlist, int t, int n_s)
{
    NodeOffs<IN_SLOTS> o;
    o.v = -1; o.o0 = o.o1 = o.i0 = o.i1 = 0;
    if (t < n_s) {
        o.v = vlist[t];
        o.o0 = g.optr[o.v]; o.o1 = g.optr[o.v + 1];
        o.i0 = g.iptr[o.v]; o.i1 = g.iptr[o.v + 1];
    }
    return o;
}
template <int FT, int IN_SLOTS>
__device__ __forceinline__ NodeEdges<IN_SLOTS> load_edges(const FastGraph& g, const NodeOffs<IN_SLOTS>& o, int tid)
{
    NodeEdges<IN_SLOTS> e;
    e.v = o.v;
    const int eo = o.o0 + tid;
    const bool ho = eo < o.o1;                  // o.v < 0 has o0 == o1 == 0
    e.uo = ho ? g.odst[eo] : -1;
    e.wo = ho ? g.ow[eo] : 0u;
#pragma unroll
    for (int s = 0; s < IN_SLOTS; ++s) {
        const int ei = o.i0 + s * FT + tid;
        const bool hi = ei < o.i1;
        e.ui[s] = hi ? g.isrc[ei] : -1;
        e.wi[s] = hi ? g.iw[ei] : 0u;
    }
    return e;
}

// requires (checked on the host): out-degree <= FT and in-degree <= IN_SLOTS * FT for every node
template <int FT, int IN_SLOTS>
__global__ void __launch_bounds__(FT)
graph_cover_fast_kernel(FastGraph g, int32_t n_local, const int32_t* __restrict__ resamp_ids, int32_t n_s,
                        const int32_t* __restrict__ samples, const uint64_t* __restrict__ seeds, int32_t min_size,
                        double cooling, int32_t burn_in, int32_t max_sweeps, int32_t ncl_cap, int32_t out_warps,
                        int32_t* __restrict__ mc_all, int32_t* __restrict__ f_all, int32_t* __restrict__ sweeps_out,
                        int32_t* __restrict__ status)
{
    constexpr int FW = FT / 32;
    extern __shared__ uint32_t dsm[];
    int32_t* size = (int32_t*)dsm;                  // [ncl_cap]
    uint32_t* wout = dsm + ncl_cap;                 // [2][ncl_cap]
    uint32_t* win = dsm + 3 * ncl_cap;              // [2][ncl_cap]
    int32_t* vlist = (int32_t*)(dsm + 5 * ncl_cap); // [n_s] visit order of the current sweep
    int16_t* mc_s = (int16_t*)(vlist + n_s);        // [N]
    __shared__ unsigned long long s_part[FT];
    __shared__ double s_bsc[FW];
    __shared__ int s_bk[FW];
    __shared__ int s_pick, s_flag;
    __shared__ unsigned long long s_total;

    const int N = g.N;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int lr = blockIdx.x; lr < n_local; lr += gridDim.x) {
        const int rid = resamp_ids[lr];
        const int32_t* S = samples + (int64_t)rid * n_s;
        const uint64_t seed = seeds[rid];
        const uint32_t seed_lo = (uint32_t)seed, seed_hi = (uint32_t)(seed >> 32);
        int32_t* f = f_all + (int64_t)lr * N;

        for (int v = tid; v < N; v += FT) { mc_s[v] = -2; f[v] = 0; }
        for (int k = tid; k < ncl_cap; k += FT) { size[k] = 0; wout[k] = 0; wout[ncl_cap + k] = 0; win[k] = 0; win[ncl_cap + k] = 0; }
        if (tid < FW) { s_bsc[tid] = 0.0; s_bk[tid] = -1; }
        __syncthreads();
        for (int q = tid; q < n_s; q += FT) mc_s[S[q]] = -1;
        __syncthreads();
        for (int q = tid; q < n_s; q += FT) {
            const int v = S[q];
            int c = 0;
            for (int e = g.optr[v]; e < g.optr[v + 1]; ++e) c += (mc_s[g.odst[e]] == -1);
            f[v] = c;
        }
        __syncthreads();

        // ---- seeding (as graph_cover_kernel) ------------------------------------------------------
        int ncl = 0;
        int err = 0;
        const int chunk = (n_s + FT - 1) / FT;
        const int q0 = tid * chunk, q1 = min(n_s, q0 + chunk);
        for (uint32_t iter = 0;; ++iter) {
            unsigned long long part = 0;
            for (int q = q0; q < q1; ++q) {
                const int v = S[q];
                const int fv = f[v];
                if (mc_s[v] == -1 && fv >= min_size && fv > 0) part += (unsigned long long)fv * fv * fv;
            }
            s_part[tid] = part;
            __syncthreads();
            if (tid < 32) {        // block total by one warp
                unsigned long long tot = 0;
                for (int q = lane; q < FT; q += 32) tot += s_part[q];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
                if (lane == 0) { s_total = tot; s_pick = -1; }
            }
            __syncthreads();
            const unsigned long long total = s_total;
            if (total == 0) break;
            if (ncl >= ncl_cap) { err = 3; break; }
            const unsigned long long r = cover_rng_u64(seed_lo, seed_hi, STREAM_SEED, iter, 0) % total;
            unsigned long long excl = 0;
            for (int q = 0; q < tid; ++q) excl += s_part[q];
            if (r >= excl && r < excl + part) {
                unsigned long long cum = excl;
                for (int q = q0; q < q1; ++q) {
                    const int v = S[q];
                    const int fv = f[v];
                    if (mc_s[v] == -1 && fv >= min_size && fv > 0) {
                        cum += (unsigned long long)fv * fv * fv;
                        if (cum > r) { s_pick = v; break; }
                    }
                }
            }
            __syncthreads();
            const int pick = s_pick;
            const int k = ncl++;
            if (tid == 0) { mc_s[pick] = (int16_t)k; s_flag = 1; }
            __syncthreads();
            const int oe0 = g.optr[pick], oe1 = g.optr[pick + 1];
            for (int e = oe0 + tid; e < oe1; e += FT) {
                const int u = g.odst[e];
                if (mc_s[u] == -1) { mc_s[u] = (int16_t)k; atomicAdd(&s_flag, 1); }
            }
            __syncthreads();
            if (tid == 0) size[k] = s_flag;
            for (int e = oe0 - 1 + wid; e < oe1; e += FW) {
                const int u = (e < oe0) ? pick : g.odst[e];
                if (mc_s[u] != k || (e >= oe0 && u == pick)) continue;
                for (int e2 = g.iptr[u] + lane; e2 < g.iptr[u + 1]; e2 += 32) atomicSub(&f[g.isrc[e2]], 1);
            }
            __syncthreads();
        }
        if (err) {
            if (tid == 0) { status[lr] = err; sweeps_out[lr] = 0; }
            for (int v = tid; v < N; v += FT) mc_all[(int64_t)lr * N + v] = mc_s[v];
            __syncthreads();
            continue;
        }

        // ---- reassignment sweeps ------------------------------------------------------------------
        uint32_t hb = 1;
        while ((1ull << (2 * hb)) < (unsigned long long)n_s) ++hb;
        double factor = 1.0;
        int sweep = 0;
        for (; sweep < max_sweeps; ++sweep) {
            if (sweep > burn_in) factor = factor * cooling;
            uint32_t rk[4];
            {
                const uint64_t w0 = cover_rng_u64(seed_lo, seed_hi, STREAM_PERM, (uint32_t)sweep, 0u);
                const uint64_t w1 = cover_rng_u64(seed_lo, seed_hi, STREAM_PERM, (uint32_t)sweep, 1u);
                rk[0] = (uint32_t)w0; rk[1] = (uint32_t)(w0 >> 32); rk[2] = (uint32_t)w1; rk[3] = (uint32_t)(w1 >> 32);
            }
            __syncthreads();                                         // last visit of the previous sweep is done with vlist
            for (int t = tid; t < n_s; t += FT) vlist[t] = S[perm_index((uint32_t)t, (uint32_t)n_s, hb, rk)];
            __syncthreads();
            int moved = 0;
            int pend_v = -1, pend_k = -1;
            NodeEdges<IN_SLOTS> cur = load_edges<FT, IN_SLOTS>(g, load_offs<IN_SLOTS>(g, vlist, 0, n_s), tid);
            NodeOffs<IN_SLOTS> nxt = load_offs<IN_SLOTS>(g, vlist, 1, n_s);
            for (int t = 0; t < n_s; ++t) {
                const NodeEdges<IN_SLOTS> pre = load_edges<FT, IN_SLOTS>(g, nxt, tid);   // node t + 1, in flight during this visit
                nxt = load_offs<IN_SLOTS>(g, vlist, t + 2, n_s);                          // node t + 2
                const int v = cur.v;
                const int curk = (v == pend_v) ? pend_k : (int)mc_s[v];
                uint32_t* wo_b = wout + (t & 1) * ncl_cap;
                uint32_t* wi_b = win + (t & 1) * ncl_cap;
                // [A] association towards every neighbouring cluster
                int ko = -1, ki[IN_SLOTS];
                if (cur.uo >= 0) { ko = (cur.uo == pend_v) ? pend_k : (int)mc_s[cur.uo]; if (ko >= 0) atomicAdd(&wo_b[ko], cur.wo); }
#pragma unroll
                for (int s = 0; s < IN_SLOTS; ++s) {
                    ki[s] = -1;
                    if (cur.ui[s] >= 0) { ki[s] = (cur.ui[s] == pend_v) ? pend_k : (int)mc_s[cur.ui[s]]; if (ki[s] >= 0) atomicAdd(&wi_b[ki[s]], cur.wi[s]); }
                }
                __syncthreads();
                // [B] scores of the clusters reached by an out edge (only the first out_warps warps hold out edges)
                if (wid < out_warps) {
                    double bsc = 0.0; int bk = -1;
                    if (ko >= 0) {
                        const uint32_t wo = wo_b[ko], wi = wi_b[ko];
                        if (wo != 0 && wi != 0) {
                            const double sz = (double)size[ko];
                            double sc = (double)((unsigned long long)wo * (unsigned long long)wi) / (sz * sz);
                            if (ko == curk) sc = sc * factor;
                            bsc = sc; bk = ko;
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const double osc = __shfl_xor_sync(0xffffffffu, bsc, o);
                        const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
                        if (ok >= 0 && better(osc, ok, bsc, bk)) { bsc = osc; bk = ok; }
                    }
                    if (lane == 0) { s_bsc[wid] = bsc; s_bk[wid] = bk; }
                }
                __syncthreads();
                double bsc = 0.0; int bk = -1;
                for (int w = 0; w < out_warps; ++w) {
                    const double osc = s_bsc[w]; const int ok = s_bk[w];
                    if (ok >= 0 && better(osc, ok, bsc, bk)) { bsc = osc; bk = ok; }
                }
                // [C] clear this visit's accumulator buffer (it is next used two visits from now)
                if (ko >= 0) wo_b[ko] = 0;
#pragma unroll
                for (int s = 0; s < IN_SLOTS; ++s) if (ki[s] >= 0) wi_b[ki[s]] = 0;
                // [D] apply the move; the other threads see it through (pend_v, pend_k) until the next barrier
                pend_v = -1;
                if (bk >= 0 && bk != curk) {
                    if (tid == 0) {
                        if (curk >= 0) size[curk]--;
                        size[bk]++;
                        mc_s[v] = (int16_t)bk;
                    }
                    pend_v = v; pend_k = bk;
                    ++moved;
                }
                cur = pre;
            }
            if (moved == 0) { ++sweep; break; }
        }
        __syncthreads();
        for (int v = tid; v < N; v += FT) mc_all[(int64_t)lr * N + v] = mc_s[v];
        if (tid == 0) { sweeps_out[lr] = sweep; status[lr] = 0; }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// co-occurrence: for every cluster of every local resample, cnt[min(a,b)][max(a,b)] += 1 for all
// member pairs; nsamp[a] += 1 for every sampled node (mc[a] != -2).
// One CTA per resample; members are grouped by a counting sort in global workspace.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cooccur_kernel(int32_t N, int32_t n_local, const int32_t* __restrict__ mc_all, int32_t ncl_cap,
               uint32_t* __restrict__ cnt, int32_t* __restrict__ nsamp, int32_t* __restrict__ ws_order,
               int32_t* __restrict__ ws_start)
{
    extern __shared__ int32_t sfill[];     // [ncl_cap + 1] cursor
    for (int lr = blockIdx.x; lr < n_local; lr += gridDim.x) {
        const int32_t* mc = mc_all + (int64_t)lr * N;
        int32_t* order = ws_order + (int64_t)lr * N;
        int32_t* start = ws_start + (int64_t)lr * (ncl_cap + 1);
        for (int k = threadIdx.x; k <= ncl_cap; k += blockDim.x) sfill[k] = 0;
        __syncthreads();
        for (int v = threadIdx.x; v < N; v += blockDim.x) {
            const int k = mc[v];
            if (k != -2) atomicAdd(&nsamp[v], 1);
            if (k >= 0) atomicAdd(&sfill[k + 1], 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int run = 0;
            for (int k = 0; k <= ncl_cap; ++k) { run += sfill[k]; start[k] = run; sfill[k] = run; }
        }
        __syncthreads();
        // sfill[k] now = start of cluster k (exclusive prefix incl. shift); use as cursor
        for (int v = threadIdx.x; v < N; v += blockDim.x) {
            const int k = mc[v];
            if (k >= 0) { const int pos = atomicAdd(&sfill[k], 1); order[pos] = v; }
        }
        __syncthreads();
        const int total = start[ncl_cap];
        // one thread per member a; pairs with the later members of the same cluster
        for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
            const int a = order[idx];
            const int k = mc[a];
            const int end = start[k + 1];
            for (int j = idx + 1; j < end; ++j) {
                const int b = order[j];
                const int lo = a < b ? a : b, hi = a < b ? b : a;
                atomicAdd(&cnt[(int64_t)lo * N + hi], 1u);
            }
        }
        __syncthreads();
    }
}

// row_counts[a] = #{b > a : cnt[a][b] > 0}
__global__ void count_pairs_kernel(int32_t N, const uint32_t* __restrict__ cnt, int64_t* __restrict__ row_counts)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t a = warp; a < N; a += nwarps) {
        int c = 0;
        for (int64_t b = a + 1 + lane; b < N; b += 32) c += cnt[a * N + b] != 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) row_counts[a] = c;
    }
}
// ordered extraction: one warp per row, ballot-compacted so rows stay sorted by node2
__global__ void extract_pairs_kernel(int32_t N, const uint32_t* __restrict__ cnt, const int64_t* __restrict__ row_offs,
                                     int32_t* __restrict__ n1, int32_t* __restrict__ n2, int32_t* __restrict__ c)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t a = warp; a < N; a += nwarps) {
        int64_t pos = row_offs[a];
        for (int64_t b0 = a + 1; b0 < N; b0 += 32) {
            const int64_t b = b0 + lane;
            const uint32_t v = b < N ? cnt[a * N + b] : 0u;
            const unsigned m = __ballot_sync(0xffffffffu, v != 0);
            if (v) {
                const int64_t o = pos + __popc(m & ((1u << lane) - 1u));
                n1[o] = (int32_t)a; n2[o] = (int32_t)b; c[o] = (int32_t)v;
            }
            pos += __popc(m);
        }
    }
}

template <int FT, int IN_SLOTS>
static void launch_cover_fast(mcb_ctx* ctx, const FastGraph& fg, int32_t n_local, const int32_t* resamp_ids, int32_t n_s,
                              const int32_t* samples, const uint64_t* seeds, int32_t min_size, double cooling, int32_t burn_in,
                              int32_t max_sweeps, int32_t ncl_cap, int32_t max_out, size_t smem, int32_t* mc, int32_t* f_ws,
                              int32_t* sweeps, int32_t* status)
{
    static size_t configured = 0;
    if (smem > configured) {
        MCB_CUDA(cudaFuncSetAttribute(graph_cover_fast_kernel<FT, IN_SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int32_t out_warps = std::max(1, (int)ceil_div(max_out, 32));
    graph_cover_fast_kernel<FT, IN_SLOTS><<<n_local, FT, smem, ctx->stream>>>(fg, n_local, resamp_ids, n_s, samples, seeds,
                                                                               min_size, cooling, burn_in, max_sweeps, ncl_cap,
                                                                               out_warps, mc, f_ws, sweeps, status);
    MCB_LAUNCH_CHECK();
}

void launch_graph_cover(mcb_ctx* ctx, const CoverGraph& g, const int32_t* optr32, const int32_t* iptr32, int32_t max_out,
                        int32_t max_in, int32_t n_local, const int32_t* resamp_ids, int32_t n_s,
                        const int32_t* samples, const uint64_t* seeds, int32_t min_size, double cooling,
                        int32_t burn_in, int32_t max_sweeps, int32_t ncl_cap, int32_t* mc, int32_t* f_ws,
                        int32_t* sweeps, int32_t* status)
{
    if (n_local == 0) return;
    // fast kernels: size + 2 x (wout, win) + visit order + int16 mc[] in shared memory, edges of a node in registers
    const size_t fast_smem = ((size_t)5 * ncl_cap + (size_t)n_s) * sizeof(uint32_t) + (size_t)g.N * sizeof(int16_t) + 16;
    if (optr32 && fast_smem <= 200 * 1024 && ncl_cap < 32000 && !getenv("MCB_COVER_SLOW")) {
        const FastGraph fg{g.N, optr32, g.odst, g.ow, iptr32, g.isrc, g.iw};
#define FAST(FT, SL) launch_cover_fast<FT, SL>(ctx, fg, n_local, resamp_ids, n_s, samples, seeds, min_size, cooling, burn_in, \
                                               max_sweeps, ncl_cap, max_out, fast_smem, mc, f_ws, sweeps, status)
        if (max_out <= 128 && max_in <= 3 * 128) { FAST(128, 3); return; }
        if (max_out <= 256 && max_in <= 3 * 256) { FAST(256, 3); return; }
        if (max_out <= 256 && max_in <= 6 * 256) { FAST(256, 6); return; }
#undef FAST
    }
    const size_t smem = (size_t)ncl_cap * 3 * sizeof(uint32_t);
    MCB_CHECK(smem <= 200 * 1024, "MC-ERR: graph cover: too many potential clusters for shared memory");
    static size_t configured = 0;
    if (smem > configured) {
        MCB_CUDA(cudaFuncSetAttribute(graph_cover_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    graph_cover_kernel<<<n_local, CC_THREADS, smem, ctx->stream>>>(g, n_local, resamp_ids, n_s, samples, seeds, min_size,
                                                                    cooling, burn_in, max_sweeps, ncl_cap, mc, f_ws,
                                                                    sweeps, status);
    MCB_LAUNCH_CHECK();
}

void launch_cooccur(mcb_ctx* ctx, int32_t N, int32_t n_local, const int32_t* mc, int32_t ncl_cap, uint32_t* cnt,
                    int32_t* nsamp, int32_t* ws_order, int32_t* ws_start)
{
    if (n_local == 0) return;
    const size_t smem = (size_t)(ncl_cap + 1) * sizeof(int32_t);
    static size_t configured = 0;
    if (smem > configured && smem > 48 * 1024) {
        MCB_CUDA(cudaFuncSetAttribute(cooccur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    cooccur_kernel<<<n_local, 256, smem, ctx->stream>>>(N, n_local, mc, ncl_cap, cnt, nsamp, ws_order, ws_start);
    MCB_LAUNCH_CHECK();
}

void launch_count_pairs(mcb_ctx* ctx, int32_t N, const uint32_t* cnt, int64_t* row_counts)
{
    if (N == 0) return;
    int grid = (int)std::min<int64_t>(ceil_div(N, 8), (int64_t)ctx->num_sms * 16);
    count_pairs_kernel<<<grid, 256, 0, ctx->stream>>>(N, cnt, row_counts);
    MCB_LAUNCH_CHECK();
}
void launch_extract_pairs(mcb_ctx* ctx, int32_t N, const uint32_t* cnt, const int64_t* row_offs, int32_t* n1,
                          int32_t* n2, int32_t* c)
{
    if (N == 0) return;
    int grid = (int)std::min<int64_t>(ceil_div(N, 8), (int64_t)ctx->num_sms * 16);
    extract_pairs_kernel<<<grid, 256, 0, ctx->stream>>>(N, cnt, row_offs, n1, n2, c);
    MCB_LAUNCH_CHECK();
}

}  // namespace mcb
