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
#include <string>

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
                   int32_t* __restrict__ f_all, int32_t* __restrict__ sweeps_out, int32_t* __restrict__ status,
                   int32_t knn, const uint16_t* __restrict__ ipos, uint16_t* __restrict__ ocut_all, uint32_t* __restrict__ acc_ws)
{
    // acc_ws != null: the cluster accumulators (size, wout, win: 3 * ncl_cap words per CTA) live in global memory because
    // they exceed shared memory (10^6-node graphs with small clusters)
    // knn > 0 (`knn` of tgs_graph_cover_resample): a node uses only its first knn out edges, in (weight descending,
    // target ascending) order, among those whose target is sampled.  Out segments are stored in that order, so the
    // usable part of node v's segment is its first ocut[v] positions; the in edge u -> v is usable iff its position in
    // u's out segment, ipos[e], is below ocut[u].
    extern __shared__ uint32_t dsm_[];
    uint32_t* dsm = acc_ws ? acc_ws + (size_t)blockIdx.x * 3 * ncl_cap : dsm_;
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
        uint16_t* ocut = knn > 0 ? ocut_all + (int64_t)lr * N : nullptr;

        for (int v = tid; v < N; v += CC_THREADS) { mc[v] = -2; f[v] = 0; }
        for (int k = tid; k < ncl_cap; k += CC_THREADS) { size[k] = 0; wout[k] = 0; win[k] = 0; }
        __syncthreads();
        for (int q = tid; q < n_s; q += CC_THREADS) mc[S[q]] = -1;
        __syncthreads();
        for (int q = tid; q < n_s; q += CC_THREADS) {
            const int v = S[q];
            int c = 0;
            int64_t e = g.optr[v];
            for (; e < g.optr[v + 1]; ++e) {
                if (mc[g.odst[e]] == -1) { if (knn > 0 && c == knn) break; ++c; }
            }
            f[v] = c;
            if (knn > 0) ocut[v] = (uint16_t)(e - g.optr[v]);
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
            const int64_t oe0 = g.optr[pick], oe1 = knn > 0 ? oe0 + ocut[pick] : g.optr[pick + 1];
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
                for (int64_t e2 = g.iptr[u] + lane; e2 < g.iptr[u + 1]; e2 += 32) {
                    const int sn = g.isrc[e2];
                    if (knn > 0 && ipos[e2] >= ocut[sn]) continue;
                    atomicSub(&f[sn], 1);
                }
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
                const int64_t oe0 = g.optr[v], oe1 = knn > 0 ? oe0 + ocut[v] : g.optr[v + 1];
                const int64_t ie0 = g.iptr[v], ie1 = g.iptr[v + 1];
                // [A] accumulate association towards every neighbouring cluster
                for (int64_t e = oe0 + tid; e < oe1; e += CC_THREADS) {
                    const int u = g.odst[e];
                    const int k = (u == pend_v) ? pend_k : mc[u];
                    if (k >= 0) atomicAdd(&wout[k], g.ow[e]);
                }
                for (int64_t e = ie0 + tid; e < ie1; e += CC_THREADS) {
                    const int u = g.isrc[e];
                    if (knn > 0 && ipos[e] >= ocut[u]) continue;
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
                for (int64_t e = ie0 + tid; e < ie1; e += CC_THREADS) { const int k = mc[g.isrc[e]]; if (k >= 0) win[k] = 0; }   // (clearing an untouched slot is harmless)
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
    // packed copies for the speculative batch kernel: one 16-byte descriptor {o0, o1, i0, i1} per node and one 8-byte
    // {neighbour, weight} per edge, so a visit's prefetch is 1 + OUT_SLOTS + IN_SLOTS loads instead of twice as many
    // 4-byte loads with their own address arithmetic (the prefetch was 23 % of the kernel's instructions)
    const int4* nd; const int2* oe; const int2* ie;
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

template <int OUT_SLOTS, int IN_SLOTS> struct BatchEdges { int v; int uo[OUT_SLOTS]; uint32_t wo[OUT_SLOTS]; int ui[IN_SLOTS]; uint32_t wi[IN_SLOTS]; };
// TRUNC (per-node edge truncation, `knn`): ocut[v] = usable leading positions of v's out segment in this resample;
// an in edge is usable iff its position in its source's out segment (ipos) is below the source's ocut
template <int IN_SLOTS, typename VT, bool TRUNC = false>
__device__ __forceinline__ NodeOffs<IN_SLOTS> load_batch_offs(const FastGraph& g, const VT* vlist, int t, int n_s, const uint16_t* ocut = nullptr)
{
    NodeOffs<IN_SLOTS> o;
    o.v = -1; o.o0 = o.o1 = o.i0 = o.i1 = 0;
    if (t < n_s) {
        o.v = (int)vlist[t];
        const int4 d = __ldg(g.nd + o.v);
        o.o0 = d.x; o.o1 = TRUNC ? d.x + (int)ocut[o.v] : d.y; o.i0 = d.z; o.i1 = d.w;
    }
    return o;
}
template <int FT, int OUT_SLOTS, int IN_SLOTS, bool TRUNC = false>
__device__ __forceinline__ BatchEdges<OUT_SLOTS, IN_SLOTS> load_batch_edges(const FastGraph& g, const NodeOffs<IN_SLOTS>& o, int tid,
                                                                            const uint16_t* ipos = nullptr, const uint16_t* ocut = nullptr)
{
    BatchEdges<OUT_SLOTS, IN_SLOTS> e;
    e.v = o.v;
#pragma unroll
    for (int s = 0; s < OUT_SLOTS; ++s) {
        const int eo = o.o0 + s * FT + tid;
        int2 p = make_int2(-1, 0);
        if (eo < o.o1) p = __ldg(g.oe + eo);          // o.v < 0 has o0 == o1 == 0
        e.uo[s] = p.x; e.wo[s] = (uint32_t)p.y;
    }
#pragma unroll
    for (int s = 0; s < IN_SLOTS; ++s) {
        const int ei = o.i0 + s * FT + tid;
        int2 p = make_int2(-1, 0);
        if (ei < o.i1) {
            p = __ldg(g.ie + ei);
            if (TRUNC && ipos[ei] >= ocut[p.x]) p.x = -1;
        }
        e.ui[s] = p.x; e.wi[s] = (uint32_t)p.y;
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
// Speculative batch variant.  The sweep is a sequential Gauss-Seidel pass, but after the first sweeps a
// visit rarely changes anything: the decision for node t+1 equals the one computed from the state before
// node t unless node t moved.  So B consecutive nodes of the visit order are evaluated concurrently, one
// per 128-thread sub-block, all against the state at the start of the batch; the first node of the batch
// whose decision is a move is applied and the batch restarts right after it (later results are dropped).
// Nodes before the first mover saw exactly the state the sequential pass shows them, so the assignment is
// bit-identical to graph_cover_fast_kernel / orc_graph_cover; only the wall-clock order of evaluation
// differs.  Warp arg-max uses redux on the score's bit pattern (scores are positive doubles, so their
// bits order like unsigned integers) instead of a five-level shuffle tree.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// VT: element type of the shared-memory visit order (uint16_t when N < 65536: at the C5 shape this is what lets two
// CTAs of the wide-slot instantiation share an SM)
template <int B, int OUT_SLOTS, int IN_SLOTS, typename VT, bool TRUNC>
__global__ void __launch_bounds__(128 * B, (OUT_SLOTS == 1 && IN_SLOTS <= 3) ? 8 / B : (sizeof(VT) == 2 ? 4 / B : 1))
graph_cover_batch_kernel(FastGraph g, int32_t n_local, const int32_t* __restrict__ resamp_ids, int32_t n_s,
                         const int32_t* __restrict__ samples, const uint64_t* __restrict__ seeds, int32_t min_size,
                         double cooling, int32_t burn_in, int32_t max_sweeps, int32_t ncl_cap, int32_t out_warps,
                         int32_t* __restrict__ mc_all, int32_t* __restrict__ f_all, int32_t* __restrict__ sweeps_out,
                         int32_t* __restrict__ status, int32_t knn, const uint16_t* __restrict__ ipos, uint16_t* __restrict__ ocut_all)
{
    constexpr int FT = 128;            // threads per sub-block (one visited node each)
    constexpr int NT = FT * B;
    constexpr int NW = NT / 32;
    static_assert(B <= 32, "one lane per node of a batch");
    extern __shared__ uint32_t dsm[];
    int32_t* size = (int32_t*)dsm;                               // [ncl_cap]
    uint32_t* wout = dsm + ncl_cap;                              // [B][2][ncl_cap]
    uint32_t* win = wout + (size_t)2 * B * ncl_cap;              // [B][2][ncl_cap]
    VT* vlist = (VT*)(win + (size_t)2 * B * ncl_cap);            // [n_s]
    int16_t* mc_s = (int16_t*)(vlist + ((n_s + 1) & ~1));        // [N]
    __shared__ unsigned long long s_part[NT];
    __shared__ unsigned long long s_pbits[2][B];       // per-node arg-max (score bits, cluster), double buffered by batch parity
    __shared__ int s_pk[2][B];
    // distinct clusters reached by the visited node's out edges (its ~100 edges point into a handful of clusters): the
    // first edge that adds to a cluster's accumulator lists it, and ONE warp scores the list -- the fp64 division of
    // the score used to be issued by all four warps of the node, a dozen times per cluster
    constexpr int CL_MAX = 128 * OUT_SLOTS;
    __shared__ int16_t s_clist[2][B][CL_MAX];
    __shared__ int s_ccount[2][B];
    __shared__ int s_pick, s_flag;
    __shared__ unsigned long long s_total;

    const int N = g.N;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int sb = tid >> 7, ltid = tid & (FT - 1), lwid = ltid >> 5;
    if (tid < 2 * B) s_ccount[tid / B][tid % B] = 0;
    for (int lr = blockIdx.x; lr < n_local; lr += gridDim.x) {
        const int rid = resamp_ids[lr];
        const int32_t* S = samples + (int64_t)rid * n_s;
        const uint64_t seed = seeds[rid];
        const uint32_t seed_lo = (uint32_t)seed, seed_hi = (uint32_t)(seed >> 32);
        int32_t* f = f_all + (int64_t)lr * N;
        uint16_t* ocut = TRUNC ? ocut_all + (int64_t)lr * N : nullptr;

        for (int v = tid; v < N; v += NT) { mc_s[v] = -2; f[v] = 0; }
        for (int k = tid; k < ncl_cap; k += NT) size[k] = 0;
        for (int k = tid; k < 2 * B * ncl_cap; k += NT) { wout[k] = 0; win[k] = 0; }
        if (tid < B) { s_pbits[0][tid] = s_pbits[1][tid] = 0ull; s_pk[0][tid] = s_pk[1][tid] = 0x7fffffff; }
        __syncthreads();
        for (int q = tid; q < n_s; q += NT) mc_s[S[q]] = -1;
        __syncthreads();
        for (int q = tid; q < n_s; q += NT) {
            const int v = S[q];
            int c = 0;
            int e = g.optr[v];
            for (; e < g.optr[v + 1]; ++e) {
                if (mc_s[g.odst[e]] == -1) { if (TRUNC && c == knn) break; ++c; }
            }
            f[v] = c;
            if (TRUNC) ocut[v] = (uint16_t)(e - g.optr[v]);
        }
        __syncthreads();

        // ---- seeding (as graph_cover_kernel, NT threads) -------------------------------------------
        int ncl = 0;
        int err = 0;
        const int chunk = (n_s + NT - 1) / NT;
        const int q0 = min(n_s, tid * chunk), q1 = min(n_s, q0 + chunk);
        for (uint32_t iter = 0;; ++iter) {
            unsigned long long part = 0;
            for (int q = q0; q < q1; ++q) {
                const int v = S[q];
                const int fv = f[v];
                if (mc_s[v] == -1 && fv >= min_size && fv > 0) part += (unsigned long long)fv * fv * fv;
            }
            s_part[tid] = part;
            __syncthreads();
            if (tid < 32) {
                unsigned long long tot = 0;
                for (int q = lane; q < NT; q += 32) tot += s_part[q];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
                if (lane == 0) { s_total = tot; s_pick = -1; }
            }
            __syncthreads();
            const unsigned long long total = s_total;
            if (total == 0) break;
            if (ncl >= ncl_cap) { err = 3; break; }
            const unsigned long long r = cover_rng_u64(seed_lo, seed_hi, STREAM_SEED, iter, 0) % total;
            // exclusive prefix of s_part at tid: warp-level partial sums first, then the tail inside the warp
            unsigned long long excl = 0;
            for (int q = 0; q < (tid & ~31); ++q) excl += s_part[q];
            for (int q = (tid & ~31); q < tid; ++q) excl += s_part[q];
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
            const int oe0 = g.optr[pick], oe1 = TRUNC ? oe0 + (int)ocut[pick] : g.optr[pick + 1];
            for (int e = oe0 + tid; e < oe1; e += NT) {
                const int u = g.odst[e];
                if (mc_s[u] == -1) { mc_s[u] = (int16_t)k; atomicAdd(&s_flag, 1); }
            }
            __syncthreads();
            if (tid == 0) size[k] = s_flag;
            for (int e = oe0 - 1 + wid; e < oe1; e += NW) {
                const int u = (e < oe0) ? pick : g.odst[e];
                if (mc_s[u] != k || (e >= oe0 && u == pick)) continue;
                for (int e2 = g.iptr[u] + lane; e2 < g.iptr[u + 1]; e2 += 32) {
                    const int sn = g.isrc[e2];
                    if (TRUNC && ipos[e2] >= ocut[sn]) continue;
                    atomicSub(&f[sn], 1);
                }
            }
            __syncthreads();
        }
        if (err) {
            if (tid == 0) { status[lr] = err; sweeps_out[lr] = 0; }
            for (int v = tid; v < N; v += NT) mc_all[(int64_t)lr * N + v] = mc_s[v];
            __syncthreads();
            continue;
        }

        // ---- reassignment sweeps, B speculative visits per batch ------------------------------------
        uint32_t hb = 1;
        while ((1ull << (2 * hb)) < (unsigned long long)n_s) ++hb;
        double factor = 1.0;
        int sweep = 0;
        uint32_t parity = 0;
        for (; sweep < max_sweeps; ++sweep) {
            if (sweep > burn_in) factor = factor * cooling;
            uint32_t rk[4];
            {
                const uint64_t w0 = cover_rng_u64(seed_lo, seed_hi, STREAM_PERM, (uint32_t)sweep, 0u);
                const uint64_t w1 = cover_rng_u64(seed_lo, seed_hi, STREAM_PERM, (uint32_t)sweep, 1u);
                rk[0] = (uint32_t)w0; rk[1] = (uint32_t)(w0 >> 32); rk[2] = (uint32_t)w1; rk[3] = (uint32_t)(w1 >> 32);
            }
            __syncthreads();
            for (int t = tid; t < n_s; t += NT) vlist[t] = (VT)S[perm_index((uint32_t)t, (uint32_t)n_s, hb, rk)];
            __syncthreads();
            int moved = 0;
            int t = 0;
            int cur_t = -1, nxt_t = -1;
            BatchEdges<OUT_SLOTS, IN_SLOTS> cur;
            NodeOffs<IN_SLOTS> nxt;
            cur.v = -1; nxt.v = -1;
            uint32_t* const wo_b0 = wout + (size_t)(2 * sb) * ncl_cap; uint32_t* const wo_b1 = wo_b0 + ncl_cap;
            uint32_t* const wi_b0 = win + (size_t)(2 * sb) * ncl_cap;  uint32_t* const wi_b1 = wi_b0 + ncl_cap;
            while (t < n_s) {
                const int my_t = t + sb;
                if (cur_t != my_t) {              // start of sweep, or the previous batch ended on a move: reload
                    cur = load_batch_edges<FT, OUT_SLOTS, IN_SLOTS, TRUNC>(g, load_batch_offs<IN_SLOTS, VT, TRUNC>(g, vlist, my_t, n_s, ocut), ltid, ipos, ocut);
                    cur_t = my_t;
                    nxt_t = my_t + B;
                    nxt = load_batch_offs<IN_SLOTS, VT, TRUNC>(g, vlist, nxt_t, n_s, ocut);
                }
                const BatchEdges<OUT_SLOTS, IN_SLOTS> pre = load_batch_edges<FT, OUT_SLOTS, IN_SLOTS, TRUNC>(g, nxt, ltid, ipos, ocut);   // node my_t + B, in flight
                const NodeOffs<IN_SLOTS> nxt2 = load_batch_offs<IN_SLOTS, VT, TRUNC>(g, vlist, nxt_t + B, n_s, ocut);
                uint32_t* wo_b = (parity & 1) ? wo_b1 : wo_b0;
                uint32_t* wi_b = (parity & 1) ? wi_b1 : wi_b0;
                const int v = cur.v;               // -1 beyond the end of the sweep
                const int nb = min(B, n_s - t);
                // current cluster of the batch's nodes, read before anything in this batch can move
                int rcur = -3;
                if (lane < nb) rcur = (int)mc_s[vlist[t + lane]];
                // [A] association of this sub-block's node towards its neighbouring clusters
                int ko[OUT_SLOTS], ki[IN_SLOTS];
#pragma unroll
                for (int s = 0; s < OUT_SLOTS; ++s) {
                    ko[s] = -1;
                    if (cur.uo[s] >= 0) {
                        ko[s] = (int)mc_s[cur.uo[s]];
                        // weights are at least one unit: a zero accumulator means this edge is the cluster's first
                        if (ko[s] >= 0 && atomicAdd(&wo_b[ko[s]], cur.wo[s]) == 0u)
                            s_clist[parity & 1][sb][atomicAdd(&s_ccount[parity & 1][sb], 1)] = (int16_t)ko[s];
                    }
                }
#pragma unroll
                for (int s = 0; s < IN_SLOTS; ++s) {
                    ki[s] = -1;
                    if (cur.ui[s] >= 0) { ki[s] = (int)mc_s[cur.ui[s]]; if (ki[s] >= 0) atomicAdd(&wi_b[ki[s]], cur.wi[s]); }
                }
                named_bar_sync(1 + sb, FT);
                // [B] best cluster among those reached by an out edge: the node's first warp scores the distinct clusters
                if (lwid == 0) {
                    unsigned long long bits = 0ull; int bko = 0x7fffffff;
                    const int ncand = s_ccount[parity & 1][sb];
                    const int curk = v >= 0 ? (int)mc_s[v] : -1;
                    for (int c = lane; c < ncand; c += 32) {
                        const int k = (int)s_clist[parity & 1][sb][c];
                        const uint32_t wo = wo_b[k], wi = wi_b[k];
                        if (wi != 0) {
                            const double sz = (double)size[k];
                            double sc = (double)((unsigned long long)wo * (unsigned long long)wi) / (sz * sz);
                            if (k == curk) sc = sc * factor;
                            const unsigned long long b2 = (unsigned long long)__double_as_longlong(sc);
                            if (b2 > bits || (b2 == bits && k < bko)) { bits = b2; bko = k; }
                        }
                    }
                    const uint32_t hi = (uint32_t)(bits >> 32), lo = (uint32_t)bits;
                    const uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
                    const uint32_t ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
                    const bool top = bits != 0ull && hi == mh && lo == ml;
                    const uint32_t mk = __reduce_min_sync(0xffffffffu, top ? (uint32_t)bko : 0x7fffffffu);
                    if (lane == 0) {
                        s_pbits[parity & 1][sb] = ((unsigned long long)mh << 32) | ml; s_pk[parity & 1][sb] = (int)mk;
                        s_ccount[parity & 1][sb] = 0;              // next used two batches from now
                    }
                }
                __syncthreads();
                // [C] clear this batch's accumulators (next used two batches from now)
#pragma unroll
                for (int s = 0; s < OUT_SLOTS; ++s) if (ko[s] >= 0) wo_b[ko[s]] = 0;
#pragma unroll
                for (int s = 0; s < IN_SLOTS; ++s) if (ki[s] >= 0) wi_b[ki[s]] = 0;
                // [D] resolve: every warp finds the first node of the batch that moves (lane j <-> node j)
                unsigned long long rb = 0ull; int rk_ = 0x7fffffff;
                if (lane < nb) { rb = s_pbits[parity & 1][lane]; rk_ = s_pk[parity & 1][lane]; }
                const bool mv = lane < nb && rb != 0ull && rk_ != rcur;
                const unsigned mvmask = __ballot_sync(0xffffffffu, mv);
                parity ^= 1u;
                cur = pre; cur_t = nxt_t; nxt = nxt2; nxt_t += B;
                if (mvmask == 0u) {
                    t += nb;
                } else {
                    const int js = __ffs(mvmask) - 1;          // the first moving node
                    const int bk = __shfl_sync(0xffffffffu, rk_, js);
                    const int curk = __shfl_sync(0xffffffffu, rcur, js);
                    if (tid == 0) {
                        if (curk >= 0) size[curk]--;
                        size[bk]++;
                        mc_s[vlist[t + js]] = (int16_t)bk;
                    }
                    ++moved;
                    t += js + 1;
                    __syncthreads();
                }
            }
            if (moved == 0) { ++sweep; break; }
        }
        __syncthreads();
        for (int v = tid; v < N; v += NT) mc_all[(int64_t)lr * N + v] = mc_s[v];
        if (tid == 0) { sweeps_out[lr] = sweep; status[lr] = 0; }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Warp-per-node speculative variant.  graph_cover_batch_kernel is issue-bound: four warps per visited
// node execute ~1,300 warp instructions per visit although a node has only ~150-200 edges.  Here each
// WARP evaluates one node of the batch (B = warps per CTA nodes per batch): edges are walked 32 at a time
// from L1 (the CTA touches the edge lines of the following batch while it works on the current one),
// the accumulators are warp-private so no barrier separates accumulation from scoring, and one block
// barrier per batch resolves the first mover exactly as in graph_cover_batch_kernel.  Same arithmetic,
// bit-identical assignment.  Requires out-degree <= 32 * OUT_SLOTS (host-checked).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void touch_line(const void* p)
{
    uint32_t sink;
    asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(sink) : "l"(p));
    (void)sink;
}

template <int B, int OUT_SLOTS>
__global__ void __launch_bounds__(32 * B, 1024 / (32 * B))
graph_cover_warp_kernel(FastGraph g, int32_t n_local, const int32_t* __restrict__ resamp_ids, int32_t n_s,
                        const int32_t* __restrict__ samples, const uint64_t* __restrict__ seeds, int32_t min_size,
                        double cooling, int32_t burn_in, int32_t max_sweeps, int32_t ncl_cap, int32_t kin_stride,
                        int32_t* __restrict__ mc_all, int32_t* __restrict__ f_all, int32_t* __restrict__ sweeps_out,
                        int32_t* __restrict__ status)
{
    constexpr int NT = 32 * B;
    static_assert(B <= 32, "one lane per node partial");
    extern __shared__ uint32_t dsm[];
    int32_t* size = (int32_t*)dsm;                               // [ncl_cap]
    uint32_t* accO = dsm + ncl_cap;                              // [B][ncl_cap]
    uint32_t* accI = accO + (size_t)B * ncl_cap;                 // [B][ncl_cap]
    int32_t* vlist = (int32_t*)(accI + (size_t)B * ncl_cap);     // [n_s]
    int16_t* mc_s = (int16_t*)(vlist + n_s);                     // [N]
    int16_t* kin = mc_s + ((g.N + 1) & ~1);                      // [B][kin_stride] clusters of the node's in-neighbours
    __shared__ unsigned long long s_part[NT];
    __shared__ unsigned long long s_pbits[2][B];
    __shared__ int s_pk[2][B];
    __shared__ int s_pick, s_flag;
    __shared__ unsigned long long s_total;
    __shared__ int s_kcnt[B];

    const int N = g.N;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int* mycnt = &s_kcnt[wid];
    uint32_t* myO = accO + (size_t)wid * ncl_cap;
    uint32_t* myI = accI + (size_t)wid * ncl_cap;
    int16_t* mykin = kin + (size_t)wid * kin_stride;
    for (int lr = blockIdx.x; lr < n_local; lr += gridDim.x) {
        const int rid = resamp_ids[lr];
        const int32_t* S = samples + (int64_t)rid * n_s;
        const uint64_t seed = seeds[rid];
        const uint32_t seed_lo = (uint32_t)seed, seed_hi = (uint32_t)(seed >> 32);
        int32_t* f = f_all + (int64_t)lr * N;

        for (int v = tid; v < N; v += NT) { mc_s[v] = -2; f[v] = 0; }
        for (int k = tid; k < ncl_cap; k += NT) size[k] = 0;
        for (int k = tid; k < B * ncl_cap; k += NT) { accO[k] = 0; accI[k] = 0; }
        if (tid < B) { s_pbits[0][tid] = s_pbits[1][tid] = 0ull; s_pk[0][tid] = s_pk[1][tid] = 0x7fffffff; s_kcnt[tid] = 0; }
        __syncthreads();
        for (int q = tid; q < n_s; q += NT) mc_s[S[q]] = -1;
        __syncthreads();
        for (int q = tid; q < n_s; q += NT) {
            const int v = S[q];
            int c = 0;
            for (int e = g.optr[v]; e < g.optr[v + 1]; ++e) c += (mc_s[g.odst[e]] == -1);
            f[v] = c;
        }
        __syncthreads();

        // ---- seeding (as graph_cover_kernel, NT threads) -------------------------------------------
        int ncl = 0;
        int err = 0;
        const int chunk = (n_s + NT - 1) / NT;
        const int q0 = min(n_s, tid * chunk), q1 = min(n_s, q0 + chunk);
        for (uint32_t iter = 0;; ++iter) {
            unsigned long long part = 0;
            for (int q = q0; q < q1; ++q) {
                const int v = S[q];
                const int fv = f[v];
                if (mc_s[v] == -1 && fv >= min_size && fv > 0) part += (unsigned long long)fv * fv * fv;
            }
            s_part[tid] = part;
            __syncthreads();
            if (tid < 32) {
                unsigned long long tot = 0;
                for (int q = lane; q < NT; q += 32) tot += s_part[q];
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
            for (int e = oe0 + tid; e < oe1; e += NT) {
                const int u = g.odst[e];
                if (mc_s[u] == -1) { mc_s[u] = (int16_t)k; atomicAdd(&s_flag, 1); }
            }
            __syncthreads();
            if (tid == 0) size[k] = s_flag;
            for (int e = oe0 - 1 + wid; e < oe1; e += B) {
                const int u = (e < oe0) ? pick : g.odst[e];
                if (mc_s[u] != k || (e >= oe0 && u == pick)) continue;
                for (int e2 = g.iptr[u] + lane; e2 < g.iptr[u + 1]; e2 += 32) atomicSub(&f[g.isrc[e2]], 1);
            }
            __syncthreads();
        }
        if (err) {
            if (tid == 0) { status[lr] = err; sweeps_out[lr] = 0; }
            for (int v = tid; v < N; v += NT) mc_all[(int64_t)lr * N + v] = mc_s[v];
            __syncthreads();
            continue;
        }

        // ---- reassignment sweeps, one node per warp, B speculative visits per batch -----------------
        uint32_t hb = 1;
        while ((1ull << (2 * hb)) < (unsigned long long)n_s) ++hb;
        double factor = 1.0;
        int sweep = 0;
        uint32_t parity = 0;
        for (; sweep < max_sweeps; ++sweep) {
            if (sweep > burn_in) factor = factor * cooling;
            uint32_t rk[4];
            {
                const uint64_t w0 = cover_rng_u64(seed_lo, seed_hi, STREAM_PERM, (uint32_t)sweep, 0u);
                const uint64_t w1 = cover_rng_u64(seed_lo, seed_hi, STREAM_PERM, (uint32_t)sweep, 1u);
                rk[0] = (uint32_t)w0; rk[1] = (uint32_t)(w0 >> 32); rk[2] = (uint32_t)w1; rk[3] = (uint32_t)(w1 >> 32);
            }
            __syncthreads();
            for (int t = tid; t < n_s; t += NT) vlist[t] = S[perm_index((uint32_t)t, (uint32_t)n_s, hb, rk)];
            __syncthreads();
            int moved = 0;
            int t = 0;
            while (t < n_s) {
                const int nb = min(B, n_s - t);
                int rcur = -3;                     // current cluster of the batch's nodes, before anything can move
                if (lane < nb) rcur = (int)mc_s[vlist[t + lane]];
                const int my_t = t + wid;
                unsigned long long bbits = 0ull; int bk = 0x7fffffff;
                if (my_t < n_s) {
                    const int v = vlist[my_t];
                    const int o0 = __ldg(g.optr + v), o1 = __ldg(g.optr + v + 1);
                    const int i0 = __ldg(g.iptr + v), i1 = __ldg(g.iptr + v + 1);
                    // [A] association towards neighbouring clusters.  The lane whose add finds the accumulator at
                    // zero owns that cluster for this visit: it scores it in [B] and clears it in [C].
                    int ko[OUT_SLOTS];
                    unsigned own = 0u;
#pragma unroll
                    for (int s = 0; s < OUT_SLOTS; ++s) {
                        ko[s] = -1;
                        const int e = o0 + s * 32 + lane;
                        if (e < o1) {
                            const int u = __ldg(g.odst + e);
                            const uint32_t w = __ldg(g.ow + e);
                            ko[s] = (int)mc_s[u];
                            if (ko[s] >= 0 && atomicAdd(&myO[ko[s]], w) == 0u) own |= 1u << s;
                        }
                    }
                    for (int e = i0 + lane; e < i1; e += 32) {
                        const int u = __ldg(g.isrc + e);
                        const uint32_t w = __ldg(g.iw + e);
                        const int k = (int)mc_s[u];
                        if (k >= 0 && atomicAdd(&myI[k], w) == 0u) mykin[atomicAdd(mycnt, 1)] = (int16_t)k;
                    }
                    // touch the edge lines of the node this warp expects in the next batch
                    if (my_t + B < n_s) {
                        const int vn = vlist[my_t + B];
                        const int p0 = __ldg(g.optr + vn), p1 = __ldg(g.optr + vn + 1);
                        const int r0 = __ldg(g.iptr + vn), r1 = __ldg(g.iptr + vn + 1);
                        const int eo = p0 + 32 * lane, ei = r0 + 32 * lane;
                        if (eo < p1 + 31 && p1 > p0) { const int a = min(eo, p1 - 1); touch_line(g.odst + a); touch_line(g.ow + a); }
                        if (ei < r1 + 31 && r1 > r0) { const int a = min(ei, r1 - 1); touch_line(g.isrc + a); touch_line(g.iw + a); }
                    }
                    __syncwarp();
                    // [B] best cluster among those reached by an out edge
                    const int curk = (int)mc_s[v];
#pragma unroll
                    for (int s = 0; s < OUT_SLOTS; ++s) {
                        const bool mine = (own >> s) & 1u;
                        if (__any_sync(0xffffffffu, mine)) {
                            if (mine) {
                                const uint32_t wo = myO[ko[s]], wi = myI[ko[s]];
                                if (wi != 0) {
                                    const double sz = (double)size[ko[s]];
                                    double sc = (double)((unsigned long long)wo * (unsigned long long)wi) / (sz * sz);
                                    if (ko[s] == curk) sc = sc * factor;
                                    const unsigned long long bits = (unsigned long long)__double_as_longlong(sc);
                                    if (bits > bbits || (bits == bbits && ko[s] < bk)) { bbits = bits; bk = ko[s]; }
                                }
                            }
                        }
                    }
                    const uint32_t hi = (uint32_t)(bbits >> 32), lo = (uint32_t)bbits;
                    const uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
                    const uint32_t ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
                    const bool top = bbits != 0ull && hi == mh && lo == ml;
                    const uint32_t mk = __reduce_min_sync(0xffffffffu, top ? (uint32_t)bk : 0x7fffffffu);
                    if (lane == 0) { s_pbits[parity & 1][wid] = ((unsigned long long)mh << 32) | ml; s_pk[parity & 1][wid] = (int)mk; }
                    __syncwarp();
                    // [C] clear the warp's accumulators
#pragma unroll
                    for (int s = 0; s < OUT_SLOTS; ++s) if ((own >> s) & 1u) myO[ko[s]] = 0;
                    const int nin = *mycnt;
                    for (int q = lane; q < nin; q += 32) myI[mykin[q]] = 0;
                    __syncwarp();
                    if (lane == 0) *mycnt = 0;
                }
                __syncthreads();
                // [D] resolve: every warp finds the first node of the batch that moves
                unsigned long long rb = 0ull; int rk_ = 0x7fffffff;
                if (lane < nb) { rb = s_pbits[parity & 1][lane]; rk_ = s_pk[parity & 1][lane]; }
                const bool mv = lane < nb && rb != 0ull && rk_ != rcur;
                const unsigned mvmask = __ballot_sync(0xffffffffu, mv);
                parity ^= 1u;
                if (mvmask == 0u) {
                    t += nb;
                } else {
                    const int js = __ffs(mvmask) - 1;
                    const int nk = __shfl_sync(0xffffffffu, rk_, js);
                    const int curk = __shfl_sync(0xffffffffu, rcur, js);
                    if (tid == 0) {
                        if (curk >= 0) size[curk]--;
                        size[nk]++;
                        mc_s[vlist[t + js]] = (int16_t)nk;
                    }
                    ++moved;
                    t += js + 1;
                    __syncthreads();
                }
            }
            if (moved == 0) { ++sweep; break; }
        }
        __syncthreads();
        for (int v = tid; v < N; v += NT) mc_all[(int64_t)lr * N + v] = mc_s[v];
        if (tid == 0) { sweeps_out[lr] = sweep; status[lr] = 0; }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Incremental variant.  The kernels above re-read all ~400 edges of a node at every visit, although in a converged cover
// almost no visit changes anything (C3: 90,000 visits and 8,700 moves per resample).  Here every node carries two dense
// rows in global memory, wout[v][k] = summed weight of v's usable out edges into cluster k and win[v][k] = summed weight
// of its in edges from cluster k (row stride = the number of clusters the seeding made, rounded up to 32).  A visit is ONE
// WARP reading the node's two rows (coalesced, address known from the visit order alone) and scoring the clusters present
// in both; a batch is up to B visits evaluated speculatively side by side as in graph_cover_batch_kernel.  A move of
// node v from cluster a to b edits the rows of v's neighbours, one thread per edge of v and two reductions each at
// addresses that need no lookup: an out edge v -> u changes win[u][a], win[u][b]; an in edge u -> v changes wout[u][a],
// wout[u][b].  All sums are integers: the scores, and with them every decision, are the ones orc_graph_cover computes
// from the edges (tests/cover_incr_model.c is the host model of this structure, checked against the oracle on the CPU;
// the kernel is checked bit-exact against the oracle on the GPU).  The number of speculative visits adapts to the move
// rate (2 while most visits move, B when hardly any does).  Resamples are handed to CTAs through an atomic counter, so
// the row workspace is per resident CTA, not per resample.
// ---------------------------------------------------------------------------------------------
struct IncrWs {
    uint32_t* rows;                   // [slots][2][N * stride_max]: wout rows, then win rows
    int32_t* next;                    // work counter, zeroed before the launch
    int64_t slot_words;               // N * stride_max
};

template <int B, typename VT, bool TRUNC>
__global__ void __launch_bounds__(32 * B, B >= 16 ? 2 : 4)
graph_cover_incr_kernel(FastGraph g, int32_t n_local, const int32_t* __restrict__ resamp_ids, int32_t n_s,
                        const int32_t* __restrict__ samples, const uint64_t* __restrict__ seeds, int32_t min_size,
                        double cooling, int32_t burn_in, int32_t max_sweeps, int32_t ncl_cap,
                        int32_t* __restrict__ mc_all, int32_t* __restrict__ f_all, int32_t* __restrict__ sweeps_out,
                        int32_t* __restrict__ status, int32_t knn, const uint16_t* __restrict__ ipos, uint16_t* __restrict__ ocut_all,
                        IncrWs ws)
{
    constexpr int NT = 32 * B;
    static_assert(B <= 32, "one lane per node of a batch");
    extern __shared__ uint32_t dsm[];
    int32_t* size = (int32_t*)dsm;                               // [ncl_cap]
    VT* vlist = (VT*)(dsm + ncl_cap);                            // [n_s]
    int16_t* mc_s = (int16_t*)(vlist + ((n_s + 1) & ~1));        // [N]
    __shared__ unsigned long long s_part[NT];
    __shared__ int s_pk[2][B];                // per-node decision (best cluster, or none), double buffered by batch parity
    __shared__ int4 s_nd[2][B];               // the node's CSR descriptor, fetched beside the rows for the update that may follow
    __shared__ int4 s_cand[B][32];            // per warp: {cluster, wout, win} of the clusters present in both rows of its node
    __shared__ int s_pick, s_flag, s_lr;
    __shared__ unsigned long long s_total;
    constexpr int NONE = 0x7fffffff;

    const int N = g.N;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t* const wout = ws.rows + (size_t)blockIdx.x * 2 * ws.slot_words;
    uint32_t* const win = wout + ws.slot_words;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_lr = atomicAdd(ws.next, 1);
        __syncthreads();
        const int lr = s_lr;
        if (lr >= n_local) break;
#ifdef MCB_COVER_PROF
        long long pc_t0 = clock64(), pc_seed = 0, pc_build = 0, pc_eval = 0, pc_upd = 0, pc_perm = 0, pc_tmp;
        int pc_batches = 0, pc_moves = 0, pc_visits = 0, pc_exact = 0;
#endif
        const int rid = resamp_ids[lr];
        const int32_t* S = samples + (int64_t)rid * n_s;
        const uint64_t seed = seeds[rid];
        const uint32_t seed_lo = (uint32_t)seed, seed_hi = (uint32_t)(seed >> 32);
        int32_t* f = f_all + (int64_t)lr * N;
        uint16_t* ocut = TRUNC ? ocut_all + (int64_t)lr * N : nullptr;

        for (int v = tid; v < N; v += NT) { mc_s[v] = -2; f[v] = 0; }
        for (int k = tid; k < ncl_cap; k += NT) size[k] = 0;
        if (tid < B) s_pk[0][tid] = s_pk[1][tid] = NONE;
        __syncthreads();
        for (int q = tid; q < n_s; q += NT) mc_s[S[q]] = -1;
        __syncthreads();
        for (int q = tid; q < n_s; q += NT) {
            const int v = S[q];
            int c = 0;
            int e = g.optr[v];
            for (; e < g.optr[v + 1]; ++e) {
                if (mc_s[g.odst[e]] == -1) { if (TRUNC && c == knn) break; ++c; }
            }
            f[v] = c;
            if (TRUNC) ocut[v] = (uint16_t)(e - g.optr[v]);
        }
        __syncthreads();

        // ---- seeding (as graph_cover_batch_kernel) ---------------------------------------------------
        int ncl = 0;
        int err = 0;
        const int chunk = (n_s + NT - 1) / NT;
        const int q0 = min(n_s, tid * chunk), q1 = min(n_s, q0 + chunk);
        for (uint32_t iter = 0;; ++iter) {
            unsigned long long part = 0;
            for (int q = q0; q < q1; ++q) {
                const int v = S[q];
                const int fv = f[v];
                if (mc_s[v] == -1 && fv >= min_size && fv > 0) part += (unsigned long long)fv * fv * fv;
            }
            s_part[tid] = part;
            __syncthreads();
            if (tid < 32) {
                unsigned long long tot = 0;
                for (int q = lane; q < NT; q += 32) tot += s_part[q];
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
            const int oe0 = g.optr[pick], oe1 = TRUNC ? oe0 + (int)ocut[pick] : g.optr[pick + 1];
            for (int e = oe0 + tid; e < oe1; e += NT) {
                const int u = g.odst[e];
                if (mc_s[u] == -1) { mc_s[u] = (int16_t)k; atomicAdd(&s_flag, 1); }
            }
            __syncthreads();
            if (tid == 0) size[k] = s_flag;
            for (int e = oe0 - 1 + wid; e < oe1; e += B) {
                const int u = (e < oe0) ? pick : g.odst[e];
                if (mc_s[u] != k || (e >= oe0 && u == pick)) continue;
                for (int e2 = g.iptr[u] + lane; e2 < g.iptr[u + 1]; e2 += 32) {
                    const int sn = g.isrc[e2];
                    if (TRUNC && ipos[e2] >= ocut[sn]) continue;
                    atomicSub(&f[sn], 1);
                }
            }
            __syncthreads();
        }
        if (err) {
            if (tid == 0) { status[lr] = err; sweeps_out[lr] = 0; }
            for (int v = tid; v < N; v += NT) mc_all[(int64_t)lr * N + v] = mc_s[v];
            continue;
        }
#ifdef MCB_COVER_PROF
        pc_seed = clock64() - pc_t0; pc_tmp = clock64();
#endif
        // ---- rows of every node from the seeded state: zero, then one warp per sampled node walks its usable out edges
        // (the edge v -> u with both ends sampled feeds wout[v][cluster of u] and win[u][cluster of v])
        const int stride = max(32, (ncl + 31) & ~31);
        {
            uint4* const z = reinterpret_cast<uint4*>(wout);            // slot_words is a multiple of 32: both halves stay 16-byte aligned
            const size_t nz = (size_t)N * stride / 4;
            for (size_t i = tid; i < nz; i += NT) z[i] = make_uint4(0u, 0u, 0u, 0u);
            uint4* const z2 = reinterpret_cast<uint4*>(win);
            for (size_t i = tid; i < nz; i += NT) z2[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();
        for (int q = wid; q < n_s; q += B) {
            const int v = S[q];
            const int4 d = __ldg(g.nd + v);
            const int nout = TRUNC ? (int)ocut[v] : d.y - d.x;
            const int kv = (int)mc_s[v];
            for (int i = lane; i < nout; i += 32) {
                const int2 p = __ldg(g.oe + d.x + i);
                const int ku = (int)mc_s[p.x];
                if (ku == -2) continue;
                if (ku >= 0) atomicAdd(wout + (size_t)v * stride + ku, (uint32_t)p.y);
                if (kv >= 0) atomicAdd(win + (size_t)p.x * stride + kv, (uint32_t)p.y);
            }
        }
#ifdef MCB_COVER_PROF
        __syncthreads();
        pc_build = clock64() - pc_tmp;
#endif
        // ---- reassignment sweeps: up to B speculative visits per batch, one warp each ---------------------
        uint32_t hb = 1;
        while ((1ull << (2 * hb)) < (unsigned long long)n_s) ++hb;
        double factor = 1.0;
        int sweep = 0;
        uint32_t parity = 0;
        int nbc = B;                          // visits of the next batch
        for (; sweep < max_sweeps; ++sweep) {
            if (sweep > burn_in) factor = factor * cooling;
            const float factor_f = (float)factor;
            uint32_t rk[4];
            {
                const uint64_t w0 = cover_rng_u64(seed_lo, seed_hi, STREAM_PERM, (uint32_t)sweep, 0u);
                const uint64_t w1 = cover_rng_u64(seed_lo, seed_hi, STREAM_PERM, (uint32_t)sweep, 1u);
                rk[0] = (uint32_t)w0; rk[1] = (uint32_t)(w0 >> 32); rk[2] = (uint32_t)w1; rk[3] = (uint32_t)(w1 >> 32);
            }
            __syncthreads();
#ifdef MCB_COVER_PROF
            pc_tmp = clock64();
#endif
            for (int t = tid; t < n_s; t += NT) vlist[t] = (VT)S[perm_index((uint32_t)t, (uint32_t)n_s, hb, rk)];
            __syncthreads();
#ifdef MCB_COVER_PROF
            pc_perm += clock64() - pc_tmp;
#endif
            int moved = 0;
            int t = 0;
            while (t < n_s) {
                const int nb = min(nbc, n_s - t);
                const uint32_t par = parity & 1u;
#ifdef MCB_COVER_PROF
                pc_tmp = clock64(); ++pc_batches; pc_visits += nb;
#endif
                // current cluster of the batch's nodes, read before anything in this batch can move (lane j <-> node j)
                int rcur = -3;
                if (lane < nb) rcur = (int)mc_s[vlist[t + lane]];
                if (wid < nb) {
                    const int v = (int)vlist[t + wid];
                    const int curk = (int)mc_s[v];
                    const uint32_t* const ro = wout + (size_t)v * stride + lane;
                    const uint32_t* const ri = win + (size_t)v * stride + lane;
                    if (lane == 0) s_nd[par][wid] = __ldg(g.nd + v);
                    // Scores in fp32 first: the arg-max is decided by them whenever the runner-up is more than 1e-5 (relative)
                    // behind, far above the 1e-6 the fp32 evaluation can be off by -- the exact fp64 scores (one correctly
                    // rounded division per candidate) then order the same way and are not needed.  Closer calls, ties
                    // included, take the exact path below.
                    // the handful of clusters present in both rows are first packed into the warp's candidate list, one
                    // per lane, so the score arithmetic runs once per visit instead of once per 32-cluster group
                    int ncand = 0;
                    for (int k0 = 0; k0 < stride; k0 += 128) {           // four 32-cluster groups per round, eight loads in flight
                        uint32_t wo[4], wi[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const bool in = k0 + 32 * j < stride;
                            wo[j] = in ? __ldcg(ro + k0 + 32 * j) : 0u;
                            wi[j] = in ? __ldcg(ri + k0 + 32 * j) : 0u;
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const bool c = wo[j] != 0u && wi[j] != 0u;
                            const unsigned m = __ballot_sync(0xffffffffu, c);
                            const int slot = ncand + __popc(m & lt_mask);
                            if (c && slot < 32) s_cand[wid][slot] = make_int4(k0 + 32 * j + lane, (int)wo[j], (int)wi[j], 0);
                            ncand += __popc(m);
                        }
                    }
                    __syncwarp();
                    float sc = 0.f; int kc = NONE;
                    if (lane < ncand) {                                  // (more than 32 candidates: the exact path below)
                        const int4 c = s_cand[wid][lane];
                        const float sz = (float)size[c.x];
                        sc = __fdividef((float)(uint32_t)c.y * (float)(uint32_t)c.z, sz * sz);
                        if (c.x == curk) sc = sc * factor_f;
                        kc = c.x;
                    }
                    const float mx = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(sc)));   // scores are >= 0
                    const float thr = mx * (1.0f - 1e-5f);
                    const unsigned m1 = __ballot_sync(0xffffffffu, sc >= thr);
                    int bk_node = NONE;
                    if (mx > 0.f) {
                        if (ncand <= 32 && (m1 & (m1 - 1u)) == 0u) bk_node = __shfl_sync(0xffffffffu, kc, __ffs(m1) - 1);
                        else {
#ifdef MCB_COVER_PROF
                            ++pc_exact;
#endif
                            unsigned long long bits = 0ull; int bko = NONE;
                            for (int k = lane; k < stride; k += 32) {
                                const uint32_t wo = __ldcg(ro + k - lane), wi = __ldcg(ri + k - lane);
                                if (wo != 0u && wi != 0u) {
                                    const double sz = (double)size[k];
                                    double dsc = (double)((unsigned long long)wo * (unsigned long long)wi) / (sz * sz);
                                    if (k == curk) dsc = dsc * factor;
                                    const unsigned long long b2 = (unsigned long long)__double_as_longlong(dsc);
                                    if (b2 > bits || (b2 == bits && k < bko)) { bits = b2; bko = k; }
                                }
                            }
                            const uint32_t hi = (uint32_t)(bits >> 32), lo = (uint32_t)bits;
                            const uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
                            const uint32_t ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
                            const bool top = bits != 0ull && hi == mh && lo == ml;
                            bk_node = (int)__reduce_min_sync(0xffffffffu, top ? (uint32_t)bko : (uint32_t)NONE);
                        }
                    }
                    if (lane == 0) s_pk[par][wid] = bk_node;
                }
                __syncthreads();
#ifdef MCB_COVER_PROF
                pc_eval += clock64() - pc_tmp; pc_tmp = clock64();
#endif
                // resolve: every warp finds the first node of the batch that moves
                int rk_ = NONE;
                if (lane < nb) rk_ = s_pk[par][lane];
                const bool mv = lane < nb && rk_ != NONE && rk_ != rcur;
                const unsigned mvmask = __ballot_sync(0xffffffffu, mv);
                parity ^= 1u;
                if (mvmask == 0u) {
                    t += nb;
                    nbc = min(B, 2 * nbc);
                    continue;
                }
                const int js = __ffs(mvmask) - 1;              // the first moving node
                const int bk = __shfl_sync(0xffffffffu, rk_, js);
                const int ck = __shfl_sync(0xffffffffu, rcur, js);
                const int v = (int)vlist[t + js];
                if (tid == 0) {
                    if (ck >= 0) size[ck]--;
                    size[bk]++;
                    mc_s[v] = (int16_t)bk;
                }
                {   // v's out edges are in edges of their targets, v's in edges are out edges of their sources
                    const int4 d = s_nd[par][js];
                    const int nout = TRUNC ? (int)ocut[v] : d.y - d.x;
                    const int nin = d.w - d.z;
                    for (int it = tid; it < nout + nin; it += NT) {
                        uint32_t* row;
                        int2 p;
                        if (it < nout) {
                            p = __ldg(g.oe + d.x + it);
                            if (mc_s[p.x] == -2) continue;
                            row = win + (size_t)p.x * stride;
                        } else {
                            const int ei = d.z + it - nout;
                            p = __ldg(g.ie + ei);
                            if (mc_s[p.x] == -2 || (TRUNC && ipos[ei] >= ocut[p.x])) continue;
                            row = wout + (size_t)p.x * stride;
                        }
                        if (ck >= 0) atomicSub(row + ck, (uint32_t)p.y);
                        atomicAdd(row + bk, (uint32_t)p.y);
                    }
                }
                ++moved;
                t += js + 1;
                nbc = max(2, min(B, 2 * (js + 1)));
                __syncthreads();
#ifdef MCB_COVER_PROF
                pc_upd += clock64() - pc_tmp; ++pc_moves;
#endif
            }
            if (moved == 0) { ++sweep; break; }
        }
        __syncthreads();
        for (int v = tid; v < N; v += NT) mc_all[(int64_t)lr * N + v] = mc_s[v];
        if (tid == 0) { sweeps_out[lr] = sweep; status[lr] = 0; }
#ifdef MCB_COVER_PROF
        if (tid == 0 && lr == 0)
            printf("incr<%d> lr 0: total %lld cyc; seed %lld build %lld perm %lld eval %lld upd %lld; batches %d visits %d moves %d sweeps %d stride %d; exact scores in %d of warp 0's visits\n", B,
                   clock64() - pc_t0, pc_seed, pc_build, pc_perm, pc_eval, pc_upd, pc_batches, pc_visits, pc_moves, sweep, stride, pc_exact);
#endif
    }
}

// ---------------------------------------------------------------------------------------------
// CSR (out and in) of the host's edge table, built on the device: degree count, scan, fill.  The order of a node's
// edges inside its IN segment is arbitrary (atomic cursors); every consumer only sums integer weights or counts over a
// segment, so the cover does not depend on it.  OUT segments are ordered by (weight descending, target ascending) when
// per-node edge truncation (`knn`) is requested.  Also validates what the host loops used to validate.
// err bits: 1 = edge endpoint out of range, 2 = sampled node out of range, 4 = sample set not strictly ascending,
// 8 = out segment too long to order, 16 = a (source, target) pair listed twice.
// ---------------------------------------------------------------------------------------------
__global__ void edge_degree_kernel(int64_t ne, const int32_t* __restrict__ src, const int32_t* __restrict__ dst, int32_t N,
                                   int32_t* __restrict__ odeg, int32_t* __restrict__ ideg, int32_t* __restrict__ err)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
        const int32_t a = src[e], b = dst[e];
        if (a < 0 || a >= N || b < 0 || b >= N) { atomicOr(err, 1); continue; }
        atomicAdd(&odeg[a], 1); atomicAdd(&ideg[b], 1);
    }
}
__global__ void edge_fill_out_kernel(int64_t ne, const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                     const double* __restrict__ w, int32_t N, const int32_t* __restrict__ optr,
                                     int32_t* __restrict__ ocur, int32_t* __restrict__ odst, uint32_t* __restrict__ ow)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
        const int32_t a = src[e], b = dst[e];
        if (a < 0 || a >= N || b < 0 || b >= N) continue;
        const double wf = rint(w[e] * 65536.0);                 // 16-bit fixed point, at least one unit (as orc_graph_to_csr)
        const uint32_t wi = wf < 1.0 ? 1u : (uint32_t)wf;
        const int po = optr[a] + atomicAdd(&ocur[a], 1);
        odst[po] = b; ow[po] = wi;
    }
}
// One warp per node: (SORT) order the node's out segment by (weight descending, target ascending) -- the order in which
// tgs_graph_cover_resample's `knn` keeps a node's edges -- then write the packed out edges and scatter the node's edges
// into the in-CSR of their targets, each with its POSITION in the source's out segment (ipos).  err bit 8: a segment
// longer than COVER_SORT_MAX could not be ordered.
constexpr int COVER_SORT_MAX = 1024;
template <bool SORT>
__global__ void __launch_bounds__(128)
finish_csr_kernel(int32_t N, const int32_t* __restrict__ optr, const int32_t* __restrict__ iptr, int32_t* __restrict__ icur,
                  int32_t* __restrict__ odst, uint32_t* __restrict__ ow, int2* __restrict__ oe, int32_t* __restrict__ isrc,
                  uint32_t* __restrict__ iw, int2* __restrict__ ie, uint16_t* __restrict__ ipos, int32_t* __restrict__ err)
{
    __shared__ int32_t sd[SORT ? 4 : 1][SORT ? COVER_SORT_MAX : 1];
    __shared__ uint32_t sw[SORT ? 4 : 1][SORT ? COVER_SORT_MAX : 1];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int v = blockIdx.x * 4 + wid; v < N; v += gridDim.x * 4) {
        const int o0 = optr[v], d = optr[v + 1] - o0;
        if (SORT && d > 1) {
            if (d <= COVER_SORT_MAX) {
                for (int q = lane; q < d; q += 32) { sd[wid][q] = odst[o0 + q]; sw[wid][q] = ow[o0 + q]; }
                __syncwarp();
                for (int q = lane; q < d; q += 32) {
                    const int32_t t = sd[wid][q]; const uint32_t x = sw[wid][q];
                    int rank = 0;
                    for (int r = 0; r < d; ++r) rank += (sw[wid][r] > x || (sw[wid][r] == x && sd[wid][r] < t)) ? 1 : 0;
                    odst[o0 + rank] = t; ow[o0 + rank] = x;          // (src, dst) pairs are unique: ranks are a permutation
                }
                __syncwarp();
            } else if (lane == 0) atomicOr(err, 8);
        }
        for (int q = lane; q < d; q += 32) {
            const int32_t t = odst[o0 + q]; const uint32_t x = ow[o0 + q];
            oe[o0 + q] = make_int2(t, (int)x);
            const int pi = iptr[t] + atomicAdd(&icur[t], 1);
            isrc[pi] = v; iw[pi] = x; ie[pi] = make_int2(v, (int)x); ipos[pi] = (uint16_t)min(q, 65535);
        }
        __syncwarp();
    }
}
// err bit 16: some node has two out edges to the same target.  Every cover kernel claims a seed's neighbours with one
// thread per edge and relies on (source, target) pairs being unique, as they are in any cell graph; a table that repeats
// a pair is refused instead of being covered differently from the sequential statement.  (Segments longer than
// COVER_DUP_MAX are not checked.)
constexpr int COVER_DUP_MAX = 4096;
__global__ void __launch_bounds__(128)
dup_edge_kernel(int32_t N, const int32_t* __restrict__ optr, const int32_t* __restrict__ odst, int32_t* __restrict__ err)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int v = blockIdx.x * 4 + wid; v < N; v += gridDim.x * 4) {
        const int o0 = optr[v], d = optr[v + 1] - o0;
        bool dup = false;
        if (d <= COVER_DUP_MAX)
            for (int q = lane; q < d; q += 32) {
                const int32_t t = odst[o0 + q];
                for (int r = 0; r < q; ++r) dup |= odst[o0 + r] == t;
            }
        if (__any_sync(0xffffffffu, dup) && lane == 0) atomicOr(err, 16);
    }
}
__global__ void node_desc_kernel(int32_t N, const int32_t* __restrict__ optr, const int32_t* __restrict__ iptr, int4* __restrict__ nd)
{
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < N; v += gridDim.x * blockDim.x)
        nd[v] = make_int4(optr[v], optr[v + 1], iptr[v], iptr[v + 1]);
}
__global__ void max_degree_kernel(int32_t N, const int32_t* __restrict__ odeg, const int32_t* __restrict__ ideg, int32_t* __restrict__ out2)
{
    int mo = 0, mi = 0;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < N; v += gridDim.x * blockDim.x) { mo = max(mo, odeg[v]); mi = max(mi, ideg[v]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mo = max(mo, __shfl_xor_sync(0xffffffffu, mo, o)); mi = max(mi, __shfl_xor_sync(0xffffffffu, mi, o)); }
    if ((threadIdx.x & 31) == 0) { atomicMax(&out2[0], mo); atomicMax(&out2[1], mi); }
}
__global__ void sample_check_kernel(int64_t total, int32_t n_s, const int32_t* __restrict__ samples, int32_t N, int32_t* __restrict__ err)
{
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        const int32_t v = samples[q];
        if (v < 0 || v >= N) atomicOr(err, 2);
        if (q % n_s != 0 && v <= samples[q - 1]) atomicOr(err, 4);
    }
}

void launch_build_cover_csr(mcb_ctx* ctx, int32_t N, int64_t ne, const int32_t* src, const int32_t* dst, const double* w,
                            int32_t* odeg /*[N+1] zeroed*/, int32_t* ideg /*[N+1] zeroed*/, int32_t* optr /*[N+1]*/,
                            int32_t* iptr /*[N+1]*/, int32_t* odst, uint32_t* ow, int32_t* isrc, uint32_t* iw,
                            void* oe /* int2[ne] */, void* ie /* int2[ne] */, void* nd /* int4[N] */, uint16_t* ipos /* [ne] */,
                            bool sort_out, int32_t* scratch4 /* [0] total, [1] err, [2] max out, [3] max in; zeroed */)
{
    const int grid = (int)std::min<int64_t>(std::max<int64_t>(ceil_div(ne, 256), 1), (int64_t)ctx->num_sms * 8);
    if (ne > 0) { edge_degree_kernel<<<grid, 256, 0, ctx->stream>>>(ne, src, dst, N, odeg, ideg, scratch4 + 1); MCB_LAUNCH_CHECK(); }
    max_degree_kernel<<<std::min(ctx->num_sms, (int)ceil_div(N, 256)), 256, 0, ctx->stream>>>(N, odeg, ideg, scratch4 + 2);
    MCB_LAUNCH_CHECK();
    launch_exclusive_scan_i32(ctx, odeg, optr, (int64_t)N + 1, scratch4);
    launch_exclusive_scan_i32(ctx, ideg, iptr, (int64_t)N + 1, scratch4);
    MCB_CUDA(cudaMemsetAsync(odeg, 0, sizeof(int32_t) * ((size_t)N + 1), ctx->stream));     // reused as fill cursors
    MCB_CUDA(cudaMemsetAsync(ideg, 0, sizeof(int32_t) * ((size_t)N + 1), ctx->stream));
    if (ne > 0) {
        edge_fill_out_kernel<<<grid, 256, 0, ctx->stream>>>(ne, src, dst, w, N, optr, odeg, odst, ow);
        MCB_LAUNCH_CHECK();
        const int fgrid = (int)std::min<int64_t>(ceil_div(N, 4), (int64_t)ctx->num_sms * 16);
        if (sort_out) finish_csr_kernel<true><<<fgrid, 128, 0, ctx->stream>>>(N, optr, iptr, ideg, odst, ow, (int2*)oe, isrc, iw, (int2*)ie, ipos, scratch4 + 1);
        else finish_csr_kernel<false><<<fgrid, 128, 0, ctx->stream>>>(N, optr, iptr, ideg, odst, ow, (int2*)oe, isrc, iw, (int2*)ie, ipos, scratch4 + 1);
        MCB_LAUNCH_CHECK();
    }
    node_desc_kernel<<<std::min(ctx->num_sms, (int)ceil_div(N, 256)), 256, 0, ctx->stream>>>(N, optr, iptr, (int4*)nd);
    MCB_LAUNCH_CHECK();
    if (ne > 0) {
        dup_edge_kernel<<<(int)std::min<int64_t>(ceil_div(N, 4), (int64_t)ctx->num_sms * 16), 128, 0, ctx->stream>>>(N, optr, odst, scratch4 + 1);
        MCB_LAUNCH_CHECK();
    }
}
__global__ void iota_i32_kernel(int32_t* __restrict__ v, int32_t n)
{
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = i;
}
void launch_iota_i32(mcb_ctx* ctx, int32_t* v, int32_t n)
{
    if (n == 0) return;
    iota_i32_kernel<<<std::min(ctx->num_sms * 4, (int)ceil_div(n, 256)), 256, 0, ctx->stream>>>(v, n);
    MCB_LAUNCH_CHECK();
}

void launch_sample_check(mcb_ctx* ctx, int64_t total, int32_t n_s, const int32_t* samples, int32_t N, int32_t* err)
{
    if (total == 0) return;
    const int grid = (int)std::min<int64_t>(ceil_div(total, 256), (int64_t)ctx->num_sms * 8);
    sample_check_kernel<<<grid, 256, 0, ctx->stream>>>(total, n_s, samples, N, err);
    MCB_LAUNCH_CHECK();
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

// ---------------------------------------------------------------------------------------------
// method = "edges": co-occurrence counted only for the pairs that are edges of the graph, one counter per edge in the
// order of the out-CSR (which is ordered by (source, weight descending, target ascending) for this mode).  Memory is
// O(edges) instead of O(N^2): this is the variant that reaches cell counts where the N x N matrix of method = "full"
// cannot exist (10^6 cells: 4 TB).  One warp per (resample, node): edges into the node's own cluster are counted.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cooccur_edges_kernel(int32_t N, int32_t n_local, const int32_t* __restrict__ mc_all, const int32_t* __restrict__ optr,
                     const int32_t* __restrict__ odst, uint32_t* __restrict__ cnt_e, int32_t* __restrict__ nsamp)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t items = (int64_t)n_local * N;
    for (int64_t it = warp0; it < items; it += nwarps) {        // resample-major: one resample's assignment stays in L2
        const int64_t lr = it / N;
        const int32_t v = (int32_t)(it - lr * N);
        const int32_t* mc = mc_all + lr * N;
        const int32_t k = mc[v];
        if (k == -2) continue;
        if (lane == 0) atomicAdd(&nsamp[v], 1);
        if (k < 0) continue;
        for (int32_t e = optr[v] + lane; e < optr[v + 1]; e += 32)
            if (mc[odst[e]] == k) atomicAdd(&cnt_e[e], 1u);
    }
}
__global__ void expand_src_kernel(int32_t N, const int32_t* __restrict__ optr, int32_t* __restrict__ esrc)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t v = warp0; v < N; v += nwarps)
        for (int32_t e = optr[v] + lane; e < optr[v + 1]; e += 32) esrc[e] = (int32_t)v;
}
void launch_cooccur_edges(mcb_ctx* ctx, int32_t N, int32_t n_local, const int32_t* mc, const int32_t* optr32, const int32_t* odst,
                          uint32_t* cnt_e, int32_t* nsamp)
{
    if (n_local == 0 || N == 0) return;
    const int grid = (int)std::min<int64_t>(ceil_div((int64_t)n_local * N, 8), (int64_t)ctx->num_sms * 8);
    cooccur_edges_kernel<<<grid, 256, 0, ctx->stream>>>(N, n_local, mc, optr32, odst, cnt_e, nsamp);
    MCB_LAUNCH_CHECK();
}
void launch_expand_src(mcb_ctx* ctx, int32_t N, const int32_t* optr32, int32_t* esrc)
{
    if (N == 0) return;
    const int grid = (int)std::min<int64_t>(ceil_div(N, 8), (int64_t)ctx->num_sms * 8);
    expand_src_kernel<<<grid, 256, 0, ctx->stream>>>(N, optr32, esrc);
    MCB_LAUNCH_CHECK();
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
    static SmemOptIn optin;
    ensure_dyn_smem(optin, ctx, graph_cover_fast_kernel<FT, IN_SLOTS>, smem);
    const int32_t out_warps = std::max(1, (int)ceil_div(max_out, 32));
    graph_cover_fast_kernel<FT, IN_SLOTS><<<n_local, FT, smem, ctx->stream>>>(fg, n_local, resamp_ids, n_s, samples, seeds,
                                                                               min_size, cooling, burn_in, max_sweeps, ncl_cap,
                                                                               out_warps, mc, f_ws, sweeps, status);
    MCB_LAUNCH_CHECK();
}

template <int B, int OUT_SLOTS, int IN_SLOTS, typename VT, bool TRUNC>
static void launch_cover_batch_t(mcb_ctx* ctx, const FastGraph& fg, int32_t n_local, const int32_t* resamp_ids, int32_t n_s,
                                 const int32_t* samples, const uint64_t* seeds, int32_t min_size, double cooling, int32_t burn_in,
                                 int32_t max_sweeps, int32_t ncl_cap, int32_t max_out, size_t smem, int32_t* mc, int32_t* f_ws,
                                 int32_t* sweeps, int32_t* status, int32_t knn, const uint16_t* ipos, uint16_t* ocut)
{
    static SmemOptIn optin;
    ensure_dyn_smem(optin, ctx, graph_cover_batch_kernel<B, OUT_SLOTS, IN_SLOTS, VT, TRUNC>, smem);
    const int32_t out_warps = std::max(1, (int)ceil_div(max_out, 32));
    graph_cover_batch_kernel<B, OUT_SLOTS, IN_SLOTS, VT, TRUNC><<<n_local, 128 * B, smem, ctx->stream>>>(fg, n_local, resamp_ids, n_s, samples, seeds,
                                                                                    min_size, cooling, burn_in, max_sweeps,
                                                                                    ncl_cap, out_warps, mc, f_ws, sweeps, status, knn, ipos, ocut);
    MCB_LAUNCH_CHECK();
}
template <int B, int OUT_SLOTS, int IN_SLOTS, typename VT>
static void launch_cover_batch(mcb_ctx* ctx, const FastGraph& fg, int32_t n_local, const int32_t* resamp_ids, int32_t n_s,
                               const int32_t* samples, const uint64_t* seeds, int32_t min_size, double cooling, int32_t burn_in,
                               int32_t max_sweeps, int32_t ncl_cap, int32_t max_out, size_t smem, int32_t* mc, int32_t* f_ws,
                               int32_t* sweeps, int32_t* status, int32_t knn, const uint16_t* ipos, uint16_t* ocut)
{
    if (knn > 0) launch_cover_batch_t<B, OUT_SLOTS, IN_SLOTS, VT, true>(ctx, fg, n_local, resamp_ids, n_s, samples, seeds, min_size, cooling, burn_in,
                                                                        max_sweeps, ncl_cap, max_out, smem, mc, f_ws, sweeps, status, knn, ipos, ocut);
    else launch_cover_batch_t<B, OUT_SLOTS, IN_SLOTS, VT, false>(ctx, fg, n_local, resamp_ids, n_s, samples, seeds, min_size, cooling, burn_in,
                                                                 max_sweeps, ncl_cap, max_out, smem, mc, f_ws, sweeps, status, 0, nullptr, nullptr);
}

template <int B, int OUT_SLOTS>
static void launch_cover_warp(mcb_ctx* ctx, const FastGraph& fg, int32_t n_local, const int32_t* resamp_ids, int32_t n_s,
                              const int32_t* samples, const uint64_t* seeds, int32_t min_size, double cooling, int32_t burn_in,
                              int32_t max_sweeps, int32_t ncl_cap, int32_t kin_stride, size_t smem, int32_t* mc, int32_t* f_ws,
                              int32_t* sweeps, int32_t* status)
{
    static SmemOptIn optin;
    ensure_dyn_smem(optin, ctx, graph_cover_warp_kernel<B, OUT_SLOTS>, smem);
    graph_cover_warp_kernel<B, OUT_SLOTS><<<n_local, 32 * B, smem, ctx->stream>>>(fg, n_local, resamp_ids, n_s, samples, seeds,
                                                                                   min_size, cooling, burn_in, max_sweeps, ncl_cap,
                                                                                   kin_stride, mc, f_ws, sweeps, status);
    MCB_LAUNCH_CHECK();
}

// rows: 2 x N x stride_max words per resident CTA.  The CTA count is cut (down to one per SM) until the workspace fits
// `budget` bytes; returns false when even that does not fit (the caller then takes an edge-reading kernel).
template <int B, typename VT, bool TRUNC>
static int cover_incr_occupancy(mcb_ctx* ctx, size_t smem)
{
    static SmemOptIn optin;
    auto* kern = graph_cover_incr_kernel<B, VT, TRUNC>;
    ensure_dyn_smem(optin, ctx, kern, smem);
    int occ = 0;
    MCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * B, smem));
    return std::max(occ, 1);
}
template <int B, typename VT, bool TRUNC>
static bool launch_cover_incr_t(mcb_ctx* ctx, const FastGraph& fg, int32_t n_local, const int32_t* resamp_ids, int32_t n_s,
                                const int32_t* samples, const uint64_t* seeds, int32_t min_size, double cooling, int32_t burn_in,
                                int32_t max_sweeps, int32_t ncl_cap, size_t smem, int32_t* mc, int32_t* f_ws, int32_t* sweeps,
                                int32_t* status, int32_t knn, const uint16_t* ipos, uint16_t* ocut, size_t budget)
{
    const int occ = cover_incr_occupancy<B, VT, TRUNC>(ctx, smem);
    const size_t slot_words = (size_t)fg.N * (size_t)round_up((int64_t)std::max(ncl_cap, 1), 32);
    int slots = std::max(1, std::min(n_local, occ * ctx->num_sms));
    const size_t fit = budget / (slot_words * 2 * sizeof(uint32_t));
    if (fit < (size_t)std::min(n_local, ctx->num_sms)) return false;
    slots = (int)std::min<size_t>((size_t)slots, fit);
    DevBuf<uint32_t> rows(ctx, (size_t)slots * 2 * slot_words);
    DevBuf<int32_t> next(ctx, 1);
    next.zero();
    const IncrWs ws{rows.p, next.p, (int64_t)slot_words};
    graph_cover_incr_kernel<B, VT, TRUNC><<<slots, 32 * B, smem, ctx->stream>>>(fg, n_local, resamp_ids, n_s, samples, seeds, min_size, cooling,
                                                                                 burn_in, max_sweeps, ncl_cap, mc, f_ws, sweeps, status, knn,
                                                                                 ipos, ocut, ws);
    MCB_LAUNCH_CHECK();
    return true;
}
// want_b: 16 or 8 forced, 0 = by occupancy: 16 warps per CTA unless the resamples need more CTA slots than that gives AND
// the 8-warp instantiation doubles the CTAs per SM (C3: 500 resamples over 296 slots would run in two waves; C5: both
// instantiations fit two CTAs per SM, so the larger one keeps twice the warps busy)
template <typename VT, bool TRUNC>
static bool launch_cover_incr_v(mcb_ctx* ctx, int want_b, const FastGraph& fg, int32_t n_local, const int32_t* resamp_ids, int32_t n_s,
                                const int32_t* samples, const uint64_t* seeds, int32_t min_size, double cooling, int32_t burn_in,
                                int32_t max_sweeps, int32_t ncl_cap, size_t smem, int32_t* mc, int32_t* f_ws, int32_t* sweeps,
                                int32_t* status, int32_t knn, const uint16_t* ipos, uint16_t* ocut, size_t budget)
{
    int b = want_b;
    if (b == 0) {
        const int occ16 = cover_incr_occupancy<16, VT, TRUNC>(ctx, smem), occ8 = cover_incr_occupancy<8, VT, TRUNC>(ctx, smem);
        b = (n_local > occ16 * ctx->num_sms && occ8 >= 2 * occ16) ? 8 : 16;
    }
    if (b == 8) return launch_cover_incr_t<8, VT, TRUNC>(ctx, fg, n_local, resamp_ids, n_s, samples, seeds, min_size, cooling, burn_in, max_sweeps,
                                                         ncl_cap, smem, mc, f_ws, sweeps, status, knn, ipos, ocut, budget);
    return launch_cover_incr_t<16, VT, TRUNC>(ctx, fg, n_local, resamp_ids, n_s, samples, seeds, min_size, cooling, burn_in, max_sweeps,
                                              ncl_cap, smem, mc, f_ws, sweeps, status, knn, ipos, ocut, budget);
}
template <typename VT>
static bool launch_cover_incr(mcb_ctx* ctx, int want_b, const FastGraph& fg, int32_t n_local, const int32_t* resamp_ids, int32_t n_s,
                              const int32_t* samples, const uint64_t* seeds, int32_t min_size, double cooling, int32_t burn_in,
                              int32_t max_sweeps, int32_t ncl_cap, size_t smem, int32_t* mc, int32_t* f_ws, int32_t* sweeps,
                              int32_t* status, int32_t knn, const uint16_t* ipos, uint16_t* ocut, size_t budget)
{
    if (knn > 0) return launch_cover_incr_v<VT, true>(ctx, want_b, fg, n_local, resamp_ids, n_s, samples, seeds, min_size, cooling, burn_in,
                                                      max_sweeps, ncl_cap, smem, mc, f_ws, sweeps, status, knn, ipos, ocut, budget);
    return launch_cover_incr_v<VT, false>(ctx, want_b, fg, n_local, resamp_ids, n_s, samples, seeds, min_size, cooling, burn_in,
                                          max_sweeps, ncl_cap, smem, mc, f_ws, sweeps, status, 0, nullptr, nullptr, budget);
}

void launch_graph_cover(mcb_ctx* ctx, const CoverGraph& g, const int32_t* optr32, const int32_t* iptr32, const void* nd,
                        const void* oe, const void* ie, int32_t max_out, int32_t max_in, int32_t n_local, const int32_t* resamp_ids, int32_t n_s,
                        const int32_t* samples, const uint64_t* seeds, int32_t min_size, double cooling,
                        int32_t burn_in, int32_t max_sweeps, int32_t ncl_cap, int32_t* mc, int32_t* f_ws,
                        int32_t* sweeps, int32_t* status, int32_t knn, const uint16_t* ipos, uint16_t* ocut_ws)
{
    // knn > 0: per-node edge truncation is active (the caller passes 0 when knn >= the largest out-degree, where it can never
    // bind).  Only the speculative batch kernels and the general kernel implement it.
    if (n_local == 0) return;
    // Kernel choice (MCB_COVER_KERNEL overrides for tests/diagnostics: incr8|incr16|batch2|batch4|batch8|warp8|warp16|warp32|fast|slow):
    //   incr16  per-node cluster rows edited on moves, 16 speculative visits of one warp each (any degree; needs
    //           2 x N x clusters words of workspace per resident CTA); incr8 when that doubles the CTAs per SM and the
    //           resamples need the slots
    //   batch4  speculative batches of 4 visits, 4 warps per node  (out-degree <= 128, in-degree <= 384)
    //   warp16  speculative batches of 16 visits, 1 warp per node  (out-degree <= 256, any in-degree)
    //   fast    sequential visits, state in shared memory
    //   slow    sequential visits, mc[] in global memory (any size)
    const char* ksel = getenv("MCB_COVER_KERNEL");
    const std::string want = ksel ? ksel : (getenv("MCB_COVER_SLOW") ? "slow" : "auto");
    const bool small_state = optr32 && nd && oe && ie && ncl_cap < 32000;
    if (small_state && want != "slow") {
        const FastGraph fg{g.N, optr32, g.odst, g.ow, iptr32, g.isrc, g.iw, (const int4*)nd, (const int2*)oe, (const int2*)ie};
        const bool auto_ = want == "auto";
        // batches of 4 visits minimise one resample's latency (37.7 ms at C3, two 512-thread CTAs per SM); once more
        // resamples are queued than those 2 x num_sms slots the machine is issue-bound and batches of 2 (less
        // speculative waste, four 256-thread CTAs per SM) give more throughput: 500 resamples 99 ms vs 112 ms
        const int auto_b = n_local > 2 * ctx->num_sms ? 2 : 4;
        {
            // the row workspace may take a quarter of the memory that is free or held by the context's own block cache (so a
            // repeated call sizes it the same way and finds its block again)
            size_t free_b = 0, total_b = 0;
            MCB_CUDA(cudaMemGetInfo(&free_b, &total_b));
            const size_t budget = (free_b + ctx->cached_bytes) / 4;
            const int want_b = want == "incr16" ? 16 : want == "incr8" ? 8 : auto_ ? 0 : -1;
            if (want_b >= 0) {
                const bool small_ids = g.N < 65536;          // 16-bit visit order
                const size_t ismem = (size_t)ncl_cap * sizeof(uint32_t) + (size_t)((n_s + 1) & ~1) * (small_ids ? 2 : 4) +
                                     (size_t)g.N * sizeof(int16_t) + 16;
                if (ismem <= 200 * 1024) {
                    const bool ok = small_ids
                        ? launch_cover_incr<uint16_t>(ctx, want_b, fg, n_local, resamp_ids, n_s, samples, seeds, min_size, cooling, burn_in,
                                                      max_sweeps, ncl_cap, ismem, mc, f_ws, sweeps, status, knn, ipos, ocut_ws, budget)
                        : launch_cover_incr<int32_t>(ctx, want_b, fg, n_local, resamp_ids, n_s, samples, seeds, min_size, cooling, burn_in,
                                                     max_sweeps, ncl_cap, ismem, mc, f_ws, sweeps, status, knn, ipos, ocut_ws, budget);
                    if (ok) return;
                }
            }
        }
#define BATCH(BB, OS, IS, VT, AUTO)                                                                                       \
    if ((auto_ && (AUTO)) || want == "batch" #BB) {                                                                      \
        const size_t bsmem = (size_t)(1 + 4 * BB) * ncl_cap * sizeof(uint32_t) + (size_t)((n_s + 1) & ~1) * sizeof(VT) +  \
                             (size_t)g.N * sizeof(int16_t) + 16;                                                          \
        if (max_out <= OS * 128 && max_in <= IS * 128 && bsmem <= 200 * 1024 && (sizeof(VT) == 4 || g.N < 65536)) {       \
            launch_cover_batch<BB, OS, IS, VT>(ctx, fg, n_local, resamp_ids, n_s, samples, seeds, min_size, cooling,      \
                                               burn_in, max_sweeps, ncl_cap, max_out, bsmem, mc, f_ws, sweeps, status,    \
                                               knn, ipos, ocut_ws);                                                       \
            return;                                                                                                       \
        }                                                                                                                 \
    }
        BATCH(4, 1, 3, int32_t, auto_b == 4) BATCH(2, 1, 3, int32_t, auto_b == 2) BATCH(8, 1, 3, int32_t, false)
        // K = 150 graphs: out-degree <= 256, in-degree <= 512.  With a 16-bit visit order two CTAs of the 2-visit
        // instantiation fit an SM (C5: 96 KB each); otherwise one CTA of the 4-visit one
        BATCH(2, 2, 4, uint16_t, n_local > ctx->num_sms)
        BATCH(4, 2, 4, int32_t, true)
        BATCH(2, 1, 3, int32_t, true)             // batches of 4 did not fit in shared memory
#undef BATCH
        const int32_t kin_stride = (int32_t)round_up((int64_t)std::max(std::min(max_in, ncl_cap), 1), 32) + 2;
#define WARPK(BB, OS)                                                                                                    \
    if ((auto_ && BB == 16) || want == "warp" #BB) {                                                                      \
        const size_t wsmem = ((size_t)(1 + 2 * BB) * ncl_cap + (size_t)n_s) * sizeof(uint32_t) +                         \
                             ((size_t)((g.N + 1) & ~1) + (size_t)BB * kin_stride) * sizeof(int16_t) + 16;                \
        if (max_out <= 32 * OS && wsmem <= 200 * 1024) {                                                                  \
            launch_cover_warp<BB, OS>(ctx, fg, n_local, resamp_ids, n_s, samples, seeds, min_size, cooling, burn_in,     \
                                      max_sweeps, ncl_cap, kin_stride, wsmem, mc, f_ws, sweeps, status);                  \
            return;                                                                                                       \
        }                                                                                                                 \
    }
        if (knn == 0) {
            if (max_out <= 128) { WARPK(16, 4) WARPK(8, 4) WARPK(32, 4) }
            else { WARPK(16, 8) WARPK(8, 8) }
        }
#undef WARPK
        // fast kernels: size + 2 x (wout, win) + visit order + int16 mc[] in shared memory, edges of a node in registers
        const size_t fast_smem = ((size_t)5 * ncl_cap + (size_t)n_s) * sizeof(uint32_t) + (size_t)g.N * sizeof(int16_t) + 16;
        if (fast_smem <= 200 * 1024 && knn == 0) {
#define FAST(FT, SL) launch_cover_fast<FT, SL>(ctx, fg, n_local, resamp_ids, n_s, samples, seeds, min_size, cooling, burn_in, \
                                               max_sweeps, ncl_cap, max_out, fast_smem, mc, f_ws, sweeps, status)
            if (max_out <= 128 && max_in <= 3 * 128) { FAST(128, 3); return; }
            if (max_out <= 256 && max_in <= 3 * 256) { FAST(256, 3); return; }
            if (max_out <= 256 && max_in <= 6 * 256) { FAST(256, 6); return; }
#undef FAST
        }
    }
    size_t smem = (size_t)ncl_cap * 3 * sizeof(uint32_t);
    DevBuf<uint32_t> acc_ws;                   // accumulators of every CTA in global memory when they exceed shared memory
    int grid = n_local;
    if (smem > 200 * 1024) {
        grid = std::min(n_local, ctx->num_sms * 4);
        acc_ws.alloc(ctx, (size_t)grid * 3 * ncl_cap);
        smem = 0;
    }
    static SmemOptIn optin;
    ensure_dyn_smem(optin, ctx, graph_cover_kernel, smem);
    graph_cover_kernel<<<grid, CC_THREADS, smem, ctx->stream>>>(g, n_local, resamp_ids, n_s, samples, seeds, min_size,
                                                                 cooling, burn_in, max_sweeps, ncl_cap, mc, f_ws,
                                                                 sweeps, status, knn, ipos, ocut_ws, acc_ws.p);
    MCB_LAUNCH_CHECK();
}

void launch_cooccur(mcb_ctx* ctx, int32_t N, int32_t n_local, const int32_t* mc, int32_t ncl_cap, uint32_t* cnt,
                    int32_t* nsamp, int32_t* ws_order, int32_t* ws_start)
{
    if (n_local == 0) return;
    const size_t smem = (size_t)(ncl_cap + 1) * sizeof(int32_t);
    static SmemOptIn optin;
    ensure_dyn_smem(optin, ctx, cooccur_kernel, smem);
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
