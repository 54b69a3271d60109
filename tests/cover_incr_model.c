/* cover_incr_model.c -- TEST INFRASTRUCTURE.  Host model of the data structure graph_cover_incr_kernel
 * (metacell_b200/csrc/coclust.cu) keeps: instead of re-reading a visited node's edges, every node carries two dense rows
 * wout[v][k] = summed weight of its out edges into cluster k, win[v][k] = summed weight of its in edges from cluster k;
 * a move of node v from cluster a to cluster b edits two entries in the row of each of v's neighbours.  All sums are
 * integers, so the scores -- and with them every decision -- are the ones orc_graph_cover computes from the edges:
 * tests/test_oracle.py asserts identical assignments and sweep counts, and reads the statistic the kernel rests on
 * (moves per sweep: most visits move nothing).  It also replays the kernel's fp32 pre-decision beside every exact
 * decision: whenever the best fp32 score leads the runner-up by more than 1e-5 (relative) the fp32 arg-max must be the
 * exact one; the visits that do not clear that margin are the ones the kernel scores in fp64.
 *
 * Includes the oracle's translation unit for its static helpers (Philox stream, keyed permutation).  Seeding is the
 * oracle's own code path restated; only the sweeps differ.
 */
#include "../oracle/mc_oracle.c"

int model_graph_cover_incr(int32_t N, const int64_t* optr, const int32_t* odst, const uint32_t* ow,
                           const int64_t* iptr, const int32_t* isrc, const uint32_t* iw,
                           int32_t n_s, const int32_t* sample, int32_t min_size, double cooling, int32_t burn_in,
                           uint64_t seed, int32_t max_sweeps, int32_t* mc, int32_t* n_sweeps_out,
                           int64_t* moves_per_sweep /* [max_sweeps] or NULL */,
                           int64_t* fp32_stats /* [3] or NULL: visits with a candidate, visits inside the 1e-5 margin (exact
                                                  path), visits where a clear fp32 decision differs from the exact one */)
{
    int32_t* f = (int32_t*)calloc((size_t)N, sizeof(int32_t));
    int32_t ncl_cap = n_s / (min_size > 0 ? min_size + 1 : 1) + 2;
    int32_t* size = (int32_t*)calloc((size_t)ncl_cap, sizeof(int32_t));
    if (!f || !size) return 2;
    for (int32_t v = 0; v < N; ++v) mc[v] = -2;
    for (int32_t q = 0; q < n_s; ++q) mc[sample[q]] = -1;
    for (int32_t q = 0; q < n_s; ++q) {
        int32_t v = sample[q], c = 0;
        for (int64_t e = optr[v]; e < optr[v + 1]; ++e) c += (mc[odst[e]] == -1);
        f[v] = c;
    }
    int32_t ncl = 0;
    for (uint32_t iter = 0;; ++iter) {           /* seeding: as orc_graph_cover */
        uint64_t total = 0;
        for (int32_t v = 0; v < N; ++v)
            if (mc[v] == -1 && f[v] >= min_size && f[v] > 0) total += (uint64_t)f[v] * f[v] * f[v];
        if (total == 0) break;
        uint64_t r = rng_u64(seed, ORC_STREAM_SEED, iter, 0) % total;
        int32_t pick = -1; uint64_t cum = 0;
        for (int32_t v = 0; v < N; ++v) {
            if (mc[v] == -1 && f[v] >= min_size && f[v] > 0) {
                cum += (uint64_t)f[v] * f[v] * f[v];
                if (cum > r) { pick = v; break; }
            }
        }
        if (ncl >= ncl_cap) return 3;
        int32_t k = ncl++;
        mc[pick] = k; size[k] = 1;
        for (int64_t e = iptr[pick]; e < iptr[pick + 1]; ++e) f[isrc[e]]--;
        for (int64_t e = optr[pick]; e < optr[pick + 1]; ++e) {
            int32_t u = odst[e];
            if (mc[u] != -1) continue;
            mc[u] = k; size[k]++;
            for (int64_t e2 = iptr[u]; e2 < iptr[u + 1]; ++e2) f[isrc[e2]]--;
        }
    }
    /* rows of every node from the seeded state; the row stride is the number of clusters the seeding made */
    const size_t stride = (size_t)(ncl > 0 ? ncl : 1);
    uint32_t* wout = (uint32_t*)calloc((size_t)N * stride, sizeof(uint32_t));
    uint32_t* win = (uint32_t*)calloc((size_t)N * stride, sizeof(uint32_t));
    if (!wout || !win) return 2;
    for (int32_t q = 0; q < n_s; ++q) {
        int32_t v = sample[q];
        for (int64_t e = optr[v]; e < optr[v + 1]; ++e) {        /* the edge v -> u feeds wout[v][cluster of u], win[u][cluster of v] */
            int32_t u = odst[e];
            if (mc[u] == -2) continue;
            if (mc[u] >= 0) wout[(size_t)v * stride + mc[u]] += ow[e];
            if (mc[v] >= 0) win[(size_t)u * stride + mc[v]] += ow[e];
        }
    }
    const uint32_t hb = half_bits_for((uint32_t)n_s);
    double factor = 1.0;
    int32_t sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        if (sweep > burn_in) factor = factor * cooling;
        uint32_t rk[4];
        for (int r = 0; r < 4; r += 2) {
            uint64_t w = rng_u64(seed, ORC_STREAM_PERM, (uint32_t)sweep, (uint32_t)(r >> 1));
            rk[r] = (uint32_t)w; rk[r + 1] = (uint32_t)(w >> 32);
        }
        int32_t moved = 0;
        for (int32_t t = 0; t < n_s; ++t) {
            int32_t v = sample[perm_index((uint32_t)t, (uint32_t)n_s, hb, rk)];
            int32_t best = -1; double best_score = 0.0;
            float fbest = 0.f, fsecond = 0.f; int32_t fk = -1;                       /* the kernel's fp32 pre-decision */
            for (int32_t k = 0; k < ncl; ++k) {
                const uint32_t wo = wout[(size_t)v * stride + k], wi = win[(size_t)v * stride + k];
                if (wo == 0 || wi == 0) continue;
                double sz = (double)size[k];
                double sc = (double)((int64_t)wo * (int64_t)wi) / (sz * sz);
                if (k == mc[v]) sc = sc * factor;
                if (best < 0 || sc > best_score) { best = k; best_score = sc; }      /* ascending k: ties keep the lower cluster */
                float szf = (float)size[k];
                float fs = ((float)wo * (float)wi) / (szf * szf);
                if (k == mc[v]) fs = fs * (float)factor;
                if (fs > fbest) { fsecond = fbest; fbest = fs; fk = k; } else if (fs > fsecond) fsecond = fs;
            }
            if (fp32_stats && best >= 0) {
                fp32_stats[0]++;
                if (fsecond >= fbest * (1.0f - 1e-5f)) fp32_stats[1]++;
                else if (fk != best) fp32_stats[2]++;
            }
            if (best >= 0 && best != mc[v]) {
                const int32_t a = mc[v], b = best;
                if (a >= 0) size[a]--;
                mc[v] = b; size[b]++; ++moved;
                /* v's out edges are in edges of their targets; v's in edges are out edges of their sources */
                for (int64_t e = optr[v]; e < optr[v + 1]; ++e) {
                    uint32_t* row = win + (size_t)odst[e] * stride;
                    if (mc[odst[e]] == -2) continue;
                    if (a >= 0) row[a] -= ow[e];
                    row[b] += ow[e];
                }
                for (int64_t e = iptr[v]; e < iptr[v + 1]; ++e) {
                    uint32_t* row = wout + (size_t)isrc[e] * stride;
                    if (mc[isrc[e]] == -2) continue;
                    if (a >= 0) row[a] -= iw[e];
                    row[b] += iw[e];
                }
            }
        }
        if (moves_per_sweep) moves_per_sweep[sweep] = moved;
        if (moved == 0) { ++sweep; break; }
    }
    if (n_sweeps_out) *n_sweeps_out = sweep;
    free(f); free(size); free(wout); free(win);
    return 0;
}
