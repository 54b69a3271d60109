// gstats.cu -- per-gene statistics of scm_gene_stat (reference R/gstats.r:65-205): the step
// right before the hot path (it selects the feature genes) and already a client of scm_downsamp (:99).
//
// What the reference computes per gene over the raw matrix (all n cells) and the downsampled one (n_ds kept
// cells), restated from R/gstats.r:108-131:
//   tot = rowSums, var = rowVars (n-1 denominator, zeros included), is_on_count = #(x > 0),
//   niche_stat = (10 + sum of the round(n*q) largest values) / (10 + tot)            (quant_mean, :83-87)
//   n_mean = rowMeans of x * K_std_n / colSums(mat)                                   (:102, :120)
//   ds_top1/2/3 = three largest downsampled values, ds_mean, ds_var, ds_is_on_count  (:121-127)
//   sz_cor = cor(x[g, subs], colSums[subs])                                           (:143-151)
// The rolling-median trends after that (:153-190) are O(genes) host work and stay with the caller.
//
// Layout: the CSC is cell-major, the statistics are per gene, so the entries are first bucketed by gene
// (count, scan, scatter: values, cell ids and the downsampled values of the same entry land in gene-major
// arrays), then one CTA per gene streams its contiguous segment.  Integer sums are exact (64/128-bit); only
// n_mean is a floating-point sum.  HBM-bound: 12 B per entry read by the scatter, 12 B written, 12 B read by
// the per-gene pass, plus the radix passes of the niche selection over the gene's (L2-resident) segment.
#include "internal.h"

namespace mcb {

typedef unsigned __int128 u128;

// Entries per gene.  Hot genes receive one update per cell, so the counters are privatised per CTA in shared
// memory (ngenes * 4 B) and flushed once; gene_count_global_kernel is the fallback for very long gene lists.
__global__ void gene_count_global_kernel(int64_t nnz, const int32_t* __restrict__ gi, uint32_t* __restrict__ cnt)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&cnt[gi[e]], 1u);
}
__global__ void __launch_bounds__(512)
gene_count_kernel(int32_t ngenes, int64_t nnz, const int32_t* __restrict__ gi, uint32_t* __restrict__ cnt)
{
    extern __shared__ uint32_t sh[];
    for (int g = threadIdx.x; g < ngenes; g += blockDim.x) sh[g] = 0;
    __syncthreads();
    const int64_t per = (nnz + gridDim.x - 1) / gridDim.x;
    const int64_t e0 = (int64_t)blockIdx.x * per, e1 = min(nnz, e0 + per);
    for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) atomicAdd(&sh[gi[e]], 1u);
    __syncthreads();
    for (int g = threadIdx.x; g < ngenes; g += blockDim.x) { const uint32_t v = sh[g]; if (v) atomicAdd(&cnt[g], v); }
}

// single block: start[g] = exclusive prefix of cnt (int64), start[ngenes] = total
__global__ void __launch_bounds__(1024) gene_scan_kernel(int32_t ngenes, const uint32_t* __restrict__ cnt, int64_t* __restrict__ start)
{
    __shared__ long long warp_tot[33];
    __shared__ long long s_run;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < ngenes; base += 1024) {
        const int g = base + threadIdx.x;
        const long long v = g < ngenes ? (long long)cnt[g] : 0;
        long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const long long u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
        if (lane == 31) warp_tot[w] = inc;
        __syncthreads();
        long long pre = 0, all = 0;
        for (int q = 0; q < 32; ++q) { const long long t = warp_tot[q]; if (q < w) pre += t; all += t; }
        const long long run = s_run;
        if (g < ngenes) start[g] = run + pre + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) s_run = run + all;
        __syncthreads();
    }
    if (threadIdx.x == 0) start[ngenes] = s_run;
}

// Scatter into the gene segments: one warp per cell, positions handed out by a global cursor per gene.  Because
// the cursors advance in global time order, the write frontier is one or two sectors per gene and array (a few MB
// in total), stays L2-resident and the 4-byte stores merge there.  (A per-CTA run reservation was tried: it removes
// the atomic hot spots but multiplies the frontier by the CTA count, thrashes L2 and ran 3x slower.)
__global__ void gene_scatter_kernel(int64_t ncells, const int64_t* __restrict__ p, const int32_t* __restrict__ gi,
                                           const uint32_t* __restrict__ x, const uint32_t* __restrict__ dsx,
                                           const int64_t* __restrict__ start, uint32_t* __restrict__ cursor,
                                           uint32_t* __restrict__ tx, int32_t* __restrict__ tc, uint32_t* __restrict__ tds)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp; c < ncells; c += nwarps) {
        for (int64_t e = p[c] + lane; e < p[c + 1]; e += 32) {
            const int32_t g = gi[e];
            const int64_t pos = start[g] + (int64_t)atomicAdd(&cursor[g], 1u);
            tx[pos] = x[e];
            tc[pos] = (int32_t)c;
            if (tds) tds[pos] = dsx[e];
        }
    }
}
constexpr int GS_THREADS = 256;

__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v, unsigned long long* sm /* [8] */)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long t = 0;
#pragma unroll
    for (int q = 0; q < GS_THREADS / 32; ++q) t += sm[q];
    return t;
}
__device__ __forceinline__ u128 block_sum_u128(u128 v, unsigned long long* sm /* [16] */)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long lo = __shfl_xor_sync(0xffffffffu, (unsigned long long)v, o);
        const unsigned long long hi = __shfl_xor_sync(0xffffffffu, (unsigned long long)(v >> 64), o);
        v += ((u128)hi << 64) | lo;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { sm[2 * (threadIdx.x >> 5)] = (unsigned long long)v; sm[2 * (threadIdx.x >> 5) + 1] = (unsigned long long)(v >> 64); }
    __syncthreads();
    u128 t = 0;
#pragma unroll
    for (int q = 0; q < GS_THREADS / 32; ++q) t += ((u128)sm[2 * q + 1] << 64) | sm[2 * q];
    return t;
}
__device__ __forceinline__ double block_sum_f64(double v, double* sm /* [8] */)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < GS_THREADS / 32; ++q) t += sm[q];
    return t;
}
__device__ __forceinline__ void top3_push(uint32_t t[3], uint32_t v)
{
    if (v > t[0]) { t[2] = t[1]; t[1] = t[0]; t[0] = v; }
    else if (v > t[1]) { t[2] = t[1]; t[1] = v; }
    else if (v > t[2]) t[2] = v;
}

// one CTA per gene (grid-stride).  niche_r = round(n * niche_quantile) (host, R rounding).
__global__ void __launch_bounds__(GS_THREADS)
gene_stats_kernel(int32_t ngenes, const int64_t* __restrict__ start, const uint32_t* __restrict__ tx,
                  const int32_t* __restrict__ tc, const uint32_t* __restrict__ tds, const CellInfo* __restrict__ cinfo,
                  long long niche_r,
                  GeneRaw* __restrict__ out)
{
    __shared__ unsigned long long sm64[16];
    __shared__ double smd[8];
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_top[GS_THREADS / 32][3];
    __shared__ unsigned long long s_prefix;
    __shared__ long long s_k;
    for (int g = blockIdx.x; g < ngenes; g += gridDim.x) {
        const int64_t s = start[g], e = start[g + 1];
        unsigned long long tot = 0, sx_s = 0, dtot = 0;
        u128 sumsq = 0, sxx = 0, sxt = 0, dsumsq = 0;
        double nm = 0.0;
        unsigned long long on = 0, d_on = 0;
        uint32_t vmax = 0;
        uint32_t t3[3] = {0u, 0u, 0u};
        for (int64_t q = s + threadIdx.x; q < e; q += GS_THREADS) {
            const unsigned long long v = tx[q];
            const int32_t c = tc[q];
            tot += v; sumsq += (u128)(v * v); on += v > 0;
            const CellInfo ci = cinfo[c];                // one 16-byte gather: K_std_n / total and total (negated if outside subs)
            nm += (double)v * ci.scale;
            vmax = max(vmax, (uint32_t)v);
            if (ci.tot >= 0) { sx_s += v; sxx += (u128)(v * v); sxt += (u128)v * (u128)(unsigned long long)ci.tot; }
            if (tds) {
                const unsigned long long d = tds[q];
                dtot += d; dsumsq += (u128)(d * d); d_on += d > 0;
                top3_push(t3, (uint32_t)d);
            }
        }
        GeneRaw r;
        r.tot = block_sum_u64(tot, sm64);
        { const u128 t = block_sum_u128(sumsq, sm64); r.sumsq_lo = (unsigned long long)t; r.sumsq_hi = (unsigned long long)(t >> 64); }
        r.on = (uint32_t)block_sum_u64(on, sm64);
        r.nmean_sum = block_sum_f64(nm, smd);
        r.sx_s = block_sum_u64(sx_s, sm64);
        { const u128 t = block_sum_u128(sxx, sm64); r.sxx_lo = (unsigned long long)t; r.sxx_hi = (unsigned long long)(t >> 64); }
        { const u128 t = block_sum_u128(sxt, sm64); r.sxt_lo = (unsigned long long)t; r.sxt_hi = (unsigned long long)(t >> 64); }
        r.dtot = block_sum_u64(dtot, sm64);
        { const u128 t = block_sum_u128(dsumsq, sm64); r.dsumsq_lo = (unsigned long long)t; r.dsumsq_hi = (unsigned long long)(t >> 64); }
        r.d_on = (uint32_t)block_sum_u64(d_on, sm64);
        // top-3 downsampled values: warp merge, then warp leaders through shared memory
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint32_t a = __shfl_xor_sync(0xffffffffu, t3[0], o), b = __shfl_xor_sync(0xffffffffu, t3[1], o),
                           c3 = __shfl_xor_sync(0xffffffffu, t3[2], o);
            top3_push(t3, a); top3_push(t3, b); top3_push(t3, c3);
        }
        __syncthreads();
        if ((threadIdx.x & 31) == 0) { s_top[threadIdx.x >> 5][0] = t3[0]; s_top[threadIdx.x >> 5][1] = t3[1]; s_top[threadIdx.x >> 5][2] = t3[2]; }
        __syncthreads();
        r.top[0] = r.top[1] = r.top[2] = 0u;
        for (int q = 0; q < GS_THREADS / 32; ++q) { top3_push(r.top, s_top[q][0]); top3_push(r.top, s_top[q][1]); top3_push(r.top, s_top[q][2]); }
        // niche: sum of the niche_r largest values (zeros of the other cells included, so at most all of tot)
        {
            uint32_t m = vmax;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
            __syncthreads();
            if ((threadIdx.x & 31) == 0) sm64[threadIdx.x >> 5] = m;
            __syncthreads();
            for (int q = 0; q < GS_THREADS / 32; ++q) m = max(m, (uint32_t)sm64[q]);
            vmax = m;
        }
        const long long L = (long long)(e - s);
        unsigned long long up;
        if (niche_r <= 0) up = 0;
        else if (niche_r >= (long long)r.on) up = r.tot;
        else {
            // radix select of the niche_r-th largest = (L - niche_r)-th smallest (0-based) 32-bit value
            int top_byte = 3;
            while (top_byte > 0 && (vmax >> (8 * top_byte)) == 0) --top_byte;
            if (threadIdx.x == 0) { s_prefix = 0ull; s_k = L - niche_r; }
            __syncthreads();
            for (int b = top_byte; b >= 0; --b) {
                const int shift = 8 * b;
                const uint32_t decided = b == 3 ? 0u : (0xffffffffu << (shift + 8));
                const uint32_t prefix = (uint32_t)s_prefix;
                hist[threadIdx.x] = 0;
                __syncthreads();
                for (int64_t q = s + threadIdx.x; q < e; q += GS_THREADS) {
                    const uint32_t v = tx[q];
                    if ((v & decided) == prefix) atomicAdd(&hist[(v >> shift) & 255u], 1u);
                }
                __syncthreads();
                if (threadIdx.x < 32) {
                    int dg; long long cum;
                    const long long k = s_k;
                    warp_find_bucket<8>(hist, k, dg, cum);
                    __syncwarp();
                    if (threadIdx.x == 0) { s_prefix = (unsigned long long)(prefix | ((uint32_t)dg << shift)); s_k = k - cum; }
                }
                __syncthreads();
            }
            const uint32_t thr = (uint32_t)s_prefix;          // the niche_r-th largest value
            unsigned long long gsum = 0, gcnt = 0;
            for (int64_t q = s + threadIdx.x; q < e; q += GS_THREADS) {
                const uint32_t v = tx[q];
                if (v > thr) { gsum += v; ++gcnt; }
            }
            gsum = block_sum_u64(gsum, sm64);
            gcnt = block_sum_u64(gcnt, sm64);
            up = gsum + ((unsigned long long)niche_r - gcnt) * (unsigned long long)thr;
        }
        r.up = up;
        r.pad = 0;
        if (threadIdx.x == 0) out[g] = r;
        __syncthreads();
    }
}

void launch_gene_bucket(mcb_ctx* ctx, int32_t ngenes, int64_t ncells, int64_t nnz, const int64_t* p, const int32_t* gi,
                        const uint32_t* x, const uint32_t* dsx, uint32_t* cnt, int64_t* start, uint32_t* tx, int32_t* tc,
                        uint32_t* tds)
{
    const size_t smem = (size_t)ngenes * sizeof(uint32_t);
    const bool priv = smem <= 200 * 1024;              // per-CTA counters fit in shared memory
    if (priv) {
        static SmemOptIn optin;
        ensure_dyn_smem(optin, ctx, gene_count_kernel, smem);
    }
    const int per_sm = priv ? (int)std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / std::max<size_t>(smem, 1))) : 16;
    MCB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(uint32_t) * (size_t)std::max(ngenes, 1), ctx->stream));
    if (nnz > 0) {
        if (priv) {
            const int grid = (int)std::min<int64_t>(ceil_div(nnz, 4096), (int64_t)ctx->num_sms * per_sm);
            gene_count_kernel<<<grid, 512, smem, ctx->stream>>>(ngenes, nnz, gi, cnt);
        } else {
            const int grid = (int)std::min<int64_t>(ceil_div(nnz, 256), (int64_t)ctx->num_sms * 16);
            gene_count_global_kernel<<<grid, 256, 0, ctx->stream>>>(nnz, gi, cnt);
        }
        MCB_LAUNCH_CHECK();
    }
    gene_scan_kernel<<<1, 1024, 0, ctx->stream>>>(ngenes, cnt, start);
    MCB_LAUNCH_CHECK();
    MCB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(uint32_t) * (size_t)std::max(ngenes, 1), ctx->stream));    // reused as cursor
    if (nnz > 0) {
        const int grid = (int)std::min<int64_t>(ceil_div(ncells, 8), (int64_t)ctx->num_sms * 16);
        gene_scatter_kernel<<<grid, 256, 0, ctx->stream>>>(ncells, p, gi, x, dsx, start, cnt, tx, tc, tds);
        MCB_LAUNCH_CHECK();
    }
}

void launch_gene_stats(mcb_ctx* ctx, int32_t ngenes, const int64_t* start, const uint32_t* tx, const int32_t* tc,
                       const uint32_t* tds, const CellInfo* cinfo, int64_t niche_r, void* out_raw)
{
    if (ngenes == 0) return;
    const int grid = std::min(ngenes, ctx->num_sms * 8);
    gene_stats_kernel<<<grid, GS_THREADS, 0, ctx->stream>>>(ngenes, start, tx, tc, tds, cinfo, (long long)niche_r,
                                                            (GeneRaw*)out_raw);
    MCB_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------
// Per-metacell gene summaries: mc_compute_fp / mc_compute_e_gc / mc_compute_cov_gc (reference R/mc.r:164-241,
// SURVEY 8f rank 3).  All three are tgs_matrix_tapply(us, mc@mc, f) with f = exp(mean(log(1+y)))-1 or mean(x>0):
// per (gene, metacell) a sum of log(1+u) over the metacell's cells and a count of non-zero cells.  One pass over
// the cell-major CSC; each entry adds to a dense [n_mc][ngenes] accumulator pair.  log(1+u) is accumulated as a
// 64-bit fixed-point integer (2^-40 units, rounding error <= 2^-41 per entry) so the sums do not depend on the
// order the atomics land in: the result is deterministic, and within 1e-12 of the fp64 restatement for any
// realistic metacell size.  HBM/L2-atomic bound: 12 B read + two atomics per stored entry.
// ---------------------------------------------------------------------------------------------
constexpr double MC_LOG_FIX = 1099511627776.0;      // 2^40

__global__ void mc_tapply_kernel(int64_t ncells, int32_t ngenes, const int64_t* __restrict__ p, const int32_t* __restrict__ gi,
                                 const uint32_t* __restrict__ x, const int32_t* __restrict__ mc_of_cell,
                                 unsigned long long* __restrict__ logsum, uint32_t* __restrict__ on,
                                 unsigned long long* __restrict__ mc_tot)
{
    __shared__ long long lut[1024];                 // round(log(1+u) * 2^40) for the small counts that make up almost all entries
    for (int u = threadIdx.x; u < 1024; u += blockDim.x) lut[u] = __double2ll_rn(log1p((double)u) * MC_LOG_FIX);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp; c < ncells; c += nwarps) {
        const int32_t k = mc_of_cell[c];
        if (k < 0) continue;
        unsigned long long tot = 0;
        for (int64_t e = p[c] + lane; e < p[c + 1]; e += 32) {
            const uint32_t v = x[e];
            tot += v;
            if (v == 0) continue;
            const long long f = v < 1024u ? lut[v] : __double2ll_rn(log1p((double)v) * MC_LOG_FIX);
            const size_t a = (size_t)k * ngenes + gi[e];
            atomicAdd(&logsum[a], (unsigned long long)f);
            atomicAdd(&on[a], 1u);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (lane == 0) atomicAdd(&mc_tot[k], tot);
    }
}

// geomean[k][g] = exp(logsum / (2^40 * size_k)) - 1, frac_on[k][g] = on / size_k
__global__ void mc_tapply_finish_kernel(int64_t n, int32_t ngenes, const unsigned long long* __restrict__ logsum,
                                        const uint32_t* __restrict__ on, const int64_t* __restrict__ mc_size,
                                        double* __restrict__ geomean, double* __restrict__ frac_on)
{
    for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < n; a += (int64_t)gridDim.x * blockDim.x) {
        const double sz = (double)mc_size[a / ngenes];
        geomean[a] = sz > 0 ? expm1((double)logsum[a] / MC_LOG_FIX / sz) : nan("");
        frac_on[a] = sz > 0 ? (double)on[a] / sz : nan("");
    }
}

void launch_mc_tapply(mcb_ctx* ctx, int64_t ncells, int32_t ngenes, const int64_t* p, const int32_t* gi, const uint32_t* x,
                      const int32_t* mc_of_cell, int32_t n_mc, const int64_t* mc_size, unsigned long long* logsum, uint32_t* on,
                      unsigned long long* mc_tot, double* geomean, double* frac_on)
{
    const int64_t n = (int64_t)n_mc * ngenes;
    if (n == 0) return;
    MCB_CUDA(cudaMemsetAsync(logsum, 0, sizeof(unsigned long long) * (size_t)n, ctx->stream));
    MCB_CUDA(cudaMemsetAsync(on, 0, sizeof(uint32_t) * (size_t)n, ctx->stream));
    MCB_CUDA(cudaMemsetAsync(mc_tot, 0, sizeof(unsigned long long) * (size_t)n_mc, ctx->stream));
    if (ncells > 0) {
        const int grid = (int)std::min<int64_t>(ceil_div(ncells, 8), (int64_t)ctx->num_sms * 8);
        mc_tapply_kernel<<<grid, 256, 0, ctx->stream>>>(ncells, ngenes, p, gi, x, mc_of_cell, logsum, on, mc_tot);
        MCB_LAUNCH_CHECK();
    }
    const int grid = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)ctx->num_sms * 16);
    mc_tapply_finish_kernel<<<grid, 256, 0, ctx->stream>>>(n, ngenes, logsum, on, mc_size, geomean, frac_on);
    MCB_LAUNCH_CHECK();
}

}  // namespace mcb
