// select.cu -- per-row selection kernels around the correlation contraction of tgs_cor_knn
// (reference call site R/utils.r:119: the K*k_expand best-correlated cells of every cell, ranked
// 1..Kc with the cell itself at rank 1, R/atlas.r:475-476).
//
//   sample_threshold : r-th largest of each row of the sampled correlation block  -> filter threshold
//   topk_rows        : sort one row's filtered candidates, emit the ranked kNN list
//   topk_big         : exact selection over a full correlation row (rows whose window was missed)
//
// All three are shared-memory radix-select / bitonic-sort kernels, one CTA per row; their HBM
// traffic is one read of the candidates and one write of the list.
#include "internal.h"
#include <cuda_fp16.h>

namespace mcb {

constexpr int SEL_THREADS = 256;

// ---------------------------------------------------------------------------------------------
// thr[row] = (r-th largest of C[row][0..S)) - eps ; -2 if r > S, from a 4096-bin linear histogram of the
// correlations over [-1, 1] (well spread, so the shared-memory atomics do not pile onto a few exponent bins)
// ---------------------------------------------------------------------------------------------
constexpr int THR_BINS = 4096;

__device__ __forceinline__ int lin_bin(float c)
{
    int b = (int)((c + 1.0f) * (THR_BINS / 2));
    return b < 0 ? 0 : (b >= THR_BINS ? THR_BINS - 1 : b);
}

// The filter threshold is a statistical estimate with slack (api.cu::sample_rule), so the selection does not have
// to be exact: thr = the r-th largest sampled correlation located to within a fraction of a 1/2048-wide histogram bin
// (linear interpolation inside the bin; the rank error is far below the rule's 4 sigma slack, and rows that still
// miss their window are redone exactly).  One streaming pass and 16 KB of shared memory
// (8 CTAs per SM), and replaced an exact two-level select that kept the row's keys in shared memory (4.8 ms at C2).
// T = float (CUDA-core validation path) or __half (tensor-core path, STORE16).  Instruction-issue bound at ~0.9
// instructions per sampled value (ncu: 11 K warp instructions per row of 12,800 values, 51 % issue-active); the
// shared-memory atomics compile to ATOMS.POPC.INC, which aggregates same-bin lanes, and are not the limiter: three
// variants that pre-filtered the values against mean + sd before binning all ran slower than binning everything.
template <typename T>
__global__ void __launch_bounds__(SEL_THREADS, 8)
sample_threshold_kernel(const T* __restrict__ C, int64_t ldc, int64_t rows, int32_t S, int32_t r, float eps,
                        float* __restrict__ thr)
{
    __shared__ uint32_t bins[THR_BINS];
    __shared__ uint32_t part[SEL_THREADS];
    __shared__ uint32_t chunk[32];
    constexpr int BPT = THR_BINS / SEL_THREADS;     // bins per thread
    static_assert(BPT <= 32 && (BPT & (BPT - 1)) == 0, "one lane per bin of a chunk; power of two for the rotated read");
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        if (r > S) { if (threadIdx.x == 0) thr[row] = -2.0f; continue; }
        const T* c = C + row * ldc;
#pragma unroll
        for (int q = 0; q < BPT; ++q) bins[q * SEL_THREADS + threadIdx.x] = 0;
        __syncthreads();
        if constexpr (sizeof(T) == 2) {
            // rows are 16-byte aligned (ldc % 8 == 0): 8 values per load, the loads of a batch issued together
            const uint4* c8 = reinterpret_cast<const uint4*>(c);
            const int nv = S / 8;
            constexpr int NLD = 4;
            for (int j0 = 0; j0 < nv; j0 += NLD * SEL_THREADS) {
                uint4 v[NLD];
#pragma unroll
                for (int q = 0; q < NLD; ++q) {
                    const int j = j0 + q * SEL_THREADS + threadIdx.x;
                    v[q] = j < nv ? c8[j] : make_uint4(0u, 0u, 0u, 0u);
                }
#pragma unroll
                for (int q = 0; q < NLD; ++q) {
                    if (j0 + q * SEL_THREADS + threadIdx.x >= nv) continue;
                    const uint32_t w[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 p = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
                        atomicAdd(&bins[lin_bin(p.x)], 1u);
                        atomicAdd(&bins[lin_bin(p.y)], 1u);
                    }
                }
            }
            for (int j = nv * 8 + threadIdx.x; j < S; j += SEL_THREADS) atomicAdd(&bins[lin_bin(__half2float(c[j]))], 1u);
        } else {
            for (int j = threadIdx.x; j < S; j += SEL_THREADS) atomicAdd(&bins[lin_bin((float)c[j])], 1u);
        }
        __syncthreads();
        // part[t] = count of the thread's BPT consecutive bins; the read order is rotated per thread so that a warp
        // touches 16 banks at a time instead of 2 (the straight loop was a 16-way bank conflict)
        uint32_t mine = 0;
#pragma unroll
        for (int q = 0; q < BPT; ++q) mine += bins[threadIdx.x * BPT + ((q + threadIdx.x) & (BPT - 1))];
        part[threadIdx.x] = mine;
        __syncthreads();
        if (threadIdx.x < 32) {
            // the r-th largest is the (S - r)-th smallest (0-based): chunk of BPT bins, then the bin inside it
            int t, b; long long before_t, before_b;
            warp_find_bucket<SEL_THREADS / 32>(part, (long long)(S - r), t, before_t);
            chunk[threadIdx.x] = threadIdx.x < BPT ? bins[t * BPT + threadIdx.x] : 0u;
            __syncwarp();
            warp_find_bucket<1>(chunk, (long long)(S - r) - before_t, b, before_b);
            if (threadIdx.x == 0) {
                // linear interpolation inside the bin: the value at ascending position `pos` of its `cnt` members
                const float pos = (float)((long long)(S - r) - before_t - before_b);
                const float cnt = (float)chunk[b];
                const float frac = cnt > 0.f ? (pos + 0.5f) / cnt : 0.f;
                thr[row] = ((float)(t * BPT + b) + frac) * (2.0f / THR_BINS) - 1.0f - eps;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// candidates of one row -> ranked kNN list.  cand keys are (cor desc, col asc) ascending u64.
// fail_flag[row] = 1 when the row's candidate count fell outside [kc_out-1, cap].
// ---------------------------------------------------------------------------------------------
template <int IPT>      // sorts up to 256 * IPT candidates per row
__global__ void __launch_bounds__(SEL_THREADS)
topk_rows_kernel(const uint64_t* __restrict__ cand, const uint32_t* __restrict__ cnt, int32_t cap, int64_t row0,
                 int64_t rows, int32_t kc_out, int32_t Kc, int32_t* __restrict__ knn_col, float* __restrict__ knn_cor,
                 int32_t* __restrict__ fail_flag)
{
    extern __shared__ uint64_t sk[];        // [256 * IPT]
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const uint32_t n = cnt[row];
        int32_t* oc = knn_col + row * (int64_t)Kc;
        float* ov = knn_cor + row * (int64_t)Kc;
        if (n > (uint32_t)cap || (int64_t)n < (int64_t)kc_out - 1) {
            if (threadIdx.x == 0) fail_flag[row] = 1;
            continue;
        }
        if (threadIdx.x == 0) fail_flag[row] = 0;
        const uint64_t* src = cand + row * (int64_t)cap;
        uint64_t key[IPT];
#pragma unroll
        for (int r = 0; r < IPT; ++r) {
            const int e = threadIdx.x * IPT + r;
            key[r] = e < (int)n ? src[e] : ~0ull;
        }
        bitonic_sort_regs_u64<IPT>(key, sk);
        __syncthreads();
#pragma unroll
        for (int r = 0; r < IPT; ++r) sk[threadIdx.x * IPT + r] = key[r];
        __syncthreads();
        for (int q = threadIdx.x; q < Kc; q += SEL_THREADS) {
            if (q == 0) { oc[0] = (int32_t)(row0 + row); ov[0] = 1.0f; }
            else if (q < kc_out) { const uint64_t k = sk[q - 1]; oc[q] = (int32_t)cand_key_col(k); ov[q] = cand_key_cor(k); }
            else { oc[q] = -1; ov[q] = 0.0f; }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// exact selection over a full correlation row held in global memory (fp32, M values):
// the kc_out-1 best columns != self by (cor desc, col asc).  One CTA per listed row.
// ---------------------------------------------------------------------------------------------
template <bool WITH_SELF>    // WITH_SELF: x == y, the row's own column is forced to rank 1 and skipped in the scan
__global__ void __launch_bounds__(SEL_THREADS)
topk_big_kernel(const float* __restrict__ C, int64_t ldc, int64_t M, const int32_t* __restrict__ rows, int32_t nrows,
                int64_t row0, int32_t kc_out, int32_t Kc, const float* __restrict__ thr_rows, int32_t np2s,
                int32_t* __restrict__ knn_col, float* __restrict__ knn_cor)
{
    extern __shared__ uint64_t sk[];        // [np2s >= pow2(kc_out)]
    __shared__ uint32_t hist[256];
    __shared__ unsigned long long s_prefix, s_thr;
    __shared__ long long s_k;
    __shared__ int s_done, s_fill;
    for (int q = blockIdx.x; q < nrows; q += gridDim.x) {
        const int64_t lrow = rows[q];               // local row (relative to row0)
        const int64_t self = WITH_SELF ? row0 + lrow : -1;
        const float* c = C + (int64_t)q * ldc;
        const int want = WITH_SELF ? kc_out - 1 : kc_out;
        int32_t* oc = knn_col + lrow * (int64_t)Kc;
        float* ov = knn_cor + lrow * (int64_t)Kc;
        if (want <= 0) {
            for (int t = threadIdx.x; t < Kc; t += SEL_THREADS) { oc[t] = (WITH_SELF && t == 0) ? (int32_t)self : -1; ov[t] = (WITH_SELF && t == 0) ? 1.f : 0.f; }
            continue;
        }
        if (threadIdx.x == 0) { s_prefix = 0ull; s_k = want - 1; s_done = 0; s_thr = ~0ull; s_fill = 0; }
        __syncthreads();
        // common case (the row overflowed its candidate capacity): everything wanted lies above the row's filter
        // threshold, and only a few thousand values do -> collect them in one pass and sort in shared memory
        if (thr_rows) {
            const float t = thr_rows[lrow];
            for (int64_t j0 = 0; j0 < M; j0 += SEL_THREADS) {
                const int64_t j = j0 + threadIdx.x;
                const bool in = j < M && j != self && c[j] >= t;
                if (in) { const int pos = atomicAdd(&s_fill, 1); if (pos < np2s) sk[pos] = cand_key(c[j], (uint32_t)j); }
            }
            __syncthreads();
            const int got = s_fill;
            __syncthreads();
            if (got >= want && got <= np2s) {
                const int np2 = next_pow2(got < 2 ? 2 : got);
                for (int tq = got + threadIdx.x; tq < np2; tq += SEL_THREADS) sk[tq] = ~0ull;
                __syncthreads();
                bitonic_sort_u64<SEL_THREADS>(sk, np2);
                for (int tq = threadIdx.x; tq < Kc; tq += SEL_THREADS) {
                    if (WITH_SELF && tq == 0) { oc[0] = (int32_t)self; ov[0] = 1.0f; }
                    else if (tq < kc_out) { const uint64_t k = sk[tq - (WITH_SELF ? 1 : 0)]; oc[tq] = (int32_t)cand_key_col(k); ov[tq] = cand_key_cor(k); }
                    else { oc[tq] = -1; ov[tq] = 0.0f; }
                }
                __syncthreads();
                continue;
            }
            if (threadIdx.x == 0) s_fill = 0;
            __syncthreads();
        }
        for (int pass = 0; pass < 8; ++pass) {
            const int shift = 56 - 8 * pass;
            const unsigned long long decided = pass == 0 ? 0ull : (~0ull << (shift + 8));
            const unsigned long long prefix = s_prefix;
            hist[threadIdx.x] = 0;
            __syncthreads();
            for (int64_t j = threadIdx.x; j < M; j += SEL_THREADS) {
                if (j == self) continue;
                const uint64_t key = cand_key(c[j], (uint32_t)j);
                if ((key & decided) == prefix) atomicAdd(&hist[(uint32_t)(key >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (threadIdx.x < 32) {
                int dg; long long cum;
                const long long k = s_k;
                warp_find_bucket<8>(hist, k, dg, cum);
                __syncwarp();
                if (threadIdx.x == 0) {
                    const unsigned long long np = prefix | ((unsigned long long)dg << shift);
                    s_prefix = np; s_k = k - cum;
                    if (k - cum == (long long)hist[dg] - 1 || pass == 7) {
                        s_thr = np | (shift ? ((1ull << shift) - 1ull) : 0ull);
                        s_done = 1;
                    }
                }
            }
            __syncthreads();
            if (s_done) break;
        }
        const uint64_t thr = s_thr;
        const int np2 = next_pow2(want < 2 ? 2 : want);
        for (int t = threadIdx.x; t < np2; t += SEL_THREADS) sk[t] = ~0ull;
        __syncthreads();
        for (int64_t j = threadIdx.x; j < M; j += SEL_THREADS) {
            if (j == self) continue;
            const uint64_t key = cand_key(c[j], (uint32_t)j);
            if (key <= thr) { const int pos = atomicAdd(&s_fill, 1); if (pos < np2) sk[pos] = key; }
        }
        __syncthreads();
        bitonic_sort_u64<SEL_THREADS>(sk, np2);
        for (int t = threadIdx.x; t < Kc; t += SEL_THREADS) {
            if (WITH_SELF && t == 0) { oc[0] = (int32_t)self; ov[0] = 1.0f; }
            else if (t < kc_out) { const uint64_t k = sk[t - (WITH_SELF ? 1 : 0)]; oc[t] = (int32_t)cand_key_col(k); ov[t] = cand_key_cor(k); }
            else { oc[t] = -1; ov[t] = 0.0f; }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// records of the filter epilogue -> per-row candidate lists.  One CTA per pool chunk; the returning
// atomics that hand out list positions are independent across threads, so their latency is hidden by
// parallelism here instead of stalling the GEMM epilogue.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
scatter_candidates_kernel(const uint4* __restrict__ records, const uint32_t* __restrict__ cursor,
                          const uint32_t* __restrict__ chunk_fill, uint32_t pool_chunks, uint32_t chunk_records,
                          uint64_t* __restrict__ cand, uint32_t* __restrict__ cnt, int32_t cap)
{
    const uint32_t used = min(*cursor, pool_chunks);
    for (uint32_t ch = blockIdx.x; ch < used; ch += gridDim.x) {
        const uint32_t n = min(chunk_fill[ch], chunk_records);
        const uint4* r = records + (size_t)ch * chunk_records;
        // four records per thread in flight: the load -> returning L2 atomic -> store chain is pure latency
        // (ncu: 253 long-scoreboard stall cycles per issued instruction with one record per thread)
        for (uint32_t e0 = 0; e0 < n; e0 += 4 * blockDim.x) {
            uint4 rec[4]; uint32_t pos[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t e = e0 + q * blockDim.x + threadIdx.x;
                rec[q] = e < n ? r[e] : make_uint4(0u, 0u, 0xffffffffu, 0u);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) pos[q] = rec[q].z != 0xffffffffu ? atomicAdd(&cnt[rec[q].z], 1u) : 0xffffffffu;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (pos[q] < (uint32_t)cap) cand[(size_t)rec[q].z * cap + pos[q]] = ((uint64_t)rec[q].y << 32) | rec[q].x;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// symmetric multi-GPU: a rank's filter pass yields records for rows owned by every rank.  Records are
// grouped by owner (row / rows_per_rank) into one send buffer (NCCL all-to-all by the host plumbing),
// and the received records are binned like local ones.
// ---------------------------------------------------------------------------------------------
constexpr int MAX_RANKS = 64;

__global__ void __launch_bounds__(256)
count_records_by_dest_kernel(const uint4* __restrict__ records, const uint32_t* __restrict__ cursor,
                             const uint32_t* __restrict__ chunk_fill, uint32_t pool_chunks, uint32_t chunk_records,
                             uint32_t rows_per_rank, int nranks, unsigned long long* __restrict__ counts)
{
    __shared__ uint32_t sc[MAX_RANKS];
    const uint32_t used = min(*cursor, pool_chunks);
    for (uint32_t ch = blockIdx.x; ch < used; ch += gridDim.x) {
        if (threadIdx.x < MAX_RANKS) sc[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t n = min(chunk_fill[ch], chunk_records);
        const uint4* r = records + (size_t)ch * chunk_records;
        for (uint32_t e = threadIdx.x; e < n; e += blockDim.x) atomicAdd(&sc[r[e].z / rows_per_rank], 1u);
        __syncthreads();
        if (threadIdx.x < nranks && sc[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)sc[threadIdx.x]);
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256)
partition_records_kernel(const uint4* __restrict__ records, const uint32_t* __restrict__ cursor,
                         const uint32_t* __restrict__ chunk_fill, uint32_t pool_chunks, uint32_t chunk_records,
                         uint32_t rows_per_rank, int nranks, const unsigned long long* __restrict__ counts,
                         unsigned long long* __restrict__ cur, uint4* __restrict__ send)
{
    // counts[d] = records for destination d (count_records_by_dest_kernel); a destination's segment of `send` starts at the
    // sum of the counts before it -- derived here, so the host never has to hand the prefix back
    __shared__ uint32_t sc[MAX_RANKS];
    __shared__ unsigned long long sb[MAX_RANKS], base[MAX_RANKS];
    if (threadIdx.x == 0) { unsigned long long run = 0; for (int d = 0; d < nranks; ++d) { base[d] = run; run += counts[d]; } }
    __syncthreads();
    const uint32_t used = min(*cursor, pool_chunks);
    for (uint32_t ch = blockIdx.x; ch < used; ch += gridDim.x) {
        if (threadIdx.x < MAX_RANKS) sc[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t n = min(chunk_fill[ch], chunk_records);
        const uint4* r = records + (size_t)ch * chunk_records;
        for (uint32_t e = threadIdx.x; e < n; e += blockDim.x) atomicAdd(&sc[r[e].z / rows_per_rank], 1u);
        __syncthreads();
        if (threadIdx.x < nranks) {      // reserve this chunk's range in every destination segment
            sb[threadIdx.x] = base[threadIdx.x] + (sc[threadIdx.x] ? atomicAdd(&cur[threadIdx.x], (unsigned long long)sc[threadIdx.x]) : 0ull);
            sc[threadIdx.x] = 0;
        }
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < n; e += blockDim.x) {
            const uint4 rec = r[e];
            const uint32_t d = rec.z / rows_per_rank;
            send[sb[d] + atomicAdd(&sc[d], 1u)] = rec;
        }
        __syncthreads();
    }
}
__global__ void scatter_flat_kernel(const uint4* __restrict__ records, int64_t n, uint32_t row0, uint64_t* __restrict__ cand,
                                    uint32_t* __restrict__ cnt, int32_t cap)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const uint4 rec = records[e];
        const uint32_t lrow = rec.z - row0;
        const uint32_t pos = atomicAdd(&cnt[lrow], 1u);
        if (pos < (uint32_t)cap) cand[(size_t)lrow * cap + pos] = ((uint64_t)rec.y << 32) | rec.x;
    }
}

// compact the flagged rows into a list; *n_out = count
__global__ void compact_flags_kernel(const int32_t* __restrict__ flag, int64_t n, int32_t* __restrict__ list,
                                     int32_t* __restrict__ n_out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (flag[i]) { const int pos = atomicAdd(n_out, 1); list[pos] = (int32_t)i; }
}

void launch_sample_threshold(mcb_ctx* ctx, const float* C, int64_t ldc, int64_t rows, int32_t S, int32_t r,
                             float eps, float* thr)
{
    if (rows == 0) return;
    int grid = (int)std::min<int64_t>(rows, (int64_t)ctx->num_sms * 8);
    sample_threshold_kernel<float><<<grid, SEL_THREADS, 0, ctx->stream>>>(C, ldc, rows, S, r, eps, thr);
    MCB_LAUNCH_CHECK();
}
void launch_sample_threshold16(mcb_ctx* ctx, const void* C16, int64_t ldc, int64_t rows, int32_t S, int32_t r,
                               float eps, float* thr)
{
    if (rows == 0) return;
    MCB_CHECK(ldc % 8 == 0, "sample threshold: fp16 rows must be 16-byte aligned");
    static int per_sm = 0;            // resident CTAs per SM: one full wave, the kernel strides over the rows
    if (!per_sm) MCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sample_threshold_kernel<__half>, SEL_THREADS, 0));
    int grid = (int)std::min<int64_t>(rows, (int64_t)ctx->num_sms * std::max(per_sm, 1));
    sample_threshold_kernel<__half><<<grid, SEL_THREADS, 0, ctx->stream>>>(reinterpret_cast<const __half*>(C16), ldc, rows, S, r, eps, thr);
    MCB_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------
// Same result with less sorting: the row holds ~1.5 x the wanted number of candidates, so a monotone
// 2048-bin histogram of the keys first finds the bin of the (kc_out-1)-th best; only the keys up to that
// bin (a handful more than wanted) are compacted and sorted with the smaller IPT_OUT network.  Falls back
// to sorting everything when ties make the selected set exceed 256 * IPT_OUT keys.
// ---------------------------------------------------------------------------------------------
constexpr int TOPK_BINS = 2048;

#ifndef MCB_TOPK_MIN_BLOCKS
#define MCB_TOPK_MIN_BLOCKS 6        // 40 registers: six resident CTAs per SM hide the barriers of the per-row select (1.35 -> 1.25 ms at C2)
#endif
template <int IPT_IN, int IPT_OUT>
__global__ void __launch_bounds__(SEL_THREADS, IPT_IN <= 8 ? MCB_TOPK_MIN_BLOCKS : 1)
topk_select_rows_kernel(const uint64_t* __restrict__ cand, const uint32_t* __restrict__ cnt, int32_t cap, int64_t row0,
                        int64_t rows, int32_t kc_out, int32_t Kc, int32_t* __restrict__ knn_col, float* __restrict__ knn_cor,
                        int32_t* __restrict__ fail_flag)
{
    extern __shared__ uint64_t sk[];        // [256 * IPT_IN]
    __shared__ __align__(16) uint32_t hist[TOPK_BINS];
    __shared__ __align__(16) uint32_t pfx[TOPK_BINS + 4];
    __shared__ uint32_t s_min[SEL_THREADS / 32], s_max[SEL_THREADS / 32], s_wsum[SEL_THREADS / 32];
    __shared__ int s_bin, s_total, s_n;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const uint32_t n = cnt[row];
        int32_t* oc = knn_col + row * (int64_t)Kc;
        float* ov = knn_cor + row * (int64_t)Kc;
        if (n > (uint32_t)cap || (int64_t)n < (int64_t)kc_out - 1) {
            if (threadIdx.x == 0) fail_flag[row] = 1;
            continue;
        }
        if (threadIdx.x == 0) { fail_flag[row] = 0; s_n = 0; }
        const int want = kc_out - 1;
        const uint64_t* src = cand + row * (int64_t)cap;
        uint64_t key[IPT_IN];
        uint32_t kmin = 0xffffffffu, kmax = 0u;
#pragma unroll
        for (int r = 0; r < IPT_IN; ++r) {
            const int e = r * SEL_THREADS + threadIdx.x;          // coalesced; order is irrelevant before the sort
            key[r] = e < (int)n ? src[e] : ~0ull;
            if (e < (int)n) { const uint32_t h = (uint32_t)(key[r] >> 32); kmin = min(kmin, h); kmax = max(kmax, h); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
            kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
        }
        if (lane == 0) { s_min[wid] = kmin; s_max[wid] = kmax; }
        for (int b = threadIdx.x; b < TOPK_BINS; b += SEL_THREADS) hist[b] = 0;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < SEL_THREADS / 32; ++w) { kmin = min(kmin, s_min[w]); kmax = max(kmax, s_max[w]); }
        int shift = 0;
        while (((kmax - kmin) >> shift) >= (uint32_t)TOPK_BINS) ++shift;
#pragma unroll
        for (int r = 0; r < IPT_IN; ++r)
            if (key[r] != ~0ull) atomicAdd(&hist[((uint32_t)(key[r] >> 32) - kmin) >> shift], 1u);
        __syncthreads();
        if (threadIdx.x < 32) {
            int b = 0; long long before = 0;
            if (want > 0) warp_find_bucket<TOPK_BINS / 32>(hist, (long long)(want - 1), b, before);
            if (threadIdx.x == 0) { s_bin = b; s_total = want > 0 ? (int)before + (int)hist[b] : 0; }
        }
        __syncthreads();
        const int bsel = s_bin, total = s_total;
        // exclusive prefix of the histogram (thread t owns bins [8t, 8t + 8)), and the largest selected bin
        constexpr int BPT = TOPK_BINS / SEL_THREADS;
        static_assert(BPT == 8, "two 16-byte loads per thread");
        uint32_t hv[BPT];
        {
            const uint4 a = *reinterpret_cast<const uint4*>(&hist[threadIdx.x * BPT]);
            const uint4 b = *reinterpret_cast<const uint4*>(&hist[threadIdx.x * BPT + 4]);
            hv[0] = a.x; hv[1] = a.y; hv[2] = a.z; hv[3] = a.w; hv[4] = b.x; hv[5] = b.y; hv[6] = b.z; hv[7] = b.w;
        }
        uint32_t loc = 0, mx = 0;
#pragma unroll
        for (int q = 0; q < BPT; ++q) { loc += hv[q]; if ((int)threadIdx.x * BPT + q <= bsel) mx = max(mx, hv[q]); }
        uint32_t incl = loc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += up; }
        if (lane == 31) s_wsum[wid] = incl;
        __syncthreads();
        uint32_t run = incl - loc;
#pragma unroll
        for (int w = 0; w < SEL_THREADS / 32; ++w) if (w < wid) run += s_wsum[w];
        {
            uint32_t p[BPT];
#pragma unroll
            for (int q = 0; q < BPT; ++q) { p[q] = run; run += hv[q]; }
            *reinterpret_cast<uint4*>(&pfx[threadIdx.x * BPT]) = make_uint4(p[0], p[1], p[2], p[3]);
            *reinterpret_cast<uint4*>(&pfx[threadIdx.x * BPT + 4]) = make_uint4(p[4], p[5], p[6], p[7]);
            if (threadIdx.x == SEL_THREADS - 1) pfx[TOPK_BINS] = run;
        }
        const int crowded = __syncthreads_or(mx > 32u);
        const uint64_t* res = sk;
        if (total <= 256 * IPT_OUT && !crowded) {
            // counting sort: the histogram already groups the keys by their leading bits (~0.5 keys per bin), so each
            // selected key goes to its bin's segment and is then ranked among the few keys of the same bin --
            // O(n) work instead of the n log^2 n of a sorting network (which took ~70 % of this kernel's instructions)
#pragma unroll
            for (int r = 0; r < IPT_IN; ++r) {
                if (key[r] == ~0ull) continue;
                const int bin = (int)((((uint32_t)(key[r] >> 32) - kmin) >> shift));
                if (bin <= bsel) sk[pfx[bin] + (atomicSub(&hist[bin], 1u) - 1u)] = key[r];
            }
            __syncthreads();
            uint64_t* sorted = sk + 256 * IPT_OUT;          // total <= 256 * IPT_OUT <= half of sk
            for (int p = threadIdx.x; p < total; p += SEL_THREADS) {
                const uint64_t k = sk[p];
                const int bin = (int)((((uint32_t)(k >> 32) - kmin) >> shift));
                const int a = (int)pfx[bin], b = (int)pfx[bin + 1];
                int rank = 0;
                for (int q = a; q < b; ++q) rank += sk[q] < k ? 1 : 0;
                sorted[a + rank] = k;
            }
            res = sorted;
        } else if (total <= 256 * IPT_OUT) {
            // crowded bins (many near-ties): compact the keys of the bins up to bsel, pad, sort the short list
#pragma unroll
            for (int r = 0; r < IPT_IN; ++r) {
                const bool in = key[r] != ~0ull && (int)((((uint32_t)(key[r] >> 32) - kmin) >> shift)) <= bsel;
                const unsigned m = __ballot_sync(0xffffffffu, in);
                int base = 0;
                if (lane == 0 && m) base = atomicAdd(&s_n, __popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (in) sk[base + __popc(m & ((1u << lane) - 1u))] = key[r];
            }
            __syncthreads();
            for (int q = total + threadIdx.x; q < 256 * IPT_OUT; q += SEL_THREADS) sk[q] = ~0ull;
            __syncthreads();
            uint64_t k2[IPT_OUT];
#pragma unroll
            for (int r = 0; r < IPT_OUT; ++r) k2[r] = sk[r * SEL_THREADS + threadIdx.x];
            __syncthreads();
            bitonic_sort_regs_u64<IPT_OUT>(k2, sk);
            __syncthreads();
#pragma unroll
            for (int r = 0; r < IPT_OUT; ++r) sk[threadIdx.x * IPT_OUT + r] = k2[r];
        } else {
            bitonic_sort_regs_u64<IPT_IN>(key, sk);
            __syncthreads();
#pragma unroll
            for (int r = 0; r < IPT_IN; ++r) sk[threadIdx.x * IPT_IN + r] = key[r];
        }
        __syncthreads();
        for (int q = threadIdx.x; q < Kc; q += SEL_THREADS) {
            if (q == 0) { oc[0] = (int32_t)(row0 + row); ov[0] = 1.0f; }
            else if (q < kc_out) { const uint64_t k = res[q - 1]; oc[q] = (int32_t)cand_key_col(k); ov[q] = cand_key_cor(k); }
            else { oc[q] = -1; ov[q] = 0.0f; }
        }
        __syncthreads();
    }
}

template <int IPT_IN, int IPT_OUT>
static void launch_topk_select_t(mcb_ctx* ctx, const uint64_t* cand, const uint32_t* cnt, int32_t cap, int64_t row0,
                                 int64_t rows, int32_t kc_out, int32_t Kc, int32_t* knn_col, float* knn_cor, int32_t* fail_flag)
{
    const size_t smem = (size_t)IPT_IN * 256 * sizeof(uint64_t);
    static SmemOptIn optin;   // static (histogram + prefix, 16.5 KB) + dynamic shared memory can exceed the 48 KB default
    ensure_dyn_smem(optin, ctx, topk_select_rows_kernel<IPT_IN, IPT_OUT>, smem);
    static int per_sm = 0;            // resident CTAs per SM: whole waves only, the kernel strides over the rows
    if (!per_sm) MCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, topk_select_rows_kernel<IPT_IN, IPT_OUT>, SEL_THREADS, smem));
    int grid = (int)std::min<int64_t>(rows, (int64_t)ctx->num_sms * std::max(per_sm, 1));
    topk_select_rows_kernel<IPT_IN, IPT_OUT><<<grid, SEL_THREADS, smem, ctx->stream>>>(cand, cnt, cap, row0, rows, kc_out, Kc,
                                                                                        knn_col, knn_cor, fail_flag);
    MCB_LAUNCH_CHECK();
}

template <int IPT>
static void launch_topk_rows_t(mcb_ctx* ctx, const uint64_t* cand, const uint32_t* cnt, int32_t cap, int64_t row0,
                               int64_t rows, int32_t kc_out, int32_t Kc, int32_t* knn_col, float* knn_cor, int32_t* fail_flag)
{
    const size_t smem = (size_t)IPT * 256 * sizeof(uint64_t);
    static SmemOptIn optin;
    ensure_dyn_smem(optin, ctx, topk_rows_kernel<IPT>, smem);
    int grid = (int)std::min<int64_t>(rows, (int64_t)ctx->num_sms * 8);
    topk_rows_kernel<IPT><<<grid, SEL_THREADS, smem, ctx->stream>>>(cand, cnt, cap, row0, rows, kc_out, Kc, knn_col,
                                                                     knn_cor, fail_flag);
    MCB_LAUNCH_CHECK();
}

void launch_topk_rows(mcb_ctx* ctx, const uint64_t* cand, const uint32_t* cnt, int32_t cap, int64_t row0,
                      int64_t rows, int32_t kc_out, int32_t Kc, int32_t* knn_col, float* knn_cor, int32_t* fail_flag)
{
    if (rows == 0) return;
    MCB_CHECK(cap <= 4096, "MC-ERR: K * k_expand is too large relative to the number of cells for this build");
    // select-then-sort when the wanted list fits a sorting network half (or less) the size of the candidate capacity
    const int want = kc_out - 1 + 16;       // a little room for the boundary bin
    if (cap > 1024 && cap <= 2048 && want <= 1024) { launch_topk_select_t<8, 4>(ctx, cand, cnt, cap, row0, rows, kc_out, Kc, knn_col, knn_cor, fail_flag); return; }
    if (cap > 2048 && want <= 1024) { launch_topk_select_t<16, 4>(ctx, cand, cnt, cap, row0, rows, kc_out, Kc, knn_col, knn_cor, fail_flag); return; }
    if (cap > 2048 && want <= 2048) { launch_topk_select_t<16, 8>(ctx, cand, cnt, cap, row0, rows, kc_out, Kc, knn_col, knn_cor, fail_flag); return; }
    if (cap <= 512) launch_topk_rows_t<2>(ctx, cand, cnt, cap, row0, rows, kc_out, Kc, knn_col, knn_cor, fail_flag);
    else if (cap <= 1024) launch_topk_rows_t<4>(ctx, cand, cnt, cap, row0, rows, kc_out, Kc, knn_col, knn_cor, fail_flag);
    else if (cap <= 2048) launch_topk_rows_t<8>(ctx, cand, cnt, cap, row0, rows, kc_out, Kc, knn_col, knn_cor, fail_flag);
    else launch_topk_rows_t<16>(ctx, cand, cnt, cap, row0, rows, kc_out, Kc, knn_col, knn_cor, fail_flag);
}

void launch_scatter_candidates(mcb_ctx* ctx, const CandPool& pool, uint64_t* cand, uint32_t* cnt, int32_t cap)
{
    int grid = (int)std::min<int64_t>(pool.n_chunks, (int64_t)ctx->num_sms * 8);
    scatter_candidates_kernel<<<grid, 256, 0, ctx->stream>>>((const uint4*)pool.records, pool.cursor, pool.chunk_fill,
                                                             pool.n_chunks, tc_pool_chunk_records(), cand, cnt, cap);
    MCB_LAUNCH_CHECK();
}

void launch_count_records_by_dest(mcb_ctx* ctx, const CandPool& pool, int64_t rows_per_rank, int32_t nranks, uint64_t* counts)
{
    MCB_CHECK(nranks <= MAX_RANKS, "too many ranks");
    int grid = (int)std::min<int64_t>(pool.n_chunks, (int64_t)ctx->num_sms * 8);
    count_records_by_dest_kernel<<<grid, 256, 0, ctx->stream>>>((const uint4*)pool.records, pool.cursor, pool.chunk_fill, pool.n_chunks,
                                                                tc_pool_chunk_records(), (uint32_t)rows_per_rank, nranks,
                                                                (unsigned long long*)counts);
    MCB_LAUNCH_CHECK();
}
void launch_partition_records(mcb_ctx* ctx, const CandPool& pool, int64_t rows_per_rank, int32_t nranks, const uint64_t* counts,
                              uint64_t* cursor, void* send)
{
    int grid = (int)std::min<int64_t>(pool.n_chunks, (int64_t)ctx->num_sms * 8);
    partition_records_kernel<<<grid, 256, 0, ctx->stream>>>((const uint4*)pool.records, pool.cursor, pool.chunk_fill, pool.n_chunks,
                                                            tc_pool_chunk_records(), (uint32_t)rows_per_rank, nranks,
                                                            (const unsigned long long*)counts, (unsigned long long*)cursor, (uint4*)send);
    MCB_LAUNCH_CHECK();
}
void launch_scatter_flat(mcb_ctx* ctx, const void* records, int64_t n, int64_t row0, uint64_t* cand, uint32_t* cnt, int32_t cap)
{
    if (n == 0) return;
    int grid = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)ctx->num_sms * 16);
    scatter_flat_kernel<<<grid, 256, 0, ctx->stream>>>((const uint4*)records, n, (uint32_t)row0, cand, cnt, cap);
    MCB_LAUNCH_CHECK();
}

__global__ void sum_u32_kernel(const uint32_t* __restrict__ v, int64_t n, unsigned long long* __restrict__ out)
{
    unsigned long long s = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}
__global__ void fill_u32_kernel(uint32_t* __restrict__ v, int64_t n, uint32_t value)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] = value;
}
// *out (pre-zeroed by this launcher) = sum of v[0..n)
void launch_sum_u32(mcb_ctx* ctx, const uint32_t* v, int64_t n, uint64_t* out)
{
    MCB_CUDA(cudaMemsetAsync(out, 0, sizeof(uint64_t), ctx->stream));
    if (n == 0) return;
    int grid = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)ctx->num_sms * 4);
    sum_u32_kernel<<<grid, 256, 0, ctx->stream>>>(v, n, (unsigned long long*)out);
    MCB_LAUNCH_CHECK();
}
void launch_fill_u32(mcb_ctx* ctx, void* v, int64_t n, uint32_t value)
{
    if (n == 0) return;
    int grid = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)ctx->num_sms * 4);
    fill_u32_kernel<<<grid, 256, 0, ctx->stream>>>((uint32_t*)v, n, value);
    MCB_LAUNCH_CHECK();
}

void launch_compact_flags(mcb_ctx* ctx, const int32_t* flag, int64_t n, int32_t* list, int32_t* n_out)
{
    MCB_CUDA(cudaMemsetAsync(n_out, 0, sizeof(int32_t), ctx->stream));
    if (n == 0) return;
    int grid = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)ctx->num_sms * 8);
    compact_flags_kernel<<<grid, 256, 0, ctx->stream>>>(flag, n, list, n_out);
    MCB_LAUNCH_CHECK();
}

void launch_topk_big(mcb_ctx* ctx, const float* C, int64_t ldc, int64_t M, const int32_t* rows, int32_t nrows,
                     int64_t row0, int32_t kc_out, int32_t Kc, const float* thr_rows, int32_t* knn_col, float* knn_cor,
                     bool with_self)
{
    if (nrows == 0) return;
    int np2 = 2; while (np2 < kc_out) np2 <<= 1;
    if (thr_rows) np2 = std::max(np2, 4096);
    const size_t smem = (size_t)np2 * sizeof(uint64_t);
    MCB_CHECK(smem <= 200 * 1024, "topk_big: Kc too large for shared memory");
    static SmemOptIn optin_t, optin_f;
    ensure_dyn_smem(optin_t, ctx, topk_big_kernel<true>, smem);
    ensure_dyn_smem(optin_f, ctx, topk_big_kernel<false>, smem);
    if (with_self)
        topk_big_kernel<true><<<nrows, SEL_THREADS, smem, ctx->stream>>>(C, ldc, M, rows, nrows, row0, kc_out, Kc, thr_rows, np2, knn_col, knn_cor);
    else
        topk_big_kernel<false><<<nrows, SEL_THREADS, smem, ctx->stream>>>(C, ldc, M, rows, nrows, row0, kc_out, Kc, thr_rows, np2, knn_col, knn_cor);
    MCB_LAUNCH_CHECK();
}

}  // namespace mcb
