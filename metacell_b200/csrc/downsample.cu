// downsample.cu -- scm_downsamp (reference R/utils.r:30-69) and colSums on a device-resident CSC.
//
// Semantics restated from the reference:
//   umis = umis[, colSums(umis) >= n]                              (R/utils.r:34)
//   per cell: tabulate(sample(rep(1:m, times = v), size = n))      (R/utils.r:36-39)
// i.e. a uniformly random floor(n)-subset of the cell's UMIs, counted per gene.  R's Mersenne
// Twister stream is not reproduced (DESIGN.md "RNG"): UMI t of cell c gets the 64-bit key
// (Philox4x32-10(seed; t>>2, c)[t&3] << 32) | t and the floor(n) smallest keys are kept -- the same
// distribution as sample(replace = FALSE), independent of evaluation order, and bit-identical
// to the CPU oracle (oracle/mc_oracle.c orc_downsample).
//
// HBM-bound integer work: one CTA per cell, the counts are read twice (sum, final count) and
// written once.  The keys never touch HBM: a cell's Philox words are drawn once and kept in shared memory for
// the later radix passes and the final pass (cells above DS_KEYCAP UMIs redraw them per pass).
#include "internal.h"

namespace mcb {

constexpr int DS_THREADS = 256;

__global__ void convert_x_kernel(const double* __restrict__ xd, uint32_t* __restrict__ x, int64_t nnz)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
        double v = xd[e];
        x[e] = v > 0.0 ? (uint32_t)__double2ll_rn(v) : 0u;
    }
}
__global__ void x_to_double_kernel(const uint32_t* __restrict__ x, double* __restrict__ xd, int64_t nnz)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x)
        xd[e] = (double)x[e];
}

__global__ void col_sums_kernel(int64_t ncells, const int64_t* __restrict__ p, const uint32_t* __restrict__ x,
                                int64_t* __restrict__ tot)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp; c < ncells; c += nwarps) {
        long long s = 0;
        for (int64_t e = p[c] + lane; e < p[c + 1]; e += 32) s += x[e];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) tot[c] = s;
    }
}

__device__ __forceinline__ uint64_t ds_key(uint32_t o, uint32_t t) { return ((uint64_t)o << 32) | (uint64_t)t; }

// block-wide exclusive scan of one int64 per thread; returns exclusive prefix, *total = block sum
__device__ __forceinline__ long long block_excl_scan(long long v, long long* warp_tot /* [NT/32 + 1] */, long long* total)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    __syncthreads();                       // protect warp_tot reuse
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    long long base = 0, all = 0;
    const int nw = blockDim.x >> 5;
    for (int q = 0; q < nw; ++q) { long long t = warp_tot[q]; if (q < w) base += t; all += t; }
    *total = all;
    return base + inc - v;
}

// radix passes over the 64-bit key, most significant bits first: 11 bits, then 8 at a time (53 = 6 * 8 + 5).  With 2,048
// bins the first pass leaves ~1 key in the bucket of the d-th smallest for cells up to a few thousand UMIs, so the select is
// usually one or two passes over the cell's UMIs instead of three or four.
constexpr int DS_BINS = 2048;
__device__ __forceinline__ void ds_pass(int pass, int& shift, int& width)
{
    if (pass == 0) { shift = 53; width = 11; }
    else if (pass <= 6) { shift = 53 - 8 * pass; width = 8; }
    else { shift = 0; width = 5; }
}
constexpr int DS_BITMAP_WORDS = 4096;      // kept-UMI bitmap in shared memory: cells up to 131,072 UMIs take the fast final pass
constexpr int DS_KEYCAP = 5120;            // Philox words of a cell kept in shared memory after the first pass (typical cells:
                                           // a few thousand UMIs), so the later radix passes and the final pass do not redraw them

__global__ void __launch_bounds__(DS_THREADS, 4)
downsample_kernel(int64_t ncells, const int64_t* __restrict__ p, const uint32_t* __restrict__ x, double n,
                  uint32_t seed_lo, uint32_t seed_hi, int add_non_dsamp, uint32_t* __restrict__ out_x,
                  uint8_t* __restrict__ kept, int64_t cell_base)
{
    __shared__ uint32_t hist[DS_BINS];
    __shared__ uint32_t bitmap[DS_BITMAP_WORDS];
    __shared__ __align__(16) uint32_t okey[DS_KEYCAP];
    __shared__ long long warp_tot[DS_THREADS / 32 + 1];
    __shared__ unsigned long long s_prefix, s_thr;
    __shared__ long long s_krem;
    __shared__ int s_done;

    const long long d = (long long)floor(n);
    for (int64_t c = blockIdx.x; c < ncells; c += gridDim.x) {
        const int64_t e0 = p[c], e1 = p[c + 1];
        // 1. column sum.  Every thread owns a CONTIGUOUS slice of the cell's entries, so this one block scan also gives
        //    the thread the UMI offset at which its slice starts (used by the final count): one scan per cell, not
        //    one per 256 entries
        const int64_t per = (e1 - e0 + DS_THREADS - 1) / DS_THREADS;
        const int64_t ea0 = min(e1, e0 + (int64_t)threadIdx.x * per), ea1 = min(e1, ea0 + per);
        long long part = 0;
        for (int64_t e = ea0; e < ea1; ++e) part += x[e];
        long long tot;
        const long long my_start = block_excl_scan(part, warp_tot, &tot);
        if ((double)tot < n) {             // dropped cell (R/utils.r:34)
            for (int64_t e = e0 + threadIdx.x; e < e1; e += DS_THREADS) out_x[e] = add_non_dsamp ? x[e] : 0u;
            if (threadIdx.x == 0) kept[c] = 0;
            continue;
        }
        if (threadIdx.x == 0) kept[c] = 1;
        if (tot == d) {                    // every UMI is drawn
            for (int64_t e = e0 + threadIdx.x; e < e1; e += DS_THREADS) out_x[e] = x[e];
            continue;
        }
        // 2. radix select of the d-th smallest 64-bit key (index d-1)
        if (threadIdx.x == 0) { s_prefix = 0ull; s_krem = d - 1; s_done = 0; s_thr = ~0ull; }
        __syncthreads();
        // the key of a UMI names its cell by the index in the WHOLE matrix, so a cell shard draws the same subset
        const uint32_t c_lo = (uint32_t)(c + cell_base), c_hi = (uint32_t)((uint64_t)(c + cell_base) >> 32);
        const long long ngroups = (tot + 3) >> 2;
        const bool cached = ngroups * 4 <= DS_KEYCAP;          // block-uniform
        for (int pass = 0; pass < 8; ++pass) {
            int shift, width;
            ds_pass(pass, shift, width);
            const uint32_t bmask = (1u << width) - 1u;
            const unsigned long long decided = pass == 0 ? 0ull : (~0ull << (shift + width));
            const unsigned long long prefix = s_prefix;
            for (int q = threadIdx.x; q < (1 << width); q += DS_THREADS) hist[q] = 0;
            __syncthreads();
            for (long long g = threadIdx.x; g < ngroups; g += DS_THREADS) {
                uint32_t o[4];
                if (cached && pass > 0) {
                    const uint4 v = *reinterpret_cast<const uint4*>(&okey[g << 2]);
                    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
                } else {
                    philox4x32_10((uint32_t)g, c_lo, c_hi, TAG_DOWNSAMP, seed_lo, seed_hi, o);
                    if (cached) *reinterpret_cast<uint4*>(&okey[g << 2]) = make_uint4(o[0], o[1], o[2], o[3]);
                }
                if (pass == 0) {                 // nothing decided yet: the digit is the top 11 bits of the Philox word
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if ((g << 2) + q < tot) atomicAdd(&hist[o[q] >> 21], 1u);
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const long long t = (g << 2) + q;
                        if (t < tot) {
                            const uint64_t key = ds_key(o[q], (uint32_t)t);
                            if ((key & decided) == prefix) atomicAdd(&hist[(uint32_t)(key >> shift) & bmask], 1u);
                        }
                    }
                }
            }
            __syncthreads();
            if (threadIdx.x < 32) {
                int dg; long long cum;
                const long long k = s_krem;
                if (width == 11) warp_find_bucket<64>(hist, k, dg, cum);
                else if (width == 8) warp_find_bucket<8>(hist, k, dg, cum);
                else warp_find_bucket<1>(hist, k, dg, cum);
                __syncwarp();
                if (threadIdx.x == 0) {
                    const unsigned long long np = prefix | ((unsigned long long)dg << shift);
                    s_prefix = np;
                    s_krem = k - cum;
                    if (k - cum == (long long)hist[dg] - 1 || pass == 7) {
                        // the bucket's largest element is the threshold: everything in it is kept
                        s_thr = np | (shift ? ((1ull << shift) - 1ull) : 0ull);
                        s_done = 1;
                    }
                }
            }
            __syncthreads();
            if (s_done) break;
        }
        const uint64_t thr = s_thr;
        // 3. per stored entry: how many of its UMIs have key <= thr
        if (tot <= (long long)DS_BITMAP_WORDS * 32) {
            // one more uniform pass over the UMIs: thread -> one bitmap word = 8 Philox blocks = 32 UMIs (no atomics),
            // then every entry counts the bits of its own UMI range.  (The per-entry loop below makes a warp wait for
            // its most expressed gene.)
            const int tot32 = (int)tot;
            const int nwords = (tot32 + 31) >> 5;
            for (int wq = threadIdx.x; wq < nwords; wq += DS_THREADS) {
                uint32_t word = 0;
#pragma unroll
                for (int b8 = 0; b8 < 8; ++b8) {
                    const int g = (wq << 3) + b8;
                    if ((g << 2) >= tot32) break;
                    uint32_t o[4];
                    if (cached) {
                        const uint4 v = *reinterpret_cast<const uint4*>(&okey[g << 2]);
                        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
                    } else philox4x32_10((uint32_t)g, c_lo, c_hi, TAG_DOWNSAMP, seed_lo, seed_hi, o);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int t = (g << 2) + q;
                        if (t < tot32 && ds_key(o[q], (uint32_t)t) <= thr) word |= 1u << (b8 * 4 + q);
                    }
                }
                bitmap[wq] = word;
            }
            __syncthreads();
            int start = (int)my_start;                  // tot <= 131,072 here
            for (int64_t e = ea0; e < ea1; ++e) {
                const int xe = (int)x[e];
                uint32_t cnt = 0;
                if (xe) {
                    const int last = start + xe - 1;
                    const int w0 = start >> 5, w1 = last >> 5;
                    const uint32_t m0 = ~0u << (start & 31), m1 = ~0u >> (31 - (last & 31));
                    if (w0 == w1) cnt = __popc(bitmap[w0] & m0 & m1);
                    else {
                        cnt = __popc(bitmap[w0] & m0) + __popc(bitmap[w1] & m1);
                        for (int wq = w0 + 1; wq < w1; ++wq) cnt += __popc(bitmap[wq]);
                    }
                }
                out_x[e] = cnt;
                start += xe;
            }
            __syncthreads();
            continue;
        }
        long long run = 0;
        for (int64_t base = e0; base < e1; base += DS_THREADS) {
            const int64_t e = base + threadIdx.x;
            const long long xe = e < e1 ? (long long)x[e] : 0;
            long long chunk_tot;
            const long long start = run + block_excl_scan(xe, warp_tot, &chunk_tot);
            run += chunk_tot;
            if (e < e1) {
                uint32_t cnt = 0;
                long long cur_g = -1;
                uint32_t o[4] = {0, 0, 0, 0};
                for (long long t = start; t < start + xe; ++t) {
                    const long long g = t >> 2;
                    if (g != cur_g) { philox4x32_10((uint32_t)g, c_lo, c_hi, TAG_DOWNSAMP, seed_lo, seed_hi, o); cur_g = g; }
                    cnt += (ds_key(o[t & 3], (uint32_t)t) <= thr) ? 1u : 0u;
                }
                out_x[e] = cnt;
            }
        }
        __syncthreads();
    }
}

void launch_convert_x(mcb_ctx* ctx, const double* d_xd, uint32_t* d_x, int64_t nnz)
{
    if (nnz == 0) return;
    int grid = (int)std::min<int64_t>(ceil_div(nnz, 256), (int64_t)ctx->num_sms * 16);
    convert_x_kernel<<<grid, 256, 0, ctx->stream>>>(d_xd, d_x, nnz);
    MCB_LAUNCH_CHECK();
}
void launch_x_to_double(mcb_ctx* ctx, const uint32_t* d_x, double* d_xd, int64_t nnz)
{
    if (nnz == 0) return;
    int grid = (int)std::min<int64_t>(ceil_div(nnz, 256), (int64_t)ctx->num_sms * 16);
    x_to_double_kernel<<<grid, 256, 0, ctx->stream>>>(d_x, d_xd, nnz);
    MCB_LAUNCH_CHECK();
}
void launch_col_sums(mcb_ctx* ctx, int64_t ncells, const int64_t* p, const uint32_t* x, int64_t* tot)
{
    if (ncells == 0) return;
    int grid = (int)std::min<int64_t>(ceil_div(ncells, 8), (int64_t)ctx->num_sms * 16);
    col_sums_kernel<<<grid, 256, 0, ctx->stream>>>(ncells, p, x, tot);
    MCB_LAUNCH_CHECK();
}
void launch_downsample(mcb_ctx* ctx, int64_t ncells, const int64_t* p, const uint32_t* x, double n, uint64_t seed,
                       int add_non_dsamp, uint32_t* out_x, uint8_t* kept, int64_t cell_base)
{
    if (ncells == 0) return;
    MCB_CHECK(n >= 1.0, "MC-ERR: downsample depth must be >= 1");
    int grid = (int)std::min<int64_t>(ncells, (int64_t)ctx->num_sms * 8);
    downsample_kernel<<<grid, DS_THREADS, 0, ctx->stream>>>(ncells, p, x, n, (uint32_t)seed, (uint32_t)(seed >> 32),
                                                             add_non_dsamp, out_x, kept, cell_base);
    MCB_LAUNCH_CHECK();
}

}  // namespace mcb
