// balance.cu -- tgs_graph(x_knn, knn = K, k_expand, k_beta): rank-product balancing of the kNN lists
// (reference call site R/utils.r:120).  The rule is the one cormat_to_balanced_knn states in R
// (R/cgraph_knn_mcnorm.r:48-81):
//     keep pair iff  K*K*k_expand - rank(i,j)*rank(j,i) > 0                (:66-68)
//     per node keep the k_beta*K incoming edges with the best score        (:70-71)
//     per node keep the K outgoing edges with the best score               (:73-74)
// and the weights are the rank fractions 1 - (place-1)/K every consumer assumes
// (R/cgraph_knn.r:46, R/mc_confusion.r:26).  Tie-breaks follow oracle/mc_oracle.c orc_balance:
// ties on s = rank(i,j)*rank(j,i) go to the lower node index; self edges dropped unless keep_self.
//
// Data layout: the reverse rank rank(j,i) is found through a per-row open-addressing hash table
// (H = 2^k >= 2*Kc slots of u32 = (column << rank_bits) | rank0) instead of a binary search or a
// global sort.  Probing the tables directly costs one 64-byte DRAM atom per probe (11.5 GB for 10^8 probes at C2,
// profiles/r2a_ncu_stages.csv), so the join is organised by the BLOCK of rows that holds the probed tables:
//   one rank  -- BLOCKED LISTS (prepare_blocked / probe_blocked): the kernel that builds a row's table also writes the
//                row's list grouped by target block; the probe pass walks (block, row) segments block-major, so the 32 MB
//                slice of tables it probes stays in L2, and appends the qualifying pairs to compact per-row lists of
//                mutual keys that the two cuts then read (no record stream, no count / scatter passes);
//   N ranks   -- ROUTED RECORDS (route_* / probe_kernel): every list entry (i -> j, rank r) becomes an 8-byte record
//                bucketed by (owner rank, block) of j; the buckets of other ranks travel in one all-to-all and the probe
//                pass walks the received records block by block.  The score is symmetric, s(i,j) = s(j,i), so the probe
//                of row j's table by the record of (i -> j) is recorded in ROW j's score array at position rank(j,i):
//                results land at the owner of the probed table and nothing travels back; the lists never leave their owner.
//   direct    -- every entry probes wherever its target's table lies (join_impl = 1; validation, single rank).
// All three give bit-identical graphs.  Selection kernels are one CTA per row with the row's keys in registers.
#include "internal.h"
#include <cstdlib>

namespace mcb {

constexpr int BAL_THREADS = 256;
#ifndef MCB_BAL_MIN_BLOCKS
#define MCB_BAL_MIN_BLOCKS 8          // 32 registers: the cuts are latency bound (barriers, dependent loads), occupancy pays
#endif

// persistent row-striding kernels launch whole waves only: with a grid of num_sms * 8 and 3 resident CTAs per SM the
// last 0.67 wave ran at a third of the machine
template <typename Kern>
static int wave_grid(mcb_ctx* ctx, Kern kern, size_t smem, int64_t rows)
{
    int per_sm = 0;
    MCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BAL_THREADS, smem));
    return (int)std::min<int64_t>(rows, (int64_t)ctx->num_sms * std::max(per_sm, 1));
}
constexpr uint32_t HASH_EMPTY = 0xffffffffu;
constexpr int BAL_BINS = 2048;         // histogram of the rank products s used to select instead of sorting

__device__ __forceinline__ uint32_t hash_slot(uint32_t key, int log2H) { return (key * 2654435761u) >> (32 - log2H); }

__global__ void __launch_bounds__(BAL_THREADS)
build_hash_kernel(const int32_t* __restrict__ knn_col, int64_t M, int32_t Kc, int32_t H, int32_t log2H,
                  int32_t rank_bits, uint32_t* __restrict__ table)
{
    extern __shared__ uint32_t stab[];      // [H]
    for (int64_t row = blockIdx.x; row < M; row += gridDim.x) {
        for (int q = threadIdx.x; q < H; q += BAL_THREADS) stab[q] = HASH_EMPTY;
        __syncthreads();
        const int32_t* kc = knn_col + row * (int64_t)Kc;
        for (int r = threadIdx.x; r < Kc; r += BAL_THREADS) {
            const int32_t j = kc[r];
            if (j < 0) continue;
            const uint32_t val = ((uint32_t)j << rank_bits) | (uint32_t)r;
            uint32_t slot = hash_slot((uint32_t)j, log2H);
            while (atomicCAS(&stab[slot], HASH_EMPTY, val) != HASH_EMPTY) slot = (slot + 1) & (uint32_t)(H - 1);
        }
        __syncthreads();
        uint32_t* out = table + row * (int64_t)H;
        for (int q = threadIdx.x; q < H; q += BAL_THREADS) out[q] = stab[q];
        __syncthreads();
    }
}

// 1-based rank of `target` in row j's list, 0 if absent
__device__ __forceinline__ uint32_t hash_lookup(const uint32_t* __restrict__ tab, int32_t H, int32_t log2H,
                                                int32_t rank_bits, uint32_t target)
{
    uint32_t slot = hash_slot(target, log2H);
    const uint32_t rmask = (1u << rank_bits) - 1u;
    for (int probe = 0; probe < H; ++probe) {
        const uint32_t v = __ldg(tab + slot);
        if (v == HASH_EMPTY) return 0u;
        if ((v >> rank_bits) == target) return (v & rmask) + 1u;
        slot = (slot + 1) & (uint32_t)(H - 1);
    }
    return 0u;
}


// ---------------------------------------------------------------------------------------------------------------
// routed join
// ---------------------------------------------------------------------------------------------------------------
// Bucket of a target row j: (owner rank, block of `blk` rows inside the owner's row range).
struct RouteGeom {
    uint32_t rows_per_rank;     // contiguous row split of the M valid rows over the ranks (1 rank: every row is rank 0's)
    int32_t blk_shift;          // block = 2^blk_shift rows
    int32_t nbr;                // blocks per rank = ceil(rows_per_rank / blk)
    int32_t nbuckets;           // nranks * nbr
    int32_t rank_bits;
};
__device__ __forceinline__ int32_t route_bucket(const RouteGeom& g, int32_t j, uint32_t& jl)
{
    const uint32_t owner = g.nbuckets == g.nbr ? 0u : (uint32_t)j / g.rows_per_rank;          // 32-bit division, none on one rank
    const uint32_t lrow = (uint32_t)j - owner * g.rows_per_rank;
    jl = (uint32_t)lrow & ((1u << g.blk_shift) - 1u);
    return (int32_t)(owner * (uint32_t)g.nbr + (lrow >> g.blk_shift));
}
constexpr int ROUTE_THREADS = 256;

// pass 1: counts[cta][bucket] = entries of the CTA's rows that go to `bucket`.  Every CTA owns a contiguous chunk of
// the rank's rows (the same chunk in pass 2) and walks it row by row; the histogram lives in shared memory (+1 atomics
// without a return value: the hardware aggregates the lanes of a warp that hit the same bin).
__global__ void __launch_bounds__(ROUTE_THREADS)
route_count_kernel(const int32_t* __restrict__ knn_col, int64_t row0, int64_t rows, int32_t Kc, RouteGeom geo, int keep_self,
                   uint32_t* __restrict__ counts)
{
    extern __shared__ uint32_t hist[];      // [nbuckets]
    for (int q = threadIdx.x; q < geo.nbuckets; q += ROUTE_THREADS) hist[q] = 0;
    __syncthreads();
    const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * per, r1 = min(rows, r0 + per);
    for (int64_t lrow = r0; lrow < r1; ++lrow) {
        const int32_t* kc = knn_col + lrow * Kc;
        const int32_t self = keep_self ? -1 : (int32_t)(row0 + lrow);
        for (int r = threadIdx.x; r < Kc; r += ROUTE_THREADS) {
            const int32_t j = kc[r];
            if (j < 0 || j == self) continue;
            uint32_t jl;
            atomicAdd(&hist[route_bucket(geo, j, jl)], 1u);
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < geo.nbuckets; q += ROUTE_THREADS) counts[(int64_t)blockIdx.x * geo.nbuckets + q] = hist[q];
}
// per bucket: exclusive prefix of the CTAs' counts (in place) and the bucket total.  One warp per bucket, 32 CTAs per step.
__global__ void __launch_bounds__(128)
route_offsets_kernel(uint32_t* __restrict__ counts, int32_t nctas, int32_t nbuckets, int64_t* __restrict__ total,
                     int64_t* __restrict__ padded)
{
    const int lane = threadIdx.x & 31;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= nbuckets) return;
    uint32_t carry = 0;
    for (int c0 = 0; c0 < nctas; c0 += 32) {
        const int c = c0 + lane;
        const uint32_t v = c < nctas ? counts[(int64_t)c * nbuckets + b] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
        if (c < nctas) counts[(int64_t)c * nbuckets + b] = carry + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) { total[b] = carry; padded[b] = ((int64_t)carry + 1) & ~(int64_t)1; }
}
// Buckets start on 16-byte boundaries (an even number of 8-byte records): NCCL's send / recv moves 16-byte vectors only when
// source and destination are aligned alike, and the unaligned path was measured at a third of the bandwidth (0.2 GB in
// 1.7 ms instead of 0.65 ms on two ranks).  The pad slot of an odd bucket holds a record no table answers.
constexpr uint64_t ROUTE_PAD_RECORD = 0xffffffff00000000ull;
__global__ void route_pad_kernel(int32_t nbuckets, const int64_t* __restrict__ total, const int64_t* __restrict__ bucket_start,
                                 uint64_t* __restrict__ records)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nbuckets && (total[b] & 1)) records[bucket_start[b] + total[b]] = ROUTE_PAD_RECORD;
}
// pass 2: records[bucket_start[b] + within[cta][b] + k] = (source row i : 32 | row of j inside its block : * | rank0 : rank_bits).
// The lanes of a warp that target the same bucket take their slots with ONE shared-memory atomic (match_any + popc).
__global__ void __launch_bounds__(ROUTE_THREADS)
route_scatter_kernel(const int32_t* __restrict__ knn_col, int64_t row0, int64_t rows, int32_t Kc, RouteGeom geo, int keep_self,
                     const uint32_t* __restrict__ within, const int64_t* __restrict__ bucket_start, uint64_t* __restrict__ records)
{
    extern __shared__ long long rbase[];                // [nbuckets] first slot of this CTA in each bucket's stream
    uint32_t* cursor = reinterpret_cast<uint32_t*>(rbase + geo.nbuckets);      // [nbuckets] entries written so far
    for (int q = threadIdx.x; q < geo.nbuckets; q += ROUTE_THREADS) {
        rbase[q] = bucket_start[q] + (long long)within[(int64_t)blockIdx.x * geo.nbuckets + q];
        cursor[q] = 0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * per, r1 = min(rows, r0 + per);
    const int Kc_pad = (Kc + 31) & ~31;                 // whole warps stay converged for the match
    for (int64_t lrow = r0; lrow < r1; ++lrow) {
        const int32_t* kc = knn_col + lrow * Kc;
        const uint32_t i = (uint32_t)(row0 + lrow);
        const int32_t self = keep_self ? -1 : (int32_t)i;
        for (int r = threadIdx.x; r < Kc_pad; r += ROUTE_THREADS) {
            const int32_t j = r < Kc ? kc[r] : -1;
            const bool live = j >= 0 && j != self;
            uint32_t jl = 0;
            const int32_t b = live ? route_bucket(geo, j, jl) : -1 - lane;          // dead lanes match nobody
            const unsigned peers = __match_any_sync(0xffffffffu, b);
            const int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if (live && lane == leader) base = atomicAdd(&cursor[b], (uint32_t)__popc(peers));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (live) {
                const long long pos = rbase[b] + (long long)(base + (uint32_t)__popc(peers & ((1u << lane) - 1u)));
                records[pos] = ((uint64_t)i << 32) | (uint64_t)((jl << geo.rank_bits) | (uint32_t)r);
            }
        }
    }
}

// probe pass.  The records of this rank's blocks are visited block by block: segment q covers positions
// [seg_prefix[q], seg_prefix[q+1]) of the visiting order, lives at records + seg_src[q] and belongs to local block
// seg_block[q].  For record (i -> j, r): if i sits at rank r' of j's list and s = (r+1)(r'+1) passes the threshold,
// sym[j][r'] = s  (the score of j's own entry for i -- s is symmetric).
__global__ void __launch_bounds__(256)
probe_kernel(const uint64_t* __restrict__ records, int64_t n_records, const int64_t* __restrict__ seg_prefix,
             const int64_t* __restrict__ seg_src, const int32_t* __restrict__ seg_block, int32_t nseg, int32_t blk_shift,
             const uint32_t* __restrict__ table, int32_t H, int32_t log2H, int32_t rank_bits, int32_t Kc, long long smax, int strict,
             uint32_t* __restrict__ sym)
{
    extern __shared__ long long sp[];      // [nseg + 1] prefix, then seg_src [nseg], then seg_block as long long [nseg]
    long long* ssrc = sp + nseg + 1;
    long long* sblk = ssrc + nseg;
    for (int q = threadIdx.x; q <= nseg; q += 256) sp[q] = seg_prefix[q];
    for (int q = threadIdx.x; q < nseg; q += 256) { ssrc[q] = seg_src[q]; sblk[q] = seg_block[q]; }
    __syncthreads();
    const uint32_t rmask = (1u << rank_bits) - 1u;
    if (sp[nseg] < n_records) n_records = sp[nseg];          // the prefix ends at the true record count
    int lo = 0;                             // segment of the thread's current position: positions only move forward
    {
        const int64_t v0 = (int64_t)blockIdx.x * 256 + threadIdx.x;
        int hi = nseg;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (sp[mid] <= v0) lo = mid; else hi = mid; }
    }
    for (int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x; v < n_records; v += (int64_t)gridDim.x * 256) {
        while (lo + 1 < nseg && sp[lo + 1] <= v) ++lo;
        const uint64_t rec = records[ssrc[lo] + (v - sp[lo])];
        const uint32_t i = (uint32_t)(rec >> 32), w = (uint32_t)rec;
        const int64_t jrow = ((int64_t)sblk[lo] << blk_shift) + (w >> rank_bits);          // row of j among the rank's own rows
        const uint32_t rji = hash_lookup(table + jrow * H, H, log2H, rank_bits, i);          // 1-based, 0 = absent
        if (rji) {
            const long long s = (long long)((w & rmask) + 1u) * (long long)rji;
            if (strict ? (s < smax) : (s <= smax)) sym[jrow * Kc + (rji - 1u)] = (uint32_t)s;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// blocked join (single rank): the lists themselves are bucketed, no record stream
// ---------------------------------------------------------------------------------------------------------------
// prepare: one CTA per row builds the row's rank table (as build_hash_kernel) AND writes the row's list grouped by the
// block of 2^blk_shift rows that holds each target: blist[row][off[row][b] .. off[row][b+1]) = (j << rank_bits) | rank0
// for the targets j of block b (self dropped unless keep_self).  The probe pass then walks (block, row) segments
// block-major: the slice of tables it probes (<= 32 MB) stays in L2, every table atom comes from HBM once, and the
// list is read in contiguous segments -- no count pass, no scatter pass, no 8-byte records.
constexpr int MAX_JOIN_BLOCKS = 128;
__global__ void __launch_bounds__(BAL_THREADS)
prepare_blocked_kernel(const int32_t* __restrict__ knn_col, int64_t M, int32_t Kc, int32_t H, int32_t log2H, int32_t rank_bits,
                       int32_t blk_shift, int32_t nb, int keep_self, uint32_t* __restrict__ table, uint32_t* __restrict__ blist,
                       uint16_t* __restrict__ off)
{
    extern __shared__ uint32_t stab[];      // [H]
    __shared__ uint32_t bcount[MAX_JOIN_BLOCKS + 1], bcur[MAX_JOIN_BLOCKS];
    for (int64_t row = blockIdx.x; row < M; row += gridDim.x) {
        for (int q = threadIdx.x; q < H; q += BAL_THREADS) stab[q] = HASH_EMPTY;
        for (int q = threadIdx.x; q <= nb; q += BAL_THREADS) bcount[q] = 0;
        __syncthreads();
        const int32_t* kc = knn_col + row * (int64_t)Kc;
        for (int r = threadIdx.x; r < Kc; r += BAL_THREADS) {
            const int32_t j = kc[r];
            if (j < 0) continue;
            const uint32_t val = ((uint32_t)j << rank_bits) | (uint32_t)r;
            uint32_t slot = hash_slot((uint32_t)j, log2H);
            while (atomicCAS(&stab[slot], HASH_EMPTY, val) != HASH_EMPTY) slot = (slot + 1) & (uint32_t)(H - 1);
            if (j != (int32_t)row || keep_self) atomicAdd(&bcount[j >> blk_shift], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 32) {             // exclusive prefix of the block counts (nb <= 128: four per lane)
            uint32_t v[MAX_JOIN_BLOCKS / 32], loc = 0;
#pragma unroll
            for (int q = 0; q < MAX_JOIN_BLOCKS / 32; ++q) {
                const int b = (int)threadIdx.x * (MAX_JOIN_BLOCKS / 32) + q;
                v[q] = b < nb ? bcount[b] : 0u;
                loc += v[q];
            }
            uint32_t incl = loc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if ((int)threadIdx.x >= o) incl += u; }
            uint32_t run = incl - loc;
#pragma unroll
            for (int q = 0; q < MAX_JOIN_BLOCKS / 32; ++q) {
                const int b = (int)threadIdx.x * (MAX_JOIN_BLOCKS / 32) + q;
                if (b < nb) { bcur[b] = run; off[row * (int64_t)(nb + 1) + b] = (uint16_t)run; }
                run += v[q];
            }
            if (threadIdx.x == 31) off[row * (int64_t)(nb + 1) + nb] = (uint16_t)run;
        }
        __syncthreads();
        uint32_t* bl = blist + row * (int64_t)Kc;
        for (int r = threadIdx.x; r < Kc; r += BAL_THREADS) {
            const int32_t j = kc[r];
            if (j < 0 || (j == (int32_t)row && !keep_self)) continue;
            bl[atomicAdd(&bcur[j >> blk_shift], 1u)] = ((uint32_t)j << rank_bits) | (uint32_t)r;
        }
        uint32_t* out = table + row * (int64_t)H;
        for (int q = threadIdx.x; q < H; q += BAL_THREADS) out[q] = stab[q];
        __syncthreads();
    }
}

// probe, block-major: item = b * M + i, one warp per item.  For every entry (i -> j, rank r) of the segment whose
// reverse rank r' exists and whose score s = (r + 1) * r' passes the threshold, the key (s << 32) | j is appended to
// row i's COMPACT list of mutual edges (one returning atomic per warp and segment; the order inside a row's list is
// arbitrary, every later step selects by key).  Most of a row's Kc entries are not mutual or fail the threshold, so
// the later steps read a fraction of the list.
// L2 cache policies: the block's tables should stay resident while the one-pass streams (list entries in, mutual keys out)
// go through
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint32_t ld_hint(const uint32_t* p, uint64_t pol)
{
    uint32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__global__ void __launch_bounds__(256)
probe_blocked_kernel(const uint32_t* __restrict__ blist, const uint16_t* __restrict__ off, int64_t M, int32_t Kc, int32_t nb,
                     const uint32_t* __restrict__ table, int32_t H, int32_t log2H, int32_t rank_bits, long long smax, int strict,
                     uint64_t* __restrict__ mut, uint32_t* __restrict__ n_mut)
{
    const int lane = threadIdx.x & 31;
    const uint32_t warp0 = ((uint32_t)blockIdx.x * 256u + threadIdx.x) >> 5, nwarps = ((uint32_t)gridDim.x * 256u) >> 5;
    const uint32_t rmask = (1u << rank_bits) - 1u;
    const uint32_t Mu = (uint32_t)M;
    const uint64_t keep = l2_policy_evict_last(), stream = l2_policy_evict_first();
    // scores are products of two ranks <= 4096; strict: s < smax  <=>  s <= smax - 1
    const uint32_t slim = (uint32_t)(smax > 0x7fffffffll ? 0x7fffffffll : smax) - (strict ? 1u : 0u);
    // item = b * M + i, visited block-major; (b, i) advance without a division
    const uint32_t step_b = nwarps / Mu, step_i = nwarps - step_b * Mu;
    uint32_t b = warp0 / Mu, i = warp0 - b * Mu;
    for (; b < (uint32_t)nb; ) {
        const uint16_t* o = off + (size_t)i * (uint32_t)(nb + 1) + b;
        const int o0 = o[0], o1 = o[1];
        const uint32_t* bl = blist + (size_t)i * (uint32_t)Kc;
        // two entries per lane and pass: their list loads, then their first table probes, are in flight together
        // (a segment holds ~Kc / nb entries; the latency chain offsets -> entries -> probes -> slot bounds this kernel)
        for (int e0 = o0; e0 < o1; e0 += 64) {
            const int ea = e0 + lane, eb = e0 + 32 + lane;
            const uint32_t wa = ea < o1 ? ld_hint(bl + ea, stream) : 0u, wb = eb < o1 ? ld_hint(bl + eb, stream) : 0u;
            const uint32_t ja = wa >> rank_bits, jb = wb >> rank_bits;
            const uint32_t* ta = table + (size_t)ja * (uint32_t)H;
            const uint32_t* tb = table + (size_t)jb * (uint32_t)H;
            uint32_t sa = hash_slot(i, log2H), sb = sa;
            uint32_t va = ea < o1 ? ld_hint(ta + sa, keep) : HASH_EMPTY, vb = eb < o1 ? ld_hint(tb + sb, keep) : HASH_EMPTY;
            while (va != HASH_EMPTY && (va >> rank_bits) != i) { sa = (sa + 1) & (uint32_t)(H - 1); va = ld_hint(ta + sa, keep); }
            while (vb != HASH_EMPTY && (vb >> rank_bits) != i) { sb = (sb + 1) & (uint32_t)(H - 1); vb = ld_hint(tb + sb, keep); }
            uint32_t ra = va != HASH_EMPTY ? (va & rmask) + 1u : 0u, rb = vb != HASH_EMPTY ? (vb & rmask) + 1u : 0u;
            if (ea < o1 && ja == i) ra = (wa & rmask) + 1u;            // self (keep_self): its own rank
            if (eb < o1 && jb == i) rb = (wb & rmask) + 1u;
            const uint32_t sca = ((wa & rmask) + 1u) * ra, scb = ((wb & rmask) + 1u) * rb;
            const bool oka = ra != 0u && sca <= slim, okb = rb != 0u && scb <= slim;
            const unsigned ma = __ballot_sync(0xffffffffu, oka), mb = __ballot_sync(0xffffffffu, okb);
            if (ma | mb) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&n_mut[i], (uint32_t)(__popc(ma) + __popc(mb)));
                base = __shfl_sync(0xffffffffu, base, 0);
                uint64_t* dst = mut + (size_t)i * (uint32_t)Kc + base;
                const unsigned below = (1u << lane) - 1u;
                if (oka) __stcs(reinterpret_cast<unsigned long long*>(dst + __popc(ma & below)), ((unsigned long long)sca << 32) | ja);
                if (okb) __stcs(reinterpret_cast<unsigned long long*>(dst + __popc(ma) + __popc(mb & below)), ((unsigned long long)scb << 32) | jb);
            }
        }
        i += step_i; b += step_b;
        if (i >= Mu) { i -= Mu; ++b; }
    }
}

// mutual scores s(i, r) and the in-degree threshold of every owned row
template <int IPT>      // 256 * IPT >= Kc keys sorted per row, register resident
__global__ void __launch_bounds__(BAL_THREADS, IPT <= 4 ? MCB_BAL_MIN_BLOCKS : 1)
mutual_kernel(const int32_t* __restrict__ knn_col, const uint32_t* __restrict__ table, int64_t row0, int64_t rows,
              int32_t Kc, int32_t H, int32_t log2H, int32_t rank_bits, long long smax, int strict, int keep_self,
              int32_t kin, uint32_t* __restrict__ sym, uint64_t* __restrict__ thr_in, int have_scores,
              const uint64_t* __restrict__ mut, const uint32_t* __restrict__ n_mut)
{
    // have_scores = 2: the row's mutual keys are the compact list mut / n_mut (blocked join);
    // have_scores = 1: sym already holds the scores (routed join); otherwise they are computed here by probing the
    // tables of ALL rows (direct join, single rank only; kept as the cross-check of the routed one).
    // knn_col holds the rank's own rows (row lrow of the buffer = global row row0 + lrow).
    extern __shared__ uint64_t sk[];        // [256 * IPT]
    __shared__ uint32_t hist[BAL_BINS];
    __shared__ int s_n, s_sel, s_bin, s_before;
    int bin_shift = 0;
    while ((smax >> bin_shift) >= BAL_BINS) ++bin_shift;
    for (int64_t lrow = blockIdx.x; lrow < rows; lrow += gridDim.x) {
        const int64_t i = row0 + lrow;
        const int32_t* kc = knn_col + lrow * (int64_t)Kc;
        uint32_t* so = sym + lrow * (int64_t)Kc;
        uint64_t key[IPT];
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
            const int r = q * BAL_THREADS + threadIdx.x;          // coalesced over the rank list (any order sorts)
            uint64_t kv = ~0ull;
            if (r < Kc) {
                const int32_t j = have_scores == 2 ? 0 : kc[r];
                uint32_t sv = 0;
                if (have_scores == 2) {          // compact list of mutual keys (blocked join): n_mut[lrow] keys at mut
                    if (r < (int)n_mut[lrow]) kv = mut[lrow * (int64_t)Kc + r];
                } else if (have_scores) {
                    sv = so[r];
                    if (sv) kv = ((uint64_t)sv << 32) | (uint32_t)j;
                } else {
                    if (j >= 0 && (j != (int32_t)i || keep_self)) {
                        const uint32_t rji = (j == (int32_t)i) ? (uint32_t)(r + 1)
                                                               : hash_lookup(table + (int64_t)j * H, H, log2H, rank_bits, (uint32_t)i);
                        if (rji) {
                            const long long s = (long long)(r + 1) * (long long)rji;
                            if (strict ? (s < smax) : (s <= smax)) { sv = (uint32_t)s; kv = ((uint64_t)s << 32) | (uint32_t)j; }
                        }
                    }
                    so[r] = sv;
                }
            }
            key[q] = kv;
        }
        // thr_in = the kin-th smallest key when more than kin mutual edges exist.  A 2048-bin histogram of the
        // scores s locates its bin; only that bin's few keys are sorted (256-key network).  Ties that overfill
        // the bin fall back to sorting the whole row.
        int local = 0;
#pragma unroll
        for (int q = 0; q < IPT; ++q) local += key[q] != ~0ull;
        if (threadIdx.x == 0) { s_n = 0; s_sel = 0; }
        for (int q = threadIdx.x; q < BAL_BINS; q += BAL_THREADS) hist[q] = 0;
        __syncthreads();
        if (local) atomicAdd(&s_n, local);
#pragma unroll
        for (int q = 0; q < IPT; ++q) if (key[q] != ~0ull) atomicAdd(&hist[(uint32_t)(key[q] >> 32) >> bin_shift], 1u);
        __syncthreads();
        const int n_mut = s_n;
        uint64_t thr = ~0ull;
        if (kin >= 1 && n_mut > kin) {
            if (threadIdx.x < 32) {
                int bb; long long before;
                warp_find_bucket<BAL_BINS / 32>(hist, (long long)(kin - 1), bb, before);
                if (threadIdx.x == 0) { s_bin = bb; s_before = (int)before; }
            }
            __syncthreads();
            const int bsel = s_bin, m = (int)hist[bsel], idx = kin - 1 - s_before;
            if (m <= 256) {
#pragma unroll
                for (int q = 0; q < IPT; ++q)
                    if (key[q] != ~0ull && (int)((uint32_t)(key[q] >> 32) >> bin_shift) == bsel) sk[atomicAdd(&s_sel, 1)] = key[q];
                __syncthreads();
                // the bin holds a handful of keys: each finds its rank by counting (keys are distinct), no sorting network
                if ((int)threadIdx.x < m) {
                    const uint64_t k = sk[threadIdx.x];
                    int rank = 0;
                    for (int q = 0; q < m; ++q) rank += sk[q] < k ? 1 : 0;
                    if (rank == idx) thr_in[i] = k;
                }
            } else {
                bitonic_sort_regs_u64<IPT>(key, sk);
#pragma unroll
                for (int q = 0; q < IPT; ++q) if (threadIdx.x * IPT + q == kin - 1) thr_in[i] = key[q];
            }
        } else if (threadIdx.x == 0) {
            thr_in[i] = thr;
        }
        __syncthreads();
    }
}

// out-degree prune: survivors of the target's in-degree cut, best K by (s, target)
template <int IPT>
__global__ void __launch_bounds__(BAL_THREADS, IPT <= 4 ? MCB_BAL_MIN_BLOCKS : 1)
prune_kernel(const int32_t* __restrict__ knn_col, const uint32_t* __restrict__ sym, const uint64_t* __restrict__ thr_in,
             int64_t row0, int64_t rows, int32_t Kc, int32_t K, long long smax, int32_t* __restrict__ out_deg,
             int32_t* __restrict__ out_dst, int32_t* __restrict__ out_s, const uint64_t* __restrict__ mut,
             const uint32_t* __restrict__ n_mut)
{
    extern __shared__ uint64_t sk[];        // [256 * IPT]
    __shared__ uint32_t hist[BAL_BINS];
    __shared__ int s_n, s_sel, s_bin, s_before;
    int bin_shift = 0;
    while ((smax >> bin_shift) >= BAL_BINS) ++bin_shift;
    for (int64_t lrow = blockIdx.x; lrow < rows; lrow += gridDim.x) {
        const int64_t i = row0 + lrow;
        const int32_t* kc = knn_col + lrow * (int64_t)Kc;          // own rows only
        const uint32_t* so = sym + lrow * (int64_t)Kc;
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        int local = 0;
        uint64_t key[IPT];
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
            const int r = q * BAL_THREADS + threadIdx.x;
            uint64_t kv = ~0ull;
            if (mut) {                       // compact list of mutual keys (blocked join)
                if (r < (int)n_mut[lrow]) {
                    const uint64_t k = mut[lrow * (int64_t)Kc + r];
                    const uint64_t inkey = (k & 0xffffffff00000000ull) | (uint32_t)i;   // edge i -> j seen from target j
                    if (inkey <= thr_in[(uint32_t)k]) { kv = k; ++local; }
                }
            } else if (r < Kc) {
                const uint32_t s = so[r];
                if (s) {
                    const int32_t j = kc[r];
                    const uint64_t inkey = ((uint64_t)s << 32) | (uint32_t)i;   // edge i -> j seen from target j
                    if (inkey <= thr_in[j]) { kv = ((uint64_t)s << 32) | (uint32_t)j; ++local; }
                }
            }
            key[q] = kv;
        }
        for (int q = threadIdx.x; q < BAL_BINS; q += BAL_THREADS) hist[q] = 0;
        if (threadIdx.x == 0) s_sel = 0;
        __syncthreads();
        if (local) atomicAdd(&s_n, local);
#pragma unroll
        for (int q = 0; q < IPT; ++q) if (key[q] != ~0ull) atomicAdd(&hist[(uint32_t)(key[q] >> 32) >> bin_shift], 1u);
        __syncthreads();
        // the K best survivors: histogram of s -> bin of the K-th; the keys up to that bin (a few more than K) are
        // compacted and sorted with the 256-key network; more than 256 of them (ties) -> sort the whole row
        const int n_surv = s_n;
        int bsel = BAL_BINS - 1, total = n_surv;
        if (n_surv > K) {
            if (threadIdx.x < 32) {
                int bb; long long before;
                warp_find_bucket<BAL_BINS / 32>(hist, (long long)(K - 1), bb, before);
                if (threadIdx.x == 0) { s_bin = bb; s_before = (int)before + (int)hist[bb]; }
            }
            __syncthreads();
            bsel = s_bin; total = s_before;
        }
        if (total <= 256 && K <= 256) {
#pragma unroll
            for (int q = 0; q < IPT; ++q)
                if (key[q] != ~0ull && (int)((uint32_t)(key[q] >> 32) >> bin_shift) <= bsel) sk[atomicAdd(&s_sel, 1)] = key[q];
            __syncthreads();
            // ~K + a few keys: rank by counting (keys are distinct) instead of a 256-key sorting network with its barriers
            uint64_t k = ~0ull;
            int rank = 0;
            if ((int)threadIdx.x < total) {
                k = sk[threadIdx.x];
                for (int q = 0; q < total; ++q) rank += sk[q] < k ? 1 : 0;
            }
            __syncthreads();
            if ((int)threadIdx.x < total) sk[rank] = k;
        } else {
            bitonic_sort_regs_u64<IPT>(key, sk);
            __syncthreads();
#pragma unroll
            for (int q = 0; q < IPT; ++q) {
                const int e = threadIdx.x * IPT + q;
                if (e < K) sk[e] = key[q];
            }
        }
        __syncthreads();
        const int n = s_n < K ? s_n : K;
        if (threadIdx.x == 0) out_deg[lrow] = n;
        for (int q = threadIdx.x; q < K; q += BAL_THREADS) {
            const bool ok = q < n;
            out_dst[lrow * (int64_t)K + q] = ok ? (int32_t)(uint32_t)sk[q] : -1;
            out_s[lrow * (int64_t)K + q] = ok ? (int32_t)(sk[q] >> 32) : 0;
        }
        __syncthreads();
    }
}

// (mc1, mc2, w) rows of the tgCellGraph edge table (reference R/cgraph_knn.r:22), original cell ids
__global__ void emit_edges_kernel(const int32_t* __restrict__ out_deg, const int32_t* __restrict__ out_dst,
                                  const int64_t* __restrict__ offs, const int32_t* __restrict__ cell_of_row,
                                  int64_t row0, int64_t rows, int32_t K, int32_t* __restrict__ src,
                                  int32_t* __restrict__ dst, double* __restrict__ w)
{
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < rows * K;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t lrow = idx / K;
        const int q = (int)(idx % K);
        if (q >= out_deg[lrow]) continue;
        const int64_t e = offs[lrow] + q;
        src[e] = cell_of_row[row0 + lrow];
        dst[e] = cell_of_row[out_dst[lrow * (int64_t)K + q]];
        w[e] = 1.0 - (double)q / (double)K;
    }
}

__global__ void widen_i32_kernel(const int32_t* __restrict__ in, int64_t* __restrict__ out, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = in[i];
}

static int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

template <typename F>
static void ensure_smem(mcb_ctx* ctx, F* kernel, size_t smem, SmemOptIn& optin)
{
    MCB_CHECK(smem <= 200 * 1024, "balance: K*k_expand too large for shared memory");
    ensure_dyn_smem(optin, ctx, kernel, smem);
}

void launch_build_hash(mcb_ctx* ctx, const int32_t* knn_col, int64_t M, int32_t Kc, int32_t H, int32_t rank_bits,
                       uint32_t* table)
{
    if (M == 0) return;
    static SmemOptIn optin;
    const size_t smem = (size_t)H * sizeof(uint32_t);
    ensure_smem(ctx, build_hash_kernel, smem, optin);
    const int grid = wave_grid(ctx, build_hash_kernel, smem, M);
    build_hash_kernel<<<grid, BAL_THREADS, smem, ctx->stream>>>(knn_col, M, Kc, H, ilog2(H), rank_bits, table);
    MCB_LAUNCH_CHECK();
}

template <int IPT>
static void launch_mutual_t(mcb_ctx* ctx, const int32_t* knn_col, const uint32_t* table, int64_t row0, int64_t rows,
                            int32_t Kc, int32_t H, int32_t rank_bits, int64_t smax, int strict, int keep_self, int32_t kin,
                            uint32_t* sym, uint64_t* thr_in, int have_scores, const uint64_t* mut, const uint32_t* n_mut)
{
    static SmemOptIn optin;
    const size_t smem = (size_t)IPT * 256 * sizeof(uint64_t);
    ensure_smem(ctx, mutual_kernel<IPT>, smem, optin);
    const int grid = wave_grid(ctx, mutual_kernel<IPT>, smem, rows);
    mutual_kernel<IPT><<<grid, BAL_THREADS, smem, ctx->stream>>>(knn_col, table, row0, rows, Kc, H, ilog2(H), rank_bits,
                                                                  (long long)smax, strict, keep_self, kin, sym, thr_in, have_scores, mut, n_mut);
    MCB_LAUNCH_CHECK();
}
template <int IPT>
static void launch_prune_t(mcb_ctx* ctx, const int32_t* knn_col, const uint32_t* sym, const uint64_t* thr_in, int64_t row0,
                           int64_t rows, int32_t Kc, int32_t K, int64_t smax, int32_t* out_deg, int32_t* out_dst, int32_t* out_s,
                           const uint64_t* mut, const uint32_t* n_mut)
{
    static SmemOptIn optin;
    const size_t smem = (size_t)IPT * 256 * sizeof(uint64_t);
    ensure_smem(ctx, prune_kernel<IPT>, smem, optin);
    const int grid = wave_grid(ctx, prune_kernel<IPT>, smem, rows);
    prune_kernel<IPT><<<grid, BAL_THREADS, smem, ctx->stream>>>(knn_col, sym, thr_in, row0, rows, Kc, K, (long long)smax, out_deg, out_dst, out_s,
                                                                 mut, n_mut);
    MCB_LAUNCH_CHECK();
}

#define MCB_DISPATCH_IPT(Kc, CALL)                                   \
    do {                                                              \
        if ((Kc) <= 512) { CALL(2); }                                 \
        else if ((Kc) <= 1024) { CALL(4); }                           \
        else if ((Kc) <= 2048) { CALL(8); }                           \
        else { CALL(16); }                                            \
    } while (0)


// ---- routed join launchers ------------------------------------------------------------------------------------------
RouteShape route_shape(mcb_ctx* ctx, int64_t rows_per_rank, int32_t nranks, int32_t H, int32_t rank_bits)
{
    RouteShape rs;
    // block of rows whose tables take <= 16 MB (a quarter of one L2 partition, with the score rows beside them)
    int32_t shift = 13;
    while (shift > 6 && ((int64_t)H * 4 << shift) > ((int64_t)16 << 20)) --shift;
    int32_t nbr = (int32_t)std::max<int64_t>(1, ceil_div(rows_per_rank, (int64_t)1 << shift));
    while ((int64_t)nbr * nranks > 4096) { ++shift; nbr = (int32_t)std::max<int64_t>(1, ceil_div(rows_per_rank, (int64_t)1 << shift)); }
    MCB_CHECK(shift + rank_bits <= 32, "balance: block too large for the packed record");
    rs.blk_shift = shift; rs.nbr = nbr; rs.nbuckets = nbr * nranks;
    rs.grid = ctx->num_sms * 4;
    return rs;
}
void launch_route_records(mcb_ctx* ctx, const int32_t* knn_col_own, int64_t row0, int64_t rows, int32_t Kc, int64_t rows_per_rank,
                          const RouteShape& rs, int32_t rank_bits, int keep_self, uint32_t* counts /* [grid][nbuckets] */,
                          int64_t* bucket_total /* [nbuckets] true counts */, int64_t* bucket_padded /* [nbuckets] rounded up to even */,
                          int64_t* bucket_start /* [nbuckets + 1] prefix of the padded counts */, uint64_t* records /* rows * Kc + nbuckets */)
{
    RouteGeom geo{(uint32_t)rows_per_rank, rs.blk_shift, rs.nbr, rs.nbuckets, rank_bits};
    const size_t smem1 = (size_t)rs.nbuckets * sizeof(uint32_t), smem2 = (size_t)rs.nbuckets * (sizeof(long long) + sizeof(uint32_t));
    static SmemOptIn o1, o2;
    ensure_dyn_smem(o1, ctx, route_count_kernel, smem1);
    ensure_dyn_smem(o2, ctx, route_scatter_kernel, smem2);
    route_count_kernel<<<rs.grid, ROUTE_THREADS, smem1, ctx->stream>>>(knn_col_own, row0, rows, Kc, geo, keep_self, counts);
    MCB_LAUNCH_CHECK();
    route_offsets_kernel<<<(int)ceil_div(rs.nbuckets, 4), 128, 0, ctx->stream>>>(counts, rs.grid, rs.nbuckets, bucket_total, bucket_padded);
    MCB_LAUNCH_CHECK();
    launch_exclusive_scan_i64(ctx, bucket_padded, bucket_start, rs.nbuckets, bucket_start + rs.nbuckets);
    route_scatter_kernel<<<rs.grid, ROUTE_THREADS, smem2, ctx->stream>>>(knn_col_own, row0, rows, Kc, geo, keep_self, counts, bucket_start, records);
    MCB_LAUNCH_CHECK();
    route_pad_kernel<<<(int)ceil_div(rs.nbuckets, 256), 256, 0, ctx->stream>>>(rs.nbuckets, bucket_total, bucket_start, records);
    MCB_LAUNCH_CHECK();
}
void launch_probe(mcb_ctx* ctx, const uint64_t* records, int64_t n_records, const int64_t* seg_prefix, const int64_t* seg_src,
                  const int32_t* seg_block, int32_t nseg, int32_t blk_shift, const uint32_t* table_own, int32_t H, int32_t rank_bits,
                  int32_t Kc, int64_t smax, int strict, uint32_t* sym_own)
{
    if (n_records == 0 || nseg == 0) return;
    const size_t smem = (size_t)(3 * nseg + 1) * sizeof(long long);
    static SmemOptIn optin;
    ensure_dyn_smem(optin, ctx, probe_kernel, smem);
    int per_sm = 0;
    MCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, probe_kernel, 256, smem));
    const int grid = (int)std::min<int64_t>(ceil_div(n_records, 256), (int64_t)ctx->num_sms * std::max(per_sm, 1));
    probe_kernel<<<grid, 256, smem, ctx->stream>>>(records, n_records, seg_prefix, seg_src, seg_block, nseg, blk_shift, table_own, H,
                                                   ilog2(H), rank_bits, Kc, (long long)smax, strict, sym_own);
    MCB_LAUNCH_CHECK();
}

// blocked join (single rank).  Block = the largest power of two of rows whose tables fit 32 MB, at most MAX_JOIN_BLOCKS blocks
BlockedShape blocked_shape(int64_t M, int32_t H)
{
    BlockedShape bs;
    static const int64_t block_mb = getenv("MCB_JOIN_BLOCK_MB") ? std::max(1, atoi(getenv("MCB_JOIN_BLOCK_MB"))) : 32;     // tuning hook
    int32_t shift = 6;
    while (((int64_t)H * 4 << (shift + 1)) <= (block_mb << 20)) ++shift;
    while (ceil_div(std::max<int64_t>(M, 1), (int64_t)1 << shift) > MAX_JOIN_BLOCKS) ++shift;
    bs.blk_shift = shift; bs.nb = (int32_t)ceil_div(std::max<int64_t>(M, 1), (int64_t)1 << shift);
    return bs;
}
void launch_prepare_blocked(mcb_ctx* ctx, const int32_t* knn_col, int64_t M, int32_t Kc, int32_t H, int32_t rank_bits, const BlockedShape& bs,
                            int keep_self, uint32_t* table, uint32_t* blist, uint16_t* off)
{
    if (M == 0) return;
    MCB_CHECK(Kc < 65536 && bs.nb <= MAX_JOIN_BLOCKS, "balance: blocked join shape out of range");
    static SmemOptIn optin;
    const size_t smem = (size_t)H * sizeof(uint32_t);
    ensure_smem(ctx, prepare_blocked_kernel, smem, optin);
    const int grid = wave_grid(ctx, prepare_blocked_kernel, smem, M);
    prepare_blocked_kernel<<<grid, BAL_THREADS, smem, ctx->stream>>>(knn_col, M, Kc, H, ilog2(H), rank_bits, bs.blk_shift, bs.nb, keep_self,
                                                                      table, blist, off);
    MCB_LAUNCH_CHECK();
}
void launch_probe_blocked(mcb_ctx* ctx, const uint32_t* blist, const uint16_t* off, int64_t M, int32_t Kc, const BlockedShape& bs,
                          const uint32_t* table, int32_t H, int32_t rank_bits, int64_t smax, int strict, uint64_t* mut, uint32_t* n_mut)
{
    if (M == 0) return;
    int per_sm = 0;
    MCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, probe_blocked_kernel, 256, 0));
    const int grid = (int)std::min<int64_t>(ceil_div((int64_t)bs.nb * M, 8), (int64_t)ctx->num_sms * std::max(per_sm, 1));
    probe_blocked_kernel<<<grid, 256, 0, ctx->stream>>>(blist, off, M, Kc, bs.nb, table, H, ilog2(H), rank_bits, (long long)smax, strict, mut, n_mut);
    MCB_LAUNCH_CHECK();
}

void launch_mutual(mcb_ctx* ctx, const int32_t* knn_col, const uint32_t* table, int64_t row0, int64_t rows,
                   int32_t Kc, int32_t H, int32_t rank_bits, int64_t smax, int strict, int keep_self, int32_t kin,
                   uint32_t* sym, uint64_t* thr_in, int have_scores, const uint64_t* mut, const uint32_t* n_mut)
{
    if (rows == 0) return;
    MCB_CHECK(Kc <= 4096, "MC-ERR: K * k_expand above 4096 is not supported");
    MCB_CHECK(have_scores != 2 || (mut && n_mut), "balance: compact lists missing");
#define CALL(I) launch_mutual_t<I>(ctx, knn_col, table, row0, rows, Kc, H, rank_bits, smax, strict, keep_self, kin, sym, thr_in, have_scores, mut, n_mut)
    MCB_DISPATCH_IPT(Kc, CALL);
#undef CALL
}

void launch_prune(mcb_ctx* ctx, const int32_t* knn_col, const uint32_t* sym, const uint64_t* thr_in, int64_t row0,
                  int64_t rows, int32_t Kc, int32_t K, int64_t smax, int32_t* out_deg, int32_t* out_dst, int32_t* out_s,
                  const uint64_t* mut, const uint32_t* n_mut)
{
    if (rows == 0) return;
    MCB_CHECK(Kc <= 4096, "MC-ERR: K * k_expand above 4096 is not supported");
#define CALL(I) launch_prune_t<I>(ctx, knn_col, sym, thr_in, row0, rows, Kc, K, smax, out_deg, out_dst, out_s, mut, n_mut)
    MCB_DISPATCH_IPT(Kc, CALL);
#undef CALL
}

void launch_widen_i32(mcb_ctx* ctx, const int32_t* in, int64_t* out, int64_t n)
{
    if (n == 0) return;
    int grid = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)ctx->num_sms * 8);
    widen_i32_kernel<<<grid, 256, 0, ctx->stream>>>(in, out, n);
    MCB_LAUNCH_CHECK();
}

void launch_emit_edges(mcb_ctx* ctx, const int32_t* out_deg, const int32_t* out_dst, const int64_t* offs,
                       const int32_t* cell_of_row, int64_t row0, int64_t rows, int32_t K, int32_t* src, int32_t* dst,
                       double* w)
{
    if (rows == 0) return;
    int grid = (int)std::min<int64_t>(ceil_div(rows * K, 256), (int64_t)ctx->num_sms * 16);
    emit_edges_kernel<<<grid, 256, 0, ctx->stream>>>(out_deg, out_dst, offs, cell_of_row, row0, rows, K, src, dst, w);
    MCB_LAUNCH_CHECK();
}

}  // namespace mcb
