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
// global sort, so the join costs ~1.5 random 32-byte sectors per edge.  All kernels are one CTA per
// row with the row's keys sorted in shared memory; HBM-bound integer work.
#include "internal.h"

namespace mcb {

constexpr int BAL_THREADS = 256;

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

// mutual scores s(i, r) and the in-degree threshold of every owned row
template <int IPT>      // 256 * IPT >= Kc keys sorted per row, register resident
__global__ void __launch_bounds__(BAL_THREADS, IPT <= 4 ? 4 : 1)
mutual_kernel(const int32_t* __restrict__ knn_col, const uint32_t* __restrict__ table, int64_t row0, int64_t rows,
              int32_t Kc, int32_t H, int32_t log2H, int32_t rank_bits, long long smax, int strict, int keep_self,
              int32_t kin, uint32_t* __restrict__ sym, uint64_t* __restrict__ thr_in)
{
    extern __shared__ uint64_t sk[];        // [256 * IPT]
    __shared__ uint32_t hist[BAL_BINS];
    __shared__ int s_n, s_sel, s_bin, s_before;
    int bin_shift = 0;
    while ((smax >> bin_shift) >= BAL_BINS) ++bin_shift;
    for (int64_t lrow = blockIdx.x; lrow < rows; lrow += gridDim.x) {
        const int64_t i = row0 + lrow;
        const int32_t* kc = knn_col + i * (int64_t)Kc;
        uint32_t* so = sym + lrow * (int64_t)Kc;
        uint64_t key[IPT];
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
            const int r = q * BAL_THREADS + threadIdx.x;          // coalesced over the rank list (any order sorts)
            uint64_t kv = ~0ull;
            if (r < Kc) {
                const int32_t j = kc[r];
                uint32_t sv = 0;
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
                uint64_t k1[1] = {threadIdx.x < m ? sk[threadIdx.x] : ~0ull};
                __syncthreads();
                bitonic_sort_regs_u64<1>(k1, sk);
                if (threadIdx.x == idx) thr_in[i] = k1[0];
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
__global__ void __launch_bounds__(BAL_THREADS, IPT <= 4 ? 4 : 1)
prune_kernel(const int32_t* __restrict__ knn_col, const uint32_t* __restrict__ sym, const uint64_t* __restrict__ thr_in,
             int64_t row0, int64_t rows, int32_t Kc, int32_t K, long long smax, int32_t* __restrict__ out_deg,
             int32_t* __restrict__ out_dst, int32_t* __restrict__ out_s)
{
    extern __shared__ uint64_t sk[];        // [256 * IPT]
    __shared__ uint32_t hist[BAL_BINS];
    __shared__ int s_n, s_sel, s_bin, s_before;
    int bin_shift = 0;
    while ((smax >> bin_shift) >= BAL_BINS) ++bin_shift;
    for (int64_t lrow = blockIdx.x; lrow < rows; lrow += gridDim.x) {
        const int64_t i = row0 + lrow;
        const int32_t* kc = knn_col + i * (int64_t)Kc;
        const uint32_t* so = sym + lrow * (int64_t)Kc;
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        int local = 0;
        uint64_t key[IPT];
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
            const int r = q * BAL_THREADS + threadIdx.x;
            uint64_t kv = ~0ull;
            if (r < Kc) {
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
            uint64_t k1[1] = {threadIdx.x < total ? sk[threadIdx.x] : ~0ull};
            __syncthreads();
            bitonic_sort_regs_u64<1>(k1, sk);
            __syncthreads();
            sk[threadIdx.x] = k1[0];
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
                            uint32_t* sym, uint64_t* thr_in)
{
    static SmemOptIn optin;
    const size_t smem = (size_t)IPT * 256 * sizeof(uint64_t);
    ensure_smem(ctx, mutual_kernel<IPT>, smem, optin);
    const int grid = wave_grid(ctx, mutual_kernel<IPT>, smem, rows);
    mutual_kernel<IPT><<<grid, BAL_THREADS, smem, ctx->stream>>>(knn_col, table, row0, rows, Kc, H, ilog2(H), rank_bits,
                                                                  (long long)smax, strict, keep_self, kin, sym, thr_in);
    MCB_LAUNCH_CHECK();
}
template <int IPT>
static void launch_prune_t(mcb_ctx* ctx, const int32_t* knn_col, const uint32_t* sym, const uint64_t* thr_in, int64_t row0,
                           int64_t rows, int32_t Kc, int32_t K, int64_t smax, int32_t* out_deg, int32_t* out_dst, int32_t* out_s)
{
    static SmemOptIn optin;
    const size_t smem = (size_t)IPT * 256 * sizeof(uint64_t);
    ensure_smem(ctx, prune_kernel<IPT>, smem, optin);
    const int grid = wave_grid(ctx, prune_kernel<IPT>, smem, rows);
    prune_kernel<IPT><<<grid, BAL_THREADS, smem, ctx->stream>>>(knn_col, sym, thr_in, row0, rows, Kc, K, (long long)smax, out_deg, out_dst, out_s);
    MCB_LAUNCH_CHECK();
}

#define MCB_DISPATCH_IPT(Kc, CALL)                                   \
    do {                                                              \
        if ((Kc) <= 512) { CALL(2); }                                 \
        else if ((Kc) <= 1024) { CALL(4); }                           \
        else if ((Kc) <= 2048) { CALL(8); }                           \
        else { CALL(16); }                                            \
    } while (0)

void launch_mutual(mcb_ctx* ctx, const int32_t* knn_col, const uint32_t* table, int64_t row0, int64_t rows,
                   int32_t Kc, int32_t H, int32_t rank_bits, int64_t smax, int strict, int keep_self, int32_t kin,
                   uint32_t* sym, uint64_t* thr_in)
{
    if (rows == 0) return;
    MCB_CHECK(Kc <= 4096, "MC-ERR: K * k_expand above 4096 is not supported");
#define CALL(I) launch_mutual_t<I>(ctx, knn_col, table, row0, rows, Kc, H, rank_bits, smax, strict, keep_self, kin, sym, thr_in)
    MCB_DISPATCH_IPT(Kc, CALL);
#undef CALL
}

void launch_prune(mcb_ctx* ctx, const int32_t* knn_col, const uint32_t* sym, const uint64_t* thr_in, int64_t row0,
                  int64_t rows, int32_t Kc, int32_t K, int64_t smax, int32_t* out_deg, int32_t* out_dst, int32_t* out_s)
{
    if (rows == 0) return;
    MCB_CHECK(Kc <= 4096, "MC-ERR: K * k_expand above 4096 is not supported");
#define CALL(I) launch_prune_t<I>(ctx, knn_col, sym, thr_in, row0, rows, Kc, K, smax, out_deg, out_dst, out_s)
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
