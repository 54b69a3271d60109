// corknn_tc.cu -- the cell x cell Pearson correlation of tgs_cor_knn (reference call site
// R/utils.r:119) as a tcgen05 / TMEM GEMM fed by TMA, with the per-row selection fused into the
// epilogue so the N x N matrix never touches HBM.
//
//   C[i][j] = sum_g Z[i][g] * Z[j][g]          Z = standardised log-feature rows (features.cu)
//
// Precision: 3-pass split scheme on FP16 operand planes.  Z * 2^12 is stored as hi + lo with
// hi = fp16(z*S), lo = fp16(z*S - hi): the same 11-bit significands as the split-TF32 scheme, but
// kind::f16 runs at twice the kind::tf32 rate and moves half the bytes; fp16's narrow exponent is
// harmless because Z is standardised (|z| <= 1) and prescaled.  Per k-step the kernel issues
// hi*hi, hi*lo, lo*hi into FP32 TMEM accumulators and rescales by 2^-24 in the epilogue, which
// keeps the result within ~5e-6 of fp64 at G = 1000 (DESIGN.md "Precision", measured in tests).
// NPASS = 1 (hi*hi only) is used by the threshold sampling pass where 1e-3 accuracy is enough.
//
// Kernel shape (persistent, warp specialised, one CTA per SM, cta_group::1):
//   warp 0      TMA producer : cp.async.bulk.tensor 2D tiles, 128-byte swizzle, NSTAGE-deep ring
//   warp 1      MMA issuer   : one elected lane issues tcgen05.mma.kind::f16 (M=128, N=256, K=16);
//                              owns TMEM alloc/dealloc.  TMEM columns [0,256) accumulate hi*hi,
//                              columns [256,512) the two cross terms: the tensor core truncates
//                              when it adds into a large accumulator (measured ~1 ulp per MMA), so
//                              keeping the 2^-11-sized cross terms out of the main sum cuts the
//                              accumulated truncation error 3x (DESIGN.md "Precision")
//   warps 2..9  epilogue     : tcgen05.ld 32x32b.x32 of main + cross, add, then either
//                                STORE  : write the fp32 tile (sample block of the threshold pass), or
//                                FILTER : keep (i, j, c) with c >= thr[i], j != i (~1.5% of elements) and
//                                         append (row, key) records to warp-private 32 KB chunks taken
//                                         from a global pool -- no returning global atomic sits on the
//                                         tile's critical path; scatter_candidates_kernel bins the
//                                         records into the per-row candidate lists afterwards
//   The two accumulators fill TMEM, so the MMA warp waits while the 8 epilogue warps drain a tile
//   (~3% of a tile's MMA time at G = 1000).
// Operand sharing (CM = 2, the FILTER kernel): CTAs are launched as clusters of two that work on
// vertically adjacent row tiles of the same column tile.  Each CTA fetches its own A tile and HALF
// of the shared B tile, multicasting that half into both CTAs' shared memory, so every CTA reads
// 64 KB instead of 96 KB per k-block from L2 (the measured bound of this kernel is the L2 -> SM
// operand feed, not the tensor pipe; profiles/).  A stage is recycled only when both CTAs' MMAs
// have retired (tcgen05.commit multicast onto both empty barriers).
// Tile order: groups of GROUP_M row tiles sweep the column tiles together so the A panel of the
// group (GROUP_M x 1 MB) stays L2 resident and each B tile is fetched from HBM once per group.
#include "internal.h"
#include "ptx.cuh"

namespace mcb {

using namespace ptx;

constexpr int BM = 128;            // rows of C per tile (TMEM lanes)
constexpr int BN = 256;            // columns of C per tile (TMEM columns per accumulator)
constexpr int BK = 64;             // fp16 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;         // kind::f16
constexpr int ELT = 2;             // bytes per operand element
constexpr int GROUP_M = 16;
constexpr int TC_THREADS = 320;    // 10 warps: TMA, MMA, 8 epilogue
constexpr int EPI_WARPS = 8;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t POOL_CHUNK = 2048; // records (16 B each) per pool chunk; >= 32 lanes x 64 records
constexpr int FILTER_CM = 1;        // CTAs per cluster in the FILTER kernel (B tile multicast)

template <int NPASS> struct TcCfg {
    static constexpr int A_PLANES = NPASS == 3 ? 2 : 1;
    static constexpr int A_BYTES = BM * BK * ELT;                      // 16 KB per plane
    static constexpr int B_BYTES = BN * BK * ELT;                      // 32 KB per plane
    static constexpr int STAGE_BYTES = A_PLANES * (A_BYTES + B_BYTES);
    static constexpr int NSTAGE = NPASS == 3 ? 2 : 4;
    static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + EPI_WARPS * (BN / 2) * 4;
};

struct TcParams {
    int64_t a_row0;      // first row of A (global row index of local row 0), used for the self test
    int64_t a_rows;      // valid local rows
    int64_t b_rows;      // valid columns
    int32_t num_kb;      // Gpad / BK
    int32_t m_tiles, n_tiles, total_tiles;
    // STORE
    float* C; int64_t ldc;
    // FILTER: per-row thresholds in, candidate records out (chunk pool)
    const float* thr;
    uint4* pool; uint32_t* pool_cursor; uint32_t* chunk_fill; uint32_t pool_chunks; uint32_t* overflow;
};

// Tile enumeration shared by the three warp roles (and by the host to size the grid).  Row units are
// grouped GROUP_M at a time; a group sweeps its column tiles with the row unit varying fastest.  In
// the symmetric mode (square C = Z Z^T, every unordered pair visited once) a group starts at the
// first column tile that reaches the diagonal, so groups have different widths.
struct TileIter {
    int m_units, n_tiles, sym, rows_per_unit;
    int g = 0, off = 0;                         // current group and the linear id of its first tile
    __host__ __device__ TileIter(int mu, int nt, int s, int rpu) : m_units(mu), n_tiles(nt), sym(s), rows_per_unit(rpu) {}
    __host__ __device__ int gmc(int grp) const { int c = m_units - grp * GROUP_M; return c < GROUP_M ? c : GROUP_M; }
    __host__ __device__ int ns(int grp) const { return sym ? (grp * GROUP_M * rows_per_unit) / BN : 0; }
    __host__ __device__ int size(int grp) const { const int c = n_tiles - ns(grp); return gmc(grp) * (c > 0 ? c : 0); }
    __host__ __device__ int total() const
    {
        int tot = 0;
        for (int grp = 0; grp * GROUP_M < m_units; ++grp) tot += size(grp);
        return tot;
    }
    // tile ids must be presented in increasing order
    __device__ void coords(int tile, int& m_unit, int& n_blk)
    {
        while (tile >= off + size(g)) { off += size(g); ++g; }
        const int within = tile - off, c = gmc(g);
        m_unit = g * GROUP_M + within % c;
        n_blk = ns(g) + within / c;
    }
};

// symmetric mode: a tile whose last column does not exceed its first row holds no pair (i, j > i)
template <int MODE, int CM>
__device__ __forceinline__ bool below_diagonal(int m_blk, int n_blk)
{
    return MODE == 2 && CM == 1 && (n_blk + 1) * BN <= m_blk * BM + 1;
}

template <int NPASS, int MODE, int CM>   // MODE 0 = STORE, 1 = FILTER, 2 = symmetric FILTER; CM = CTAs per cluster sharing a B tile
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_cor_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
              const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
              const TcParams prm)
{
    using Cfg = TcCfg<NPASS>;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128-byte swizzle atoms
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::NSTAGE * Cfg::STAGE_BYTES);
    uint64_t* full_bar = bars;                       // [NSTAGE]
    uint64_t* empty_bar = bars + Cfg::NSTAGE;        // [NSTAGE]
    uint64_t* tfull_bar = bars + 2 * Cfg::NSTAGE;    // accumulators complete -> epilogue
    uint64_t* tempty_bar = bars + 2 * Cfg::NSTAGE + 1;   // epilogue drained -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::NSTAGE + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // work units are "cluster tiles": CM vertically adjacent row tiles x one column tile
    const int crank = CM > 1 ? (int)cluster_ctarank() : 0;
    const int m_units = (prm.m_tiles + CM - 1) / CM;
    const TileIter iter0(m_units, prm.n_tiles, MODE == 2 ? 1 : 0, CM * BM);
    const int total_tiles = prm.total_tiles;
    const int unit0 = (int)blockIdx.x / CM, unit_stride = (int)gridDim.x / CM;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_a_hi); prefetch_tensormap(&map_b_hi);
        if (NPASS == 3) { prefetch_tensormap(&map_a_lo); prefetch_tensormap(&map_b_lo); }
        for (int s = 0; s < Cfg::NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CM); }
        mbar_init(tfull_bar, 1); mbar_init(tempty_bar, EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    if (CM > 1) cluster_sync_all();        // the peer's barriers exist before anything is multicast onto them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            TileIter iter = iter0;
            for (int tile = unit0; tile < total_tiles; tile += unit_stride) {
                int m_blk, n_blk;
                iter.coords(tile, m_blk, n_blk);
                m_blk = m_blk * CM + crank;
                if (below_diagonal<MODE, CM>(m_blk, n_blk)) continue;
                const int32_t m0 = (int32_t)(prm.a_row0 + (int64_t)m_blk * BM);
                const int32_t n0 = n_blk * BN;
                for (int kb = 0; kb < prm.num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* st = smem + stage * Cfg::STAGE_BYTES;
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    tma_load_2d(st, &map_a_hi, &full_bar[stage], kb * BK, m0);
                    if (NPASS == 3) tma_load_2d(st + Cfg::A_BYTES, &map_a_lo, &full_bar[stage], kb * BK, m0);
                    if (CM == 1) {
                        tma_load_2d(st + Cfg::A_PLANES * Cfg::A_BYTES, &map_b_hi, &full_bar[stage], kb * BK, n0);
                        if (NPASS == 3) tma_load_2d(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, &map_b_lo, &full_bar[stage], kb * BK, n0);
                    } else {
                        // this CTA's share of the B tile (BN / CM rows; the B maps then have a BN / CM row box),
                        // delivered to the same offset in every CTA of the cluster
                        constexpr int SHARE = Cfg::B_BYTES / CM;
                        const int32_t nr = n0 + crank * (BN / CM);
                        const uint16_t mask = (uint16_t)((1u << CM) - 1u);
                        tma_load_2d_mc(st + Cfg::A_PLANES * Cfg::A_BYTES + crank * SHARE, &map_b_hi, &full_bar[stage], kb * BK, nr, mask);
                        if (NPASS == 3)
                            tma_load_2d_mc(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES + crank * SHARE, &map_b_lo, &full_bar[stage], kb * BK, nr, mask);
                    }
                    if (++stage == Cfg::NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(BM, BN);
            int stage = 0; uint32_t phase = 0;
            uint32_t tphase = 0;
            const uint32_t d_main = tmem_base, d_cross = tmem_base + (uint32_t)BN;
            TileIter iter = iter0;
            for (int tile = unit0; tile < total_tiles; tile += unit_stride) {
                if (MODE == 2) {
                    int m_blk, n_blk;
                    iter.coords(tile, m_blk, n_blk);
                    if (below_diagonal<MODE, CM>(m_blk * CM + crank, n_blk)) continue;
                }
                mbar_wait(tempty_bar, tphase ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < prm.num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint64_t a_hi = make_sw128_kmajor_desc(st);
                    const uint64_t b_hi = make_sw128_kmajor_desc(st + Cfg::A_PLANES * Cfg::A_BYTES);
                    const uint64_t a_lo = make_sw128_kmajor_desc(st + Cfg::A_BYTES);
                    const uint64_t b_lo = make_sw128_kmajor_desc(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t koff = (uint64_t)((k * UMMA_K * ELT) >> 4);   // 32 bytes per k step
                        const uint32_t accum = (kb | k) != 0 ? 1u : 0u;    // first k-step overwrites
                        mma_f16_ss(d_main, a_hi + koff, b_hi + koff, idesc, accum);
                        if (NPASS == 3) {
                            mma_f16_ss(d_cross, a_hi + koff, b_lo + koff, idesc, accum);
                            mma_f16_ss(d_cross, a_lo + koff, b_hi + koff, idesc, 1u);
                        }
                    }
                    // frees the smem stage (in every CTA that multicasts into it) when these MMAs retire
                    if (CM == 1) tc_commit(&empty_bar[stage]);
                    else tc_commit_mc(&empty_bar[stage], (uint16_t)((1u << CM) - 1u));
                    if (++stage == Cfg::NSTAGE) { stage = 0; phase ^= 1; }
                }
                tc_commit(tfull_bar);                      // accumulators complete -> epilogue
                tphase ^= 1;
            }
        }
    } else {
        // ===== epilogue warps 2..9 : TMEM lanes (warp % 4) * 32 .. + 31, column half (warp - 2) / 4 =====
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        uint32_t tphase = 0;
        TileIter iter = iter0;
        float* tcol = reinterpret_cast<float*>(tmem_slot + 4) + (warp - 2) * (BN / 2);    // per-warp column thresholds (MODE 2)
        uint32_t chunk_id = 0, fill = 0;          // this warp's open record chunk (warp-uniform)
        bool have_chunk = false;
        for (int tile = unit0; tile < total_tiles; tile += unit_stride) {
            int m_blk, n_blk;
            iter.coords(tile, m_blk, n_blk);
            m_blk = m_blk * CM + crank;
            if (below_diagonal<MODE, CM>(m_blk, n_blk)) continue;
            const int64_t lrow = (int64_t)m_blk * BM + quarter * 32 + lane;       // local row
            const bool row_ok = lrow < prm.a_rows;
            const int64_t n0 = (int64_t)n_blk * BN;
            float t = 0.f;
            if (MODE >= 1) t = row_ok ? prm.thr[lrow] : __int_as_float(0x7f800000);
            if (MODE == 2) {
                __syncwarp();
#pragma unroll
                for (int q = lane; q < BN / 2; q += 32) {
                    const int64_t col = n0 + half * (BN / 2) + q;
                    tcol[q] = col < prm.b_rows ? prm.thr[col] : __int_as_float(0x7f800000);
                }
                __syncwarp();
            }
            mbar_wait(tfull_bar, tphase);
            tphase ^= 1;
            tc_fence_after();
            const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
            for (int c = half * (BN / 64); c < (half + 1) * (BN / 64); ++c) {
                uint32_t v[32];
                tmem_ld_32x32(taddr0 + (uint32_t)(c * 32), v);
                if (NPASS == 3) {
                    uint32_t u[32];
                    tmem_ld_32x32(taddr0 + (uint32_t)(BN + c * 32), u);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 32; ++q)
                        v[q] = __float_as_uint((__uint_as_float(v[q]) + __uint_as_float(u[q])) * C_RESCALE);
                } else {
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 32; ++q) v[q] = __float_as_uint(__uint_as_float(v[q]) * C_RESCALE);
                }
                const int64_t col0 = n0 + c * 32;
                if (MODE == 0) {
                    if (row_ok && col0 < prm.b_rows) {
                        float* dst = prm.C + lrow * prm.ldc + col0;
                        if (col0 + 32 <= prm.b_rows) {
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                reinterpret_cast<float4*>(dst)[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                                                                __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                        } else {
#pragma unroll
                            for (int q = 0; q < 32; ++q)
                                if (col0 + q < prm.b_rows) dst[q] = __uint_as_float(v[q]);
                        }
                    }
                } else {
                    // pass masks of this lane's 32 values: bit q of mr -> record for row lrow (column col0 + q),
                    // bit q of mcm (symmetric mode) -> record for row col0 + q (column lrow)
                    uint32_t mr = 0, mcm = 0;
                    if (MODE == 1) {
                        const int64_t self = prm.a_row0 + lrow;
#pragma unroll
                        for (int q = 0; q < 32; ++q) {
                            const int64_t col = col0 + q;
                            const bool ok = row_ok && col < prm.b_rows && col != self;
                            mr |= (ok && __uint_as_float(v[q]) >= t) ? (1u << q) : 0u;
                        }
                    } else {
                        const float* tc = tcol + (c - half * (BN / 64)) * 32;
#pragma unroll
                        for (int q = 0; q < 32; ++q) {
                            const int64_t col = col0 + q;
                            const bool ok = row_ok && col > lrow && col < prm.b_rows;     // each unordered pair once
                            const float cv = __uint_as_float(v[q]);
                            mr |= (ok && cv >= t) ? (1u << q) : 0u;
                            mcm |= (ok && cv >= tc[q]) ? (1u << q) : 0u;
                        }
                    }
                    const uint32_t n = (uint32_t)(__popc(mr) + __popc(mcm));
                    uint32_t incl = n;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += up;
                    }
                    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                    if (total) {                                   // warp-uniform
                        if (!have_chunk || fill + total > POOL_CHUNK) {
                            if (have_chunk && lane == 0 && chunk_id < prm.pool_chunks) prm.chunk_fill[chunk_id] = fill;
                            uint32_t id = 0;
                            if (lane == 0) id = atomicAdd(prm.pool_cursor, 1u);      // one per ~2048 records
                            chunk_id = __shfl_sync(0xffffffffu, id, 0);
                            fill = 0; have_chunk = true;
                            if (chunk_id >= prm.pool_chunks && lane == 0) *prm.overflow = 1u;
                        }
                        if (chunk_id < prm.pool_chunks) {
                            uint4* dst = prm.pool + (size_t)chunk_id * POOL_CHUNK + fill + (incl - n);
#pragma unroll
                            for (int q = 0; q < 32; ++q)
                                if (mr & (1u << q)) {
                                    const uint64_t key = cand_key(__uint_as_float(v[q]), (uint32_t)(col0 + q));
                                    *dst++ = make_uint4((uint32_t)key, (uint32_t)(key >> 32), (uint32_t)lrow, 0u);
                                }
                            if (MODE == 2) {
#pragma unroll
                                for (int q = 0; q < 32; ++q)
                                    if (mcm & (1u << q)) {
                                        const uint64_t key = cand_key(__uint_as_float(v[q]), (uint32_t)lrow);
                                        *dst++ = make_uint4((uint32_t)key, (uint32_t)(key >> 32), (uint32_t)(col0 + q), 0u);
                                    }
                            }
                        }
                        fill += total;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar);
        }
        if (MODE >= 1 && have_chunk && lane == 0 && chunk_id < prm.pool_chunks) prm.chunk_fill[chunk_id] = fill;
    }

    tc_fence_before();
    __syncthreads();
    if (CM > 1) cluster_sync_all();        // no CTA exits while its peer may still signal its barriers
    if (warp == 1) { tc_fence_after(); tmem_dealloc<TMEM_COLS>(tmem_base); }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

void tc_init()
{
    if (g_encode) return;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    MCB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    MCB_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available in this driver");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
}

// 2D fp16 tensor [rows][Gpad] row-major; box = BK columns x box_rows rows; 128-byte swizzle
static CUtensorMap make_map(const void* base, int64_t rows, int32_t Gpad, int box_rows)
{
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)Gpad, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)Gpad * ELT};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char b[160];
        snprintf(b, sizeof(b), "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld Gpad=%d)", (int)r, (long long)rows, Gpad);
        throw Error(2, b);
    }
    return m;
}

template <int NPASS, int MODE, int CM>
static void launch_tc(mcb_ctx* ctx, const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh,
                      const CUtensorMap& bl, const TcParams& prm)
{
    using Cfg = TcCfg<NPASS>;
    static bool configured = false;
    if (!configured) {
        MCB_CUDA(cudaFuncSetAttribute(tc_cor_kernel<NPASS, MODE, CM>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured = true;
    }
    TcParams p2 = prm;
    p2.total_tiles = TileIter((int)ceil_div(prm.m_tiles, CM), prm.n_tiles, MODE == 2 ? 1 : 0, CM * BM).total();
    const int units = p2.total_tiles;
    const int grid = std::min(units, ctx->num_sms / CM) * CM;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CM; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    MCB_CUDA(cudaLaunchKernelEx(&cfg, tc_cor_kernel<NPASS, MODE, CM>, ah, al, bh, bl, p2));
    MCB_LAUNCH_CHECK();
}

void launch_tc_store(mcb_ctx* ctx, const void* Ahi, int64_t a_rows_alloc, const void* Bhi, int64_t b_rows_alloc,
                     int32_t Gpad, float* C, int64_t ldc, int64_t a_rows, int64_t b_rows)
{
    if (a_rows == 0 || b_rows == 0) return;
    tc_init();
    MCB_CHECK(Gpad % BK == 0, "Gpad must be a multiple of 64");
    const CUtensorMap ah = make_map(Ahi, a_rows_alloc, Gpad, BM);
    const CUtensorMap bh = make_map(Bhi, b_rows_alloc, Gpad, BN);
    TcParams prm{};
    prm.a_row0 = 0; prm.a_rows = a_rows; prm.b_rows = b_rows; prm.num_kb = Gpad / BK;
    prm.m_tiles = (int32_t)ceil_div(a_rows, BM); prm.n_tiles = (int32_t)ceil_div(b_rows, BN);
    prm.C = C; prm.ldc = ldc;
    launch_tc<1, 0, 1>(ctx, ah, ah, bh, bh, prm);
}

uint32_t tc_pool_chunk_records() { return POOL_CHUNK; }
uint32_t tc_pool_min_chunks(mcb_ctx* ctx) { return (uint32_t)(EPI_WARPS * ctx->num_sms + 64); }   // one open chunk per warp

void launch_tc_filter(mcb_ctx* ctx, const void* Zhi, const void* Zlo, int64_t rows_alloc, int64_t a_row0,
                      int64_t a_rows, int64_t b_rows, int32_t Gpad, const float* thr, const CandPool& pool)
{
    if (a_rows == 0 || b_rows == 0) return;
    tc_init();
    MCB_CHECK(Gpad % BK == 0, "Gpad must be a multiple of 64");
    const CUtensorMap ah = make_map(Zhi, rows_alloc, Gpad, BM);
    const CUtensorMap al = make_map(Zlo, rows_alloc, Gpad, BM);
    // clusters of FILTER_CM CTAs share each B tile: every CTA loads (and multicasts) BN / FILTER_CM of its rows
    const CUtensorMap bh = make_map(Zhi, rows_alloc, Gpad, BN / FILTER_CM);
    const CUtensorMap bl = make_map(Zlo, rows_alloc, Gpad, BN / FILTER_CM);
    TcParams prm{};
    prm.a_row0 = a_row0; prm.a_rows = a_rows; prm.b_rows = b_rows; prm.num_kb = Gpad / BK;
    prm.m_tiles = (int32_t)ceil_div(a_rows, BM); prm.n_tiles = (int32_t)ceil_div(b_rows, BN);
    prm.thr = thr;
    prm.pool = reinterpret_cast<uint4*>(pool.records); prm.pool_cursor = pool.cursor; prm.chunk_fill = pool.chunk_fill;
    prm.pool_chunks = pool.n_chunks; prm.overflow = pool.overflow;
    // the whole square on one rank: every unordered pair is contracted once and serves both rows
    if (a_row0 == 0 && a_rows == b_rows) launch_tc<3, 2, FILTER_CM>(ctx, ah, al, bh, bl, prm);
    else launch_tc<3, 1, FILTER_CM>(ctx, ah, al, bh, bl, prm);
}

// validation entry: full 3-pass product stored to C (used by tests to check the tensor-core path
// element by element against the SIMT kernel and fp64)
void launch_tc_store3(mcb_ctx* ctx, const void* Zhi, const void* Zlo, int64_t rows_alloc, int64_t a_rows,
                      int64_t b_rows, int32_t Gpad, float* C, int64_t ldc)
{
    if (a_rows == 0 || b_rows == 0) return;
    tc_init();
    const CUtensorMap ah = make_map(Zhi, rows_alloc, Gpad, BM);
    const CUtensorMap al = make_map(Zlo, rows_alloc, Gpad, BM);
    const CUtensorMap bh = make_map(Zhi, rows_alloc, Gpad, BN);
    const CUtensorMap bl = make_map(Zlo, rows_alloc, Gpad, BN);
    TcParams prm{};
    prm.a_row0 = 0; prm.a_rows = a_rows; prm.b_rows = b_rows; prm.num_kb = Gpad / BK;
    prm.m_tiles = (int32_t)ceil_div(a_rows, BM); prm.n_tiles = (int32_t)ceil_div(b_rows, BN);
    prm.C = C; prm.ldc = ldc;
    launch_tc<3, 0, 1>(ctx, ah, al, bh, bl, prm);
}

}  // namespace mcb
