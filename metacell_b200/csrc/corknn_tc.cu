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
//   warps 4..11 epilogue     : tcgen05.ld 32x32b.x32 of main + cross, add, then either
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
#include <cuda_fp16.h>
#include <cstdlib>

namespace mcb {

using namespace ptx;

constexpr int BM = 128;            // rows of C per tile (TMEM lanes)
constexpr int BN = 256;            // columns of C per tile (TMEM columns per accumulator)
constexpr int BK = 64;             // fp16 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;         // kind::f16
constexpr int ELT = 2;             // bytes per operand element
constexpr int GROUP_M = 16;
constexpr int TC_THREADS = 384;    // 12 warps = 3 warpgroups: {TMA, MMA, 2 idle}, 2 x 4 epilogue warps
constexpr int EPI_WARP0 = 4;       // first epilogue warp (warpgroup aligned: warp % 4 = TMEM lane quarter)
constexpr int EPI_WARPS = 8;
constexpr uint32_t REGS_CTRL = 56, REGS_EPI = 224;   // setmaxnreg: 128 * 56 + 256 * 224 = 384 * 168 (the CTA's pool; inc blocks forever if it asks for more)
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t POOL_CHUNK = 2048; // records (16 B each) per pool chunk; >= 32 lanes x 64 records
constexpr int FILTER_CM = 1;        // CTAs per cluster in the FILTER kernel (B tile multicast)

template <int NPASS> struct TcCfg {
    static constexpr int A_PLANES = NPASS == 3 ? 2 : 1;
    static constexpr int A_BYTES = BM * BK * ELT;                      // 16 KB per plane
    static constexpr int B_BYTES = BN * BK * ELT;                      // 32 KB per plane
    static constexpr int STAGE_BYTES = A_PLANES * (A_BYTES + B_BYTES);
    static constexpr int NSTAGE = NPASS == 3 ? 2 : 4;
    // per-epilogue-warp scratch: column thresholds (symmetric filter) or, in the 1-pass kernel, the 32 x 32 fp16
    // transpose buffer of the coalesced STORE16 epilogue (32 rows x (16 + 1 pad) words)
    static constexpr int EPI_SCRATCH_WORDS = NPASS == 1 ? 32 * 17 : BN / 2;
    static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + EPI_WARPS * EPI_SCRATCH_WORDS * 4;
};

struct TcParams {
    int64_t a_row0;      // first row of A (global row index of local row 0), used for the self test
    int64_t a_rows;      // valid local rows
    int64_t b_rows;      // valid columns
    int32_t num_kb;      // Gpad / BK
    int32_t last_ksteps; // UMMA k-steps (16 columns each) that hold data in the last k-block: G = 1000 needs 3 of its 4 (columns past G are zero)
    int32_t m_tiles, n_tiles, total_tiles;
    // rank partition of the tile list (symmetric multi-GPU): this launch takes the runs of part_block consecutive
    // tiles whose run index is congruent to `part` modulo `nparts`; local_tiles = how many that is
    int32_t part_block, nparts, part, local_tiles;
    int32_t dbg;         // diagnostics (MCB_TC_DEBUG): 1 = epilogue only waits, 2 = epilogue only reads TMEM
    // STORE (fp32) or, when C16 is set, STORE16: fp16 output of the 1-pass kernel (sampled correlations)
    float* C; int64_t ldc;
    __half* C16; int64_t ldc16;
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

// l-th tile of this launch's share of the tile list (identity when nparts == 1)
__device__ __forceinline__ int part_tile(const TcParams& prm, int l)
{
    return ((l / prm.part_block) * prm.nparts + prm.part) * prm.part_block + (l % prm.part_block);
}
static int part_count(int total, int part_block, int nparts, int part)      // host: size of that share
{
    const int runs = total / part_block, tail = total % part_block;
    int n = (runs / nparts) * part_block + ((runs % nparts) > part ? part_block : 0);
    if (tail && (runs % nparts) == part) n += tail;
    return n;
}

// symmetric mode: a tile whose last column does not exceed its first row holds no pair (i, j > i)
template <int MODE, int CM>
__device__ __forceinline__ bool below_diagonal(int m_blk, int n_blk)
{
    return MODE == 2 && CM == 1 && (n_blk + 1) * BN <= m_blk * BM + 1;
}

// ---------------------------------------------------------------------------------------------
// epilogue of one 128-row x 256-column tile for one epilogue warp (TMEM lanes quarter*32.., column half):
// wait for the accumulators, read main (+ cross), then STORE the tile or FILTER it into candidate records.
// Shared by the 1-SM and the 2-SM kernels (in the latter every CTA drains its own 128 rows).
// ---------------------------------------------------------------------------------------------
template <int NPASS, int MODE>
__device__ __forceinline__ void epilogue_tile(const TcParams& prm, uint32_t tmem_base, int quarter, int half, int lane,
                                              int m_blk, int n_blk, float* tcol, uint64_t* tfull_bar, uint32_t tphase,
                                              uint32_t& chunk_id, uint32_t& fill, bool& have_chunk)
{
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
    tc_fence_after();
    if (prm.dbg == 1) return;
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
        if (prm.dbg == 2) { if (__uint_as_float(v[lane & 31]) == 123.456f) prm.overflow[0] = 2u; continue; }
        if (MODE == 0 && NPASS == 1 && prm.C16) {
            // fp16 output, transposed through the warp's scratch so that each store instruction writes two 64-byte
            // row segments (the direct path writes 16 bytes to 32 different rows per instruction and was store-bound)
            uint32_t* sc = reinterpret_cast<uint32_t*>(tcol);
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const __half2 h = __floats2half2_rn(__uint_as_float(v[2 * k]), __uint_as_float(v[2 * k + 1]));
                sc[lane * 17 + k] = *reinterpret_cast<const uint32_t*>(&h);
            }
            __syncwarp();
            const int sub = lane >> 4, w = lane & 15;
            const int64_t col = col0 + 2 * w;
            const int64_t row_base = (int64_t)m_blk * BM + quarter * 32;
#pragma unroll
            for (int rr = 0; rr < 16; ++rr) {
                const int r = 2 * rr + sub;
                const uint32_t val = sc[r * 17 + w];
                if (row_base + r < prm.a_rows && col < prm.b_rows)
                    *reinterpret_cast<uint32_t*>(prm.C16 + (row_base + r) * prm.ldc16 + col) = val;
            }
        } else if (MODE == 0) {
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
}

// ---------------------------------------------------------------------------------------------
// FILTER epilogue of one 128-row x 256-column tile for one epilogue warp.  The accumulators fill TMEM, so
// the MMA warp cannot start the next tile until this tile is drained: the warp therefore first copies its
// whole 32 x 128 slice (main + cross summed) into registers, releases TMEM at once (`release`), and only
// then does the threshold tests and record writes -- which now overlap the next tile's MMAs.  (Measured at
// C2: filtering straight out of TMEM cost 10 ms of a 32 ms kernel.)  Thresholds are compared in
// accumulator units (x 2^24, exact), so only passing values are rescaled.
// ---------------------------------------------------------------------------------------------
template <int MODE, class Release>
__device__ __forceinline__ void epilogue_filter_tile(const TcParams& prm, uint32_t tmem_base, int quarter, int half, int lane,
                                                     int m_blk, int n_blk, float* tcol, uint64_t* tfull_bar, uint32_t tphase,
                                                     uint32_t& chunk_id, uint32_t& fill, bool& have_chunk, Release release)
{
    constexpr int NCH = BN / 64;                       // 32-column chunks per warp
    constexpr float INV = 1.0f / C_RESCALE;            // 2^24
    const float inf = __int_as_float(0x7f800000);
    const int64_t row_base = (int64_t)m_blk * BM;
    const int64_t lrow = row_base + quarter * 32 + lane;                  // local row
    const bool row_ok = lrow < prm.a_rows;
    const int64_t c_base = (int64_t)n_blk * BN + half * (BN / 2);         // first column of this warp's slice
    const float ts = row_ok ? prm.thr[lrow] * INV : inf;
    bool interior = c_base + BN / 2 <= prm.b_rows && row_base + BM <= prm.a_rows;   // warp-uniform
    if (MODE == 2) {
        __syncwarp();
#pragma unroll
        for (int q = lane; q < BN / 2; q += 32) {
            const int64_t col = c_base + q;
            tcol[q] = col < prm.b_rows ? prm.thr[col] * INV : inf;
        }
        __syncwarp();
        interior = interior && c_base >= row_base + BM;                   // every column is right of every row
    } else {
        const int64_t self0 = prm.a_row0 + row_base;
        interior = interior && (self0 + BM <= c_base || self0 >= c_base + BN / 2);   // no self column in the slice
    }
    mbar_wait(tfull_bar, tphase);
    tc_fence_after();
    if (prm.dbg == 1) { release(); return; }
    uint32_t v[NCH * 32];
    const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * (BN / 2));
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        uint32_t u[32];
        tmem_ld_32x32_p(taddr0 + (uint32_t)(c * 32), v + c * 32);
        tmem_ld_32x32(taddr0 + (uint32_t)(BN + c * 32), u);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; ++q) v[c * 32 + q] = __float_as_uint(__uint_as_float(v[c * 32 + q]) + __uint_as_float(u[q]));
    }
    release();                                          // TMEM is free: the next tile's MMAs may start
    if (prm.dbg == 2) { if (__uint_as_float(v[5]) == 123.456f) prm.overflow[0] = 2u; return; }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const uint32_t* vc = v + c * 32;
        const int64_t col0 = c_base + c * 32;
        uint32_t mr = 0, mcm = 0;
        if (interior) {
            const float4* tc4 = reinterpret_cast<const float4*>(tcol + c * 32);      // 16-byte aligned, broadcast reads
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4) {
                float4 tq = make_float4(0.f, 0.f, 0.f, 0.f);
                if (MODE == 2) tq = tc4[q4];
                const float c0 = __uint_as_float(vc[4 * q4]), c1 = __uint_as_float(vc[4 * q4 + 1]);
                const float c2 = __uint_as_float(vc[4 * q4 + 2]), c3 = __uint_as_float(vc[4 * q4 + 3]);
                mr |= (c0 >= ts ? (1u << (4 * q4)) : 0u) | (c1 >= ts ? (2u << (4 * q4)) : 0u) |
                      (c2 >= ts ? (4u << (4 * q4)) : 0u) | (c3 >= ts ? (8u << (4 * q4)) : 0u);
                if (MODE == 2)
                    mcm |= (c0 >= tq.x ? (1u << (4 * q4)) : 0u) | (c1 >= tq.y ? (2u << (4 * q4)) : 0u) |
                           (c2 >= tq.z ? (4u << (4 * q4)) : 0u) | (c3 >= tq.w ? (8u << (4 * q4)) : 0u);
            }
        } else if (MODE == 1) {
            const int64_t self = prm.a_row0 + lrow;
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const int64_t col = col0 + q;
                const bool ok = row_ok && col < prm.b_rows && col != self;
                mr |= (ok && __uint_as_float(vc[q]) >= ts) ? (1u << q) : 0u;
            }
        } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const int64_t col = col0 + q;
                const bool ok = row_ok && col > lrow && col < prm.b_rows;     // each unordered pair once
                const float cv = __uint_as_float(vc[q]);
                mr |= (ok && cv >= ts) ? (1u << q) : 0u;
                mcm |= (ok && cv >= tcol[c * 32 + q]) ? (1u << q) : 0u;
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
                        const uint64_t key = cand_key(__uint_as_float(vc[q]) * C_RESCALE, (uint32_t)(col0 + q));
                        *dst++ = make_uint4((uint32_t)key, (uint32_t)(key >> 32), (uint32_t)lrow, 0u);
                    }
                if (MODE == 2) {
#pragma unroll
                    for (int q = 0; q < 32; ++q)
                        if (mcm & (1u << q)) {
                            const uint64_t key = cand_key(__uint_as_float(vc[q]) * C_RESCALE, (uint32_t)lrow);
                            *dst++ = make_uint4((uint32_t)key, (uint32_t)(key >> 32), (uint32_t)(col0 + q), 0u);
                        }
                }
            }
            fill += total;
        }
    }
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
    // NPASS == 1 needs only one 256-column accumulator per tile, so it double-buffers TMEM (tile t drains while
    // tile t + 1 accumulates); NPASS == 3 fills all 512 columns with main + cross and uses buffer 0 only
    constexpr int NACC = NPASS == 1 ? 2 : 1;
    uint64_t* tfull_bar = bars + 2 * Cfg::NSTAGE;        // [2] accumulators complete -> epilogue
    uint64_t* tempty_bar = bars + 2 * Cfg::NSTAGE + 2;   // [2] epilogue drained -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::NSTAGE + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // work units are "cluster tiles": CM vertically adjacent row tiles x one column tile
    const int crank = CM > 1 ? (int)cluster_ctarank() : 0;
    const int m_units = (prm.m_tiles + CM - 1) / CM;
    const TileIter iter0(m_units, prm.n_tiles, MODE == 2 ? 1 : 0, CM * BM);
    const int unit0 = (int)blockIdx.x / CM, unit_stride = (int)gridDim.x / CM;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_a_hi); prefetch_tensormap(&map_b_hi);
        if (NPASS == 3) { prefetch_tensormap(&map_a_lo); prefetch_tensormap(&map_b_lo); }
        for (int s = 0; s < Cfg::NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CM); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    if (CM > 1) cluster_sync_all();        // the peer's barriers exist before anything is multicast onto them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < EPI_WARP0) {
    setmaxnreg_dec<REGS_CTRL>();
    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            TileIter iter = iter0;
            for (int l = unit0; l < prm.local_tiles; l += unit_stride) {
                const int tile = part_tile(prm, l);
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
            int it = 0;
            TileIter iter = iter0;
            for (int l = unit0; l < prm.local_tiles; l += unit_stride) {
                const int tile = part_tile(prm, l);
                if (MODE == 2) {
                    int m_blk, n_blk;
                    iter.coords(tile, m_blk, n_blk);
                    if (below_diagonal<MODE, CM>(m_blk * CM + crank, n_blk)) continue;
                }
                const int acc = it % NACC;
                const uint32_t tphase = (uint32_t)(it / NACC) & 1u;
                ++it;
                const uint32_t d_main = tmem_base + (uint32_t)(acc * BN), d_cross = tmem_base + (uint32_t)BN;
                mbar_wait(&tempty_bar[acc], tphase ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < prm.num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint64_t a_hi = make_sw128_kmajor_desc(st);
                    const uint64_t b_hi = make_sw128_kmajor_desc(st + Cfg::A_PLANES * Cfg::A_BYTES);
                    const uint64_t a_lo = make_sw128_kmajor_desc(st + Cfg::A_BYTES);
                    const uint64_t b_lo = make_sw128_kmajor_desc(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES);
                    const int nk = kb == prm.num_kb - 1 ? prm.last_ksteps : BK / UMMA_K;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        if (k >= nk) break;
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
                tc_commit(&tfull_bar[acc]);                // accumulators complete -> epilogue
            }
        }
    }
    } else {
        setmaxnreg_inc<REGS_EPI>();
        // ===== epilogue warps 4..11 : TMEM lanes (warp % 4) * 32 .. + 31, column half (warp - 4) / 4 =====
        const int quarter = warp & 3;
        const int half = (warp - EPI_WARP0) >> 2;
        int it = 0;
        TileIter iter = iter0;
        float* tcol = reinterpret_cast<float*>(tmem_slot + 4) + (warp - EPI_WARP0) * Cfg::EPI_SCRATCH_WORDS;    // per-warp scratch (column thresholds / transpose)
        uint32_t chunk_id = 0, fill = 0;          // this warp's open record chunk (warp-uniform)
        bool have_chunk = false;
        for (int l = unit0; l < prm.local_tiles; l += unit_stride) {
            const int tile = part_tile(prm, l);
            int m_blk, n_blk;
            iter.coords(tile, m_blk, n_blk);
            m_blk = m_blk * CM + crank;
            if (below_diagonal<MODE, CM>(m_blk, n_blk)) continue;
            const int acc = it % NACC;
            const uint32_t tphase = (uint32_t)(it / NACC) & 1u;
            ++it;
            auto release = [&]() { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(&tempty_bar[acc]); };
            if (MODE == 0) {
                epilogue_tile<NPASS, MODE>(prm, tmem_base + (uint32_t)(acc * BN), quarter, half, lane, m_blk, n_blk, tcol,
                                           &tfull_bar[acc], tphase, chunk_id, fill, have_chunk);
                release();
            } else {
                epilogue_filter_tile<(MODE == 0 ? 1 : MODE)>(prm, tmem_base, quarter, half, lane, m_blk, n_blk, tcol, &tfull_bar[acc],
                                                             tphase, chunk_id, fill, have_chunk, release);
            }
        }
        if (MODE >= 1 && have_chunk && lane == 0 && chunk_id < prm.pool_chunks) prm.chunk_fill[chunk_id] = fill;
    }

    tc_fence_before();
    __syncthreads();
    if (CM > 1) cluster_sync_all();        // no CTA exits while its peer may still signal its barriers
    if (warp == 1) { tc_fence_after(); tmem_dealloc<TMEM_COLS>(tmem_base); }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair version of the FILTER kernels (tcgen05 cta_group::2).  The kernel above is bound by the
// L2 -> SM operand feed, so this one halves the B bytes every SM receives: a cluster of two CTAs
// contracts a 256-row x 256-column tile with M = 256 UMMAs issued by the leader CTA; each CTA holds
// its own 128 A rows and only HALF of the B tile (the tensor core reads the other half from the peer's
// shared memory), i.e. 64 KB instead of 96 KB per k-block, which also makes room for a third stage.
// Every CTA drains its own 128 accumulator rows with the same epilogue.
//   full[s]   (leader's)  : leader's producer arrives once expecting both CTAs' bytes; both CTAs' TMA
//                           loads complete their bytes on it (cp.async.bulk.tensor ... cta_group::2)
//   empty[s]  (each CTA)  : leader's tcgen05.commit.cta_group::2 multicast when the stage's MMAs retire
//   tfull     (each CTA)  : same multicast commit after the last k-block of a tile
//   tempty    (leader's)  : 2 x EPI_WARPS arrivals (the peer's warps arrive through the cluster window)
// ---------------------------------------------------------------------------------------------
struct Tc2Cfg {
    static constexpr int A_BYTES = BM * BK * ELT;                    // 16 KB per plane (this CTA's 128 rows)
    static constexpr int B_BYTES = (BN / 2) * BK * ELT;              // 16 KB per plane (this CTA's half of B)
    static constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);      // 64 KB
    static constexpr int NSTAGE = 3;
    static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024 + 256 + EPI_WARPS * (BN / 2) * 4;
};

template <int MODE>   // 1 = FILTER, 2 = symmetric FILTER
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
tc_cor2_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo, const TcParams prm)
{
    using Cfg = Tc2Cfg;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::NSTAGE * Cfg::STAGE_BYTES);
    uint64_t* full_bar = bars;                       // [NSTAGE] used in the leader
    uint64_t* empty_bar = bars + Cfg::NSTAGE;        // [NSTAGE]
    uint64_t* tfull_bar = bars + 2 * Cfg::NSTAGE;
    uint64_t* tempty_bar = bars + 2 * Cfg::NSTAGE + 1;   // used in the leader
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::NSTAGE + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int crank = (int)cluster_ctarank();
    const bool leader = crank == 0;
    const int m_units = (prm.m_tiles + 1) / 2;       // 256-row units
    const TileIter iter0(m_units, prm.n_tiles, MODE == 2 ? 1 : 0, 2 * BM);
    const int unit0 = (int)blockIdx.x / 2, unit_stride = (int)gridDim.x / 2;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_hi); prefetch_tensormap(&map_lo);
        for (int s = 0; s < Cfg::NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tfull_bar, 1); mbar_init(tempty_bar, 2 * EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // a 256 x 256 pair tile entirely below the diagonal holds no pair (i, j > i)
    auto skip = [&](int unit, int n_blk) { return MODE == 2 && (n_blk + 1) * BN <= unit * 2 * BM + 1; };

    if (warp < EPI_WARP0) {
    setmaxnreg_dec<REGS_CTRL>();
    if (warp == 0) {
        // ===== TMA producer (both CTAs) =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            TileIter iter = iter0;
            for (int l = unit0; l < prm.local_tiles; l += unit_stride) {
                const int tile = part_tile(prm, l);
                int unit, n_blk;
                iter.coords(tile, unit, n_blk);
                if (skip(unit, n_blk)) continue;
                const int32_t m0 = (int32_t)(prm.a_row0 + ((int64_t)unit * 2 + crank) * BM);
                const int32_t n0 = n_blk * BN + crank * (BN / 2);
                for (int kb = 0; kb < prm.num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* st = smem + stage * Cfg::STAGE_BYTES;
                    if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                    const uint32_t fb = mapa_u32(smem_u32(&full_bar[stage]), 0);       // the leader's barrier
                    tma_load_2d_2sm(st, &map_hi, fb, kb * BK, m0);
                    tma_load_2d_2sm(st + Cfg::A_BYTES, &map_lo, fb, kb * BK, m0);
                    tma_load_2d_2sm(st + 2 * Cfg::A_BYTES, &map_hi, fb, kb * BK, n0);
                    tma_load_2d_2sm(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, &map_lo, fb, kb * BK, n0);
                    if (++stage == Cfg::NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        if (lane == 0 && leader) {
            constexpr uint32_t idesc = make_idesc_f16(2 * BM, BN);
            int stage = 0; uint32_t phase = 0;
            uint32_t tphase = 0;
            const uint32_t d_main = tmem_base, d_cross = tmem_base + (uint32_t)BN;
            TileIter iter = iter0;
            for (int l = unit0; l < prm.local_tiles; l += unit_stride) {
                const int tile = part_tile(prm, l);
                int unit, n_blk;
                iter.coords(tile, unit, n_blk);
                if (skip(unit, n_blk)) continue;
                mbar_wait(tempty_bar, tphase ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < prm.num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint64_t a_hi = make_sw128_kmajor_desc(st);
                    const uint64_t a_lo = make_sw128_kmajor_desc(st + Cfg::A_BYTES);
                    const uint64_t b_hi = make_sw128_kmajor_desc(st + 2 * Cfg::A_BYTES);
                    const uint64_t b_lo = make_sw128_kmajor_desc(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES);
                    const int nk = kb == prm.num_kb - 1 ? prm.last_ksteps : BK / UMMA_K;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        if (k >= nk) break;
                        const uint64_t koff = (uint64_t)((k * UMMA_K * ELT) >> 4);
                        const uint32_t accum = (kb | k) != 0 ? 1u : 0u;
                        mma_f16_ss_2sm(d_main, a_hi + koff, b_hi + koff, idesc, accum);
                        mma_f16_ss_2sm(d_cross, a_hi + koff, b_lo + koff, idesc, accum);
                        mma_f16_ss_2sm(d_cross, a_lo + koff, b_hi + koff, idesc, 1u);
                    }
                    tc_commit_2sm_mc(&empty_bar[stage], (uint16_t)3);     // both CTAs may refill the stage
                    if (++stage == Cfg::NSTAGE) { stage = 0; phase ^= 1; }
                }
                tc_commit_2sm_mc(tfull_bar, (uint16_t)3);                 // both CTAs drain their 128 rows
                tphase ^= 1;
            }
        }
    }
    } else {
        setmaxnreg_inc<REGS_EPI>();
        // ===== epilogue warps 4..11 (both CTAs) =====
        const int quarter = warp & 3;
        const int half = (warp - EPI_WARP0) >> 2;
        uint32_t tphase = 0;
        TileIter iter = iter0;
        float* tcol = reinterpret_cast<float*>(tmem_slot + 4) + (warp - EPI_WARP0) * (BN / 2);
        uint32_t chunk_id = 0, fill = 0;
        bool have_chunk = false;
        const uint32_t tempty_leader = mapa_u32(smem_u32(tempty_bar), 0);
        for (int l = unit0; l < prm.local_tiles; l += unit_stride) {
                const int tile = part_tile(prm, l);
            int unit, n_blk;
            iter.coords(tile, unit, n_blk);
            if (skip(unit, n_blk)) continue;
            auto release = [&]() { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive_cluster(tempty_leader); };
            epilogue_filter_tile<MODE>(prm, tmem_base, quarter, half, lane, unit * 2 + crank, n_blk, tcol, tfull_bar, tphase,
                                       chunk_id, fill, have_chunk, release);
            tphase ^= 1;
        }
        if (have_chunk && lane == 0 && chunk_id < prm.pool_chunks) prm.chunk_fill[chunk_id] = fill;
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) { tc_fence_after(); tmem_dealloc_2sm<TMEM_COLS>(tmem_base); }
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

// columns [G, Gpad) of the planes are zero: the k-steps of the last k-block that lie wholly past G contribute nothing and
// are not issued (G = 1000: 63 k-steps of 16 columns instead of 64)
static int32_t tc_last_ksteps(int32_t Gpad, int32_t G)
{
    if (G <= 0 || G > Gpad || G <= Gpad - BK) return BK / UMMA_K;
    return std::max(1, (G - (Gpad - BK) + UMMA_K - 1) / UMMA_K);
}

template <int NPASS, int MODE, int CM>
static void launch_tc(mcb_ctx* ctx, const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh,
                      const CUtensorMap& bl, const TcParams& prm)
{
    using Cfg = TcCfg<NPASS>;
    static SmemOptIn optin;
    ensure_dyn_smem(optin, ctx, tc_cor_kernel<NPASS, MODE, CM>, (size_t)Cfg::SMEM_BYTES);
    TcParams p2 = prm;
    p2.dbg = (MODE != 0 && getenv("MCB_TC_DEBUG")) ? atoi(getenv("MCB_TC_DEBUG")) : 0;
    p2.total_tiles = TileIter((int)ceil_div(prm.m_tiles, CM), prm.n_tiles, MODE == 2 ? 1 : 0, CM * BM).total();
    if (p2.nparts < 1) { p2.nparts = 1; p2.part = 0; }
    p2.part_block = GROUP_M;
    p2.local_tiles = part_count(p2.total_tiles, p2.part_block, p2.nparts, p2.part);
    const int units = std::max(1, p2.local_tiles);
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

// 1-pass product with fp16 output (the sampled correlations of the threshold stage): ldc16 even, >= b_rows rounded up to 2
void launch_tc_store16(mcb_ctx* ctx, const void* Ahi, int64_t a_rows_alloc, const void* Bhi, int64_t b_rows_alloc,
                       int32_t Gpad, void* C16, int64_t ldc16, int64_t a_rows, int64_t b_rows)
{
    if (a_rows == 0 || b_rows == 0) return;
    tc_init();
    MCB_CHECK(Gpad % BK == 0, "Gpad must be a multiple of 64");
    MCB_CHECK(ldc16 % 2 == 0 && ldc16 >= ((b_rows + 1) & ~(int64_t)1), "tc store16: ldc must be even and cover the columns");
    const CUtensorMap ah = make_map(Ahi, a_rows_alloc, Gpad, BM);
    const CUtensorMap bh = make_map(Bhi, b_rows_alloc, Gpad, BN);
    TcParams prm{};
    prm.a_row0 = 0; prm.a_rows = a_rows; prm.b_rows = b_rows; prm.num_kb = Gpad / BK; prm.last_ksteps = BK / UMMA_K;
    prm.m_tiles = (int32_t)ceil_div(a_rows, BM); prm.n_tiles = (int32_t)ceil_div(b_rows, BN);
    prm.C16 = reinterpret_cast<__half*>(C16); prm.ldc16 = ldc16;
    launch_tc<1, 0, 1>(ctx, ah, ah, bh, bh, prm);
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
    prm.a_row0 = 0; prm.a_rows = a_rows; prm.b_rows = b_rows; prm.num_kb = Gpad / BK; prm.last_ksteps = BK / UMMA_K;
    prm.m_tiles = (int32_t)ceil_div(a_rows, BM); prm.n_tiles = (int32_t)ceil_div(b_rows, BN);
    prm.C = C; prm.ldc = ldc;
    launch_tc<1, 0, 1>(ctx, ah, ah, bh, bh, prm);
}

template <int MODE>
static void launch_tc2(mcb_ctx* ctx, const CUtensorMap& mh, const CUtensorMap& ml, const TcParams& prm)
{
    static SmemOptIn optin;
    ensure_dyn_smem(optin, ctx, tc_cor2_kernel<MODE>, (size_t)Tc2Cfg::SMEM_BYTES);
    TcParams p2 = prm;
    p2.dbg = getenv("MCB_TC_DEBUG") ? atoi(getenv("MCB_TC_DEBUG")) : 0;
    p2.total_tiles = TileIter((int)ceil_div(prm.m_tiles, 2), prm.n_tiles, MODE == 2 ? 1 : 0, 2 * BM).total();
    if (p2.nparts < 1) { p2.nparts = 1; p2.part = 0; }
    p2.part_block = GROUP_M;
    p2.local_tiles = part_count(p2.total_tiles, p2.part_block, p2.nparts, p2.part);
    const int pairs = std::max(1, std::min(p2.local_tiles, ctx->num_sms / 2));
    tc_cor2_kernel<MODE><<<2 * pairs, TC_THREADS, Tc2Cfg::SMEM_BYTES, ctx->stream>>>(mh, ml, p2);   // __cluster_dims__(2,1,1)
    MCB_LAUNCH_CHECK();
}

uint32_t tc_pool_chunk_records() { return POOL_CHUNK; }
uint32_t tc_pool_min_chunks(mcb_ctx* ctx) { return (uint32_t)(EPI_WARPS * ctx->num_sms + 64); }   // one open chunk per warp

void launch_tc_filter(mcb_ctx* ctx, const void* Zhi, const void* Zlo, int64_t rows_alloc, int64_t a_row0,
                      int64_t a_rows, int64_t b_rows, int32_t Gpad, const float* thr, const CandPool& pool, int32_t part,
                      int32_t nparts, int32_t G)
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
    prm.a_row0 = a_row0; prm.a_rows = a_rows; prm.b_rows = b_rows; prm.num_kb = Gpad / BK; prm.last_ksteps = tc_last_ksteps(Gpad, G);
    prm.m_tiles = (int32_t)ceil_div(a_rows, BM); prm.n_tiles = (int32_t)ceil_div(b_rows, BN);
    prm.thr = thr;
    prm.pool = reinterpret_cast<uint4*>(pool.records); prm.pool_cursor = pool.cursor; prm.chunk_fill = pool.chunk_fill;
    prm.pool_chunks = pool.n_chunks; prm.overflow = pool.overflow;
    prm.part = part; prm.nparts = nparts;
    const bool sym = a_row0 == 0 && a_rows == b_rows;   // whole square on one rank: each unordered pair contracted once
    if (ctx->gemm_impl == 2) {                          // 1-SM kernel (A/B measurements)
        if (sym) launch_tc<3, 2, FILTER_CM>(ctx, ah, al, bh, bl, prm);
        else launch_tc<3, 1, FILTER_CM>(ctx, ah, al, bh, bl, prm);
        return;
    }
    if (sym) launch_tc2<2>(ctx, ah, al, prm);
    else launch_tc2<1>(ctx, ah, al, prm);
}

// 3-pass product of separate A and B operands stored to C (exact rows of the kNN stage)
void launch_tc_store3_ab(mcb_ctx* ctx, const void* Ahi, const void* Alo, int64_t a_rows_alloc, const void* Bhi, const void* Blo,
                         int64_t b_rows_alloc, int64_t a_rows, int64_t b_rows, int32_t Gpad, float* C, int64_t ldc)
{
    if (a_rows == 0 || b_rows == 0) return;
    tc_init();
    MCB_CHECK(ldc % 4 == 0, "tc store: ldc must be a multiple of 4");
    const CUtensorMap ah = make_map(Ahi, a_rows_alloc, Gpad, BM);
    const CUtensorMap al = make_map(Alo, a_rows_alloc, Gpad, BM);
    const CUtensorMap bh = make_map(Bhi, b_rows_alloc, Gpad, BN);
    const CUtensorMap bl = make_map(Blo, b_rows_alloc, Gpad, BN);
    TcParams prm{};
    prm.a_row0 = 0; prm.a_rows = a_rows; prm.b_rows = b_rows; prm.num_kb = Gpad / BK; prm.last_ksteps = BK / UMMA_K;
    prm.m_tiles = (int32_t)ceil_div(a_rows, BM); prm.n_tiles = (int32_t)ceil_div(b_rows, BN);
    prm.C = C; prm.ldc = ldc;
    launch_tc<3, 0, 1>(ctx, ah, al, bh, bl, prm);
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
    prm.a_row0 = 0; prm.a_rows = a_rows; prm.b_rows = b_rows; prm.num_kb = Gpad / BK; prm.last_ksteps = BK / UMMA_K;
    prm.m_tiles = (int32_t)ceil_div(a_rows, BM); prm.n_tiles = (int32_t)ceil_div(b_rows, BN);
    prm.C = C; prm.ldc = ldc;
    launch_tc<3, 0, 1>(ctx, ah, al, bh, bl, prm);
}

}  // namespace mcb
