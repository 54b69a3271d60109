// corsample.cu -- per-row filter thresholds of the kNN stage (tgs_cor_knn, reference call site R/utils.r:119) from a
// random column sample, WITHOUT writing the sampled correlations to HBM.
//
// The filter kernel (corknn_tc.cu) needs, for every row i, a threshold that keeps ~1.5 * Kc of the M correlations of
// the row.  It is the r-th largest correlation of row i against S sampled columns (api.cu::sample_rule).  The first
// implementation stored the rows x S block in fp16 and binned it in a second kernel (2.5 GB written and read back at
// C2: 3.0 + 1.0 ms, the GEMM bound by the L2 -> SM operand feed, LTS cap ~6,300 B / clk).  This kernel fuses the
// selection into the epilogue and shrinks the operands:
//
//   * operands: the fp16 hi plane of z * 2^12 (one pass, kind::f16; ~5e-4 worst-case error, covered by eps), read in
//     place -- the TMA maps treat the plane as a byte matrix with 128-byte k-blocks.  An FP8 (e4m3) variant of the same
//     kernel (kind::f8f6f4, twice the rate, half the bytes; F8 = true, context option sample_impl = 2) was built and
//     measured: 1.45 ms instead of ~2 ms at C2, but its rounding errors are STRUCTURED -- most genes of a cell share one
//     value (the zero-count z), so that value's e4m3 rounding error (up to 6 %) is common to ~70 % of the cell's entries
//     and shifts all of the cell's correlations together by up to ~5e-3: 12 % of the rows left their candidate window.
//     It is kept for the record only;
//   * a cluster of two CTAs (cta_group::2) contracts 256 rows x 256 sampled columns per tile, each CTA loading its own
//     128 A rows and half of the B tile; accumulators are double buffered in TMEM (2 x 256 columns), so the epilogue of
//     tile t overlaps the MMAs of tile t + 1;
//   * a CTA pair keeps ONE 256-row unit for all column tiles, so the per-row histograms live in shared memory
//     ([bin][row] u32: a warp's 32 rows hit 32 banks, no conflicts, fire-and-forget shared atomics) and never touch HBM;
//   * two phases per unit: the first n_tiles1 column tiles fill a COARSE histogram over [-0.5, 1] (bins of 1.5 / 128);
//     from it every row brackets its quantile, [lo, hi) = coarse bin edges around sample ranks c_lo / c_hi; the remaining
//     tiles fill a FINE histogram of 127 bins over the bracket (+ one bin for >= hi), i.e. ~3e-4 resolution; the r-th
//     largest is then interpolated inside its fine bin.  Rows whose quantile escapes the bracket (probability ~1e-5 by
//     the choice of c_lo / c_hi) get thr = +inf: they collect no candidates and are redone exactly by the kNN stage.
#include "internal.h"
#include "ptx.cuh"
#include <cuda_fp16.h>
#include <cuda_fp8.h>

namespace mcb {

using namespace ptx;

constexpr int SBM = 128;             // rows per CTA (TMEM lanes); a pair holds 256
constexpr int SBN = 256;             // sampled columns per tile
constexpr int SBK = 128;             // BYTES per k-block: one 128-byte swizzle row (64 fp16 or 128 e4m3 elements)
constexpr int S_UMMA_KB = 32;        // bytes per UMMA k-step (K = 16 fp16 / 32 e4m3)
constexpr int S_THREADS = 384;       // warp 0 TMA, warp 1 MMA, warps 4..11 epilogue
constexpr int S_EPI_WARP0 = 4;
constexpr int S_EPI_WARPS = 8;
constexpr int S_NB = 128;            // histogram bins per row
constexpr int S_NSTAGE = 4;
constexpr int S_A_BYTES = SBM * SBK;                 // 16 KB
constexpr int S_B_BYTES = (SBN / 2) * SBK;           // 16 KB: this CTA's half of the B tile
constexpr int S_STAGE_BYTES = S_A_BYTES + S_B_BYTES;
constexpr int S_HIST_WORDS = S_NB * SBM;
constexpr int S_SMEM_BYTES = S_NSTAGE * S_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + SBM * 16 /*row params*/ + S_HIST_WORDS * 4;
constexpr float S_CLO = -0.5f;                       // coarse range [S_CLO, S_CLO + S_NB * S_CW) = [-0.5, 1.0)
constexpr float S_CW = 1.5f / S_NB;
constexpr float Z8_SCALE = 256.0f;                   // e4m3 planes hold z * 2^8 (|z| <= 1 -> <= 256 < 448)
constexpr float ACC8_TO_COR = 1.0f / (Z8_SCALE * Z8_SCALE);
constexpr float ACC16_TO_COR = C_RESCALE;            // fp16 planes hold z * 2^12

struct SampParams {
    int64_t a_rows;
    int32_t num_kb, last_ksteps, n_units, n_tiles, n_tiles1, S, r_sel, c_hi, c_lo;
    float eps;
    float* thr;
    uint32_t* n_escaped;        // rows whose quantile left the bracket (statistics)
};

template <bool F8>
__device__ __forceinline__ void mma_ss_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate);
template <> __device__ __forceinline__ void mma_ss_2sm<false>(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    mma_f16_ss_2sm(d_tmem, a_desc, b_desc, idesc, accumulate);
}
template <> __device__ __forceinline__ void mma_ss_2sm<true>(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
        "}\n" :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// kind::f8f6f4 instruction descriptor: D = F32, A = B = E4M3 (format code 0), both K-major
__host__ __device__ constexpr uint32_t make_idesc_e4m3(uint32_t M, uint32_t N)
{
    return (1u << 4) | (0u << 7) | (0u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// hist[min(floor(t), S_NB - 1)][row] += 1 when t >= 0 (and ok): predicated fire-and-forget shared-memory reduction, no branch.
// hrow_s = shared-window address of hist[0][row]
__device__ __forceinline__ void hist_add(uint32_t hrow_s, float t, bool ok = true)
{
    const int b = min(__float2int_rd(t), S_NB - 1);
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "setp.ne.b32 q, %3, 0;\n\t"
        "setp.ge.and.f32 p, %2, 0f00000000, q;\n\t"
        "@p red.shared.add.u32 [%0], %1;\n\t"
        "}\n" :: "r"(hrow_s + (uint32_t)b * (uint32_t)(SBM * 4)), "r"(1u), "f"(t), "r"((int)ok) : "memory");
}

template <bool F8>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(S_THREADS, 1)
tc_sample_hist_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const SampParams prm)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S_NSTAGE * S_STAGE_BYTES);
    uint64_t* full_bar = bars;                          // [NSTAGE] used in the leader
    uint64_t* empty_bar = bars + S_NSTAGE;              // [NSTAGE]
    uint64_t* tfull_bar = bars + 2 * S_NSTAGE;          // [2]
    uint64_t* tempty_bar = bars + 2 * S_NSTAGE + 2;     // [2] used in the leader
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S_NSTAGE + 4);
    float4* rowp = reinterpret_cast<float4*>(smem + S_NSTAGE * S_STAGE_BYTES + 256);      // [SBM] {scale, offset, lo, width}
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem + S_NSTAGE * S_STAGE_BYTES + 256 + SBM * 16);   // [S_NB][SBM]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int crank = (int)cluster_ctarank();
    const bool leader = crank == 0;
    const int unit0 = (int)blockIdx.x / 2, unit_stride = (int)gridDim.x / 2;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_a); prefetch_tensormap(&map_b);
        for (int s = 0; s < S_NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 2 * S_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < S_EPI_WARP0) {
        if (warp == 0) {
            // ===== TMA producer (both CTAs): own 128 A rows + own half of the B tile per k-block =====
            if (lane == 0) {
                int stage = 0; uint32_t phase = 0;
                for (int unit = unit0; unit < prm.n_units; unit += unit_stride) {
                    const int32_t m0 = (unit * 2 + crank) * SBM;
                    for (int n = 0; n < prm.n_tiles; ++n) {
                        const int32_t n0 = n * SBN + crank * (SBN / 2);
                        for (int kb = 0; kb < prm.num_kb; ++kb) {
                            mbar_wait(&empty_bar[stage], phase ^ 1);
                            uint8_t* st = smem + stage * S_STAGE_BYTES;
                            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * S_STAGE_BYTES);
                            const uint32_t fb = mapa_u32(smem_u32(&full_bar[stage]), 0);       // the leader's barrier
                            tma_load_2d_2sm(st, &map_a, fb, kb * SBK, m0);
                            tma_load_2d_2sm(st + S_A_BYTES, &map_b, fb, kb * SBK, n0);
                            if (++stage == S_NSTAGE) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        } else if (warp == 1) {
            // ===== MMA issuer (leader CTA only) =====
            if (lane == 0 && leader) {
                constexpr uint32_t idesc = F8 ? make_idesc_e4m3(2 * SBM, SBN) : make_idesc_f16(2 * SBM, SBN);
                int stage = 0; uint32_t phase = 0;
                int it = 0;
                for (int unit = unit0; unit < prm.n_units; unit += unit_stride) {
                    for (int n = 0; n < prm.n_tiles; ++n, ++it) {
                        const int acc = it & 1;
                        const uint32_t tphase = (uint32_t)(it >> 1) & 1u;
                        const uint32_t d = tmem_base + (uint32_t)(acc * SBN);
                        mbar_wait(&tempty_bar[acc], tphase ^ 1);
                        tc_fence_after();
                        for (int kb = 0; kb < prm.num_kb; ++kb) {
                            mbar_wait(&full_bar[stage], phase);
                            tc_fence_after();
                            const uint32_t st = smem_u32(smem + stage * S_STAGE_BYTES);
                            const uint64_t a = make_sw128_kmajor_desc(st);
                            const uint64_t b = make_sw128_kmajor_desc(st + S_A_BYTES);
                            const int nk = kb == prm.num_kb - 1 ? prm.last_ksteps : SBK / S_UMMA_KB;      // trailing k-steps of zero columns are skipped
#pragma unroll
                            for (int k = 0; k < SBK / S_UMMA_KB; ++k) {
                                if (k >= nk) break;
                                const uint64_t koff = (uint64_t)((k * S_UMMA_KB) >> 4);      // 32 bytes per k step
                                mma_ss_2sm<F8>(d, a + koff, b + koff, idesc, (kb | k) != 0 ? 1u : 0u);
                            }
                            tc_commit_2sm_mc(&empty_bar[stage], (uint16_t)3);
                            if (++stage == S_NSTAGE) { stage = 0; phase ^= 1; }
                        }
                        tc_commit_2sm_mc(&tfull_bar[acc], (uint16_t)3);
                    }
                }
            }
        }
    } else {
        // ===== epilogue warps 4..11 (both CTAs): TMEM lanes (warp % 4) * 32 .. + 31 = rows of the CTA, column half (warp - 4) / 4 =====
        const int quarter = warp & 3;
        const int half = (warp - S_EPI_WARP0) >> 2;
        const int row = quarter * 32 + lane;                 // row inside the CTA's 128
        const int etid = (int)threadIdx.x - S_EPI_WARP0 * 32;         // 0..255
        const uint32_t tempty_leader0 = mapa_u32(smem_u32(&tempty_bar[0]), 0);
        const uint32_t tempty_leader1 = mapa_u32(smem_u32(&tempty_bar[1]), 0);
        uint32_t* hrow = hist + row;
        const uint32_t hrow_s = smem_u32(hrow);
        int it = 0;
        for (int q = etid; q < S_HIST_WORDS; q += 256) hist[q] = 0u;
        epi_bar_sync();
        for (int unit = unit0; unit < prm.n_units; unit += unit_stride) {
            const int64_t grow = ((int64_t)unit * 2 + crank) * SBM + row;      // row of A
            const bool row_ok = grow < prm.a_rows;
            // coarse phase parameters (accumulator units: c * 2^16 for e4m3, c * 2^24 for fp16 operands)
            constexpr float ACC_TO_COR = F8 ? ACC8_TO_COR : ACC16_TO_COR;
            float bs = ACC_TO_COR / S_CW, bo = -S_CLO / S_CW;
            for (int n = 0; n < prm.n_tiles; ++n, ++it) {
                if (n == prm.n_tiles1) {
                    // ---- bracket of every row from the coarse histogram, then restart the histogram for the fine phase
                    epi_bar_sync();
                    if (half == 0) {
                        uint32_t cum = 0;
                        int b_hi = S_NB, b_lo = 0;
                        for (int b = S_NB - 1; b >= 0; --b) {
                            cum += hrow[b * SBM];
                            if (cum <= (uint32_t)prm.c_hi) b_hi = b;
                            if (cum >= (uint32_t)prm.c_lo) { b_lo = b; break; }
                        }
                        const float lo = S_CLO + (float)b_lo * S_CW, hi = S_CLO + (float)b_hi * S_CW;
                        const float width = (hi - lo) / (float)(S_NB - 1);
                        rowp[row] = make_float4(ACC_TO_COR / width, -lo / width, lo, width);
                    }
                    epi_bar_sync();
                    for (int q = etid; q < S_HIST_WORDS; q += 256) hist[q] = 0u;
                    epi_bar_sync();
                    const float4 p = rowp[row];
                    bs = p.x; bo = p.y;
                }
                const int acc = it & 1;
                const uint32_t tphase = (uint32_t)(it >> 1) & 1u;
                mbar_wait(&tfull_bar[acc], tphase);
                tc_fence_after();
                const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * SBN + half * (SBN / 2));
                const int col_base = n * SBN + half * (SBN / 2);
#pragma unroll 1
                for (int c = 0; c < SBN / 64; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32(taddr0 + (uint32_t)(c * 32), v);
                    tmem_ld_wait();
                    const int col0 = col_base + c * 32;
                    if (!row_ok || col0 >= prm.S) continue;
                    if (col0 + 32 <= prm.S) {
#pragma unroll
                        for (int q = 0; q < 32; ++q) hist_add(hrow_s, fmaf(__uint_as_float(v[q]), bs, bo));
                    } else {
#pragma unroll
                        for (int q = 0; q < 32; ++q) hist_add(hrow_s, fmaf(__uint_as_float(v[q]), bs, bo), col0 + q < prm.S);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(acc ? tempty_leader1 : tempty_leader0);
            }
            // ---- r-th largest of the fine sample, interpolated inside its bin
            epi_bar_sync();
            if (half == 0 && row_ok) {
                const float4 p = rowp[row];
                uint32_t cum = hrow[(S_NB - 1) * SBM];           // >= hi
                float thr = __int_as_float(0x7f800000);             // +inf: the row is redone exactly
                if (cum < (uint32_t)prm.r_sel) {
                    for (int b = S_NB - 2; b >= 0; --b) {
                        const uint32_t cnt = hrow[b * SBM];
                        if (cum + cnt >= (uint32_t)prm.r_sel) {
                            const float pos = (float)((uint32_t)prm.r_sel - cum);          // 1 .. cnt from the top of the bin
                            thr = p.z + ((float)(b + 1) - (pos - 0.5f) / (float)cnt) * p.w - prm.eps;
                            break;
                        }
                        cum += cnt;
                    }
                }
                if (thr == __int_as_float(0x7f800000) && prm.n_escaped) atomicAdd(prm.n_escaped, 1u);
                prm.thr[grow] = thr;
            }
            epi_bar_sync();
            for (int q = etid; q < S_HIST_WORDS; q += 256) hist[q] = 0u;
            epi_bar_sync();
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) { tc_fence_after(); tmem_dealloc_2sm<512>(tmem_base); }
}

// ---- e4m3 copies of the hi plane ------------------------------------------------------------------------------------
// out[r][0..G8pad) = e4m3(Zh[src(r)][g] * 2^-4) (Zh holds z * 2^12), zero past Gpad; src(r) = rows ? rows[r] : row0 + r;
// a negative index gives a zero row (padding of the sample)
__global__ void z8_rows_kernel(const __half* __restrict__ Zh, int32_t Gpad, const int32_t* __restrict__ rows, int64_t row0, int64_t nrows,
                               int32_t G8pad, uint8_t* __restrict__ out)
{
    const int vec = G8pad / 8;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < nrows * vec; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = idx / vec;
        const int g0 = (int)(idx % vec) * 8;
        const int64_t src = rows ? (int64_t)rows[r] : row0 + r;
        uint2 o = make_uint2(0u, 0u);
        if (src >= 0 && g0 < Gpad) {            // Gpad is a multiple of 64: an 8-element vector never straddles it
            const uint4 h = *reinterpret_cast<const uint4*>(Zh + src * Gpad + g0);
            const uint32_t w[4] = {h.x, h.y, h.z, h.w};
            uint32_t f8[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
                f.x *= Z8_SCALE / Z_PRESCALE; f.y *= Z8_SCALE / Z_PRESCALE;
                f8[k] = (uint32_t)__nv_cvt_float2_to_fp8x2(f, __NV_SATFINITE, __NV_E4M3);
            }
            o.x = f8[0] | (f8[1] << 16);
            o.y = f8[2] | (f8[3] << 16);
        }
        *reinterpret_cast<uint2*>(out + r * G8pad + g0) = o;
    }
}
void launch_z8_rows(mcb_ctx* ctx, const void* Zh, int32_t Gpad, const int32_t* rows, int64_t row0, int64_t nrows, int32_t G8pad, void* out)
{
    if (nrows == 0) return;
    MCB_CHECK(Gpad % 64 == 0 && G8pad % 128 == 0 && G8pad >= Gpad, "z8 rows: bad padding");
    const int grid = (int)std::min<int64_t>(ceil_div(nrows * (G8pad / 8), 256), (int64_t)ctx->num_sms * 16);
    z8_rows_kernel<<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<const __half*>(Zh), Gpad, rows, row0, nrows, G8pad,
                                                  reinterpret_cast<uint8_t*>(out));
    MCB_LAUNCH_CHECK();
}

// ---- host side ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        MCB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres));
        MCB_CHECK(f != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available in this driver");
        fn = reinterpret_cast<EncodeTiledFn>(f);
    }
    return fn;
}
// 2D byte tensor [rows][row_bytes] row-major; box = 128 bytes x 128 rows; 128-byte swizzle; rows past the end read as zero
static CUtensorMap make_map_u8(const void* base, int64_t rows, int32_t G8pad)
{
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)G8pad, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)G8pad};
    cuuint32_t box[2] = {(cuuint32_t)SBK, (cuuint32_t)SBM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char b[160];
        snprintf(b, sizeof(b), "cuTensorMapEncodeTiled (u8) failed with CUresult %d (rows=%lld G8pad=%d)", (int)r, (long long)rows, G8pad);
        throw Error(2, b);
    }
    return m;
}

// thr[0..a_rows) from the operand rows A [a_rows][row_bytes] against the sample B [Spad][row_bytes] (Spad = S rounded up
// to 256, padding rows zero); f8 = 0: fp16 rows of z * 2^12 (the hi plane itself), 1: e4m3 rows of z * 2^8.  The first
// n_tiles1 * 256 sampled columns are the coarse phase; r_sel is the rank to locate among the remaining S - n_tiles1 * 256.
void launch_sample_thresholds_fused(mcb_ctx* ctx, int f8, const void* A, int64_t a_rows, const void* B, int64_t Spad, int32_t row_bytes,
                                    int32_t S, int32_t n_tiles1, int32_t r_sel, int32_t c_hi, int32_t c_lo, float eps, float* thr,
                                    uint32_t* n_escaped, int32_t valid_bytes)
{
    if (a_rows == 0) return;
    MCB_CHECK(row_bytes % SBK == 0 && Spad % SBN == 0 && S <= Spad, "sample thresholds: bad padding");
    MCB_CHECK(n_tiles1 >= 1 && (int64_t)n_tiles1 * SBN < S && c_lo > c_hi && c_hi >= 0 && r_sel >= 1, "sample thresholds: bad phase split");
    const CUtensorMap ma = make_map_u8(A, a_rows, row_bytes);
    const CUtensorMap mb = make_map_u8(B, Spad, row_bytes);
    SampParams prm{};
    prm.a_rows = a_rows; prm.num_kb = row_bytes / SBK; prm.n_units = (int32_t)ceil_div(a_rows, 2 * SBM);
    prm.last_ksteps = SBK / S_UMMA_KB;
    if (valid_bytes > row_bytes - SBK && valid_bytes <= row_bytes) prm.last_ksteps = std::max(1, (valid_bytes - (row_bytes - SBK) + S_UMMA_KB - 1) / S_UMMA_KB);
    prm.n_tiles = (int32_t)(Spad / SBN); prm.n_tiles1 = n_tiles1; prm.S = S; prm.r_sel = r_sel; prm.c_hi = c_hi; prm.c_lo = c_lo;
    prm.eps = eps; prm.thr = thr; prm.n_escaped = n_escaped;
    const int pairs = std::max(1, std::min(prm.n_units, ctx->num_sms / 2));
    if (f8) {
        static SmemOptIn optin;
        ensure_dyn_smem(optin, ctx, tc_sample_hist_kernel<true>, (size_t)S_SMEM_BYTES);
        tc_sample_hist_kernel<true><<<2 * pairs, S_THREADS, S_SMEM_BYTES, ctx->stream>>>(ma, mb, prm);      // __cluster_dims__(2,1,1)
    } else {
        static SmemOptIn optin;
        ensure_dyn_smem(optin, ctx, tc_sample_hist_kernel<false>, (size_t)S_SMEM_BYTES);
        tc_sample_hist_kernel<false><<<2 * pairs, S_THREADS, S_SMEM_BYTES, ctx->stream>>>(ma, mb, prm);
    }
    MCB_LAUNCH_CHECK();
}

}  // namespace mcb
