// features.cu -- gset_get_feat_mat row subset (reference R/gset.r:192-197) + the bknn transform
// log2(1 + k_nonz_exp * u) (R/cgraph_knn.r:13-14, k = scm_k_nonz_exp = 7) + per-cell
// standardisation, so that Pearson cor(i, j) (what tgs_cor_knn computes, R/utils.r:119) becomes a
// plain dot product of rows of Z.
//
// Layout in HBM: Z is [Mpad][Gpad] row-major (one row per *valid* cell, K-major for both GEMM
// operands), stored three times: Zf = fp32(z) for the CUDA-core kernels, and two FP16 planes
// Zh + Zl = z * 2^12 (hi = fp16(z*S), lo = fp16(z*S - hi)) for the 3-pass split contraction on the
// tensor cores.  Gpad is G rounded up to 64 (one 128-byte swizzle row of fp16 per k-block); padded
// columns and padded rows are zero.  Cells whose feature vector has zero variance (NA correlations in R,
// R/cgraph_knn_mcnorm.r:25) are dropped here: rowpos compacts the valid cells.
//
// HBM-bound sparse-to-dense work: read nnz * 8 B twice, write M * Gpad * 8 B once, coalesced.
// The fp16 prescale 2^12 keeps every |z| >= 2^-15 in the normal fp16 range for the lo plane.
#include "internal.h"
#include <cuda_fp16.h>

namespace mcb {

constexpr uint32_t FEAT_LUT = 256;

// one warp per cell: sum / sum of squares of the transformed feature entries in fp64
__global__ void feat_stats_kernel(int64_t ncells, const int64_t* __restrict__ p, const int32_t* __restrict__ ridx,
                                  const uint32_t* __restrict__ x, const int32_t* __restrict__ feat_map, int32_t G,
                                  double k_nonz, double* __restrict__ mu_out, double* __restrict__ inv_out,
                                  int32_t* __restrict__ valid)
{
    // UMI counts are small integers: the transform of u < 256 comes from a table built with the same expression
    // (bit-identical), which removes three fp64 log2 evaluations per stored entry from the feature stage
    __shared__ double lut[FEAT_LUT];
    for (int u = threadIdx.x; u < FEAT_LUT; u += blockDim.x) lut[u] = log2(1.0 + k_nonz * (double)u);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp; c < ncells; c += nwarps) {
        double s = 0.0;
        int cnt = 0;
        for (int64_t e = p[c] + lane; e < p[c + 1]; e += 32) {
            const uint32_t u = x[e];
            if (u == 0) continue;
            if (feat_map[ridx[e]] < 0) continue;
            s += u < FEAT_LUT ? lut[u] : log2(1.0 + k_nonz * (double)u);
            ++cnt;
        }
        s = warp_reduce_sum(s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        const double mu = s / (double)G;
        double s2 = 0.0;                     // second pass (entries are L1/L2 hot): sum (v - mu)^2
        for (int64_t e = p[c] + lane; e < p[c + 1]; e += 32) {
            const uint32_t u = x[e];
            if (u == 0) continue;
            if (feat_map[ridx[e]] < 0) continue;
            const double dlt = (u < FEAT_LUT ? lut[u] : log2(1.0 + k_nonz * (double)u)) - mu;
            s2 += dlt * dlt;
        }
        s2 = warp_reduce_sum(s2);
        if (lane == 0) {
            // the (G - cnt) zero entries each contribute mu^2
            const double ss = s2 + (double)(G - cnt) * mu * mu;
            const bool ok = cnt > 0 && ss >= 1e-20;
            mu_out[c] = mu;
            inv_out[c] = ok ? 1.0 / sqrt(ss) : 0.0;
            valid[c] = ok ? 1 : 0;
        }
    }
}

// single-block exclusive scan (n up to a few million).  Each thread owns SCAN_VPT consecutive elements per sweep, so a
// 100,000-element scan is 25 sweeps of three barriers instead of 98.
constexpr int SCAN_VPT = 4;
template <typename T>
__global__ void __launch_bounds__(1024) excl_scan_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t n,
                                                          T* __restrict__ total)
{
    __shared__ T warp_tot[32];
    __shared__ T carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int64_t base = 0; base < n; base += 1024 * SCAN_VPT) {
        const int64_t i0 = base + (int64_t)threadIdx.x * SCAN_VPT;
        T v[SCAN_VPT];
        T mine = 0;
#pragma unroll
        for (int q = 0; q < SCAN_VPT; ++q) { v[q] = i0 + q < n ? in[i0 + q] : (T)0; mine += v[q]; }
        T inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) warp_tot[w] = inc;
        __syncthreads();
        T wbase = 0, all = 0;
        for (int q = 0; q < 32; ++q) { T t = warp_tot[q]; if (q < w) wbase += t; all += t; }
        const T carry = carry_s;
        T run = carry + wbase + inc - mine;
#pragma unroll
        for (int q = 0; q < SCAN_VPT; ++q) { if (i0 + q < n) out[i0 + q] = run; run += v[q]; }
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + all;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry_s;
}

// z -> (hi, lo) fp16 planes with the 2^12 prescale
__device__ __forceinline__ void split_store(float z, __half* hi, __half* lo)
{
    const float zs = z * Z_PRESCALE;
    const __half h = __float2half_rn(zs);
    *hi = h;
    *lo = __float2half_rn(zs - __half2float(h));
}

// one CTA per cell: build the standardised row in shared memory, write the three planes coalesced
constexpr int FR_THREADS = 128;
__global__ void __launch_bounds__(FR_THREADS)
feat_rows_kernel(int64_t ncells, const int64_t* __restrict__ p, const int32_t* __restrict__ ridx,
                 const uint32_t* __restrict__ x, const int32_t* __restrict__ feat_map, int32_t G, int32_t Gpad,
                 double k_nonz, const double* __restrict__ mu_in, const double* __restrict__ inv_in,
                 const int32_t* __restrict__ valid, const int32_t* __restrict__ rowpos, float* __restrict__ Zf,
                 __half* __restrict__ Zh, __half* __restrict__ Zl, int32_t* __restrict__ cell_of_row, int64_t cell_base)
{
    extern __shared__ float srow[];        // [Gpad]
    __shared__ double lut[FEAT_LUT];
    for (int u = threadIdx.x; u < FEAT_LUT; u += FR_THREADS) lut[u] = log2(1.0 + k_nonz * (double)u);
    __syncthreads();
    for (int64_t c = blockIdx.x; c < ncells; c += gridDim.x) {
        if (!valid[c]) continue;           // uniform per block
        const double mu = mu_in[c], inv = inv_in[c];
        const float base = (float)((0.0 - mu) * inv);
        for (int g = threadIdx.x; g < Gpad; g += FR_THREADS) srow[g] = g < G ? base : 0.0f;
        __syncthreads();
        for (int64_t e = p[c] + threadIdx.x; e < p[c + 1]; e += FR_THREADS) {
            const uint32_t u = x[e];
            if (u == 0) continue;
            const int32_t f = feat_map[ridx[e]];
            if (f < 0) continue;
            const double v = u < FEAT_LUT ? lut[u] : log2(1.0 + k_nonz * (double)u);
            srow[f] = (float)((v - mu) * inv);
        }
        __syncthreads();
        const int64_t r = rowpos[c];
        for (int g = threadIdx.x; g < Gpad; g += FR_THREADS) {
            const float z = srow[g];
            if (Zf) Zf[r * (int64_t)Gpad + g] = z;          // the fp32 plane only exists on single-rank builds
            split_store(z, Zh + r * (int64_t)Gpad + g, Zl + r * (int64_t)Gpad + g);
        }
        if (threadIdx.x == 0) cell_of_row[r] = (int32_t)(cell_base + c);
        __syncthreads();
    }
}

// out[q][:] = Z[rows[q]][:] for rows of row_bytes bytes (multiple of 16); rows[q] < 0 -> zeros
__global__ void gather_rows_kernel(const uint4* __restrict__ Z, int32_t row_vecs, const int32_t* __restrict__ rows,
                                   int32_t nrows, uint4* __restrict__ out)
{
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (int64_t)nrows * row_vecs;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = idx / row_vecs;
        const int v = (int)(idx % row_vecs);
        const int32_t r = rows[q];
        uint4 val = make_uint4(0u, 0u, 0u, 0u);
        if (r >= 0) val = Z[(int64_t)r * row_vecs + v];
        out[q * (int64_t)row_vecs + v] = val;
    }
}

// ---- dense input (tgs_cor_graph(x = feat, ...), reference R/utils.r:115): x is the R matrix
// genes x cells, column-major, i.e. one contiguous fp64 vector of G values per cell -------------
__global__ void __launch_bounds__(128)
dense_stats_kernel(int64_t ncells, const double* __restrict__ x, int32_t G, double* __restrict__ mu_out,
                   double* __restrict__ inv_out, int32_t* __restrict__ valid)
{
    __shared__ double red[4];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int64_t c = blockIdx.x; c < ncells; c += gridDim.x) {
        const double* v = x + c * (int64_t)G;
        double s = 0.0;
        for (int g = threadIdx.x; g < G; g += 128) s += v[g];
        s = warp_reduce_sum(s);
        if (lane == 0) red[w] = s;
        __syncthreads();
        const double mu = (red[0] + red[1] + red[2] + red[3]) / (double)G;
        __syncthreads();
        double s2 = 0.0;
        for (int g = threadIdx.x; g < G; g += 128) { const double d = v[g] - mu; s2 += d * d; }
        s2 = warp_reduce_sum(s2);
        if (lane == 0) red[w] = s2;
        __syncthreads();
        if (threadIdx.x == 0) {
            const double ss = red[0] + red[1] + red[2] + red[3];
            const bool ok = ss >= 1e-20;      // also false for NaN
            mu_out[c] = mu; inv_out[c] = ok ? 1.0 / sqrt(ss) : 0.0; valid[c] = ok ? 1 : 0;
        }
        __syncthreads();
    }
}
__global__ void __launch_bounds__(128)
dense_rows_kernel(int64_t ncells, const double* __restrict__ x, int32_t G, int32_t Gpad, const double* __restrict__ mu_in,
                  const double* __restrict__ inv_in, const int32_t* __restrict__ valid, const int32_t* __restrict__ rowpos,
                  float* __restrict__ Zf, __half* __restrict__ Zh, __half* __restrict__ Zl, int32_t* __restrict__ cell_of_row)
{
    for (int64_t c = blockIdx.x; c < ncells; c += gridDim.x) {
        if (!valid[c]) continue;
        const double mu = mu_in[c], inv = inv_in[c];
        const double* v = x + c * (int64_t)G;
        const int64_t r = rowpos[c];
        for (int g = threadIdx.x; g < Gpad; g += 128) {
            const float z = g < G ? (float)((v[g] - mu) * inv) : 0.0f;
            Zf[r * (int64_t)Gpad + g] = z;
            split_store(z, Zh + r * (int64_t)Gpad + g, Zl + r * (int64_t)Gpad + g);
        }
        if (threadIdx.x == 0) cell_of_row[r] = (int32_t)c;
    }
}

void launch_dense_stats(mcb_ctx* ctx, int64_t ncells, const double* x, int32_t G, double* mu, double* inv, int32_t* valid)
{
    if (ncells == 0) return;
    int grid = (int)std::min<int64_t>(ncells, (int64_t)ctx->num_sms * 16);
    dense_stats_kernel<<<grid, 128, 0, ctx->stream>>>(ncells, x, G, mu, inv, valid);
    MCB_LAUNCH_CHECK();
}
void launch_dense_rows(mcb_ctx* ctx, int64_t ncells, const double* x, int32_t G, int32_t Gpad, const double* mu,
                       const double* inv, const int32_t* valid, const int32_t* rowpos, float* Zf, void* Zh, void* Zl,
                       int32_t* cell_of_row)
{
    if (ncells == 0) return;
    int grid = (int)std::min<int64_t>(ncells, (int64_t)ctx->num_sms * 16);
    dense_rows_kernel<<<grid, 128, 0, ctx->stream>>>(ncells, x, G, Gpad, mu, inv, valid, rowpos, Zf, (__half*)Zh, (__half*)Zl, cell_of_row);
    MCB_LAUNCH_CHECK();
}

void launch_feat_stats(mcb_ctx* ctx, int64_t ncells, const int64_t* p, const int32_t* ridx, const uint32_t* x,
                       const int32_t* feat_map, int32_t G, double k_nonz, double* mu, double* inv, int32_t* valid)
{
    if (ncells == 0) return;
    int grid = (int)std::min<int64_t>(ceil_div(ncells, 8), (int64_t)ctx->num_sms * 16);
    feat_stats_kernel<<<grid, 256, 0, ctx->stream>>>(ncells, p, ridx, x, feat_map, G, k_nonz, mu, inv, valid);
    MCB_LAUNCH_CHECK();
}
void launch_exclusive_scan_i32(mcb_ctx* ctx, const int32_t* in, int32_t* out, int64_t n, int32_t* total)
{
    excl_scan_kernel<int32_t><<<1, 1024, 0, ctx->stream>>>(in, out, n, total);
    MCB_LAUNCH_CHECK();
}
void launch_exclusive_scan_i64(mcb_ctx* ctx, const int64_t* in, int64_t* out, int64_t n, int64_t* total)
{
    excl_scan_kernel<long long><<<1, 1024, 0, ctx->stream>>>((const long long*)in, (long long*)out, n, (long long*)total);
    MCB_LAUNCH_CHECK();
}
void launch_feat_rows(mcb_ctx* ctx, int64_t ncells, const int64_t* p, const int32_t* ridx, const uint32_t* x,
                      const int32_t* feat_map, int32_t G, int32_t Gpad, double k_nonz, const double* mu,
                      const double* inv, const int32_t* valid, const int32_t* rowpos, float* Zf, void* Zh, void* Zl,
                      int32_t* cell_of_row, int64_t cell_base)
{
    if (ncells == 0) return;
    int grid = (int)std::min<int64_t>(ncells, (int64_t)ctx->num_sms * 16);
    feat_rows_kernel<<<grid, FR_THREADS, (size_t)Gpad * sizeof(float), ctx->stream>>>(
        ncells, p, ridx, x, feat_map, G, Gpad, k_nonz, mu, inv, valid, rowpos, Zf, (__half*)Zh, (__half*)Zl, cell_of_row, cell_base);
    MCB_LAUNCH_CHECK();
}
void launch_gather_rows(mcb_ctx* ctx, const void* Z, int64_t row_bytes, const int32_t* rows, int32_t nrows, void* out)
{
    if (nrows == 0) return;
    MCB_CHECK(row_bytes % 16 == 0, "gather_rows: row size must be a multiple of 16 bytes");
    const int32_t rv = (int32_t)(row_bytes / 16);
    int grid = (int)std::min<int64_t>(ceil_div((int64_t)nrows * rv, 256), (int64_t)ctx->num_sms * 16);
    gather_rows_kernel<<<grid, 256, 0, ctx->stream>>>((const uint4*)Z, rv, rows, nrows, (uint4*)out);
    MCB_LAUNCH_CHECK();
}

}  // namespace mcb
