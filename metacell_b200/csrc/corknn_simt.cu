// corknn_simt.cu -- CUDA-core version (fp32 inputs, fp64 accumulation) of the cell x cell correlation
// contraction with the same two epilogues as the tcgen05 kernel (corknn_tc.cu).  It exists (a) to validate the tensor-core
// kernel on the GPU against an independent implementation, (b) as the exact path for the handful
// of rows whose sampled threshold missed its window (see cgraph.cu), and (c) behind the
// "gemm_impl" = 1 option.  It is GPU code, not a CPU fallback, and is not the default path.
//
// C[i][j] = sum_g A[i][g] * B[j][g]      (tgs_cor_knn's Pearson correlation on standardised rows,
// reference call site R/utils.r:119); A, B are rows of the fp32 plane Zf.
#include "internal.h"

namespace mcb {

constexpr int ST = 64;      // tile edge
constexpr int SK = 16;      // k step
constexpr int SPAD = 4;

template <int MODE>  // 0 = store, 1 = filter
__global__ void __launch_bounds__(256)
simt_cor_kernel(const float* __restrict__ A, int64_t a_row0, int64_t a_rows,
                const float* __restrict__ B, int64_t b_rows, int32_t Gpad,
                float* __restrict__ C, int64_t ldc, const float* __restrict__ thr, uint64_t* __restrict__ cand,
                uint32_t* __restrict__ cnt, int32_t cap)
{
    __shared__ float As[SK][ST + SPAD];
    __shared__ float Bs[SK][ST + SPAD];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t i0 = (int64_t)blockIdx.y * ST, j0 = (int64_t)blockIdx.x * ST;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;

    const int lr = threadIdx.x >> 2;          // 0..63 row within tile
    const int lk = (threadIdx.x & 3) * 4;     // 0,4,8,12
    for (int k0 = 0; k0 < Gpad; k0 += SK) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (i0 + lr < a_rows) a = *reinterpret_cast<const float4*>(A + (i0 + lr) * (int64_t)Gpad + k0 + lk);
        if (j0 + lr < b_rows) b = *reinterpret_cast<const float4*>(B + (j0 + lr) * (int64_t)Gpad + k0 + lk);
        __syncthreads();
        As[lk + 0][lr] = a.x; As[lk + 1][lr] = a.y; As[lk + 2][lr] = a.z; As[lk + 3][lr] = a.w;
        Bs[lk + 0][lr] = b.x; Bs[lk + 1][lr] = b.y; Bs[lk + 2][lr] = b.z; Bs[lk + 3][lr] = b.w;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SK; ++k) {
            float av[4], bv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { av[q] = As[k][ty * 4 + q]; bv[q] = Bs[k][tx * 4 + q]; }
#pragma unroll
            for (int a2 = 0; a2 < 4; ++a2)
#pragma unroll
                for (int b2 = 0; b2 < 4; ++b2) acc[a2][b2] = fma((double)av[a2], (double)bv[b2], acc[a2][b2]);
        }
    }
#pragma unroll
    for (int a2 = 0; a2 < 4; ++a2) {
        const int64_t i = i0 + ty * 4 + a2;
        if (i >= a_rows) continue;
        if (MODE == 0) {
#pragma unroll
            for (int b2 = 0; b2 < 4; ++b2) {
                const int64_t j = j0 + tx * 4 + b2;
                if (j < b_rows) C[i * ldc + j] = (float)acc[a2][b2];
            }
        } else {
            const float t = thr[i];
#pragma unroll
            for (int b2 = 0; b2 < 4; ++b2) {
                const int64_t j = j0 + tx * 4 + b2;
                const float c = (float)acc[a2][b2];
                if (j < b_rows && j != a_row0 + i && c >= t) {
                    const uint32_t pos = atomicAdd(&cnt[i], 1u);
                    if (pos < (uint32_t)cap) cand[i * (int64_t)cap + pos] = cand_key(c, (uint32_t)j);
                }
            }
        }
    }
}

void launch_simt_store(mcb_ctx* ctx, const float* A, int64_t a_rows, const float* B, int64_t b_rows, int32_t Gpad,
                       float* C, int64_t ldc)
{
    if (a_rows == 0 || b_rows == 0) return;
    dim3 grid((unsigned)ceil_div(b_rows, ST), (unsigned)ceil_div(a_rows, ST));
    MCB_CHECK(grid.y < 65536, "simt store: too many row tiles for one launch");
    simt_cor_kernel<0><<<grid, 256, 0, ctx->stream>>>(A, 0, a_rows, B, b_rows, Gpad, C, ldc, nullptr, nullptr, nullptr, 0);
    MCB_LAUNCH_CHECK();
}
void launch_simt_filter(mcb_ctx* ctx, const float* A, int64_t a_row0, int64_t a_rows, const float* B, int64_t b_rows,
                        int32_t Gpad, const float* thr, uint64_t* cand, uint32_t* cnt, int32_t cap)
{
    if (a_rows == 0 || b_rows == 0) return;
    dim3 grid((unsigned)ceil_div(b_rows, ST), (unsigned)ceil_div(a_rows, ST));
    MCB_CHECK(grid.y < 65536, "simt filter: too many row tiles for one launch");
    simt_cor_kernel<1><<<grid, 256, 0, ctx->stream>>>(A, a_row0, a_rows, B, b_rows, Gpad, nullptr, 0, thr, cand, cnt, cap);
    MCB_LAUNCH_CHECK();
}

}  // namespace mcb
