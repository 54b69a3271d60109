// common.cuh -- shared device/host helpers for libmcb200 (sm_100a only).
//
// Everything in this library computes on the GPU; there is no CPU fallback.  Errors never cross
// the C ABI as exceptions: each entry point catches into mcb_set_error() and returns a status.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <stdexcept>
#include <atomic>

namespace mcb {

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
void set_error(const char* msg);

#define MCB_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            char _b[512];                                                                      \
            snprintf(_b, sizeof(_b), "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e),       \
                     __FILE__, __LINE__, cudaGetErrorString(_e));                              \
            throw ::mcb::Error(2, _b);                                                         \
        }                                                                                      \
    } while (0)

#define MCB_CHECK(cond, msg)                                                                   \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            char _b[512];                                                                      \
            snprintf(_b, sizeof(_b), "%s (%s:%d)", msg, __FILE__, __LINE__);                   \
            throw ::mcb::Error(1, _b);                                                         \
        }                                                                                      \
    } while (0)

extern std::atomic<long long> g_launches;   // kernels launched by this library (bench: gpu_launches)
#define MCB_LAUNCH_CHECK() do { ++::mcb::g_launches; MCB_CUDA(cudaGetLastError()); } while (0)

static inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11): the counter-based generator shared with the CPU oracle
// so host-supplied seeds give identical draws on both sides.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// fp16 operand planes hold z * 2^12 (features.cu); the contraction is rescaled by 2^-24 in the epilogue
constexpr float Z_PRESCALE = 4096.0f;
constexpr float C_RESCALE = 1.0f / (4096.0f * 4096.0f);

constexpr uint32_t TAG_DOWNSAMP = 0x4453u;   // 'DS'
constexpr uint32_t TAG_COVER    = 0x4743u;   // 'GC'
constexpr uint32_t STREAM_SEED  = 1u;
constexpr uint32_t STREAM_PERM  = 2u;

// ---------------------------------------------------------------------------------------------
// float <-> order-preserving uint32
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t float_to_ordered(float f)
{
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    uint32_t b; memcpy(&b, &f, 4);
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t o)
{
    uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f; memcpy(&f, &b, 4); return f;
#endif
}
// candidate key: ascending u64 order == (correlation descending, column ascending)
__host__ __device__ __forceinline__ uint64_t cand_key(float c, uint32_t col)
{
    return ((uint64_t)(~float_to_ordered(c)) << 32) | (uint64_t)col;
}
__host__ __device__ __forceinline__ float cand_key_cor(uint64_t k) { return ordered_to_float(~(uint32_t)(k >> 32)); }
__host__ __device__ __forceinline__ uint32_t cand_key_col(uint64_t k) { return (uint32_t)k; }

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// block-wide bitonic sort of n (power of two) u64 keys held in shared memory, ascending.
// All NT threads of the block must call it.
// ---------------------------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ void bitonic_sort_u64(uint64_t* s, int n)
{
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (n >> 1); t += NT) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int ixj = i | j;
                const bool up = ((i & k) == 0);
                const uint64_t a = s[i], b = s[ixj];
                if ((a > b) == up) { s[i] = b; s[ixj] = a; }
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// register-resident bitonic sort of N = 256 * IPT u64 keys by a 256-thread block, ascending.
// Element e = tid * IPT + r lives in key[r] of thread tid.  Compare-exchange partners closer than IPT
// are in the same thread, closer than 32 * IPT in the same warp (shuffles); only the log2(N / (32 IPT))
// widest strides go through shared memory (N u64 of scratch, conflict-free [r][tid] layout).  The
// all-shared-memory version above is bound by shared-memory bandwidth (4 accesses per exchange).
// ---------------------------------------------------------------------------------------------
template <int IPT>
__device__ __forceinline__ void bitonic_sort_regs_u64(uint64_t (&key)[IPT], uint64_t* scratch)
{
    constexpr int N = IPT * 256;
    const int tid = threadIdx.x, lane = tid & 31;
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j < IPT) {
#pragma unroll
                for (int r = 0; r < IPT; ++r) {
                    if ((r & j) == 0) {
                        const bool up = (((tid * IPT + r) & k) == 0);
                        const uint64_t a = key[r], b = key[r | j];
                        if ((a > b) == up) { key[r] = b; key[r | j] = a; }
                    }
                }
            } else if (j < IPT * 32) {
                const int lj = j / IPT;
                const bool lower = (lane & lj) == 0;
#pragma unroll
                for (int r = 0; r < IPT; ++r) {
                    const bool up = (((tid * IPT + r) & k) == 0);
                    const uint64_t other = __shfl_xor_sync(0xffffffffu, (unsigned long long)key[r], lj);
                    const bool keep_min = (lower == up);
                    key[r] = keep_min ? (key[r] < other ? key[r] : other) : (key[r] > other ? key[r] : other);
                }
            } else {
                const int tj = j / IPT;                      // partner thread distance
                __syncthreads();
#pragma unroll
                for (int r = 0; r < IPT; ++r) scratch[r * 256 + tid] = key[r];
                __syncthreads();
                const bool lower = (tid & tj) == 0;
#pragma unroll
                for (int r = 0; r < IPT; ++r) {
                    const bool up = (((tid * IPT + r) & k) == 0);
                    const uint64_t other = scratch[r * 256 + (tid ^ tj)];
                    const bool keep_min = (lower == up);
                    key[r] = keep_min ? (key[r] < other ? key[r] : other) : (key[r] > other ? key[r] : other);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// one warp: smallest bucket d of counts[0 .. 32 * PER) whose inclusive cumulative count exceeds k;
// `before` = cumulative count of the buckets below d.  Result valid in every lane.  Replaces the
// 256-step single-thread scans of the radix-select kernels (each step a dependent smem load).
// ---------------------------------------------------------------------------------------------
template <int PER>
__device__ __forceinline__ void warp_find_bucket(const uint32_t* counts, long long k, int& d, long long& before)
{
    static_assert((PER & (PER - 1)) == 0, "PER must be a power of two");
    const int lane = threadIdx.x & 31;
    // level 1: sum of each lane's PER consecutive counts.  The read order is rotated by the lane index: reading
    // counts[lane * PER + q] straight is a PER-way (up to 32-way) shared-memory bank conflict.
    long long s = 0;
#pragma unroll
    for (int q = 0; q < PER; ++q) s += counts[lane * PER + ((q + lane) & (PER - 1))];
    long long incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    const long long excl = incl - s;
    const unsigned owner_mask = __ballot_sync(0xffffffffu, excl <= k && k < incl);
    if (owner_mask == 0u) {                      // k is past the total: last bucket, everything before it
        d = 32 * PER - 1;
        before = __shfl_sync(0xffffffffu, incl, 31);
        return;
    }
    const int owner = __ffs(owner_mask) - 1;
    const long long base = __shfl_sync(0xffffffffu, excl, owner);
    // level 2: the warp scans the owner's PER counts together (conflict-free, no serial dependent loads)
    constexpr int E = PER >= 32 ? PER / 32 : 1;
    uint32_t c2[E];
    long long s2 = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int idx = lane * E + e;
        c2[e] = idx < PER ? counts[owner * PER + idx] : 0u;
        s2 += c2[e];
    }
    long long incl2 = s2;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long up = __shfl_up_sync(0xffffffffu, incl2, o);
        if (lane >= o) incl2 += up;
    }
    const long long excl2 = incl2 - s2;
    const long long kk = k - base;
    const unsigned m2 = __ballot_sync(0xffffffffu, excl2 <= kk && kk < incl2);
    const int owner2 = m2 ? __ffs(m2) - 1 : 31;
    int dd = owner * PER + PER - 1;
    long long bb = base + excl2;
    if (lane == owner2) {
        long long run = excl2;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            if (run + c2[e] > kk) { dd = owner * PER + lane * E + e; bb = base + run; break; }
            run += c2[e];
            bb = base + run;
        }
    }
    d = __shfl_sync(0xffffffffu, dd, owner2);
    before = __shfl_sync(0xffffffffu, bb, owner2);
}

__device__ __forceinline__ int next_pow2(int n)
{
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

// round an fp32 value to TF32 (10 explicit mantissa bits, round-to-nearest-even), low 13 bits zero
__device__ __forceinline__ float tf32_round(float x)
{
    uint32_t b = __float_as_uint(x);
    const uint32_t lsb = (b >> 13) & 1u;
    b += 0x0fffu + lsb;
    b &= 0xffffe000u;
    return __uint_as_float(b);
}

__device__ __forceinline__ float warp_reduce_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_reduce_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

}  // namespace mcb
