// narrow.cpp -- see narrow.h.  The scalar loops are the definition; the AVX2 loops are the same arithmetic eight
// values at a time (cvttpd2dq / packusdw) and fall back to the scalar loop for any block that is not all small
// non-negative integers, so both paths return identical results.
#include "narrow.h"
#include <immintrin.h>

namespace mcb {

static int counts_scalar(const double* src, size_t n, uint16_t* out)
{
    int rc = 0;
    for (size_t q = 0; q < n; ++q) {
        const double v = src[q];
        if (!(v >= 0.0 && v < 4294967296.0)) return 2;          // also NaN
        const uint32_t u = (uint32_t)v;
        if ((double)u != v) return 2;
        if (u >= 65535u) { out[q] = 65535u; rc = 1; }
        else out[q] = (uint16_t)u;
    }
    return rc;
}
static int rows_scalar(const int32_t* src, size_t n, uint16_t* out)
{
    uint32_t acc = 0;
    for (size_t q = 0; q < n; ++q) { acc |= (uint32_t)src[q]; out[q] = (uint16_t)src[q]; }
    return acc > 65535u ? 2 : 0;
}

__attribute__((target("avx2"))) static int counts_avx2(const double* src, size_t n, uint16_t* out)
{
    int rc = 0;
    size_t q = 0;
    const __m128i lim = _mm_set1_epi32(65534);
    for (; q + 8 <= n; q += 8) {
        const __m256d a = _mm256_loadu_pd(src + q), b = _mm256_loadu_pd(src + q + 4);
        const __m128i ia = _mm256_cvttpd_epi32(a), ib = _mm256_cvttpd_epi32(b);           // out of range / NaN -> INT_MIN
        const __m256d ra = _mm256_cvtepi32_pd(ia), rb = _mm256_cvtepi32_pd(ib);
        const int exact = _mm256_movemask_pd(_mm256_cmp_pd(a, ra, _CMP_EQ_OQ)) & _mm256_movemask_pd(_mm256_cmp_pd(b, rb, _CMP_EQ_OQ));
        // small = 0 <= v <= 65534 in every lane (signed compares: negatives and INT_MIN fail the first test)
        const __m128i neg = _mm_or_si128(_mm_srai_epi32(ia, 31), _mm_srai_epi32(ib, 31));
        const __m128i big = _mm_or_si128(_mm_cmpgt_epi32(ia, lim), _mm_cmpgt_epi32(ib, lim));
        if (exact == 0xf && _mm_testz_si128(_mm_or_si128(neg, big), _mm_set1_epi32(-1))) {
            _mm_storeu_si128((__m128i*)(out + q), _mm_packus_epi32(ia, ib));
        } else {
            const int r = counts_scalar(src + q, 8, out + q);
            if (r == 2) return 2;
            rc |= r;
        }
    }
    if (q < n) {
        const int r = counts_scalar(src + q, n - q, out + q);
        if (r == 2) return 2;
        rc |= r;
    }
    return rc;
}
__attribute__((target("avx2"))) static int rows_avx2(const int32_t* src, size_t n, uint16_t* out)
{
    __m128i acc = _mm_setzero_si128();
    size_t q = 0;
    for (; q + 8 <= n; q += 8) {
        const __m128i a = _mm_loadu_si128((const __m128i*)(src + q)), b = _mm_loadu_si128((const __m128i*)(src + q + 4));
        acc = _mm_or_si128(acc, _mm_or_si128(a, b));
        const __m128i m = _mm_set1_epi32(0xffff);
        _mm_storeu_si128((__m128i*)(out + q), _mm_packus_epi32(_mm_and_si128(a, m), _mm_and_si128(b, m)));
    }
    uint32_t lanes[4];
    _mm_storeu_si128((__m128i*)lanes, acc);
    uint32_t all = lanes[0] | lanes[1] | lanes[2] | lanes[3];
    for (; q < n; ++q) { all |= (uint32_t)src[q]; out[q] = (uint16_t)src[q]; }
    return all > 65535u ? 2 : 0;
}

static bool has_avx2()
{
    static const bool v = __builtin_cpu_supports("avx2");
    return v;
}

int narrow_counts_u16(const double* src, size_t n, uint16_t* out) { return has_avx2() ? counts_avx2(src, n, out) : counts_scalar(src, n, out); }
int narrow_rows_u16(const int32_t* src, size_t n, uint16_t* out) { return has_avx2() ? rows_avx2(src, n, out) : rows_scalar(src, n, out); }

}  // namespace mcb
