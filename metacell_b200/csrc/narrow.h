// narrow.h -- host-side narrowing of the dgCMatrix slots while they are staged into pinned memory (staging.cu).
// Plain C++ (narrow.cpp, compiled by g++ with AVX2 code paths selected at run time); no CUDA.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace mcb {

// counts: double -> uint16.  Returns 0 when every value was an integer in [0, 65535): all of out[] is valid.
// Returns 1 when the chunk holds counts >= 65535 (legal: out[] carries 65535 at those positions, the caller escapes them)
// Returns 2 when some value is not an integer in [0, 2^32) (NaN, negative, fractional, too large): the input is invalid.
int narrow_counts_u16(const double* src, size_t n, uint16_t* out);
// row indices: int32 -> uint16; returns 0 when every value was in [0, 65536), else 2
int narrow_rows_u16(const int32_t* src, size_t n, uint16_t* out);

}  // namespace mcb
