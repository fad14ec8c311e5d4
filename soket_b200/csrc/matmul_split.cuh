// matmul_split.cuh -- fp16 hi/lo operand split with K-invariant power-of-two scales.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace sk {

struct SplitOperand {
  __half *hi = nullptr;        // (outer, ld) fp16, same major-ness as the source
  __half *lo = nullptr;
  float *inv_scale = nullptr;  // 2^-e per mn index (exact powers of two)
  int64_t ld = 0;
  void release();
};

// row_mul (MN-major case only): X[r, :] is multiplied by row_mul[r] (an exact power of two) before the split
// colsum_out (row-scaled case only, rows of up to 8192 elements: split_colsum_supported): receives the column
// sums of X -- the bias gradient when X is the adjoint of a Linear output -- from the same pass
int split_f16(const float *x, int64_t ldx, int64_t outer, int64_t inner, bool scale_rows, SplitOperand &out,
              const float *row_mul = nullptr, float *colsum_out = nullptr);
bool split_colsum_supported(int64_t inner);

}  // namespace sk
