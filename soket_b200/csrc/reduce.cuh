// reduce.cuh -- fp32 reduction entry points shared between translation units.
#pragma once
#include "common.cuh"
namespace sk {
// out[c] = sum_r in[r * row_stride + c]
int reduce_cols_sum_f32(const float *in, int64_t row_stride, float *out, int64_t R, int64_t C);
// out[r] = op_c in[r * row_stride + c]   (op: sk_reduce_op)
int reduce_rows_f32(int op, const float *in, int64_t row_stride, float *out, int64_t R, int64_t C);
}  // namespace sk
