// matmul_split.cuh -- fp16 hi/lo operand split with K-invariant power-of-two scales.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace sk {

// amax -> (scale, 1/scale) as exact powers of two with amax * scale in [2^14, 2^15): the top of the
// fp16 range with one bit of headroom for rounding.  Zero / non-finite maxima are left alone.
__device__ __forceinline__ void pow2_scale(float amax, float &scale, float &inv) {
  const uint32_t bits = __float_as_uint(amax);
  const int ef = (int)((bits >> 23) & 0xFF);
  if (ef == 0 || ef == 0xFF) { scale = 1.f; inv = 1.f; return; }
  int shift = 14 - (ef - 127);
  if (shift > 126) shift = 126;       // rows below 2^-112: products underflow fp32 anyway
  scale = __uint_as_float((uint32_t)(shift + 127) << 23);
  inv = __uint_as_float((uint32_t)(127 - shift) << 23);
}

// four fp32 values -> four fp16 hi + four fp16 lo (x * s = hi + lo), packed for 8-byte stores
__device__ __forceinline__ void split4(const float4 &v, float s, uint2 &hi, uint2 &lo) {
  const float x[4] = {v.x * s, v.y * s, v.z * s, v.w * s};
  uint32_t h[2], l[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const __half h0 = __float2half_rn(x[2 * j]), h1 = __float2half_rn(x[2 * j + 1]);
    const __half l0 = __float2half_rn(x[2 * j] - __half2float(h0));
    const __half l1 = __float2half_rn(x[2 * j + 1] - __half2float(h1));
    h[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
    l[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
  }
  hi = make_uint2(h[0], h[1]);
  lo = make_uint2(l[0], l[1]);
}

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

// ONE scale for the whole matrix (sk_split_f16): hi / lo / scale are caller-owned.  amax_bits: device
// word holding the bit pattern of max |x| (or of any upper bound of it); NULL = computed here first.
// scale4: device float[4], receives {scale, 1/scale, amax, 0}.
int split_f16_tensor(const float *x, int64_t ldx, int64_t rows, int64_t cols, const uint32_t *amax_bits, __half *hi,
                     __half *lo, int64_t ldh, float *scale4, float *colsum_out);

}  // namespace sk
