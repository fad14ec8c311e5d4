// matmul_split.cu -- operand preparation for the fp16x3 GEMM (SK_MM_F16X3).
//
// An fp32 operand X (mn x K) is rewritten as two fp16 matrices and one power-of-two
// scale per mn index:   X[i, k] * 2^e[i]  =  hi[i, k] + lo[i, k]  (+ <= 2^-24 relative)
//   e[i]   = 14 - floor(log2(max_k |X[i, k]|))  -> the scaled row maximum lies in
//            [2^14, 2^15): the top of the fp16 range, no overflow after rounding
//   hi     = fp16_rn(X * 2^e)                    (11 significant bits)
//   lo     = fp16_rn(X * 2^e - hi)               (the next 11 bits; exact difference)
// The scale is constant along K, so A@B = 2^-(ea[m] + eb[n]) * (Ahi Bhi + Ahi Blo +
// Alo Bhi) with every product exact in the fp32 accumulator (11 x 11 bits); the
// dropped Alo Blo term is <= 2^-24 relative.  Elements more than 2^15 below their
// row maximum fall into the fp16 subnormal range and keep an ABSOLUTE error of
// 2^-39 of the row maximum instead (documented in DESIGN.md section 4.1).
//
// Two storage cases, matching what TMA / tcgen05 consume in place:
//   K-major   `outer` = mn rows of K contiguous elements -> one scale per stored row
//   MN-major  `outer` = K rows of mn contiguous elements -> one scale per stored column
#include "common.cuh"
#include "matmul_split.cuh"

namespace sk {

__device__ __forceinline__ float absmax4(float m, const float4 &v) {
  return fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
}

struct Half8 { uint4 v; };

__device__ __forceinline__ void split8(const float4 &a, const float4 &b, float s, uint4 &hi, uint4 &lo) {
  const float x[8] = {a.x * s, a.y * s, a.z * s, a.w * s, b.x * s, b.y * s, b.z * s, b.w * s};
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __half h0 = __float2half_rn(x[2 * j]), h1 = __float2half_rn(x[2 * j + 1]);
    const __half l0 = __float2half_rn(x[2 * j] - __half2float(h0));
    const __half l1 = __float2half_rn(x[2 * j + 1] - __half2float(h1));
    h[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
    l[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// ---------------------------------------------------------------- K-major operands
// One thread group of TPR threads per stored row; a thread owns 8-element units
// t, t + TPR, ... (VPT of them, held in registers between the max and the split).
// Columns >= inner (inside the 16-byte pitch padding) read as zero.
template <int TPR, int VPT>
__global__ void __launch_bounds__(256)
split_rows_kernel(const float *__restrict__ x, int64_t ldx, __half *__restrict__ hi, __half *__restrict__ lo,
                  int64_t ldh, float *__restrict__ inv_scale, int64_t outer, int inner,
                  float *__restrict__ cs_part, const uint32_t *__restrict__ tensor_amax) {
  __shared__ float red[8];
  // tensor_amax: ONE scale for the whole matrix, from the bit pattern of (a bound of) its |max|; block 0
  // publishes {scale, 1/scale, amax, 0} in inv_scale[0..3] instead of one inverse scale per row
  float ts = 1.f, tinv = 1.f;
  if (tensor_amax) {
    pow2_scale(__uint_as_float(*tensor_amax), ts, tinv);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      inv_scale[0] = ts; inv_scale[1] = tinv; inv_scale[2] = __uint_as_float(*tensor_amax); inv_scale[3] = 0.f;
    }
  }
  constexpr int RPB = 256 / TPR;
  const int t = threadIdx.x % TPR, grp = threadIdx.x / TPR;
  const int units = (int)(ldh >> 3);
  // by-product for Linear backward: column sums of the UNSCALED operand (the bias gradient,
  // autodiff.pyx:84) -- a thread always visits the same columns, so it keeps their running sums
  // and writes one partial row per thread group at the end (cs_part: (gridDim.x * RPB, ldh))
  float4 ca[VPT], cb[VPT];
#pragma unroll
  for (int j = 0; j < VPT; ++j) { ca[j] = make_float4(0.f, 0.f, 0.f, 0.f); cb[j] = ca[j]; }
  for (int64_t row0 = (int64_t)blockIdx.x * RPB; row0 < outer; row0 += (int64_t)gridDim.x * RPB) {
    const int64_t row = row0 + grp;
    const bool live = row < outer;
    const float *xr = x + row * ldx;
    float4 a[VPT], b[VPT];
    float m = 0.f;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int u = t + j * TPR, c = u * 8;
      a[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      b[j] = a[j];
      if (live && u < units) {
        if (c + 4 <= inner) a[j] = ld_stream(reinterpret_cast<const float4 *>(xr + c));
        else {
          if (c + 0 < inner) a[j].x = xr[c + 0];
          if (c + 1 < inner) a[j].y = xr[c + 1];
          if (c + 2 < inner) a[j].z = xr[c + 2];
        }
        if (c + 8 <= inner) b[j] = ld_stream(reinterpret_cast<const float4 *>(xr + c + 4));
        else {
          if (c + 4 < inner) b[j].x = xr[c + 4];
          if (c + 5 < inner) b[j].y = xr[c + 5];
          if (c + 6 < inner) b[j].z = xr[c + 6];
        }
        m = absmax4(absmax4(m, a[j]), b[j]);
        if (cs_part) {
          ca[j].x += a[j].x; ca[j].y += a[j].y; ca[j].z += a[j].z; ca[j].w += a[j].w;
          cb[j].x += b[j].x; cb[j].y += b[j].y; cb[j].z += b[j].z; cb[j].w += b[j].w;
        }
      }
    }
    float s = ts, inv = tinv;
    if (!tensor_amax) {
      m = warp_max(m);
      if (TPR > 32) {
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
        __syncthreads();
        m = red[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
      }
      pow2_scale(m, s, inv);
      if (live && t == 0) inv_scale[row] = inv;
    }
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int u = t + j * TPR;
      if (live && u < units) {
        uint4 h, l;
        split8(a[j], b[j], s, h, l);
        *reinterpret_cast<uint4 *>(hi + row * ldh + (int64_t)u * 8) = h;   // re-read soon by TMA: keep in L2
        *reinterpret_cast<uint4 *>(lo + row * ldh + (int64_t)u * 8) = l;
      }
    }
  }
  if (cs_part) {
    float *pr = cs_part + ((int64_t)blockIdx.x * RPB + grp) * ldh;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int u = t + j * TPR;
      if (u < units) {
        *reinterpret_cast<float4 *>(pr + (int64_t)u * 8) = ca[j];
        *reinterpret_cast<float4 *>(pr + (int64_t)u * 8 + 4) = cb[j];
      }
    }
  }
}

// column sums of the (P x ld) partial matrix above (L2 resident: just written).  Block = 8 float4
// column groups x 32 row lanes; the 32 lane sums are added in a fixed order (deterministic).
__global__ void __launch_bounds__(256)
colsum_partials_kernel(const float *__restrict__ part, int64_t P, int64_t ld, float *__restrict__ out, int64_t C) {
  __shared__ float4 sm[32][8];
  const int cg = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int64_t c = ((int64_t)blockIdx.x * 8 + cg) * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < ld) {
    for (int64_t r = rl; r < P; r += 32) {
      const float4 v = *reinterpret_cast<const float4 *>(part + r * ld + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  sm[rl][cg] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    const int g = threadIdx.x >> 2, k = threadIdx.x & 3;
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += reinterpret_cast<const float *>(&sm[i][g])[k];
    const int64_t cc = ((int64_t)blockIdx.x * 8 + g) * 4 + k;
    if (cc < C) out[cc] = t;
  }
}

// rows longer than the register cache: one block per row, the row is read twice (the
// second time from L1/L2).
__global__ void __launch_bounds__(256)
split_rows_long_kernel(const float *__restrict__ x, int64_t ldx, __half *__restrict__ hi, __half *__restrict__ lo,
                       int64_t ldh, float *__restrict__ inv_scale, int64_t outer, int64_t inner,
                       const uint32_t *__restrict__ tensor_amax) {
  __shared__ float red[8];
  const int64_t units = ldh >> 3;
  float ts = 1.f, tinv = 1.f;
  if (tensor_amax) {
    pow2_scale(__uint_as_float(*tensor_amax), ts, tinv);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      inv_scale[0] = ts; inv_scale[1] = tinv; inv_scale[2] = __uint_as_float(*tensor_amax); inv_scale[3] = 0.f;
    }
  }
  for (int64_t row = blockIdx.x; row < outer; row += gridDim.x) {
    const float *xr = x + row * ldx;
    float s = ts, inv = tinv;
    if (!tensor_amax) {
      float m = 0.f;
      for (int64_t c = (int64_t)threadIdx.x * 4; c < inner; c += 1024) {
        if (c + 4 <= inner) m = absmax4(m, *reinterpret_cast<const float4 *>(xr + c));
        else
          for (int64_t k = c; k < inner; ++k) m = fmaxf(m, fabsf(xr[k]));
      }
      m = warp_max(m);
      __syncthreads();
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
      __syncthreads();
      m = red[0];
#pragma unroll
      for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
      pow2_scale(m, s, inv);
      if (threadIdx.x == 0) inv_scale[row] = inv;
    }
    for (int64_t u = threadIdx.x; u < units; u += 256) {
      const int64_t c = u * 8;
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = (c + k < inner) ? xr[c + k] : 0.f;
      uint4 h, l;
      split8(make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7]), s, h, l);
      *reinterpret_cast<uint4 *>(hi + row * ldh + c) = h;
      *reinterpret_cast<uint4 *>(lo + row * ldh + c) = l;
    }
  }
}

// ---------------------------------------------------------------- MN-major operands
// pass 1: column |max| of a (outer x inner) matrix into colmax (uint32 bit patterns of
// non-negative floats order like the floats: atomicMax is exact and order-independent).
__global__ void __launch_bounds__(256)
absmax_cols_kernel(const float *__restrict__ x, int64_t ldx, const float *__restrict__ row_mul,
                   uint32_t *__restrict__ colmax, int64_t outer, int64_t inner, int64_t rows_per_slab) {
  __shared__ float sm[8][129];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = ((int64_t)blockIdx.x * 32 + tx) * 4;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_slab;
  const int64_t r1 = (r0 + rows_per_slab < outer) ? r0 + rows_per_slab : outer;
  float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c + 4 <= inner) {
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      const float4 v = __ldg(reinterpret_cast<const float4 *>(x + r * ldx + c));   // read again by the split pass
      const float w = row_mul ? fabsf(__ldg(row_mul + r)) : 1.f;
      m.x = fmaxf(m.x, fabsf(v.x) * w); m.y = fmaxf(m.y, fabsf(v.y) * w);
      m.z = fmaxf(m.z, fabsf(v.z) * w); m.w = fmaxf(m.w, fabsf(v.w) * w);
    }
  } else if (c < inner) {
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      const float *p = x + r * ldx + c;
      const float w = row_mul ? fabsf(__ldg(row_mul + r)) : 1.f;
      m.x = fmaxf(m.x, fabsf(p[0]) * w);
      if (c + 1 < inner) m.y = fmaxf(m.y, fabsf(p[1]) * w);
      if (c + 2 < inner) m.z = fmaxf(m.z, fabsf(p[2]) * w);
    }
  }
  sm[ty][tx * 4 + 0] = m.x; sm[ty][tx * 4 + 1] = m.y; sm[ty][tx * 4 + 2] = m.z; sm[ty][tx * 4 + 3] = m.w;
  __syncthreads();
  if (threadIdx.x < 128) {
    float v = sm[0][threadIdx.x];
#pragma unroll
    for (int j = 1; j < 8; ++j) v = fmaxf(v, sm[j][threadIdx.x]);
    const int64_t cc = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (cc < inner && v > 0.f) atomicMax(colmax + cc, __float_as_uint(v));
  }
}

// pass 2: thread = one 8-column unit x a strided set of rows.
__global__ void __launch_bounds__(256)
split_cols_kernel(const float *__restrict__ x, int64_t ldx, const float *__restrict__ row_mul,
                  const uint32_t *__restrict__ colmax, __half *__restrict__ hi, __half *__restrict__ lo, int64_t ldh, float *__restrict__ inv_scale,
                  int64_t outer, int64_t inner, int64_t rows_per_slab) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t u = (int64_t)blockIdx.x * 32 + tx, c = u * 8;
  if (c >= ldh) return;
  float s[8], inv[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float m = (c + k < inner) ? __uint_as_float(colmax[c + k]) : 0.f;
    pow2_scale(m, s[k], inv[k]);
  }
  if (blockIdx.y == 0 && ty == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (c + k < inner) inv_scale[c + k] = inv[k];
  }
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_slab;
  const int64_t r1 = (r0 + rows_per_slab < outer) ? r0 + rows_per_slab : outer;
  for (int64_t r = r0 + ty; r < r1; r += 8) {
    const float *p = x + r * ldx + c;
    float v[8];
    if (c + 8 <= inner) {
      const float4 a = ld_stream(reinterpret_cast<const float4 *>(p));
      const float4 b = ld_stream(reinterpret_cast<const float4 *>(p + 4));
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = (c + k < inner) ? p[k] : 0.f;
    }
    if (row_mul) {   // exact: the multipliers are powers of two
      const float w = __ldg(row_mul + r);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] *= w;
    }
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float x0 = v[2 * j] * s[2 * j], x1 = v[2 * j + 1] * s[2 * j + 1];
      const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
      const __half l0 = __float2half_rn(x0 - __half2float(h0)), l1 = __float2half_rn(x1 - __half2float(h1));
      h[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
      l[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
    }
    *reinterpret_cast<uint4 *>(hi + r * ldh + c) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4 *>(lo + r * ldh + c) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// ---------------------------------------------------------------- host side
// row-major (outer x inner) fp32 -> hi / lo (pitch ldh); per-row scales into inv_scale[outer], or with
// tensor_amax one scale for the whole matrix into inv_scale[0..3]
static int launch_split_rows(const float *x, int64_t ldx, int64_t outer, int64_t inner, __half *hi, __half *lo,
                             int64_t ldh, float *inv_scale, float *colsum_out, const uint32_t *tensor_amax) {
  int rc;
  const int64_t units = ldh / 8;
  float *cs_part = nullptr;
  int64_t cs_rows = 0;
  // with column sums: fewer, longer-running blocks keep the partial matrix small (4 per SM)
#define ROWS(TPR, VPT)                                                                                  \
  do {                                                                                                  \
    const int grid = grid_for(outer, 256 / TPR, colsum_out ? 4 : 16);                                   \
    if (colsum_out) {                                                                                   \
      cs_rows = (int64_t)grid * (256 / TPR);                                                            \
      if ((rc = sk_malloc((size_t)(cs_rows * ldh) * sizeof(float), (void **)&cs_part))) return rc;      \
    }                                                                                                   \
    split_rows_kernel<TPR, VPT><<<grid, 256, 0, stream()>>>(x, ldx, hi, lo, ldh, inv_scale, outer,      \
                                                            (int)inner, cs_part, tensor_amax);          \
  } while (0)
  if (units <= 32) ROWS(32, 1);
  else if (units <= 64) ROWS(32, 2);
  else if (units <= 128) ROWS(32, 4);
  else if (units <= 256) ROWS(256, 1);
  else if (units <= 512) ROWS(256, 2);
  else if (units <= 1024) ROWS(256, 4);
  else {
    const int grid = grid_for(outer, 1, 8);
    split_rows_long_kernel<<<grid, 256, 0, stream()>>>(x, ldx, hi, lo, ldh, inv_scale, outer, inner, tensor_amax);
  }
#undef ROWS
  SK_LAUNCH_CHECK();
  if (cs_part) {
    colsum_partials_kernel<<<(unsigned)((ldh / 4 + 7) / 8), 256, 0, stream()>>>(cs_part, cs_rows, ldh, colsum_out, inner);
    SK_LAUNCH_CHECK();
    sk_free(cs_part);   // stream-ordered
  }
  return SK_OK;
}

// max |x| over a (rows x cols) matrix as a bit pattern (non-negative floats order like their bits)
__global__ void __launch_bounds__(256)
absmax_tensor_kernel(const float *__restrict__ x, int64_t ldx, int64_t rows, int64_t cols, uint32_t *__restrict__ out) {
  __shared__ float red[8];
  float m = 0.f;
  const int64_t c4 = cols >> 2, n4 = rows * c4;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const int64_t r = i / c4, c = (i - r * c4) * 4;
    m = absmax4(m, __ldg(reinterpret_cast<const float4 *>(x + r * ldx + c)));   // read again by the split pass
  }
  if ((cols & 3) && blockIdx.x == 0)
    for (int64_t r = threadIdx.x; r < rows; r += 256)
      for (int64_t c = c4 * 4; c < cols; ++c) m = fmaxf(m, fabsf(x[r * ldx + c]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
    if (m > 0.f) atomicMax(out, __float_as_uint(m));
  }
}

int split_f16_tensor(const float *x, int64_t ldx, int64_t rows, int64_t cols, const uint32_t *amax_bits, __half *hi,
                     __half *lo, int64_t ldh, float *scale4, float *colsum_out) {
  if (colsum_out && !split_colsum_supported(cols)) {
    set_error("split_f16_tensor: column sums come with rows of up to 8192 elements only");
    return SK_ERR_ARG;
  }
  int rc;
  uint32_t *own = nullptr;
  if (!amax_bits) {
    if ((rc = sk_malloc(sizeof(uint32_t), (void **)&own))) return rc;
    SK_CUDA(cudaMemsetAsync(own, 0, sizeof(uint32_t), stream()));
    const int grid = grid_for(rows * (cols >> 2) + 1, 256, 8);
    absmax_tensor_kernel<<<grid, 256, 0, stream()>>>(x, ldx, rows, cols, own);
    SK_LAUNCH_CHECK();
    amax_bits = own;
  }
  rc = launch_split_rows(x, ldx, rows, cols, hi, lo, ldh, scale4, colsum_out, amax_bits);
  if (own) sk_free(own);   // stream-ordered
  return rc;
}

void SplitOperand::release() {
  if (hi) sk_free(hi);
  if (inv_scale) sk_free(inv_scale);
  hi = nullptr; lo = nullptr; inv_scale = nullptr;
}

// x: `outer` stored rows of `inner` contiguous fp32 (pitch ldx, 16-byte aligned base and
// pitch).  scale_rows: one scale per stored row (K-major operand), else per stored column.
bool split_colsum_supported(int64_t inner) { return (inner + 7) / 8 <= 1024; }

int split_f16(const float *x, int64_t ldx, int64_t outer, int64_t inner, bool scale_rows, SplitOperand &out,
              const float *row_mul, float *colsum_out) {
  if (colsum_out && (!scale_rows || !split_colsum_supported(inner))) {
    set_error("split_f16: column sums come with the row-scaled split of rows up to 8192 elements only");
    return SK_ERR_ARG;
  }
  if (scale_rows && row_mul) {
    set_error("split_f16: row multipliers apply to the column-scaled (MN-major) case only");
    return SK_ERR_ARG;
  }
  const int64_t ldh = (inner + 7) / 8 * 8;
  const int64_t n_scale = scale_rows ? outer : inner;
  const size_t mat_bytes = ((size_t)(outer * ldh) * sizeof(__half) + 255) / 256 * 256;
  int rc;
  void *p = nullptr;
  if ((rc = sk_malloc(2 * mat_bytes, &p))) return rc;     // hi and lo share one block
  out.hi = (__half *)p;
  out.lo = (__half *)((char *)p + mat_bytes);
  out.ld = ldh;
  // inverse scales, then (MN-major) the column-max scratch words
  if ((rc = sk_malloc((size_t)n_scale * (scale_rows ? 4 : 8), (void **)&out.inv_scale))) { out.release(); return rc; }
  if (scale_rows) {
    if ((rc = launch_split_rows(x, ldx, outer, inner, out.hi, out.lo, ldh, out.inv_scale, colsum_out, nullptr))) {
      out.release();
      return rc;
    }
  } else {
    uint32_t *colmax = (uint32_t *)(out.inv_scale + n_scale);
    SK_CUDA(cudaMemsetAsync(colmax, 0, (size_t)n_scale * 4, stream()));
    const int64_t col_blocks128 = (inner + 127) / 128;
    int64_t slabs = ((int64_t)ctx().num_sms * 8 + col_blocks128 - 1) / col_blocks128;
    if (slabs > (outer + 7) / 8) slabs = (outer + 7) / 8;
    if (slabs < 1) slabs = 1;
    if (slabs > 65535) slabs = 65535;
    int64_t rps = (outer + slabs - 1) / slabs;
    dim3 g1((unsigned)col_blocks128, (unsigned)((outer + rps - 1) / rps));
    absmax_cols_kernel<<<g1, 256, 0, stream()>>>(x, ldx, row_mul, colmax, outer, inner, rps);
    SK_LAUNCH_CHECK();
    const int64_t col_blocks256 = (ldh / 8 + 31) / 32;
    slabs = ((int64_t)ctx().num_sms * 8 + col_blocks256 - 1) / col_blocks256;
    if (slabs > (outer + 7) / 8) slabs = (outer + 7) / 8;
    if (slabs < 1) slabs = 1;
    if (slabs > 65535) slabs = 65535;
    rps = (outer + slabs - 1) / slabs;
    dim3 g2((unsigned)col_blocks256, (unsigned)((outer + rps - 1) / rps));
    split_cols_kernel<<<g2, 256, 0, stream()>>>(x, ldx, row_mul, colmax, out.hi, out.lo, ldh, out.inv_scale, outer, inner, rps);
    SK_LAUNCH_CHECK();
  }
  return SK_OK;
}

}  // namespace sk
