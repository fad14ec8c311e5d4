// reduce.cu -- sum / mean / max / min over arbitrary axis sets, argmax / argmin.
//
// Replaces intern-table slots _SUM/_MEAN/_MAX/_MIN/_ARGMAX/_ARGMIN
// (soket/tensor/ops/intern.pyx:52-57), called as f(x, axes, dtype, out, keepdims)
// from soket/tensor/ops/forward.pyx:128-170, :224-271, backward.pyx:1061-1113
// and autodiff.pyx:84.
//
// fp32 fast paths (HBM-bound, 4 B/elem read):
//   rows : reduced axes innermost+contiguous -> warp-shuffle per row (short
//          rows) or block-per-(row,chunk) with a shared-memory tree (long rows
//          and the full reduction), two passes when a row is split;
//   cols : kept axis innermost -> float4 column tiles x row slabs, partials
//          summed by a second pass of the same kernel.
// Everything else goes through a strided one-thread-per-output kernel.
#include <float.h>
#include <math.h>

#include "common.cuh"
#include "reduce.cuh"

namespace sk {

template <int OP>
struct Red {
  __device__ __forceinline__ static float identity() {
    if (OP == SK_RED_MAX) return -INFINITY;
    if (OP == SK_RED_MIN) return INFINITY;
    return 0.f;
  }
  __device__ __forceinline__ static float combine(float a, float b) {
    if (OP == SK_RED_MAX) return (a != a) ? a : ((b != b) ? b : fmaxf(a, b));  // NaN propagates
    if (OP == SK_RED_MIN) return (a != a) ? a : ((b != b) ? b : fminf(a, b));
    return a + b;
  }
  __device__ __forceinline__ static float warp(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = combine(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
  }
  __device__ __forceinline__ static float combine4(float acc, const float4 &v) {
    return combine(acc, combine(combine(v.x, v.y), combine(v.z, v.w)));
  }
};

constexpr int kRT = 256;  // threads per block for the reduction kernels

// block-wide combine of one value per thread (kRT threads); result valid in thread 0
template <int OP>
__device__ __forceinline__ float block_reduce(float v, float *smem /* >= 8 floats */) {
  v = Red<OP>::warp(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < (kRT / 32) ? smem[lane] : Red<OP>::identity();
    v = Red<OP>::warp(v);
  }
  __syncthreads();
  return v;
}

// ---- rows, one warp per row ---------------------------------------------------
template <int OP, bool VEC>
__global__ void __launch_bounds__(kRT)
reduce_rows_warp(const float *__restrict__ in, int64_t row_stride, float *__restrict__ out,
                 int64_t R, int64_t C, float divisor) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (kRT / 32);
  for (int64_t r = (int64_t)blockIdx.x * (kRT / 32) + (threadIdx.x >> 5); r < R; r += warps_total) {
    const float *row = in + r * row_stride;
    float acc = Red<OP>::identity();
    if (VEC) {
      const float4 *row4 = reinterpret_cast<const float4 *>(row);
      const int64_t C4 = C >> 2;
      int64_t i = lane;
      // 4 independent 128-bit loads in flight per lane
      for (; i + 96 < C4; i += 128) {
        float4 v0 = ld_stream(row4 + i), v1 = ld_stream(row4 + i + 32);
        float4 v2 = ld_stream(row4 + i + 64), v3 = ld_stream(row4 + i + 96);
        float a0 = Red<OP>::combine4(Red<OP>::identity(), v0);
        float a1 = Red<OP>::combine4(Red<OP>::identity(), v1);
        float a2 = Red<OP>::combine4(Red<OP>::identity(), v2);
        float a3 = Red<OP>::combine4(Red<OP>::identity(), v3);
        acc = Red<OP>::combine(acc, Red<OP>::combine(Red<OP>::combine(a0, a1), Red<OP>::combine(a2, a3)));
      }
      for (; i < C4; i += 32) acc = Red<OP>::combine4(acc, ld_stream(row4 + i));
    } else {
      for (int64_t i = lane; i < C; i += 32) acc = Red<OP>::combine(acc, row[i]);
    }
    acc = Red<OP>::warp(acc);
    if (lane == 0) out[r] = (OP == SK_RED_MEAN) ? acc / divisor : acc;
  }
}

// ---- rows, one block per (row, chunk) --------------------------------------------
// grid = R * S blocks; block (r, s) reduces columns [s*chunk, min(C, (s+1)*chunk)).
template <int OP, bool VEC>
__global__ void __launch_bounds__(kRT)
reduce_rows_block(const float *__restrict__ in, int64_t row_stride, float *__restrict__ out,
                  int64_t R, int64_t C, int S, int64_t chunk, float divisor) {
  __shared__ float smem[8];
  const int64_t total = R * S;
  for (int64_t b = blockIdx.x; b < total; b += gridDim.x) {
    const int64_t r = b / S;
    const int s = (int)(b - r * S);
    const int64_t c0 = (int64_t)s * chunk;
    const int64_t c1 = (c0 + chunk < C) ? c0 + chunk : C;
    const float *row = in + r * row_stride;
    float acc = Red<OP>::identity();
    if (VEC) {
      const float4 *row4 = reinterpret_cast<const float4 *>(row);
      const int64_t e4 = c1 >> 2;  // chunk and c0 are multiples of 4
      int64_t i = (c0 >> 2) + threadIdx.x;
      for (; i + 3 * kRT < e4; i += 4 * kRT) {
        float4 v0 = ld_stream(row4 + i), v1 = ld_stream(row4 + i + kRT);
        float4 v2 = ld_stream(row4 + i + 2 * kRT), v3 = ld_stream(row4 + i + 3 * kRT);
        float a0 = Red<OP>::combine4(Red<OP>::identity(), v0);
        float a1 = Red<OP>::combine4(Red<OP>::identity(), v1);
        float a2 = Red<OP>::combine4(Red<OP>::identity(), v2);
        float a3 = Red<OP>::combine4(Red<OP>::identity(), v3);
        acc = Red<OP>::combine(acc, Red<OP>::combine(Red<OP>::combine(a0, a1), Red<OP>::combine(a2, a3)));
      }
      for (; i < e4; i += kRT) acc = Red<OP>::combine4(acc, ld_stream(row4 + i));
      // scalar tail of the last chunk
      for (int64_t j = (e4 << 2) + threadIdx.x; j < c1; j += kRT) acc = Red<OP>::combine(acc, row[j]);
    } else {
      for (int64_t j = c0 + threadIdx.x; j < c1; j += kRT) acc = Red<OP>::combine(acc, row[j]);
    }
    acc = block_reduce<OP>(acc, smem);
    if (threadIdx.x == 0) out[b] = (OP == SK_RED_MEAN) ? acc / divisor : acc;
  }
}

// ---- columns ------------------------------------------------------------------------
// in: (R, C) with unit column stride.  block = 32 column-groups (float4) x 8 row
// lanes; grid.x = column tiles (128 columns each), grid.y = row slabs.
// out[slab, c] partial; a second pass (R = slabs) finishes.
template <int OP, bool VEC>
__global__ void __launch_bounds__(kRT)
reduce_cols(const float *__restrict__ in, int64_t row_stride, float *__restrict__ out,
            int64_t R, int64_t C, int64_t rows_per_slab, float divisor) {
  constexpr int W = VEC ? 4 : 1;
  __shared__ float smem[8][32 * W + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = ((int64_t)blockIdx.x * 32 + tx) * W;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_slab;
  const int64_t r1 = (r0 + rows_per_slab < R) ? r0 + rows_per_slab : R;
  float acc[W];
#pragma unroll
  for (int k = 0; k < W; ++k) acc[k] = Red<OP>::identity();
  if (c < C) {
    int64_t r = r0 + ty;
    if (VEC) {
      for (; r + 24 < r1; r += 32) {
        float4 v0 = ld_stream(reinterpret_cast<const float4 *>(in + r * row_stride + c));
        float4 v1 = ld_stream(reinterpret_cast<const float4 *>(in + (r + 8) * row_stride + c));
        float4 v2 = ld_stream(reinterpret_cast<const float4 *>(in + (r + 16) * row_stride + c));
        float4 v3 = ld_stream(reinterpret_cast<const float4 *>(in + (r + 24) * row_stride + c));
        acc[0] = Red<OP>::combine(acc[0], Red<OP>::combine(Red<OP>::combine(v0.x, v1.x), Red<OP>::combine(v2.x, v3.x)));
        acc[1 % W] = Red<OP>::combine(acc[1 % W], Red<OP>::combine(Red<OP>::combine(v0.y, v1.y), Red<OP>::combine(v2.y, v3.y)));
        acc[2 % W] = Red<OP>::combine(acc[2 % W], Red<OP>::combine(Red<OP>::combine(v0.z, v1.z), Red<OP>::combine(v2.z, v3.z)));
        acc[3 % W] = Red<OP>::combine(acc[3 % W], Red<OP>::combine(Red<OP>::combine(v0.w, v1.w), Red<OP>::combine(v2.w, v3.w)));
      }
      for (; r < r1; r += 8) {
        float4 v = ld_stream(reinterpret_cast<const float4 *>(in + r * row_stride + c));
        acc[0] = Red<OP>::combine(acc[0], v.x);
        acc[1 % W] = Red<OP>::combine(acc[1 % W], v.y);
        acc[2 % W] = Red<OP>::combine(acc[2 % W], v.z);
        acc[3 % W] = Red<OP>::combine(acc[3 % W], v.w);
      }
    } else {
      for (; r < r1; r += 8) acc[0] = Red<OP>::combine(acc[0], in[r * row_stride + c]);
    }
  }
#pragma unroll
  for (int k = 0; k < W; ++k) smem[ty][tx * W + k] = acc[k];
  __syncthreads();
  // 32*W columns, 8 partials each: threads 0..32*W-1 finish
  if (threadIdx.x < 32 * W) {
    float v = smem[0][threadIdx.x];
#pragma unroll
    for (int j = 1; j < 8; ++j) v = Red<OP>::combine(v, smem[j][threadIdx.x]);
    int64_t cc = (int64_t)blockIdx.x * 32 * W + threadIdx.x;
    if (cc < C) out[(int64_t)blockIdx.y * C + cc] = (OP == SK_RED_MEAN) ? v / divisor : v;
  }
}

// ---- generic -------------------------------------------------------------------------
struct RedDesc {
  const void *in;
  void *out;
  int in_dt, out_dt, op;
  int nk, nr;  // kept dims, reduced dims
  int64_t n_out, n_red;
  int64_t kshape[SK_MAX_NDIM], kstride[SK_MAX_NDIM];
  int64_t rshape[SK_MAX_NDIM], rstride[SK_MAX_NDIM];
};

template <typename C>
__device__ __forceinline__ C red_identity(int op);
template <> __device__ __forceinline__ float red_identity<float>(int op) {
  return op == SK_RED_MAX ? -INFINITY : (op == SK_RED_MIN ? INFINITY : 0.f);
}
template <> __device__ __forceinline__ double red_identity<double>(int op) {
  return op == SK_RED_MAX ? -(double)INFINITY : (op == SK_RED_MIN ? (double)INFINITY : 0.0);
}
template <> __device__ __forceinline__ int64_t red_identity<int64_t>(int op) {
  return op == SK_RED_MAX ? INT64_MIN : (op == SK_RED_MIN ? INT64_MAX : 0);
}
template <typename C>
__device__ __forceinline__ C red_combine(int op, C a, C b) {
  if (op == SK_RED_MAX) return (a != a) ? a : ((b != b) ? b : (a > b ? a : b));
  if (op == SK_RED_MIN) return (a != a) ? a : ((b != b) ? b : (a < b ? a : b));
  return a + b;
}

template <typename C> __device__ __forceinline__ C ld_generic(const void *p, int dt, int64_t i);
// (definitions shared with ewise.cu would need rdc; keep a local copy)
template <typename C>
__device__ __forceinline__ C ld_generic(const void *p, int dt, int64_t i) {
  switch (dt) {
    case SK_BOOL: return (C)(((const uint8_t *)p)[i] != 0);
    case SK_I8: return (C)((const int8_t *)p)[i];
    case SK_U8: return (C)((const uint8_t *)p)[i];
    case SK_I16: return (C)((const int16_t *)p)[i];
    case SK_U16: return (C)((const uint16_t *)p)[i];
    case SK_I32: return (C)((const int32_t *)p)[i];
    case SK_U32: return (C)((const uint32_t *)p)[i];
    case SK_I64: return (C)((const int64_t *)p)[i];
    case SK_U64: return (C)((const uint64_t *)p)[i];
    case SK_F16: return (C)__half2float(((const __half *)p)[i]);
    case SK_BF16: return (C)__bfloat162float(((const __nv_bfloat16 *)p)[i]);
    case SK_F32: return (C)((const float *)p)[i];
    default: return (C)((const double *)p)[i];
  }
}
template <typename C>
__device__ __forceinline__ void st_generic(void *p, int dt, int64_t i, C v) {
  switch (dt) {
    case SK_BOOL: ((uint8_t *)p)[i] = (v != (C)0) ? 1 : 0; break;
    case SK_I8: ((int8_t *)p)[i] = (int8_t)(int64_t)v; break;
    case SK_U8: ((uint8_t *)p)[i] = (uint8_t)(int64_t)v; break;
    case SK_I16: ((int16_t *)p)[i] = (int16_t)(int64_t)v; break;
    case SK_U16: ((uint16_t *)p)[i] = (uint16_t)(int64_t)v; break;
    case SK_I32: ((int32_t *)p)[i] = (int32_t)(int64_t)v; break;
    case SK_U32: ((uint32_t *)p)[i] = (uint32_t)(int64_t)v; break;
    case SK_I64: ((int64_t *)p)[i] = (int64_t)v; break;
    case SK_U64: ((uint64_t *)p)[i] = (uint64_t)(int64_t)v; break;
    case SK_F16: ((__half *)p)[i] = __float2half_rn((float)v); break;
    case SK_BF16: ((__nv_bfloat16 *)p)[i] = __float2bfloat16_rn((float)v); break;
    case SK_F32: ((float *)p)[i] = (float)v; break;
    default: ((double *)p)[i] = (double)v; break;
  }
}

// one warp per output element; lanes stride over the reduced index space
template <typename C>
__global__ void __launch_bounds__(kRT) reduce_generic(const RedDesc d) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (kRT / 32);
  for (int64_t o = (int64_t)blockIdx.x * (kRT / 32) + (threadIdx.x >> 5); o < d.n_out; o += warps_total) {
    int64_t rem = o, base = 0;
#pragma unroll 1
    for (int k = d.nk - 1; k >= 0; --k) {
      int64_t q = rem / d.kshape[k];
      base += (rem - q * d.kshape[k]) * d.kstride[k];
      rem = q;
    }
    C acc = red_identity<C>(d.op);
    for (int64_t j = lane; j < d.n_red; j += 32) {
      int64_t rr = j, off = base;
#pragma unroll 1
      for (int k = d.nr - 1; k >= 0; --k) {
        int64_t q = rr / d.rshape[k];
        off += (rr - q * d.rshape[k]) * d.rstride[k];
        rr = q;
      }
      acc = red_combine<C>(d.op, acc, ld_generic<C>(d.in, d.in_dt, off));
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      C other = __shfl_xor_sync(0xffffffffu, acc, s);
      acc = red_combine<C>(d.op, acc, other);
    }
    if (lane == 0) {
      if (d.op == SK_RED_MEAN) acc = acc / (C)d.n_red;
      st_generic<C>(d.out, d.out_dt, o, acc);
    }
  }
}

// argmax / argmin: one thread per output, first occurrence wins, NaN wins (NumPy)
struct ArgDesc {
  const void *in;
  void *out;
  int in_dt, out_dt, is_min;
  int nk;
  int64_t n_out, len, axis_stride;
  int64_t kshape[SK_MAX_NDIM], kstride[SK_MAX_NDIM];
};
template <typename C>
__global__ void __launch_bounds__(kRT) argreduce_kernel(const ArgDesc d) {
  const int64_t stride = (int64_t)gridDim.x * kRT;
  for (int64_t o = (int64_t)blockIdx.x * kRT + threadIdx.x; o < d.n_out; o += stride) {
    int64_t rem = o, base = 0;
#pragma unroll 1
    for (int k = d.nk - 1; k >= 0; --k) {
      int64_t q = rem / d.kshape[k];
      base += (rem - q * d.kshape[k]) * d.kstride[k];
      rem = q;
    }
    C best = ld_generic<C>(d.in, d.in_dt, base);
    int64_t best_i = 0;
    if (!(best != best)) {
      for (int64_t j = 1; j < d.len; ++j) {
        C v = ld_generic<C>(d.in, d.in_dt, base + j * d.axis_stride);
        if (v != v) { best_i = j; break; }
        if (d.is_min ? (v < best) : (v > best)) { best = v; best_i = j; }
      }
    }
    if (d.out_dt == SK_I64) ((int64_t *)d.out)[o] = best_i;
    else ((int32_t *)d.out)[o] = (int32_t)best_i;
  }
}

// --------------------------------------------------------------------- host side
static inline bool aligned16(const void *p) { return (((uintptr_t)p) & 15) == 0; }

struct DimList {
  int n = 0;
  int64_t shape[SK_MAX_NDIM], stride[SK_MAX_NDIM];
  void push(int64_t s, int64_t st) {
    if (s == 1) return;
    if (n > 0 && stride[n - 1] == st * s) {  // merge with the previous (outer) dim
      shape[n - 1] *= s;
      stride[n - 1] = st;
      return;
    }
    shape[n] = s;
    stride[n] = st;
    ++n;
  }
  int64_t count() const {
    int64_t c = 1;
    for (int i = 0; i < n; ++i) c *= shape[i];
    return c;
  }
};

template <int OP>
static int launch_rows(const float *in, int64_t row_stride, float *out, int64_t R, int64_t C,
                       float divisor) {
  const bool vec = (C % 4 == 0) && (row_stride % 4 == 0 || R == 1) && aligned16(in);
  const int sms = ctx().num_sms;
  if (C <= 8192 && R >= (int64_t)sms * 2) {
    int grid = grid_for(R, kRT / 32, 8);
    if (vec) reduce_rows_warp<OP, true><<<grid, kRT, 0, stream()>>>(in, row_stride, out, R, C, divisor);
    else reduce_rows_warp<OP, false><<<grid, kRT, 0, stream()>>>(in, row_stride, out, R, C, divisor);
    SK_LAUNCH_CHECK();
    return SK_OK;
  }
  // block per (row, chunk): pick S so that R*S fills the machine, chunk multiple of 4096
  int64_t want_blocks = (int64_t)sms * 8;
  int64_t S = 1;
  if (R < want_blocks) {
    S = (want_blocks + R - 1) / R;
    int64_t max_s = (C + 4095) / 4096;  // at least 4096 columns per chunk
    if (S > max_s) S = max_s;
    if (S < 1) S = 1;
  }
  int64_t chunk = (C + S - 1) / S;
  chunk = (chunk + 4095) / 4096 * 4096;
  S = (C + chunk - 1) / chunk;
  if (S == 1) {
    int grid = (int)(R < want_blocks ? R : want_blocks);
    if (vec) reduce_rows_block<OP, true><<<grid, kRT, 0, stream()>>>(in, row_stride, out, R, C, 1, chunk, divisor);
    else reduce_rows_block<OP, false><<<grid, kRT, 0, stream()>>>(in, row_stride, out, R, C, 1, chunk, divisor);
    SK_LAUNCH_CHECK();
    return SK_OK;
  }
  // two passes: partial (R, S) in scratch, then finish.  MEAN divides once, at the end.
  float *partial = nullptr;
  int rc = sk_malloc((size_t)(R * S) * sizeof(float), (void **)&partial);
  if (rc) return rc;
  constexpr int OP1 = (OP == SK_RED_MEAN) ? SK_RED_SUM : OP;
  int grid = (int)(R * S < want_blocks ? R * S : want_blocks);
  if (vec) reduce_rows_block<OP1, true><<<grid, kRT, 0, stream()>>>(in, row_stride, partial, R, C, (int)S, chunk, 1.f);
  else reduce_rows_block<OP1, false><<<grid, kRT, 0, stream()>>>(in, row_stride, partial, R, C, (int)S, chunk, 1.f);
  note_launch();
  int grid2 = grid_for(R, kRT / 32, 8);
  reduce_rows_warp<OP, false><<<grid2, kRT, 0, stream()>>>(partial, S, out, R, S, divisor);
  SK_LAUNCH_CHECK();
  return sk_free(partial);
}

template <int OP>
static int launch_cols(const float *in, int64_t row_stride, float *out, int64_t R, int64_t C,
                       float divisor) {
  const bool vec = (C % 4 == 0) && (row_stride % 4 == 0) && aligned16(in) && aligned16(out);
  const int W = vec ? 4 : 1;
  const int64_t col_tiles = (C + 32 * W - 1) / (32 * W);
  int64_t want_blocks = (int64_t)ctx().num_sms * 8;
  int64_t slabs = (want_blocks + col_tiles - 1) / col_tiles;
  int64_t max_slabs = (R + 63) / 64;  // at least 64 rows per slab
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  if (slabs > 65535) slabs = 65535;
  int64_t rows_per_slab = (R + slabs - 1) / slabs;
  rows_per_slab = (rows_per_slab + 7) / 8 * 8;
  slabs = (R + rows_per_slab - 1) / rows_per_slab;
  dim3 grid((unsigned)col_tiles, (unsigned)slabs);
  if (slabs == 1) {
    if (vec) reduce_cols<OP, true><<<grid, kRT, 0, stream()>>>(in, row_stride, out, R, C, rows_per_slab, divisor);
    else reduce_cols<OP, false><<<grid, kRT, 0, stream()>>>(in, row_stride, out, R, C, rows_per_slab, divisor);
    SK_LAUNCH_CHECK();
    return SK_OK;
  }
  float *partial = nullptr;
  int rc = sk_malloc((size_t)(slabs * C) * sizeof(float), (void **)&partial);
  if (rc) return rc;
  constexpr int OP1 = (OP == SK_RED_MEAN) ? SK_RED_SUM : OP;
  if (vec) reduce_cols<OP1, true><<<grid, kRT, 0, stream()>>>(in, row_stride, partial, R, C, rows_per_slab, 1.f);
  else reduce_cols<OP1, false><<<grid, kRT, 0, stream()>>>(in, row_stride, partial, R, C, rows_per_slab, 1.f);
  note_launch();
  dim3 grid2((unsigned)col_tiles, 1);
  int64_t rps2 = (slabs + 7) / 8 * 8;
  if (vec) reduce_cols<OP, true><<<grid2, kRT, 0, stream()>>>(partial, C, out, slabs, C, rps2, divisor);
  else reduce_cols<OP, false><<<grid2, kRT, 0, stream()>>>(partial, C, out, slabs, C, rps2, divisor);
  SK_LAUNCH_CHECK();
  return sk_free(partial);
}

#define SK_RED_SWITCH(op, CALL)                 \
  switch (op) {                                 \
    case SK_RED_SUM: rc = CALL(SK_RED_SUM); break;   \
    case SK_RED_MEAN: rc = CALL(SK_RED_MEAN); break; \
    case SK_RED_MAX: rc = CALL(SK_RED_MAX); break;   \
    default: rc = CALL(SK_RED_MIN); break;           \
  }

int reduce_cols_sum_f32(const float *in, int64_t row_stride, float *out, int64_t R, int64_t C) {
  return launch_cols<SK_RED_SUM>(in, row_stride, out, R, C, 1.f);
}
int reduce_rows_f32(int op, const float *in, int64_t row_stride, float *out, int64_t R, int64_t C) {
  int rc;
  const float divisor = (float)C;
#define CALL(OP) launch_rows<OP>(in, row_stride, out, R, C, divisor)
  SK_RED_SWITCH(op, CALL)
#undef CALL
  return rc;
}

}  // namespace sk

using namespace sk;

extern "C" {

int sk_reduce(int op, const sk_array *in, uint32_t axes_mask, sk_array *out) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(in && out, "sk_reduce: null array");
  SK_REQUIRE(op >= SK_RED_SUM && op <= SK_RED_MIN, "sk_reduce: bad op %d", op);
  SK_REQUIRE(in->ndim <= SK_MAX_NDIM, "sk_reduce: ndim too large");
  DimList kept, red;
  int64_t n_out = 1, n_red = 1;
  for (int i = 0; i < in->ndim; ++i) {
    if (axes_mask & (1u << i)) { red.push(in->shape[i], in->strides[i]); n_red *= in->shape[i]; }
    else { kept.push(in->shape[i], in->strides[i]); n_out *= in->shape[i]; }
  }
  SK_REQUIRE(numel(out) == n_out, "sk_reduce: output has %lld elements, expected %lld",
             (long long)numel(out), (long long)n_out);
  SK_REQUIRE(is_contiguous(out), "sk_reduce: output must be contiguous");
  if (n_out == 0) return SK_OK;
  if (n_red == 0) {
    SK_REQUIRE(op == SK_RED_SUM || op == SK_RED_MEAN, "zero-size reduction has no identity for max/min");
    return sk_fill(out, op == SK_RED_MEAN ? NAN : 0.0, 0, 0);
  }

  if (in->dtype == SK_F32 && out->dtype == SK_F32) {
    const float *ip = (const float *)in->data;
    float *op_ = (float *)out->data;
    const float divisor = (float)n_red;
    // rows: reduced dims collapse to one unit-stride dim; kept dims to <= 1 dim
    if (red.n <= 1 && kept.n <= 1 && (red.n == 0 || red.stride[0] == 1)) {
      int64_t R = kept.n ? kept.shape[0] : 1, C = red.n ? red.shape[0] : 1;
      int64_t rs = kept.n ? kept.stride[0] : 0;
      if (red.n == 0) {  // nothing reduced (all reduced axes have size 1): copy
        sk_array src = *in;
        src.ndim = out->ndim;
        int j = 0;
        for (int i = 0; i < in->ndim; ++i)
          if (!(axes_mask & (1u << i))) { src.shape[j] = in->shape[i]; src.strides[j] = in->strides[i]; ++j; }
        return sk_copy(&src, out);
      }
      if (rs >= 0) {
        ProfScope ps(SK_PROF_REDUCE, (double)R * C * 4.0);
#define CALL(OP) launch_rows<OP>(ip, rs, op_, R, C, divisor)
        SK_RED_SWITCH(op, CALL)
#undef CALL
        return rc;
      }
    }
    // cols: kept dim has unit stride, one reduced dim with a positive row stride
    if (red.n == 1 && kept.n == 1 && kept.stride[0] == 1 && red.stride[0] >= kept.shape[0]) {
      ProfScope ps(SK_PROF_REDUCE, (double)red.shape[0] * kept.shape[0] * 4.0);
#define CALL(OP) launch_cols<OP>(ip, red.stride[0], op_, red.shape[0], kept.shape[0], divisor)
      SK_RED_SWITCH(op, CALL)
#undef CALL
      return rc;
    }
  }

  RedDesc d;
  memset(&d, 0, sizeof(d));
  d.in = in->data; d.out = out->data;
  d.in_dt = in->dtype; d.out_dt = out->dtype; d.op = op;
  d.nk = kept.n; d.nr = red.n;
  d.n_out = n_out; d.n_red = n_red;
  for (int i = 0; i < kept.n; ++i) { d.kshape[i] = kept.shape[i]; d.kstride[i] = kept.stride[i]; }
  for (int i = 0; i < red.n; ++i) { d.rshape[i] = red.shape[i]; d.rstride[i] = red.stride[i]; }
  int grid = grid_for(n_out, kRT / 32, 8);
  // accumulate in the class of the OUTPUT for sum/mean (NumPy: dtype= selects the
  // accumulator), of the INPUT for max/min
  int probe = (op == SK_RED_MAX || op == SK_RED_MIN) ? in->dtype : out->dtype;
  if (probe == SK_F64) reduce_generic<double><<<grid, kRT, 0, stream()>>>(d);
  else if (dtype_is_float(probe)) reduce_generic<float><<<grid, kRT, 0, stream()>>>(d);
  else reduce_generic<int64_t><<<grid, kRT, 0, stream()>>>(d);
  SK_LAUNCH_CHECK();
  return SK_OK;
}

int sk_argreduce(int is_min, const sk_array *in, int axis, sk_array *out) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(in && out, "sk_argreduce: null array");
  SK_REQUIRE(out->dtype == SK_I64 || out->dtype == SK_I32, "sk_argreduce: out must be int64/int32");
  SK_REQUIRE(is_contiguous(out), "sk_argreduce: output must be contiguous");
  ArgDesc d;
  memset(&d, 0, sizeof(d));
  d.in = in->data; d.out = out->data;
  d.in_dt = in->dtype; d.out_dt = out->dtype; d.is_min = is_min;
  if (axis < 0) {
    SK_REQUIRE(is_contiguous(in), "sk_argreduce: flattened argreduce needs a contiguous input");
    d.nk = 0; d.n_out = 1; d.len = numel(in); d.axis_stride = 1;
  } else {
    SK_REQUIRE(axis < in->ndim, "sk_argreduce: axis %d out of range", axis);
    DimList kept;
    int64_t n_out = 1;
    for (int i = 0; i < in->ndim; ++i)
      if (i != axis) { kept.push(in->shape[i], in->strides[i]); n_out *= in->shape[i]; }
    d.nk = kept.n; d.n_out = n_out;
    for (int i = 0; i < kept.n; ++i) { d.kshape[i] = kept.shape[i]; d.kstride[i] = kept.stride[i]; }
    d.len = in->shape[axis]; d.axis_stride = in->strides[axis];
  }
  SK_REQUIRE(numel(out) == d.n_out, "sk_argreduce: output size mismatch");
  SK_REQUIRE(d.len > 0, "sk_argreduce: attempt to get argmax of an empty sequence");
  if (d.n_out == 0) return SK_OK;
  int grid = grid_for(d.n_out, kRT, 8);
  if (in->dtype == SK_F64) argreduce_kernel<double><<<grid, kRT, 0, stream()>>>(d);
  else if (dtype_is_float(in->dtype)) argreduce_kernel<float><<<grid, kRT, 0, stream()>>>(d);
  else argreduce_kernel<int64_t><<<grid, kRT, 0, stream()>>>(d);
  SK_LAUNCH_CHECK();
  return SK_OK;
}

}  // extern "C"
