// ewise.cu -- elementwise / scalar / unary / compare / copy(cast, compaction) / fill.
//
// Replaces the intern-table slots _ADD.._POW, _NEG, _MAXIMUM, _EXP, _LOG,
// _EQUAL.._LESS_EQUAL, _ARRAY/_COPY and the materialisation of _BCASTTO /
// _TRANSPOSE / _RESHAPE views (soket/tensor/ops/intern.pyx:45-76; call shapes
// soket/tensor/ops/forward.pyx:7-221).
//
// Two tiers:
//   * fp32 fast paths (the hot path): contiguous float4 streaming kernels and a
//     2-D row/column-broadcast kernel -- HBM-bound, 12 B/elem (binary),
//     8 B/elem (scalar / unary / broadcast operand), 128-bit accesses, 4 loads
//     in flight per operand per thread, grid = multiple of the SM count.
//   * a generic strided kernel for every other dtype / stride combination
//     (bit-exact compaction, casts, comparisons, weak-scalar semantics).
#include <math.h>

#include <stdlib.h>
#include "common.cuh"

namespace sk {

// ------------------------------------------------------------------ functors
template <int OP>
struct BinF32 {
  __device__ __forceinline__ static float apply(float a, float b) {
    if (OP == SK_OP_ADD) return a + b;
    if (OP == SK_OP_SUB) return a - b;
    if (OP == SK_OP_MUL) return a * b;
    if (OP == SK_OP_DIV) return a / b;
    if (OP == SK_OP_POW) return powf(a, b);
    // np.maximum / np.minimum propagate NaN
    if (OP == SK_OP_MAXIMUM) return (a != a || b != b) ? (a + b) : fmaxf(a, b);
    if (OP == SK_OP_MINIMUM) return (a != a || b != b) ? (a + b) : fminf(a, b);
    return 0.f;
  }
};

template <int UOP>
struct UnF32 {
  __device__ __forceinline__ static float apply(float a) {
    if (UOP == SK_UOP_NEG) return -a;
    if (UOP == SK_UOP_EXP) return expf(a);
    if (UOP == SK_UOP_LOG) return logf(a);
    if (UOP == SK_UOP_SQRT) return sqrtf(a);
    if (UOP == SK_UOP_RELU) return (a != a) ? a : fmaxf(a, 0.f);
    if (UOP == SK_UOP_ABS) return fabsf(a);
    return 0.f;
  }
};

// Scalar-exponent power: exact forms for the exponents the reference uses
// (forward.pyx:299 `pow(xs, 2)`, :321-324 `pow(var+eps, -0.5)`,
// optim.pyx:262 `pow(v, 0.5)`).
enum { POW_GENERIC = 0, POW_SQUARE, POW_SQRT, POW_RSQRT, POW_RECIP, POW_ID };
__device__ __forceinline__ float pow_special(float a, float e, int kind) {
  switch (kind) {
    case POW_SQUARE: return a * a;
    case POW_SQRT: return sqrtf(a);
    case POW_RSQRT: return 1.0f / sqrtf(a);
    case POW_RECIP: return 1.0f / a;
    case POW_ID: return a;
    default: return powf(a, e);
  }
}
__device__ __forceinline__ int pow_kind_dev(float e) {
  if (e == 2.0f) return POW_SQUARE;
  if (e == 0.5f) return POW_SQRT;
  if (e == -0.5f) return POW_RSQRT;
  if (e == -1.0f) return POW_RECIP;
  if (e == 1.0f) return POW_ID;
  return POW_GENERIC;
}
static int pow_kind(double e) {
  if (e == 2.0) return POW_SQUARE;
  if (e == 0.5) return POW_SQRT;
  if (e == -0.5) return POW_RSQRT;
  if (e == -1.0) return POW_RECIP;
  if (e == 1.0) return POW_ID;
  return POW_GENERIC;
}

// ------------------------------------------------------------- fp32 fast paths
constexpr int kThreads = 256;
constexpr int kUnroll = 4;

template <int OP>
__global__ void __launch_bounds__(kThreads)
binary_f32_vec(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ out,
               int64_t n) {
  const int64_t n4 = n >> 2;
  const float4 *a4 = reinterpret_cast<const float4 *>(a);
  const float4 *b4 = reinterpret_cast<const float4 *>(b);
  float4 *o4 = reinterpret_cast<float4 *>(out);
  const int64_t stride = (int64_t)gridDim.x * kThreads * kUnroll;
  for (int64_t base = (int64_t)blockIdx.x * kThreads * kUnroll + threadIdx.x; base < n4;
       base += stride) {
    float4 va[kUnroll], vb[kUnroll];
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      int64_t i = base + (int64_t)j * kThreads;
      if (i < n4) {
        va[j] = ld_stream(a4 + i);
        vb[j] = ld_stream(b4 + i);
      }
    }
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      int64_t i = base + (int64_t)j * kThreads;
      if (i < n4) {
        float4 r;
        r.x = BinF32<OP>::apply(va[j].x, vb[j].x);
        r.y = BinF32<OP>::apply(va[j].y, vb[j].y);
        r.z = BinF32<OP>::apply(va[j].z, vb[j].z);
        r.w = BinF32<OP>::apply(va[j].w, vb[j].w);
        st_stream(o4 + i, r);
      }
    }
  }
  // tail (< 4 elements)
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    int64_t i = (n4 << 2) + threadIdx.x;
    out[i] = BinF32<OP>::apply(a[i], b[i]);
  }
}

// out = a (op) s  or  s (op) a
template <int OP, bool REVERSE>
__global__ void __launch_bounds__(kThreads)
scalar_f32_vec(const float *__restrict__ a, float s, int powk, float *__restrict__ out, int64_t n) {
  const int64_t n4 = n >> 2;
  const float4 *a4 = reinterpret_cast<const float4 *>(a);
  float4 *o4 = reinterpret_cast<float4 *>(out);
  auto f = [&](float x) -> float {
    if (OP == SK_OP_POW && !REVERSE) return pow_special(x, s, powk);
    return REVERSE ? BinF32<OP>::apply(s, x) : BinF32<OP>::apply(x, s);
  };
  const int64_t stride = (int64_t)gridDim.x * kThreads * kUnroll;
  for (int64_t base = (int64_t)blockIdx.x * kThreads * kUnroll + threadIdx.x; base < n4;
       base += stride) {
    float4 va[kUnroll];
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      int64_t i = base + (int64_t)j * kThreads;
      if (i < n4) va[j] = ld_stream(a4 + i);
    }
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      int64_t i = base + (int64_t)j * kThreads;
      if (i < n4) {
        float4 r;
        r.x = f(va[j].x); r.y = f(va[j].y); r.z = f(va[j].z); r.w = f(va[j].w);
        st_stream(o4 + i, r);
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    int64_t i = (n4 << 2) + threadIdx.x;
    out[i] = f(a[i]);
  }
}

template <int UOP>
__global__ void __launch_bounds__(kThreads)
unary_f32_vec(const float *__restrict__ a, float *__restrict__ out, int64_t n) {
  const int64_t n4 = n >> 2;
  const float4 *a4 = reinterpret_cast<const float4 *>(a);
  float4 *o4 = reinterpret_cast<float4 *>(out);
  const int64_t stride = (int64_t)gridDim.x * kThreads * kUnroll;
  for (int64_t base = (int64_t)blockIdx.x * kThreads * kUnroll + threadIdx.x; base < n4;
       base += stride) {
    float4 va[kUnroll];
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      int64_t i = base + (int64_t)j * kThreads;
      if (i < n4) va[j] = ld_stream(a4 + i);
    }
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      int64_t i = base + (int64_t)j * kThreads;
      if (i < n4) {
        float4 r;
        r.x = UnF32<UOP>::apply(va[j].x); r.y = UnF32<UOP>::apply(va[j].y);
        r.z = UnF32<UOP>::apply(va[j].z); r.w = UnF32<UOP>::apply(va[j].w);
        st_stream(o4 + i, r);
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    int64_t i = (n4 << 2) + threadIdx.x;
    out[i] = UnF32<UOP>::apply(a[i]);
  }
}

// 2-D broadcast: out (R, C) contiguous; each operand is either a full (R, C)
// contiguous matrix, a row vector (stride (0,1)), a column vector (stride (1,0)
// or (s,0)) or a scalar (0,0).  C % 4 == 0.  Covers bias add (B,H)+(H,)
// (prototypes.pyx:113), x - mean(B,1), xs * rstd(B,1), gamma(H,) * norm
// (forward.pyx:296-353).
template <int OP>
__global__ void __launch_bounds__(kThreads)
binary_f32_bcast2d(const float *__restrict__ a, int64_t a_s0, int a_s1,
                   const float *__restrict__ b, int64_t b_s0, int b_s1,
                   float *__restrict__ out, int64_t R, int64_t C4) {
  const int64_t total = R * C4;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  float4 *o4 = reinterpret_cast<float4 *>(out);
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += stride) {
    int64_t r = i / C4;
    int64_t c = (i - r * C4) << 2;
    auto load_operand = [&](const float *p, int64_t s0, int s1) -> float4 {
      if (s1) {
        if (s0) return ld_stream(reinterpret_cast<const float4 *>(p + r * s0 + c));  // full matrix
        return __ldg(reinterpret_cast<const float4 *>(p + c));  // row vector: keep cached
      }
      float s = __ldg(p + r * s0);  // column vector / scalar
      return make_float4(s, s, s, s);
    };
    const float4 va = load_operand(a, a_s0, a_s1);
    const float4 vb = load_operand(b, b_s0, b_s1);
    float4 r4;
    r4.x = BinF32<OP>::apply(va.x, vb.x);
    r4.y = BinF32<OP>::apply(va.y, vb.y);
    r4.z = BinF32<OP>::apply(va.z, vb.z);
    r4.w = BinF32<OP>::apply(va.w, vb.w);
    st_stream(o4 + i, r4);
  }
}

// relu backward in one pass: out = (x > 0) * adj   (backward.pyx:849-874)
__global__ void __launch_bounds__(kThreads)
relu_bwd_f32_vec(const float *__restrict__ x, const float *__restrict__ adj,
                 float *__restrict__ out, int64_t n) {
  const int64_t n4 = n >> 2;
  const float4 *x4 = reinterpret_cast<const float4 *>(x);
  const float4 *g4 = reinterpret_cast<const float4 *>(adj);
  float4 *o4 = reinterpret_cast<float4 *>(out);
  const int64_t stride = (int64_t)gridDim.x * kThreads * kUnroll;
  for (int64_t base = (int64_t)blockIdx.x * kThreads * kUnroll + threadIdx.x; base < n4;
       base += stride) {
    float4 vx[kUnroll], vg[kUnroll];
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      int64_t i = base + (int64_t)j * kThreads;
      if (i < n4) { vx[j] = ld_stream(x4 + i); vg[j] = ld_stream(g4 + i); }
    }
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      int64_t i = base + (int64_t)j * kThreads;
      if (i < n4) {
        float4 r;
        r.x = (vx[j].x > 0.f ? 1.f : 0.f) * vg[j].x;
        r.y = (vx[j].y > 0.f ? 1.f : 0.f) * vg[j].y;
        r.z = (vx[j].z > 0.f ? 1.f : 0.f) * vg[j].z;
        r.w = (vx[j].w > 0.f ? 1.f : 0.f) * vg[j].w;
        st_stream(o4 + i, r);
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    int64_t i = (n4 << 2) + threadIdx.x;
    out[i] = (x[i] > 0.f ? 1.f : 0.f) * adj[i];
  }
}

// fill (fp32 contiguous)
__global__ void __launch_bounds__(kThreads) fill_f32_vec(float *__restrict__ out, float v, int64_t n) {
  const int64_t n4 = n >> 2;
  float4 *o4 = reinterpret_cast<float4 *>(out);
  const float4 v4 = make_float4(v, v, v, v);
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride)
    st_stream(o4 + i, v4);
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) out[(n4 << 2) + threadIdx.x] = v;
}

// ---------------------------------------------------------------- generic path
struct GenDesc {
  const void *a;
  const void *b;
  void *out;
  int a_dt, b_dt, out_dt;
  int op;    // sk_binary_op / sk_unary_op / -1 copy
  int mode;  // 0: a op b ; 1: a op scalar ; 2: scalar op a ; 3: unary(a) ; 4: copy/cast ; 5: fill
  int ndim;
  int powk;
  int64_t n;
  int64_t shape[SK_MAX_NDIM];
  int64_t sa[SK_MAX_NDIM], sb[SK_MAX_NDIM], so[SK_MAX_NDIM];
  double fscalar;
  int64_t iscalar;
};

template <typename C>
__device__ __forceinline__ C load_as(const void *p, int dt, int64_t i) {
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
__device__ __forceinline__ void store_from(void *p, int dt, int64_t i, C v) {
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

__device__ __forceinline__ int64_t ipow(int64_t a, int64_t e) {
  if (e < 0) return (a == 1) ? 1 : ((a == -1) ? ((e & 1) ? -1 : 1) : 0);
  int64_t r = 1;
  while (e) {
    if (e & 1) r *= a;
    a *= a;
    e >>= 1;
  }
  return r;
}

template <typename C>
__device__ __forceinline__ C gen_binary(int op, C a, C b);
template <>
__device__ __forceinline__ float gen_binary<float>(int op, float a, float b) {
  switch (op) {
    case SK_OP_ADD: return a + b;
    case SK_OP_SUB: return a - b;
    case SK_OP_MUL: return a * b;
    case SK_OP_DIV: return a / b;
    case SK_OP_POW: return powf(a, b);
    case SK_OP_MAXIMUM: return (a != a || b != b) ? (a + b) : fmaxf(a, b);
    case SK_OP_MINIMUM: return (a != a || b != b) ? (a + b) : fminf(a, b);
    default: return 0.f;
  }
}
template <>
__device__ __forceinline__ double gen_binary<double>(int op, double a, double b) {
  switch (op) {
    case SK_OP_ADD: return a + b;
    case SK_OP_SUB: return a - b;
    case SK_OP_MUL: return a * b;
    case SK_OP_DIV: return a / b;
    case SK_OP_POW: return pow(a, b);
    case SK_OP_MAXIMUM: return (a != a || b != b) ? (a + b) : fmax(a, b);
    case SK_OP_MINIMUM: return (a != a || b != b) ? (a + b) : fmin(a, b);
    default: return 0.0;
  }
}
template <>
__device__ __forceinline__ int64_t gen_binary<int64_t>(int op, int64_t a, int64_t b) {
  switch (op) {
    case SK_OP_ADD: return a + b;
    case SK_OP_SUB: return a - b;
    case SK_OP_MUL: return a * b;
    case SK_OP_DIV: {  // floor division (only reached for integer out dtype)
      if (b == 0) return 0;
      int64_t q = a / b;
      if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
      return q;
    }
    case SK_OP_POW: return ipow(a, b);
    case SK_OP_MAXIMUM: return a > b ? a : b;
    case SK_OP_MINIMUM: return a < b ? a : b;
    default: return 0;
  }
}

template <typename C>
__device__ __forceinline__ bool gen_compare(int op, C a, C b) {
  switch (op) {
    case SK_OP_EQ: return a == b;
    case SK_OP_NE: return a != b;
    case SK_OP_GT: return a > b;
    case SK_OP_GE: return a >= b;
    case SK_OP_LT: return a < b;
    default: return a <= b;
  }
}

template <typename C>
__device__ __forceinline__ C gen_unary(int op, C a);
template <>
__device__ __forceinline__ float gen_unary<float>(int op, float a) {
  switch (op) {
    case SK_UOP_NEG: return -a;
    case SK_UOP_EXP: return expf(a);
    case SK_UOP_LOG: return logf(a);
    case SK_UOP_SQRT: return sqrtf(a);
    case SK_UOP_RELU: return (a != a) ? a : fmaxf(a, 0.f);
    default: return fabsf(a);
  }
}
template <>
__device__ __forceinline__ double gen_unary<double>(int op, double a) {
  switch (op) {
    case SK_UOP_NEG: return -a;
    case SK_UOP_EXP: return exp(a);
    case SK_UOP_LOG: return log(a);
    case SK_UOP_SQRT: return sqrt(a);
    case SK_UOP_RELU: return (a != a) ? a : fmax(a, 0.0);
    default: return fabs(a);
  }
}
template <>
__device__ __forceinline__ int64_t gen_unary<int64_t>(int op, int64_t a) {
  switch (op) {
    case SK_UOP_NEG: return -a;
    case SK_UOP_RELU: return a > 0 ? a : 0;
    case SK_UOP_ABS: return a < 0 ? -a : a;
    default: return a;
  }
}

template <typename C>
__device__ __forceinline__ C scalar_of(const GenDesc &d);
template <> __device__ __forceinline__ float scalar_of<float>(const GenDesc &d) { return (float)d.fscalar; }
template <> __device__ __forceinline__ double scalar_of<double>(const GenDesc &d) { return d.fscalar; }
template <> __device__ __forceinline__ int64_t scalar_of<int64_t>(const GenDesc &d) { return d.iscalar; }

// One output element per thread-iteration, consecutive threads -> consecutive
// (row-major) output elements, so contiguous outputs / unit-inner-stride inputs
// coalesce.
template <typename C>
__global__ void __launch_bounds__(kThreads) generic_ewise(const GenDesc d) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  const bool is_cmp = d.op >= SK_OP_EQ && d.mode <= 2;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < d.n; i += stride) {
    int64_t rem = i, oa = 0, ob = 0, oo = 0;
#pragma unroll 1
    for (int k = d.ndim - 1; k >= 0; --k) {
      int64_t q = rem / d.shape[k];
      int64_t idx = rem - q * d.shape[k];
      rem = q;
      oa += idx * d.sa[k];
      ob += idx * d.sb[k];
      oo += idx * d.so[k];
    }
    C r;
    if (d.mode == 4) {
      r = load_as<C>(d.a, d.a_dt, oa);
    } else if (d.mode == 5) {
      r = scalar_of<C>(d);
    } else if (d.mode == 3) {
      r = gen_unary<C>(d.op, load_as<C>(d.a, d.a_dt, oa));
    } else {
      C x, y;
      if (d.mode == 0) { x = load_as<C>(d.a, d.a_dt, oa); y = load_as<C>(d.b, d.b_dt, ob); }
      else if (d.mode == 1) { x = load_as<C>(d.a, d.a_dt, oa); y = scalar_of<C>(d); }
      else { x = scalar_of<C>(d); y = load_as<C>(d.a, d.a_dt, oa); }
      if (is_cmp) r = (C)gen_compare<C>(d.op, x, y);
      else r = gen_binary<C>(d.op, x, y);
    }
    store_from<C>(d.out, d.out_dt, oo, r);
  }
}

// float->float copies with identical dtype must be bit-exact (NaN payloads,
// -0.0, subnormals): move raw bytes instead of converting.
template <typename T>
__global__ void __launch_bounds__(kThreads) generic_copy_raw(const GenDesc d) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  const T *src = (const T *)d.a;
  T *dst = (T *)d.out;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < d.n; i += stride) {
    int64_t rem = i, oa = 0, oo = 0;
#pragma unroll 1
    for (int k = d.ndim - 1; k >= 0; --k) {
      int64_t q = rem / d.shape[k];
      int64_t idx = rem - q * d.shape[k];
      rem = q;
      oa += idx * d.sa[k];
      oo += idx * d.so[k];
    }
    dst[oo] = src[oa];
  }
}

// 2-D transpose-style compaction through shared memory: src has unit stride
// along dim0 of the collapsed space, dst along dim1 (e.g. materialising x.T).
// Both sides coalesced; 8 B/elem.
template <typename T>
__global__ void __launch_bounds__(256)
transpose_tiled(const T *__restrict__ src, T *__restrict__ dst, int64_t R, int64_t C,
                int64_t src_col_stride, int64_t dst_row_stride) {
  // dst[r, c] = src[r + c * src_col_stride]  (src unit stride along r)
  __shared__ T tile[32][33];
  const int64_t tiles_c = (C + 31) / 32;
  const int64_t tiles_r = (R + 31) / 32;
  const int64_t total = tiles_c * tiles_r;
  for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
    int64_t tr = t / tiles_c, tc = t - tr * tiles_c;
    int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      int64_t c = tc * 32 + ty + j;
      int64_t r = tr * 32 + tx;
      if (r < R && c < C) tile[ty + j][tx] = src[r + c * src_col_stride];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      int64_t r = tr * 32 + ty + j;
      int64_t c = tc * 32 + tx;
      if (r < R && c < C) dst[r * dst_row_stride + c] = tile[tx][ty + j];
    }
    __syncthreads();
  }
}


// 2-D compaction with contiguous destination rows, 4-byte elements: dst[r, c] =
// src[r * lds + c * sin] with sin in {0, 1} and any lds (0 = every row the same):
// materialising broadcast_to views (forward.pyx:120-126; zero strides) and pitched
// slices.  128-bit stores; TPR (a power of two) threads walk one row, no index division.
// Algorithmic traffic: 4 B/elem written + the source once.
template <int SIN>
__global__ void __launch_bounds__(256)
copy_rows_u32(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, int64_t R, int64_t C,
              int64_t lds, int64_t ldd, int tpr_log2, int vec_src) {
  const int tpr = 1 << tpr_log2;
  const int tx = threadIdx.x & (tpr - 1), ty = threadIdx.x >> tpr_log2;
  const int rpb = 256 >> tpr_log2;
  const int64_t C4 = C >> 2;
  for (int64_t r = (int64_t)blockIdx.x * rpb + ty; r < R; r += (int64_t)gridDim.x * rpb) {
    const uint32_t *srow = src + r * lds;
    uint4 *drow = reinterpret_cast<uint4 *>(dst + r * ldd);
    if (SIN == 0) {
      const uint32_t v = __ldg(srow);
      const uint4 v4 = make_uint4(v, v, v, v);
#pragma unroll 4
      for (int64_t c4 = tx; c4 < C4; c4 += tpr) drow[c4] = v4;
    } else if (vec_src) {
      const uint4 *s4 = reinterpret_cast<const uint4 *>(srow);
#pragma unroll 4
      for (int64_t c4 = tx; c4 < C4; c4 += tpr) drow[c4] = __ldg(s4 + c4);
    } else {
#pragma unroll 2
      for (int64_t c4 = tx; c4 < C4; c4 += tpr) {
        const uint32_t *q = srow + c4 * 4;
        drow[c4] = make_uint4(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3));
      }
    }
  }
}

// 64 x 64 tiled transpose of 4-byte elements with 128-bit global accesses on both sides.
// dst[r * ldd + c] = src[c * lds + r]  (src unit stride along r, dst along c).
// Shared tile: 64 "c" rows of 16 uint4 (= 64 r values); the uint4 slot of (c, r) is
// (r / 4) ^ ((c / 4) & 7), which makes both the 16-byte writes of the load phase and the
// 4-byte reads of the store phase bank-conflict free.  Requires R, C multiples of 4,
// lds, ldd multiples of 4 and 16-byte aligned bases (else transpose_tiled).
__global__ void __launch_bounds__(256)
transpose64_u32(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, int64_t R, int64_t C,
                int64_t lds, int64_t ldd) {
  __shared__ uint4 tile[64][16];
  const int64_t tiles_r = (R + 63) / 64, tiles_c = (C + 63) / 64;
  const int64_t total = tiles_r * tiles_c;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // load: half-warp = one source column (16 x uint4 = 64 r); a warp covers 2 columns per j
  const int r4 = lane & 15, cl0 = warp * 8 + (lane >> 4);
  uint4 v[4];
  auto fetch = [&](int64_t t) {
    const int64_t tc = t / tiles_r, tr = t - tc * tiles_r;
    const int64_t r = tr * 64 + r4 * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t c = tc * 64 + cl0 + j * 2;
      v[j] = (c < C && r < R) ? __ldg(reinterpret_cast<const uint4 *>(src + c * lds + r)) : make_uint4(0, 0, 0, 0);
    }
  };
  int64_t t = blockIdx.x;
  if (t < total) fetch(t);
  for (; t < total; t += gridDim.x) {
    const int64_t tc = t / tiles_r, tr = t - tc * tiles_r;
    const int64_t r0 = tr * 64, c0 = tc * 64;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cl = cl0 + j * 2;
      tile[cl][r4 ^ ((cl >> 2) & 7)] = v[j];
    }
    __syncthreads();
    // the next tile's loads fly while this one is written out (latency-bound otherwise:
    // ncu showed 68 % of stall cycles waiting on the global loads at 3.9 TB/s)
    if (t + gridDim.x < total) fetch(t + gridDim.x);
    // store: a warp covers 4 consecutive r x 8 uint4 of c (32 c); 2 x 16 such pieces per tile
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int piece = warp * 4 + j;             // 0..31: (r block of 4) x (c half)
      const int rl = (piece >> 1) * 4 + (lane & 3);
      const int c4 = (piece & 1) * 8 + (lane >> 2);
      const uint32_t *tw = reinterpret_cast<const uint32_t *>(&tile[0][0]);
      uint32_t e[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int cl = c4 * 4 + k;
        e[k] = tw[(cl * 16 + ((rl >> 2) ^ ((cl >> 2) & 7))) * 4 + (rl & 3)];
      }
      const int64_t r = r0 + rl, c = c0 + c4 * 4;
      if (r < R && c < C) *reinterpret_cast<uint4 *>(dst + r * ldd + c) = make_uint4(e[0], e[1], e[2], e[3]);
    }
    __syncthreads();
  }
}

// ----------------------------------------------------------------- host side
static inline bool aligned16(const void *p) { return (((uintptr_t)p) & 15) == 0; }

static int compute_class_for(int out_dt) {
  // 0: float, 1: double, 2: int64
  if (out_dt == SK_F32 || out_dt == SK_F16 || out_dt == SK_BF16) return 0;
  if (out_dt == SK_F64) return 1;
  return 2;
}

static int launch_generic(GenDesc &d, int cls) {
  if (d.n == 0) return SK_OK;
  int grid = grid_for(d.n, kThreads, 16);
  if (cls == 0) generic_ewise<float><<<grid, kThreads, 0, stream()>>>(d);
  else if (cls == 1) generic_ewise<double><<<grid, kThreads, 0, stream()>>>(d);
  else generic_ewise<int64_t><<<grid, kThreads, 0, stream()>>>(d);
  SK_LAUNCH_CHECK();
  return SK_OK;
}

static int check_array(const sk_array *a, const char *name) {
  SK_REQUIRE(a != nullptr, "%s: null array", name);
  SK_REQUIRE(a->ndim >= 0 && a->ndim <= SK_MAX_NDIM, "%s: ndim %d out of range", name, a->ndim);
  SK_REQUIRE(a->dtype >= SK_BOOL && a->dtype <= SK_BF16, "%s: bad dtype %d", name, a->dtype);
  for (int i = 0; i < a->ndim; ++i) SK_REQUIRE(a->shape[i] >= 0, "%s: negative dim", name);
  return SK_OK;
}

template <int N>
static void fill_desc(GenDesc &d, const Collapsed<N> &c, int ia, int ib, int io) {
  d.ndim = c.ndim;
  d.n = 1;
  for (int i = 0; i < c.ndim; ++i) {
    d.shape[i] = c.shape[i];
    d.n *= c.shape[i];
    d.sa[i] = ia >= 0 ? c.strides[ia][i] : 0;
    d.sb[i] = ib >= 0 ? c.strides[ib][i] : 0;
    d.so[i] = io >= 0 ? c.strides[io][i] : 0;
  }
}

template <template <int> class K>
struct DispatchBinary;


// ------------------------------------------------------------------ fused elementwise chains (lazy mode)
// One thread walks the accumulator program for U groups of W elements.  The program is uniform across
// the grid, so the per-step switch costs no divergence; what is saved is one HBM round trip per fused
// node (8-12 B/elem each).  Every arithmetic step calls the SAME functors as the unfused kernels.
template <int W>
struct FVec { float v[W]; };

template <int W>
__device__ __forceinline__ FVec<W> fused_operand(const sk_fused_program &p, int k, int64_t e, const FVec<W> (&tmp)[3]) {
  FVec<W> o;
  const int src = p.src[k], idx = p.idx[k];
  if (src == SK_F_CONST) {
#pragma unroll
    for (int c = 0; c < W; ++c) o.v[c] = p.cst[k];
  } else if (src == SK_F_TEMP) {
    o = idx == 0 ? tmp[0] : (idx == 1 ? tmp[1] : tmp[2]);
  } else {
    const float *in = p.in[idx];
    const int kind = p.in_kind[idx];
    if (kind == SK_F_SINGLE) {
      const float s = __ldg(in);
#pragma unroll
      for (int c = 0; c < W; ++c) o.v[c] = s;
    } else {
      const int64_t at = kind == SK_F_FULL ? e : e % p.cols;
      if (W == 4) {
        const float4 q = kind == SK_F_FULL ? ld_stream(reinterpret_cast<const float4 *>(in + at))
                                           : __ldg(reinterpret_cast<const float4 *>(in + at));
        o.v[0] = q.x; o.v[1 % W] = q.y; o.v[2 % W] = q.z; o.v[3 % W] = q.w;
      } else {
        o.v[0] = in[at];
      }
    }
  }
  return o;
}

template <int OP, int W>
__device__ __forceinline__ void fused_bin(FVec<W> &acc, const FVec<W> &o, bool rev, float cst, bool scalar_pow) {
#pragma unroll
  for (int c = 0; c < W; ++c) {
    if (OP == SK_OP_POW && scalar_pow && !rev) acc.v[c] = pow_special(acc.v[c], cst, pow_kind_dev(cst));
    else acc.v[c] = rev ? BinF32<OP>::apply(o.v[c], acc.v[c]) : BinF32<OP>::apply(acc.v[c], o.v[c]);
  }
}

template <int W, int U>
__global__ void __launch_bounds__(kThreads)
fused_ewise_kernel(const __grid_constant__ sk_fused_program p, float *__restrict__ out) {
  const int64_t groups = p.n / W;
  const int64_t stride = (int64_t)gridDim.x * kThreads * U;
  for (int64_t base = (int64_t)blockIdx.x * kThreads * U + threadIdx.x; base < groups; base += stride) {
    FVec<W> acc[U], tmp[U][3];
#pragma unroll
    for (int j = 0; j < U; ++j)
#pragma unroll
      for (int c = 0; c < W; ++c) { acc[j].v[c] = 0.f; tmp[j][0].v[c] = 0.f; tmp[j][1].v[c] = 0.f; tmp[j][2].v[c] = 0.f; }
    for (int k = 0; k < p.n_ops; ++k) {
      const int code = p.code[k], sub = p.sub[k];
      if (code == SK_F_STORE) {
        const int idx = p.idx[k];
#pragma unroll
        for (int j = 0; j < U; ++j) {
          if (idx == 0) tmp[j][0] = acc[j];
          else if (idx == 1) tmp[j][1] = acc[j];
          else tmp[j][2] = acc[j];
        }
        continue;
      }
      if (code == SK_F_UN) {
#pragma unroll
        for (int j = 0; j < U; ++j)
#pragma unroll
          for (int c = 0; c < W; ++c) {
            float a = acc[j].v[c];
            switch (sub) {
              case SK_UOP_NEG: a = UnF32<SK_UOP_NEG>::apply(a); break;
              case SK_UOP_EXP: a = UnF32<SK_UOP_EXP>::apply(a); break;
              case SK_UOP_LOG: a = UnF32<SK_UOP_LOG>::apply(a); break;
              case SK_UOP_SQRT: a = UnF32<SK_UOP_SQRT>::apply(a); break;
              case SK_UOP_RELU: a = UnF32<SK_UOP_RELU>::apply(a); break;
              default: a = UnF32<SK_UOP_ABS>::apply(a); break;
            }
            acc[j].v[c] = a;
          }
        continue;
      }
      FVec<W> o[U];
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const int64_t g = base + (int64_t)j * kThreads;
        if (g < groups) o[j] = fused_operand<W>(p, k, g * W, tmp[j]);
        else {
#pragma unroll
          for (int c = 0; c < W; ++c) o[j].v[c] = 1.f;
        }
      }
      if (code == SK_F_LOAD) {
#pragma unroll
        for (int j = 0; j < U; ++j) acc[j] = o[j];
        continue;
      }
      const bool rev = p.rev[k] != 0, spow = p.src[k] == SK_F_CONST;
      const float cst = p.cst[k];
#pragma unroll
      for (int j = 0; j < U; ++j) {
        switch (sub) {
          case SK_OP_ADD: fused_bin<SK_OP_ADD, W>(acc[j], o[j], rev, cst, spow); break;
          case SK_OP_SUB: fused_bin<SK_OP_SUB, W>(acc[j], o[j], rev, cst, spow); break;
          case SK_OP_MUL: fused_bin<SK_OP_MUL, W>(acc[j], o[j], rev, cst, spow); break;
          case SK_OP_DIV: fused_bin<SK_OP_DIV, W>(acc[j], o[j], rev, cst, spow); break;
          case SK_OP_POW: fused_bin<SK_OP_POW, W>(acc[j], o[j], rev, cst, spow); break;
          case SK_OP_MAXIMUM: fused_bin<SK_OP_MAXIMUM, W>(acc[j], o[j], rev, cst, spow); break;
          default: fused_bin<SK_OP_MINIMUM, W>(acc[j], o[j], rev, cst, spow); break;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int64_t g = base + (int64_t)j * kThreads;
      if (g < groups) {
        if (W == 4) st_stream(reinterpret_cast<float4 *>(out) + g, make_float4(acc[j].v[0], acc[j].v[1 % W], acc[j].v[2 % W], acc[j].v[3 % W]));
        else out[g] = acc[j].v[0];
      }
    }
  }
}

#define SK_BIN_SWITCH(op, CALL)                  \
  switch (op) {                                  \
    case SK_OP_ADD: CALL(SK_OP_ADD); break;      \
    case SK_OP_SUB: CALL(SK_OP_SUB); break;      \
    case SK_OP_MUL: CALL(SK_OP_MUL); break;      \
    case SK_OP_DIV: CALL(SK_OP_DIV); break;      \
    case SK_OP_POW: CALL(SK_OP_POW); break;      \
    case SK_OP_MAXIMUM: CALL(SK_OP_MAXIMUM); break; \
    case SK_OP_MINIMUM: CALL(SK_OP_MINIMUM); break; \
    default: break;                              \
  }

}  // namespace sk

using namespace sk;

extern "C" {

int sk_ewise_binary(int op, const sk_array *a, const sk_array *b, sk_array *out) {
  int rc;
  if ((rc = ensure_init())) return rc;
  if ((rc = check_array(a, "a")) || (rc = check_array(b, "b")) || (rc = check_array(out, "out")))
    return rc;
  SK_REQUIRE(same_shape(a, out) && same_shape(b, out),
             "sk_ewise_binary: operands must be pre-broadcast to the output shape");
  const bool is_cmp = op >= SK_OP_EQ && op <= SK_OP_LE;
  SK_REQUIRE(is_cmp || (op >= SK_OP_ADD && op <= SK_OP_MINIMUM), "sk_ewise_binary: bad op %d", op);
  const int64_t n = numel(out);
  if (n == 0) return SK_OK;

  const int64_t *strs[3] = {a->strides, b->strides, out->strides};
  Collapsed<3> c;
  collapse_dims<3>(out->ndim, out->shape, strs, c);

  const bool all_f32 = a->dtype == SK_F32 && b->dtype == SK_F32 && out->dtype == SK_F32 && !is_cmp;
  if (all_f32 && aligned16(a->data) && aligned16(b->data) && aligned16(out->data)) {
    if (c.ndim == 1 && c.strides[0][0] == 1 && c.strides[1][0] == 1 && c.strides[2][0] == 1) {
      int grid = grid_for((n + 3) / 4, kThreads * kUnroll, 8);
      ProfScope ps(SK_PROF_EWISE, (double)n * 12.0);
#define CALL(OP) binary_f32_vec<OP><<<grid, kThreads, 0, stream()>>>((const float *)a->data, (const float *)b->data, (float *)out->data, n)
      SK_BIN_SWITCH(op, CALL)
#undef CALL
      SK_LAUNCH_CHECK();
      return SK_OK;
    }
    // (R, C) with each operand full / row / column / scalar broadcast
    int nd = c.ndim;
    if (nd <= 2) {
      int64_t R = nd == 2 ? c.shape[0] : 1, C = c.shape[nd - 1];
      auto s0 = [&](int k) { return nd == 2 ? c.strides[k][0] : 0; };
      auto s1 = [&](int k) { return c.strides[k][nd - 1]; };
      bool out_ok = s1(2) == 1 && (nd == 1 || s0(2) == C);
      auto opnd_ok = [&](int k) {
        int64_t i1 = s1(k), i0 = s0(k);
        if (i1 != 0 && i1 != 1) return false;
        if (i1 == 1 && i0 != 0 && (i0 % 4 != 0)) return false;
        return i0 >= 0;
      };
      if (out_ok && (C % 4 == 0) && opnd_ok(0) && opnd_ok(1)) {
        int grid = grid_for(R * (C / 4), kThreads, 8);
        ProfScope ps(SK_PROF_EWISE, (double)n * 4.0 + ((s0(0) && s1(0)) ? (double)n * 4.0 : 0.0) + ((s0(1) && s1(1)) ? (double)n * 4.0 : 0.0));
#define CALL(OP) binary_f32_bcast2d<OP><<<grid, kThreads, 0, stream()>>>((const float *)a->data, s0(0), (int)s1(0), (const float *)b->data, s0(1), (int)s1(1), (float *)out->data, R, C / 4)
        SK_BIN_SWITCH(op, CALL)
#undef CALL
        SK_LAUNCH_CHECK();
        return SK_OK;
      }
    }
  }

  GenDesc d;
  memset(&d, 0, sizeof(d));
  d.a = a->data; d.b = b->data; d.out = out->data;
  d.a_dt = a->dtype; d.b_dt = b->dtype; d.out_dt = out->dtype;
  d.op = op; d.mode = 0;
  fill_desc<3>(d, c, 0, 1, 2);
  int cls;
  if (is_cmp) cls = (dtype_is_float(a->dtype) || dtype_is_float(b->dtype)) ? 1 : 2;
  else cls = compute_class_for(out->dtype);
  return launch_generic(d, cls);
}

int sk_ewise_scalar(int op, const sk_array *a, double fscalar, int64_t iscalar, int scalar_is_int,
                    int reverse, sk_array *out) {
  int rc;
  if ((rc = ensure_init())) return rc;
  if ((rc = check_array(a, "a")) || (rc = check_array(out, "out"))) return rc;
  SK_REQUIRE(same_shape(a, out), "sk_ewise_scalar: shape mismatch");
  const bool is_cmp = op >= SK_OP_EQ && op <= SK_OP_LE;
  SK_REQUIRE(is_cmp || (op >= SK_OP_ADD && op <= SK_OP_MINIMUM), "sk_ewise_scalar: bad op %d", op);
  const int64_t n = numel(out);
  if (n == 0) return SK_OK;
  if (scalar_is_int) fscalar = (double)iscalar;
  else iscalar = (int64_t)fscalar;

  const int64_t *strs[2] = {a->strides, out->strides};
  Collapsed<2> c;
  collapse_dims<2>(out->ndim, out->shape, strs, c);

  if (a->dtype == SK_F32 && out->dtype == SK_F32 && !is_cmp && c.ndim == 1 &&
      c.strides[0][0] == 1 && c.strides[1][0] == 1 && aligned16(a->data) && aligned16(out->data)) {
    int grid = grid_for((n + 3) / 4, kThreads * kUnroll, 8);
    ProfScope ps(SK_PROF_EWISE, (double)n * 8.0);
    float s = (float)fscalar;
    int pk = pow_kind(fscalar);
    if (op == SK_OP_MAXIMUM && s == 0.f) {
      unary_f32_vec<SK_UOP_RELU><<<grid, kThreads, 0, stream()>>>((const float *)a->data, (float *)out->data, n);
    } else if (reverse) {
#define CALL(OP) scalar_f32_vec<OP, true><<<grid, kThreads, 0, stream()>>>((const float *)a->data, s, pk, (float *)out->data, n)
      SK_BIN_SWITCH(op, CALL)
#undef CALL
    } else {
#define CALL(OP) scalar_f32_vec<OP, false><<<grid, kThreads, 0, stream()>>>((const float *)a->data, s, pk, (float *)out->data, n)
      SK_BIN_SWITCH(op, CALL)
#undef CALL
    }
    SK_LAUNCH_CHECK();
    return SK_OK;
  }

  GenDesc d;
  memset(&d, 0, sizeof(d));
  d.a = a->data; d.out = out->data;
  d.a_dt = a->dtype; d.out_dt = out->dtype;
  d.op = op; d.mode = reverse ? 2 : 1;
  d.fscalar = fscalar; d.iscalar = iscalar;
  fill_desc<2>(d, c, 0, -1, 1);
  int cls;
  if (is_cmp) cls = (dtype_is_float(a->dtype) || !scalar_is_int) ? 1 : 2;
  else cls = compute_class_for(out->dtype);
  // weak float scalar against fp32 data compares in fp32 (NEP 50): round it first
  if (is_cmp && a->dtype == SK_F32) d.fscalar = (double)(float)fscalar;
  return launch_generic(d, cls);
}

int sk_ewise_unary(int op, const sk_array *a, sk_array *out) {
  int rc;
  if ((rc = ensure_init())) return rc;
  if ((rc = check_array(a, "a")) || (rc = check_array(out, "out"))) return rc;
  SK_REQUIRE(same_shape(a, out), "sk_ewise_unary: shape mismatch");
  SK_REQUIRE(op >= SK_UOP_NEG && op <= SK_UOP_ABS, "sk_ewise_unary: bad op %d", op);
  const int64_t n = numel(out);
  if (n == 0) return SK_OK;
  const int64_t *strs[2] = {a->strides, out->strides};
  Collapsed<2> c;
  collapse_dims<2>(out->ndim, out->shape, strs, c);
  if (a->dtype == SK_F32 && out->dtype == SK_F32 && c.ndim == 1 && c.strides[0][0] == 1 &&
      c.strides[1][0] == 1 && aligned16(a->data) && aligned16(out->data)) {
    int grid = grid_for((n + 3) / 4, kThreads * kUnroll, 8);
    ProfScope ps(SK_PROF_EWISE, (double)n * 8.0);
    const float *ap = (const float *)a->data;
    float *op_ = (float *)out->data;
    switch (op) {
      case SK_UOP_NEG: unary_f32_vec<SK_UOP_NEG><<<grid, kThreads, 0, stream()>>>(ap, op_, n); break;
      case SK_UOP_EXP: unary_f32_vec<SK_UOP_EXP><<<grid, kThreads, 0, stream()>>>(ap, op_, n); break;
      case SK_UOP_LOG: unary_f32_vec<SK_UOP_LOG><<<grid, kThreads, 0, stream()>>>(ap, op_, n); break;
      case SK_UOP_SQRT: unary_f32_vec<SK_UOP_SQRT><<<grid, kThreads, 0, stream()>>>(ap, op_, n); break;
      case SK_UOP_RELU: unary_f32_vec<SK_UOP_RELU><<<grid, kThreads, 0, stream()>>>(ap, op_, n); break;
      default: unary_f32_vec<SK_UOP_ABS><<<grid, kThreads, 0, stream()>>>(ap, op_, n); break;
    }
    SK_LAUNCH_CHECK();
    return SK_OK;
  }
  GenDesc d;
  memset(&d, 0, sizeof(d));
  d.a = a->data; d.out = out->data;
  d.a_dt = a->dtype; d.out_dt = out->dtype;
  d.op = op; d.mode = 3;
  fill_desc<2>(d, c, 0, -1, 1);
  return launch_generic(d, compute_class_for(out->dtype));
}

int sk_copy(const sk_array *src, sk_array *dst) {
  int rc;
  if ((rc = ensure_init())) return rc;
  if ((rc = check_array(src, "src")) || (rc = check_array(dst, "dst"))) return rc;
  SK_REQUIRE(same_shape(src, dst), "sk_copy: src must be pre-broadcast to dst's shape");
  const int64_t n = numel(dst);
  if (n == 0) return SK_OK;
  const int64_t *strs[2] = {src->strides, dst->strides};
  Collapsed<2> c;
  collapse_dims<2>(dst->ndim, dst->shape, strs, c);

  if (src->dtype == dst->dtype) {
    const int esz = dtype_size(src->dtype);
    // fully contiguous on both sides: plain D2D copy
    if (c.ndim == 1 && c.strides[0][0] == 1 && c.strides[1][0] == 1) {
      SK_CUDA(cudaMemcpyAsync(dst->data, src->data, (size_t)n * esz, cudaMemcpyDeviceToDevice, stream()));
      note_launch();
      return SK_OK;
    }
    // contiguous destination rows, source inner stride 0 or 1 (broadcast views, pitched slices)
    if (esz == 4 && c.ndim <= 2 && c.strides[1][c.ndim - 1] == 1 &&
        (c.strides[0][c.ndim - 1] == 0 || c.strides[0][c.ndim - 1] == 1) && c.shape[c.ndim - 1] % 4 == 0 &&
        aligned16(dst->data) && (c.ndim == 1 || c.strides[1][0] % 4 == 0)) {
      const int64_t R = c.ndim == 2 ? c.shape[0] : 1, C = c.shape[c.ndim - 1];
      const int64_t lds = c.ndim == 2 ? c.strides[0][0] : 0, ldd = c.ndim == 2 ? c.strides[1][0] : C;
      const int sin = (int)c.strides[0][c.ndim - 1];
      const int vec_src = sin == 1 && aligned16(src->data) && lds % 4 == 0;
      int tpr_log2 = 0;
      while (tpr_log2 < 8 && (1 << tpr_log2) < C / 4) ++tpr_log2;
      const int rpb = 256 >> tpr_log2;
      const int grid = grid_for(R, rpb, 16);
      ProfScope ps(SK_PROF_COPY, (double)n * 4.0 + (double)(sin ? (lds ? n : C) : R) * 4.0);
      if (sin == 0) copy_rows_u32<0><<<grid, 256, 0, stream()>>>((const uint32_t *)src->data, (uint32_t *)dst->data, R, C, lds, ldd, tpr_log2, 0);
      else copy_rows_u32<1><<<grid, 256, 0, stream()>>>((const uint32_t *)src->data, (uint32_t *)dst->data, R, C, lds, ldd, tpr_log2, vec_src);
      SK_LAUNCH_CHECK();
      return SK_OK;
    }
    // 2-D transpose pattern -> shared-memory tiled transpose (coalesced both sides)
    if (c.ndim == 2 && c.strides[0][0] == 1 && c.strides[1][1] == 1 && c.strides[1][0] >= c.shape[1] &&
        c.strides[0][1] >= c.shape[0] && (esz == 4 || esz == 8 || esz == 2 || esz == 1)) {
      int64_t R = c.shape[0], C = c.shape[1];
      if (esz == 4 && R % 4 == 0 && C % 4 == 0 && c.strides[0][1] % 4 == 0 && c.strides[1][0] % 4 == 0 &&
          aligned16(src->data) && aligned16(dst->data)) {
        const int64_t tiles64 = ((R + 63) / 64) * ((C + 63) / 64);
        static int occ64 = 0;
        if (!occ64) {
          if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ64, transpose64_u32, 256, 0) != cudaSuccess || occ64 < 1) occ64 = 4;
        }
        const int64_t cap64 = (int64_t)ctx().num_sms * occ64;   // one resident wave: the tile loop is software-pipelined
        const int grid64 = (int)(tiles64 < cap64 ? tiles64 : cap64);
        ProfScope ps(SK_PROF_COPY, (double)n * esz * 2.0);
        transpose64_u32<<<grid64, 256, 0, stream()>>>((const uint32_t *)src->data, (uint32_t *)dst->data, R, C,
                                                      c.strides[0][1], c.strides[1][0]);
        SK_LAUNCH_CHECK();
        return SK_OK;
      }
      int64_t tiles = ((R + 31) / 32) * ((C + 31) / 32);
      int grid = (int)(tiles < (int64_t)ctx().num_sms * 16 ? tiles : (int64_t)ctx().num_sms * 16);
      ProfScope ps(SK_PROF_COPY, (double)n * esz * 2.0);
      if (esz == 4) transpose_tiled<uint32_t><<<grid, 256, 0, stream()>>>((const uint32_t *)src->data, (uint32_t *)dst->data, R, C, c.strides[0][1], c.strides[1][0]);
      else if (esz == 8) transpose_tiled<uint64_t><<<grid, 256, 0, stream()>>>((const uint64_t *)src->data, (uint64_t *)dst->data, R, C, c.strides[0][1], c.strides[1][0]);
      else if (esz == 2) transpose_tiled<uint16_t><<<grid, 256, 0, stream()>>>((const uint16_t *)src->data, (uint16_t *)dst->data, R, C, c.strides[0][1], c.strides[1][0]);
      else transpose_tiled<uint8_t><<<grid, 256, 0, stream()>>>((const uint8_t *)src->data, (uint8_t *)dst->data, R, C, c.strides[0][1], c.strides[1][0]);
      SK_LAUNCH_CHECK();
      return SK_OK;
    }
    GenDesc d;
    memset(&d, 0, sizeof(d));
    d.a = src->data; d.out = dst->data;
    d.a_dt = src->dtype; d.out_dt = dst->dtype;
    d.mode = 4;
    fill_desc<2>(d, c, 0, -1, 1);
    int grid = grid_for(d.n, kThreads, 16);
    ProfScope ps(SK_PROF_COPY, (double)d.n * esz * 2.0);
    if (esz == 4) generic_copy_raw<uint32_t><<<grid, kThreads, 0, stream()>>>(d);
    else if (esz == 8) generic_copy_raw<uint64_t><<<grid, kThreads, 0, stream()>>>(d);
    else if (esz == 2) generic_copy_raw<uint16_t><<<grid, kThreads, 0, stream()>>>(d);
    else generic_copy_raw<uint8_t><<<grid, kThreads, 0, stream()>>>(d);
    SK_LAUNCH_CHECK();
    return SK_OK;
  }

  GenDesc d;
  memset(&d, 0, sizeof(d));
  d.a = src->data; d.out = dst->data;
  d.a_dt = src->dtype; d.out_dt = dst->dtype;
  d.mode = 4;
  fill_desc<2>(d, c, 0, -1, 1);
  // cast through the widest class that holds both sides exactly
  int cls;
  if (dtype_is_float(src->dtype) || dtype_is_float(dst->dtype)) {
    bool wide = src->dtype == SK_F64 || dst->dtype == SK_F64 || src->dtype == SK_I64 ||
                src->dtype == SK_U64 || src->dtype == SK_I32 || src->dtype == SK_U32 ||
                dst->dtype == SK_I64 || dst->dtype == SK_U64 || dst->dtype == SK_I32 ||
                dst->dtype == SK_U32;
    cls = wide ? 1 : 0;
  } else {
    cls = 2;
  }
  return launch_generic(d, cls);
}

int sk_fill(sk_array *dst, double fvalue, int64_t ivalue, int value_is_int) {
  int rc;
  if ((rc = ensure_init())) return rc;
  if ((rc = check_array(dst, "dst"))) return rc;
  const int64_t n = numel(dst);
  if (n == 0) return SK_OK;
  if (value_is_int) fvalue = (double)ivalue;
  else ivalue = (int64_t)fvalue;
  if (is_contiguous(dst)) {
    const int esz = dtype_size(dst->dtype);
    bool zero = value_is_int ? (ivalue == 0) : (fvalue == 0.0 && !signbit(fvalue));
    if (zero) {
      SK_CUDA(cudaMemsetAsync(dst->data, 0, (size_t)n * esz, stream()));
      note_launch();
      return SK_OK;
    }
    if (dst->dtype == SK_F32 && aligned16(dst->data)) {
      int grid = grid_for((n + 3) / 4, kThreads, 8);
      fill_f32_vec<<<grid, kThreads, 0, stream()>>>((float *)dst->data, (float)fvalue, n);
      SK_LAUNCH_CHECK();
      return SK_OK;
    }
  }
  const int64_t *strs[1] = {dst->strides};
  Collapsed<1> c;
  collapse_dims<1>(dst->ndim, dst->shape, strs, c);
  GenDesc d;
  memset(&d, 0, sizeof(d));
  d.out = dst->data;
  d.out_dt = dst->dtype;
  d.mode = 5;
  d.fscalar = fvalue; d.iscalar = ivalue;
  fill_desc<1>(d, c, -1, -1, 0);
  int cls = dtype_is_float(dst->dtype) ? 1 : 2;
  return launch_generic(d, cls);
}

int sk_ewise_fused(const sk_fused_program *prog, float *out) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(prog && out, "sk_ewise_fused: null pointer");
  const sk_fused_program &p = *prog;
  SK_REQUIRE(p.n_ops >= 1 && p.n_ops <= SK_FUSED_MAX_OPS && p.n_in >= 0 && p.n_in <= SK_FUSED_MAX_INPUTS,
             "sk_ewise_fused: program of %d ops / %d inputs out of range", p.n_ops, p.n_in);
  SK_REQUIRE(p.n >= 0 && p.cols >= 1, "sk_ewise_fused: bad sizes");
  SK_REQUIRE(p.code[0] == SK_F_LOAD, "sk_ewise_fused: a program starts with a load");
  bool vec = p.n % 4 == 0 && aligned16(out);
  for (int k = 0; k < p.n_ops; ++k) {
    SK_REQUIRE(p.code[k] <= SK_F_UN, "sk_ewise_fused: bad code at op %d", k);
    if (p.code[k] == SK_F_STORE) { SK_REQUIRE(p.idx[k] < 3, "sk_ewise_fused: 3 temporaries"); continue; }
    if (p.code[k] == SK_F_UN) { SK_REQUIRE(p.sub[k] <= SK_UOP_ABS, "sk_ewise_fused: bad unary op"); continue; }
    if (p.code[k] == SK_F_BIN) SK_REQUIRE(p.sub[k] <= SK_OP_MINIMUM, "sk_ewise_fused: bad binary op");
    SK_REQUIRE(p.src[k] >= SK_F_IN && p.src[k] <= SK_F_TEMP, "sk_ewise_fused: bad operand source at op %d", k);
    if (p.src[k] == SK_F_TEMP) SK_REQUIRE(p.idx[k] < 3, "sk_ewise_fused: 3 temporaries");
    if (p.src[k] == SK_F_IN) SK_REQUIRE(p.idx[k] < p.n_in, "sk_ewise_fused: input index out of range at op %d", k);
  }
  for (int i = 0; i < p.n_in; ++i) {
    SK_REQUIRE(p.in[i] != nullptr && p.in_kind[i] >= SK_F_FULL && p.in_kind[i] <= SK_F_SINGLE, "sk_ewise_fused: bad input %d", i);
    if (p.in_kind[i] != SK_F_SINGLE) vec = vec && aligned16(p.in[i]);
    if (p.in_kind[i] == SK_F_VECTOR) vec = vec && p.cols % 4 == 0;
  }
  if (p.n == 0) return SK_OK;
  double bytes = 4.0 * (double)p.n;
  for (int i = 0; i < p.n_in; ++i) bytes += p.in_kind[i] == SK_F_FULL ? 4.0 * (double)p.n : 0.0;
  ProfScope ps(SK_PROF_EWISE, bytes);
  if (vec) {
    int grid = grid_for(p.n / 4, kThreads * 2, 8);
    fused_ewise_kernel<4, 2><<<grid, kThreads, 0, stream()>>>(p, out);
  } else {
    int grid = grid_for(p.n, kThreads * 2, 8);
    fused_ewise_kernel<1, 2><<<grid, kThreads, 0, stream()>>>(p, out);
  }
  SK_LAUNCH_CHECK();
  return SK_OK;
}

int sk_relu_bwd(const sk_array *x, const sk_array *adj, sk_array *out) {
  int rc;
  if ((rc = ensure_init())) return rc;
  if ((rc = check_array(x, "x")) || (rc = check_array(adj, "adj")) || (rc = check_array(out, "out")))
    return rc;
  SK_REQUIRE(x->dtype == SK_F32 && adj->dtype == SK_F32 && out->dtype == SK_F32,
             "sk_relu_bwd: fp32 only");
  SK_REQUIRE(same_shape(x, out) && same_shape(adj, out), "sk_relu_bwd: shape mismatch");
  SK_REQUIRE(is_contiguous(x) && is_contiguous(adj) && is_contiguous(out),
             "sk_relu_bwd: contiguous operands required");
  SK_REQUIRE(aligned16(x->data) && aligned16(adj->data) && aligned16(out->data),
             "sk_relu_bwd: 16-byte aligned operands required");
  const int64_t n = numel(out);
  if (n == 0) return SK_OK;
  int grid = grid_for((n + 3) / 4, kThreads * kUnroll, 8);
  ProfScope ps(SK_PROF_EWISE, (double)n * 12.0);
  relu_bwd_f32_vec<<<grid, kThreads, 0, stream()>>>((const float *)x->data, (const float *)adj->data, (float *)out->data, n);
  SK_LAUNCH_CHECK();
  return SK_OK;
}

int sk_cast_bf16(const sk_array *src, sk_array *dst) {
  SK_REQUIRE(src && dst && src->dtype == SK_F32 && dst->dtype == SK_BF16,
             "sk_cast_bf16: expects float32 -> bfloat16");
  return sk_copy(src, dst);  // __float2bfloat16_rn: round-to-nearest-even
}

}  // extern "C"
