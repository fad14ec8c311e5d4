// common.cuh -- shared plumbing for libsoketb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/soket_b200.h"

namespace sk {

// ---- error plumbing -------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define SK_CUDA(expr)                                                          \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) return sk::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define SK_REQUIRE(cond, ...)                                                  \
  do {                                                                         \
    if (!(cond)) {                                                             \
      sk::set_error(__VA_ARGS__);                                              \
      return SK_ERR_ARG;                                                       \
    }                                                                          \
  } while (0)

// Checks the launch itself (async execution errors surface at sk_sync()).
#define SK_LAUNCH_CHECK()                                                      \
  do {                                                                         \
    sk::note_launch();                                                         \
    cudaError_t _e = cudaGetLastError();                                       \
    if (_e != cudaSuccess) return sk::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
  } while (0)

// ---- context ----------------------------------------------------------------
struct Context {
  bool ready = false;
  int device = -1;
  int num_sms = 148;
  size_t l2_bytes = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t comm_stream = nullptr;
  cudaStream_t copy_stream = nullptr;   // host -> device input prefetch (sk_h2d_prefetch)
  cudaStream_t opt_stream = nullptr;    // per-bucket optimizer updates of data-parallel training
  cudaStream_t launch = nullptr;        // where kernels go: `stream` unless sk_launch_stream() says otherwise
};
Context &ctx();
int ensure_init();
void note_launch();
inline cudaStream_t stream() { return ctx().launch; }
cudaStream_t stream_by_id(int id);   // nullptr for an unknown id
// device-resident counter mixed into dropout seeds at RUN time; a captured graph advances it per replay
const uint64_t *rng_epoch_ptr();
void dropout_reseed(uint64_t seed);
// Sticky device-side error word in mapped pinned host memory: a kernel that meets an invalid index
// ORs a code into it (the reference raises IndexError at the call; kernels are asynchronous, so it
// is raised at the next sync point -- sk_sync / sk_d2h / sk_event_sync -- instead).
enum { SK_DEVERR_LABEL_RANGE = 1 };
unsigned int *dev_error_ptr();        // device address, for kernels
int check_dev_error();                // host: SK_OK, or SK_ERR_INDEX after clearing the word   // nn_fused.cu: sk_rng_seed also restarts the dropout draw sequence

// ---- per-family profiling (sk_prof_*) -------------------------------------------
// Usage inside a launcher:  ProfScope ps(SK_PROF_GEMM_TC, flops);  ... launch ...
bool prof_on();
void prof_begin(int family);
void prof_end(int family, double work);
struct ProfScope {
  int family;
  double work;
  bool on;
  ProfScope(int f, double w) : family(f), work(w), on(prof_on()) {
    if (on) prof_begin(family);
  }
  ~ProfScope() {
    if (on) prof_end(family, work);
  }
};

// ---- dtype helpers ----------------------------------------------------------
__host__ __device__ inline int dtype_size(int dt) {
  switch (dt) {
    case SK_BOOL: case SK_I8: case SK_U8: return 1;
    case SK_I16: case SK_U16: case SK_F16: case SK_BF16: return 2;
    case SK_I32: case SK_U32: case SK_F32: return 4;
    default: return 8;
  }
}
__host__ __device__ inline bool dtype_is_float(int dt) {
  return dt == SK_F16 || dt == SK_F32 || dt == SK_F64 || dt == SK_BF16;
}

inline int64_t numel(const sk_array *a) {
  int64_t n = 1;
  for (int i = 0; i < a->ndim; ++i) n *= a->shape[i];
  return n;
}
inline bool is_contiguous(const sk_array *a) {
  int64_t expect = 1;
  for (int i = a->ndim - 1; i >= 0; --i) {
    if (a->shape[i] == 1) continue;
    if (a->strides[i] != expect) return false;
    expect *= a->shape[i];
  }
  return true;
}
inline bool same_shape(const sk_array *a, const sk_array *b) {
  if (a->ndim != b->ndim) return false;
  for (int i = 0; i < a->ndim; ++i)
    if (a->shape[i] != b->shape[i]) return false;
  return true;
}

// Collapsed iteration space shared by N operands: merges adjacent dims that are
// jointly contiguous for every operand and drops size-1 dims.
template <int N>
struct Collapsed {
  int ndim;
  int64_t shape[SK_MAX_NDIM];
  int64_t strides[N][SK_MAX_NDIM];
};

template <int N>
inline void collapse_dims(int ndim, const int64_t *shape, const int64_t *const strides[N],
                          Collapsed<N> &out) {
  int64_t shp[SK_MAX_NDIM];
  int64_t str[N][SK_MAX_NDIM];
  int nd = 0;
  for (int i = 0; i < ndim; ++i) {
    if (shape[i] == 1) continue;
    shp[nd] = shape[i];
    for (int k = 0; k < N; ++k) str[k][nd] = strides[k][i];
    ++nd;
  }
  // merge from the innermost outwards: dims (i, i+1) merge if str[i] == str[i+1]*shape[i+1]
  int w = 0;
  for (int i = 0; i < nd; ++i) {
    if (w > 0) {
      bool ok = true;
      for (int k = 0; k < N; ++k)
        if (str[k][w - 1] != str[k][i] * shp[i]) { ok = false; break; }
      if (ok) {
        shp[w - 1] *= shp[i];
        for (int k = 0; k < N; ++k) str[k][w - 1] = str[k][i];
        continue;
      }
    }
    shp[w] = shp[i];
    for (int k = 0; k < N; ++k) str[k][w] = str[k][i];
    ++w;
  }
  if (w == 0) {
    w = 1;
    shp[0] = 1;
    for (int k = 0; k < N; ++k) str[k][0] = 1;
  }
  out.ndim = w;
  for (int i = 0; i < w; ++i) {
    out.shape[i] = shp[i];
    for (int k = 0; k < N; ++k) out.strides[k][i] = str[k][i];
  }
}

// Grid sizing: multiples of the SM count (148 on B200), capped by the work.
inline int grid_for(int64_t work_items, int per_block, int blocks_per_sm = 8) {
  int64_t need = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)ctx().num_sms * blocks_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ---- device helpers -----------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// streaming 128-bit accesses: data touched exactly once should not pollute L1
__device__ __forceinline__ float4 ld_stream(const float4 *p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4 *p, const float4 &v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

}  // namespace sk
