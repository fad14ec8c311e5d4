// optim.cu -- multi-tensor SGD and Adam: one launch updates every parameter in place.
//
// Replaces the per-parameter loops of unfused array calls in
// soket/optim.pyx:82-131 (SGD.step, 5 calls/param) and :201-269 (Adam.step,
// 14 calls/param).  Every multiply/add/divide/sqrt is issued as a separately
// rounded IEEE operation (__fmul_rn / __fadd_rn / __fdiv_rn / __fsqrt_rn, never
// contracted into FMA) in the reference's order, so given identical gradients
// the update is bit-identical to the NumPy path.  12 B/param (SGD), 28 B/param
// (Adam) of HBM traffic.
#include <stdlib.h>

#include "common.cuh"
#include "matmul_split.cuh"
#include "optim.cuh"

namespace sk {

constexpr int kOT = 256;
constexpr int kMaxTensors = 48;          // per launch (kernel-argument space)
constexpr int64_t kChunk = kOT * 4 * 4;  // elements per block-iteration

struct SgdArgs {
  float *p[kMaxTensors];
  const float *g[kMaxTensors];
  int64_t size[kMaxTensors];
  int block_start[kMaxTensors + 1];
  int n;
  float lr, wd, grad_scale;
  int have_wd, have_scale;
};

struct AdamArgs {
  float *p[kMaxTensors];
  const float *g[kMaxTensors];
  float *m[kMaxTensors];
  float *v[kMaxTensors];
  int64_t size[kMaxTensors];
  int block_start[kMaxTensors + 1];
  int n;
  float lr, beta1, beta2, omb1, omb2, eps, wd, bc1, bc2, grad_scale;
  int have_wd, have_scale, first;
  // capturable variant: {beta1^t, beta2^t} as doubles in DEVICE memory (advanced by
  // adam_bias_advance_kernel), so that a CUDA-graph replay sees the current bias corrections
  const double *bias_state;
  // optional by-product: the bit pattern of max |p_new| per tensor (atomicMax into a zeroed word), from which
  // the weight's fp16x3 operand split takes its scale without a pass of its own (sk_split_f16 amax_bits)
  unsigned int *amax[kMaxTensors];
  // sk_adam_step_split: per tensor two persistent words {max |p| before this update, accumulator (zero)}
  // (`amax` above then points at the accumulator).  With hi / lo the NEW weights also leave the kernel as
  // the fp16 hi / lo operand split of the fp16x3 GEMM, scaled by a power of two chosen BEFORE the update
  // from max |p_old| + update_bound (|p_new - p_old| <= update_bound, see sk_adam_step_split): the
  // weight's split costs 4 B/element of writes here instead of a 8 B/element pass of its own.  The last
  // block to finish rotates the words (cur = accumulator, accumulator = 0) for the next step.
  unsigned int *amax_cur[kMaxTensors];
  __half *hi[kMaxTensors];
  __half *lo[kMaxTensors];
  float *scale4[kMaxTensors];
  float update_bound;
  unsigned int *done;  // launch-wide counter of finished blocks (zero before and after the launch), or null
  int total_blocks;   // chunks over all tensors; the grid may be smaller (capped) and strides over them
};

template <typename A>
__device__ __forceinline__ int find_tensor(const A &a, int b) {
  int lo = 0, hi = a.n;  // block_start[lo] <= b < block_start[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (a.block_start[mid] <= b) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ float sgd_one(float p, float g, const SgdArgs &a) {
  if (a.have_scale) g = __fmul_rn(g, a.grad_scale);
  if (a.have_wd) g = __fadd_rn(g, __fmul_rn(p, a.wd));  // optim.pyx:105
  // quirk Q2 (optim.pyx:72,108-125): the momentum branch only runs when momentum == 0,
  // where u = 0*u + 1*g = g; every configuration reduces to plain SGD.
  return __fsub_rn(p, __fmul_rn(a.lr, g));  // optim.pyx:131
}

__global__ void __launch_bounds__(kOT) sgd_kernel(const __grid_constant__ SgdArgs a) {
  const int t = find_tensor(a, blockIdx.x);
  const int64_t base = (int64_t)(blockIdx.x - a.block_start[t]) * kChunk;
  float *p = a.p[t];
  const float *g = a.g[t];
  const int64_t n = a.size[t];
  const bool vec = ((((uintptr_t)p) | ((uintptr_t)g)) & 15) == 0;
  if (vec) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t i = base + ((int64_t)j * kOT + threadIdx.x) * 4;
      if (i + 3 < n) {
        float4 pv = *reinterpret_cast<float4 *>(p + i);
        float4 gv = ld_stream(reinterpret_cast<const float4 *>(g + i));
        pv.x = sgd_one(pv.x, gv.x, a); pv.y = sgd_one(pv.y, gv.y, a);
        pv.z = sgd_one(pv.z, gv.z, a); pv.w = sgd_one(pv.w, gv.w, a);
        *reinterpret_cast<float4 *>(p + i) = pv;
      } else {
        for (int64_t k = i; k < n && k < i + 4; ++k) p[k] = sgd_one(p[k], g[k], a);
      }
    }
  } else {
    for (int64_t i = base + threadIdx.x; i < n && i < base + kChunk; i += kOT) p[i] = sgd_one(p[i], g[i], a);
  }
}

__global__ void __launch_bounds__(kOT) adam_kernel(const __grid_constant__ AdamArgs a) {
 for (int blk = blockIdx.x; blk < a.total_blocks; blk += gridDim.x) {
  const int t = find_tensor(a, blk);
  const int64_t base = (int64_t)(blk - a.block_start[t]) * kChunk;
  float amax = 0.f;
  float *p = a.p[t];
  const float *g = a.g[t];
  float *m = a.m[t];
  float *v = a.v[t];
  const int64_t n = a.size[t];
  // 1 - beta^t: the host's double subtraction rounded to float32 (optim.pyx:266-269 + NEP 50),
  // from the kernel arguments or -- capturable -- recomputed identically from device state
  const float bc1 = a.bias_state ? (float)__dsub_rn(1.0, a.bias_state[0]) : a.bc1;
  const float bc2 = a.bias_state ? (float)__dsub_rn(1.0, a.bias_state[1]) : a.bc2;
  const bool vec = ((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v)) & 15) == 0;
  __half *hi = a.hi[t], *lo = a.lo[t];
  float sc = 1.f;
  if (hi) {   // the host only passes hi / lo for 16-byte aligned tensors of a multiple of 4 elements
    float inv;
    const float bound = __fadd_ru(__uint_as_float(*a.amax_cur[t]), a.update_bound);
    pow2_scale(bound, sc, inv);
    if (blk == a.block_start[t] && threadIdx.x == 0) {
      float *s4 = a.scale4[t];
      s4[0] = sc; s4[1] = inv; s4[2] = bound; s4[3] = 0.f;
    }
  }
  if (vec) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t i = base + ((int64_t)j * kOT + threadIdx.x) * 4;
      if (i + 3 < n) {
        float4 pv = *reinterpret_cast<float4 *>(p + i);
        float4 gv = ld_stream(reinterpret_cast<const float4 *>(g + i));
        float4 mv = a.first ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<float4 *>(m + i);
        float4 vv = a.first ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<float4 *>(v + i);
        adam_one(pv.x, gv.x, mv.x, vv.x, a, bc1, bc2); adam_one(pv.y, gv.y, mv.y, vv.y, a, bc1, bc2);
        adam_one(pv.z, gv.z, mv.z, vv.z, a, bc1, bc2); adam_one(pv.w, gv.w, mv.w, vv.w, a, bc1, bc2);
        amax = fmaxf(fmaxf(amax, fmaxf(fabsf(pv.x), fabsf(pv.y))), fmaxf(fabsf(pv.z), fabsf(pv.w)));
        *reinterpret_cast<float4 *>(p + i) = pv;
        *reinterpret_cast<float4 *>(m + i) = mv;
        *reinterpret_cast<float4 *>(v + i) = vv;
        if (hi) {
          uint2 h, l;
          split4(pv, sc, h, l);
          *reinterpret_cast<uint2 *>(hi + i) = h;     // read next by the forward GEMM's TMA loads
          *reinterpret_cast<uint2 *>(lo + i) = l;
        }
      } else {
        for (int64_t k = i; k < n && k < i + 4; ++k) { adam_one(p[k], g[k], m[k], v[k], a, bc1, bc2); amax = fmaxf(amax, fabsf(p[k])); }
      }
    }
  } else {
    for (int64_t i = base + threadIdx.x; i < n && i < base + kChunk; i += kOT) { adam_one(p[i], g[i], m[i], v[i], a, bc1, bc2); amax = fmaxf(amax, fabsf(p[i])); }
  }
  if (a.amax[t]) {
    amax = warp_max(amax);
    // same-address atomics serialise in L2 (32 k warps per 4096 x 4096 weight): look first, and only the few
    // warps that would raise the running maximum issue one
    if ((threadIdx.x & 31) == 0 && amax > 0.f) {
      const unsigned int bits = __float_as_uint(amax);
      if (bits > *(volatile unsigned int *)a.amax[t]) atomicMax(a.amax[t], bits);
    }
  }
 }
  if (a.done) {   // the last block to finish: cur = accumulator, accumulator = 0 (nobody reads `cur` any more)
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      last = atomicAdd(a.done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
      __threadfence();
      for (int t = threadIdx.x; t < a.n; t += kOT)
        if (a.amax_cur[t]) {
          *a.amax_cur[t] = *(volatile unsigned int *)a.amax[t];
          *a.amax[t] = 0u;
        }
      if (threadIdx.x == 0) *a.done = 0u;
    }
  }
}

// optim.pyx:266-267: beta1_t *= beta1; beta2_t *= beta2 (Python floats = IEEE doubles)
__global__ void adam_bias_advance_kernel(double *state, double beta1, double beta2) {
  state[0] = __dmul_rn(state[0], beta1);
  state[1] = __dmul_rn(state[1], beta2);
}

}  // namespace sk

using namespace sk;

extern "C" {

int sk_sgd_step(int n_tensors, float *const *params, const float *const *grads, const int64_t *sizes,
                double lr, double weight_decay, double grad_scale) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(n_tensors >= 0 && (n_tensors == 0 || (params && grads && sizes)), "sk_sgd_step: null list");
  int i = 0;
  while (i < n_tensors) {
    SgdArgs a;
    memset(&a, 0, sizeof(a));
    int n = 0, blocks = 0;
    for (; i < n_tensors && n < kMaxTensors; ++i) {
      SK_REQUIRE(params[i] && grads[i] && sizes[i] >= 0, "sk_sgd_step: tensor %d has a null pointer", i);
      if (sizes[i] == 0) continue;
      a.p[n] = params[i]; a.g[n] = grads[i]; a.size[n] = sizes[i];
      a.block_start[n] = blocks;
      blocks += (int)((sizes[i] + kChunk - 1) / kChunk);
      ++n;
    }
    a.block_start[n] = blocks;
    a.n = n;
    a.lr = (float)lr; a.wd = (float)weight_decay; a.grad_scale = (float)grad_scale;
    a.have_wd = weight_decay != 0.0; a.have_scale = grad_scale != 1.0;
    if (blocks == 0) continue;
    double elems = 0;
    for (int k = 0; k < n; ++k) elems += (double)a.size[k];
    ProfScope ps(SK_PROF_OPTIM, elems * 12.0);
    sgd_kernel<<<blocks, kOT, 0, stream()>>>(a);
    SK_LAUNCH_CHECK();
  }
  return SK_OK;
}

static int adam_step_impl(int n_tensors, float *const *params, const float *const *grads, float *const *m,
                          float *const *v, const int64_t *sizes, double lr, double beta1, double beta2,
                          double eps, double weight_decay, double one_minus_beta1_t,
                          double one_minus_beta2_t, int first_step, double grad_scale,
                          const double *bias_state, unsigned int *const *amax,
                          const sk_adam_split *splits = nullptr, double update_bound = 0.0) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(n_tensors >= 0 && (n_tensors == 0 || (params && grads && m && v && sizes)), "sk_adam_step: null list");
  // launch-wide "blocks finished" counters for the |max| word rotation: a ring, so that two launches in flight
  // (compute and optimizer streams) never share one; each launch leaves its counter at zero
  static unsigned int *done_ring = nullptr;
  static unsigned int done_next = 0;
  constexpr unsigned int kDoneRing = 64;
  if (splits && !done_ring) {
    SK_CUDA(cudaMalloc((void **)&done_ring, kDoneRing * sizeof(unsigned int)));
    SK_CUDA(cudaMemset(done_ring, 0, kDoneRing * sizeof(unsigned int)));
  }
  int i = 0;
  while (i < n_tensors) {
    AdamArgs a;
    memset(&a, 0, sizeof(a));
    int n = 0, blocks = 0;
    double split_elems = 0;
    for (; i < n_tensors && n < kMaxTensors; ++i) {
      SK_REQUIRE(params[i] && grads[i] && m[i] && v[i] && sizes[i] >= 0, "sk_adam_step: tensor %d has a null pointer", i);
      if (sizes[i] == 0) continue;
      a.p[n] = params[i]; a.g[n] = grads[i]; a.m[n] = m[i]; a.v[n] = v[i]; a.size[n] = sizes[i];
      a.amax[n] = amax ? amax[i] : nullptr;
      if (splits && splits[i].amax2) {
        const sk_adam_split &sp = splits[i];
        a.amax_cur[n] = sp.amax2;
        a.amax[n] = sp.amax2 + 1;
        if (sp.hi) {
          SK_REQUIRE(sp.lo && sp.scale4, "sk_adam_step_split: tensor %d has hi without lo / scale4", i);
          SK_REQUIRE(sizes[i] % 4 == 0 && ((((uintptr_t)params[i]) | ((uintptr_t)grads[i]) | ((uintptr_t)m[i]) |
                                            ((uintptr_t)v[i])) & 15) == 0 &&
                         ((((uintptr_t)sp.hi) | ((uintptr_t)sp.lo)) & 7) == 0,
                     "sk_adam_step_split: tensor %d: the fused split needs 16-byte aligned arrays of a multiple of 4 elements", i);
          a.hi[n] = (__half *)sp.hi; a.lo[n] = (__half *)sp.lo; a.scale4[n] = sp.scale4;
          split_elems += (double)sizes[i];
        }
      }
      a.block_start[n] = blocks;
      blocks += (int)((sizes[i] + kChunk - 1) / kChunk);
      ++n;
    }
    a.block_start[n] = blocks;
    a.n = n;
    // Python floats meet float32 arrays as float32 scalars (NEP 50): optim.pyx:222-263
    a.lr = (float)lr; a.beta1 = (float)beta1; a.beta2 = (float)beta2;
    a.omb1 = (float)(1.0 - beta1); a.omb2 = (float)(1.0 - beta2);
    a.eps = (float)eps; a.wd = (float)weight_decay;
    a.bc1 = (float)one_minus_beta1_t; a.bc2 = (float)one_minus_beta2_t;
    a.grad_scale = (float)grad_scale;
    a.have_wd = weight_decay != 0.0; a.have_scale = grad_scale != 1.0; a.first = first_step;
    a.bias_state = bias_state;
    if (blocks == 0) continue;
    double elems = 0;
    for (int k = 0; k < n; ++k) elems += (double)a.size[k];
    a.total_blocks = blocks;
    // SOKET_B200_OPT_GRID_CAP = blocks per SM (0 = one block per chunk): a small persistent grid leaves issue
    // slots to the GEMMs this update overlaps with under data parallelism
    static const int cap_env = getenv("SOKET_B200_OPT_GRID_CAP") ? atoi(getenv("SOKET_B200_OPT_GRID_CAP")) : 0;
    const int cap = cap_env > 0 ? cap_env * ctx().num_sms : blocks;
    if (splits) {
      a.update_bound = (float)update_bound;
      a.done = done_ring + (done_next++ % kDoneRing);
    }
    ProfScope ps(SK_PROF_OPTIM, elems * (first_step ? 20.0 : 28.0) + split_elems * 4.0);
    adam_kernel<<<blocks < cap ? blocks : cap, kOT, 0, stream()>>>(a);
    SK_LAUNCH_CHECK();
  }
  return SK_OK;
}

int sk_adam_step(int n_tensors, float *const *params, const float *const *grads, float *const *m,
                 float *const *v, const int64_t *sizes, double lr, double beta1, double beta2,
                 double eps, double weight_decay, double one_minus_beta1_t,
                 double one_minus_beta2_t, int first_step, double grad_scale) {
  return adam_step_impl(n_tensors, params, grads, m, v, sizes, lr, beta1, beta2, eps, weight_decay,
                        one_minus_beta1_t, one_minus_beta2_t, first_step, grad_scale, nullptr, nullptr);
}

int sk_adam_step_amax(int n_tensors, float *const *params, const float *const *grads, float *const *m,
                      float *const *v, const int64_t *sizes, double lr, double beta1, double beta2,
                      double eps, double weight_decay, double one_minus_beta1_t, double one_minus_beta2_t,
                      const double *bias_state, int first_step, double grad_scale, unsigned int *const *amax) {
  SK_REQUIRE(amax, "sk_adam_step_amax: null word list");
  return adam_step_impl(n_tensors, params, grads, m, v, sizes, lr, beta1, beta2, eps, weight_decay,
                        one_minus_beta1_t, one_minus_beta2_t, first_step, grad_scale, bias_state, amax);
}

int sk_adam_step_split(int n_tensors, float *const *params, const float *const *grads, float *const *m,
                       float *const *v, const int64_t *sizes, double lr, double beta1, double beta2,
                       double eps, double weight_decay, double one_minus_beta1_t, double one_minus_beta2_t,
                       const double *bias_state, int first_step, double grad_scale, const sk_adam_split *splits,
                       double update_bound) {
  SK_REQUIRE(splits, "sk_adam_step_split: null split list");
  SK_REQUIRE(update_bound >= 0.0 && update_bound < 3.0e38, "sk_adam_step_split: update_bound must be a finite bound of |p_new - p_old|");
  return adam_step_impl(n_tensors, params, grads, m, v, sizes, lr, beta1, beta2, eps, weight_decay,
                        one_minus_beta1_t, one_minus_beta2_t, first_step, grad_scale, bias_state, nullptr, splits,
                        update_bound);
}

int sk_adam_step_dev(int n_tensors, float *const *params, const float *const *grads, float *const *m,
                     float *const *v, const int64_t *sizes, double lr, double beta1, double beta2,
                     double eps, double weight_decay, const double *bias_state, int first_step,
                     double grad_scale) {
  SK_REQUIRE(bias_state, "sk_adam_step_dev: null bias state");
  return adam_step_impl(n_tensors, params, grads, m, v, sizes, lr, beta1, beta2, eps, weight_decay,
                        0.0, 0.0, first_step, grad_scale, bias_state, nullptr);
}

int sk_adam_bias_advance(double *bias_state, double beta1, double beta2) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(bias_state, "sk_adam_bias_advance: null bias state");
  adam_bias_advance_kernel<<<1, 1, 0, stream()>>>(bias_state, beta1, beta2);
  SK_LAUNCH_CHECK();
  return SK_OK;
}

}  // extern "C"
