// nn_fused.cu -- fused LayerNorm / BatchNorm / softmax-cross-entropy / ReLU-residual /
// dropout / bias-gradient kernels.
//
// Each kernel replaces a SEQUENCE of backend array calls in the reference
// (soket/tensor/ops/forward.pyx:224-353, backward.pyx:849-1132,
// soket/nn/prototypes.pyx:272-273,746-760, soket/autodiff.pyx:30-101) by one or
// two HBM-bound passes.  The arithmetic follows the reference's formulas
// (biased variance, (var+eps)^-0.5, dgamma/dbeta summed over axis 0, 1/B as a
// float scalar) so results agree to fp32 rounding (<= 1e-5 rel).
//
// Algorithmic bytes (fp32): LN fwd 8 B/elem, LN bwd 12 B/elem (+4 with a stored
// ReLU mask), BN fwd 12 B/elem, BN bwd 20 B/elem, CE 8 B/elem, add+relu 12 B/elem.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "matmul_split.cuh"
#include "reduce.cuh"

namespace sk {

constexpr int kNT = 256;
static inline bool al16(const void *p) { return (((uintptr_t)p) & 15) == 0; }

// sum over a group of TPR threads (32 = one warp, 256 = the whole block);
// every thread of the group receives the result.
template <int TPR>
__device__ __forceinline__ float group_sum(float v, float *smem) {
  v = warp_sum(v);
  if (TPR == 32) return v;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < kNT / 32; ++i) t += smem[i];
  __syncthreads();
  return t;
}

// Philox is private to rng.cu; dropout uses a cheap counter hash (PCG-style
// output permutation over a Weyl sequence keyed by seed): one draw per element.
__device__ __forceinline__ uint32_t hash_u32(uint64_t idx, uint64_t seed) {
  uint64_t z = idx * 0x9E3779B97F4A7C15ull + seed;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (uint32_t)(z >> 32);
}

// Dropout folded into the LayerNorm kernels (Linear - LayerNorm - ReLU - Dropout of the residual
// block, model.py:24-31): the same per-element Bernoulli draw as dropout_kernel / dropout_bwd_kernel
// (element index = row * C + column, seed mixed with the graph-replay epoch), applied to the
// LayerNorm(+ReLU) OUTPUT in the forward epilogue and to the incoming adjoint in the backward
// prologue.  keep >= 1 switches it off.
struct DropSpec {
  float keep, r_keep;
  uint64_t seed;
  const uint64_t *epoch;
};
__device__ __forceinline__ uint64_t drop_seed(const DropSpec &d) {
  return d.keep < 1.f ? d.seed + *d.epoch * 0xD1B54A32D192ED03ull : 0ull;
}
// ONE 64-bit hash per aligned group of four elements, 16 bits per element: the hash (three 64-bit
// multiplies) was the visible cost once the draw moved into the LayerNorm kernels.  P(keep) is
// floor(keep * 65536) / 65536, within 1.6e-5 of `keep`.  e0 is a multiple of 4.
__device__ __forceinline__ float4 drop_mask4(int64_t e0, uint64_t seed, float keep) {
  uint64_t z = ((uint64_t)e0 >> 2) * 0x9E3779B97F4A7C15ull + seed;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  const uint32_t thr = (uint32_t)(keep * 65536.0f);
  const uint32_t lo = (uint32_t)z, hi = (uint32_t)(z >> 32);
  float4 m;
  m.x = (lo & 0xFFFFu) < thr ? 1.f : 0.f;
  m.y = (lo >> 16) < thr ? 1.f : 0.f;
  m.z = (hi & 0xFFFFu) < thr ? 1.f : 0.f;
  m.w = (hi >> 16) < thr ? 1.f : 0.f;
  return m;
}
// forward: (y * mask) * (1/keep) (prototypes.pyx:758); backward: (adj * (1/keep)) * mask
__device__ __forceinline__ void drop_fwd4(float4 &o, const float4 &m, float r_keep) {
  o.x = (o.x * m.x) * r_keep; o.y = (o.y * m.y) * r_keep; o.z = (o.z * m.z) * r_keep; o.w = (o.w * m.w) * r_keep;
}
__device__ __forceinline__ void drop_bwd4(float4 &a, const float4 &m, float r_keep) {
  a.x = (a.x * r_keep) * m.x; a.y = (a.y * r_keep) * m.y; a.z = (a.z * r_keep) * m.z; a.w = (a.w * r_keep) * m.w;
}

__device__ __forceinline__ float affine(float xs, float r, float g, float b) {
  // gamma * (xs * r) + beta, in the reference's order (forward.pyx:325,343,352)
  return __fadd_rn(__fmul_rn(g, __fmul_rn(xs, r)), b);
}

// Optional by-products of the LayerNorm kernels for the fp16x3 GEMM that consumes their result
// (sk_layernorm_*_ex).  Forward: the output also as fp16 hi / lo with ONE power-of-two scale, chosen
// BEFORE any element is computed from a bound of the output:
//     |gamma * norm + beta| <= max|gamma| * sqrt(C) + max|beta|       (|norm| < sqrt(C): one element can
//     carry at most the whole variance), plus the residual's own bound, times 1/keep under dropout
// -- no pass over the data, no second kernel.  Backward: the bit pattern of max |dx| into a device
// word (atomicMax), from which the adjoint's split pass takes its scale without a pass of its own.
struct LnExtras {
  __half *hi, *lo;            // (R, C) fp16, or null
  float *scale;               // out: {scale, 1/scale, bound, 0}
  const float *res_scale;     // the residual's scale4 (its [2] = bound of |residual|)
  unsigned int *dx_amax;      // backward only, or null
};

// max over the block of two values; every thread gets both.  scratch: 16 floats.
__device__ __forceinline__ void block_max2(float &a, float &b, float *scratch) {
  a = warp_max(a); b = warp_max(b);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { scratch[warp] = a; scratch[8 + warp] = b; }
  __syncthreads();
  a = scratch[0]; b = scratch[8];
#pragma unroll
  for (int i = 1; i < kNT / 32; ++i) { a = fmaxf(a, scratch[i]); b = fmaxf(b, scratch[8 + i]); }
  __syncthreads();
}

// the scale of the output split (all blocks compute the same value; block 0 publishes it)
__device__ __forceinline__ float ln_split_scale(const LnExtras &ex, const float *gamma, const float *beta, int C,
                                                bool has_residual, float r_keep, float *scratch) {
  float gm = gamma ? 0.f : 1.f, bm = 0.f;
  for (int i = threadIdx.x; i < C; i += kNT) {
    if (gamma) gm = fmaxf(gm, fabsf(__ldg(gamma + i)));
    if (beta) bm = fmaxf(bm, fabsf(__ldg(beta + i)));
  }
  block_max2(gm, bm, scratch);
  float bound = gm * sqrtf((float)C) + bm;
  if (has_residual) bound += ex.res_scale[2];
  bound *= 1.0001f * r_keep;          // fp32 rounding of the output; dropout scales kept values by 1/keep
  float sc, inv;
  pow2_scale(bound, sc, inv);
  if (blockIdx.x == 0 && threadIdx.x == 0) { ex.scale[0] = sc; ex.scale[1] = inv; ex.scale[2] = bound; ex.scale[3] = 0.f; }
  return sc;
}

__device__ __forceinline__ void ln_store_split(const LnExtras &ex, int64_t elem, const float4 &o, float sc) {
  uint2 h, l;
  split4(o, sc, h, l);
  *reinterpret_cast<uint2 *>(ex.hi + elem) = h;     // read next by the GEMM's TMA loads: keep in L2
  *reinterpret_cast<uint2 *>(ex.lo + elem) = l;
}

// ------------------------------------------------------------------- LayerNorm
// TPR threads cooperate on one row; each holds VPT float4 (the row lives in
// registers between the statistics and the normalisation: x is read once).
template <int TPR, int VPT>
__global__ void __launch_bounds__(kNT)
ln_fwd_kernel(const float *__restrict__ x, const float *__restrict__ gamma,
              const float *__restrict__ beta, const float *__restrict__ residual,
              float *__restrict__ y, float *__restrict__ mean_out, float *__restrict__ rstd_out,
              int64_t R, int C, float eps, int relu, const DropSpec drop, const LnExtras ex) {
  __shared__ float smem[16];
  const uint64_t dseed = drop_seed(drop);
  const float sc = ex.hi ? ln_split_scale(ex, gamma, beta, C, residual != nullptr, drop.keep < 1.f ? drop.r_keep : 1.f, smem) : 1.f;
  constexpr int RPB = kNT / TPR;  // rows per block
  const int t = threadIdx.x % TPR;
  const int C4 = C >> 2;
  const float n_obs = (float)C;
  for (int64_t row0 = (int64_t)blockIdx.x * RPB; row0 < R; row0 += (int64_t)gridDim.x * RPB) {
    const int64_t row = row0 + threadIdx.x / TPR;
    const bool live = row < R;  // whole group shares `live`; block-level syncs still reached
    const float4 *xr = reinterpret_cast<const float4 *>(x + row * C);
    float4 v[VPT];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      int i = t + j * TPR;
      if (live && i < C4) {
        v[j] = ld_stream(xr + i);
        s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
      } else {
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const float mean = group_sum<TPR>(s, smem) / n_obs;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      int i = t + j * TPR;
      if (live && i < C4) {
        v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
        q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
      }
    }
    const float var = group_sum<TPR>(q, smem) / n_obs;
    const float r = 1.0f / sqrtf(var + eps);
    if (live && t == 0) {
      mean_out[row] = mean;
      rstd_out[row] = r;
    }
    float4 *yr = reinterpret_cast<float4 *>(y + row * C);
    const float4 *rr = reinterpret_cast<const float4 *>(residual ? residual + row * C : nullptr);
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      int i = t + j * TPR;
      if (live && i < C4) {
        float4 g = gamma ? __ldg(reinterpret_cast<const float4 *>(gamma) + i) : make_float4(1.f, 1.f, 1.f, 1.f);
        float4 b = beta ? __ldg(reinterpret_cast<const float4 *>(beta) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 o;
        o.x = affine(v[j].x, r, g.x, b.x); o.y = affine(v[j].y, r, g.y, b.y);
        o.z = affine(v[j].z, r, g.z, b.z); o.w = affine(v[j].w, r, g.w, b.w);
        if (residual) {
          float4 rs = ld_stream(rr + i);
          o.x = __fadd_rn(rs.x, o.x); o.y = __fadd_rn(rs.y, o.y);
          o.z = __fadd_rn(rs.z, o.z); o.w = __fadd_rn(rs.w, o.w);
        }
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        if (drop.keep < 1.f) drop_fwd4(o, drop_mask4(row * C + 4 * (int64_t)i, dseed, drop.keep), drop.r_keep);
        st_stream(yr + i, o);
        if (ex.hi) ln_store_split(ex, row * C + 4 * (int64_t)i, o, sc);
      }
    }
  }
}

// Backward.  mask_mode: 0 none; 1 ReLU mask recomputed from (x, mean, rstd, gamma,
// beta) -- LN followed directly by ReLU; 2 ReLU mask read from y_out (fused
// residual+ReLU output).  dgamma/dbeta partials: one row of `part_g`/`part_b` per
// thread group, column-reduced afterwards.
template <int TPR, int VPT>
__global__ void __launch_bounds__(kNT)
ln_bwd_kernel(const float *__restrict__ adj, const float *__restrict__ x,
              const float *__restrict__ gamma, const float *__restrict__ beta,
              const float *__restrict__ mean_in, const float *__restrict__ rstd_in,
              const float *__restrict__ y_out, int mask_mode, float *__restrict__ dx,
              float *__restrict__ dresidual, float *__restrict__ part_g,
              float *__restrict__ part_b, int64_t R, int C, const DropSpec drop, unsigned int *__restrict__ dx_amax) {
  __shared__ float smem[kNT / 32];
  const uint64_t dseed = drop_seed(drop);
  float amax = 0.f;
  constexpr int RPB = kNT / TPR;
  const int t = threadIdx.x % TPR;
  const int grp = threadIdx.x / TPR;
  const int C4 = C >> 2;
  const float inv_n = 1.0f / (float)C;
  float4 ag[VPT], ab[VPT];  // dgamma / dbeta accumulators for this thread's columns
#pragma unroll
  for (int j = 0; j < VPT; ++j) { ag[j] = make_float4(0.f, 0.f, 0.f, 0.f); ab[j] = ag[j]; }

  for (int64_t row0 = (int64_t)blockIdx.x * RPB; row0 < R; row0 += (int64_t)gridDim.x * RPB) {
    const int64_t row = row0 + grp;
    const bool live = row < R;
    const float mean = live ? mean_in[row] : 0.f;
    const float r = live ? rstd_in[row] : 0.f;
    const float4 *ar = reinterpret_cast<const float4 *>(adj + row * C);
    const float4 *xr = reinterpret_cast<const float4 *>(x + row * C);
    const float4 *yr = reinterpret_cast<const float4 *>(y_out ? y_out + row * C : nullptr);
    float4 a[VPT], xs[VPT];
    float s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      int i = t + j * TPR;
      if (live && i < C4) {
        a[j] = ld_stream(ar + i);
        if (drop.keep < 1.f) drop_bwd4(a[j], drop_mask4(row * C + 4 * (int64_t)i, dseed, drop.keep), drop.r_keep);
        xs[j] = ld_stream(xr + i);
        xs[j].x -= mean; xs[j].y -= mean; xs[j].z -= mean; xs[j].w -= mean;
        float4 g = gamma ? __ldg(reinterpret_cast<const float4 *>(gamma) + i) : make_float4(1.f, 1.f, 1.f, 1.f);
        if (mask_mode == 1) {
          float4 b = beta ? __ldg(reinterpret_cast<const float4 *>(beta) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (!(affine(xs[j].x, r, g.x, b.x) > 0.f)) a[j].x = 0.f;
          if (!(affine(xs[j].y, r, g.y, b.y) > 0.f)) a[j].y = 0.f;
          if (!(affine(xs[j].z, r, g.z, b.z) > 0.f)) a[j].z = 0.f;
          if (!(affine(xs[j].w, r, g.w, b.w) > 0.f)) a[j].w = 0.f;
        } else if (mask_mode == 2) {
          float4 yo = ld_stream(yr + i);
          if (!(yo.x > 0.f)) a[j].x = 0.f;
          if (!(yo.y > 0.f)) a[j].y = 0.f;
          if (!(yo.z > 0.f)) a[j].z = 0.f;
          if (!(yo.w > 0.f)) a[j].w = 0.f;
        }
        if (dresidual) st_stream(reinterpret_cast<float4 *>(dresidual + row * C) + i, a[j]);
        // dgamma += norm * adj ; dbeta += adj   (backward.pyx:1058-1078)
        ag[j].x += (xs[j].x * r) * a[j].x; ag[j].y += (xs[j].y * r) * a[j].y;
        ag[j].z += (xs[j].z * r) * a[j].z; ag[j].w += (xs[j].w * r) * a[j].w;
        ab[j].x += a[j].x; ab[j].y += a[j].y; ab[j].z += a[j].z; ab[j].w += a[j].w;
        // dxn = adj * gamma (kept in a[])
        a[j].x *= g.x; a[j].y *= g.y; a[j].z *= g.z; a[j].w *= g.w;
        s1 += (a[j].x * xs[j].x + a[j].y * xs[j].y) + (a[j].z * xs[j].z + a[j].w * xs[j].w);
        s2 += (a[j].x + a[j].y) + (a[j].z + a[j].w);
        s3 += (xs[j].x + xs[j].y) + (xs[j].z + xs[j].w);
      } else {
        a[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        xs[j] = a[j];
      }
    }
    s1 = group_sum<TPR>(s1, smem);
    s2 = group_sum<TPR>(s2, smem);
    s3 = group_sum<TPR>(s3, smem);
    // backward.pyx:1094-1126
    const float dvar = s1 * (-0.5f * ((r * r) * r));
    const float dmean = (-r) * s2 + dvar * (inv_n * (-2.0f * s3));
    const float c0 = inv_n * dmean;
    const float c2 = dvar * (2.0f * inv_n);
    float4 *dr = reinterpret_cast<float4 *>(dx + row * C);
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      int i = t + j * TPR;
      if (live && i < C4) {
        float4 o;
        o.x = c0 + (a[j].x * r + c2 * xs[j].x); o.y = c0 + (a[j].y * r + c2 * xs[j].y);
        o.z = c0 + (a[j].z * r + c2 * xs[j].z); o.w = c0 + (a[j].w * r + c2 * xs[j].w);
        st_stream(dr + i, o);
        amax = fmaxf(fmaxf(amax, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
      }
    }
  }
  if (dx_amax) {
    amax = warp_max(amax);
    if ((threadIdx.x & 31) == 0 && amax > 0.f) atomicMax(dx_amax, __float_as_uint(amax));
  }
  if (part_g) {
    const int64_t prow = (int64_t)blockIdx.x * RPB + grp;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      int i = t + j * TPR;
      if (i < C4) {
        reinterpret_cast<float4 *>(part_g + prow * C)[i] = ag[j];
        reinterpret_cast<float4 *>(part_b + prow * C)[i] = ab[j];
      }
    }
  }
}


// ------------------------------------------------------------------- LayerNorm, staged rows
// Same arithmetic as ln_fwd_kernel / ln_bwd_kernel for rows longer than 512 columns, with the
// input rows brought into shared memory by the bulk-copy engine (cp.async.bulk + mbarrier
// complete_tx) a few rows AHEAD of the row being reduced.  The register-resident kernels are
// latency bound -- one row per 256-thread block in flight, 120 registers in the backward
// (ncu: 3.3-3.5 TB/s, 25 % occupancy) -- whereas here the bytes in flight are set by the
// stage count, not by occupancy.  One block walks rows blockIdx.x, +gridDim.x, ...
__device__ __forceinline__ uint32_t ln_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ln_mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ln_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void ln_mbar_expect(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ln_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ln_mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(ln_smem_u32(bar)), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();   // a protocol bug surfaces as a CUDA error instead of a hang
}
__device__ __forceinline__ void ln_bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(ln_smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(ln_smem_u32(bar)) : "memory");
}

constexpr int kLnMaxStages = 4;
constexpr int kLnHeader = 256;   // reduction scratch (2 x 3 x 8 floats) + kLnMaxStages mbarriers

// Block-wide sums of up to three values with ONE __syncthreads: the scratch alternates
// between two buffers (`buf`), so the next reduction never overwrites values a slower
// warp is still reading (a warp cannot run two barriers ahead of another).
template <int N>
__device__ __forceinline__ void block_sum_n(float (&v)[N], float *red, int buf) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float *r = red + buf * 24;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    v[k] = warp_sum(v[k]);
    if (lane == 0) r[k * 8 + warp] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kNT / 32; ++i) t += r[k * 8 + i];
    v[k] = t;
  }
}

template <int VPT>
__global__ void __launch_bounds__(kNT, 2)
ln_bwd_staged_kernel(const float *__restrict__ adj, const float *__restrict__ x,
                     const float *__restrict__ gamma, const float *__restrict__ beta,
                     const float *__restrict__ mean_in, const float *__restrict__ rstd_in,
                     const float *__restrict__ y_out, int mask_mode, float *__restrict__ dx,
                     float *__restrict__ dresidual, float *__restrict__ part_g,
                     float *__restrict__ part_b, int64_t R, int C, int stages, int *__restrict__ sched,
                     const DropSpec drop, unsigned int *__restrict__ dx_amax) {
  extern __shared__ __align__(128) uint8_t ln_sm[];
  const uint64_t dseed = drop_seed(drop);
  float amax = 0.f;
  float *red = reinterpret_cast<float *>(ln_sm);
  uint64_t *full = reinterpret_cast<uint64_t *>(ln_sm + 192);
  volatile int *row_ring = reinterpret_cast<volatile int *>(ln_sm + 224);   // [kLnMaxStages] claimed rows
  float *data = reinterpret_cast<float *>(ln_sm + kLnHeader);
  const int nbuf = mask_mode == 2 ? 3 : 2;
  const uint32_t row_bytes = (uint32_t)C * 4u;
  const int t = threadIdx.x;
  const int C4 = C >> 2;
  const float inv_n = 1.0f / (float)C;

  // Rows are CLAIMED from a global counter as their loads are issued (the first one is
  // static): a block that starts late -- its SM was running the overlapped NCCL all-reduce --
  // takes fewer rows instead of stretching the kernel (static striding cost +0.8 ms per step
  // under data parallelism).  The claimed row travels to the consumers in `row_ring`, published
  // by the same mbarrier phase that says its bytes have landed; -1 ends the walk.
  int64_t claimed = 0;   // thread 0: rows claimed so far
  auto issue = [&](int64_t it) {   // thread 0 only; `it` counts this block's claims
    const int sl = (int)(it % stages);
    int64_t row = it == 0 ? (int64_t)blockIdx.x : (int64_t)gridDim.x + atomicAdd(sched, 1);
    if (row >= R) {
      row_ring[sl] = -1;
      ln_mbar_expect(&full[sl], 0);   // plain arrive: the phase completes with no bytes
      return false;
    }
    row_ring[sl] = (int)row;
    float *dst = data + (size_t)sl * nbuf * C;
    ln_mbar_expect(&full[sl], row_bytes * nbuf);
    ln_bulk_load(dst, adj + row * C, row_bytes, &full[sl]);
    ln_bulk_load(dst + C, x + row * C, row_bytes, &full[sl]);
    if (nbuf == 3) ln_bulk_load(dst + 2 * C, y_out + row * C, row_bytes, &full[sl]);
    return true;
  };
  if (t == 0) {
    for (int i = 0; i < stages; ++i) ln_mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  bool more = true;      // thread 0: the counter has not run out yet
  if (t == 0)
    for (; claimed < stages - 1 && more; ++claimed) more = issue(claimed);

  float4 ag[VPT], ab[VPT];
#pragma unroll
  for (int j = 0; j < VPT; ++j) { ag[j] = make_float4(0.f, 0.f, 0.f, 0.f); ab[j] = ag[j]; }

  for (int64_t it = 0;; ++it) {
    // the slot refilled here was read in iteration it-1, before that iteration's block sync
    if (t == 0 && more) { more = issue(claimed); ++claimed; }
    const int sl = (int)(it % stages);
    ln_mbar_wait(&full[sl], (uint32_t)((it / stages) & 1));
    const int64_t row = row_ring[sl];
    if (row < 0) break;
    const float mean = mean_in[row];
    const float r = rstd_in[row];
    const float4 *ar = reinterpret_cast<const float4 *>(data + (size_t)sl * nbuf * C);
    const float4 *xr = ar + C4;
    const float4 *yr = ar + 2 * C4;
    float4 a[VPT], xs[VPT];
    float s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int i = t + j * kNT;
      if (i < C4) {
        a[j] = ar[i];
        if (drop.keep < 1.f) drop_bwd4(a[j], drop_mask4(row * C + 4 * (int64_t)i, dseed, drop.keep), drop.r_keep);
        xs[j] = xr[i];
        xs[j].x -= mean; xs[j].y -= mean; xs[j].z -= mean; xs[j].w -= mean;
        float4 g = gamma ? __ldg(reinterpret_cast<const float4 *>(gamma) + i) : make_float4(1.f, 1.f, 1.f, 1.f);
        if (mask_mode == 1) {
          float4 b = beta ? __ldg(reinterpret_cast<const float4 *>(beta) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (!(affine(xs[j].x, r, g.x, b.x) > 0.f)) a[j].x = 0.f;
          if (!(affine(xs[j].y, r, g.y, b.y) > 0.f)) a[j].y = 0.f;
          if (!(affine(xs[j].z, r, g.z, b.z) > 0.f)) a[j].z = 0.f;
          if (!(affine(xs[j].w, r, g.w, b.w) > 0.f)) a[j].w = 0.f;
        } else if (mask_mode == 2) {
          const float4 yo = yr[i];
          if (!(yo.x > 0.f)) a[j].x = 0.f;
          if (!(yo.y > 0.f)) a[j].y = 0.f;
          if (!(yo.z > 0.f)) a[j].z = 0.f;
          if (!(yo.w > 0.f)) a[j].w = 0.f;
        }
        if (dresidual) st_stream(reinterpret_cast<float4 *>(dresidual + row * C) + i, a[j]);
        // dgamma += norm * adj ; dbeta += adj   (backward.pyx:1058-1078)
        ag[j].x += (xs[j].x * r) * a[j].x; ag[j].y += (xs[j].y * r) * a[j].y;
        ag[j].z += (xs[j].z * r) * a[j].z; ag[j].w += (xs[j].w * r) * a[j].w;
        ab[j].x += a[j].x; ab[j].y += a[j].y; ab[j].z += a[j].z; ab[j].w += a[j].w;
        a[j].x *= g.x; a[j].y *= g.y; a[j].z *= g.z; a[j].w *= g.w;
        s1 += (a[j].x * xs[j].x + a[j].y * xs[j].y) + (a[j].z * xs[j].z + a[j].w * xs[j].w);
        s2 += (a[j].x + a[j].y) + (a[j].z + a[j].w);
        s3 += (xs[j].x + xs[j].y) + (xs[j].z + xs[j].w);
      } else {
        a[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        xs[j] = a[j];
      }
    }
    float sums[3] = {s1, s2, s3};
    block_sum_n<3>(sums, red, (int)(it & 1));
    s1 = sums[0]; s2 = sums[1]; s3 = sums[2];
    // backward.pyx:1094-1126
    const float dvar = s1 * (-0.5f * ((r * r) * r));
    const float dmean = (-r) * s2 + dvar * (inv_n * (-2.0f * s3));
    const float c0 = inv_n * dmean;
    const float c2 = dvar * (2.0f * inv_n);
    float4 *dr = reinterpret_cast<float4 *>(dx + row * C);
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int i = t + j * kNT;
      if (i < C4) {
        float4 o;
        o.x = c0 + (a[j].x * r + c2 * xs[j].x); o.y = c0 + (a[j].y * r + c2 * xs[j].y);
        o.z = c0 + (a[j].z * r + c2 * xs[j].z); o.w = c0 + (a[j].w * r + c2 * xs[j].w);
        st_stream(dr + i, o);
        amax = fmaxf(fmaxf(amax, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
      }
    }
  }
  if (dx_amax) {
    amax = warp_max(amax);
    if ((threadIdx.x & 31) == 0 && amax > 0.f) atomicMax(dx_amax, __float_as_uint(amax));
  }
  if (part_g) {
    const int64_t prow = blockIdx.x;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int i = t + j * kNT;
      if (i < C4) {
        reinterpret_cast<float4 *>(part_g + prow * C)[i] = ag[j];
        reinterpret_cast<float4 *>(part_b + prow * C)[i] = ab[j];
      }
    }
  }
  if (t == 0) {   // the last block to finish re-arms the row counter for the next launch
    __threadfence();
    if (atomicAdd(sched + 1, 1) == (int)gridDim.x - 1) {
      sched[0] = 0;
      sched[1] = 0;
      __threadfence();
    }
  }
}

template <int VPT>
__global__ void __launch_bounds__(kNT, 2)
ln_fwd_staged_kernel(const float *__restrict__ x, const float *__restrict__ gamma,
                     const float *__restrict__ beta, const float *__restrict__ residual,
                     float *__restrict__ y, float *__restrict__ mean_out, float *__restrict__ rstd_out,
                     int64_t R, int C, float eps, int relu, int stages, const DropSpec drop, const LnExtras ex) {
  extern __shared__ __align__(128) uint8_t ln_sm[];
  const uint64_t dseed = drop_seed(drop);
  float *red = reinterpret_cast<float *>(ln_sm);
  const float sc = ex.hi ? ln_split_scale(ex, gamma, beta, C, residual != nullptr, drop.keep < 1.f ? drop.r_keep : 1.f, red) : 1.f;
  uint64_t *full = reinterpret_cast<uint64_t *>(ln_sm + 192);
  float *data = reinterpret_cast<float *>(ln_sm + kLnHeader);
  const int nbuf = residual ? 2 : 1;
  const uint32_t row_bytes = (uint32_t)C * 4u;
  const int t = threadIdx.x;
  const int C4 = C >> 2;
  const float n_obs = (float)C;
  const int64_t n_it = (R - blockIdx.x + gridDim.x - 1) / gridDim.x;
  auto issue = [&](int64_t it) {
    const int sl = (int)(it % stages);
    const int64_t row = blockIdx.x + it * gridDim.x;
    float *dst = data + (size_t)sl * nbuf * C;
    ln_mbar_expect(&full[sl], row_bytes * nbuf);
    ln_bulk_load(dst, x + row * C, row_bytes, &full[sl]);
    if (nbuf == 2) ln_bulk_load(dst + C, residual + row * C, row_bytes, &full[sl]);
  };
  if (t == 0) {
    for (int i = 0; i < stages; ++i) ln_mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (t == 0)
    for (int64_t it = 0; it < stages - 1 && it < n_it; ++it) issue(it);
  for (int64_t it = 0; it < n_it; ++it) {
    if (t == 0 && it + stages - 1 < n_it) issue(it + stages - 1);
    const int sl = (int)(it % stages);
    const int64_t row = blockIdx.x + it * gridDim.x;
    ln_mbar_wait(&full[sl], (uint32_t)((it / stages) & 1));
    const float4 *xr = reinterpret_cast<const float4 *>(data + (size_t)sl * nbuf * C);
    const float4 *rr = xr + C4;
    float4 v[VPT], rs[VPT];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int i = t + j * kNT;
      if (i < C4) {
        v[j] = xr[i];
        if (nbuf == 2) rs[j] = rr[i];
        s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
      } else {
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    float s_[1] = {s};
    block_sum_n<1>(s_, red, 0);
    const float mean = s_[0] / n_obs;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int i = t + j * kNT;
      if (i < C4) {
        v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
        q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
      }
    }
    float q_[1] = {q};
    block_sum_n<1>(q_, red, 1);
    const float var = q_[0] / n_obs;
    const float r = 1.0f / sqrtf(var + eps);
    if (t == 0) {
      mean_out[row] = mean;
      rstd_out[row] = r;
    }
    float4 *yr = reinterpret_cast<float4 *>(y + row * C);
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int i = t + j * kNT;
      if (i < C4) {
        float4 g = gamma ? __ldg(reinterpret_cast<const float4 *>(gamma) + i) : make_float4(1.f, 1.f, 1.f, 1.f);
        float4 b = beta ? __ldg(reinterpret_cast<const float4 *>(beta) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 o;
        o.x = affine(v[j].x, r, g.x, b.x); o.y = affine(v[j].y, r, g.y, b.y);
        o.z = affine(v[j].z, r, g.z, b.z); o.w = affine(v[j].w, r, g.w, b.w);
        if (nbuf == 2) {
          o.x = __fadd_rn(rs[j].x, o.x); o.y = __fadd_rn(rs[j].y, o.y);
          o.z = __fadd_rn(rs[j].z, o.z); o.w = __fadd_rn(rs[j].w, o.w);
        }
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        if (drop.keep < 1.f) drop_fwd4(o, drop_mask4(row * C + 4 * (int64_t)i, dseed, drop.keep), drop.r_keep);
        st_stream(yr + i, o);
        if (ex.hi) ln_store_split(ex, row * C + 4 * (int64_t)i, o, sc);
      }
    }
  }
}

// stage count / blocks per SM for the staged kernels; stages == 0: use the register kernels
static void ln_stage_plan(int nbuf, int C, int &stages, int &blocks_per_sm, size_t &smem, int max_blocks = 2) {
  static const int off = getenv("SOKET_B200_LN_STAGED") ? !atoi(getenv("SOKET_B200_LN_STAGED")) : 0;
  const size_t stage_bytes = (size_t)nbuf * C * 4;
  stages = 0;
  if (off || C <= 512) return;
  const size_t budget2 = 100 * 1024, budget1 = 200 * 1024, budget3 = 66 * 1024;
  if (max_blocks >= 3 && stage_bytes * 3 + kLnHeader <= budget3) {
    blocks_per_sm = 3;
    stages = (int)((budget3 - kLnHeader) / stage_bytes);
  } else if (stage_bytes * 2 + kLnHeader <= budget2) {
    blocks_per_sm = 2;
    stages = (int)((budget2 - kLnHeader) / stage_bytes);
  } else if (stage_bytes * 2 + kLnHeader <= budget1) {
    blocks_per_sm = 1;
    stages = (int)((budget1 - kLnHeader) / stage_bytes);
  } else {
    return;
  }
  if (stages > kLnMaxStages) stages = kLnMaxStages;
  smem = stages * stage_bytes + kLnHeader;
}


// dgamma / dbeta from the per-block partial rows of the LayerNorm backward: ONE launch
// column-sums both (P x C) matrices (they are L2 resident: just written).  Block = 8 float4
// column groups x 32 row lanes; the 32 lane sums are added in a fixed order (deterministic).
__global__ void __launch_bounds__(kNT)
ln_param_grads_kernel(const float *__restrict__ part_g, const float *__restrict__ part_b,
                      float *__restrict__ dgamma, float *__restrict__ dbeta, int64_t P, int C) {
  __shared__ float4 sm[32][8];
  const float *part = blockIdx.y == 0 ? part_g : part_b;
  float *out = blockIdx.y == 0 ? dgamma : dbeta;
  if (out == nullptr) return;
  const int cg = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c = (blockIdx.x * 8 + cg) * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < C) {
    for (int64_t r = rl; r < P; r += 32) {
      const float4 v = *reinterpret_cast<const float4 *>(part + r * C + c);
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
    const int cc = (blockIdx.x * 8 + g) * 4 + k;
    if (cc < C) out[cc] = t;
  }
}

template <int TPR, int VPT>
static int ln_fwd_launch(const float *x, const float *gamma, const float *beta, const float *residual,
                         float *y, float *mean, float *rstd, int64_t R, int C, float eps, int relu,
                         const DropSpec drop, const LnExtras ex) {
  constexpr int RPB = kNT / TPR;
  if (TPR == kNT) {
    int stages, bps = 1;
    size_t smem = 0;
    ln_stage_plan(residual ? 2 : 1, C, stages, bps, smem, VPT <= 4 ? 3 : 2);
    if (stages >= 2) {
      auto kern = ln_fwd_staged_kernel<VPT>;
      static bool attr_set = false;
      if (!attr_set) {
        SK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + kLnHeader));
        attr_set = true;
      }
      const int64_t cap = (int64_t)ctx().num_sms * bps;
      const int grid = (int)(R < cap ? R : cap);
      ProfScope ps(SK_PROF_LN_FWD, (double)R * C * ((residual ? 12.0 : 8.0) + (ex.hi ? 4.0 : 0.0)));
      kern<<<grid, kNT, smem, stream()>>>(x, gamma, beta, residual, y, mean, rstd, R, C, eps, relu, stages, drop, ex);
      SK_LAUNCH_CHECK();
      return SK_OK;
    }
  }
  int grid = grid_for(R, RPB, 8);
  ProfScope ps(SK_PROF_LN_FWD, (double)R * C * ((residual ? 12.0 : 8.0) + (ex.hi ? 4.0 : 0.0)));
  ln_fwd_kernel<TPR, VPT><<<grid, kNT, 0, stream()>>>(x, gamma, beta, residual, y, mean, rstd, R, C, eps, relu, drop, ex);
  SK_LAUNCH_CHECK();
  return SK_OK;
}

template <int TPR, int VPT>
static int ln_bwd_launch(const float *adj, const float *x, const float *gamma, const float *beta,
                         const float *mean, const float *rstd, const float *y_out, int mask_mode,
                         float *dx, float *dresidual, float *dgamma, float *dbeta, int64_t R, int C,
                         const DropSpec drop, unsigned int *dx_amax) {
  constexpr int RPB = kNT / TPR;
  // persistent: each group walks many rows so the dgamma/dbeta partial matrix stays small
  int64_t need = (R + RPB - 1) / RPB;
  int64_t cap = (int64_t)ctx().num_sms * (TPR == 32 ? 4 : 2);
  int stages = 0, bps = 1;
  size_t smem = 0;
  if (TPR == kNT && al16(adj) && al16(x) && (mask_mode != 2 || al16(y_out))) ln_stage_plan(mask_mode == 2 ? 3 : 2, C, stages, bps, smem);
  if (stages >= 2) cap = (int64_t)ctx().num_sms * bps;
  int grid = (int)(need < cap ? need : cap);
  float *part = nullptr;
  const int64_t P = (int64_t)grid * RPB;
  const bool want_params = dgamma != nullptr || dbeta != nullptr;
  int rc;
  if (want_params) {
    if ((rc = sk_malloc((size_t)(2 * P * C) * sizeof(float), (void **)&part))) return rc;
  }
  {
    ProfScope ps(SK_PROF_LN_BWD, (double)R * C * (12.0 + (mask_mode == 2 ? 4.0 : 0.0) + (dresidual ? 4.0 : 0.0)));
    if (stages >= 2) {
      auto kern = ln_bwd_staged_kernel<VPT>;
      static bool attr_set = false;
      if (!attr_set) {
        SK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + kLnHeader));
        attr_set = true;
      }
      static int *sched_dev = nullptr;
      if (!sched_dev) {
        SK_CUDA(cudaMalloc((void **)&sched_dev, 2 * sizeof(int)));
        SK_CUDA(cudaMemsetAsync(sched_dev, 0, 2 * sizeof(int), stream()));
      }
      kern<<<grid, kNT, smem, stream()>>>(adj, x, gamma, beta, mean, rstd, y_out, mask_mode, dx, dresidual, part,
                                          part ? part + P * C : nullptr, R, C, stages, sched_dev, drop, dx_amax);
    } else {
      ln_bwd_kernel<TPR, VPT><<<grid, kNT, 0, stream()>>>(adj, x, gamma, beta, mean, rstd, y_out, mask_mode, dx,
                                                         dresidual, part, part ? part + P * C : nullptr, R, C, drop,
                                                         dx_amax);
    }
  }
  SK_LAUNCH_CHECK();
  if (want_params) {
    dim3 pg((unsigned)((C / 4 + 7) / 8), 2);
    ln_param_grads_kernel<<<pg, kNT, 0, stream()>>>(part, part + P * C, dgamma, dbeta, P, C);
    SK_LAUNCH_CHECK();
    return sk_free(part);
  }
  return SK_OK;
}

#define LN_DISPATCH(FN, ...)                                                          \
  do {                                                                                \
    const int c4 = (int)(cols >> 2);                                                  \
    if (c4 <= 32) return FN<32, 1>(__VA_ARGS__);                                      \
    if (c4 <= 64) return FN<32, 2>(__VA_ARGS__);                                      \
    if (c4 <= 128) return FN<32, 4>(__VA_ARGS__);                                     \
    if (c4 <= 256) return FN<256, 1>(__VA_ARGS__);                                    \
    if (c4 <= 512) return FN<256, 2>(__VA_ARGS__);                                    \
    if (c4 <= 1024) return FN<256, 4>(__VA_ARGS__);                                   \
    if (c4 <= 2048) return FN<256, 8>(__VA_ARGS__);                                   \
  } while (0)

// ------------------------------------------------------------------- BatchNorm1d
// Column statistics with a per-column shift K = x[0, c]:  S1 = sum(x-K), S2 = sum((x-K)^2)
// add across row slabs; mean = K + S1/R, var = S2/R - (S1/R)^2 (biased, forward.pyx:298-301).
__global__ void __launch_bounds__(kNT)
bn_stats_kernel(const float *__restrict__ x, float *__restrict__ part, int64_t R, int64_t C,
                int64_t rows_per_slab) {
  // block = 32 float4 column groups x 8 row lanes (same shape as reduce_cols)
  __shared__ float sm[2][8][129];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = ((int64_t)blockIdx.x * 32 + tx) * 4;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_slab;
  const int64_t r1 = (r0 + rows_per_slab < R) ? r0 + rows_per_slab : R;
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  if (c < C) {
    const float4 k = __ldg(reinterpret_cast<const float4 *>(x + c));
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      float4 v = ld_stream(reinterpret_cast<const float4 *>(x + r * C + c));
      v.x -= k.x; v.y -= k.y; v.z -= k.z; v.w -= k.w;
      s1.x += v.x; s1.y += v.y; s1.z += v.z; s1.w += v.w;
      s2.x += v.x * v.x; s2.y += v.y * v.y; s2.z += v.z * v.z; s2.w += v.w * v.w;
    }
  }
  sm[0][ty][tx * 4 + 0] = s1.x; sm[0][ty][tx * 4 + 1] = s1.y; sm[0][ty][tx * 4 + 2] = s1.z; sm[0][ty][tx * 4 + 3] = s1.w;
  sm[1][ty][tx * 4 + 0] = s2.x; sm[1][ty][tx * 4 + 1] = s2.y; sm[1][ty][tx * 4 + 2] = s2.z; sm[1][ty][tx * 4 + 3] = s2.w;
  __syncthreads();
  const int which = threadIdx.x >> 7, col = threadIdx.x & 127;  // 256 threads: 2 x 128
  float v = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) v += sm[which][j][col];
  const int64_t cc = (int64_t)blockIdx.x * 128 + col;
  if (cc < C) part[((int64_t)blockIdx.y * 2 + which) * C + cc] = v;
}

__global__ void __launch_bounds__(kNT)
bn_finalize_kernel(const float *__restrict__ x, const float *__restrict__ part, int64_t slabs,
                   int64_t R, int64_t C, float eps, float momentum, float *__restrict__ mean_out,
                   float *__restrict__ rstd_out, float *__restrict__ running_mean,
                   float *__restrict__ running_var) {
  const int64_t c = (int64_t)blockIdx.x * kNT + threadIdx.x;
  if (c >= C) return;
  float s1 = 0.f, s2 = 0.f;
  for (int64_t s = 0; s < slabs; ++s) {
    s1 += part[(s * 2 + 0) * C + c];
    s2 += part[(s * 2 + 1) * C + c];
  }
  const float inv_n = 1.0f / (float)R;
  const float d = s1 * inv_n;
  const float mean = x[c] + d;
  float var = s2 * inv_n - d * d;
  var = fmaxf(var, 0.f);
  mean_out[c] = mean;
  rstd_out[c] = 1.0f / sqrtf(var + eps);
  if (running_mean) {  // forward.pyx:308-318: rm*(1-m) + mean*m
    running_mean[c] = __fadd_rn(__fmul_rn(running_mean[c], 1.0f - momentum), __fmul_rn(mean, momentum));
    running_var[c] = __fadd_rn(__fmul_rn(running_var[c], 1.0f - momentum), __fmul_rn(var, momentum));
  }
}

__global__ void __launch_bounds__(kNT)
bn_apply_kernel(const float *__restrict__ x, const float *__restrict__ gamma,
                const float *__restrict__ beta, const float *__restrict__ mean,
                const float *__restrict__ rstd, float *__restrict__ y, int64_t R, int64_t C4,
                int relu) {
  const int64_t total = R * C4;
  const int64_t stride = (int64_t)gridDim.x * kNT;
  for (int64_t i = (int64_t)blockIdx.x * kNT + threadIdx.x; i < total; i += stride) {
    const int64_t c = i % C4;
    float4 v = ld_stream(reinterpret_cast<const float4 *>(x) + i);
    const float4 m = __ldg(reinterpret_cast<const float4 *>(mean) + c);
    const float4 r = __ldg(reinterpret_cast<const float4 *>(rstd) + c);
    const float4 g = gamma ? __ldg(reinterpret_cast<const float4 *>(gamma) + c) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 b = beta ? __ldg(reinterpret_cast<const float4 *>(beta) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 o;
    o.x = affine(v.x - m.x, r.x, g.x, b.x); o.y = affine(v.y - m.y, r.y, g.y, b.y);
    o.z = affine(v.z - m.z, r.z, g.z, b.z); o.w = affine(v.w - m.w, r.w, g.w, b.w);
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    st_stream(reinterpret_cast<float4 *>(y) + i, o);
  }
}

// backward pass 1: per column S1 = sum(adj*xs), S2 = sum(adj), S3 = sum(xs)
__global__ void __launch_bounds__(kNT)
bn_bwd_stats_kernel(const float *__restrict__ adj, const float *__restrict__ x,
                    const float *__restrict__ gamma, const float *__restrict__ beta,
                    const float *__restrict__ mean, const float *__restrict__ rstd,
                    const float *__restrict__ y_out, int mask_mode, float *__restrict__ part,
                    int64_t R, int64_t C, int64_t rows_per_slab) {
  __shared__ float sm[3][8][129];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = ((int64_t)blockIdx.x * 32 + tx) * 4;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_slab;
  const int64_t r1 = (r0 + rows_per_slab < R) ? r0 + rows_per_slab : R;
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1, s3 = s1;
  if (c < C) {
    const float4 m = __ldg(reinterpret_cast<const float4 *>(mean + c));
    const float4 rr = __ldg(reinterpret_cast<const float4 *>(rstd + c));
    const float4 g = gamma ? __ldg(reinterpret_cast<const float4 *>(gamma + c)) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 b = beta ? __ldg(reinterpret_cast<const float4 *>(beta + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      float4 a = ld_stream(reinterpret_cast<const float4 *>(adj + r * C + c));
      float4 v = ld_stream(reinterpret_cast<const float4 *>(x + r * C + c));
      v.x -= m.x; v.y -= m.y; v.z -= m.z; v.w -= m.w;
      if (mask_mode == 1) {
        if (!(affine(v.x, rr.x, g.x, b.x) > 0.f)) a.x = 0.f;
        if (!(affine(v.y, rr.y, g.y, b.y) > 0.f)) a.y = 0.f;
        if (!(affine(v.z, rr.z, g.z, b.z) > 0.f)) a.z = 0.f;
        if (!(affine(v.w, rr.w, g.w, b.w) > 0.f)) a.w = 0.f;
      } else if (mask_mode == 2) {
        float4 yo = ld_stream(reinterpret_cast<const float4 *>(y_out + r * C + c));
        if (!(yo.x > 0.f)) a.x = 0.f;
        if (!(yo.y > 0.f)) a.y = 0.f;
        if (!(yo.z > 0.f)) a.z = 0.f;
        if (!(yo.w > 0.f)) a.w = 0.f;
      }
      s1.x += a.x * v.x; s1.y += a.y * v.y; s1.z += a.z * v.z; s1.w += a.w * v.w;
      s2.x += a.x; s2.y += a.y; s2.z += a.z; s2.w += a.w;
      s3.x += v.x; s3.y += v.y; s3.z += v.z; s3.w += v.w;
    }
  }
  sm[0][ty][tx * 4 + 0] = s1.x; sm[0][ty][tx * 4 + 1] = s1.y; sm[0][ty][tx * 4 + 2] = s1.z; sm[0][ty][tx * 4 + 3] = s1.w;
  sm[1][ty][tx * 4 + 0] = s2.x; sm[1][ty][tx * 4 + 1] = s2.y; sm[1][ty][tx * 4 + 2] = s2.z; sm[1][ty][tx * 4 + 3] = s2.w;
  sm[2][ty][tx * 4 + 0] = s3.x; sm[2][ty][tx * 4 + 1] = s3.y; sm[2][ty][tx * 4 + 2] = s3.z; sm[2][ty][tx * 4 + 3] = s3.w;
  __syncthreads();
  for (int idx = threadIdx.x; idx < 3 * 128; idx += kNT) {
    const int which = idx >> 7, col = idx & 127;
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) v += sm[which][j][col];
    const int64_t cc = (int64_t)blockIdx.x * 128 + col;
    if (cc < C) part[((int64_t)blockIdx.y * 3 + which) * C + cc] = v;
  }
}

// backward pass 2 (per column): coefficients for dX, and dgamma / dbeta
__global__ void __launch_bounds__(kNT)
bn_bwd_finalize_kernel(const float *__restrict__ part, int64_t slabs, int64_t R, int64_t C,
                       const float *__restrict__ gamma, const float *__restrict__ rstd,
                       float *__restrict__ coef, float *__restrict__ dgamma,
                       float *__restrict__ dbeta) {
  const int64_t c = (int64_t)blockIdx.x * kNT + threadIdx.x;
  if (c >= C) return;
  float s1 = 0.f, s2 = 0.f, s3 = 0.f;
  for (int64_t s = 0; s < slabs; ++s) {
    s1 += part[(s * 3 + 0) * C + c];
    s2 += part[(s * 3 + 1) * C + c];
    s3 += part[(s * 3 + 2) * C + c];
  }
  const float g = gamma ? gamma[c] : 1.f;
  const float r = rstd[c];
  const float inv_n = 1.0f / (float)R;
  if (dgamma) dgamma[c] = r * s1;  // sum(norm * adj), norm = xs * r
  if (dbeta) dbeta[c] = s2;
  const float dvar = (g * s1) * (-0.5f * ((r * r) * r));
  const float dmean = (-r) * (g * s2) + dvar * (inv_n * (-2.0f * s3));
  coef[c] = inv_n * dmean;              // c0
  coef[C + c] = g * r;                  // multiplies adj
  coef[2 * C + c] = dvar * (2.0f * inv_n);  // multiplies xs
}

__global__ void __launch_bounds__(kNT)
bn_bwd_apply_kernel(const float *__restrict__ adj, const float *__restrict__ x,
                    const float *__restrict__ gamma, const float *__restrict__ beta,
                    const float *__restrict__ mean, const float *__restrict__ rstd,
                    const float *__restrict__ y_out, int mask_mode, const float *__restrict__ coef,
                    float *__restrict__ dx, int64_t R, int64_t C) {
  const int64_t C4 = C >> 2;
  const int64_t total = R * C4;
  const int64_t stride = (int64_t)gridDim.x * kNT;
  for (int64_t i = (int64_t)blockIdx.x * kNT + threadIdx.x; i < total; i += stride) {
    const int64_t c = i % C4;
    float4 a = ld_stream(reinterpret_cast<const float4 *>(adj) + i);
    float4 v = ld_stream(reinterpret_cast<const float4 *>(x) + i);
    const float4 m = __ldg(reinterpret_cast<const float4 *>(mean) + c);
    v.x -= m.x; v.y -= m.y; v.z -= m.z; v.w -= m.w;
    if (mask_mode == 1) {
      const float4 rr = __ldg(reinterpret_cast<const float4 *>(rstd) + c);
      const float4 g = gamma ? __ldg(reinterpret_cast<const float4 *>(gamma) + c) : make_float4(1.f, 1.f, 1.f, 1.f);
      const float4 b = beta ? __ldg(reinterpret_cast<const float4 *>(beta) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (!(affine(v.x, rr.x, g.x, b.x) > 0.f)) a.x = 0.f;
      if (!(affine(v.y, rr.y, g.y, b.y) > 0.f)) a.y = 0.f;
      if (!(affine(v.z, rr.z, g.z, b.z) > 0.f)) a.z = 0.f;
      if (!(affine(v.w, rr.w, g.w, b.w) > 0.f)) a.w = 0.f;
    } else if (mask_mode == 2) {
      float4 yo = ld_stream(reinterpret_cast<const float4 *>(y_out) + i);
      if (!(yo.x > 0.f)) a.x = 0.f;
      if (!(yo.y > 0.f)) a.y = 0.f;
      if (!(yo.z > 0.f)) a.z = 0.f;
      if (!(yo.w > 0.f)) a.w = 0.f;
    }
    const float4 c0 = __ldg(reinterpret_cast<const float4 *>(coef) + c);
    const float4 c1 = __ldg(reinterpret_cast<const float4 *>(coef + C) + c);
    const float4 c2 = __ldg(reinterpret_cast<const float4 *>(coef + 2 * C) + c);
    float4 o;
    o.x = c0.x + (a.x * c1.x + c2.x * v.x); o.y = c0.y + (a.y * c1.y + c2.y * v.y);
    o.z = c0.z + (a.z * c1.z + c2.z * v.z); o.w = c0.w + (a.w * c1.w + c2.w * v.w);
    st_stream(reinterpret_cast<float4 *>(dx) + i, o);
  }
}

static void bn_slabs(int64_t R, int64_t C, int64_t &slabs, int64_t &rows_per_slab, int64_t &col_tiles) {
  col_tiles = (C + 127) / 128;
  int64_t want_blocks = (int64_t)ctx().num_sms * 8;
  slabs = (want_blocks + col_tiles - 1) / col_tiles;
  int64_t max_slabs = (R + 63) / 64;
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  if (slabs > 4096) slabs = 4096;
  rows_per_slab = (R + slabs - 1) / slabs;
  rows_per_slab = (rows_per_slab + 7) / 8 * 8;
  slabs = (R + rows_per_slab - 1) / rows_per_slab;
}

// ----------------------------------------------------------- softmax cross-entropy
__device__ __forceinline__ int64_t load_label(const void *p, int dt, int64_t i) {
  switch (dt) {
    case SK_BOOL: case SK_U8: return ((const uint8_t *)p)[i];
    case SK_I8: return ((const int8_t *)p)[i];
    case SK_I16: return ((const int16_t *)p)[i];
    case SK_U16: return ((const uint16_t *)p)[i];
    case SK_I32: return ((const int32_t *)p)[i];
    case SK_U32: return ((const uint32_t *)p)[i];
    default: return ((const int64_t *)p)[i];
  }
}

// one warp per row: m = max, s = sum(exp(x-m)), lse = log(s) + m (forward.pyx:224-247),
// row_loss = lse - x[y]; dx = (exp(x-m)/s - onehot) * inv_b (backward.pyx:980-996).
__global__ void __launch_bounds__(kNT)
softmax_ce_kernel(const float *__restrict__ logits, const void *__restrict__ labels, int label_dt,
                  float *__restrict__ row_loss, float *__restrict__ dlogits, int64_t R, int C,
                  float inv_b, unsigned int *__restrict__ dev_err) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (kNT / 32);
  for (int64_t row = (int64_t)blockIdx.x * (kNT / 32) + (threadIdx.x >> 5); row < R; row += warps_total) {
    const float *xr = logits + row * C;
    float m = -INFINITY;
    for (int i = lane; i < C; i += 32) m = fmaxf(m, xr[i]);
    m = warp_max(m);
    float s = 0.f;
    for (int i = lane; i < C; i += 32) s += expf(xr[i] - m);
    s = warp_sum(s);
    int64_t y = load_label(labels, label_dt, row);
    if (y < 0) y += C;
    // eye(C)[labels] raises IndexError on the reference (device.pyx:239); here the row's loss becomes
    // NaN, its gradient keeps no one-hot term and the sticky error word makes the next sync raise
    const bool bad = y < 0 || y >= C;
    if (bad && lane == 0) atomicOr(dev_err, (unsigned int)SK_DEVERR_LABEL_RANGE);
    if (lane == 0 && row_loss) row_loss[row] = bad ? __int_as_float(0x7fc00000) : (logf(s) + m) - xr[y];
    if (dlogits) {
      float *dr = dlogits + row * C;
      for (int i = lane; i < C; i += 32) {
        float p = expf(xr[i] - m) / s;
        dr[i] = (p - (i == y ? 1.f : 0.f)) * inv_b;
      }
    }
  }
}

// ----------------------------------------------------------------- small fused ops
__global__ void __launch_bounds__(kNT)
add_relu_kernel(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ out, int64_t n) {
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * kNT * 4;
  for (int64_t base = (int64_t)blockIdx.x * kNT * 4 + threadIdx.x; base < n4; base += stride) {
    float4 va[4], vb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t i = base + (int64_t)j * kNT;
      if (i < n4) { va[j] = ld_stream(reinterpret_cast<const float4 *>(a) + i); vb[j] = ld_stream(reinterpret_cast<const float4 *>(b) + i); }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t i = base + (int64_t)j * kNT;
      if (i < n4) {
        float4 o;
        o.x = fmaxf(va[j].x + vb[j].x, 0.f); o.y = fmaxf(va[j].y + vb[j].y, 0.f);
        o.z = fmaxf(va[j].z + vb[j].z, 0.f); o.w = fmaxf(va[j].w + vb[j].w, 0.f);
        st_stream(reinterpret_cast<float4 *>(out) + i, o);
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    int64_t i = (n4 << 2) + threadIdx.x;
    out[i] = fmaxf(a[i] + b[i], 0.f);
  }
}

__global__ void __launch_bounds__(kNT)
accumulate_kernel(float *__restrict__ acc, const float *__restrict__ part, int64_t n) {
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * kNT * 4;
  for (int64_t base = (int64_t)blockIdx.x * kNT * 4 + threadIdx.x; base < n4; base += stride) {
    float4 va[4], vb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t i = base + (int64_t)j * kNT;
      if (i < n4) { va[j] = reinterpret_cast<const float4 *>(acc)[i]; vb[j] = ld_stream(reinterpret_cast<const float4 *>(part) + i); }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t i = base + (int64_t)j * kNT;
      if (i < n4) {
        float4 o;
        o.x = va[j].x + vb[j].x; o.y = va[j].y + vb[j].y; o.z = va[j].z + vb[j].z; o.w = va[j].w + vb[j].w;
        reinterpret_cast<float4 *>(acc)[i] = o;
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    int64_t i = (n4 << 2) + threadIdx.x;
    acc[i] += part[i];
  }
}

__global__ void __launch_bounds__(kNT)
dropout_kernel(const float *__restrict__ x, float *__restrict__ out, float *__restrict__ mask,
               int64_t n, float keep, float r_keep, uint64_t seed, const uint64_t *__restrict__ epoch) {
  seed += *epoch * 0xD1B54A32D192ED03ull;   // replay counter of a captured graph (0 in eager mode)
  const int64_t stride = (int64_t)gridDim.x * kNT;
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * kNT + threadIdx.x; i < n4; i += stride) {
    float4 v = ld_stream(reinterpret_cast<const float4 *>(x) + i);
    const float4 m = drop_mask4(4 * i, seed, keep);
    float4 o;  // (x * mask) * (1/keep): prototypes.pyx:758
    o.x = (v.x * m.x) * r_keep; o.y = (v.y * m.y) * r_keep; o.z = (v.z * m.z) * r_keep; o.w = (v.w * m.w) * r_keep;
    st_stream(reinterpret_cast<float4 *>(out) + i, o);
    if (mask) st_stream(reinterpret_cast<float4 *>(mask) + i, m);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    int64_t i = (n4 << 2) + threadIdx.x;
    float m = (float)(hash_u32(i, seed) >> 8) * (1.0f / 16777216.0f) < keep ? 1.f : 0.f;
    out[i] = (x[i] * m) * r_keep;
    if (mask) mask[i] = m;
  }
}

// backward with the mask REGENERATED from the forward's seed (same counter hash): no mask
// array is stored or read -- 8 B/elem instead of the reference's two passes over three
// arrays (prototypes.pyx:746-760 backward: adj * (1/keep) * mask).
__global__ void __launch_bounds__(kNT)
dropout_bwd_kernel(const float *__restrict__ adj, float *__restrict__ out, int64_t n, float keep,
                   float r_keep, uint64_t seed, const uint64_t *__restrict__ epoch) {
  seed += *epoch * 0xD1B54A32D192ED03ull;
  const int64_t stride = (int64_t)gridDim.x * kNT;
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * kNT + threadIdx.x; i < n4; i += stride) {
    const float4 a = ld_stream(reinterpret_cast<const float4 *>(adj) + i);
    const float4 m = drop_mask4(4 * i, seed, keep);
    float4 o;
    o.x = (a.x * r_keep) * m.x; o.y = (a.y * r_keep) * m.y; o.z = (a.z * r_keep) * m.z; o.w = (a.w * r_keep) * m.w;
    st_stream(reinterpret_cast<float4 *>(out) + i, o);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    const float m = (float)(hash_u32(i, seed) >> 8) * (1.0f / 16777216.0f) < keep ? 1.f : 0.f;
    out[i] = (adj[i] * r_keep) * m;
  }
}

// column sum with optional ReLU mask: out[c] = sum_r (y_out[r,c] > 0 ? adj[r,c] : 0)
__global__ void __launch_bounds__(kNT)
colsum_mask_kernel(const float *__restrict__ adj, const float *__restrict__ y_out,
                   float *__restrict__ part, int64_t R, int64_t C, int64_t rows_per_slab) {
  __shared__ float sm[8][129];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = ((int64_t)blockIdx.x * 32 + tx) * 4;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_slab;
  const int64_t r1 = (r0 + rows_per_slab < R) ? r0 + rows_per_slab : R;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < C) {
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      float4 a = ld_stream(reinterpret_cast<const float4 *>(adj + r * C + c));
      float4 y = ld_stream(reinterpret_cast<const float4 *>(y_out + r * C + c));
      s.x += y.x > 0.f ? a.x : 0.f; s.y += y.y > 0.f ? a.y : 0.f;
      s.z += y.z > 0.f ? a.z : 0.f; s.w += y.w > 0.f ? a.w : 0.f;
    }
  }
  sm[ty][tx * 4 + 0] = s.x; sm[ty][tx * 4 + 1] = s.y; sm[ty][tx * 4 + 2] = s.z; sm[ty][tx * 4 + 3] = s.w;
  __syncthreads();
  if (threadIdx.x < 128) {
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) v += sm[j][threadIdx.x];
    const int64_t cc = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (cc < C) part[(int64_t)blockIdx.y * C + cc] = v;
  }
}

static uint64_t g_dropout_seed = 0x0d15ea5e;
static uint64_t g_dropout_calls = 0;
void dropout_reseed(uint64_t seed) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull;            // splitmix64 of the user seed
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  g_dropout_seed = z ^ (z >> 31);
  g_dropout_calls = 0;
}


}  // namespace sk

using namespace sk;

extern "C" {

static int ln_extras_fwd(const sk_ln_extras *in, const float *residual, int64_t cols, LnExtras &ex, const char *who) {
  ex.hi = nullptr; ex.lo = nullptr; ex.scale = nullptr; ex.res_scale = nullptr; ex.dx_amax = nullptr;
  if (!in || !in->split_hi) return SK_OK;
  SK_REQUIRE(in->split_lo && in->split_scale, "%s: split_hi needs split_lo and split_scale", who);
  SK_REQUIRE(!residual || in->residual_scale, "%s: the output split of a residual LayerNorm needs residual_scale "
             "(the bound of |residual|)", who);
  SK_REQUIRE(cols % 8 == 0, "%s: the output split needs cols %% 8 == 0 (got %lld)", who, (long long)cols);
  SK_REQUIRE(al16(in->split_hi) && al16(in->split_lo), "%s: split buffers must be 16-byte aligned", who);
  ex.hi = (__half *)in->split_hi; ex.lo = (__half *)in->split_lo;
  ex.scale = in->split_scale; ex.res_scale = residual ? in->residual_scale : nullptr;
  return SK_OK;
}

int sk_layernorm_fwd_ex(const float *x, const float *gamma, const float *beta, const float *residual,
                        float *y, float *mean, float *rstd, int64_t rows, int64_t cols, float eps,
                        int relu, const sk_ln_extras *extras) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(x && y && mean && rstd, "sk_layernorm_fwd: null pointer");
  SK_REQUIRE(cols > 0 && cols % 4 == 0 && cols <= 8192,
             "sk_layernorm_fwd: cols must be a multiple of 4 and <= 8192 (got %lld)", (long long)cols);
  SK_REQUIRE(al16(x) && al16(y) && (!gamma || al16(gamma)) && (!beta || al16(beta)) &&
                 (!residual || al16(residual)),
             "sk_layernorm_fwd: pointers must be 16-byte aligned");
  LnExtras ex;
  if ((rc = ln_extras_fwd(extras, residual, cols, ex, "sk_layernorm_fwd_ex"))) return rc;
  if (rows == 0) return SK_OK;
  const DropSpec off = {1.f, 1.f, 0ull, nullptr};
  LN_DISPATCH(ln_fwd_launch, x, gamma, beta, residual, y, mean, rstd, rows, (int)cols, eps, relu, off, ex);
  return SK_ERR_UNSUPPORTED;
}

int sk_layernorm_fwd(const float *x, const float *gamma, const float *beta, const float *residual,
                     float *y, float *mean, float *rstd, int64_t rows, int64_t cols, float eps,
                     int relu) {
  return sk_layernorm_fwd_ex(x, gamma, beta, residual, y, mean, rstd, rows, cols, eps, relu, nullptr);
}

int sk_layernorm_dropout_fwd_ex(const float *x, const float *gamma, const float *beta, float *y, float *mean,
                                float *rstd, int64_t rows, int64_t cols, float eps, int relu, float keep,
                                uint64_t *seed_out, const sk_ln_extras *extras) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(x && y && mean && rstd && seed_out, "sk_layernorm_dropout_fwd: null pointer");
  SK_REQUIRE(cols > 0 && cols % 4 == 0 && cols <= 8192,
             "sk_layernorm_dropout_fwd: cols must be a multiple of 4 and <= 8192 (got %lld)", (long long)cols);
  SK_REQUIRE(keep > 0.f && keep < 1.f, "sk_layernorm_dropout_fwd: keep rate must be in (0, 1)");
  SK_REQUIRE(al16(x) && al16(y) && (!gamma || al16(gamma)) && (!beta || al16(beta)),
             "sk_layernorm_dropout_fwd: pointers must be 16-byte aligned");
  LnExtras ex;
  if ((rc = ln_extras_fwd(extras, nullptr, cols, ex, "sk_layernorm_dropout_fwd_ex"))) return rc;
  // one draw per call from the same sequence as sk_dropout_fwd_seeded
  const uint64_t seed = g_dropout_seed + 0x632BE59BD9B4E019ull * (++g_dropout_calls);
  *seed_out = seed;
  if (rows == 0) return SK_OK;
  const DropSpec drop = {keep, (float)(1.0 / (double)keep), seed, rng_epoch_ptr()};
  LN_DISPATCH(ln_fwd_launch, x, gamma, beta, nullptr, y, mean, rstd, rows, (int)cols, eps, relu, drop, ex);
  return SK_ERR_UNSUPPORTED;
}

int sk_layernorm_dropout_fwd(const float *x, const float *gamma, const float *beta, float *y, float *mean,
                             float *rstd, int64_t rows, int64_t cols, float eps, int relu, float keep,
                             uint64_t *seed_out) {
  return sk_layernorm_dropout_fwd_ex(x, gamma, beta, y, mean, rstd, rows, cols, eps, relu, keep, seed_out, nullptr);
}

int sk_layernorm_bwd_ex(const float *adj, const float *x, const float *gamma, const float *beta,
                        const float *mean, const float *rstd, const float *y_out, int mask_mode,
                        float *dx, float *dgamma, float *dbeta, float *dresidual, int64_t rows,
                        int64_t cols, const sk_ln_extras *extras) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(adj && x && mean && rstd && dx, "sk_layernorm_bwd: null pointer");
  SK_REQUIRE(cols > 0 && cols % 4 == 0 && cols <= 8192,
             "sk_layernorm_bwd: cols must be a multiple of 4 and <= 8192 (got %lld)", (long long)cols);
  SK_REQUIRE(mask_mode >= 0 && mask_mode <= 2, "sk_layernorm_bwd: bad mask_mode");
  SK_REQUIRE(mask_mode != 2 || y_out, "sk_layernorm_bwd: mask_mode 2 needs y_out");
  SK_REQUIRE(al16(adj) && al16(x) && al16(dx), "sk_layernorm_bwd: pointers must be 16-byte aligned");
  if (rows == 0) return SK_OK;
  const DropSpec off = {1.f, 1.f, 0ull, nullptr};
  unsigned int *dx_amax = extras ? extras->dx_absmax : nullptr;
  LN_DISPATCH(ln_bwd_launch, adj, x, gamma, beta, mean, rstd, y_out, mask_mode, dx, dresidual, dgamma,
              dbeta, rows, (int)cols, off, dx_amax);
  return SK_ERR_UNSUPPORTED;
}

int sk_layernorm_bwd(const float *adj, const float *x, const float *gamma, const float *beta,
                     const float *mean, const float *rstd, const float *y_out, int mask_mode,
                     float *dx, float *dgamma, float *dbeta, float *dresidual, int64_t rows,
                     int64_t cols) {
  return sk_layernorm_bwd_ex(adj, x, gamma, beta, mean, rstd, y_out, mask_mode, dx, dgamma, dbeta, dresidual, rows,
                             cols, nullptr);
}

int sk_layernorm_dropout_bwd(const float *adj, const float *x, const float *gamma, const float *beta,
                             const float *mean, const float *rstd, int relu, float keep, float r_keep,
                             uint64_t seed, float *dx, float *dgamma, float *dbeta, int64_t rows,
                             int64_t cols) {
  return sk_layernorm_dropout_bwd_ex(adj, x, gamma, beta, mean, rstd, relu, keep, r_keep, seed, dx, dgamma, dbeta,
                                     rows, cols, nullptr);
}

int sk_layernorm_dropout_bwd_ex(const float *adj, const float *x, const float *gamma, const float *beta,
                                const float *mean, const float *rstd, int relu, float keep, float r_keep,
                                uint64_t seed, float *dx, float *dgamma, float *dbeta, int64_t rows,
                                int64_t cols, const sk_ln_extras *extras) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(adj && x && mean && rstd && dx, "sk_layernorm_dropout_bwd: null pointer");
  SK_REQUIRE(cols > 0 && cols % 4 == 0 && cols <= 8192,
             "sk_layernorm_dropout_bwd: cols must be a multiple of 4 and <= 8192 (got %lld)", (long long)cols);
  SK_REQUIRE(keep > 0.f && keep < 1.f, "sk_layernorm_dropout_bwd: keep rate must be in (0, 1)");
  SK_REQUIRE(al16(adj) && al16(x) && al16(dx), "sk_layernorm_dropout_bwd: pointers must be 16-byte aligned");
  if (rows == 0) return SK_OK;
  const DropSpec drop = {keep, r_keep, seed, rng_epoch_ptr()};
  LN_DISPATCH(ln_bwd_launch, adj, x, gamma, beta, mean, rstd, nullptr, relu ? 1 : 0, dx, nullptr, dgamma,
              dbeta, rows, (int)cols, drop, extras ? extras->dx_absmax : nullptr);
  return SK_ERR_UNSUPPORTED;
}

int sk_batchnorm_fwd(const float *x, const float *gamma, const float *beta, float *y, float *mean,
                     float *rstd, float *running_mean, float *running_var, int64_t rows,
                     int64_t cols, float eps, float momentum, int relu) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(x && y && mean && rstd, "sk_batchnorm_fwd: null pointer");
  SK_REQUIRE(cols > 0 && cols % 4 == 0, "sk_batchnorm_fwd: cols must be a multiple of 4");
  SK_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "sk_batchnorm_fwd: running stats come in pairs");
  SK_REQUIRE(al16(x) && al16(y) && al16(mean) && al16(rstd), "sk_batchnorm_fwd: pointers must be 16-byte aligned");
  if (rows == 0) return SK_OK;
  int64_t slabs, rps, col_tiles;
  bn_slabs(rows, cols, slabs, rps, col_tiles);
  float *part = nullptr;
  if ((rc = sk_malloc((size_t)(slabs * 2 * cols) * sizeof(float), (void **)&part))) return rc;
  ProfScope ps(SK_PROF_BN, (double)rows * cols * 12.0);
  bn_stats_kernel<<<dim3((unsigned)col_tiles, (unsigned)slabs), kNT, 0, stream()>>>(x, part, rows, cols, rps);
  note_launch();
  bn_finalize_kernel<<<(unsigned)((cols + kNT - 1) / kNT), kNT, 0, stream()>>>(x, part, slabs, rows, cols, eps, momentum, mean, rstd, running_mean, running_var);
  note_launch();
  int grid = grid_for(rows * (cols / 4), kNT, 8);
  bn_apply_kernel<<<grid, kNT, 0, stream()>>>(x, gamma, beta, mean, rstd, y, rows, cols / 4, relu);
  SK_LAUNCH_CHECK();
  return sk_free(part);
}

int sk_batchnorm_bwd(const float *adj, const float *x, const float *gamma, const float *beta,
                     const float *mean, const float *rstd, const float *y_out, int mask_mode,
                     float *dx, float *dgamma, float *dbeta, int64_t rows, int64_t cols) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(adj && x && mean && rstd && dx, "sk_batchnorm_bwd: null pointer");
  SK_REQUIRE(cols > 0 && cols % 4 == 0, "sk_batchnorm_bwd: cols must be a multiple of 4");
  SK_REQUIRE(mask_mode >= 0 && mask_mode <= 2, "sk_batchnorm_bwd: bad mask_mode");
  SK_REQUIRE(mask_mode != 2 || y_out, "sk_batchnorm_bwd: mask_mode 2 needs y_out");
  if (rows == 0) return SK_OK;
  int64_t slabs, rps, col_tiles;
  bn_slabs(rows, cols, slabs, rps, col_tiles);
  float *part = nullptr;
  if ((rc = sk_malloc((size_t)((slabs * 3 + 3) * cols) * sizeof(float), (void **)&part))) return rc;
  float *coef = part + slabs * 3 * cols;
  ProfScope ps(SK_PROF_BN, (double)rows * cols * (20.0 + (mask_mode == 2 ? 8.0 : 0.0)));
  bn_bwd_stats_kernel<<<dim3((unsigned)col_tiles, (unsigned)slabs), kNT, 0, stream()>>>(adj, x, gamma, beta, mean, rstd, y_out, mask_mode, part, rows, cols, rps);
  note_launch();
  bn_bwd_finalize_kernel<<<(unsigned)((cols + kNT - 1) / kNT), kNT, 0, stream()>>>(part, slabs, rows, cols, gamma, rstd, coef, dgamma, dbeta);
  note_launch();
  int grid = grid_for(rows * (cols / 4), kNT, 8);
  bn_bwd_apply_kernel<<<grid, kNT, 0, stream()>>>(adj, x, gamma, beta, mean, rstd, y_out, mask_mode, coef, dx, rows, cols);
  SK_LAUNCH_CHECK();
  return sk_free(part);
}

int sk_softmax_ce_fwd_bwd(const float *logits, const void *labels, int label_dtype, float *loss,
                          float *dlogits, float *row_loss, int64_t rows, int64_t classes) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(logits && labels, "sk_softmax_ce: null pointer");
  SK_REQUIRE(label_dtype >= SK_BOOL && label_dtype <= SK_U64, "sk_softmax_ce: labels must be integers");
  SK_REQUIRE(rows > 0 && classes > 0, "sk_softmax_ce: empty input");
  float *rl = row_loss;
  bool own = false;
  if (loss && !rl) {
    if ((rc = sk_malloc((size_t)rows * sizeof(float), (void **)&rl))) return rc;
    own = true;
  }
  // 1/B is a C double in the reference (backward.pyx:995) that NumPy applies as a float32
  const float inv_b = (float)(1.0 / (double)rows);
  int grid = grid_for(rows, kNT / 32, 8);
  ProfScope ps(SK_PROF_LOSS, (double)rows * classes * (dlogits ? 8.0 : 4.0));
  softmax_ce_kernel<<<grid, kNT, 0, stream()>>>(logits, labels, label_dtype, rl, dlogits, rows, (int)classes, inv_b,
                                                dev_error_ptr());
  SK_LAUNCH_CHECK();
  if (loss) {
    if ((rc = reduce_rows_f32(SK_RED_MEAN, rl, rows, loss, 1, rows))) return rc;
  }
  if (own) return sk_free(rl);
  return SK_OK;
}

int sk_add_relu(const float *a, const float *b, float *out, int64_t n) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(a && b && out, "sk_add_relu: null pointer");
  SK_REQUIRE(al16(a) && al16(b) && al16(out), "sk_add_relu: pointers must be 16-byte aligned");
  if (n == 0) return SK_OK;
  int grid = grid_for((n + 3) / 4, kNT * 4, 8);
  ProfScope ps(SK_PROF_EWISE, (double)n * 12.0);
  add_relu_kernel<<<grid, kNT, 0, stream()>>>(a, b, out, n);
  SK_LAUNCH_CHECK();
  return SK_OK;
}

int sk_accumulate(float *acc, const float *part, int64_t n) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(acc && part, "sk_accumulate: null pointer");
  SK_REQUIRE(al16(acc) && al16(part), "sk_accumulate: pointers must be 16-byte aligned");
  if (n == 0) return SK_OK;
  int grid = grid_for((n + 3) / 4, kNT * 4, 8);
  ProfScope ps(SK_PROF_EWISE, (double)n * 12.0);
  accumulate_kernel<<<grid, kNT, 0, stream()>>>(acc, part, n);
  SK_LAUNCH_CHECK();
  return SK_OK;
}

int sk_dropout_fwd(const float *x, float *out, float *mask, int64_t n, float keep) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(x && out, "sk_dropout_fwd: null pointer");
  SK_REQUIRE(keep > 0.f && keep <= 1.f, "sk_dropout_fwd: keep rate must be in (0, 1]");
  SK_REQUIRE(al16(x) && al16(out) && (!mask || al16(mask)), "sk_dropout_fwd: pointers must be 16-byte aligned");
  if (n == 0) return SK_OK;
  const float r_keep = (float)(1.0 / (double)keep);
  int grid = grid_for((n + 3) / 4, kNT, 8);
  uint64_t seed = g_dropout_seed + 0x632BE59BD9B4E019ull * (++g_dropout_calls);
  ProfScope ps(SK_PROF_EWISE, (double)n * (mask ? 12.0 : 8.0));
  dropout_kernel<<<grid, kNT, 0, stream()>>>(x, out, mask, n, keep, r_keep, seed, rng_epoch_ptr());
  SK_LAUNCH_CHECK();
  return SK_OK;
}

int sk_dropout_fwd_seeded(const float *x, float *out, int64_t n, float keep, uint64_t *seed_out) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(x && out && seed_out, "sk_dropout_fwd_seeded: null pointer");
  SK_REQUIRE(keep > 0.f && keep <= 1.f, "sk_dropout_fwd_seeded: keep rate must be in (0, 1]");
  SK_REQUIRE(al16(x) && al16(out), "sk_dropout_fwd_seeded: pointers must be 16-byte aligned");
  const uint64_t seed = g_dropout_seed + 0x632BE59BD9B4E019ull * (++g_dropout_calls);
  *seed_out = seed;
  if (n == 0) return SK_OK;
  const float r_keep = (float)(1.0 / (double)keep);
  int grid = grid_for((n + 3) / 4, kNT, 8);
  ProfScope ps(SK_PROF_EWISE, (double)n * 8.0);
  dropout_kernel<<<grid, kNT, 0, stream()>>>(x, out, nullptr, n, keep, r_keep, seed, rng_epoch_ptr());
  SK_LAUNCH_CHECK();
  return SK_OK;
}

int sk_dropout_bwd(const float *adj, float *out, int64_t n, float keep, float r_keep, uint64_t seed) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(adj && out, "sk_dropout_bwd: null pointer");
  SK_REQUIRE(al16(adj) && al16(out), "sk_dropout_bwd: pointers must be 16-byte aligned");
  if (n == 0) return SK_OK;
  int grid = grid_for((n + 3) / 4, kNT, 8);
  ProfScope ps(SK_PROF_EWISE, (double)n * 8.0);
  dropout_bwd_kernel<<<grid, kNT, 0, stream()>>>(adj, out, n, keep, r_keep, seed, rng_epoch_ptr());
  SK_LAUNCH_CHECK();
  return SK_OK;
}

int sk_colsum(const float *adj, const float *y_out, float *out, int64_t rows, int64_t cols) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(adj && out, "sk_colsum: null pointer");
  if (rows == 0 || cols == 0) return SK_OK;
  if (!y_out) return reduce_cols_sum_f32(adj, cols, out, rows, cols);
  SK_REQUIRE(cols % 4 == 0 && al16(adj) && al16(y_out), "sk_colsum: masked form needs cols % 4 == 0 and aligned pointers");
  int64_t slabs, rps, col_tiles;
  bn_slabs(rows, cols, slabs, rps, col_tiles);
  float *part = nullptr;
  if ((rc = sk_malloc((size_t)(slabs * cols) * sizeof(float), (void **)&part))) return rc;
  colsum_mask_kernel<<<dim3((unsigned)col_tiles, (unsigned)slabs), kNT, 0, stream()>>>(adj, y_out, part, rows, cols, rps);
  SK_LAUNCH_CHECK();
  if ((rc = reduce_cols_sum_f32(part, cols, out, slabs, cols))) return rc;
  return sk_free(part);
}

}  // extern "C"
