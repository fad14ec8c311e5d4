// rng.cu -- counter-based RNG (Philox4x32-10) for uniform / normal / Bernoulli fills.
// Replaces cupy.random.uniform/normal/binomial(1,p) as interned by
// soket/backend/device.pyx:64-66 and used by Device._rand/_randn/_randb
// (:204-224; Dropout mask soket/nn/prototypes.pyx:751-757).
// Streams are not bit-compatible with NumPy's MT19937 (nor was CuPy's): parity
// runs initialise on the host and upload (RNG row).
// 4 B/elem write, one 128-bit store per Philox call.
#include "common.cuh"

namespace sk {

static uint64_t g_seed = 0x5eed5eedULL;
static uint64_t g_offset = 0;

__device__ __forceinline__ void philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3,
                                             uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
  uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
  uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
  c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}

__device__ __forceinline__ uint4 philox4x32_10(uint64_t counter, uint64_t seed) {
  uint32_t c0 = (uint32_t)counter, c1 = (uint32_t)(counter >> 32), c2 = 0x9E3779B9u, c3 = 0;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }  // [0,1)

enum { RNG_UNIFORM = 0, RNG_NORMAL = 1, RNG_BERNOULLI = 2 };

template <typename T, int KIND>
__global__ void __launch_bounds__(256)
rng_kernel(T *out, int64_t n, uint64_t seed, uint64_t offset, float p0, float p1) {
  const int64_t n4 = (n + 3) >> 2;
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += stride) {
    uint4 r = philox4x32_10(offset + (uint64_t)i, seed);
    float v[4];
    if (KIND == RNG_UNIFORM) {
      v[0] = p0 + (p1 - p0) * u01(r.x); v[1] = p0 + (p1 - p0) * u01(r.y);
      v[2] = p0 + (p1 - p0) * u01(r.z); v[3] = p0 + (p1 - p0) * u01(r.w);
    } else if (KIND == RNG_NORMAL) {
      // Box-Muller on (0,1] x [0,1)
      float u1 = 1.0f - u01(r.x), u2 = u01(r.y), u3 = 1.0f - u01(r.z), u4 = u01(r.w);
      float ra = sqrtf(-2.0f * logf(u1)), rb = sqrtf(-2.0f * logf(u3));
      float s1, c1, s2, c2;
      sincospif(2.0f * u2, &s1, &c1);
      sincospif(2.0f * u4, &s2, &c2);
      v[0] = p0 + p1 * ra * c1; v[1] = p0 + p1 * ra * s1;
      v[2] = p0 + p1 * rb * c2; v[3] = p0 + p1 * rb * s2;
    } else {
      v[0] = u01(r.x) < p0 ? 1.f : 0.f; v[1] = u01(r.y) < p0 ? 1.f : 0.f;
      v[2] = u01(r.z) < p0 ? 1.f : 0.f; v[3] = u01(r.w) < p0 ? 1.f : 0.f;
    }
    const int64_t base = i << 2;
    if (sizeof(T) == 4 && base + 3 < n) {
      st_stream(reinterpret_cast<float4 *>(out) + i, make_float4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (base + k < n) out[base + k] = (T)v[k];
    }
  }
}

template <int KIND>
static int launch_rng(sk_array *out, float p0, float p1) {
  SK_REQUIRE(out != nullptr, "rng: null array");
  SK_REQUIRE(is_contiguous(out), "rng: output must be contiguous");
  const int64_t n = numel(out);
  if (n == 0) return SK_OK;
  int grid = grid_for((n + 3) / 4, 256, 8);
  switch (out->dtype) {
    case SK_F32: rng_kernel<float, KIND><<<grid, 256, 0, stream()>>>((float *)out->data, n, g_seed, g_offset, p0, p1); break;
    case SK_F64: rng_kernel<double, KIND><<<grid, 256, 0, stream()>>>((double *)out->data, n, g_seed, g_offset, p0, p1); break;
    case SK_I64: rng_kernel<int64_t, KIND><<<grid, 256, 0, stream()>>>((int64_t *)out->data, n, g_seed, g_offset, p0, p1); break;
    case SK_I32: rng_kernel<int32_t, KIND><<<grid, 256, 0, stream()>>>((int32_t *)out->data, n, g_seed, g_offset, p0, p1); break;
    case SK_U8: case SK_BOOL:
      rng_kernel<uint8_t, KIND><<<grid, 256, 0, stream()>>>((uint8_t *)out->data, n, g_seed, g_offset, p0, p1); break;
    default:
      set_error("rng: unsupported output dtype %d", out->dtype);
      return SK_ERR_UNSUPPORTED;
  }
  g_offset += (uint64_t)((n + 3) / 4);
  SK_LAUNCH_CHECK();
  return SK_OK;
}

}  // namespace sk

using namespace sk;

extern "C" {

int sk_rng_seed(uint64_t seed) {
  g_seed = seed;
  g_offset = 0;
  dropout_reseed(seed);   // Dropout masks follow the seed too (per-rank seeds => per-rank masks)
  return SK_OK;
}
int sk_rng_uniform(sk_array *out, double low, double high) {
  int rc;
  if ((rc = ensure_init())) return rc;
  return launch_rng<RNG_UNIFORM>(out, (float)low, (float)high);
}
int sk_rng_normal(sk_array *out, double mean, double std) {
  int rc;
  if ((rc = ensure_init())) return rc;
  return launch_rng<RNG_NORMAL>(out, (float)mean, (float)std);
}
int sk_rng_bernoulli(sk_array *out, double p) {
  int rc;
  if ((rc = ensure_init())) return rc;
  return launch_rng<RNG_BERNOULLI>(out, (float)p, 0.f);
}

}  // extern "C"
