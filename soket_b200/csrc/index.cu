// index.cu -- integer-array gather, eye, one-hot.
// Replaces ndarray.__getitem__(int array) / eye as used by Device._one_hot
// (soket/backend/device.pyx:236-239: `eye(C, None, 0, dtype)[labels]`) and the
// row gathers an on-device dataset needs.  Integer / byte
// work: bit-exact by construction (raw element moves).
#include "common.cuh"

namespace sk {

__device__ __forceinline__ int64_t load_index(const void *p, int dt, int64_t i) {
  switch (dt) {
    case SK_BOOL: case SK_U8: return ((const uint8_t *)p)[i];
    case SK_I8: return ((const int8_t *)p)[i];
    case SK_I16: return ((const int16_t *)p)[i];
    case SK_U16: return ((const uint16_t *)p)[i];
    case SK_I32: return ((const int32_t *)p)[i];
    case SK_U32: return ((const uint32_t *)p)[i];
    case SK_I64: return ((const int64_t *)p)[i];
    default: return (int64_t)((const uint64_t *)p)[i];
  }
}

static inline bool aligned16(const void *p) { return (((uintptr_t)p) & 15) == 0; }

struct GatherDesc {
  const void *src;
  const void *index;
  void *out;
  int index_dt;
  int index_ndim;
  int64_t index_shape[SK_MAX_NDIM], index_stride[SK_MAX_NDIM];
  int64_t n_index, inner, n_rows, src_row_stride;
};

// out[i, j] = src[index[i], j] ; the inner (row) part is contiguous on both sides.
template <typename T>
__global__ void __launch_bounds__(256) gather_rows_kernel(const GatherDesc d) {
  const int64_t total = d.n_index * d.inner;
  const int64_t stride = (int64_t)gridDim.x * 256;
  const T *src = (const T *)d.src;
  T *out = (T *)d.out;
  for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += stride) {
    int64_t i = t / d.inner, j = t - i * d.inner;
    int64_t rem = i, off = 0;
#pragma unroll 1
    for (int k = d.index_ndim - 1; k >= 0; --k) {
      int64_t q = rem / d.index_shape[k];
      off += (rem - q * d.index_shape[k]) * d.index_stride[k];
      rem = q;
    }
    int64_t row = load_index(d.index, d.index_dt, off);
    if (row < 0) row += d.n_rows;           // NumPy negative-index wrap
    row = row < 0 ? 0 : (row >= d.n_rows ? d.n_rows - 1 : row);  // never fault on bad labels
    out[t] = src[row * d.src_row_stride + j];
  }
}

template <typename T>
__global__ void __launch_bounds__(256) eye_kernel(T *out, int64_t rows, int64_t cols, int64_t k, T one) {
  const int64_t total = rows * cols;
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += stride) {
    int64_t r = t / cols, c = t - r * cols;
    out[t] = (c - r == k) ? one : (T)0;
  }
}

__global__ void __launch_bounds__(256)
one_hot_f32_kernel(const void *labels, int label_dt, float *out, int64_t rows, int64_t classes) {
  const int64_t total = rows * classes;
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += stride) {
    int64_t r = t / classes, c = t - r * classes;
    int64_t y = load_index(labels, label_dt, r);
    if (y < 0) y += classes;
    out[t] = (c == y) ? 1.f : 0.f;
  }
}

}  // namespace sk

using namespace sk;

extern "C" {

int sk_gather_rows(const sk_array *src, const sk_array *index, sk_array *out) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(src && index && out, "sk_gather_rows: null array");
  SK_REQUIRE(src->ndim >= 1, "sk_gather_rows: src must have at least one axis");
  SK_REQUIRE(!dtype_is_float(index->dtype), "sk_gather_rows: index must be an integer array");
  SK_REQUIRE(src->dtype == out->dtype, "sk_gather_rows: dtype mismatch");
  SK_REQUIRE(is_contiguous(out), "sk_gather_rows: out must be contiguous");
  // inner part of src (axes 1..) must be contiguous
  int64_t inner = 1;
  for (int i = src->ndim - 1; i >= 1; --i) {
    SK_REQUIRE(src->shape[i] == 1 || src->strides[i] == inner,
               "sk_gather_rows: src rows must be contiguous (compact first)");
    inner *= src->shape[i];
  }
  GatherDesc d;
  memset(&d, 0, sizeof(d));
  d.src = src->data; d.index = index->data; d.out = out->data;
  d.index_dt = index->dtype; d.index_ndim = index->ndim;
  d.n_index = 1;
  for (int i = 0; i < index->ndim; ++i) {
    d.index_shape[i] = index->shape[i];
    d.index_stride[i] = index->strides[i];
    d.n_index *= index->shape[i];
  }
  d.inner = inner; d.n_rows = src->shape[0]; d.src_row_stride = src->strides[0];
  SK_REQUIRE(numel(out) == d.n_index * inner, "sk_gather_rows: output size mismatch");
  if (d.n_index * inner == 0) return SK_OK;
  SK_REQUIRE(d.n_rows > 0, "sk_gather_rows: index into an empty axis");
  // rows that are whole 16-byte units on both sides (MNIST: 784 floats = 196 units) move as uint4
  const int64_t row_bytes = inner * dtype_size(src->dtype);
  if (row_bytes % 16 == 0 && (src->strides[0] * dtype_size(src->dtype)) % 16 == 0 &&
      aligned16(src->data) && aligned16(out->data)) {
    d.inner = row_bytes / 16;
    d.src_row_stride = src->strides[0] * dtype_size(src->dtype) / 16;
    gather_rows_kernel<uint4><<<grid_for(d.n_index * d.inner, 256, 8), 256, 0, stream()>>>(d);
    SK_LAUNCH_CHECK();
    return SK_OK;
  }
  int grid = grid_for(d.n_index * inner, 256, 8);
  switch (dtype_size(src->dtype)) {
    case 1: gather_rows_kernel<uint8_t><<<grid, 256, 0, stream()>>>(d); break;
    case 2: gather_rows_kernel<uint16_t><<<grid, 256, 0, stream()>>>(d); break;
    case 4: gather_rows_kernel<uint32_t><<<grid, 256, 0, stream()>>>(d); break;
    default: gather_rows_kernel<uint64_t><<<grid, 256, 0, stream()>>>(d); break;
  }
  SK_LAUNCH_CHECK();
  return SK_OK;
}

int sk_eye(sk_array *out, int64_t k) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(out && out->ndim == 2 && is_contiguous(out), "sk_eye: out must be a contiguous matrix");
  const int64_t rows = out->shape[0], cols = out->shape[1];
  if (rows * cols == 0) return SK_OK;
  int grid = grid_for(rows * cols, 256, 8);
  switch (out->dtype) {
    case SK_F32: eye_kernel<float><<<grid, 256, 0, stream()>>>((float *)out->data, rows, cols, k, 1.f); break;
    case SK_F64: eye_kernel<double><<<grid, 256, 0, stream()>>>((double *)out->data, rows, cols, k, 1.0); break;
    case SK_F16: eye_kernel<__half><<<grid, 256, 0, stream()>>>((__half *)out->data, rows, cols, k, __float2half(1.f)); break;
    case SK_BOOL: case SK_I8: case SK_U8:
      eye_kernel<uint8_t><<<grid, 256, 0, stream()>>>((uint8_t *)out->data, rows, cols, k, 1); break;
    case SK_I16: case SK_U16:
      eye_kernel<uint16_t><<<grid, 256, 0, stream()>>>((uint16_t *)out->data, rows, cols, k, 1); break;
    case SK_I32: case SK_U32:
      eye_kernel<uint32_t><<<grid, 256, 0, stream()>>>((uint32_t *)out->data, rows, cols, k, 1); break;
    case SK_I64: case SK_U64:
      eye_kernel<uint64_t><<<grid, 256, 0, stream()>>>((uint64_t *)out->data, rows, cols, k, 1); break;
    default:
      set_error("sk_eye: unsupported dtype %d", out->dtype);
      return SK_ERR_UNSUPPORTED;
  }
  SK_LAUNCH_CHECK();
  return SK_OK;
}

int sk_one_hot(const sk_array *labels, sk_array *out) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(labels && out, "sk_one_hot: null array");
  SK_REQUIRE(labels->ndim == 1 && (labels->shape[0] <= 1 || labels->strides[0] == 1),
             "sk_one_hot: labels must be a contiguous vector");
  SK_REQUIRE(!dtype_is_float(labels->dtype), "sk_one_hot: labels must be integers");
  SK_REQUIRE(out->ndim == 2 && out->dtype == SK_F32 && is_contiguous(out) &&
                 out->shape[0] == labels->shape[0],
             "sk_one_hot: out must be a contiguous (rows, classes) fp32 matrix");
  const int64_t rows = out->shape[0], classes = out->shape[1];
  if (rows * classes == 0) return SK_OK;
  int grid = grid_for(rows * classes, 256, 8);
  one_hot_f32_kernel<<<grid, 256, 0, stream()>>>(labels->data, labels->dtype, (float *)out->data, rows, classes);
  SK_LAUNCH_CHECK();
  return SK_OK;
}

}  // extern "C"
