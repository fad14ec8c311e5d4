// dp_p2p.cu -- data-parallel Adam over NVLink peer memory, no collective library on the data path:
// per gradient bucket, copy engines pull this rank's shard of every peer's gradients, ONE local kernel does
// the W-term sum + Adam on the shard + the GEMM weights' fp16x3 operand split, and copy engines push the new
// weights (fp32 + hi / lo) into every replica -- reduce-scatter + optimizer + all-gather with the SMs touching
// only local HBM, and 1/W of the replicated update's optimizer traffic.
//
// New relative to the reference (it has no collective at all).  Every rank maps
// its peers' arenas with CUDA IPC (one cudaMalloc block each, same layout on every rank):
//   G   fp32 gradients (the slots backward writes into)         pulled from all peers, shard only
//   P   fp32 parameters (the tensors' own storage)              pushed to all peers, shard only
//   HI, LO  fp16 hi / lo operand split of the Linear weights    pushed to all peers, shard only
//   F   uint32 flags: ready[bucket][rank], done[bucket][rank], |max| parts[2][tensor][rank]
// A bucket's arena range is cut into `world` contiguous shards of a multiple of 64 elements.  The owner of a
// shard sums the `world` gradient shards in rank order (so the result does not depend on who computes it),
// applies the optimizer's element update (optim.cuh: the reference's arithmetic) and emits, for GEMM weights,
// the hi / lo split under a power-of-two scale every rank derives from the same numbers: max over ranks of
// last step's shard |max| + the optimizer's update bound.  Why copy engines: an SM sustains only a few GB/s
// of remote loads / stores (measured: 96 CTAs of direct peer ld/st moved 370 GB/s and fought the backward
// GEMMs for SMs), a copy engine moves 8 MB shards at 520-560 GB/s beside them (profiles/r2_p2p_copy_bench.txt).
//
// Streams per bucket: S0 (the caller's launch stream: ready kernel, update kernel), one pull stream, two push
// streams, one flag stream; events chain them, S0 never waits for the pushes (bucket k+1's pull overlaps
// bucket k's push).  Flags live in peer-mapped memory, values = step number, monotonic:
//   ready: a 1-block kernel on S0, stream-ordered after the bucket's last gradient kernel, stores `step` into
//          every peer's ready[bucket][me], then spins until its own ready[bucket][*] == step.
//   done:  a 1-block kernel on the flag stream, after this rank's pushes, stores `step` into every peer's
//          done[bucket][me]; sk_dp_p2p_wait (1 block on the compute stream) spins on done[*][*] before the
//          next forward reads the weights.  A peer's `done` also means it has finished PULLING this rank's
//          gradient shard, so the next backward may overwrite the slots.
#include <cuda.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "matmul_split.cuh"
#include "optim.cuh"

namespace sk {

constexpr int kP2pThreads = 256;
constexpr int kP2pMaxTensors = 32;
constexpr int64_t kP2pChunk = kP2pThreads * 4 * 4;   // elements per block-iteration

struct P2pHyper {
  float lr, beta1, beta2, omb1, omb2, eps, wd, grad_scale;
  int have_wd, have_scale, first;
};

struct P2pTensor {
  int64_t off;     // element offset of the tensor in G / P / HI / LO (same on every rank)
  int64_t start;   // this rank's shard: elements [start, start + count) of the tensor
  int64_t count;
  float *m, *v;    // optimizer state of the shard (local, `count` elements)
  float *scale4;   // local float[4] {scale, 1/scale, bound, 0} of the weight's operand split, or null
  int slot;        // row of the |max| parts table
  int first;       // no optimizer state yet (optim.pyx:224-238)
};

struct P2pArgs {
  float *G, *P;                        // this rank's arenas
  __half *HI, *LO;
  const float *S;                      // staging: peer q's copy of this rank's shard at S + q * shard_stride
  int64_t shard_start, shard_stride;   // first arena element of this rank's shard; elements between staged copies
  unsigned int *F[SK_P2P_MAX_WORLD];
  P2pTensor t[kP2pMaxTensors];
  int block_start[kP2pMaxTensors + 1];
  int n, total_blocks, world, rank, bucket;
  unsigned int step;
  int done_off, parts_off, n_slots;    // offsets (in words) into F
  int share_grads;                     // also leave the reduced gradient in G (the caller pushes it to the replicas)
  float bc1, bc2, update_bound;
  P2pHyper h;
  unsigned int *counter;               // local: finished blocks of this launch (zero before and after)
  unsigned int *amax_acc;              // local: per slot accumulator of max |p_new| over the shard (zero before and after)
};

struct P2pReadyArgs {
  unsigned int *F[SK_P2P_MAX_WORLD];
  int world, rank, bucket, ready_off, parts_off, n_slots;
  unsigned int step;
  int n;
  int slot[kP2pMaxTensors];
  float *scale4[kP2pMaxTensors];
  float update_bound;
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_peer(const float *p) {      // L2 only: peer data must never sit in this SM's L1
  float4 r;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float ld_peer1(const float *p) {
  float r;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

// the bucket's gradients are complete here (stream order): tell every peer, wait for every peer, and set the
// scale of each GEMM weight's operand split for THIS update from last step's shard maxima
__global__ void p2p_ready_kernel(const __grid_constant__ P2pReadyArgs a) {
  const int t = threadIdx.x;
  if (t < a.world) {
    __threadfence_system();
    st_release_sys(a.F[t] + a.ready_off + a.bucket * SK_P2P_MAX_WORLD + a.rank, a.step);
    const unsigned int *mine = a.F[a.rank] + a.ready_off + a.bucket * SK_P2P_MAX_WORLD + t;
    while (ld_acquire_sys(mine) < a.step) __nanosleep(200);
  }
  __syncthreads();
  if (t < a.n && a.scale4[t]) {
    // parts[(step & 1)][slot][rank]: written by the owners' kernels of the previous step (or by the host at start)
    const unsigned int *parts = a.F[a.rank] + a.parts_off + ((a.step & 1u) * a.n_slots + a.slot[t]) * SK_P2P_MAX_WORLD;
    unsigned int mx = 0;
    for (int q = 0; q < a.world; ++q) {
      const unsigned int b = ld_acquire_sys(parts + q);
      mx = b > mx ? b : mx;
    }
    const float bound = __fadd_ru(__uint_as_float(mx), a.update_bound);
    float sc, inv;
    pow2_scale(bound, sc, inv);
    float *s4 = a.scale4[t];
    s4[0] = sc; s4[1] = inv; s4[2] = bound; s4[3] = 0.f;
  }
}

__device__ __forceinline__ int p2p_find(const P2pArgs &a, int b) {
  int lo = 0, hi = a.n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (a.block_start[mid] <= b) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kP2pThreads) p2p_adam_kernel(const __grid_constant__ P2pArgs a) {
  const int W = a.world;
  for (int blk = blockIdx.x; blk < a.total_blocks; blk += gridDim.x) {
    const int t = p2p_find(a, blk);
    const P2pTensor &T = a.t[t];
    const int64_t base = (int64_t)(blk - a.block_start[t]) * kP2pChunk;
    const int64_t e0 = T.off + T.start;           // arena element of the shard's first element
    P2pHyper h = a.h;
    h.first = T.first;
    const bool split = T.scale4 != nullptr;
    const float sc = split ? T.scale4[0] : 1.f;   // set by p2p_ready_kernel of this bucket (stream order)
    float amax = 0.f;
    const bool vec = ((T.count | T.start) & 3) == 0;
    if (vec) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t i = base + ((int64_t)j * kP2pThreads + threadIdx.x) * 4;
        if (i < T.count) {
          const int64_t e = e0 + i;
          const int64_t k = e - a.shard_start;            // element of the shard (staging index)
          float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int q = 0; q < W; ++q) {                   // rank order whoever owns the shard
            const float4 x = q == a.rank ? ld_stream(reinterpret_cast<const float4 *>(a.G + e))
                                         : ld_stream(reinterpret_cast<const float4 *>(a.S + q * a.shard_stride + k));
            if (q == 0) g = x;
            else { g.x = __fadd_rn(g.x, x.x); g.y = __fadd_rn(g.y, x.y); g.z = __fadd_rn(g.z, x.z); g.w = __fadd_rn(g.w, x.w); }
          }
          float4 pv = *reinterpret_cast<const float4 *>(a.P + e);
          float4 mv = h.first ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<float4 *>(T.m + i);
          float4 vv = h.first ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<float4 *>(T.v + i);
          if (a.share_grads) *reinterpret_cast<float4 *>(a.G + e) = g;
          adam_one(pv.x, g.x, mv.x, vv.x, h, a.bc1, a.bc2); adam_one(pv.y, g.y, mv.y, vv.y, h, a.bc1, a.bc2);
          adam_one(pv.z, g.z, mv.z, vv.z, h, a.bc1, a.bc2); adam_one(pv.w, g.w, mv.w, vv.w, h, a.bc1, a.bc2);
          amax = fmaxf(fmaxf(amax, fmaxf(fabsf(pv.x), fabsf(pv.y))), fmaxf(fabsf(pv.z), fabsf(pv.w)));
          *reinterpret_cast<float4 *>(T.m + i) = mv;
          *reinterpret_cast<float4 *>(T.v + i) = vv;
          *reinterpret_cast<float4 *>(a.P + e) = pv;
          if (split) {
            uint2 hh, ll;
            split4(pv, sc, hh, ll);
            *reinterpret_cast<uint2 *>(a.HI + e) = hh;
            *reinterpret_cast<uint2 *>(a.LO + e) = ll;
          }
        }
      }
    } else {
      for (int64_t i = base + threadIdx.x; i < T.count && i < base + kP2pChunk; i += kP2pThreads) {
        const int64_t e = e0 + i;
        const int64_t k = e - a.shard_start;
        float g = 0.f;
        for (int q = 0; q < W; ++q) {
          const float x = q == a.rank ? a.G[e] : a.S[q * a.shard_stride + k];
          g = q == 0 ? x : __fadd_rn(g, x);
        }
        float pv = a.P[e];
        float mv = h.first ? 0.f : T.m[i], vv = h.first ? 0.f : T.v[i];
        if (a.share_grads) a.G[e] = g;
        adam_one(pv, g, mv, vv, h, a.bc1, a.bc2);
        amax = fmaxf(amax, fabsf(pv));
        T.m[i] = mv; T.v[i] = vv;
        a.P[e] = pv;
        if (split) {
          const float x = pv * sc;
          const __half hv = __float2half_rn(x);
          a.HI[e] = hv;
          a.LO[e] = __float2half_rn(x - __half2float(hv));
        }
      }
    }
    amax = warp_max(amax);
    if ((threadIdx.x & 31) == 0 && amax > 0.f) {
      const unsigned int bits = __float_as_uint(amax);
      unsigned int *acc = a.amax_acc + T.slot;
      if (bits > *(volatile unsigned int *)acc) atomicMax(acc, bits);
    }
  }
  // the last block publishes the shard maxima to every rank's parts table for the next step
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(a.counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  const unsigned int next = (a.step + 1u) & 1u;
  for (int k = threadIdx.x; k < a.n * W; k += kP2pThreads) {
    const int t = k / W, q = k % W;
    const P2pTensor &T = a.t[t];
    const unsigned int bits = T.count > 0 ? *(volatile unsigned int *)(a.amax_acc + T.slot) : 0u;
    st_release_sys(a.F[q] + a.parts_off + (next * a.n_slots + T.slot) * SK_P2P_MAX_WORLD + a.rank, bits);
  }
  __threadfence_system();
  __syncthreads();
  for (int t = threadIdx.x; t < a.n; t += kP2pThreads) a.amax_acc[a.t[t].slot] = 0u;
  if (threadIdx.x == 0) *a.counter = 0u;
}

// after this rank's pushes of the bucket (stream order): tell every peer
__global__ void p2p_done_kernel(const __grid_constant__ P2pReadyArgs a) {
  if ((int)threadIdx.x < a.world) {
    __threadfence_system();
    st_release_sys(a.F[threadIdx.x] + a.ready_off + a.bucket * SK_P2P_MAX_WORLD + a.rank, a.step);   // ready_off = done_off here
  }
}

struct P2pWaitArgs {
  const unsigned int *flags;      // this rank's F
  int done_off, n_buckets, world;
  unsigned int step;
  unsigned int mask[SK_P2P_MAX_BUCKETS / 32];    // buckets that were launched this step
};

__global__ void p2p_wait_kernel(const __grid_constant__ P2pWaitArgs a) {
  for (int k = threadIdx.x; k < a.n_buckets * a.world; k += blockDim.x) {
    const int b = k / a.world, q = k % a.world;
    if (!((a.mask[b >> 5] >> (b & 31)) & 1u)) continue;
    const unsigned int *f = a.flags + a.done_off + b * SK_P2P_MAX_WORLD + q;
    while (ld_acquire_sys(f) < a.step) __nanosleep(200);
  }
  __threadfence_system();
}

static std::vector<void *> g_ipc_open;

}  // namespace sk

using namespace sk;

extern "C" {

int sk_ipc_export(const void *ptr, char handle[SK_IPC_HANDLE_BYTES], int64_t *offset) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(ptr && handle && offset, "sk_ipc_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == SK_IPC_HANDLE_BYTES, "IPC handle size");
  CUdeviceptr base = 0;
  size_t size = 0;
  // driver entry point at run time: the library does not link libcuda (it must load on a box without a driver)
  typedef CUresult (*RangeFn)(CUdeviceptr *, size_t *, CUdeviceptr);
  static RangeFn range_fn = nullptr;
  if (!range_fn) {
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess || !fp) {
      cudaGetLastError();
      set_error("sk_ipc_export: cuMemGetAddressRange is not available from this driver");
      return SK_ERR_CUDA;
    }
    range_fn = (RangeFn)fp;
  }
  CUresult r = range_fn(&base, &size, (CUdeviceptr)ptr);
  if (r != CUDA_SUCCESS) {
    set_error("sk_ipc_export: cuMemGetAddressRange failed (%d)", (int)r);
    return SK_ERR_CUDA;
  }
  cudaIpcMemHandle_t h;
  SK_CUDA(cudaIpcGetMemHandle(&h, (void *)base));
  memcpy(handle, &h, SK_IPC_HANDLE_BYTES);
  *offset = (int64_t)((CUdeviceptr)ptr - base);
  return SK_OK;
}

int sk_ipc_open(const char handle[SK_IPC_HANDLE_BYTES], int64_t offset, void **ptr) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(handle && ptr && offset >= 0, "sk_ipc_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, SK_IPC_HANDLE_BYTES);
  void *base = nullptr;
  SK_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  g_ipc_open.push_back(base);
  *ptr = (char *)base + offset;
  return SK_OK;
}

int sk_ipc_close_all(void) {
  for (void *b : g_ipc_open) cudaIpcCloseMemHandle(b);
  g_ipc_open.clear();
  return SK_OK;
}

// streams and events of the peer-memory update (created on first use)
struct P2pStreams {
  static constexpr int kRing = 64, kMax = SK_P2P_MAX_WORLD - 1;
  cudaStream_t pull[kMax], push[kMax], flag = nullptr;
  int n_pull = 1, n_push = 2;
  cudaEvent_t ready[kRing], updated[kRing], pulled[kRing][kMax], pushed[kRing][kMax];
  unsigned int next = 0;
  bool ok = false;
};
static P2pStreams g_p2p;

static int p2p_streams_init() {
  if (g_p2p.ok) return SK_OK;
  // SOKET_B200_P2P_PULL_STREAMS / _PUSH_STREAMS: copies to different peers may run on different copy engines
  const char *e;
  if ((e = getenv("SOKET_B200_P2P_PULL_STREAMS"))) g_p2p.n_pull = atoi(e);
  if ((e = getenv("SOKET_B200_P2P_PUSH_STREAMS"))) g_p2p.n_push = atoi(e);
  g_p2p.n_pull = g_p2p.n_pull < 1 ? 1 : g_p2p.n_pull > P2pStreams::kMax ? P2pStreams::kMax : g_p2p.n_pull;
  g_p2p.n_push = g_p2p.n_push < 1 ? 1 : g_p2p.n_push > P2pStreams::kMax ? P2pStreams::kMax : g_p2p.n_push;
  // the transfers must not queue behind compute: highest priority
  int lo = 0, hi = 0;
  SK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  for (int i = 0; i < P2pStreams::kMax; ++i) {
    SK_CUDA(cudaStreamCreateWithPriority(&g_p2p.pull[i], cudaStreamNonBlocking, hi));
    SK_CUDA(cudaStreamCreateWithPriority(&g_p2p.push[i], cudaStreamNonBlocking, hi));
  }
  SK_CUDA(cudaStreamCreateWithPriority(&g_p2p.flag, cudaStreamNonBlocking, hi));
  for (int i = 0; i < P2pStreams::kRing; ++i) {
    SK_CUDA(cudaEventCreateWithFlags(&g_p2p.ready[i], cudaEventDisableTiming));
    SK_CUDA(cudaEventCreateWithFlags(&g_p2p.updated[i], cudaEventDisableTiming));
    for (int j = 0; j < P2pStreams::kMax; ++j) {
      SK_CUDA(cudaEventCreateWithFlags(&g_p2p.pulled[i][j], cudaEventDisableTiming));
      SK_CUDA(cudaEventCreateWithFlags(&g_p2p.pushed[i][j], cudaEventDisableTiming));
    }
  }
  g_p2p.ok = true;
  return SK_OK;
}

int sk_dp_p2p_update(const sk_p2p_peers *peers, int bucket, unsigned int step, int64_t bucket_start, int64_t bucket_len,
                     float *staging, int n_tensors, const sk_p2p_tensor *tensors, const sk_p2p_adam *hyper,
                     unsigned int *scratch) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(peers && tensors && hyper && scratch && staging, "sk_dp_p2p_update: null argument");
  SK_REQUIRE(peers->world >= 2 && peers->world <= SK_P2P_MAX_WORLD && peers->rank >= 0 && peers->rank < peers->world,
             "sk_dp_p2p_update: bad rank %d / world %d", peers->rank, peers->world);
  SK_REQUIRE(bucket >= 0 && bucket < peers->n_buckets && peers->n_buckets <= SK_P2P_MAX_BUCKETS && step >= 1,
             "sk_dp_p2p_update: bad bucket %d / step %u", bucket, step);
  SK_REQUIRE(n_tensors >= 1 && n_tensors <= kP2pMaxTensors, "sk_dp_p2p_update: 1..%d tensors per bucket (got %d)",
             kP2pMaxTensors, n_tensors);
  SK_REQUIRE(bucket_start >= 0 && bucket_len > 0 && bucket_start % 64 == 0, "sk_dp_p2p_update: bad bucket range");
  if ((rc = p2p_streams_init())) return rc;
  const int W = peers->world, R = peers->rank;
  // this rank's shard of the bucket: `world` contiguous pieces of a multiple of 64 elements (the last may be short / empty)
  const int64_t stride = sk_p2p_shard_len(bucket_len, W);
  const int64_t my_start = bucket_start + (int64_t)R * stride;
  int64_t my_len = bucket_start + bucket_len - my_start;
  if (my_len > stride) my_len = stride;
  if (my_len < 0) my_len = 0;
  const int ready_off = 0, done_off = peers->n_buckets * SK_P2P_MAX_WORLD;
  const int parts_off = 2 * peers->n_buckets * SK_P2P_MAX_WORLD;
  P2pReadyArgs r;
  memset(&r, 0, sizeof(r));
  P2pArgs a;
  memset(&a, 0, sizeof(a));
  for (int q = 0; q < W; ++q) {
    SK_REQUIRE(peers->grads[q] && peers->params[q] && peers->flags[q], "sk_dp_p2p_update: peer %d has a null arena", q);
    a.F[q] = peers->flags[q]; r.F[q] = peers->flags[q];
  }
  a.G = peers->grads[R]; a.P = peers->params[R];
  a.HI = (__half *)peers->hi[R]; a.LO = (__half *)peers->lo[R];
  a.S = staging; a.shard_start = my_start; a.shard_stride = stride;
  int blocks = 0;
  bool any_split = false;
  double elems = 0, split_elems = 0;
  for (int i = 0; i < n_tensors; ++i) {
    const sk_p2p_tensor &s = tensors[i];
    SK_REQUIRE(s.offset >= 0 && s.start >= 0 && s.count >= 0 && s.slot >= 0 && s.slot < peers->n_slots,
               "sk_dp_p2p_update: tensor %d has a bad range / slot", i);
    SK_REQUIRE(s.count == 0 || (s.m && s.v), "sk_dp_p2p_update: tensor %d has no optimizer state arrays", i);
    SK_REQUIRE(s.count == 0 || (s.offset + s.start >= my_start && s.offset + s.start + s.count <= my_start + my_len),
               "sk_dp_p2p_update: tensor %d: [%lld, +%lld) is outside this rank's shard of the bucket", i,
               (long long)(s.offset + s.start), (long long)s.count);
    SK_REQUIRE(!s.scale4 || (peers->hi[0] && peers->lo[0]), "sk_dp_p2p_update: tensor %d wants a split but there is no hi / lo arena", i);
    P2pTensor &T = a.t[i];
    T.off = s.offset; T.start = s.start; T.count = s.count; T.m = s.m; T.v = s.v; T.scale4 = s.scale4;
    T.slot = s.slot; T.first = s.first;
    if (((s.count | s.start) & 3) == 0)
      SK_REQUIRE(s.count == 0 || ((((uintptr_t)s.m | (uintptr_t)s.v) & 15) == 0 && (s.offset & 3) == 0),
                 "sk_dp_p2p_update: tensor %d: vector path needs 16-byte aligned state and arena offset", i);
    a.block_start[i] = blocks;
    blocks += (int)((s.count + kP2pChunk - 1) / kP2pChunk);
    elems += (double)s.count;
    if (s.scale4) { split_elems += (double)s.count; any_split = true; }
    r.slot[i] = s.slot; r.scale4[i] = s.scale4;
  }
  a.block_start[n_tensors] = blocks;
  a.n = n_tensors; a.total_blocks = blocks; a.world = W; a.rank = R; a.bucket = bucket; a.step = step;
  a.done_off = done_off; a.parts_off = parts_off; a.n_slots = peers->n_slots;
  a.share_grads = hyper->share_grads;
  a.bc1 = (float)hyper->one_minus_beta1_t; a.bc2 = (float)hyper->one_minus_beta2_t;
  a.update_bound = (float)hyper->update_bound;
  a.h.lr = (float)hyper->lr; a.h.beta1 = (float)hyper->beta1; a.h.beta2 = (float)hyper->beta2;
  a.h.omb1 = (float)(1.0 - hyper->beta1); a.h.omb2 = (float)(1.0 - hyper->beta2);
  a.h.eps = (float)hyper->eps; a.h.wd = (float)hyper->weight_decay; a.h.grad_scale = (float)hyper->grad_scale;
  a.h.have_wd = hyper->weight_decay != 0.0; a.h.have_scale = hyper->grad_scale != 1.0; a.h.first = 0;
  a.counter = scratch + bucket;                       // one counter per bucket: launches of different buckets may overlap
  a.amax_acc = scratch + SK_P2P_MAX_BUCKETS;
  r.world = W; r.rank = R; r.bucket = bucket; r.ready_off = ready_off; r.parts_off = parts_off;
  r.n_slots = peers->n_slots; r.step = step; r.n = n_tensors; r.update_bound = a.update_bound;
  cudaStream_t s0 = stream();
  const unsigned int slot = g_p2p.next++ % P2pStreams::kRing;
  // 1. every rank's gradients of this bucket are complete
  p2p_ready_kernel<<<1, 32, 0, s0>>>(r);
  SK_LAUNCH_CHECK();
  // 2. copy engines pull this rank's shard of every peer's gradients into the staging rows
  if (my_len > 0) {
    SK_CUDA(cudaEventRecord(g_p2p.ready[slot], s0));
    for (int j = 0; j < g_p2p.n_pull; ++j) SK_CUDA(cudaStreamWaitEvent(g_p2p.pull[j], g_p2p.ready[slot], 0));
    for (int k = 1; k < W; ++k) {
      const int q = (R + k) % W;                      // every rank starts with a different peer
      SK_CUDA(cudaMemcpyAsync(staging + (int64_t)q * stride, peers->grads[q] + my_start, (size_t)my_len * 4,
                              cudaMemcpyDefault, g_p2p.pull[(k - 1) % g_p2p.n_pull]));
    }
    for (int j = 0; j < g_p2p.n_pull; ++j) {
      SK_CUDA(cudaEventRecord(g_p2p.pulled[slot][j], g_p2p.pull[j]));
      SK_CUDA(cudaStreamWaitEvent(s0, g_p2p.pulled[slot][j], 0));
    }
  }
  // 3. the local update: sum in rank order, Adam, operand split, shard maxima to every rank's parts table
  {
    static const int grid_env = getenv("SOKET_B200_P2P_GRID") ? atoi(getenv("SOKET_B200_P2P_GRID")) : 0;
    // SOKET_B200_P2P_GRID (default two CTAs per SM): the kernel is HBM-bound on a 1/W shard and runs on a
    // high-priority stream beside backward's GEMMs -- it borrows the SMs briefly (measured at 2 ranks: 74 CTAs
    // stretch it to 0.5 ms per bucket and the step from 19.7 to 21.2 ms)
    int grid = grid_env > 0 ? grid_env : 2 * ctx().num_sms;
    if (grid > blocks) grid = blocks;
    if (grid < 1) grid = 1;                           // a rank without a shard still publishes (zero) maxima
    // per element of the shard: world gradient reads, p / m / v read + write, hi / lo written
    ProfScope ps(SK_PROF_OPTIM, elems * (4.0 * W + 24.0 + (a.share_grads ? 4.0 : 0.0)) + split_elems * 4.0);
    p2p_adam_kernel<<<grid, kP2pThreads, 0, s0>>>(a);
    SK_LAUNCH_CHECK();
  }
  // 4. copy engines push this rank's piece of the new weights into every replica; S0 does not wait.  A GEMM
  //    weight travels as its hi / lo operand split (what forward and backward read); its fp32 master copy too
  //    unless hyper->lazy_master (then sk_dp_p2p_gather refreshes the replicas' fp32 copies on demand)
  SK_CUDA(cudaEventRecord(g_p2p.updated[slot], s0));
  for (int j = 0; j < g_p2p.n_push; ++j) SK_CUDA(cudaStreamWaitEvent(g_p2p.push[j], g_p2p.updated[slot], 0));
  if (my_len > 0) {
    for (int k = 1; k < W; ++k) {
      const int q = (R + k) % W;
      cudaStream_t ps = g_p2p.push[(k - 1) % g_p2p.n_push];
      if (!hyper->lazy_master) {
        SK_CUDA(cudaMemcpyAsync(peers->params[q] + my_start, peers->params[R] + my_start, (size_t)my_len * 4,
                                cudaMemcpyDefault, ps));
      }
      for (int i = 0; i < n_tensors; ++i) {
        const sk_p2p_tensor &t = tensors[i];
        if (t.count == 0) continue;
        const int64_t e = t.offset + t.start;
        if (t.scale4) {
          SK_CUDA(cudaMemcpyAsync((__half *)peers->hi[q] + e, (const __half *)peers->hi[R] + e, (size_t)t.count * 2,
                                  cudaMemcpyDefault, ps));
          SK_CUDA(cudaMemcpyAsync((__half *)peers->lo[q] + e, (const __half *)peers->lo[R] + e, (size_t)t.count * 2,
                                  cudaMemcpyDefault, ps));
        } else if (hyper->lazy_master) {
          SK_CUDA(cudaMemcpyAsync(peers->params[q] + e, peers->params[R] + e, (size_t)t.count * 4, cudaMemcpyDefault, ps));
        }
      }
      if (a.share_grads)
        SK_CUDA(cudaMemcpyAsync(peers->grads[q] + my_start, peers->grads[R] + my_start, (size_t)my_len * 4,
                                cudaMemcpyDefault, ps));
    }
  }
  // 5. done[bucket][me] on every rank once the push streams have drained
  for (int j = 0; j < g_p2p.n_push; ++j) {
    SK_CUDA(cudaEventRecord(g_p2p.pushed[slot][j], g_p2p.push[j]));
    SK_CUDA(cudaStreamWaitEvent(g_p2p.flag, g_p2p.pushed[slot][j], 0));
  }
  P2pReadyArgs d = r;
  d.ready_off = done_off;
  p2p_done_kernel<<<1, 32, 0, g_p2p.flag>>>(d);
  SK_LAUNCH_CHECK();
  return SK_OK;
}

// the fp32 master copy of this rank's piece of a bucket into every replica (lazy_master: before anything reads
// the parameters as fp32 -- a checkpoint, a checksum, p.numpy()); on the current launch stream
int sk_dp_p2p_gather(const sk_p2p_peers *peers, int64_t bucket_start, int64_t bucket_len) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(peers && peers->world >= 2 && peers->world <= SK_P2P_MAX_WORLD && bucket_start >= 0 && bucket_len > 0,
             "sk_dp_p2p_gather: bad argument");
  const int W = peers->world, R = peers->rank;
  const int64_t stride = sk_p2p_shard_len(bucket_len, W);
  const int64_t my_start = bucket_start + (int64_t)R * stride;
  int64_t my_len = bucket_start + bucket_len - my_start;
  if (my_len > stride) my_len = stride;
  for (int k = 1; k < W && my_len > 0; ++k) {
    const int q = (R + k) % W;
    SK_CUDA(cudaMemcpyAsync(peers->params[q] + my_start, peers->params[R] + my_start, (size_t)my_len * 4,
                            cudaMemcpyDefault, stream()));
  }
  return SK_OK;
}

int64_t sk_p2p_shard_len(int64_t bucket_len, int world) {
  if (world < 1 || bucket_len <= 0) return 0;
  const int64_t per = (bucket_len + world - 1) / world;
  return (per + 63) / 64 * 64;
}

// copy-engine probe: `reps` x `n_copies` cudaMemcpyAsync of `bytes` each, copy k on stream k % n_streams,
// dst / src advance by `bytes` per copy; CUDA-event time of the whole batch in *ms
int sk_p2p_copy_probe(void *dst, const void *src, size_t bytes, int n_copies, int n_streams, int reps, float *ms) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(dst && src && ms && bytes > 0 && n_copies >= 1 && n_streams >= 1 && n_streams <= 8 && reps >= 1,
             "sk_p2p_copy_probe: bad argument");
  static cudaStream_t st[8] = {nullptr};
  static cudaEvent_t e0 = nullptr, e1 = nullptr, fin[8] = {nullptr};
  if (!e0) {
    SK_CUDA(cudaEventCreate(&e0));
    SK_CUDA(cudaEventCreate(&e1));
    for (int i = 0; i < 8; ++i) {
      SK_CUDA(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
      SK_CUDA(cudaEventCreateWithFlags(&fin[i], cudaEventDisableTiming));
    }
  }
  SK_CUDA(cudaDeviceSynchronize());
  SK_CUDA(cudaEventRecord(e0, st[0]));
  for (int i = 1; i < n_streams; ++i) SK_CUDA(cudaStreamWaitEvent(st[i], e0, 0));
  for (int r = 0; r < reps; ++r)
    for (int k = 0; k < n_copies; ++k)
      SK_CUDA(cudaMemcpyAsync((char *)dst + (size_t)k * bytes, (const char *)src + (size_t)k * bytes, bytes,
                              cudaMemcpyDefault, st[k % n_streams]));
  for (int i = 1; i < n_streams; ++i) {
    SK_CUDA(cudaEventRecord(fin[i], st[i]));
    SK_CUDA(cudaStreamWaitEvent(st[0], fin[i], 0));
  }
  SK_CUDA(cudaEventRecord(e1, st[0]));
  SK_CUDA(cudaEventSynchronize(e1));
  SK_CUDA(cudaEventElapsedTime(ms, e0, e1));
  return SK_OK;
}

int sk_dp_p2p_wait(const unsigned int *flags, int n_buckets, int world, unsigned int step, const unsigned int *bucket_mask) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(flags && bucket_mask && n_buckets >= 1 && n_buckets <= SK_P2P_MAX_BUCKETS && world >= 2 && world <= SK_P2P_MAX_WORLD,
             "sk_dp_p2p_wait: bad argument");
  P2pWaitArgs a;
  memset(&a, 0, sizeof(a));
  a.flags = flags; a.done_off = n_buckets * SK_P2P_MAX_WORLD; a.n_buckets = n_buckets; a.world = world; a.step = step;
  for (int i = 0; i < (n_buckets + 31) / 32; ++i) a.mask[i] = bucket_mask[i];
  p2p_wait_kernel<<<1, 256, 0, stream()>>>(a);
  SK_LAUNCH_CHECK();
  return SK_OK;
}

}  // extern "C"
