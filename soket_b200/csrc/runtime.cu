// runtime.cu -- context, stream, caching allocator, copies, events, graphs.
// Stands in for what CuPy's runtime did for the reference
// (soket/backend/device.pyx:56-58,188-198; soket/tensor/tensor.pyx:384-442).
#include <stdarg.h>
#include <string.h>

#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace sk {

static thread_local char g_err[1024] = "";
static uint64_t g_launches = 0;

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
  const char *base = strrchr(file, '/');
  set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e),
            base ? base + 1 : file, line, what);
  if (e == cudaErrorMemoryAllocation) return SK_ERR_OOM;
  return SK_ERR_CUDA;
}

void note_launch() { ++g_launches; }

Context &ctx() {
  static Context c;
  return c;
}

// ---- caching allocator ------------------------------------------------------
// One compute stream => a freed block may be handed out again immediately
// (stream order protects it).  Blocks are cached by rounded size; nothing is
// returned to the driver until sk_empty_cache() or an OOM retry.
// Arenas: arena 0 is the process-wide cache.  A CUDA-graph capture allocates from a
// private arena (sk_arena_*): the blocks its kernels were captured with stay reserved
// for the graph -- they return to the arena's own free lists, which nobody else draws
// from -- so replays never collide with later eager allocations.
static bool g_capturing = false;
struct Allocator {
  struct Block { size_t size; int arena; };
  std::map<int, std::map<size_t, std::vector<void *>>> free_blocks;   // arena -> size -> blocks
  std::unordered_map<void *, Block> live;
  size_t in_use = 0, reserved = 0, peak = 0;
  int cur_arena = 0, next_arena = 1;

  static size_t round_size(size_t n) {
    if (n == 0) n = 1;
    if (n < (1u << 20)) return (n + 511) & ~size_t(511);
    return (n + (size_t(2) << 20) - 1) & ~((size_t(2) << 20) - 1);
  }
  int alloc(size_t nbytes, void **out) {
    size_t sz = round_size(nbytes);
    auto &fl = free_blocks[cur_arena];
    auto it = fl.find(sz);
    if (it != fl.end() && !it->second.empty()) {
      *out = it->second.back();
      it->second.pop_back();
    } else {
      if (g_capturing) {
        set_error("sk_malloc: %zu bytes not available in the capture arena -- the allocation pattern of the "
                  "captured step differs from its dry runs", sz);
        return SK_ERR_UNSUPPORTED;
      }
      cudaError_t e = cudaMalloc(out, sz);
      if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        release_cached();
        e = cudaMalloc(out, sz);
      }
      if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc", __FILE__, __LINE__);
      reserved += sz;
    }
    live[*out] = Block{sz, cur_arena};
    in_use += sz;
    if (in_use > peak) peak = in_use;
    return SK_OK;
  }
  int release(void *p) {
    auto it = live.find(p);
    if (it == live.end()) {
      set_error("sk_free: pointer %p was not allocated by sk_malloc", p);
      return SK_ERR_ARG;
    }
    const Block b = it->second;
    live.erase(it);
    in_use -= b.size;
    free_blocks[b.arena][b.size].push_back(p);
    return SK_OK;
  }
  void release_arena_cached(int arena) {
    auto ia = free_blocks.find(arena);
    if (ia == free_blocks.end()) return;
    for (auto &kv : ia->second) {
      for (void *p : kv.second) {
        cudaFree(p);
        reserved -= kv.first;
      }
    }
    free_blocks.erase(ia);
  }
  // arena 0 only: graph arenas keep their blocks until sk_arena_destroy
  void release_cached() {
    cudaStreamSynchronize(ctx().stream);
    release_arena_cached(0);
  }
};
static Allocator &allocator() {
  static Allocator a;
  return a;
}

int ensure_init() {
  if (ctx().ready) return SK_OK;
  return sk_init(0);
}

// ---- profiler ---------------------------------------------------------------------
struct ProfRec { cudaEvent_t a, b; double work; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof[SK_PROF_NUM];
static std::vector<cudaEvent_t> g_prof_pool;
static cudaEvent_t g_prof_open[SK_PROF_NUM];

static cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) {
    cudaEvent_t e = g_prof_pool.back();
    g_prof_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
bool prof_on() { return g_prof_on; }
void prof_begin(int family) {
  cudaEvent_t e = prof_event();
  cudaEventRecord(e, ctx().launch);
  g_prof_open[family] = e;
}
void prof_end(int family, double work) {
  cudaEvent_t e = prof_event();
  cudaEventRecord(e, ctx().launch);
  g_prof[family].push_back({g_prof_open[family], e, work});
}

static uint64_t *g_epoch_dev = nullptr;
const uint64_t *rng_epoch_ptr() { return g_epoch_dev; }
__global__ void epoch_advance_kernel(uint64_t *e) { *e += 1; }

cudaStream_t stream_by_id(int id) {
  Context &c = ctx();
  switch (id) {
    case SK_STREAM_COMPUTE: return c.stream;
    case SK_STREAM_COMM: return c.comm_stream;
    case SK_STREAM_COPY: return c.copy_stream;
    case SK_STREAM_OPT: return c.opt_stream;
    default: return nullptr;
  }
}

static volatile unsigned int *g_deverr_host = nullptr;
static unsigned int *g_deverr_dev = nullptr;
unsigned int *dev_error_ptr() { return g_deverr_dev; }
int check_dev_error() {
  if (!g_deverr_host || *g_deverr_host == 0) return SK_OK;
  const unsigned int w = *g_deverr_host;
  *g_deverr_host = 0;
  if (w & SK_DEVERR_LABEL_RANGE)
    set_error("softmax cross-entropy: a target label is outside [-classes, classes) (index out of bounds)");
  else
    set_error("a kernel reported an out-of-range index (code %u)", w);
  return SK_ERR_INDEX;
}

static void *g_flush_buf = nullptr;
static size_t g_flush_bytes = 0;

}  // namespace sk

using namespace sk;

extern "C" {

const char *sk_last_error(void) { return g_err; }
const char *sk_version(void) { return "soket_b200 0.1 (sm_100a)"; }
uint64_t sk_launch_count(void) { return g_launches; }

int sk_device_count(int *count) {
  SK_REQUIRE(count != nullptr, "sk_device_count: null out pointer");
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) {
    cudaGetLastError();
    *count = 0;
  }
  return SK_OK;
}

int sk_init(int device) {
  Context &c = ctx();
  if (c.ready) {
    if (c.device == device) return SK_OK;
    set_error("sk_init: already initialised on device %d (one process per GPU)", c.device);
    return SK_ERR_ARG;
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_error("sk_init: no CUDA device visible (%s) -- soket_b200 has no CPU fallback",
              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return SK_ERR_CUDA;
  }
  SK_REQUIRE(device >= 0 && device < n, "sk_init: device %d out of range [0,%d)", device, n);
  SK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  SK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("sk_init: device %d is sm_%d%d; this library is built for sm_100a only", device,
              prop.major, prop.minor);
    return SK_ERR_UNSUPPORTED;
  }
  c.num_sms = prop.multiProcessorCount;
  c.l2_bytes = (size_t)prop.l2CacheSize;
  SK_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  {
    // the collective and the per-bucket optimizer update run BESIDE backward's persistent GEMMs, which hold
    // every SM for a whole kernel: at each kernel boundary the pending CTAs of a higher-priority stream are
    // placed first, so a bucket's all-reduce / update starts after at most one GEMM instead of queueing behind
    // several.  Measured (strong scaling, 8192 rows over W ranks, profiles/r2_dp_scaling.md): at W = 2 (350 us
    // GEMMs) the time from "gradients complete" to "bucket reduced" drops from 1.2 to 0.35 ms and the step by
    // ~2 % (20.7 / 21.2 vs 21.3 / 21.6 ms); at W = 4 nothing changes (12.72 vs 12.66 ms); at W = 8 (100 us
    // GEMMs) there is no queueing to remove and the earlier NCCL CTAs cost 2.7 % (9.68 vs 9.43 ms).
    // Default: on up to 2 ranks (WORLD_SIZE as the launcher exports it); SOKET_B200_STREAM_PRIORITY = 0 / 1
    // overrides.
    int lo = 0, hi = 0;
    SK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    const char *ws = getenv("WORLD_SIZE");
    bool prio = !ws || atoi(ws) <= 2;
    if (getenv("SOKET_B200_STREAM_PRIORITY")) prio = atoi(getenv("SOKET_B200_STREAM_PRIORITY")) != 0;
    const int p = prio ? hi : 0;
    SK_CUDA(cudaStreamCreateWithPriority(&c.comm_stream, cudaStreamNonBlocking, p));
    SK_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    SK_CUDA(cudaStreamCreateWithPriority(&c.opt_stream, cudaStreamNonBlocking, p));
  }
  c.launch = c.stream;
  SK_CUDA(cudaMalloc((void **)&g_epoch_dev, sizeof(uint64_t)));
  SK_CUDA(cudaMemset(g_epoch_dev, 0, sizeof(uint64_t)));
  {
    void *h = nullptr;
    SK_CUDA(cudaHostAlloc(&h, sizeof(unsigned int), cudaHostAllocMapped));
    g_deverr_host = (volatile unsigned int *)h;
    *g_deverr_host = 0;
    SK_CUDA(cudaHostGetDevicePointer((void **)&g_deverr_dev, h, 0));
  }
  c.device = device;
  c.ready = true;
  return SK_OK;
}

int sk_current_device(int *device) {
  SK_REQUIRE(device != nullptr, "null out pointer");
  *device = ctx().ready ? ctx().device : -1;
  return SK_OK;
}

void *sk_stream(void) { return (void *)ctx().stream; }

int sk_sync(void) {
  if (!ctx().ready) return SK_OK;
  SK_CUDA(cudaStreamSynchronize(ctx().stream));
  SK_CUDA(cudaStreamSynchronize(ctx().comm_stream));
  SK_CUDA(cudaStreamSynchronize(ctx().copy_stream));
  SK_CUDA(cudaStreamSynchronize(ctx().opt_stream));
  return check_dev_error();
}

int sk_launch_stream(int stream_id) {
  int rc;
  if ((rc = ensure_init())) return rc;
  cudaStream_t s = stream_by_id(stream_id);
  SK_REQUIRE(s != nullptr, "sk_launch_stream: unknown stream id %d", stream_id);
  ctx().launch = s;
  return SK_OK;
}
int sk_event_record_on(void *ev, int stream_id) {
  cudaStream_t s = stream_by_id(stream_id);
  SK_REQUIRE(ev && s, "sk_event_record_on: null event or unknown stream id %d", stream_id);
  SK_CUDA(cudaEventRecord((cudaEvent_t)ev, s));
  return SK_OK;
}
int sk_stream_wait_event(int stream_id, void *ev) {
  cudaStream_t s = stream_by_id(stream_id);
  SK_REQUIRE(ev && s, "sk_stream_wait_event: null event or unknown stream id %d", stream_id);
  SK_CUDA(cudaStreamWaitEvent(s, (cudaEvent_t)ev, 0));
  return SK_OK;
}

int sk_malloc(size_t nbytes, void **ptr) {
  SK_REQUIRE(ptr != nullptr, "sk_malloc: null out pointer");
  int rc = ensure_init();
  if (rc) return rc;
  return allocator().alloc(nbytes, ptr);
}

int sk_free(void *ptr) {
  if (ptr == nullptr) return SK_OK;
  return allocator().release(ptr);
}

int sk_empty_cache(void) {
  if (!ctx().ready) return SK_OK;
  allocator().release_cached();
  return SK_OK;
}

int sk_arena_create(int *arena) {
  SK_REQUIRE(arena != nullptr, "sk_arena_create: null out pointer");
  *arena = allocator().next_arena++;
  return SK_OK;
}
int sk_arena_begin(int arena) {
  SK_REQUIRE(arena > 0 && arena < allocator().next_arena, "sk_arena_begin: unknown arena %d", arena);
  SK_REQUIRE(allocator().cur_arena == 0, "sk_arena_begin: arena %d is already active", allocator().cur_arena);
  allocator().cur_arena = arena;
  return SK_OK;
}
int sk_arena_end(void) {
  allocator().cur_arena = 0;
  return SK_OK;
}
int sk_arena_destroy(int arena) {
  SK_REQUIRE(arena > 0, "sk_arena_destroy: arena 0 is the process cache (use sk_empty_cache)");
  Allocator &a = allocator();
  SK_REQUIRE(a.cur_arena != arena, "sk_arena_destroy: arena %d is active", arena);
  if (ctx().ready) SK_CUDA(cudaStreamSynchronize(ctx().stream));
  a.release_arena_cached(arena);
  for (auto &kv : a.live)       // blocks still held by the caller fall back to the process cache when freed
    if (kv.second.arena == arena) kv.second.arena = 0;
  return SK_OK;
}
int sk_rng_epoch_advance(void) {
  int rc = ensure_init();
  if (rc) return rc;
  epoch_advance_kernel<<<1, 1, 0, ctx().stream>>>(g_epoch_dev);
  SK_LAUNCH_CHECK();
  return SK_OK;
}

int sk_mem_stats(size_t *in_use, size_t *reserved, size_t *peak_in_use) {
  Allocator &a = allocator();
  if (in_use) *in_use = a.in_use;
  if (reserved) *reserved = a.reserved;
  if (peak_in_use) *peak_in_use = a.peak;
  return SK_OK;
}

int sk_host_alloc(size_t nbytes, void **ptr) {
  SK_REQUIRE(ptr != nullptr, "sk_host_alloc: null out pointer");
  int rc = ensure_init();
  if (rc) return rc;
  SK_CUDA(cudaMallocHost(ptr, nbytes ? nbytes : 1));
  return SK_OK;
}
int sk_host_free(void *ptr) {
  if (ptr) SK_CUDA(cudaFreeHost(ptr));
  return SK_OK;
}

int sk_h2d(void *dst, const void *src, size_t nbytes) {
  int rc = ensure_init();
  if (rc) return rc;
  if (nbytes == 0) return SK_OK;
  SK_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyHostToDevice, ctx().stream));
  SK_CUDA(cudaStreamSynchronize(ctx().stream));
  return SK_OK;
}
int sk_d2h(void *dst, const void *src, size_t nbytes) {
  int rc = ensure_init();
  if (rc) return rc;
  if (nbytes == 0) return SK_OK;
  SK_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToHost, ctx().stream));
  SK_CUDA(cudaStreamSynchronize(ctx().stream));
  return check_dev_error();
}
int sk_h2d_async(void *dst, const void *src, size_t nbytes) {
  int rc = ensure_init();
  if (rc) return rc;
  if (nbytes == 0) return SK_OK;
  SK_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyHostToDevice, ctx().stream));
  return SK_OK;
}
int sk_d2h_async(void *dst, const void *src, size_t nbytes) {
  int rc = ensure_init();
  if (rc) return rc;
  if (nbytes == 0) return SK_OK;
  SK_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToHost, ctx().stream));
  return SK_OK;
}
// Input prefetch: the copy runs on a dedicated stream, ordered AFTER everything already queued
// on the compute stream (so a staging buffer the previous step read is safe to overwrite) and
// concurrently with whatever is queued next; sk_prefetch_wait() makes the compute stream wait
// for the most recent prefetch.  With two staging buffers the batch of step k+1 crosses PCIe
// while step k computes.
static cudaEvent_t g_ev_pf_compute = nullptr, g_ev_pf_done = nullptr;
int sk_h2d_prefetch(void *dst, const void *src, size_t nbytes) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!g_ev_pf_compute) {
    SK_CUDA(cudaEventCreateWithFlags(&g_ev_pf_compute, cudaEventDisableTiming));
    SK_CUDA(cudaEventCreateWithFlags(&g_ev_pf_done, cudaEventDisableTiming));
  }
  SK_CUDA(cudaEventRecord(g_ev_pf_compute, ctx().stream));
  SK_CUDA(cudaStreamWaitEvent(ctx().copy_stream, g_ev_pf_compute, 0));
  if (nbytes) SK_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyHostToDevice, ctx().copy_stream));
  SK_CUDA(cudaEventRecord(g_ev_pf_done, ctx().copy_stream));
  return SK_OK;
}
int sk_prefetch_wait(void) {
  if (!ctx().ready || !g_ev_pf_done) return SK_OK;
  SK_CUDA(cudaStreamWaitEvent(ctx().stream, g_ev_pf_done, 0));
  return SK_OK;
}

int sk_d2d(void *dst, const void *src, size_t nbytes) {
  int rc = ensure_init();
  if (rc) return rc;
  if (nbytes == 0) return SK_OK;
  SK_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToDevice, ctx().launch));
  return SK_OK;
}
int sk_memset(void *dst, int byte, size_t nbytes) {
  int rc = ensure_init();
  if (rc) return rc;
  if (nbytes == 0) return SK_OK;
  SK_CUDA(cudaMemsetAsync(dst, byte, nbytes, ctx().launch));
  return SK_OK;
}

// ---- events -------------------------------------------------------------------
int sk_event_create(void **ev) {
  SK_REQUIRE(ev != nullptr, "null out pointer");
  int rc = ensure_init();
  if (rc) return rc;
  cudaEvent_t e;
  SK_CUDA(cudaEventCreate(&e));
  *ev = (void *)e;
  return SK_OK;
}
int sk_event_record(void *ev) {
  SK_CUDA(cudaEventRecord((cudaEvent_t)ev, ctx().stream));
  return SK_OK;
}
int sk_event_sync(void *ev) {
  SK_CUDA(cudaEventSynchronize((cudaEvent_t)ev));
  return check_dev_error();
}
int sk_event_elapsed_ms(void *start, void *stop, float *ms) {
  SK_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
  return SK_OK;
}
int sk_event_destroy(void *ev) {
  SK_CUDA(cudaEventDestroy((cudaEvent_t)ev));
  return SK_OK;
}

int sk_flush_l2(void) {
  int rc = ensure_init();
  if (rc) return rc;
  size_t want = ctx().l2_bytes * 2;
  if (want < (size_t(256) << 20)) want = size_t(256) << 20;
  if (g_flush_bytes < want) {
    if (g_flush_buf) cudaFree(g_flush_buf);
    SK_CUDA(cudaMalloc(&g_flush_buf, want));
    g_flush_bytes = want;
  }
  SK_CUDA(cudaMemsetAsync(g_flush_buf, 0, g_flush_bytes, ctx().stream));
  return SK_OK;
}

// ---- profiler ---
int sk_prof_enable(int on) {
  g_prof_on = on != 0;
  return SK_OK;
}
int sk_prof_reset(void) {
  for (int f = 0; f < SK_PROF_NUM; ++f) {
    for (auto &r : g_prof[f]) { g_prof_pool.push_back(r.a); g_prof_pool.push_back(r.b); }
    g_prof[f].clear();
  }
  return SK_OK;
}
int sk_prof_collect(int family, int64_t *launches, double *total_ms, double *total_work) {
  SK_REQUIRE(family >= 0 && family < SK_PROF_NUM, "sk_prof_collect: bad family %d", family);
  if (ctx().ready) SK_CUDA(cudaStreamSynchronize(ctx().stream));
  double ms = 0, work = 0;
  for (auto &r : g_prof[family]) {
    float t = 0;
    SK_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    ms += t;
    work += r.work;
  }
  if (launches) *launches = (int64_t)g_prof[family].size();
  if (total_ms) *total_ms = ms;
  if (total_work) *total_work = work;
  return SK_OK;
}

// ---- graphs -------------------------------------------------------------------
int sk_graph_begin(void) {
  int rc = ensure_init();
  if (rc) return rc;
  SK_REQUIRE(!g_capturing, "sk_graph_begin: a capture is already in progress");
  SK_CUDA(cudaStreamBeginCapture(ctx().stream, cudaStreamCaptureModeRelaxed));
  g_capturing = true;   // sk_malloc may only be served from the active arena's free lists now
  return SK_OK;
}
int sk_graph_capturing(void) { return g_capturing ? 1 : 0; }
int sk_graph_end(void **graph_exec) {
  SK_REQUIRE(graph_exec != nullptr, "null out pointer");
  cudaGraph_t g = nullptr;
  g_capturing = false;
  SK_CUDA(cudaStreamEndCapture(ctx().stream, &g));
  cudaGraphExec_t ge = nullptr;
  cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__);
  *graph_exec = (void *)ge;
  return SK_OK;
}
int sk_graph_launch(void *graph_exec) {
  SK_CUDA(cudaGraphLaunch((cudaGraphExec_t)graph_exec, ctx().stream));
  return SK_OK;
}
int sk_graph_destroy(void *graph_exec) {
  if (graph_exec) SK_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
  return SK_OK;
}

}  // extern "C"
