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
struct Allocator {
  std::map<size_t, std::vector<void *>> free_blocks;
  std::unordered_map<void *, size_t> live;
  size_t in_use = 0, reserved = 0, peak = 0;

  static size_t round_size(size_t n) {
    if (n == 0) n = 1;
    if (n < (1u << 20)) return (n + 511) & ~size_t(511);
    return (n + (size_t(2) << 20) - 1) & ~((size_t(2) << 20) - 1);
  }
  int alloc(size_t nbytes, void **out) {
    size_t sz = round_size(nbytes);
    auto it = free_blocks.find(sz);
    if (it != free_blocks.end() && !it->second.empty()) {
      *out = it->second.back();
      it->second.pop_back();
    } else {
      cudaError_t e = cudaMalloc(out, sz);
      if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        release_cached();
        e = cudaMalloc(out, sz);
      }
      if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc", __FILE__, __LINE__);
      reserved += sz;
    }
    live[*out] = sz;
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
    size_t sz = it->second;
    live.erase(it);
    in_use -= sz;
    free_blocks[sz].push_back(p);
    return SK_OK;
  }
  void release_cached() {
    cudaStreamSynchronize(ctx().stream);
    for (auto &kv : free_blocks) {
      for (void *p : kv.second) {
        cudaFree(p);
        reserved -= kv.first;
      }
      kv.second.clear();
    }
    free_blocks.clear();
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
  cudaEventRecord(e, ctx().stream);
  g_prof_open[family] = e;
}
void prof_end(int family, double work) {
  cudaEvent_t e = prof_event();
  cudaEventRecord(e, ctx().stream);
  g_prof[family].push_back({g_prof_open[family], e, work});
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
  SK_CUDA(cudaStreamCreateWithFlags(&c.comm_stream, cudaStreamNonBlocking));
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
  return SK_OK;
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
int sk_d2d(void *dst, const void *src, size_t nbytes) {
  int rc = ensure_init();
  if (rc) return rc;
  if (nbytes == 0) return SK_OK;
  SK_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToDevice, ctx().stream));
  return SK_OK;
}
int sk_memset(void *dst, int byte, size_t nbytes) {
  int rc = ensure_init();
  if (rc) return rc;
  if (nbytes == 0) return SK_OK;
  SK_CUDA(cudaMemsetAsync(dst, byte, nbytes, ctx().stream));
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
  return SK_OK;
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
  SK_CUDA(cudaStreamBeginCapture(ctx().stream, cudaStreamCaptureModeThreadLocal));
  return SK_OK;
}
int sk_graph_end(void **graph_exec) {
  SK_REQUIRE(graph_exec != nullptr, "null out pointer");
  cudaGraph_t g = nullptr;
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
