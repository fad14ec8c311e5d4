// dp.cu -- data-parallel gradient all-reduce over NCCL (NVLink 5 / NVSwitch).
//
// New relative to the reference (it has no collective: `grep -ri nccl` -> 0 hits,
// SURVEY.md section 2).  One process per GPU; the only collective is
// ncclAllReduce(sum, fp32) on gradient buckets (section 8e), issued on a dedicated
// comm stream so it overlaps the rest of backward, with 1/W folded into the
// optimizer kernel (grad_scale).  NCCL is bound at run time with dlopen so the
// library loads (and every other entry point works) on a box without it.
#include <dlfcn.h>
#include <stdlib.h>

#include "common.cuh"

namespace sk {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[SK_NCCL_ID_BYTES]; } ncclUniqueId;
enum { ncclFloat32 = 7, ncclSum = 0 };

struct NcclApi {
  void *handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*CommAbort)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool tried = false;
};
static NcclApi g_nccl;
static ncclComm_t g_comm = nullptr;
static int g_world = 1, g_rank = 0;
static cudaEvent_t g_ev_compute = nullptr, g_ev_comm = nullptr;

static bool load_nccl() {
  if (g_nccl.tried) return g_nccl.handle != nullptr;
  g_nccl.tried = true;
  const char *env = getenv("SOKET_B200_NCCL_LIB");
  const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    if (!n || !*n) continue;
    g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  if (!g_nccl.handle) return false;
#define LOAD(field, sym)                                                    \
  *(void **)(&g_nccl.field) = dlsym(g_nccl.handle, sym);                    \
  if (!g_nccl.field) { dlclose(g_nccl.handle); g_nccl.handle = nullptr; return false; }
  LOAD(GetUniqueId, "ncclGetUniqueId")
  LOAD(CommInitRank, "ncclCommInitRank")
  LOAD(AllReduce, "ncclAllReduce")
  LOAD(Broadcast, "ncclBroadcast")
  LOAD(CommDestroy, "ncclCommDestroy")
  LOAD(CommAbort, "ncclCommAbort")
  LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
  return true;
}

// Fail fast: a rank that hits an NCCL error aborts its communicator, which makes the peers'
// pending collectives return an error instead of waiting for this rank forever.
static int nccl_fail(int code, const char *what) {
  set_error("NCCL error %d (%s) in %s; communicator aborted", code,
            g_nccl.GetErrorString ? g_nccl.GetErrorString(code) : "?", what);
  if (g_comm && g_nccl.CommAbort) {
    ncclComm_t c = g_comm;
    g_comm = nullptr;
    g_nccl.CommAbort(c);
  }
  return SK_ERR_NCCL;
}
#define SK_NCCL(expr)                                  \
  do {                                                 \
    int _r = (expr);                                   \
    if (_r != 0) return nccl_fail(_r, #expr);          \
  } while (0)

}  // namespace sk

using namespace sk;

extern "C" {

int sk_nccl_available(void) { return load_nccl() ? 1 : 0; }

int sk_nccl_unique_id(char id[SK_NCCL_ID_BYTES]) {
  SK_REQUIRE(id != nullptr, "sk_nccl_unique_id: null buffer");
  if (!load_nccl()) {
    set_error("sk_nccl_unique_id: libnccl.so.2 not found (set SOKET_B200_NCCL_LIB)");
    return SK_ERR_NCCL;
  }
  ncclUniqueId uid;
  SK_NCCL(g_nccl.GetUniqueId(&uid));
  memcpy(id, uid.internal, SK_NCCL_ID_BYTES);
  return SK_OK;
}

int sk_nccl_init(int rank, int world, const char id[SK_NCCL_ID_BYTES]) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(world >= 1 && rank >= 0 && rank < world, "sk_nccl_init: bad rank %d / world %d", rank, world);
  SK_REQUIRE(g_comm == nullptr, "sk_nccl_init: communicator already initialised");
  if (!load_nccl()) {
    set_error("sk_nccl_init: libnccl.so.2 not found (set SOKET_B200_NCCL_LIB)");
    return SK_ERR_NCCL;
  }
  ncclUniqueId uid;
  memcpy(uid.internal, id, SK_NCCL_ID_BYTES);
  SK_NCCL(g_nccl.CommInitRank(&g_comm, world, uid, rank));
  g_world = world;
  g_rank = rank;
  SK_CUDA(cudaEventCreateWithFlags(&g_ev_compute, cudaEventDisableTiming));
  SK_CUDA(cudaEventCreateWithFlags(&g_ev_comm, cudaEventDisableTiming));
  return SK_OK;
}

// on_comm_stream != 0: the comm stream first waits for everything already queued
// on the compute stream (the bucket's producer kernels), then runs the
// all-reduce concurrently with later compute; sk_nccl_wait() joins it back.
int sk_nccl_allreduce(float *buf, size_t count, int on_comm_stream) {
  SK_REQUIRE(g_comm != nullptr, "sk_nccl_allreduce: call sk_nccl_init first");
  if (count == 0) return SK_OK;
  cudaStream_t s = ctx().stream;
  if (on_comm_stream) {
    SK_CUDA(cudaEventRecord(g_ev_compute, ctx().stream));
    SK_CUDA(cudaStreamWaitEvent(ctx().comm_stream, g_ev_compute, 0));
    s = ctx().comm_stream;
  }
  SK_NCCL(g_nccl.AllReduce(buf, buf, count, ncclFloat32, ncclSum, g_comm, s));
  note_launch();
  return SK_OK;
}

int sk_nccl_allreduce_on(float *buf, size_t count, int stream_id) {
  SK_REQUIRE(g_comm != nullptr, "sk_nccl_allreduce_on: call sk_nccl_init first");
  cudaStream_t s = stream_by_id(stream_id);
  SK_REQUIRE(s != nullptr, "sk_nccl_allreduce_on: unknown stream id %d", stream_id);
  if (count == 0) return SK_OK;
  SK_NCCL(g_nccl.AllReduce(buf, buf, count, ncclFloat32, ncclSum, g_comm, s));
  note_launch();
  return SK_OK;
}

int sk_nccl_abort(void) {
  if (g_comm && g_nccl.CommAbort) {
    ncclComm_t c = g_comm;
    g_comm = nullptr;
    g_nccl.CommAbort(c);
  }
  return SK_OK;
}

int sk_nccl_broadcast(float *buf, size_t count, int root) {
  SK_REQUIRE(g_comm != nullptr, "sk_nccl_broadcast: call sk_nccl_init first");
  if (count == 0) return SK_OK;
  SK_NCCL(g_nccl.Broadcast(buf, buf, count, ncclFloat32, root, g_comm, ctx().stream));
  note_launch();
  return SK_OK;
}

int sk_nccl_wait(void) {
  if (g_comm == nullptr) return SK_OK;
  SK_CUDA(cudaEventRecord(g_ev_comm, ctx().comm_stream));
  SK_CUDA(cudaStreamWaitEvent(ctx().stream, g_ev_comm, 0));
  return SK_OK;
}

int sk_nccl_destroy(void) {
  if (g_comm) {
    cudaStreamSynchronize(ctx().comm_stream);
    cudaStreamSynchronize(ctx().opt_stream);
    cudaStreamSynchronize(ctx().stream);
    g_nccl.CommDestroy(g_comm);
    g_comm = nullptr;
  }
  return SK_OK;
}

}  // extern "C"
