// Data-parallel gradient exchange inside the library: NCCL all-reduce of the gradient arena in buckets, on a
// communication stream of its own, fired while the backward pass is still running.
//
// Replaces l3embedding/training_utils.py:141-170 (the reference builds one TF graph over N GPUs and lets autodiff sum
// the replica gradients into shared variables).  Here every GPU has its own process and context; samples are
// independent, so the only exchange of a step is the sum of the gradients (plus two loss scalars).
//
// NCCL is bound at run time (dlopen of libnccl.so.2, the soname both the system library and the one PyTorch bundles
// carry): the library has no link-time dependency on it, a single-GPU host never loads it, and inside a PyTorch
// process the already-loaded NCCL is reused.  Only ABI-stable entry points and constants are used.
#include <dlfcn.h>
#include <string.h>
#include "kernels.h"

namespace l3 {

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;           // NCCL_UNIQUE_ID_BYTES
typedef int ncclResult_t;                                      // ncclSuccess == 0
enum { kNcclSum = 0, kNcclFloat32 = 7, kNcclFloat64 = 8 };     // ncclRedOp_t / ncclDataType_t values (nccl.h)

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};

NcclApi* nccl() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (api.handle) {
#define L3_SYM(field, sym) *(void**)(&api.field) = dlsym(api.handle, sym)
      L3_SYM(GetUniqueId, "ncclGetUniqueId");
      L3_SYM(CommInitRank, "ncclCommInitRank");
      L3_SYM(CommDestroy, "ncclCommDestroy");
      L3_SYM(AllReduce, "ncclAllReduce");
      L3_SYM(GroupStart, "ncclGroupStart");
      L3_SYM(GroupEnd, "ncclGroupEnd");
      L3_SYM(GetErrorString, "ncclGetErrorString");
      L3_SYM(GetVersion, "ncclGetVersion");
#undef L3_SYM
      if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce || !api.GroupStart || !api.GroupEnd) {
        dlclose(api.handle);
        api.handle = nullptr;
      }
    }
  }
  return api.handle ? &api : nullptr;
}

#define L3_CHECK_NCCL(expr)                                                                        \
  do {                                                                                             \
    ncclResult_t _r = (expr);                                                                      \
    if (_r != 0) {                                                                                 \
      NcclApi* _a = nccl();                                                                        \
      set_error("%s:%d %s -> NCCL error %d (%s)", __FILE__, __LINE__, #expr, (int)_r,              \
                (_a && _a->GetErrorString) ? _a->GetErrorString(_r) : "?");                        \
      return -1;                                                                                   \
    }                                                                                              \
  } while (0)
}  // namespace

struct DpState {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  cudaStream_t stream = nullptr;       // communication stream (highest priority: collectives are latency-bound)
  cudaEvent_t ev_ready[kDpMaxBuckets];  // producer stream -> communication stream, per bucket of a step
  cudaEvent_t ev_done;                  // all collectives of the step complete
  int n_issued = 0;
};

int dp_unique_id(char out[128]) {
  NcclApi* a = nccl();
  L3_REQUIRE(a != nullptr, "NCCL (libnccl.so.2) could not be loaded: %s", dlerror() ? dlerror() : "not found");
  ncclUniqueId id;
  L3_CHECK_NCCL(a->GetUniqueId(&id));
  memcpy(out, id.internal, 128);
  return 0;
}

int dp_nccl_version() {
  NcclApi* a = nccl();
  int v = 0;
  if (a && a->GetVersion) a->GetVersion(&v);
  return v;
}

DpState* dp_create(const char id_bytes[128], int rank, int nranks) {
  NcclApi* a = nccl();
  if (!a) {
    set_error("NCCL (libnccl.so.2) could not be loaded");
    return nullptr;
  }
  if (nranks < 2 || rank < 0 || rank >= nranks) {
    set_error("dp_init: rank %d of %d", rank, nranks);
    return nullptr;
  }
  DpState* d = new DpState();
  d->rank = rank;
  d->nranks = nranks;
  ncclUniqueId id;
  memcpy(id.internal, id_bytes, 128);
  ncclResult_t r = a->CommInitRank(&d->comm, nranks, id, rank);
  if (r != 0) {
    set_error("ncclCommInitRank failed: %d (%s)", (int)r, a->GetErrorString ? a->GetErrorString(r) : "?");
    delete d;
    return nullptr;
  }
  int least = 0, greatest = 0;
  bool ok = cudaDeviceGetStreamPriorityRange(&least, &greatest) == cudaSuccess &&
            cudaStreamCreateWithPriority(&d->stream, cudaStreamNonBlocking, greatest) == cudaSuccess &&
            cudaEventCreateWithFlags(&d->ev_done, cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; i < kDpMaxBuckets && ok; ++i)
    ok = cudaEventCreateWithFlags(&d->ev_ready[i], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    set_error("dp_init: stream / events: %s", cudaGetErrorString(cudaGetLastError()));
    a->CommDestroy(d->comm);
    delete d;
    return nullptr;
  }
  return d;
}

void dp_destroy(DpState* d) {
  if (!d) return;
  NcclApi* a = nccl();
  if (d->stream) cudaStreamSynchronize(d->stream);
  if (a && d->comm) a->CommDestroy(d->comm);
  if (d->stream) {
    for (int i = 0; i < kDpMaxBuckets; ++i) cudaEventDestroy(d->ev_ready[i]);
    cudaEventDestroy(d->ev_done);
    cudaStreamDestroy(d->stream);
  }
  delete d;
}

int dp_rank(const DpState* d) { return d ? d->rank : 0; }
int dp_nranks(const DpState* d) { return d ? d->nranks : 1; }

void dp_begin_step(DpState* d) { d->n_issued = 0; }

// In-place sum over the ranks of `n_ranges` disjoint ranges (one grouped launch), ordered after everything enqueued so
// far on `producer`.  fp32 unless is_f64.
int dp_allreduce_ranges(DpState* d, cudaStream_t producer, void* const* ptrs, const long long* counts, int n_ranges,
                        int is_f64) {
  NcclApi* a = nccl();
  L3_REQUIRE(a && d && d->comm, "data parallelism is not initialised (l3_dp_init)");
  L3_REQUIRE(d->n_issued < kDpMaxBuckets, "too many gradient buckets in one step");
  cudaEvent_t ev = d->ev_ready[d->n_issued++];
  L3_CHECK_CUDA(cudaEventRecord(ev, producer));
  L3_CHECK_CUDA(cudaStreamWaitEvent(d->stream, ev, 0));
  if (n_ranges > 1) L3_CHECK_NCCL(a->GroupStart());
  for (int i = 0; i < n_ranges; ++i)
    if (counts[i] > 0)
      L3_CHECK_NCCL(a->AllReduce(ptrs[i], ptrs[i], (size_t)counts[i], is_f64 ? kNcclFloat64 : kNcclFloat32, kNcclSum, d->comm,
                                 d->stream));
  if (n_ranges > 1) L3_CHECK_NCCL(a->GroupEnd());
  return 0;
}

// `consumer` continues only after every collective issued in this step has completed
int dp_join(DpState* d, cudaStream_t consumer) {
  L3_CHECK_CUDA(cudaEventRecord(d->ev_done, d->stream));
  L3_CHECK_CUDA(cudaStreamWaitEvent(consumer, d->ev_done, 0));
  return 0;
}

}  // namespace l3
