// Collective entries of the C ABI: the exchange steps of the pretraining step, enqueued on the
// caller's stream through a communicator the caller owns (ncclComm_t passed by handle in the descriptor).
//   mr_all_gather      y_all = all_gather(cat(y, y_extra), 'batch')          pretrain/pretrain_model.py:288-290
//   mr_reduce_scatter  its transpose for dL/dy_all in the backward pass       (jax.value_and_grad of :290)
//   mr_all_reduce      pmean of the bf16 gradients and of the metrics         pretrain/pretrain_model.py:329,333
//
// libnccl is NOT a link-time dependency: the functions are resolved with dlopen("libnccl.so.2") at the first
// collective call, preferring the copy the host process has already loaded (a communicator must be used with the
// library that created it).  On a machine without NCCL every entry returns MR_ERR_NCCL_UNAVAILABLE and the rest of
// the library works unchanged.  The handful of NCCL types below are restated from the public nccl.h (NCCL 2.x ABI:
// ncclUniqueId is 128 opaque bytes passed by value; ncclFloat32 = 7, ncclBfloat16 = 9; ncclSum = 0, ncclAvg = 4).
#include <dlfcn.h>

#include <mutex>

#include "common.cuh"

namespace {

struct NcclUniqueId { char internal[128]; };
using NcclComm = void*;
enum { kNcclFloat32 = 7, kNcclBfloat16 = 9, kNcclSum = 0, kNcclAvg = 4 };

struct NcclApi {
  int (*GetVersion)(int*) = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*CommCount)(NcclComm, int*) = nullptr;
  int (*CommUserRank)(NcclComm, int*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*ReduceScatter)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

NcclApi g_nccl;
std::once_flag g_nccl_once;

template <class F>
bool resolve(void* h, const char* name, F& fn) {
  fn = reinterpret_cast<F>(dlsym(h, name));
  return fn != nullptr;
}

const NcclApi& nccl() {
  std::call_once(g_nccl_once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the host's own copy, if it has one
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) return;
    NcclApi& a = g_nccl;
    a.ok = resolve(h, "ncclGetVersion", a.GetVersion) & resolve(h, "ncclGetUniqueId", a.GetUniqueId) &
           resolve(h, "ncclCommInitRank", a.CommInitRank) & resolve(h, "ncclCommDestroy", a.CommDestroy) &
           resolve(h, "ncclCommCount", a.CommCount) & resolve(h, "ncclCommUserRank", a.CommUserRank) &
           resolve(h, "ncclAllGather", a.AllGather) & resolve(h, "ncclReduceScatter", a.ReduceScatter) &
           resolve(h, "ncclAllReduce", a.AllReduce) & resolve(h, "ncclGetErrorString", a.GetErrorString);
  });
  return g_nccl;
}

int32_t nccl_status(int r) { return r == 0 ? MR_OK : MR_ERR_NCCL_BASE - r; }

// shared argument checks; returns the NCCL datatype through *dt
int32_t check(const MrCollectiveDesc* d, const void* const* in, void* const* out, int* dt, size_t* esize) {
  if (!d || !in || !out || !in[0] || !out[0]) return MR_ERR_NULL_POINTER;
  if (!d->comm) return MR_ERR_NULL_POINTER;
  if (d->count < 0) return MR_ERR_BAD_SHAPE;
  if (d->dtype == MR_DT_BF16) {
    *dt = kNcclBfloat16;
    *esize = 2;
  } else if (d->dtype == MR_DT_F32) {
    *dt = kNcclFloat32;
    *esize = 4;
  } else {
    return MR_ERR_BAD_SHAPE;
  }
  if ((reinterpret_cast<uintptr_t>(in[0]) | reinterpret_cast<uintptr_t>(out[0])) & 15) return MR_ERR_MISALIGNED;
  if (!nccl().ok) return MR_ERR_NCCL_UNAVAILABLE;
  return MR_OK;
}

}  // namespace

namespace mr {
const char* nccl_error_string(int code) {
  const NcclApi& a = nccl();
  return a.ok ? a.GetErrorString(code) : "NCCL error (libnccl not loaded)";
}
}  // namespace mr

extern "C" int32_t mr_nccl_version(void) {
  int v = 0;
  if (!nccl().ok || nccl().GetVersion(&v) != 0) return 0;
  return v;
}

extern "C" int32_t mr_nccl_unique_id(void* id128) {
  if (!id128) return MR_ERR_NULL_POINTER;
  if (!nccl().ok) return MR_ERR_NCCL_UNAVAILABLE;
  return nccl_status(nccl().GetUniqueId(static_cast<NcclUniqueId*>(id128)));
}

extern "C" int32_t mr_nccl_comm_init(void** comm, int32_t world, int32_t rank, const void* id128) {
  if (!comm || !id128) return MR_ERR_NULL_POINTER;
  if (world < 1 || rank < 0 || rank >= world) return MR_ERR_BAD_SHAPE;
  if (!nccl().ok) return MR_ERR_NCCL_UNAVAILABLE;
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  return nccl_status(nccl().CommInitRank(comm, world, id, rank));
}

extern "C" int32_t mr_nccl_comm_destroy(void* comm) {
  if (!comm) return MR_ERR_NULL_POINTER;
  if (!nccl().ok) return MR_ERR_NCCL_UNAVAILABLE;
  return nccl_status(nccl().CommDestroy(comm));
}

extern "C" int32_t mr_nccl_comm_info(void* comm, int32_t* world, int32_t* rank) {
  if (!comm || !world || !rank) return MR_ERR_NULL_POINTER;
  if (!nccl().ok) return MR_ERR_NCCL_UNAVAILABLE;
  int w = 0, r = 0;
  int st = nccl().CommCount(comm, &w);
  if (st == 0) st = nccl().CommUserRank(comm, &r);
  *world = w;
  *rank = r;
  return nccl_status(st);
}

// in[0] send [count]; out[0] recv [world * count], rank-major (= all_gather(y, 'batch').reshape(-1, H))
extern "C" int32_t mr_all_gather(mr_stream_t stream, const MrCollectiveDesc* d, const void* const* in, void* const* out,
                                 void*, size_t) {
  int dt = 0;
  size_t es = 0;
  if (int32_t st = check(d, in, out, &dt, &es)) return st;
  if (d->count == 0) return MR_OK;
  return nccl_status(nccl().AllGather(in[0], out[0], static_cast<size_t>(d->count), dt, d->comm,
                                      static_cast<cudaStream_t>(stream)));
}

// in[0] send [world * count], rank-major; out[0] recv [count] = sum (or mean) over ranks of this rank's block
extern "C" int32_t mr_reduce_scatter(mr_stream_t stream, const MrCollectiveDesc* d, const void* const* in,
                                     void* const* out, void*, size_t) {
  int dt = 0;
  size_t es = 0;
  if (int32_t st = check(d, in, out, &dt, &es)) return st;
  if (d->count == 0) return MR_OK;
  return nccl_status(nccl().ReduceScatter(in[0], out[0], static_cast<size_t>(d->count), dt,
                                          d->average ? kNcclAvg : kNcclSum, d->comm, static_cast<cudaStream_t>(stream)));
}

// in[0] send [count]; out[0] recv [count] (may alias in[0]): sum, or mean when d->average (jax.lax.pmean)
extern "C" int32_t mr_all_reduce(mr_stream_t stream, const MrCollectiveDesc* d, const void* const* in, void* const* out,
                                 void*, size_t) {
  int dt = 0;
  size_t es = 0;
  if (int32_t st = check(d, in, out, &dt, &es)) return st;
  if (d->count == 0) return MR_OK;
  return nccl_status(nccl().AllReduce(in[0], out[0], static_cast<size_t>(d->count), dt,
                                      d->average ? kNcclAvg : kNcclSum, d->comm, static_cast<cudaStream_t>(stream)));
}
