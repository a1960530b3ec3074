// C-ABI collective hooks for non-PyTorch callers (tp_comm_init / tp_comm_destroy / tp_allreduce_planes).
// The Python host uses torch.distributed for the same exchange (dist.py); a C / C++ / Go caller binds these.
// NCCL is opened at run time (dlopen "libnccl.so.2": the copy already loaded in the process if there is one), so
// libtriplane.so has no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>

#include <mutex>

#include "tp_common.cuh"

namespace tp {

// the handful of NCCL declarations used (stable ABI since NCCL 2.0)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { kNcclSum = 0, kNcclMax = 2, kNcclInt32 = 2, kNcclFloat32 = 7 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

static NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle) api.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle) return;
    auto sym = [&](const char* n) { return dlsym(api.handle, n); };
    api.GetUniqueId = (int (*)(ncclUniqueId*))sym("ncclGetUniqueId");
    api.CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))sym("ncclCommInitRank");
    api.CommDestroy = (int (*)(ncclComm_t))sym("ncclCommDestroy");
    api.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))sym("ncclAllReduce");
    api.GroupStart = (int (*)())sym("ncclGroupStart");
    api.GroupEnd = (int (*)())sym("ncclGroupEnd");
    api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GroupStart && api.GroupEnd;
  });
  return api;
}

static int nccl_fail(int rc, const char* what) {
  NcclApi& n = nccl();
  return fail(1000 + rc, "%s: NCCL error %d (%s)", what, rc, n.GetErrorString ? n.GetErrorString(rc) : "?");
}

struct Comm {
  ncclComm_t comm;
  int world, rank;
};

}  // namespace tp

using namespace tp;

extern "C" int tp_comm_unique_id(void* id_out_128_bytes) {
  if (!id_out_128_bytes) return fail(TP_E_NULL, "tp_comm_unique_id: null argument");
  NcclApi& n = nccl();
  if (!n.ok) return fail(TP_E_NULL, "tp_comm_unique_id: libnccl.so.2 could not be loaded");
  if (int rc = n.GetUniqueId(reinterpret_cast<ncclUniqueId*>(id_out_128_bytes))) return nccl_fail(rc, "ncclGetUniqueId");
  return 0;
}

extern "C" int tp_comm_init(void** comm_out, int32_t world, int32_t rank, const void* unique_id_128_bytes) {
  if (!comm_out || !unique_id_128_bytes) return fail(TP_E_NULL, "tp_comm_init: null argument");
  if (world < 1 || rank < 0 || rank >= world) return fail(TP_E_SHAPE, "tp_comm_init: bad world=%d rank=%d", world, rank);
  NcclApi& n = nccl();
  if (!n.ok) return fail(TP_E_NULL, "tp_comm_init: libnccl.so.2 could not be loaded");
  ncclUniqueId id;
  memcpy(&id, unique_id_128_bytes, sizeof(id));
  Comm* c = new Comm{nullptr, world, rank};
  if (int rc = n.CommInitRank(&c->comm, world, id, rank)) {
    delete c;
    return nccl_fail(rc, "ncclCommInitRank");
  }
  *comm_out = c;
  return 0;
}

extern "C" int tp_comm_destroy(void* comm) {
  if (!comm) return 0;
  Comm* c = reinterpret_cast<Comm*>(comm);
  const int rc = nccl().CommDestroy(c->comm);
  delete c;
  return rc ? nccl_fail(rc, "ncclCommDestroy") : 0;
}

extern "C" int tp_allreduce_planes(void* comm, float* planes, int64_t n_floats, int32_t reduce, int32_t* cell_count,
                                   int64_t n_counts, void* stream) {
  if (!comm) return fail(TP_E_NULL, "tp_allreduce_planes: null communicator");
  if (n_floats < 0 || n_counts < 0 || (n_floats && !planes) || (n_counts && !cell_count))
    return fail(TP_E_NULL, "tp_allreduce_planes: null buffer");
  int op;
  if (reduce == TP_REDUCE_MAX || reduce == TP_REDUCE_MAX_PARTIAL) op = kNcclMax;
  else if (reduce == TP_REDUCE_SUM || reduce == TP_REDUCE_MEAN) op = kNcclSum;
  else return fail(TP_E_ENUM, "tp_allreduce_planes: unknown reduce %d", reduce);
  Comm* c = reinterpret_cast<Comm*>(comm);
  NcclApi& n = nccl();
  cudaStream_t s = (cudaStream_t)stream;
  if (int rc = n.GroupStart()) return nccl_fail(rc, "ncclGroupStart");
  int rc1 = 0, rc2 = 0;
  if (n_floats) rc1 = n.AllReduce(planes, planes, (size_t)n_floats, kNcclFloat32, op, c->comm, s);
  if (n_counts) rc2 = n.AllReduce(cell_count, cell_count, (size_t)n_counts, kNcclInt32, kNcclSum, c->comm, s);
  if (int rc = n.GroupEnd()) return nccl_fail(rc, "ncclGroupEnd");
  if (rc1) return nccl_fail(rc1, "ncclAllReduce(planes)");
  if (rc2) return nccl_fail(rc2, "ncclAllReduce(counts)");
  return 0;
}
