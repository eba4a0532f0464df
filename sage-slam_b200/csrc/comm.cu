// comm.cu -- run-time binding of the six NCCL entry points the batched problem needs (see comm.h).
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

#include "comm.h"

namespace sage
{
namespace
{
struct Nccl
{
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

Nccl &nccl()
{
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    // the copy already mapped into the process first (RTLD_NOLOAD), then the default search path
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names)
      if (!n.lib)
        n.lib = dlopen(nm, RTLD_NOW | RTLD_NOLOAD);
    for (const char *nm : names)
      if (!n.lib)
        n.lib = dlopen(nm, RTLD_NOW);
    if (!n.lib)
    {
      n.err = std::string("libnccl not found: ") + dlerror();
      return;
    }
    auto sym = [&](const char *name) -> void * {
      void *f = dlsym(n.lib, name);
      if (!f && n.err.empty())
        n.err = std::string("libnccl lacks ") + name;
      return f;
    };
    n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(sym("ncclGetUniqueId"));
    n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(sym("ncclCommInitRank"));
    n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(sym("ncclCommDestroy"));
    n.AllGather = reinterpret_cast<decltype(n.AllGather)>(sym("ncclAllGather"));
    n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(sym("ncclAllReduce"));
    n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("ncclGetErrorString"));
  });
  return n;
}

std::string check(ncclResult_t r, const char *what)
{
  if (r == ncclSuccess)
    return "";
  Nccl &n = nccl();
  return std::string(what) + ": " + (n.GetErrorString ? n.GetErrorString(r) : "NCCL error");
}
} // namespace

std::string nccl_unique_id(char out[128])
{
  Nccl &n = nccl();
  if (!n.err.empty())
    return n.err;
  ncclUniqueId id;
  std::string e = check(n.GetUniqueId(&id), "ncclGetUniqueId");
  if (e.empty())
    for (int i = 0; i < 128; ++i)
      out[i] = id.internal[i];
  return e;
}

std::string nccl_comm_create(const char idb[128], int rank, int world, void **comm)
{
  Nccl &n = nccl();
  if (!n.err.empty())
    return n.err;
  ncclUniqueId id;
  for (int i = 0; i < 128; ++i)
    id.internal[i] = idb[i];
  ncclComm_t c = nullptr;
  std::string e = check(n.CommInitRank(&c, world, id, rank), "ncclCommInitRank");
  if (e.empty())
    *comm = c;
  return e;
}

std::string nccl_comm_destroy(void *comm)
{
  Nccl &n = nccl();
  if (!n.err.empty() || !comm)
    return n.err;
  return check(n.CommDestroy(static_cast<ncclComm_t>(comm)), "ncclCommDestroy");
}

std::string nccl_allgather_inplace(void *comm, float *buf, size_t seg, int rank, cudaStream_t s)
{
  Nccl &n = nccl();
  if (!n.err.empty())
    return n.err;
  return check(n.AllGather(buf + (size_t)rank * seg, buf, seg, ncclFloat32, static_cast<ncclComm_t>(comm), s), "ncclAllGather");
}

std::string nccl_allreduce_sum(void *comm, float *buf, size_t count, cudaStream_t s)
{
  Nccl &n = nccl();
  if (!n.err.empty())
    return n.err;
  return check(n.AllReduce(buf, buf, count, ncclFloat32, ncclSum, static_cast<ncclComm_t>(comm), s), "ncclAllReduce");
}

} // namespace sage
