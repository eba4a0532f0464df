// comm.h -- NCCL collectives issued by the library itself on the context stream (no host-language callback in the LM loop).
// libnccl is bound at run time (dlopen): the copy already loaded in the process (e.g. PyTorch's) is preferred, so the
// library has no link-time NCCL dependency and still loads on a machine without it.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <string>

struct sage_ba_comm
{
  void *nccl = nullptr; // ncclComm_t
  int rank = 0, world = 1;
  bool owned = false; // created by sage_ba_comm_create (destroyed with the handle) or borrowed from the caller
};

namespace sage
{
// every function returns an empty string on success, else the error text
std::string nccl_unique_id(char out[128]);
std::string nccl_comm_create(const char id[128], int rank, int world, void **comm);
std::string nccl_comm_destroy(void *comm);
// in place: every rank contributes buf[rank * seg, (rank + 1) * seg) and receives all segments
std::string nccl_allgather_inplace(void *comm, float *buf, size_t seg, int rank, cudaStream_t s);
std::string nccl_allreduce_sum(void *comm, float *buf, size_t count, cudaStream_t s);
} // namespace sage
