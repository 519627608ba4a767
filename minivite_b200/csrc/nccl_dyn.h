// Runtime binding to libnccl.so.2 (dlopen): the library must not carry a link-time dependency on a
// particular NCCL build, because inside a torch process the torch-bundled NCCL is already loaded and a
// second copy would clash; dlopen by soname resolves to whatever the process already has.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <string>

namespace mvnccl {

struct Api {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

inline bool load(Api &a, std::string &err) {
  if (a.handle) return true;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (a.handle) break;
  }
  if (!a.handle) { err = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return false; }
#define MV_SYM(field, name)                                                     \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, name));         \
  if (!a.field) { err = std::string("libnccl lacks ") + name; return false; }
  MV_SYM(GetUniqueId, "ncclGetUniqueId")
  MV_SYM(CommInitRank, "ncclCommInitRank")
  MV_SYM(CommDestroy, "ncclCommDestroy")
  MV_SYM(AllReduce, "ncclAllReduce")
  MV_SYM(AllGather, "ncclAllGather")
  MV_SYM(Send, "ncclSend")
  MV_SYM(Recv, "ncclRecv")
  MV_SYM(GroupStart, "ncclGroupStart")
  MV_SYM(GroupEnd, "ncclGroupEnd")
  MV_SYM(GetErrorString, "ncclGetErrorString")
#undef MV_SYM
  return true;
}

}  // namespace mvnccl
