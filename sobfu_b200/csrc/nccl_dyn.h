// nccl_dyn.h -- NCCL entry points resolved at run time (dlopen), so that libsobfu_b200.so has no link-time dependency on
// a particular libnccl: inside a torch process the already-loaded libnccl.so.2 (torch's bundled copy) is found first.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <string>

namespace sb {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t *, ncclConfig_t *) = nullptr;   // optional (NCCL >= 2.18)
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
    std::string err;
};

inline NcclApi &nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    const char *cands[] = {getenv("SOBFU_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *c : cands) {
        if (!c) continue;
        h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { api.err = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?"); return api; }
#define SB_SYM(field, name)                                              \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name));   \
    if (!api.field) { api.err = std::string("missing NCCL symbol ") + name; return api; }
    SB_SYM(GetUniqueId, "ncclGetUniqueId")
    SB_SYM(CommInitRank, "ncclCommInitRank")
    SB_SYM(CommDestroy, "ncclCommDestroy")
    SB_SYM(Send, "ncclSend")
    SB_SYM(Recv, "ncclRecv")
    SB_SYM(GroupStart, "ncclGroupStart")
    SB_SYM(GroupEnd, "ncclGroupEnd")
    SB_SYM(AllReduce, "ncclAllReduce")
    SB_SYM(AllGather, "ncclAllGather")
    SB_SYM(GetErrorString, "ncclGetErrorString")
#undef SB_SYM
    api.CommSplit = reinterpret_cast<decltype(api.CommSplit)>(dlsym(h, "ncclCommSplit"));
    api.ok = true;
    return api;
}

}  // namespace sb
