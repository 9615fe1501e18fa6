// Run-time binding to NCCL for the one collective on the path: the allgather of per-candidate
// log-likelihoods after a multi-device grid fit (README.md:202,285 `pmap` gather -> ncclAllGather over
// NVLink).  NCCL is dlopen()ed so that the library neither links against nor requires it
// for single-GPU use, and so that it can coexist with the NCCL that PyTorch bundles.
#include "gpcc_internal.h"
#include <dlfcn.h>
#include <cstdlib>

namespace gpcc {

typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;
enum { ncclFloat64 = 8 };

struct NcclBridge {
    void* lib = nullptr;
    std::vector<int> devs;
    std::vector<ncclComm_t> comms;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclBridge* nccl_bridge_create(const std::vector<int>& devs, std::string& err) {
    NcclBridge* b = new NcclBridge();
    const char* names[] = {std::getenv("GPCC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n) continue;
        b->lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (b->lib) break;
    }
    if (!b->lib) { err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); delete b; return nullptr; }
#define LOAD(field, sym)                                                          \
    b->field = reinterpret_cast<decltype(b->field)>(dlsym(b->lib, sym));          \
    if (!b->field) { err = std::string("missing symbol ") + sym; dlclose(b->lib); delete b; return nullptr; }
    LOAD(CommInitAll, "ncclCommInitAll")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(AllGather, "ncclAllGather")
    LOAD(GroupStart, "ncclGroupStart")
    LOAD(GroupEnd, "ncclGroupEnd")
    LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
    b->devs = devs;
    b->comms.resize(devs.size());
    ncclResult_t r = b->CommInitAll(b->comms.data(), (int)devs.size(), devs.data());
    if (r != 0) { err = std::string("ncclCommInitAll: ") + b->GetErrorString(r); dlclose(b->lib); delete b; return nullptr; }
    return b;
}

void nccl_bridge_destroy(NcclBridge* b) {
    if (!b) return;
    for (auto c : b->comms) if (c) b->CommDestroy(c);
    if (b->lib) dlclose(b->lib);
    delete b;
}

int nccl_bridge_allgather(NcclBridge* b, const std::vector<double*>& send, const std::vector<double*>& recv, int count,
                          const std::vector<cudaStream_t>& streams, std::string& err) {
    ncclResult_t r = b->GroupStart();
    if (r) { err = b->GetErrorString(r); return 1; }
    for (size_t i = 0; i < b->comms.size(); ++i) {
        cudaSetDevice(b->devs[i]);
        r = b->AllGather(send[i], recv[i], (size_t)count, ncclFloat64, b->comms[i], streams[i]);
        if (r) { err = b->GetErrorString(r); b->GroupEnd(); return 1; }
    }
    r = b->GroupEnd();
    if (r) { err = b->GetErrorString(r); return 1; }
    for (size_t i = 0; i < b->comms.size(); ++i) {
        cudaSetDevice(b->devs[i]);
        cudaError_t e = cudaStreamSynchronize(streams[i]);
        if (e != cudaSuccess) { err = cudaGetErrorString(e); return 1; }
    }
    return 0;
}

}  // namespace gpcc
