// Run-time binding to NCCL for the one collective on the path: the allgather of per-candidate
// log-likelihoods after a multi-device grid fit (README.md:202,285 `pmap` gather -> ncclAllGather over
// NVLink).  NCCL is dlopen()ed so that the library neither links against nor requires it
// for single-GPU use, and so that it can coexist with the NCCL that PyTorch bundles.
#include "gpcc_internal.h"
#include <dlfcn.h>
#include <cstdlib>
#include <cstring>

namespace gpcc {

typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;
enum { ncclFloat64 = 8 };

struct ncclUniqueIdT { char internal[128]; };   // == ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128), passed by value

struct NcclBridge {
    void* lib = nullptr;
    std::vector<int> devs;
    std::vector<ncclComm_t> comms;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueIdT*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueIdT, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

// The library handle is process-wide and never closed: ncclGetUniqueId starts the bootstrap root thread inside libnccl, so
// unloading the library between gpcc_comm_unique_id and gpcc_ctx_comm_init_rank would take that thread away and every rank
// would wait for it forever.
static void* nccl_handle(std::string& err) {
    static void* lib = nullptr;
    if (lib) return lib;
    const char* names[] = {std::getenv("GPCC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n) continue;
        lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (lib) break;
    }
    if (!lib) err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
    return lib;
}

static NcclBridge* nccl_bridge_load(std::string& err) {
    NcclBridge* b = new NcclBridge();
    b->lib = nccl_handle(err);
    if (!b->lib) { delete b; return nullptr; }
#define LOAD(field, sym)                                                          \
    b->field = reinterpret_cast<decltype(b->field)>(dlsym(b->lib, sym));          \
    if (!b->field) { err = std::string("missing symbol ") + sym; delete b; return nullptr; }
    LOAD(CommInitAll, "ncclCommInitAll")
    LOAD(GetUniqueId, "ncclGetUniqueId")
    LOAD(CommInitRank, "ncclCommInitRank")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(AllGather, "ncclAllGather")
    LOAD(GroupStart, "ncclGroupStart")
    LOAD(GroupEnd, "ncclGroupEnd")
    LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
    return b;
}

// one process, several devices (gpcc_ctx_create(ndev > 1)): ncclCommInitAll
NcclBridge* nccl_bridge_create(const std::vector<int>& devs, std::string& err) {
    NcclBridge* b = nccl_bridge_load(err);
    if (!b) return nullptr;
    b->devs = devs;
    b->comms.resize(devs.size());
    ncclResult_t r = b->CommInitAll(b->comms.data(), (int)devs.size(), devs.data());
    if (r != 0) { err = std::string("ncclCommInitAll: ") + b->GetErrorString(r); delete b; return nullptr; }
    return b;
}

// one process per device (torchrun / one Julia worker per GPU): rank 0 draws an id, the host program hands the 128 bytes to
// every rank, every rank joins with its one device
int nccl_bridge_unique_id(char* out128, std::string& err) {
    NcclBridge* b = nccl_bridge_load(err);
    if (!b) return 1;
    ncclUniqueIdT id;
    ncclResult_t r = b->GetUniqueId(&id);
    if (r != 0) err = std::string("ncclGetUniqueId: ") + b->GetErrorString(r);
    else std::memcpy(out128, id.internal, 128);
    delete b;
    return r != 0;
}

NcclBridge* nccl_bridge_create_rank(int dev, int world, int rank, const char* id128, std::string& err) {
    NcclBridge* b = nccl_bridge_load(err);
    if (!b) return nullptr;
    b->devs = {dev};
    b->comms.resize(1);
    ncclUniqueIdT id;
    std::memcpy(id.internal, id128, 128);
    cudaSetDevice(dev);
    ncclResult_t r = b->CommInitRank(&b->comms[0], world, id, rank);
    if (r != 0) { err = std::string("ncclCommInitRank: ") + b->GetErrorString(r); delete b; return nullptr; }
    return b;
}

void nccl_bridge_destroy(NcclBridge* b) {
    if (!b) return;
    for (auto c : b->comms) if (c) b->CommDestroy(c);
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
