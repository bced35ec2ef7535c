// mpcb_nccl.cu -- cross-rank reconciliation of a split tree: the ranks own contiguous ranges of the first
// control, each reduces its share to one 16-byte (float64 cost, int64 index) record per solve, and ONE collective --
// an ncclAllGather of those records over NVLink -- hands every rank all of them; the lexicographic minimum is then a
// single local kernel.  Exact for float64 costs and 63-bit indices (a packed 64-bit word cannot hold both).  The only
// coupling between the leaves of a tree is the running minimum of math_model.py:195-198; this is its multi-GPU form.
// NCCL is bound at run time with dlopen so that the library loads on hosts without NCCL and shares the copy PyTorch
// already mapped (same SONAME).
#include "../../include/mpcb200.h"
#include "mpcb_handle.cuh"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <climits>
#include <cstring>

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int *) = nullptr;
    ncclResult_t (*CommUserRank)(const ncclComm_t, int *) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    bool ok = false;
};

NcclApi &api() {
    static NcclApi a;
    if (a.lib) return a;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
        a.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (a.lib) break;
    }
    if (!a.lib) return a;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.lib, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.lib, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.lib, "ncclCommDestroy");
    a.CommCount = (decltype(a.CommCount))dlsym(a.lib, "ncclCommCount");
    a.CommUserRank = (decltype(a.CommUserRank))dlsym(a.lib, "ncclCommUserRank");
    a.AllGather = (decltype(a.AllGather))dlsym(a.lib, "ncclAllGather");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.CommCount && a.CommUserRank && a.AllGather;
    return a;
}

struct Rec { double cost; long long index; };

__global__ void pack_kernel(const double *cost, const long long *index, Rec *rec) {
    rec->cost = *cost;
    rec->index = *index;
}

// lexicographic (cost, index) minimum over the ranks that hold a leaf (index >= 0, cost not NaN); if no rank does,
// index -1 and the smallest cost reported (NaN if every cost is NaN)
__global__ void pick_kernel(const Rec *all, int nranks, double *cost, long long *index) {
    double bJ = NAN, lowest = NAN;
    long long bj = -1;
    for (int r = 0; r < nranks; ++r) {
        const double c = all[r].cost;
        const long long j = all[r].index;
        if (c == c && !(lowest <= c)) lowest = c;
        if (j >= 0 && c == c && (bj < 0 || c < bJ || (c == bJ && j < bj))) { bJ = c; bj = j; }
    }
    *cost = bj >= 0 ? bJ : lowest;
    *index = bj;
}

}  // namespace

namespace mpcb {

// internal: used by the split-tree solve of mpcb_api.cu
bool nccl_available() { return api().ok; }

int nccl_comm_info(void *comm, int *nranks, int *rank) {
    if (!comm || !api().ok) return MPCB_ERR_NCCL;
    if (api().CommCount((ncclComm_t)comm, nranks) != ncclSuccess) return MPCB_ERR_NCCL;
    if (api().CommUserRank((ncclComm_t)comm, rank) != ncclSuccess) return MPCB_ERR_NCCL;
    return MPCB_OK;
}

int nccl_allgather_bytes(void *comm, const void *send, void *recv, size_t bytes, cudaStream_t st) {
    if (!comm || !api().ok) return MPCB_ERR_NCCL;
    return api().AllGather(send, recv, bytes, ncclInt8, (ncclComm_t)comm, st) == ncclSuccess ? MPCB_OK : MPCB_ERR_NCCL;
}

}  // namespace mpcb

extern "C" {

int mpcb_nccl_unique_id(void *id128) {
    if (!id128 || !api().ok) return MPCB_ERR_NCCL;
    ncclUniqueId id;
    if (api().GetUniqueId(&id) != ncclSuccess) return MPCB_ERR_NCCL;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    memcpy(id128, &id, 128);
    return MPCB_OK;
}

int mpcb_nccl_comm_create(mpcb_handle *h, int nranks, int rank, const void *id128, void **comm_out) {
    if (!h || !id128 || !comm_out || !api().ok) return MPCB_ERR_NCCL;
    if (cudaSetDevice(h->device) != cudaSuccess) return MPCB_ERR_CUDA;      // the communicator binds to the handle's device
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm;
    if (api().CommInitRank(&comm, nranks, id, rank) != ncclSuccess) return MPCB_ERR_NCCL;
    *comm_out = (void *)comm;
    return MPCB_OK;
}

int mpcb_nccl_comm_destroy(void *comm) {
    if (!comm || !api().ok) return MPCB_ERR_NCCL;
    return api().CommDestroy((ncclComm_t)comm) == ncclSuccess ? MPCB_OK : MPCB_ERR_NCCL;
}

int mpcb_allreduce_min(mpcb_handle *h, void *nccl_comm, double *cost_dev, int64_t *index_dev) {
    if (!h || !nccl_comm || !cost_dev || !index_dev || !api().ok) return MPCB_ERR_NCCL;
    if (cudaSetDevice(h->device) != cudaSuccess) return MPCB_ERR_CUDA;
    int nranks = 0, rank = 0;
    if (mpcb::nccl_comm_info(nccl_comm, &nranks, &rank) != MPCB_OK) return MPCB_ERR_NCCL;
    // scratch of the handle (its device, freed by mpcb_destroy): own record, then one slot per rank
    if (h->nccl_scratch.ensure(sizeof(Rec) * (size_t)(nranks + 1)) != cudaSuccess) return MPCB_ERR_CUDA;
    Rec *mine = h->nccl_scratch.as<Rec>(), *all = mine + 1;
    cudaStream_t st = h->stream;
    pack_kernel<<<1, 1, 0, st>>>(cost_dev, (const long long *)index_dev, mine);
    if (mpcb::nccl_allgather_bytes(nccl_comm, mine, all, sizeof(Rec), st) != MPCB_OK) return MPCB_ERR_NCCL;
    pick_kernel<<<1, 1, 0, st>>>(all, nranks, cost_dev, (long long *)index_dev);
    return cudaGetLastError() == cudaSuccess ? MPCB_OK : MPCB_ERR_CUDA;
}

}  // extern "C"
