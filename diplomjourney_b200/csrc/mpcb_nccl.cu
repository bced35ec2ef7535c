// mpcb_nccl.cu -- cross-rank reconciliation of a split tree: the ranks own
// contiguous ranges of the first control, each reduces its share to one (cost, index)
// record, and the lexicographic minimum over ranks is taken with two 8-byte NCCL
// all-reduce(min) rounds: float64 cost first, then int64 index among the ranks that hold
// that cost.  Exact for float64 costs and 63-bit indices (a single packed 64-bit word
// cannot hold both).  NCCL is bound at run time with dlopen so that the library loads on
// hosts without NCCL and shares the copy PyTorch already mapped (same SONAME).
#include "../../include/mpcb200.h"

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
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
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
    a.AllReduce = (decltype(a.AllReduce))dlsym(a.lib, "ncclAllReduce");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce;
    return a;
}

// scratch[0] = canonical cost (NaN -> +inf), later the global minimum; scratch[1] = index contribution
__global__ void canon_kernel(const double *cost, double *scratch) {
    double c = *cost;
    scratch[0] = (c == c) ? c : INFINITY;
}
__global__ void contrib_kernel(const double *cost, const long long *index, const double *gmin, long long *contrib) {
    double c = *cost;
    *contrib = (c == *gmin && *index >= 0) ? *index : LLONG_MAX;
}
__global__ void writeback_kernel(const double *gmin, const long long *gidx, double *cost, long long *index) {
    *cost = *gmin;
    *index = (*gidx == LLONG_MAX) ? -1 : *gidx;
}

}  // namespace

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
    (void)h;
    if (!id128 || !comm_out || !api().ok) return MPCB_ERR_NCCL;
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
    cudaStream_t st = (cudaStream_t)mpcb_stream(h);
    static thread_local void *scratch = nullptr;   // 2 x 8 bytes per host thread, never freed
    if (!scratch && cudaMalloc(&scratch, 16) != cudaSuccess) return MPCB_ERR_CUDA;
    double *gmin = (double *)scratch;
    long long *contrib = (long long *)scratch + 1;
    ncclComm_t comm = (ncclComm_t)nccl_comm;
    canon_kernel<<<1, 1, 0, st>>>(cost_dev, gmin);
    if (api().AllReduce(gmin, gmin, 1, ncclFloat64, ncclMin, comm, st) != ncclSuccess) return MPCB_ERR_NCCL;
    contrib_kernel<<<1, 1, 0, st>>>(cost_dev, (const long long *)index_dev, gmin, contrib);
    if (api().AllReduce(contrib, contrib, 1, ncclInt64, ncclMin, comm, st) != ncclSuccess) return MPCB_ERR_NCCL;
    writeback_kernel<<<1, 1, 0, st>>>(gmin, contrib, cost_dev, (long long *)index_dev);
    return cudaGetLastError() == cudaSuccess ? MPCB_OK : MPCB_ERR_CUDA;
}

}  // extern "C"
