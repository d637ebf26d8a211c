// riemann_b200 -- NCCL plumbing for the one data-path exchange the engine has: the row-sharded likelihood
// (data-parallel over the N data rows when X does not fit one GPU).  Every rank holds
// the same K chains and a slice of the rows; per likelihood sweep the per-chain partial sums
// (log-likelihood, gradient, metric) are all-reduced over NVLink on the sampler's stream.
//
// NCCL is resolved at run time with dlopen (first the copy the host process already loaded -- the one
// torch.distributed uses -- then the system one), so the library has no link-time NCCL dependency and still
// loads on a box without it; the calls fail loudly if it is absent.
#include "common.cuh"

#include <dlfcn.h>
#include <cstdlib>
#include <cstring>

namespace {

// the part of nccl.h this file needs (stable ABI since NCCL 2.0)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0;

struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& api() {
    static NcclApi a;
    static bool tried = false;
    if (tried) return a;
    tried = true;
    const char* env = getenv("RMN_NCCL_LIB");
    if (env && *env) a.h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!a.h) a.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);        // already in the process (torch)
    if (!a.h) a.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!a.h) return a;
#define RMN_SYM(field, name) *(void**)(&a.field) = dlsym(a.h, name)
    RMN_SYM(GetUniqueId, "ncclGetUniqueId");
    RMN_SYM(CommInitRank, "ncclCommInitRank");
    RMN_SYM(CommDestroy, "ncclCommDestroy");
    RMN_SYM(AllReduce, "ncclAllReduce");
    RMN_SYM(GroupStart, "ncclGroupStart");
    RMN_SYM(GroupEnd, "ncclGroupEnd");
    RMN_SYM(GetErrorString, "ncclGetErrorString");
#undef RMN_SYM
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.GroupStart && a.GroupEnd &&
           a.GetErrorString;
    return a;
}

int need_api() {
    if (api().ok) return RMN_OK;
    rmn_set_error("NCCL (libnccl.so.2) could not be loaded: the row-sharded data mode needs it "
                  "(set RMN_NCCL_LIB to its path)");
    return RMN_ERR_UNSUPPORTED;
}

#define RMN_NCCL(call)                                                                   \
    do {                                                                                 \
        const ncclResult_t r_ = (call);                                                  \
        if (r_ != 0) {                                                                   \
            rmn_set_error("NCCL error %d (%s) at %s:%d", r_, api().GetErrorString(r_), __FILE__, __LINE__); \
            return RMN_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

}  // namespace

int rmn_rowcomm_unique_id(void* out, size_t nbytes) {
    if (!out || nbytes < sizeof(ncclUniqueId)) {
        rmn_set_error("rmn_nccl_unique_id: need a buffer of at least %d bytes", (int)sizeof(ncclUniqueId));
        return RMN_ERR_PARAM;
    }
    if (int rc = need_api()) return rc;
    ncclUniqueId id;
    RMN_NCCL(api().GetUniqueId(&id));
    memcpy(out, &id, sizeof(id));
    return RMN_OK;
}

int rmn_rowcomm_init(RowComm* rc, const void* unique_id, size_t nbytes, int rank, int world) {
    if (!unique_id || nbytes < sizeof(ncclUniqueId) || world < 1 || rank < 0 || rank >= world) {
        rmn_set_error("row communicator: need a %d-byte NCCL unique id and 0 <= rank < world (got rank %d of %d)",
                      (int)sizeof(ncclUniqueId), rank, world);
        return RMN_ERR_PARAM;
    }
    if (rc->comm) { rmn_set_error("row communicator is already attached to this sampler"); return RMN_ERR_PARAM; }
    if (int e = need_api()) return e;
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    ncclComm_t c = nullptr;
    RMN_NCCL(api().CommInitRank(&c, world, id, rank));
    rc->comm = c; rc->rank = rank; rc->world = world;
    return RMN_OK;
}

int rmn_rowcomm_allreduce_f64(RowComm* rc, double* const* bufs, const size_t* counts, int nbuf, cudaStream_t stream) {
    if (!rc->comm) return RMN_OK;
    RMN_NCCL(api().GroupStart());
    for (int i = 0; i < nbuf; ++i)
        if (counts[i])
            RMN_NCCL(api().AllReduce(bufs[i], bufs[i], counts[i], NCCL_FLOAT64, NCCL_SUM, (ncclComm_t)rc->comm, stream));
    RMN_NCCL(api().GroupEnd());
    return RMN_OK;
}

void rmn_rowcomm_destroy(RowComm* rc) {
    if (rc->comm && api().ok) api().CommDestroy((ncclComm_t)rc->comm);
    rc->comm = nullptr;
}
