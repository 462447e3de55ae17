// The one exchange step of the path: when a frequency sweep is sharded over several GPUs (one process each), every rank needs the
// same room modes before it can start - the modal set-up (modal.cu) deals its column panels to the ranks and exchanges them
// over NCCL (NVLink / NVSwitch).  The solve itself has no collective (frequencies are independent).
//
// NCCL is opened lazily with dlopen when a communicator is first asked for: a single-GPU user never loads it, and a process that
// already holds torch's bundled libnccl.so.2 gets that same copy (one NCCL per process).
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "common.cuh"
#include "solver.cuh"

namespace fdb {

namespace {
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        auto sym = [&](const char* n) { return dlsym(h, n); };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Broadcast && api.AllReduce && api.GroupStart && api.GroupEnd) api.handle = h;
    });
    FDB_REQUIRE(api.handle != nullptr, "NCCL (libnccl.so.2) is not available: the multi-GPU modal set-up needs it");
    return api;
}

void check(ncclResult_t r, const char* what) {
    if (r != ncclSuccess) {
        const char* msg = nccl().GetErrorString ? nccl().GetErrorString(r) : "?";
        throw std::runtime_error(std::string("NCCL: ") + what + ": " + msg);
    }
}
}  // namespace

struct RankComm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};

static_assert(sizeof(ncclUniqueId) == 128, "fdb_comm_unique_id hands out 128 bytes");

void comm_unique_id(char* out128) {
    ncclUniqueId id;
    check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(out128, &id, 128);
}

RankComm* comm_create(const char* id128, int rank, int world) {
    FDB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank / world");
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    RankComm* c = new RankComm;
    c->rank = rank;
    c->world = world;
    try {
        check(nccl().CommInitRank(&c->comm, world, id, rank), "ncclCommInitRank");
    } catch (...) {
        delete c;
        throw;
    }
    return c;
}

void comm_destroy(RankComm* c) {
    if (!c) return;
    if (c->comm) nccl().CommDestroy(c->comm);
    delete c;
}

int comm_rank(const RankComm* c) { return c ? c->rank : 0; }
int comm_world(const RankComm* c) { return c ? c->world : 1; }

// in place: rank r holds [offs[r], offs[r + 1]) of buf (in doubles); afterwards every rank holds all of it
void comm_allgatherv(RankComm* c, double* buf, const std::vector<size_t>& offs, cudaStream_t s) {
    if (!c || c->world == 1) return;
    check(nccl().GroupStart(), "ncclGroupStart");
    for (int r = 0; r < c->world; ++r) {
        const size_t n = offs[r + 1] - offs[r];
        if (n == 0) continue;
        check(nccl().Broadcast(buf + offs[r], buf + offs[r], n, ncclDouble, r, c->comm, s), "ncclBroadcast");
    }
    check(nccl().GroupEnd(), "ncclGroupEnd");
}

void comm_bcast(RankComm* c, double* buf, size_t n, int root, cudaStream_t s) {
    if (!c || c->world == 1) return;
    check(nccl().Broadcast(buf, buf, n, ncclDouble, root, c->comm, s), "ncclBroadcast");
}

void comm_allreduce_sum(RankComm* c, double* buf, size_t n, cudaStream_t s) {
    if (!c || c->world == 1) return;
    check(nccl().AllReduce(buf, buf, n, ncclDouble, ncclSum, c->comm, s), "ncclAllReduce");
}

}  // namespace fdb
