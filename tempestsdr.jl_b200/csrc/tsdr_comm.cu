// tsdr_comm.cu -- the one collective of the path, behind the C ABI: partial frame accumulators of a long
// integration (imageOut of src/GUI.jl:175, sharded over GPUs by contiguous frame blocks) are
// combined with an NCCL all-reduce over NVLink.  A Julia host reaches it through ccall like every other entry
// point -- no Python, no torch.distributed on the data path.
//
// NCCL is bound at RUN time (dlopen "libnccl.so.2"; TEMPEST_B200_NCCL overrides the path): the library itself
// links only cudart, loads on hosts without NCCL, and in a process that already holds a copy of NCCL (PyTorch
// bundles one under the same SONAME) the loader hands back that copy.  Only ABI-stable NCCL 2.x entry points
// are used; the handful of types below restate nccl.h (2.11+: ncclRedOpCreatePreMulSum).
//
// The EMA pre-weight a^(frames after the block) is folded into the collective itself: the reduction operator
// is NCCL's PreMulSum (every rank's input is multiplied by its own scalar before the sum), so no separate
// scaling kernel touches the accumulator.
#include "tsdr_internal.cuh"

#include <dlfcn.h>
#include <cstdlib>
#include <mutex>
#include <new>

namespace {

// ---- nccl.h, restated (ABI of NCCL 2.x) -------------------------------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[TSDR_COMM_ID_BYTES]; } ncclUniqueId;   // NCCL_UNIQUE_ID_BYTES = 128
typedef int ncclResult_t;                                             // ncclSuccess = 0
typedef int ncclRedOp_t;                                              // ncclSum = 0
typedef int ncclDataType_t;                                           // ncclUint8 = 1, ncclFloat32 = 7
typedef int ncclScalarResidence_t;                                    // ncclScalarDevice = 0, ncclScalarHostImmediate = 1
constexpr ncclRedOp_t kNcclSum = 0;
constexpr ncclDataType_t kNcclUint8 = 1, kNcclFloat32 = 7;
constexpr ncclScalarResidence_t kNcclScalarHostImmediate = 1;

struct NcclApi {
    void* so = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*RedOpCreatePreMulSum)(ncclRedOp_t*, void*, ncclDataType_t, ncclScalarResidence_t, ncclComm_t) = nullptr;
    ncclResult_t (*RedOpDestroy)(ncclRedOp_t, ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    char where[256] = "";
};

NcclApi g_nccl;
std::once_flag g_nccl_once;
char g_nccl_why[384] = "";

void nccl_load_once() {
    const char* env = getenv("TEMPEST_B200_NCCL");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    void* so = nullptr;
    for (const char* nm : names) {
        if (!nm || !*nm) continue;
        so = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (so) { snprintf(g_nccl.where, sizeof(g_nccl.where), "%s", nm); break; }
        snprintf(g_nccl_why, sizeof(g_nccl_why), "%s", dlerror());
    }
    if (!so) return;
    bool ok = true;
    auto sym = [&](const char* nm) { void* p = dlsym(so, nm); if (!p) { ok = false; snprintf(g_nccl_why, sizeof(g_nccl_why), "symbol %s missing in %s", nm, g_nccl.where); } return p; };
    g_nccl.GetVersion = (decltype(g_nccl.GetVersion))sym("ncclGetVersion");
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
    g_nccl.CommInitAll = (decltype(g_nccl.CommInitAll))sym("ncclCommInitAll");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))sym("ncclAllReduce");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))sym("ncclAllGather");
    g_nccl.RedOpCreatePreMulSum = (decltype(g_nccl.RedOpCreatePreMulSum))sym("ncclRedOpCreatePreMulSum");
    g_nccl.RedOpDestroy = (decltype(g_nccl.RedOpDestroy))sym("ncclRedOpDestroy");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))sym("ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))sym("ncclGroupEnd");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
    if (ok) g_nccl.so = so;
    else dlclose(so);
}

int nccl_api(const NcclApi** api) {
    std::call_once(g_nccl_once, nccl_load_once);
    if (!g_nccl.so) {
        tsdr::set_error("NCCL is not available (%s); the multi-GPU combine needs libnccl.so.2 (set TEMPEST_B200_NCCL to its path)",
                        g_nccl_why[0] ? g_nccl_why : "dlopen failed");
        return TSDR_ERR_UNSUPPORTED;
    }
    *api = &g_nccl;
    return TSDR_OK;
}

int nccl_fail(const NcclApi* api, ncclResult_t r, const char* what) {
    tsdr::set_error("NCCL error %d (%s) in %s", (int)r, api->GetErrorString ? api->GetErrorString(r) : "?", what);
    return TSDR_ERR_NCCL;
}
#define TSDR_NCCL(api, call) do { ncclResult_t _r = (call); if (_r != 0) return nccl_fail(api, _r, #call); } while (0)

}  // namespace

struct tsdr_comm {
    ncclComm_t comm;
    int device, nranks, rank;
    uint64_t collectives;
};

using namespace tsdr;

extern "C" {

int tsdr_comm_available(int* nccl_version) {
    const NcclApi* api = nullptr;
    int rc = nccl_api(&api);
    if (rc) { if (nccl_version) *nccl_version = 0; return rc; }
    int v = 0;
    TSDR_NCCL(api, api->GetVersion(&v));
    if (nccl_version) *nccl_version = v;
    return TSDR_OK;
}

int tsdr_comm_get_unique_id(unsigned char id[TSDR_COMM_ID_BYTES]) {
    TSDR_REQUIRE(id, "id is NULL");
    const NcclApi* api = nullptr;
    int rc = nccl_api(&api); if (rc) return rc;
    ncclUniqueId u;
    TSDR_NCCL(api, api->GetUniqueId(&u));
    memcpy(id, u.internal, TSDR_COMM_ID_BYTES);
    return TSDR_OK;
}

int tsdr_comm_init_rank(tsdr_comm** out, int device, int nranks, int rank, const unsigned char id[TSDR_COMM_ID_BYTES]) {
    TSDR_REQUIRE(out && id, "NULL argument");
    *out = nullptr;
    TSDR_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "rank %d out of range for %d ranks", rank, nranks);
    const NcclApi* api = nullptr;
    int rc = nccl_api(&api); if (rc) return rc;
    int ndev = 0;
    tsdr_device_count(&ndev);
    if (ndev == 0) { set_error("no CUDA device available; libtempest_b200 has no CPU fallback"); return TSDR_ERR_CUDA; }
    TSDR_REQUIRE(device >= 0 && device < ndev, "device %d out of range (%d devices)", device, ndev);
    TSDR_DEVICE(device);
    tsdr_comm* c = new (std::nothrow) tsdr_comm();
    if (!c) return TSDR_ERR_NOMEM;
    c->device = device; c->nranks = nranks; c->rank = rank; c->collectives = 0; c->comm = nullptr;
    ncclUniqueId u;
    memcpy(u.internal, id, TSDR_COMM_ID_BYTES);
    ncclResult_t r = api->CommInitRank(&c->comm, nranks, u, rank);
    if (r != 0) { delete c; return nccl_fail(api, r, "ncclCommInitRank"); }
    *out = c;
    return TSDR_OK;
}

int tsdr_comm_init_all(tsdr_comm** out, int n_devices, const int* devices) {
    TSDR_REQUIRE(out && n_devices >= 1 && n_devices <= 64, "need 1..64 devices");
    const NcclApi* api = nullptr;
    int rc = nccl_api(&api); if (rc) return rc;
    int ndev = 0;
    tsdr_device_count(&ndev);
    if (ndev == 0) { set_error("no CUDA device available; libtempest_b200 has no CPU fallback"); return TSDR_ERR_CUDA; }
    int devs[64];
    for (int i = 0; i < n_devices; ++i) {
        devs[i] = devices ? devices[i] : i;
        TSDR_REQUIRE(devs[i] >= 0 && devs[i] < ndev, "device %d out of range (%d devices)", devs[i], ndev);
    }
    DeviceScope scope;   // ncclCommInitAll moves the current device around
    if ((rc = scope.enter(devs[0]))) return rc;
    ncclComm_t comms[64];
    TSDR_NCCL(api, api->CommInitAll(comms, n_devices, devs));
    for (int i = 0; i < n_devices; ++i) {
        tsdr_comm* c = new (std::nothrow) tsdr_comm();
        if (!c) return TSDR_ERR_NOMEM;
        c->comm = comms[i]; c->device = devs[i]; c->nranks = n_devices; c->rank = i; c->collectives = 0;
        out[i] = c;
    }
    cudaSetDevice(devs[0]);
    return TSDR_OK;
}

int tsdr_comm_info(const tsdr_comm* c, int* device, int* nranks, int* rank, uint64_t* collectives) {
    TSDR_REQUIRE(c, "comm is NULL");
    if (device) *device = c->device;
    if (nranks) *nranks = c->nranks;
    if (rank) *rank = c->rank;
    if (collectives) *collectives = c->collectives;
    return TSDR_OK;
}

int tsdr_comm_group_start(void) {
    const NcclApi* api = nullptr;
    int rc = nccl_api(&api); if (rc) return rc;
    TSDR_NCCL(api, api->GroupStart());
    return TSDR_OK;
}

int tsdr_comm_group_end(void) {
    const NcclApi* api = nullptr;
    int rc = nccl_api(&api); if (rc) return rc;
    TSDR_NCCL(api, api->GroupEnd());
    return TSDR_OK;
}

int tsdr_comm_allreduce_f32(tsdr_comm* c, float* buf_dev, size_t n, float weight, void* stream) {
    TSDR_REQUIRE(c && buf_dev, "NULL argument");
    const NcclApi* api = nullptr;
    int rc = nccl_api(&api); if (rc) return rc;
    TSDR_DEVICE(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    // out = sum over ranks of fl(weight_rank * in_rank): the block's EMA tail weight rides inside the collective.
    // EVERY rank takes the PreMulSum operator, also the one whose weight is 1 (the product is exact): NCCL picks its
    // algorithm per call from (size, type, operator), a built-in sum may go through NVLS on an NVSwitch box where a
    // user-defined operator may not, and ranks that disagree on the algorithm never meet.
    ncclRedOp_t op;
    TSDR_NCCL(api, api->RedOpCreatePreMulSum(&op, &weight, kNcclFloat32, kNcclScalarHostImmediate, c->comm));
    ncclResult_t r = api->AllReduce(buf_dev, buf_dev, n, kNcclFloat32, op, c->comm, st);
    api->RedOpDestroy(op, c->comm);   // NCCL keeps it alive until the enqueued collective has used it
    if (r != 0) return nccl_fail(api, r, "ncclAllReduce(PreMulSum)");
    c->collectives += 1;
    return TSDR_OK;
}

int tsdr_comm_allgather(tsdr_comm* c, const void* send_dev, void* recv_dev, size_t bytes_per_rank, void* stream) {
    TSDR_REQUIRE(c && send_dev && recv_dev, "NULL argument");
    const NcclApi* api = nullptr;
    int rc = nccl_api(&api); if (rc) return rc;
    TSDR_DEVICE(c->device);
    TSDR_NCCL(api, api->AllGather(send_dev, recv_dev, bytes_per_rank, kNcclUint8, c->comm, (cudaStream_t)stream));
    c->collectives += 1;
    return TSDR_OK;
}

int tsdr_chain_allreduce(tsdr_chain* chain, tsdr_comm* c, float weight) {
    TSDR_REQUIRE(chain && c, "NULL argument");
    void* acc = nullptr; size_t n = 0; void* st = nullptr;
    int rc = tsdr_chain_accumulator(chain, &acc, &n);   // joins the chain's auxiliary stream into its primary stream
    if (rc) return rc;
    if ((rc = tsdr_chain_stream(chain, &st))) return rc;
    return tsdr_comm_allreduce_f32(c, (float*)acc, n, weight, st);
}

int tsdr_chain_integrate_device(tsdr_chain* chain, const float* halo_dev, size_t halo_samples, const float* const* bufs_dev,
                                const size_t* n_samples, int n_bufs, tsdr_comm* comm, float weight, int* n_frames) {
    TSDR_REQUIRE(chain && (n_bufs == 0 || (bufs_dev && n_samples)) && n_bufs >= 0, "NULL argument");
    int rc = tsdr_chain_reset(chain);
    if (rc) return rc;
    if (halo_dev && halo_samples && (rc = tsdr_chain_prime_device(chain, halo_dev, halo_samples))) return rc;
    int total = 0;
    for (int i = 0; i < n_bufs; ++i) {
        int nf = 0;
        if ((rc = tsdr_chain_push_device(chain, bufs_dev[i], n_samples[i], &nf))) return rc;
        total += nf;
    }
    if (n_frames) *n_frames = total;
    if (comm && comm->nranks > 1) return tsdr_chain_allreduce(chain, comm, weight);
    if (weight != 1.0f) return tsdr_chain_scale_accumulator(chain, weight);
    void* acc; size_t n;
    return tsdr_chain_accumulator(chain, &acc, &n);   // join the auxiliary stream: the primary stream now covers the block
}

int tsdr_comm_destroy(tsdr_comm* c) {
    if (!c) return TSDR_OK;
    const NcclApi* api = nullptr;
    if (nccl_api(&api) == TSDR_OK && c->comm) {
        DeviceScope scope;
        scope.enter(c->device);
        api->CommDestroy(c->comm);
    }
    delete c;
    return TSDR_OK;
}

}  // extern "C"
