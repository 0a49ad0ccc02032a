// mp_comm.cu -- multi-GPU entry points of libmidaspom_cuda.so: NCCL behind the C ABI.
//
// Replaces the MPI plumbing of MIDASPOM_MPI (main_MIDASPOM_MPI.c:344-372 MPI_Init / rank / row split, :483-505 the
// MPI_Send / MPI_Recv gather of the per-rank result rows on rank 0) for the two ways this engine uses several GPUs:
//   * independent chains per rank (chain_offset): mp_gather_draws all-gathers the recorded draws so that every rank
//     holds all chains (R-hat / ESS need them together);
//   * one chain sharded over the ranks (mp_sweep_sharded, in mp_engine.cu): all-gather of the owned connectivity
//     columns and broadcast of the owned year rows, on the engine stream, no host synchronisation inside a sweep.
// NCCL is resolved with dlopen at mp_comm_init, so the library itself loads on boxes without NCCL and, inside a
// process that already holds a libnccl.so.2 (PyTorch), uses that copy.  One process per GPU (mp_comm_init with a
// shared unique id) or one process driving several engines (mp_comm_init_all) are both supported.
#include "mp_host.h"
#include "mp_comm.h"

#include <dlfcn.h>
#include <nccl.h>

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};
NcclApi g_nccl;

bool load_nccl()
{
    if (g_nccl.lib) return true;
    // MP_NCCL_LIB names the copy to use.  A process that will also load PyTorch must take PyTorch's bundled NCCL: two libraries
    // with the SONAME libnccl.so.2 cannot coexist, and libtorch_cuda.so needs symbols of its own (newer) version -- the Python
    // mirror sets the variable to that copy before the first mp_comm_* call (engine.py: _prefer_bundled_nccl).
    const char *names[] = { getenv("MP_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
    void *lib = nullptr;
    for (const char *nm : names) if (nm && *nm && (lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL)) != nullptr) break;
    if (!lib) { g_nccl.err = std::string("libnccl.so.2 not found: ") + dlerror(); return false; }
    bool ok = true;
    auto sym = [&](const char *nm) { void *p = dlsym(lib, nm); if (!p) { ok = false; g_nccl.err = std::string("NCCL symbol missing: ") + nm; } return p; };
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
    g_nccl.CommInitAll = (decltype(g_nccl.CommInitAll))sym("ncclCommInitAll");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))sym("ncclAllGather");
    g_nccl.Broadcast = (decltype(g_nccl.Broadcast))sym("ncclBroadcast");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))sym("ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))sym("ncclGroupEnd");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
    if (!ok) { dlclose(lib); return false; }
    g_nccl.lib = lib;
    return true;
}

#define NK(call)                                                                                         \
    do {                                                                                                 \
        ncclResult_t r_ = (call);                                                                        \
        if (r_ != ncclSuccess) {                                                                         \
            h->err = std::string(#call) + ": " + g_nccl.GetErrorString(r_);                              \
            return MP_ERR_CUDA;                                                                          \
        }                                                                                                \
    } while (0)

int attach(mp_engine *h, ncclComm_t comm, int nranks, int rank)
{
    h->comm = comm; h->comm_size = nranks; h->comm_rank = rank;
    // one chain over the ranks: rank r evaluates the connectivity of the scan-order slots [r per, (r + 1) per) (whole
    // k_conn CTAs: multiples of 256) and scans the (chain, year) tasks r, r + nranks, ...
    const int N = h->cfg.n_patches;
    int per = (N + nranks - 1) / nranks;
    per = (per + 255) / 256 * 256;
    h->comm_per = per;
    return MP_OK;
}

}  // namespace

// ---- used by mp_engine.cu (mp_sweep_sharded)
int mp_comm_group_start(mp_engine *h) { NK(g_nccl.GroupStart()); return MP_OK; }
int mp_comm_group_end(mp_engine *h) { NK(g_nccl.GroupEnd()); return MP_OK; }
int mp_comm_allgather(mp_engine *h, const void *send, void *recv, size_t bytes_per_rank)
{
    NK(g_nccl.AllGather(send, recv, bytes_per_rank, ncclChar, (ncclComm_t)h->comm, h->stream));
    return MP_OK;
}
int mp_comm_broadcast(mp_engine *h, void *buf, size_t bytes, int root)
{
    NK(g_nccl.Broadcast(buf, buf, bytes, ncclChar, root, (ncclComm_t)h->comm, h->stream));
    return MP_OK;
}

extern "C" {

int mp_comm_unique_id(void *id128)
{
    if (!id128) return MP_ERR_ARG;
    static_assert(sizeof(ncclUniqueId) == MP_COMM_ID_BYTES, "MP_COMM_ID_BYTES must be sizeof(ncclUniqueId)");
    if (!load_nccl()) return MP_ERR_UNSUPPORTED;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return MP_ERR_CUDA;
    memcpy(id128, &id, sizeof id);
    return MP_OK;
}

const char *mp_comm_last_error(void) { return g_nccl.err.c_str(); }

int mp_comm_init(mp_engine *h, int nranks, int rank, const void *id128)
{
    if (!h || !id128) return MP_ERR_ARG;
    REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, MP_ERR_ARG, "mp_comm_init: need 0 <= rank < nranks");
    REQUIRE(!h->comm, MP_ERR_STATE, "mp_comm_init: the engine already has a communicator");
    REQUIRE(load_nccl(), MP_ERR_UNSUPPORTED, "mp_comm_init: " + g_nccl.err);
    CK(cudaSetDevice(h->cfg.device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclComm_t comm = nullptr;
    NK(g_nccl.CommInitRank(&comm, nranks, id, rank));
    return attach(h, comm, nranks, rank);
}

int mp_comm_init_all(mp_engine **hs, int n)
{
    if (!hs || n < 1) return MP_ERR_ARG;
    mp_engine *h = hs[0];
    if (!h) return MP_ERR_ARG;
    REQUIRE(load_nccl(), MP_ERR_UNSUPPORTED, "mp_comm_init_all: " + g_nccl.err);
    std::vector<int> devs(n);
    for (int i = 0; i < n; i++) {
        REQUIRE(hs[i] && !hs[i]->comm, MP_ERR_STATE, "mp_comm_init_all: null engine or engine with a communicator");
        devs[i] = hs[i]->cfg.device;
        for (int j = 0; j < i; j++) REQUIRE(devs[j] != devs[i], MP_ERR_ARG, "mp_comm_init_all: one engine per device");
    }
    std::vector<ncclComm_t> comms(n);
    NK(g_nccl.CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; i++) attach(hs[i], comms[i], n, i);
    return MP_OK;
}

int mp_comm_destroy(mp_engine *h)
{
    if (!h) return MP_ERR_ARG;
    if (!h->comm) return MP_OK;
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->stream));
    NK(g_nccl.CommDestroy((ncclComm_t)h->comm));
    h->comm = nullptr; h->comm_size = 1; h->comm_rank = 0;
    return MP_OK;
}

int mp_comm_rank(mp_engine *h) { return h ? h->comm_rank : MP_ERR_ARG; }
int mp_comm_size(mp_engine *h) { return h ? h->comm_size : MP_ERR_ARG; }

// Every rank's recorded draws [first, first + count): out[rank][sweep][chain][MP_NDRAW] on every rank -- the MPI_Send /
// MPI_Recv collection of main_MIDASPOM_MPI.c:483-505 as one NCCL all-gather.  All ranks must hold the same n_chains.
int mp_gather_draws(mp_engine *h, int first, int count, double *out)
{
    if (!h || !out) return MP_ERR_ARG;
    CK(cudaSetDevice(h->cfg.device));
    REQUIRE(first >= 0 && count >= 0 && first + count <= h->ndraws, MP_ERR_ARG, "mp_gather_draws: range outside recorded draws");
    const size_t row = nC(h) * MP_NDRAW * 8, bytes = (size_t)count * row;
    if (bytes == 0) return MP_OK;
    if (!h->comm || h->comm_size == 1) {
        CK(cudaMemcpyAsync(out, (const char *)h->d_draws + (size_t)first * row, bytes, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return MP_OK;
    }
    void *d_all = nullptr;
    CK(cudaMalloc(&d_all, bytes * (size_t)h->comm_size));
    ncclResult_t r = g_nccl.AllGather((const char *)h->d_draws + (size_t)first * row, d_all, bytes, ncclChar, (ncclComm_t)h->comm, h->stream);
    cudaError_t e = cudaSuccess;
    if (r == ncclSuccess) {
        e = cudaMemcpyAsync(out, d_all, bytes * (size_t)h->comm_size, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    }
    cudaFree(d_all);
    if (r != ncclSuccess) { h->err = std::string("ncclAllGather: ") + g_nccl.GetErrorString(r); return MP_ERR_CUDA; }
    if (e != cudaSuccess) { h->err = std::string("mp_gather_draws: ") + cudaGetErrorString(e); return MP_ERR_CUDA; }
    return MP_OK;
}

// The same gather for one process driving all engines of the communicator (mp_comm_init_all): the all-gathers of the
// local engines form one NCCL group; out (host) receives the copy assembled on engines[0].
int mp_gather_draws_all(mp_engine **hs, int n, int first, int count, double *out)
{
    if (!hs || n < 1 || !hs[0] || !out) return MP_ERR_ARG;
    mp_engine *h = hs[0];
    const size_t row = nC(h) * MP_NDRAW * 8, bytes = (size_t)count * row;
    for (int i = 0; i < n; i++) {
        REQUIRE(hs[i] && hs[i]->comm && hs[i]->comm_size == n, MP_ERR_STATE, "mp_gather_draws_all: every engine needs the communicator of mp_comm_init_all");
        REQUIRE(hs[i]->cfg.n_chains == h->cfg.n_chains, MP_ERR_ARG, "mp_gather_draws_all: all engines must hold the same number of chains");
        REQUIRE(first >= 0 && count >= 0 && first + count <= hs[i]->ndraws, MP_ERR_ARG, "mp_gather_draws_all: range outside recorded draws");
    }
    if (bytes == 0) return MP_OK;
    std::vector<void *> d_all(n, nullptr);
    int rc = MP_OK;
    for (int i = 0; i < n && rc == MP_OK; i++) {
        if (cudaSetDevice(hs[i]->cfg.device) != cudaSuccess || cudaMalloc(&d_all[i], bytes * (size_t)n) != cudaSuccess) { h->err = "mp_gather_draws_all: cudaMalloc failed"; rc = MP_ERR_CUDA; }
    }
    if (rc == MP_OK) {
        ncclResult_t r = g_nccl.GroupStart();
        for (int i = 0; i < n && r == ncclSuccess; i++)
            r = g_nccl.AllGather((const char *)hs[i]->d_draws + (size_t)first * row, d_all[i], bytes, ncclChar, (ncclComm_t)hs[i]->comm, hs[i]->stream);
        const ncclResult_t r2 = g_nccl.GroupEnd();
        if (r == ncclSuccess) r = r2;
        if (r != ncclSuccess) { h->err = std::string("mp_gather_draws_all: ") + g_nccl.GetErrorString(r); rc = MP_ERR_CUDA; }
    }
    for (int i = 0; i < n && rc == MP_OK; i++) {
        cudaSetDevice(hs[i]->cfg.device);
        if (i == 0 && cudaMemcpyAsync(out, d_all[0], bytes * (size_t)n, cudaMemcpyDeviceToHost, hs[0]->stream) != cudaSuccess) rc = MP_ERR_CUDA;
        if (cudaStreamSynchronize(hs[i]->stream) != cudaSuccess) { h->err = "mp_gather_draws_all: stream error"; rc = MP_ERR_CUDA; }
    }
    for (int i = 0; i < n; i++) if (d_all[i]) { cudaSetDevice(hs[i]->cfg.device); cudaFree(d_all[i]); }
    return rc;
}

}  // extern "C"
