// dind_comm.cu — the optional all-gather of outputs after a sharded evaluation (north star (5)):
// dind_comm_get_unique_id / dind_comm_init / dind_comm_destroy / dind_allgather_out of include/dind.h.
//
// The reference has no multi-device path; the evaluation itself shards with no exchange
// (one process per GPU, u and knots replicated, points split by dind_shard_range), so this is the ONLY collective
// of the library and it runs after the kernels, never inside them.
//
// NCCL is bound at run time (dlopen("libnccl.so.2") on the first dind_comm_* call): single-GPU users need no NCCL,
// and a host process that already loaded an NCCL (e.g. torch's bundled one, same soname) shares it.
//
// Layout.  Outputs are column-major (points..., out...): per component column the points are contiguous and the
// columns are n_points apart, so a rank's shard is `n_comp` disjoint segments of the global array.  Each segment
// travels as its own NCCL call straight from the local array into its final place in the global array, all calls of
// one gather inside ncclGroupStart/End (one fused operation on the wire): no staging buffer, no padding, no
// re-packing pass.  Even shards: ncclAllGather per column.  Uneven shards (n_total % world != 0): one ncclBroadcast
// per (rank, column), since ncclAllGather needs equal counts.
#include <dlfcn.h>

#include <mutex>

#include "dind_handle.h"

namespace dind {

// The slice of the NCCL ABI this file uses (stable since NCCL 2.0; types restated so that no NCCL header is
// needed at build time).
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
static_assert(sizeof(ncclUniqueId) == DIND_UNIQUE_ID_BYTES, "unique id size");
enum { ncclSuccess = 0 };
enum { ncclUint8 = 1, ncclFloat32 = 7, ncclFloat64 = 8 };

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    std::string load_error;
};

static NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) {
            const char* e = dlerror();
            api.load_error = e ? e : "dlopen(libnccl.so.2) failed";
            return;
        }
        bool ok = true;
        auto sym = [&](const char* name) {
            void* p = dlsym(api.lib, name);
            if (!p) { ok = false; api.load_error = std::string("missing NCCL symbol ") + name; }
            return p;
        };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
        if (!ok) { dlclose(api.lib); api.lib = nullptr; }
    });
    return api;
}

static int need_nccl() {
    NcclApi& a = nccl();
    if (!a.lib) return fail(DIND_ERR_NCCL, "NCCL is not available: %s", a.load_error.c_str());
    return DIND_OK;
}

#define NCCL_TRY(expr)                                                                                              \
    do {                                                                                                            \
        int r_ = (expr);                                                                                            \
        if (r_ != ncclSuccess)                                                                                      \
            return fail(DIND_ERR_NCCL, "%s failed: %s (%s:%d)", #expr, nccl().GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

struct CommState {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};

void comm_release(dind_interp_s* h) {
    if (h->comm) {
        if (h->comm->comm) nccl().CommDestroy(h->comm->comm);
        delete h->comm;
        h->comm = nullptr;
    }
}

}  // namespace dind

using namespace dind;

extern "C" {

int dind_comm_get_unique_id(void* id_out) {
    if (!id_out) return fail(DIND_ERR_INVALID_ARGUMENT, "id_out is NULL");
    int rc = need_nccl();
    if (rc) return rc;
    ncclUniqueId id;
    NCCL_TRY(nccl().GetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    return DIND_OK;
}

int dind_comm_init(dind_handle_t h, const void* unique_id, int rank, int world) {
    if (!h) return fail(DIND_ERR_INVALID_ARGUMENT, "null handle");
    if (!unique_id) return fail(DIND_ERR_INVALID_ARGUMENT, "unique_id is NULL");
    if (world < 1 || rank < 0 || rank >= world) return fail(DIND_ERR_INVALID_ARGUMENT, "bad rank/world (%d of %d)", rank, world);
    int rc = need_nccl();
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(h->device));
    comm_release(h);
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof id);
    CommState* c = new CommState();
    c->rank = rank;
    c->world = world;
    int r = nccl().CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        delete c;
        return fail(DIND_ERR_NCCL, "ncclCommInitRank(rank %d of %d) failed: %s", rank, world, nccl().GetErrorString(r));
    }
    h->comm = c;
    return DIND_OK;
}

int dind_comm_destroy(dind_handle_t h) {
    if (!h) return fail(DIND_ERR_INVALID_ARGUMENT, "null handle");
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    comm_release(h);
    return DIND_OK;
}

int dind_comm_info(dind_handle_t h, int* rank, int* world, int* nccl_version) {
    if (!h) return fail(DIND_ERR_INVALID_ARGUMENT, "null handle");
    if (rank) *rank = h->comm ? h->comm->rank : 0;
    if (world) *world = h->comm ? h->comm->world : 1;
    if (nccl_version) {
        *nccl_version = 0;
        if (h->comm && nccl().GetVersion) nccl().GetVersion(nccl_version);
    }
    return DIND_OK;
}

int dind_allgather_out(dind_handle_t h, const void* out_local, void* out_global, int64_t n_total, int64_t unit,
                       int64_t n_comp, unsigned flags) {
    if (!h) return fail(DIND_ERR_INVALID_ARGUMENT, "null handle");
    if (n_total < 0 || unit < 1 || n_comp < 1) return fail(DIND_ERR_INVALID_ARGUMENT, "bad all-gather sizes (n_total=%lld unit=%lld n_comp=%lld)", (long long)n_total, (long long)unit, (long long)n_comp);
    if (!out_global || (!out_local && n_total > 0)) return fail(DIND_ERR_INVALID_ARGUMENT, "out_local / out_global is NULL");
    CUDA_TRY(cudaSetDevice(h->device));
    const int world = h->comm ? h->comm->world : 1, rank = h->comm ? h->comm->rank : 0;
    const size_t esz = h->esz;
    int64_t lo = 0, hi = 0;
    dind_shard_range(n_total, rank, world, &lo, &hi);
    const int64_t n_local = hi - lo;
    const char* src = (const char*)out_local;
    char* dst = (char*)out_global;
    if (world == 1) {   // no communicator needed: the local array IS the global one
        if (src != dst && n_total > 0)
            CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)n_total * unit * n_comp * esz, cudaMemcpyDeviceToDevice, h->stream));
    } else {
        if (n_comp * (n_total % world ? world : 1) > 65536)
            return fail(DIND_ERR_UNSUPPORTED, "dind_allgather_out: too many column segments (%lld components)", (long long)n_comp);
        const int dt = h->dtype == DIND_F32 ? ncclFloat32 : ncclFloat64;
        NcclApi& N = nccl();
        NCCL_TRY(N.GroupStart());
        int r = ncclSuccess;
        if (n_total % world == 0) {
            const size_t count = (size_t)n_local * unit;
            for (int64_t c = 0; c < n_comp && r == ncclSuccess; ++c)
                r = N.AllGather(src + (size_t)c * count * esz, dst + (size_t)c * n_total * unit * esz, count, dt, h->comm->comm, h->stream);
        } else {
            for (int root = 0; root < world && r == ncclSuccess; ++root) {
                int64_t rlo, rhi;
                dind_shard_range(n_total, root, world, &rlo, &rhi);
                const size_t count = (size_t)(rhi - rlo) * unit;
                if (!count) continue;
                for (int64_t c = 0; c < n_comp && r == ncclSuccess; ++c) {
                    char* seg = dst + ((size_t)c * n_total + rlo) * unit * esz;
                    const char* mine = root == rank ? src + (size_t)c * n_local * unit * esz : seg;
                    r = N.Broadcast(mine, seg, count, dt, root, h->comm->comm, h->stream);
                }
            }
        }
        const int rg = N.GroupEnd();
        if (r != ncclSuccess) return fail(DIND_ERR_NCCL, "NCCL all-gather of outputs failed: %s", N.GetErrorString(r));
        NCCL_TRY(rg);
        h->launches += 1;
    }
    if (!(flags & DIND_FLAG_ASYNC)) CUDA_TRY(cudaStreamSynchronize(h->stream));
    return DIND_OK;
}

}  // extern "C"
