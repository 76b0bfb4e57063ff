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

#include <algorithm>
#include <mutex>
#include <vector>

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
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
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
        api.Send = (decltype(api.Send))sym("ncclSend");
        api.Recv = (decltype(api.Recv))sym("ncclRecv");
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
    // dind_eval_unstructured_allgather: the exchanges run on their own stream, ordered against the handle's stream by events
    cudaStream_t stream = nullptr;
    std::vector<cudaEvent_t> chunk_ready;
    cudaEvent_t done = nullptr;
};

void comm_release(dind_interp_s* h) {
    if (h->comm) {
        if (h->comm->stream) {
            cudaStreamSynchronize(h->comm->stream);
            cudaStreamDestroy(h->comm->stream);
        }
        for (cudaEvent_t e : h->comm->chunk_ready) cudaEventDestroy(e);
        if (h->comm->done) cudaEventDestroy(h->comm->done);
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

int dind_eval_unstructured_allgather(dind_handle_t h, void* out_global, int64_t n_total, const int* derivative_orders,
                                     int n_chunks, unsigned flags) {
    if (!h) return fail(DIND_ERR_INVALID_ARGUMENT, "null handle");
    if (n_total < 0) return fail(DIND_ERR_INVALID_ARGUMENT, "n_total must be non-negative");
    if (!out_global) return fail(DIND_ERR_INVALID_ARGUMENT, "out_global is NULL");
    CUDA_TRY(cudaSetDevice(h->device));
    {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, out_global) != cudaSuccess || a.type != cudaMemoryTypeDevice) {
            cudaGetLastError();
            return fail(DIND_ERR_INVALID_ARGUMENT, "out_global must be a device array");
        }
    }
    CommState* cs = h->comm;
    const int world = cs ? cs->world : 1, rank = cs ? cs->rank : 0;
    int64_t lo = 0, hi = 0;
    dind_shard_range(n_total, rank, world, &lo, &hi);
    for (int d = 0; d < h->n_in; ++d) {
        if (!h->dims[d].teval_set) return fail(DIND_ERR_STATE, "t_eval of dimension %d has not been set (dind_set_t_eval)", d + 1);
        if (h->dims[d].n_eval != hi - lo)
            return fail(DIND_ERR_SIZE_MISMATCH, "rank %d of %d owns %lld of %lld points, t_eval of dimension %d has %lld",
                        rank, world, (long long)(hi - lo), (long long)n_total, d + 1, (long long)h->dims[d].n_eval);
    }
    // one-time work (component-fastest / cell-record copies of u) is sized by the whole shard, not by a chunk
    int rc = dind_build_caches(h, 0, derivative_orders, DIND_FLAG_ASYNC);
    if (rc) return rc;
    const size_t esz = h->esz;
    const int64_t n_comp = h->n_comp;
    char* dst = (char*)out_global;
    const int64_t per_rank = (n_total + world - 1) / std::max(world, 1);
    int K = n_chunks > 0 ? n_chunks : (int)std::min<int64_t>(16, std::max<int64_t>(1, per_rank / 2000000));
    if (world == 1) K = 1;
    K = std::min(K, 64);
    if (world > 1) {
        if (!cs->stream) CUDA_TRY(cudaStreamCreateWithFlags(&cs->stream, cudaStreamNonBlocking));
        if (!cs->done) CUDA_TRY(cudaEventCreateWithFlags(&cs->done, cudaEventDisableTiming));
        while ((int)cs->chunk_ready.size() < K) {
            cudaEvent_t e;
            CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            cs->chunk_ready.push_back(e);
        }
    }
    auto cut = [K](int64_t n, int j) { return (int64_t)((__int128)n * j / K); };   // chunk j of a shard of n rows: [cut(j), cut(j + 1))
    NcclApi& N = nccl();
    const int dt = h->dtype == DIND_F32 ? ncclFloat32 : ncclFloat64;
    for (int j = 0; j < K; ++j) {
        const int64_t a = cut(hi - lo, j), b = cut(hi - lo, j + 1);
        if (b > a && (rc = eval_point_rows(h, dst + (size_t)(lo + a) * esz, a, b, n_total, derivative_orders))) return rc;
        if (world == 1) continue;
        CUDA_TRY(cudaEventRecord(cs->chunk_ready[j], h->stream));
        CUDA_TRY(cudaStreamWaitEvent(cs->stream, cs->chunk_ready[j], 0));
        NCCL_TRY(N.GroupStart());
        int r = ncclSuccess;
        for (int step = 1; step < world && r == ncclSuccess; ++step) {
            const int to = (rank + step) % world, from = (rank - step + world) % world;
            int64_t flo, fhi;
            dind_shard_range(n_total, from, world, &flo, &fhi);
            const int64_t fa = cut(fhi - flo, j), fb = cut(fhi - flo, j + 1);
            for (int64_t c = 0; c < n_comp && r == ncclSuccess; ++c) {
                if (b > a) r = N.Send(dst + ((size_t)c * n_total + lo + a) * esz, (size_t)(b - a), dt, to, cs->comm, cs->stream);
                if (fb > fa && r == ncclSuccess)
                    r = N.Recv(dst + ((size_t)c * n_total + flo + fa) * esz, (size_t)(fb - fa), dt, from, cs->comm, cs->stream);
            }
        }
        const int rg = N.GroupEnd();
        if (r != ncclSuccess) return fail(DIND_ERR_NCCL, "NCCL exchange of output rows failed: %s", N.GetErrorString(r));
        NCCL_TRY(rg);
        h->launches += 1;
    }
    if (world > 1) {
        CUDA_TRY(cudaEventRecord(cs->done, cs->stream));
        CUDA_TRY(cudaStreamWaitEvent(h->stream, cs->done, 0));   // whatever follows on the handle's stream sees the global output
    }
    if (!(flags & DIND_FLAG_ASYNC)) CUDA_TRY(cudaStreamSynchronize(h->stream));
    return DIND_OK;
}

}  // extern "C"
