// dind_handle.h — the state behind dind_handle_t (include/dind.h) and the error plumbing, shared by the
// translation units that implement the C ABI (dind_api.cu: evaluation; dind_comm.cu: NCCL all-gather).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/dind.h"
#include "dind_kernels.h"

namespace dind {

// sets the calling thread's dind_last_error() message and returns `code`
int fail(int code, const char* fmt, ...);

#define CUDA_TRY(expr)                                                                                \
    do {                                                                                              \
        cudaError_t e_ = (expr);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            cudaGetLastError();                                                                       \
            return ::dind::fail(DIND_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                             \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap && p) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        if (bytes == 0) bytes = 16;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct AxisTable {
    int order = -1;
    DevBuf i0, w, flag;
};

struct DimState {
    bool set = false;
    int kind = 0, degree = 0, left = 1;
    std::vector<double> t;          // distinct knots (exact for both element types)
    std::vector<int64_t> mult;
    std::vector<double> knots;      // t, or knots_all for BSpline
    int width = 0;
    int64_t n_u = 0;                // expected size(u, dim)
    DevBuf d_knots, d_keys, d_recip, d_lut;
    int lut_n = 0;              // buckets of the search table (0 = none)
    double lut_t0 = 0, lut_scale = 0;
    // t_eval
    bool teval_set = false;
    const void* d_teval = nullptr;  // device pointer (owned copy or borrowed)
    DevBuf teval_owned;
    int64_t n_eval = 0;
    int max_deriv = 0;
    DevBuf d_idx;
    bool idx_valid = false;
    AxisTable table;
    void invalidate_caches() {
        idx_valid = false;
        table.order = -1;
    }
    void release() {
        d_knots.release(); d_keys.release(); d_recip.release(); d_lut.release(); teval_owned.release();
        d_idx.release(); table.i0.release(); table.w.release(); table.flag.release();
    }
};


struct CommState;   // dind_comm.cu

// dind_api.cu, for dind_comm.cu: unstructured evaluation of the cached points [a, b) of this handle into rows that are
// `comp_stride` elements apart per output component (out_rows = address of row a of component 0); queued, not waited for
int eval_point_rows(dind_interp_s* h, void* out_rows, int64_t a, int64_t b, int64_t comp_stride, const int* derivative_orders);

}  // namespace dind

struct dind_interp_s {
    dind::Knobs knobs;   // DIND_* experiment switches, read from the environment once in dind_create
    int n_in = 0, dtype = 0, device = 0;
    size_t esz = 4;
    cudaStream_t stream = nullptr;
    dind::DimState dims[dind::MAXN];
    // u
    bool u_set = false;
    const void* d_u = nullptr;
    dind::DevBuf u_owned;
    std::vector<int64_t> u_dims;
    int64_t comp_stride = 0, n_comp = 1;
    // NURBS weights
    bool w_set = false;
    std::vector<int64_t> w_dims;
    const void* d_w = nullptr;
    dind::DevBuf w_owned;
    dind::DevBuf uw;          // u pre-multiplied by the weights
    bool uw_valid = false;
    dind::DevBuf split_scratch;   // stage-1 results of the two-stage 4-D grid evaluation
    // NURBS: the stage-1 result of the WEIGHTS (the denominator's T) depends on weights, knots and t_eval but not on u —
    // "repeated with a new u" re-uses it.  den_epoch counts everything that invalidates it; split_den_epoch is the epoch
    // at which the copy in split_scratch was computed (0 = none).
    uint64_t den_epoch = 1, split_den_epoch = 0;
    dind::DevBuf u_aos;       // component-fastest copy of u (vector-valued unstructured evaluation)
    bool u_aos_valid = false;
    dind::DevBuf u_rec;       // cell-record copy of u_aos (multilinear tables larger than L2 at iid points)
    bool u_rec_valid = false, u_rec_declined = false;   // declined: no room for it — do not ask again until the next dind_set_u
    // cell-sorted plan for eval_unstructured (dind_plan.cu)
    bool plan_valid = false;
    int64_t plan_n = 0;
    dind::DevBuf plan_perm, plan_stage, plan_scratch, plan_slot2, plan_dest2, plan_cells;
    bool plan_has_cells = false;   // plan_cells holds the packed cell words of the current plan
    bool plan_buckets_valid = false;
    // work items over the plan (tile kernels), one set for evaluations [0] and one for fused gradients [1] (they cut
    // the plan into items of different lengths): n_items + 1 start positions, item-major copies of the cell words and
    // of the sorted t_eval
    struct PlanItems {
        dind::DevBuf items, cells, t[dind::MAXN];
        bool valid = false;
        int pts = 0;
        int64_t n = 0;
        void release() {
            items.release(); cells.release();
            for (int d = 0; d < dind::MAXN; ++d) t[d].release();
            valid = false;
        }
    } plan_items[2];
    dind::DevBuf plan_t[dind::MAXN];
    // staging
    dind::DevBuf out_stage[2];            // host `out`: device staging, double buffered for the pipelined (ASYNC) path
    int out_toggle = 0;
    cudaStream_t copy_stream = nullptr;   // D2H of results, so that it overlaps the next step's H2D + kernel
    cudaEvent_t ev_kernel = nullptr, ev_d2h[2] = {nullptr, nullptr};
    dind::DevBuf pts_stage[dind::MAXN];
    void* small_host = nullptr;     // pinned, device-mapped staging of the small-batch (latency) path of dind_eval_points
    int64_t out_comp_stride_override = 0;   // != 0: component columns of `out` are this many elements apart (eval_point_rows)
    std::string last_kernel = "none";
    int64_t launches = 0;
    dind::CommState* comm = nullptr;   // NCCL communicator of dind_comm_init (dind_comm.cu), null = single GPU
};
