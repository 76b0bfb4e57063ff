// dind_kernels.h — host-callable launchers of the CUDA kernels (one explicit instantiation per
// element type in the .cu that defines them).
#pragma once
#include "dind_device.cuh"

namespace dind {

// set_idx_kernel (src/interpolation_utils.jl:156-161): idx_out[i] = get_idx(dim, t_eval[i]), 1-based.
template <typename T> cudaError_t launch_idx(const DimParams<T>& D, int32_t* idx_out, cudaStream_t s);
// basis_function_eval_kernel (src/spline_utils.jl:174-191): out[(i, k, r)] column-major.
template <typename T> cudaError_t launch_basis_cache(const DimParams<T>& D, int max_deriv, T* out, cudaStream_t s);
// per-axis stencil tables for grid evaluation at one derivative order.
template <typename T> cudaError_t launch_axis_table(const DimParams<T>& D, int order, int32_t* tab_i0, T* tab_w,
                                                    uint8_t* tab_flag, cudaStream_t s);
// bucket table of the interval search: lut[b] = #{i : lut_bucket(knots[i]) < b}, b = 0..n_buckets
template <typename T> cudaError_t launch_lut(const T* knots, int n_knots, T t0, T scale, int n_buckets, int2* lut, cudaStream_t s);
// uw[i + c*n] = u[i + c*n] * w[i]  (NURBS numerator, src/interpolation_methods.jl:124-131)
template <typename T> cudaError_t launch_premultiply(const T* u, const T* w, T* uw, int64_t n, int n_comp, cudaStream_t s);
// generic evaluation: any N_in <= MAXN, any kinds, any degree <= MAXW-1
template <typename T> cudaError_t launch_eval_generic(const EvalParams<T>& P, bool grid, cudaStream_t s);

// Cell-sorted plan for eval_unstructured (dind_plan.cu): perm = point indices sorted by the
// offset of their first stencil node in u (points of one cell become neighbours, cells follow
// memory order), t_sorted[d] = t_eval[d][perm].  cells != nullptr: the sort key is the packed per-dimension
// first nodes (DimParams::cell_shift, cell_bits bits in all) and the sorted keys are kept in cells[].
// scratch_bytes in/out: pass *scratch = nullptr to query the scratch size.
// recs (optional, n_points * plan_rec_elems(n_in) elements, 32-byte aligned): room for one coordinate record per point;
// with it the sorted copies of t_eval cost one random read per point instead of one per point and dimension.
template <typename T> int plan_rec_elems(int n_in);
template <typename T>
cudaError_t plan_sort_points(const EvalParams<T>& P, uint32_t* perm, T* const* t_sorted, uint32_t* cells, int cell_bits,
                             void* scratch, size_t* scratch_bytes, T* recs, cudaStream_t s);

// Work items over the sorted cell words (dind_plan.cu): the plan positions k whose distance from the start of their
// run of equal cells is a multiple of `pts` — every item is <= pts consecutive points of one cell.  Two calls:
// items == nullptr counts (n_items_out), then the same call with items (n_items + 1 entries; the last one is n).
// scratch: n * 4 bytes + CUB temp (scratch_bytes in/out; *scratch = nullptr queries).
cudaError_t plan_build_items(const uint32_t* cells, int64_t n, int pts, uint32_t* items, int64_t* n_items_out, void* scratch,
                             size_t* scratch_bytes, cudaStream_t s);
// Cell-sorted evaluation over work items (dind_fast_tile.cu): one thread = one item.
template <typename T> int tile_points_per_item(const EvalParams<T>& P, bool gradient);   // 0 = no tile kernel for this configuration
// item-major copies of the plan's inputs: item_cells[i], item_t[d][i * pts + q] (padding slots repeat the item's last point)
template <typename T> cudaError_t plan_item_layout(const EvalParams<T>& P, int pts, uint32_t* item_cells, T* const* item_t, cudaStream_t s);
// P for try_eval_fast_tile: P.cells / P.dim[d].t_eval are the ITEM-MAJOR arrays, P.items the item starts
template <typename T> bool try_eval_fast_tile(const EvalParams<T>& P, bool gradient, cudaStream_t s, const char** name,
                                              cudaError_t* err, int* launches);

// Component-fastest repack of u for vector-valued unstructured evaluation (dind_fast_unstructured_aos.cu).
// elements per node of the component-fastest copy: 3 components are padded to 4 (one aligned vector per node)
__host__ __device__ inline int aos_node_stride(int n_comp) { return n_comp == 3 ? 4 : n_comp; }
template <typename T> cudaError_t launch_aos_repack(const T* u, T* u_aos, int64_t n, int n_comp, cudaStream_t s);
template <typename T> bool aos_supported(const EvalParams<T>& P);
template <typename T> bool try_eval_fast_unstructured_aos(const EvalParams<T>& P, cudaStream_t s, const char** name,
                                                          cudaError_t* err, int* launches);

// Cell-record copy of u for multilinear tables larger than L2, evaluated at iid points (dind_fast_unstructured_rec.cu):
// u_rec[cell(i_2..i_N)][i_1][corner(k_2..k_N)][component] — the 2^N corners of a point are one contiguous block.
template <typename T> bool rec_supported(const EvalParams<T>& P);
template <typename T> size_t rec_table_bytes(const EvalParams<T>& P);
template <typename T> void rec_fill_strides(EvalParams<T>& P);
template <typename T> cudaError_t launch_rec_build(const EvalParams<T>& P, T* u_rec, cudaStream_t s);   // from P.u_aos
template <typename T> bool try_eval_fast_unstructured_rec(const EvalParams<T>& P, cudaStream_t s, const char** name,
                                                          cudaError_t* err, int* launches);

// Fused value + first partial derivatives (dind_fast_gradient.cu): out[(point, comp, slot)], slot 0 = value.
template <typename T> bool try_eval_fast_gradient(const EvalParams<T>& P, cudaStream_t s, const char** name,
                                                  cudaError_t* err, int* launches);

// Two-stage separable evaluation of 4-D B-spline / NURBS grids (dind_fast_grid_split.cu).
template <typename T> bool grid_split_supported(const EvalParams<T>& P);
template <typename T> bool grid_split_slice_ok(const EvalParams<T>& P);   // then NURBS may pass the raw u (P.u_is_raw): u*w is formed in the kernel
template <typename T> size_t grid_split_scratch_bytes(const EvalParams<T>& P);
template <typename T> cudaError_t launch_grid_split(const EvalParams<T>& P, void* scratch, cudaStream_t s, const char** name, int* launches);

// Two-level un-permutation (dind_plan.cu): slot2[k] = staging slot of plan position k (slots grouped by the window of
// the destination, handed out per window in arrival order from a cursor: one atomic per point, no sort),
// dest2[j] = destination of slot j; plan_unbucket copies slot j home.  scratch: (n >> PLAN_WINDOW_SHIFT) + 1 counters.
cudaError_t plan_bucket_schedule(const uint32_t* perm, uint32_t* slot2, uint32_t* dest2, int64_t n, void* scratch,
                                 size_t* scratch_bytes, cudaStream_t s);
constexpr int PLAN_WINDOW_SHIFT = 11;   // windows of 2048 original points
// Un-permutation through a staging buffer in ORIGINAL order: the kernel wrote the record of plan position k to
// staged[perm[k]]; this copies record j to out[j + c * n], reads and writes coalesced.
template <typename T> cudaError_t plan_unstage(const T* staged, T* out, int64_t n, int ct, int rec_stride, cudaStream_t s);
template <typename T> cudaError_t plan_unbucket(const T* staged, const uint32_t* dest2, T* out, int64_t n, int ct, int rec_stride, cudaStream_t s);

// Specialised kernels.  Return true when they handled the evaluation (err holds the launch
// status), false when the configuration is outside their envelope.
template <typename T> bool try_eval_fast_unstructured(const EvalParams<T>& P, cudaStream_t s, const char** name,
                                                      cudaError_t* err, int* launches);
template <typename T> bool try_eval_fast_grid(const EvalParams<T>& P, cudaStream_t s, const char** name,
                                              cudaError_t* err, int* launches);

}  // namespace dind
