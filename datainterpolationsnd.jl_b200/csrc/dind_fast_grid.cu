// eval_grid! fast-path dispatch: forwards to the per-type / per-dimension translation units.
#include "dind_kernels.h"
namespace dind {
bool try_grid_f32_lo(const EvalParams<float>&, cudaStream_t, const char**, cudaError_t*, int*);
bool try_grid_f32_hi(const EvalParams<float>&, cudaStream_t, const char**, cudaError_t*, int*);
bool try_grid_f64_lo(const EvalParams<double>&, cudaStream_t, const char**, cudaError_t*, int*);
bool try_grid_f64_hi(const EvalParams<double>&, cudaStream_t, const char**, cudaError_t*, int*);

template <>
bool try_eval_fast_grid<float>(const EvalParams<float>& P, cudaStream_t s, const char** name, cudaError_t* err, int* launches) {
    return try_grid_f32_lo(P, s, name, err, launches) || try_grid_f32_hi(P, s, name, err, launches);
}
template <>
bool try_eval_fast_grid<double>(const EvalParams<double>& P, cudaStream_t s, const char** name, cudaError_t* err, int* launches) {
    return try_grid_f64_lo(P, s, name, err, launches) || try_grid_f64_hi(P, s, name, err, launches);
}
}  // namespace dind
