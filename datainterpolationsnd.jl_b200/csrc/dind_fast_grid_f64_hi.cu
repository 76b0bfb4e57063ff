// grid pencil kernels: double, 4 <= N_in <= 5 (see dind_fast_grid.cuh)
#include "dind_fast_grid.cuh"
namespace dind {
bool try_grid_f64_hi(const EvalParams<double>& P, cudaStream_t s, const char** name, cudaError_t* err, int* launches) {
    return try_grid_range<double, 4, 5>(P, s, name, err, launches);
}
}  // namespace dind
