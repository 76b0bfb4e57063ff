// grid pencil kernels: double, 2 <= N_in <= 3 (see dind_fast_grid.cuh)
#include "dind_fast_grid.cuh"
namespace dind {
bool try_grid_f64_lo(const EvalParams<double>& P, cudaStream_t s, const char** name, cudaError_t* err, int* launches) {
    return try_grid_range<double, 2, 3>(P, s, name, err, launches);
}
}  // namespace dind
