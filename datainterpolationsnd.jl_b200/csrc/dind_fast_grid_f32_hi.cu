// grid pencil kernels: float, 4 <= N_in <= 5 (see dind_fast_grid.cuh)
#include "dind_fast_grid.cuh"
namespace dind {
bool try_grid_f32_hi(const EvalParams<float>& P, cudaStream_t s, const char** name, cudaError_t* err, int* launches) {
    return try_grid_range<float, 4, 5>(P, s, name, err, launches);
}
}  // namespace dind
