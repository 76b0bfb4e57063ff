// placeholder: specialised unstructured kernels (filled in below in later commits)
#include "dind_kernels.h"
namespace dind {
template <typename T>
bool try_eval_fast_unstructured(const EvalParams<T>&, cudaStream_t, const char**, cudaError_t*, int*) { return false; }
template bool try_eval_fast_unstructured<float>(const EvalParams<float>&, cudaStream_t, const char**, cudaError_t*, int*);
template bool try_eval_fast_unstructured<double>(const EvalParams<double>&, cudaStream_t, const char**, cudaError_t*, int*);
}
