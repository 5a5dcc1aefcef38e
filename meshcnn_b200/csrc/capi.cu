// Library-level entry points and the two diagnostics globals.
#include "common.cuh"

namespace mcnn {
unsigned long long g_launches = 0;
cudaError_t g_last_error = cudaSuccess;
}  // namespace mcnn

extern "C" int mcnn_version(void) { return 100; }

extern "C" unsigned long long mcnn_launch_count(void) { return __atomic_load_n(&mcnn::g_launches, __ATOMIC_RELAXED); }

extern "C" const char* mcnn_last_cuda_error(void) { return cudaGetErrorString(mcnn::g_last_error); }
