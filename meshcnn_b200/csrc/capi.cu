// Library-level entry points and the two diagnostics globals.
#include <stdlib.h>

#include "common.cuh"

namespace mcnn {
unsigned long long g_launches = 0;
cudaError_t g_last_error = cudaSuccess;
static int pdl_default() {
    const char* e = getenv("MESHCNN_B200_PDL");
    return (e != nullptr && e[0] == '1') ? 1 : 0;
}
int g_pdl = pdl_default();
}  // namespace mcnn

extern "C" int mcnn_version(void) { return 100; }

extern "C" unsigned long long mcnn_launch_count(void) { return __atomic_load_n(&mcnn::g_launches, __ATOMIC_RELAXED); }

extern "C" const char* mcnn_last_cuda_error(void) { return cudaGetErrorString(mcnn::g_last_error); }

extern "C" int mcnn_debug_set_pdl(int on) { mcnn::g_pdl = on ? 1 : 0; return MCNN_OK; }
