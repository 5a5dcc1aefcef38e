// Shared helpers for the meshcnn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/meshcnn_b200.h"

namespace mcnn {

extern unsigned long long g_launches;      // monotonically increasing launch counter (diagnostics only)
extern cudaError_t g_last_error;

inline int after_launch() {
    __atomic_add_fetch(&g_launches, 1ULL, __ATOMIC_RELAXED);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        g_last_error = e;
        (void)cudaGetLastError();
        return MCNN_E_CUDA;
    }
    return MCNN_OK;
}

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// torch.abs backward: sign with sign(0) == 0
__device__ __forceinline__ float sgn(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }

}  // namespace mcnn
