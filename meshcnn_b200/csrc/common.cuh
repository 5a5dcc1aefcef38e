// Shared helpers for the meshcnn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/meshcnn_b200.h"
#include "../../include/meshcnn_b200_debug.h"

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

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: one guard per launch site remembers what
// has been configured on each device, so a process that drives several GPUs (or several host threads) configures every one.
// Racing threads may both set the attribute; that is idempotent.
constexpr int MCNN_MAX_DEVICES = 64;
struct DynSmemGuard {
    size_t per_dev[MCNN_MAX_DEVICES] = {};
    template <typename F>
    int ensure(F* func, size_t bytes) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MCNN_MAX_DEVICES) dev = 0;
        if (__atomic_load_n(&per_dev[dev], __ATOMIC_ACQUIRE) >= bytes) return MCNN_OK;
        cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) { g_last_error = e; return MCNN_E_CUDA; }
        __atomic_store_n(&per_dev[dev], bytes, __ATOMIC_RELEASE);
        return MCNN_OK;
    }
};

// SM count of the CURRENT device (cached per device)
inline int sm_count() {
    static int cached[MCNN_MAX_DEVICES] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MCNN_MAX_DEVICES) dev = 0;
    int n = __atomic_load_n(&cached[dev], __ATOMIC_RELAXED);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        __atomic_store_n(&cached[dev], n, __ATOMIC_RELAXED);
    }
    return n;
}

// torch.abs backward: sign with sign(0) == 0
__device__ __forceinline__ float sgn(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }

}  // namespace mcnn
