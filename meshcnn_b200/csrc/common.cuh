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

// Programmatic dependent launch.  A training step is a chain of ~140 short dependent kernels; with the attribute below the
// next kernel of a stream is launched (CTAs scheduled, as far as the SM resources allow) while its predecessor still runs, and
// its threads block in pdl_enter() until the predecessor grid has completed and its memory is visible.  Every kernel calls
// pdl_enter() as its FIRST statement, so nothing is read or written early: only launch latency and CTA ramp-up overlap.
// Without the attribute both instructions are no-ops.  OFF by default: measured on the cubes step it changes nothing (2.71 ms
// with and without -- the replayed graph has no launch gaps worth hiding; the step is the sum of its kernels' durations).
// mcnn_debug_set_pdl(1) / MESHCNN_B200_PDL=1 switch it on.
extern int g_pdl;
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory");
}

// launch with a thread-block cluster of `cluster` CTAs (grid dimensions must be multiples of it)
template <class... KArgs, class... Args>
inline void launch_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, dim3 cluster, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster.x;
    attr[0].val.clusterDim.y = cluster.y;
    attr[0].val.clusterDim.z = cluster.z;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_pdl ? 2 : 1;
    (void)cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

template <class... KArgs, class... Args>
inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_pdl ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);     // errors are picked up by after_launch()
}

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
