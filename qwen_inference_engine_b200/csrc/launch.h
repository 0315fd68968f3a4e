// launch.h -- kernel launch helper with programmatic dependent launch (PDL).
// Every kernel of the decode step starts with pdl_wait() (griddepcontrol.wait: all
// prerequisite grids complete, their writes visible) before it touches global memory, and
// calls pdl_trigger() (griddepcontrol.launch_dependents) right away so the NEXT kernel's
// CTAs can be scheduled, set up their barriers/TMEM/tensor-map prefetch and -- for the
// GEMMs -- start streaming weights while this kernel is still running.  Launched without
// the attribute both instructions are no-ops, so the same kernels work eagerly.
#pragma once
#include <cuda_runtime.h>

#include <utility>

namespace qie {

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

extern bool g_use_pdl;  // engine.cu; set per process from QIE_PDL (default on)

// Function attributes (opt-in dynamic shared memory, cluster size) are per DEVICE: a process that creates engines on
// several devices must set them on each.  One flag per device ordinal instead of a process-wide `static bool`.
struct PerDeviceOnce {
  unsigned long long mask[4] = {0ull, 0ull, 0ull, 0ull};
  bool need() const {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 256) return true;
    return ((mask[d >> 6] >> (d & 63)) & 1ull) == 0ull;
  }
  void done() {
    int d = 0;
    if (cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < 256) mask[d >> 6] |= 1ull << (d & 63);
  }
};

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace qie
