// Programmatic dependent launch (PDL): a kernel of the pipeline may be made resident while its predecessor drains.
// Device side, every kernel launched through launch_pdl() calls pdl_trigger() at its top (lets the NEXT kernel start
// launching once all our CTAs have started) and pdl_wait() before it first touches data written by earlier kernels
// (returns when the predecessor grid has completed and its writes are visible).  What runs before pdl_wait() may
// only read constants (weights) and set up shared memory / TMEM.  Ordering is transitive: a grid cannot complete
// before its own pdl_wait() has returned.  (-DTMPNN_DEV builds: TMPNN_PDL=0 turns the launch attribute off, for A/B timing.)
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>
#include <utility>

namespace tmpnn {

__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

inline bool pdl_enabled() {
#ifdef TMPNN_DEV
    static const bool on = [] { const char* e = std::getenv("TMPNN_PDL"); return !(e && std::atoi(e) == 0); }();
    return on;
#else
    return true;
#endif
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace tmpnn
