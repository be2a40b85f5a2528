// pdl.cuh — programmatic dependent launch (sm_90+) for the engine's single-stream step chain.
// Every kernel of the hot path is launched with cudaLaunchAttributeProgrammaticStreamSerialization: the grid of step k+1
// may become resident while step k is still draining, runs whatever does not depend on step k (launch latency, CTA
// scheduling, barrier / TMEM / descriptor set-up, loads of static weights), and blocks in pdl_wait() until step k has
// completed and its memory is visible.  Rules that keep this race-free:
//   * a kernel launched through pdl_launch() executes pdl_wait() before it reads or writes ANY memory that another
//     kernel of the stream produces or consumes (activations, image tables, scratch) — by default its first statement;
//   * pdl_trigger() comes after pdl_wait(), so the pre-wait part of step k+1 can only overlap step k, never step k-1.
// VSE_PDL=0 in the environment launches everything fully serialised (A/B switch, debugging).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>

namespace vse {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("VSE_PDL"); return !(e && e[0] == '0'); }();
    return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t pdl_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace vse
