// Kernel launch helper: programmatic dependent launch (PDL).  Consecutive kernels of one query
// (hash -> fused gather -> merge) are launched with the programmatic-stream-serialization
// attribute, so a kernel's launch and prologue overlap its predecessor's tail; every kernel calls
// grid_dependency_wait() before it touches anything its predecessor produces (or still reads).
#pragma once
#include <cuda_runtime.h>

namespace bigsi {

// blocks until the preceding grid in the stream has completed and its writes are visible
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// lets the next grid in the stream start launching (it still waits for our completion)
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace bigsi
