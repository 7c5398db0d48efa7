// Kernel launch helpers.
//
// Programmatic dependent launch (PDL): consecutive kernels of a query stream are launched with the
// programmatic-stream-serialization attribute, so a kernel's launch and prologue overlap its
// predecessor's tail; every kernel calls grid_dependency_wait() before it touches anything its
// predecessor produces (or still reads).
//
// Cooperative launch: the fused query kernel runs its merge phase behind a grid-wide barrier, which
// needs every CTA resident at once; the cooperative attribute makes the driver check that.
#pragma once
#include <cuda_runtime.h>

namespace bigsi {

// blocks until the preceding grid in the stream has completed and its writes are visible
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// lets the next grid in the stream start launching (it still waits for our completion)
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                             bool cooperative, Args &&...args)
{
    static bool coop_with_pdl_ok = true;  // some drivers refuse the combination: fall back to cooperative only
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    for (int attempt = 0; attempt < 2; ++attempt) {
        unsigned n = 0;
        if (cooperative) {
            attr[n].id = cudaLaunchAttributeCooperative;
            attr[n].val.cooperative = 1;
            ++n;
        }
        if (pdl && (!cooperative || coop_with_pdl_ok)) {
            attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[n].val.programmaticStreamSerializationAllowed = 1;
            ++n;
        }
        cfg.attrs = attr;
        cfg.numAttrs = n;
        const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
        if (e == cudaSuccess || !(cooperative && pdl && coop_with_pdl_ok)) return e;
        (void)cudaGetLastError();
        coop_with_pdl_ok = false;  // retry once without PDL
    }
    return cudaErrorUnknown;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args &&...args)
{
    return launch_ex(kernel, grid, block, smem, stream, true, false, static_cast<Args &&>(args)...);
}

}  // namespace bigsi
