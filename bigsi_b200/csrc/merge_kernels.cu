// Stand-alone stage-2 kernels (used when stage 1 cannot run the merge phase itself, i.e. when its
// grid is not guaranteed to be co-resident, or when the "fuse_merge" option is off).  The work is
// in merge.cuh.
#include "launch.cuh"
#include "merge.cuh"

namespace bigsi {

template <int MODE>
__global__ void __launch_bounds__(kMergeKernelThreads) merge_kernel(const __grid_constant__ QueryParams P)
{
    extern __shared__ __align__(128) uint8_t merge_smem[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    grid_dependency_wait();  // stage 1 must have completed (PDL launch)
    uint32_t phase = 0;
    merge_item<MODE>(P, blockIdx.x, merge_smem, &bar, phase);
}

cudaError_t merge_kernels_init()
{
    cudaError_t e = cudaFuncSetAttribute(merge_kernel<kModeCounts>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kMergeKernelSmem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(merge_kernel<kModeAnd>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMergeKernelSmem);
}

cudaError_t launch_merge(const QueryParams &p_in, int mode, cudaStream_t stream)
{
    QueryParams p = p_in;
    p.debug_ts = nullptr;  // the timeline buffer is sized for stage 1's grid
    if (p.n_queries == 0 || p.merge_items == 0) return cudaSuccess;
    if (p.merge_items > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    const dim3 grid((unsigned)p.merge_items), block(kMergeKernelThreads);
    if (mode == kModeAnd) return launch_pdl(merge_kernel<kModeAnd>, grid, block, p.merge_smem, stream, p);
    return launch_pdl(merge_kernel<kModeCounts>, grid, block, p.merge_smem, stream, p);
}

}  // namespace bigsi
