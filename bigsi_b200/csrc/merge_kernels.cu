// Stand-alone stage-2 kernels (used when stage 1 cannot run the merge phase itself, i.e. when its
// grid is not guaranteed to be co-resident, or when the "fuse_merge" option is off).  The work is
// in merge.cuh.
#include "launch.cuh"
#include "merge.cuh"

namespace bigsi {

constexpr int kMergeThreads = 256;

template <int NG>
__global__ void __launch_bounds__(kMergeThreads) merge_counts_kernel(const __grid_constant__ QueryParams P,
                                                                    const uint32_t gpi)
{
    __shared__ __align__(16) uint8_t smem[kMergeSmemBytes];
    grid_dependency_wait();  // stage 1 must have completed (PDL launch)
    merge_counts_item<NG>(P, blockIdx.x, gpi, smem);
}

__global__ void __launch_bounds__(kMergeThreads) merge_and_kernel(const __grid_constant__ QueryParams P)
{
    __shared__ __align__(16) uint8_t smem[1024];
    grid_dependency_wait();
    merge_and_item(P, blockIdx.x, smem);
}

cudaError_t launch_merge(const QueryParams &p, int mode, cudaStream_t stream)
{
    if (p.n_queries == 0) return cudaSuccess;
    const MergePlan m = plan_merge(p, mode, kMergeThreads);
    if (m.n_items == 0) return cudaSuccess;
    if (m.n_items > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    const dim3 grid((unsigned)m.n_items), block(kMergeThreads);
    if (mode == kModeAnd) return launch_pdl(merge_and_kernel, grid, block, 0, stream, p);
    if (m.ng == 1) return launch_pdl(merge_counts_kernel<1>, grid, block, 0, stream, p, m.gpi);
    return launch_pdl(merge_counts_kernel<4>, grid, block, 0, stream, p, m.gpi);
}

}  // namespace bigsi
