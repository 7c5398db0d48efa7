// Stand-alone stage-2 kernels (used when stage 1 cannot run the merge phase itself, i.e. when its
// grid is not guaranteed to be co-resident, or when the "fuse_merge" option is off).  The work is
// in merge.cuh.
#include "launch.cuh"
#include "merge.cuh"
#include "ptx.cuh"

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
    merge_item<MODE>(P, blockIdx.x, merge_smem, &bar, phase, CtaTeam());
}

// Stage 2 of a STREAMED query as a kernel of its own (merge.cuh:reduce_query): the FLUSH.  Back-to-back queries are
// merged by the merge warps of their successor's gather kernel; this kernel runs behind the last query of a burst
// and behind every synchronous (isolated) call.  kReduceThreads threads; shared memory = the scratch the query's
// merge was planned for (P.merge_smem).
template <int MODE>
__global__ void __launch_bounds__(kReduceThreads) reduce_kernel(const __grid_constant__ QueryParams P)
{
    extern __shared__ __align__(128) uint8_t merge_smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ int s_flag;
    grid_launch_dependents();
    if (threadIdx.x == 0) {
        BIGSI_TS(0);
        mbar_init(&bar, 1);
        fence_barrier_init();
        s_flag = ld_volatile_u64(P.abort_word) != 0ull;
    }
    __syncthreads();
    if (s_flag) return;      // the handle has aborted: nothing is reported any more
    grid_dependency_wait();  // the gather kernel has completed, its planes are visible
    if (threadIdx.x == 0) BIGSI_TS(8);
    reduce_query<MODE>(P, merge_smem, &bar, &s_flag, CtaTeam(), blockIdx.x, gridDim.x);
}

cudaError_t merge_kernels_init()
{
    cudaError_t e;
    // Every kernel asks for the LARGEST shared-memory carve-out: the driver otherwise sizes an SM's carve-out for the
    // kernel that gets there first, and the carve-out only changes when the SM is empty -- a small merge CTA on an SM
    // must not keep a 227 KB gather CTA of the next launch waiting for a reconfiguration.
#define BIGSI_SET_SMEM(K, BYTES)                                                                                   \
    e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BYTES));                       \
    if (e != cudaSuccess) return e;                                                                                \
    e = cudaFuncSetAttribute(K, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);  \
    if (e != cudaSuccess) return e;
    BIGSI_SET_SMEM((merge_kernel<kModeCounts>), kMergeKernelSmem)
    BIGSI_SET_SMEM((merge_kernel<kModeAnd>), kMergeKernelSmem)
    BIGSI_SET_SMEM((reduce_kernel<kModeCounts>), kReduceFatSmemBytes)
    BIGSI_SET_SMEM((reduce_kernel<kModeAnd>), kReduceFatSmemBytes)
#undef BIGSI_SET_SMEM
    return cudaSuccess;
}

// The flush of a streamed query: `grid` CTAs share the merge items.  debug_ts (if any) points behind the gather grid's stamps.
cudaError_t launch_reduce(const QueryParams &p, int mode, int grid, cudaStream_t stream)
{
    if (grid <= 0) return cudaErrorInvalidConfiguration;
    if (p.merge_smem > (uint32_t)kReduceFatSmemBytes) return cudaErrorInvalidConfiguration;
    const dim3 g((unsigned)grid), block(kReduceThreads);
    if (mode == kModeAnd) return launch_pdl(reduce_kernel<kModeAnd>, g, block, p.merge_smem, stream, p);
    return launch_pdl(reduce_kernel<kModeCounts>, g, block, p.merge_smem, stream, p);
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
