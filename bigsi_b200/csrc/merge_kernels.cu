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
    merge_item<MODE>(P, blockIdx.x, merge_smem, &bar, phase);
}

// Stage 2 of a STREAMED query (query_kernels.cu:gather_solo): merge + threshold + publication.  64 threads and
// 16 KB of shared memory per CTA, so that TWO of its CTAs fit beside a gather CTA (query.cuh:kReduceSlotsPerSm): it
// becomes resident while "its" gather kernel runs, sleeps in the dependency wait, and works while the NEXT query's
// gather kernel streams rows.
// The last CTA to finish publishes the hit list (host block and / or every shard's result blocks), waits -- bounded
// -- for the other shards' blocks of the same query, clears the state block of query seq + kStreamRing and advances
// the handle's completion word in query order.
// THREADS = kReduceThreads for back-to-back queries; an ISOLATED query (synchronous host call: nothing runs beside it)
// takes the fat variant, 256 threads and 64 KB, which merges in a quarter of the time.
template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS) reduce_kernel(const __grid_constant__ QueryParams P)
{
    extern __shared__ __align__(128) uint8_t merge_smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ int s_flag;
    grid_launch_dependents();
    if (threadIdx.x == 0) {
        BIGSI_TS(0);
        if (P.debug_ts) {
            unsigned int smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            P.debug_ts[(size_t)blockIdx.x * kDebugStamps + 15] = smid;
        }
        mbar_init(&bar, 1);
        fence_barrier_init();
        s_flag = ld_volatile_u64(P.abort_word) != 0ull;
        atomicAdd(&P.qstate->reduce_started, 1u);  // gather_solo's exit gate counts the resident reduce CTAs
    }
    __syncthreads();
    if (s_flag) return;      // the handle has aborted: nothing is reported any more
    grid_dependency_wait();  // the gather kernel has completed, its planes are visible
    if (threadIdx.x == 0) BIGSI_TS(8);
    uint32_t phase = 0;
    for (uint64_t item = blockIdx.x; item < P.merge_items; item += gridDim.x) merge_item<MODE>(P, item, merge_smem, &bar, phase);
    if (P.scrub_words) {
        // query front-end: its de-duplication table is dead once the gather kernel has read the k-mers; clear it
        // for the next query here, off every critical path
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.scrub_words; i += (uint64_t)gridDim.x * blockDim.x)
            P.scrub[i] = 0ull;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        BIGSI_TS(7);
        __threadfence();
        s_flag = atomicAdd(&P.qstate->reduce_arrivals, 1u) + 1u == gridDim.x;
    }
    __syncthreads();
    if (!s_flag) return;
    // ---- the last CTA of the query -------------------------------------------------------------------------
    __threadfence();
    if (P.n_sinks) publish_hits(P, P.sinks, P.n_sinks, P.sink_seq);
    // front-end words behind its table ({U}, {ticket, threshold}): every reader is done, re-arm them
    if (P.scrub_words && threadIdx.x < 2) P.scrub[P.scrub_words + threadIdx.x] = 0ull;
    if (threadIdx.x < P.n_gather) {  // all-gather: every shard's block of this query has arrived here
        const unsigned long long t0 = globaltimer_ns();
        const unsigned long long *blk = P.gather_blocks[threadIdx.x];
        bounded_wait(P.abort_word, P.host_abort, P.spin_timeout_ns, kAbortPeers, P.stream_seq,
                     [&]() { return ld_acquire_sys_u64(blk) == P.gather_seq; });
        atomicMax(&P.qstate->wait_ns, globaltimer_ns() - t0);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        BIGSI_TS(6);
        if (P.wait_ns_out) atomicAdd(P.wait_ns_out, P.qstate->wait_ns);
    }
    __syncthreads();
    if (threadIdx.x < sizeof(QState) / 8) reinterpret_cast<unsigned long long *>(P.qstate_next)[threadIdx.x] = 0ull;
    __syncthreads();
    if (threadIdx.x == 0) {
        // completion in query order: "done >= s" implies every query <= s is reduced and its ring slots are free
        bounded_wait(P.abort_word, P.host_abort, P.spin_timeout_ns, kAbortChain, P.stream_seq,
                     [&]() { return ld_acquire_gpu_u64(P.stream_done) + 1ull >= P.stream_seq; });
        __threadfence();
        st_release_gpu_u64(P.stream_done, P.stream_seq);
    }
}

cudaError_t merge_kernels_init()
{
    cudaError_t e = cudaFuncSetAttribute(merge_kernel<kModeCounts>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kMergeKernelSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(merge_kernel<kModeAnd>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMergeKernelSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(reduce_kernel<kModeCounts, kReduceThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kReduceSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(reduce_kernel<kModeAnd, kReduceThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kReduceSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(reduce_kernel<kModeCounts, kReduceFatThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kReduceFatSmemBytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(reduce_kernel<kModeAnd, kReduceFatThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kReduceFatSmemBytes);
}

// Streamed launch: the reduce kernel of a gather_solo launch.  debug_ts (if any) points behind the gather grid's stamps.
cudaError_t launch_reduce(const QueryParams &p, int mode, int grid, cudaStream_t stream)
{
    if (grid <= 0) return cudaErrorInvalidConfiguration;
    const dim3 g((unsigned)grid);
    if (p.merge_smem > (uint32_t)kReduceSmemBytes) {  // planned for the fat variant (isolated query)
        const dim3 block(kReduceFatThreads);
        if (mode == kModeAnd) return launch_pdl(reduce_kernel<kModeAnd, kReduceFatThreads>, g, block, p.merge_smem, stream, p);
        return launch_pdl(reduce_kernel<kModeCounts, kReduceFatThreads>, g, block, p.merge_smem, stream, p);
    }
    const dim3 block(kReduceThreads);
    if (mode == kModeAnd) return launch_pdl(reduce_kernel<kModeAnd, kReduceThreads>, g, block, p.merge_smem, stream, p);
    return launch_pdl(reduce_kernel<kModeCounts, kReduceThreads>, g, block, p.merge_smem, stream, p);
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
