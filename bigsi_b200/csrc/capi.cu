// C ABI of libbigsi_b200.so (include/bigsi_b200.h): index lifecycle, launch planning and the
// host-buffer entry points.  No CPU fallback: every compute entry point needs a CUDA device.
#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/bigsi_b200.h"
#include "hash.cuh"
#include "merge.cuh"
#include "query.cuh"

using namespace bigsi;

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
int fail_cuda(cudaError_t e, const char *what)
{
    const int code = e == cudaErrorMemoryAllocation
                         ? BIGSI_B200_ERR_OOM
                         : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? BIGSI_B200_ERR_NO_DEVICE
                                                                                        : BIGSI_B200_ERR_CUDA;
    (void)cudaGetLastError();  // clear the sticky-free error state
    return fail(code, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
}
#define CK(expr)                                          \
    do {                                                  \
        cudaError_t _e = (expr);                          \
        if (_e != cudaSuccess) return fail_cuda(_e, #expr); \
    } while (0)

uint32_t bits_of(uint64_t v)
{
    uint32_t b = 0;
    while (v) {
        ++b;
        v >>= 1;
    }
    return b;
}
uint64_t round_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

struct DevBuf {
    void *p = nullptr;
    uint64_t cap = 0;
    cudaError_t reserve(uint64_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) {
            cudaError_t e = cudaFree(p);
            p = nullptr;
            cap = 0;
            if (e != cudaSuccess) return e;
        }
        const uint64_t want = round_up(bytes + bytes / 4, 256);
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            e = cudaMalloc(&p, round_up(bytes, 256));
            if (e != cudaSuccess) return e;
            cap = round_up(bytes, 256);
            return cudaSuccess;
        }
        cap = want;
        return cudaSuccess;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct PinnedBuf {
    void *p = nullptr;
    uint64_t cap = 0;
    cudaError_t reserve(uint64_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaHostAlloc(&p, round_up(bytes, 4096), cudaHostAllocMapped | cudaHostAllocPortable);
        if (e == cudaSuccess) {
            cap = round_up(bytes, 4096);
            // result blocks live here and are recognised by a per-handle sequence number in their first word: a
            // recycled allocation must not carry a block some earlier handle published (same numbers, other index)
            memset(p, 0, cap);
        }
        return e;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

struct DeviceGuard {
    int prev = -1;
    cudaError_t err;
    explicit DeviceGuard(int dev)
    {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

struct TimedLaunch {
    cudaEvent_t e0, e1, e2;
};

}  // namespace

// Column-sharded search without per-query collectives: every rank owns one device block
//   [kExInboxes low-latency inboxes][result blocks: kExGenerations generations x world x block]
// that its peers map (CUDA IPC across processes, plain peer access inside one process).  Rank 0's gather
// kernel pushes the query into the peers' inboxes (LL lines); every rank's stage 2 (merge team / flush) publishes its hit
// list into slot `rank` of every rank's result blocks and waits (bounded) for the others' slots of the same
// query while the next query's gather kernel already runs.  Query s uses inbox s % kExInboxes and result
// generation s % kExGenerations.  Reuse is safe because of the streamed path's entry gate: the gather kernel of
// query s starts only after THIS rank's stage 2 of query s - kStreamRing has seen every shard's block of
// that query, i.e. after every shard has finished reading the inbox of s - kStreamRing and has launched past
// its consumers of generations <= s - kExGenerations.
struct Exchange {
    int world = 0, rank = 0;
    uint32_t spec = 0;
    uint64_t max_kmer_bytes = 0, kmers_stride = 0, ll_off = 0, sinks_off = 0, block_bytes = 0, total_bytes = 0;
    uint8_t *local = nullptr;
    uint8_t *peer[kMaxSinks] = {};
    bool ipc_opened[kMaxSinks] = {};
    bool ready = false;
    uint64_t seq = 0;            // queries launched on this shard
    // optional host results (bigsi_b200_exchange_host_results): kStreamStates blocks of mapped host memory, query s ->
    // block s % kStreamStates = [u64 seq][u64 pad][world x block_bytes]
    PinnedBuf h_gather;
    uint64_t h_gather_block = 0;
    bool h_gather_on = false;    // searches launched while it is on also land in the host blocks
    uint64_t h_gather_seq[kStreamStates] = {};  // host block b holds (or will hold) the result of this search number
};
constexpr uint64_t kExInboxes = kStreamRing, kExGenerations = 2 * kStreamRing;

struct bigsi_b200_index {
    int device = 0;
    Exchange ex;
    int sm_count = 0;
    uint64_t num_rows = 0, num_cols = 0, col_capacity = 0, col_offset = 0, pitch = 0;
    uint8_t *matrix = nullptr;
    cudaStream_t stream = nullptr;
    // options
    int64_t opt_tile_bytes = 0, opt_grid = 0, opt_kmers_per_stage = 0, opt_n_stages = 0, opt_ctas_per_sm = 0;
    bool timing = false;
    int64_t opt_debug_flags = 0;
    int64_t opt_prehash = 1, opt_fuse_merge = 1, opt_merge_chunk_bytes = 0, opt_solo = 1, opt_pool_pct = -1, opt_zero_copy = 1, opt_cooperative = 1;
    int64_t opt_inputs_ready = 0, opt_spin_timeout_ms = 10000, opt_defer = 1, opt_direct = 1, opt_batch_reuse = 1, opt_self_merge = 0, opt_push_repeat = 0, opt_push_all_warps = 0;
    // streamed single-query launches (query.cuh:kStreamRing): ring-buffered scratch + the completion / abort words
    DevBuf d_pool;            // kStreamRing x [ready flags: grid x u64][ids: grid x pool_share x h x i32]
    uint64_t pool_slot_bytes = 0;
    uint64_t pool_epoch = 0;
    DevBuf stream_partial;    // kStreamRing x partial planes
    uint64_t stream_partial_slot = 0;
    DevBuf d_stream;          // [done: 64 B][abort: 64 B][wait_ns: 64 B][kStreamStates x QState]
    PinnedBuf h_status;       // mapped: word 0 = mirror of the abort word
    DevBuf d_hits_ring;       // host-buffer paths: kStreamStates x {n_hits, cols[cap], counts[cap]}
    uint64_t stream_seq = 0;  // streamed queries launched on this handle
    // the last streamed query when its stage 2 has not been launched yet (deferred launches): the next streamed
    // launch hands it to its merge team, flush_pending() launches reduce_kernel for it
    struct PendingQuery {
        bool have = false;
        QueryParams p;
        int mode = 0, reduce_grid = 0;
        cudaStream_t stream = nullptr;
        uint64_t ticket = 0;  // sequence searches: the ticket waiting for it
    } pending;
    // sequence searches (front-end inside the gather kernel): kStreamRing de-duplication tables with epoch-tagged
    // entries, up to kSeqTickets searches in flight (submit / wait), each with its own pinned sequence buffer and
    // its own result block in mapped host memory
    DevBuf d_seq_tables;
    uint64_t seq_table_entries = 0;               // entries of ONE table
    uint64_t seq_table_uses[kStreamRing] = {};
    struct SeqTicket {
        uint64_t id = 0;
        bool pending = false, deferred = false;
        std::string seq;                          // deferred (not streamable): searched synchronously at wait()
        int k = 0, h = 0;
        double threshold = 0;
        uint64_t cap = 0, spec = 0;
        const uint8_t *d_hits = nullptr;          // device hit buffers of this search (long hit lists)
    };
    static constexpr int kSeqTickets = 8;
    SeqTicket tickets[kSeqTickets];
    PinnedBuf h_seq[kSeqTickets];
    PinnedBuf h_tsink;
    uint64_t next_ticket = 1;
    DevBuf d_barrier;                             // grid-barrier arrival counter of the fused kernel
    uint64_t barrier_target = 0, done_target = 0;
    PinnedBuf h_sink, h_kmers;   // mapped pinned: result block the kernel publishes to / staging of pageable k-mers
    uint64_t sink_seq = 0;
    // scratch
    DevBuf debug_ts;
    DevBuf d_seq, d_table;   // query front-end: sequence bytes, de-duplication table (+ counter, ticket, threshold)
    uint64_t table_clean_bytes = 0;  // leading bytes of d_table known to be zero
    DevBuf partial, d_kmers, d_rows, d_qoff, d_out, d_min, d_nhits, d_bloom, d_planted;
    DevBuf d_reuse, d_reuse_rows;  // batch reuse: de-duplication workspace, the gathered AND vectors of the distinct k-mers
    PinnedBuf h_small;
    // timing
    std::vector<TimedLaunch> timed_free, timed_used;
    // statistics
    bigsi_b200_info stats{};
    uint64_t kernel_launches = 0;
};

namespace {

constexpr uint64_t kReuseMinKmers = 16384;  // smaller batches are not worth the de-duplication pass and its host round trip
constexpr uint64_t kPoolFlagEntries = 16384, kPoolFlagBytes = kPoolFlagEntries * 8;
constexpr uint64_t kStreamStateBytes = 3 * 64 + (uint64_t)kStreamStates * sizeof(QState);
// shared memory of a streamed gather CTA in COUNTS mode: the merge team's scratch lies behind the ring
constexpr uint64_t kStreamTeamSmem = kReduceSmemBytes;

bool g_kernels_ready[64] = {};  // function attributes (dynamic shared memory opt-in) are per device

int ensure_kernels()
{
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(BIGSI_B200_ERR_INVALID, "device ordinal %d not supported", dev);
    if (g_kernels_ready[dev]) return 0;
    CK(query_kernels_init());
    g_kernels_ready[dev] = true;
    return 0;
}

// Launch plan of one query batch (DESIGN.md "Launch planning").  `have_kmers`: the caller holds
// raw k-mers (length k), so the kernel may hash them itself when the geometry allows.
int plan_query(bigsi_b200_index *ix, int mode, uint64_t n_queries, uint64_t total_kmers, uint64_t max_query_kmers,
               int h, bool have_kmers, int k, QueryParams &p, int &grid, bool isolated = false)
{
    if (h < 1 || h > kMaxH) return fail(BIGSI_B200_ERR_INVALID, "h=%d out of range [1,%d]", h, kMaxH);
    if (n_queries > 0xffffffffull) return fail(BIGSI_B200_ERR_INVALID, "too many queries");
    memset(&p, 0, sizeof p);
    const uint64_t row_bytes = (ix->num_cols + 7) / 8;
    const uint64_t smem_avail = (uint64_t)(kSmemBudget - kSmemHeaderBytes);
    p.matrix = ix->matrix;
    p.pitch = ix->pitch;
    p.n_queries = (uint32_t)n_queries;
    p.h = (uint32_t)h;
    p.total_kmers = total_kmers;
    p.num_cols = (uint32_t)ix->num_cols;
    p.num_rows = (uint32_t)ix->num_rows;
    p.mod_magic = mod_magic((uint32_t)ix->num_rows);
    p.row_bytes16 = (uint32_t)round_up(row_bytes ? row_bytes : 1, 16);

    // column tile: as wide as the consumer warps allow, but >= 3 one-k-mer stages must fit in smem
    uint64_t max_tile = smem_avail / (3ull * h) / 16 * 16;
    if (max_tile > (uint64_t)kMaxTileBytes) max_tile = kMaxTileBytes;
    if (max_tile < 16) return fail(BIGSI_B200_ERR_INVALID, "h=%d too large for the shared-memory ring", h);
    uint64_t tile = 0;
    if (ix->opt_tile_bytes > 0) {
        tile = round_up((uint64_t)ix->opt_tile_bytes, 16);
        if (tile > max_tile) tile = max_tile;
    } else {
        const uint64_t nt = (p.row_bytes16 + max_tile - 1) / max_tile;
        tile = (p.row_bytes16 + nt - 1) / nt;
        const uint64_t t128 = round_up(tile, 128);
        tile = t128 <= max_tile ? t128 : round_up(tile, 16);
    }
    if (tile > p.row_bytes16) tile = p.row_bytes16;
    p.tile_bytes = (uint32_t)tile;
    p.n_tiles = (uint32_t)((p.row_bytes16 + tile - 1) / tile);

    // slices
    p.total_items = (uint64_t)p.n_tiles * total_kmers;
    const uint64_t ctas =
        ix->opt_grid > 0 ? (uint64_t)ix->opt_grid
                         : (uint64_t)ix->sm_count * (uint64_t)(ix->opt_ctas_per_sm > 0 ? ix->opt_ctas_per_sm : 1);
    if (p.total_items == 0) {
        p.items_per_slice = 1;
        p.n_slices = 0;
        p.slices_per_cta = 1;
        grid = 0;
    } else {
        const uint64_t spc = (p.total_items + ctas * kMaxSliceItems - 1) / (ctas * kMaxSliceItems);
        uint64_t ips = (p.total_items + ctas * spc - 1) / (ctas * spc);
        if (ips > kMaxSliceItems) ips = kMaxSliceItems;
        if (ips < 1) ips = 1;
        const uint64_t n_slices = (p.total_items + ips - 1) / ips;
        if (n_slices > 0xffffffffull) return fail(BIGSI_B200_ERR_INVALID, "query batch too large");
        p.items_per_slice = (uint32_t)ips;
        p.n_slices = (uint32_t)n_slices;
        p.slices_per_cta = (uint32_t)((n_slices + ctas - 1) / ctas);
        grid = (int)((n_slices + p.slices_per_cta - 1) / p.slices_per_cta);
    }
    const uint64_t longest = max_query_kmers ? (max_query_kmers < total_kmers ? max_query_kmers : total_kmers) : total_kmers;
    if (mode == BIGSI_B200_MODE_COUNTS && longest > 0xffffffffull)
        return fail(BIGSI_B200_ERR_INVALID, "a query longer than 2^32-1 k-mers does not fit uint32 counts");
    p.max_query_kmers = longest;
    if (longest / p.items_per_slice + 2 > 65535)
        return fail(BIGSI_B200_ERR_INVALID, "a query spans more than 65 535 slices; split the batch");
    p.total_planes = mode == BIGSI_B200_MODE_COUNTS ? (bits_of(longest) ? bits_of(longest) : 1) : 1;
    const uint64_t seg_max = longest < p.items_per_slice ? longest : p.items_per_slice;
    p.planes_per_slot = mode == BIGSI_B200_MODE_COUNTS ? (bits_of(seg_max) ? bits_of(seg_max) : 1) : 1;

    // in-kernel hashing: every CTA hashes its own contiguous k-mers in the prologue (one tile, one
    // slice per CTA, and the id table + hashing scratch must fit beside a useful ring)
    p.prehash = 0;
    p.ids_bytes = 0;
    p.ids_table_bytes = 0;
    if (have_kmers && grid > 0 && ix->opt_prehash != 0 && p.n_tiles == 1 && p.slices_per_cta == 1 && k >= 1) {
        const uint64_t table = round_up((uint64_t)p.items_per_slice * h * 4, 128);
        const uint64_t scratch = round_up(hash_scratch_bytes(p.items_per_slice, (uint32_t)k) + 256, 128);
        if (table <= 16384 && table + scratch + 3ull * h * tile <= smem_avail) {
            p.prehash = 1;
            p.ids_table_bytes = (uint32_t)table;
            p.ids_bytes = (uint32_t)(table + scratch);
            p.k = (uint32_t)k;
        }
    }

    // solo = STREAMED path: one query, hashed in the kernel, one slice per CTA; stage 2 (merge, threshold, publication)
    // is left to the merge team of the next streamed launch or to the flush kernel (no grid barrier).  In COUNTS mode
    // the team's scratch lies behind the ring; an isolated query pools a share of every CTA's k-mers (tail balance).
    p.solo = p.stream = 0;
    p.pool_share = 0;
    p.solo_max_kmers = 0xffffffffu;
    const uint64_t team_smem = mode == BIGSI_B200_MODE_COUNTS ? kStreamTeamSmem : 0;
    if (p.prehash && n_queries == 1 && ix->opt_solo != 0 && grid > 0 &&
        kSmemHeaderBytes + p.ids_bytes + 3ull * h * tile + team_smem <= (uint64_t)kSmemBudget)
        p.solo = p.stream = 1;

    // ring geometry
    const uint64_t ring_avail = smem_avail - p.ids_bytes - (p.solo ? team_smem : 0);
    const uint64_t kmer_bytes = (uint64_t)h * tile;
    uint32_t G = 1;
    if (ix->opt_kmers_per_stage > 0) {
        G = (uint32_t)ix->opt_kmers_per_stage;
    } else {
        // A ring slot costs the producer warp ~0.4 us whatever it holds (empty-barrier wait, expect_tx, the serialised
        // bulk-copy issue), which caps an SM at slot bytes / 0.4 us: 18.7 KB slots (one k-mer at h = 3, N = 50 000) = 47 GB/s
        // per SM -- enough when all 148 SMs gather all the time (6.3 TB/s), but a streamed query's SMs spend 20 % of their
        // time in prologue / flush / hand-over and the others could not make up for it: 34.4 us per query; two k-mers per
        // slot (37.5 KB, 4 stages): 29.9 us = 6.27 TB/s.  (h = 1 rows of 6 KB: 2 per slot 3.6 TB/s, 4 per slot 6 TB/s.)
        for (uint32_t g = 8; g > 1; g >>= 1)
            if (g * kmer_bytes <= 40960) {
                G = g;
                break;
            }
    }
    while (G > 1 && 3ull * G * kmer_bytes > ring_avail) G >>= 1;
    uint64_t stages = ring_avail / (G * kmer_bytes);
    if (stages > (uint64_t)kMaxStages) stages = kMaxStages;
    if (ix->opt_n_stages > 0 && (uint64_t)ix->opt_n_stages < stages) stages = (uint64_t)ix->opt_n_stages;
    if (stages < 2) return fail(BIGSI_B200_ERR_INVALID, "ring does not fit (h=%d tile=%llu)", h, (unsigned long long)tile);
    p.kmers_per_stage = G;
    p.n_stages = (uint32_t)stages;

    // in-kernel merge of the generic path: needs every CTA resident at once (one CTA per SM, grid <= SM count;
    // launched cooperatively, so the driver verifies it)
    p.fuse_merge = 0;
    if (!p.solo && ix->opt_fuse_merge != 0 && grid > 0 && grid <= ix->sm_count &&
        kSmemHeaderBytes + p.ids_bytes + (uint64_t)kMergeScratchBytes <= (uint64_t)kSmemBudget)
        p.fuse_merge = 1;
    if (p.solo) {
        // The pool (a share of every CTA's k-mers claimed dynamically) evens out the CTAs' finish times when they all
        // start together, i.e. for an ISOLATED query.  Back-to-back streamed queries never use it: their CTAs start
        // whenever an SM becomes free, a claimer would wait for owners that have not even started, and there is no
        // common finish line to balance for (the SM simply goes on with the next query).
        const uint64_t c = p.items_per_slice;
        const uint64_t pct = ix->opt_pool_pct >= 0 ? (uint64_t)ix->opt_pool_pct : (isolated ? 12u : 0u);
        uint64_t pp = (c >= 16 && h <= kPoolMaxH && (uint64_t)grid <= kPoolFlagEntries) ? (c * pct + 50) / 100 : 0;
        if (pp > c) pp = c;
        if (mode == BIGSI_B200_MODE_COUNTS && pp) {
            uint32_t pps = bits_of(c + 2 * pp);
            if (pps > (uint32_t)kSegPlanes) {  // cannot widen the segment counters: keep the static split
                pp = 0;
            } else {
                p.planes_per_slot = pps;
                p.solo_max_kmers = (1u << pps) - 1;
            }
        }
        p.pool_share = (uint32_t)pp;
        p.merge_team = query_has_merge_team(p, mode) ? (uint32_t)kMergeTeamThreads : 0u;
    }
    // merge geometry; it fixes the chunk-major layout of the partial planes, so stage 1 needs it as well
    p.n_slots_total = query_n_slots(p);
    if (p.fuse_merge) {
        const uint32_t scratch = query_smem_bytes(p) - kSmemHeaderBytes - p.ids_bytes;  // the drained ring
        plan_merge(p, mode, scratch, (uint64_t)grid, (uint32_t)ix->opt_merge_chunk_bytes);
    } else if (p.stream) {
        // back-to-back queries: the scratch of a gather CTA's merge team; an isolated query is always merged by the
        // flush kernel, with more scratch (fewer, larger slot batches)
        plan_merge(p, mode, (isolated && !p.merge_team) ? kReduceFatSmemBytes : kReduceSmemBytes, (uint64_t)ix->sm_count,
                   (uint32_t)ix->opt_merge_chunk_bytes);
    } else {
        plan_merge(p, mode, kMergeKernelSmem, (uint64_t)ix->sm_count * 3, (uint32_t)ix->opt_merge_chunk_bytes);
    }
    return 0;
}

struct HitsOut {
    const uint32_t *min_kmers = nullptr;
    int32_t *cols = nullptr;
    uint32_t *counts = nullptr;
    unsigned long long *n = nullptr;
    uint64_t cap = 0;
    // single-query extras: threshold by value (no device array) and result publication by the kernel itself to
    // host / peer sinks (query.cuh:QueryParams)
    bool by_value = false;
    uint32_t min_value = 0;
    uint32_t n_sinks = 0, sink_spec = 0;
    unsigned long long *sinks[kMaxSinks] = {};
    unsigned long long sink_seq = 0;
    bool *published = nullptr;  // set when the launch will publish to the sinks
    // column-sharded exchange fused into the kernels (see Exchange above)
    bool isolated = false;        // the caller waits for this query's result before it issues the next one: nothing overlaps
                                  // it, so the CTAs start together and a pooled tail balances their finish times
    bool deferred = false;        // stage 2 may be left to the NEXT streamed launch of the handle (its merge team) or to
                                  // flush_pending(): the caller does not read the result before one of the two
    uint64_t ticket = 0;          // deferred sequence searches: the ticket that waits for this query
    bool require_stream = false;  // fail before launching unless the plan is the streamed single-query one
    bool inputs_ready = false;    // the k-mers are not produced by the preceding kernel of the stream
    uint32_t n_push = 0;
    uint32_t n_gather = 0;
    const unsigned long long *gather_blocks[kMaxSinks] = {};
    unsigned long long gather_seq = 0;
    unsigned long long *host_gather = nullptr;  // device address of the mapped host block the gathered hits are copied into
    uint32_t host_block_words = 0;
    // the number of k-mers comes from a preceding kernel (query front-end): only the streamed path can follow it;
    // run_query returns 1 without launching anything when the plan is a different one
    const unsigned long long *total_dev = nullptr;
    unsigned long long *scrub = nullptr;  // words the reduce kernel clears (front-end table)
    uint64_t scrub_words = 0;
    LlRoute ll = {};  // the k-mer bytes travel through the shards' low-latency inboxes
    // sequence front-end inside the gather kernel: the "k-mers" are a SEQUENCE of total_kmers windows; unique windows
    // and min_kmers = ceil(U * threshold) are determined on the device (needs the streamed plan, else run_query returns 1)
    bool seq_mode = false;
    double seq_threshold = 0;
};

// the sticky abort word of the handle (ptx.cuh:raise_abort), as the host sees it
unsigned long long abort_state(const bigsi_b200_index *ix)
{
    return ix->h_status.p ? *static_cast<volatile unsigned long long *>(ix->h_status.p) : 0ull;
}
int fail_aborted(unsigned long long v)
{
    static const char *what[] = {"?", "a query waited for stage 2 of an earlier query (entry gate)",
                                 "a CTA waited for another CTA's pooled row ids",
                                 "a shard waited for the query bytes of rank 0 (was rank 0's search launched?)",
                                 "a shard waited for another shard's hit list (was the search launched on every rank?)",
                                 "stage 2 of a query waited for its predecessor's (completion chain)",
                                 "a merge team waited for the gather CTAs of its own query (self-merging launch)"};
    const unsigned code = (unsigned)(v & 0xff);
    return fail(BIGSI_B200_ERR_TIMEOUT, "device-side wait timed out in query %llu: %s; the handle is unusable (destroy it)",
                (unsigned long long)(v >> 8), what[code < 7 ? code : 0]);
}

// Ring-buffered scratch of a streamed launch planned as `p` (partial planes, pool slots).  Growing a buffer
// synchronises `stream` first (earlier queries may still use the old one) and frees device memory, which waits for
// the WHOLE device -- callers that must not block there (column shards of one process sharing a GPU) reserve for
// their largest query up front (bigsi_b200_exchange_reserve).
int flush_pending(bigsi_b200_index *ix);
int reserve_stream_scratch(bigsi_b200_index *ix, const QueryParams &p, int grid, cudaStream_t stream)
{
    const uint64_t need = round_up(query_partial_bytes(p), 256);
    if (need > ix->stream_partial_slot) {
        if (int rc = flush_pending(ix)) return rc;  // its planes live in the buffer that is about to go
        CK(cudaStreamSynchronize(stream));
        cudaError_t e = ix->stream_partial.reserve(kStreamRing * (need + need / 4));
        if (e != cudaSuccess) return fail_cuda(e, "partial-plane workspace");
        ix->stream_partial_slot = ix->stream_partial.cap / kStreamRing / 256 * 256;
    }
    // pool slot: [ready flags: kPoolFlagEntries x u64][ids: grid x pool_share x h x i32].  The flag region has a FIXED
    // size: a ready word is compared with the launch epoch, so it must never alias bytes that an earlier query
    // with another grid used for row ids
    const uint64_t need_pool = round_up(kPoolFlagBytes + (uint64_t)grid * p.pool_share * p.h * 4 + 16, 256);
    if (need_pool > ix->pool_slot_bytes) {
        if (int rc = flush_pending(ix)) return rc;
        CK(cudaStreamSynchronize(stream));
        cudaError_t e = ix->d_pool.reserve(kStreamRing * (need_pool + need_pool / 4));
        if (e != cudaSuccess) return fail_cuda(e, "pool workspace");
        CK(cudaMemsetAsync(ix->d_pool.p, 0, ix->d_pool.cap, stream));
        ix->pool_slot_bytes = ix->d_pool.cap / kStreamRing / 256 * 256;
    }
    return 0;
}

// Stage 2 of the pending streamed query (if any) as a kernel of its own, on the stream it was launched on.
int flush_pending(bigsi_b200_index *ix)
{
    if (!ix->pending.have) return 0;
    bigsi_b200_index::PendingQuery &pq = ix->pending;
    pq.have = false;
    const cudaError_t e = launch_reduce(pq.p, pq.mode, pq.reduce_grid, pq.stream);
    if (e != cudaSuccess) {
        // the gather kernel ran without its stage 2: the completion chain is broken for good
        *static_cast<volatile unsigned long long *>(ix->h_status.p) = ((unsigned long long)pq.p.stream_seq << 8) | kAbortChain;
        return fail_cuda(e, "reduce_kernel launch");
    }
    ix->kernel_launches++;
    return 0;
}

// One query batch on `stream`.  Exactly one of d_rows / d_kmers is given; with k-mers the kernel
// hashes them itself when the plan allows, otherwise the hash kernel runs first into scratch rows.
int run_query(bigsi_b200_index *ix, int mode, const int32_t *d_rows, const char *d_kmers, int k, const int64_t *d_qoff,
              uint64_t n_queries, uint64_t total_kmers, uint64_t max_query_kmers, int h, void *d_out,
              uint64_t out_stride, cudaStream_t stream, const HitsOut *hits = nullptr)
{
    if (mode != BIGSI_B200_MODE_COUNTS && mode != BIGSI_B200_MODE_AND)
        return fail(BIGSI_B200_ERR_INVALID, "unknown query mode %d", mode);
    if (n_queries == 0) return 0;
    if (const unsigned long long av = abort_state(ix)) return fail_aborted(av);
    const uint64_t row_bytes = (ix->num_cols + 7) / 8;
    if (d_out && (mode == BIGSI_B200_MODE_COUNTS ? out_stride < ix->num_cols : out_stride < row_bytes))
        return fail(BIGSI_B200_ERR_INVALID, "out_stride %llu too small", (unsigned long long)out_stride);
    if (!d_out && !(hits && mode == BIGSI_B200_MODE_COUNTS)) return fail(BIGSI_B200_ERR_INVALID, "null output");
    if (d_kmers && k < 1) return fail(BIGSI_B200_ERR_INVALID, "k must be >= 1");
    if (ix->num_cols == 0) {
        if (hits) CK(cudaMemsetAsync(hits->n, 0, n_queries * sizeof(unsigned long long), stream));
        return 0;
    }
    if (int rc = ensure_kernels()) return rc;
    QueryParams p;
    int grid = 0;
    if (int rc = plan_query(ix, mode, n_queries, total_kmers, max_query_kmers, h, d_kmers != nullptr, k, p, grid,
                            hits != nullptr && hits->isolated))
        return rc;
    // a sequence (front-end inside the gather kernel) or a device-side k-mer count can only be followed by the streamed
    // plan: say so BEFORE anything is launched (the "k-mers" of a sequence search are not n x k bytes)
    if (hits && hits->seq_mode && !(p.stream && hits->n_sinks && k <= 32)) return 1;
    if (hits && hits->total_dev && !(p.stream && hits->n_sinks)) return 1;
    // ---- shared row-gather reuse (batches): de-duplicate the batch's k-mers by row-id tuple; when at most half of them
    // are distinct, gather the distinct tuples' AND vectors ONCE (lookup kernel -> scratch matrix A, one row per class)
    // and let the queries count over A with h = 1.  Bytes moved: (h + 1) * U' + T rows instead of h * T.
    uint64_t reuse_unique = 0;
    const bool reuse_candidate = n_queries > 1 && total_kmers >= kReuseMinKmers && total_kmers < 0xfffffff0ull && grid > 0 &&
                                 ix->opt_batch_reuse != 0 && !(hits && (hits->seq_mode || hits->total_dev));
    if (reuse_candidate) {
        if (int rc = flush_pending(ix)) return rc;
        const uint64_t T = total_kmers;
        cudaError_t e;
        if (d_kmers) {  // the row ids are needed up front: hash kernel (the plan below never hashes in the kernel)
            if (T * (uint64_t)h * 4 > ix->d_rows.cap) {
                CK(cudaStreamSynchronize(stream));
                if ((e = ix->d_rows.reserve(T * (uint64_t)h * 4)) != cudaSuccess) return fail_cuda(e, "row-id workspace");
            }
            CK(launch_hash_kmers(d_kmers, T, k, h, ix->num_rows, 1, static_cast<int32_t *>(ix->d_rows.p), stream));
            ix->kernel_launches++;
            d_rows = static_cast<const int32_t *>(ix->d_rows.p);
            d_kmers = nullptr;
        }
        uint64_t entries = 1024;
        while (entries < 2 * T) entries <<= 1;
        // scratch: [table: entries x u64][counter: 256 B][rep: T x u32][uid_of: T x u32][ids: T x i32][unique rows: T x h x i32]
        const uint64_t off_counter = entries * 8, off_rep = off_counter + 256, off_uid = off_rep + round_up(T * 4, 256),
                       off_ids = off_uid + round_up(T * 4, 256), off_urows = off_ids + round_up(T * 4, 256),
                       need = off_urows + round_up(T * (uint64_t)h * 4, 256);
        if (need > ix->d_reuse.cap) {
            CK(cudaStreamSynchronize(stream));
            if ((e = ix->d_reuse.reserve(need)) != cudaSuccess) return fail_cuda(e, "batch de-duplication workspace");
        }
        uint8_t *rb = static_cast<uint8_t *>(ix->d_reuse.p);
        CK(cudaMemsetAsync(rb, 0, off_counter + 256, stream));
        CK(launch_dedup_rows(d_rows, T, h, reinterpret_cast<unsigned long long *>(rb), entries, reinterpret_cast<uint32_t *>(rb + off_rep),
                             reinterpret_cast<uint32_t *>(rb + off_uid), reinterpret_cast<unsigned int *>(rb + off_counter),
                             reinterpret_cast<int32_t *>(rb + off_ids), reinterpret_cast<int32_t *>(rb + off_urows), stream));
        ix->kernel_launches += 2;
        if ((e = ix->h_small.reserve(64)) != cudaSuccess) return fail_cuda(e, "pinned staging");
        CK(cudaMemcpyAsync(ix->h_small.p, rb + off_counter, 4, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        const uint64_t U = *static_cast<const unsigned int *>(ix->h_small.p);
        const uint64_t pitch_a = round_up(row_bytes, 128);
        bool use = U > 0 && 2 * U <= T;
        if (use && U * pitch_a > ix->d_reuse_rows.cap) {
            size_t free_b = 0, total_b = 0;
            CK(cudaMemGetInfo(&free_b, &total_b));
            if (U * pitch_a + (1ull << 30) > (uint64_t)free_b + ix->d_reuse_rows.cap) use = false;  // not worth evicting anything
            else if ((e = ix->d_reuse_rows.reserve(U * pitch_a)) != cudaSuccess) {
                (void)cudaGetLastError();
                use = false;
            }
        }
        int h_run = h;
        if (use) {
            CK(launch_lookup(ix->matrix, ix->pitch, (uint32_t)row_bytes, reinterpret_cast<const int32_t *>(rb + off_urows), U, h,
                             static_cast<uint8_t *>(ix->d_reuse_rows.p), pitch_a, stream));
            ix->kernel_launches++;
            d_rows = reinterpret_cast<const int32_t *>(rb + off_ids);
            h_run = 1;
            reuse_unique = U;
        }
        // re-plan for row ids (no in-kernel hashing) and, with reuse, one "row" per k-mer
        if (int rc = plan_query(ix, mode, n_queries, total_kmers, max_query_kmers, h_run, false, k, p, grid, false)) return rc;
        if (use) {
            p.matrix = static_cast<const uint8_t *>(ix->d_reuse_rows.p);
            p.pitch = pitch_a;
        }
    }
    if (d_kmers && p.prehash) {
        p.kmers = reinterpret_cast<const uint8_t *>(d_kmers);
    } else if (d_kmers && total_kmers) {
        if (total_kmers * (uint64_t)h * 4 > ix->d_rows.cap) {
            CK(cudaStreamSynchronize(stream));
            cudaError_t e = ix->d_rows.reserve(total_kmers * (uint64_t)h * 4);
            if (e != cudaSuccess) return fail_cuda(e, "row-id workspace");
        }
        CK(launch_hash_kmers(d_kmers, total_kmers, k, h, ix->num_rows, 1, static_cast<int32_t *>(ix->d_rows.p), stream));
        ix->kernel_launches++;
        d_rows = static_cast<const int32_t *>(ix->d_rows.p);
    }
    p.rows = d_rows;
    p.qoff = d_qoff;
    p.out = d_out;
    p.out_stride = out_stride;
    p.debug_flags = (uint32_t)ix->opt_debug_flags;
    p.plain_launch = ix->opt_cooperative ? 0u : 1u;
    int reduce_grid = 0;
    if (p.stream) {
        reduce_grid = (int)std::min<uint64_t>(p.merge_items, (uint64_t)ix->sm_count);
        if (reduce_grid < 1) reduce_grid = 1;
    }
    if (p.debug_flags & 2u) {
        // timeline stamps, fetched with bigsi_b200_index_debug_read: [gather grid + reduce grid][kDebugStamps] words per
        // launch; streamed launches rotate over kStreamStates such regions (query seq uses region seq % kStreamStates),
        // so the schedule of several consecutive queries can be read back
        // (stage 2's stamps are written by whoever executes it: at most sm_count CTAs)
        const uint64_t region = (uint64_t)((grid > 0 ? grid : 1) + std::max(reduce_grid, ix->sm_count)) * kDebugStamps;
        cudaError_t de = ix->debug_ts.reserve(region * 8 * kStreamStates);
        if (de != cudaSuccess) return fail_cuda(de, "debug buffer");
        p.debug_ts = static_cast<unsigned long long *>(ix->debug_ts.p) + (p.stream ? ((ix->stream_seq + 1) % kStreamStates) * region : 0);
    }
    bool will_publish = false;
    if (hits) {
        p.min_kmers = hits->min_kmers;
        p.hit_cols = hits->cols;
        p.hit_counts = hits->counts;
        p.n_hits = hits->n;
        p.hit_cap = hits->cap;
        if (hits->by_value) {
            p.min_by_value = 1;
            p.min_kmers_value = hits->min_value;
        }
        if (hits->total_dev) {
            if (!(p.stream && hits->n_sinks)) return 1;
            p.total_dev = hits->total_dev;
            p.scrub = hits->scrub;
            p.scrub_words = hits->scrub_words;
        }
        if (hits->seq_mode && !(p.stream && hits->n_sinks && k <= 32)) return 1;
        if (hits->require_stream && !p.stream)
            return fail(BIGSI_B200_ERR_INVALID, "this query cannot run as a streamed single-query launch (prehash=%u grid=%d)",
                        p.prehash, grid);
        p.n_push = hits->n_push;
        p.push_repeat = (uint32_t)ix->opt_push_repeat;
        p.push_all_warps = (uint32_t)ix->opt_push_all_warps;
        p.n_gather = hits->n_gather;
        for (uint32_t i = 0; i < hits->n_gather; ++i) p.gather_blocks[i] = hits->gather_blocks[i];
        p.gather_seq = hits->gather_seq;
        p.host_gather = hits->host_gather;
        p.host_block_words = hits->host_block_words;
        p.ll = hits->ll;
        p.ll.kmers_base = reinterpret_cast<const uint8_t *>(d_kmers);
        p.sink_spec = hits->sink_spec;
        if (hits->published) *hits->published = false;
        if (hits->n_sinks && (p.stream || p.fuse_merge) && grid > 0 && n_queries == 1) {
            p.n_sinks = hits->n_sinks;
            p.sink_spec = hits->sink_spec;
            p.sink_seq = hits->sink_seq;
            for (uint32_t i = 0; i < hits->n_sinks; ++i) p.sinks[i] = hits->sinks[i];
            will_publish = true;
        }
        // stage 1 zeroes the hit counters; without a stage-1 launch do it here
        if (grid == 0) CK(cudaMemsetAsync(hits->n, 0, n_queries * sizeof(unsigned long long), stream));
    }
    if (p.stream) {
        // ---- streamed launch: gather kernel (+ merge team / flush kernel), ring-buffered scratch ------------------
        const uint64_t seq = ix->stream_seq + 1;
        const uint64_t slot = seq % kStreamRing;
        if (int rc = reserve_stream_scratch(ix, p, grid, stream)) return rc;
        p.partial = static_cast<uint8_t *>(ix->stream_partial.p) + slot * ix->stream_partial_slot;
        uint8_t *pb = static_cast<uint8_t *>(ix->d_pool.p) + slot * ix->pool_slot_bytes;
        p.pool_ready = reinterpret_cast<unsigned long long *>(pb);
        p.pool_ids = reinterpret_cast<int32_t *>(pb + kPoolFlagBytes);
        p.pool_epoch = ix->pool_epoch + 1;
        uint8_t *sb = static_cast<uint8_t *>(ix->d_stream.p);
        QState *states = reinterpret_cast<QState *>(sb + 3 * 64);
        p.stream_done = reinterpret_cast<unsigned long long *>(sb);
        p.abort_word = reinterpret_cast<unsigned long long *>(sb + 64);
        p.wait_ns_out = reinterpret_cast<unsigned long long *>(sb + 128);
        void *hs = nullptr;
        CK(cudaHostGetDevicePointer(&hs, ix->h_status.p, 0));
        p.host_abort = static_cast<unsigned long long *>(hs);
        p.spin_timeout_ns = (unsigned long long)ix->opt_spin_timeout_ms * 1000000ull;
        p.stream_seq = seq;
        p.qstate = states + seq % kStreamStates;
        p.qstate_next = states + (seq + kStreamRing) % kStreamStates;
        p.pool_counter = &p.qstate->pool_claims;
        if (hits && hits->seq_mode) {
            // table `slot` of the ring; live entries carry this use's 16-bit epoch, so the table is only cleared when the
            // epoch wraps (every 65 535 uses) -- stream-ordered, which serialises one query in a quarter of a million
            uint64_t T = 1024;
            while (T < 2 * total_kmers) T <<= 1;
            if (T > ix->seq_table_entries) return fail(BIGSI_B200_ERR_INVALID, "internal: sequence table not reserved");
            uint64_t &uses = ix->seq_table_uses[slot];
            unsigned long long *table = static_cast<unsigned long long *>(ix->d_seq_tables.p) + slot * ix->seq_table_entries;
            if (uses && uses % 65535 == 0) CK(cudaMemsetAsync(table, 0, ix->seq_table_entries * 8, stream));
            p.seq_mode = 1;
            p.seq_threshold = hits->seq_threshold;
            p.seq_table = table;
            p.seq_table_entries = T;
            p.seq_epoch = (uint32_t)(uses % 65535) + 1;
            ++uses;
            p.total_dev = &p.qstate->n_unique;  // reported in the sink block by the reduce kernel
        }
        p.ll.abort_word = p.abort_word;
        p.ll.host_abort = p.host_abort;
        p.ll.timeout_ns = p.spin_timeout_ns;
        p.ll.seq = seq;
        // the kernel must wait for its predecessor in the stream when that may produce its input: always for a
        // device-side k-mer count, and for caller-provided device k-mers unless the caller says otherwise
        const bool ready = (hits && hits->inputs_ready) || ix->opt_inputs_ready != 0 || p.ll.in != nullptr;  // (a peer shard reads its inbox)
        p.stream_wait_inputs = (ready && !(hits && hits->total_dev)) ? 0u : 1u;

        // the previous streamed query, if its stage 2 is still pending: this launch's merge team takes it when it can
        // (same stream, COUNTS, scratch within the team's), otherwise it is flushed first
        bigsi_b200_index::PendingQuery &pq = ix->pending;
        const bool take_prev = pq.have && p.merge_team != 0 && pq.stream == stream && pq.mode == BIGSI_B200_MODE_COUNTS &&
                               pq.p.merge_smem <= (uint32_t)kReduceSmemBytes && !ix->timing;
        if (pq.have && !take_prev)
            if (int rc = flush_pending(ix)) return rc;
        p.merge_prev = take_prev ? 1u : 0u;
        const bool defer = hits && hits->deferred && mode == BIGSI_B200_MODE_COUNTS && !ix->timing && ix->opt_defer != 0 &&
                           p.merge_smem <= (uint32_t)kReduceSmemBytes;
        // option "self_merge": a query nobody follows (synchronous calls) is merged by its OWN kernel's team once every
        // gather CTA has flushed -- needs all CTAs co-resident, so the launch is cooperative (the driver checks).  Off by
        // default: measured no faster than the flush kernel, which is already queued behind the gather kernel (PDL)
        bool self_merge = !defer && p.merge_team != 0 && !ix->timing && ix->opt_self_merge != 0 && grid <= ix->sm_count &&
                          p.merge_smem <= (uint32_t)kReduceSmemBytes;
        p.self_merge = self_merge ? 1u : 0u;

        TimedLaunch tl{};
        if (ix->timing) {
            if (ix->timed_free.empty()) {
                CK(cudaEventCreate(&tl.e0));
                CK(cudaEventCreate(&tl.e1));
                CK(cudaEventCreate(&tl.e2));
            } else {
                tl = ix->timed_free.back();
                ix->timed_free.pop_back();
            }
            CK(cudaEventRecord(tl.e0, stream));
        }
        cudaError_t e = launch_query(p, mode, grid, stream, take_prev ? &pq.p : nullptr);
        if (e != cudaSuccess && self_merge) {  // the cooperative launch was refused (SMs held by something else): two kernels
            (void)cudaGetLastError();
            self_merge = false;
            p.self_merge = 0;
            e = launch_query(p, mode, grid, stream, take_prev ? &pq.p : nullptr);
        }
        if (e != cudaSuccess) return fail_cuda(e, "gather_solo launch");  // (a pending query stays pending)
        ix->kernel_launches++;
        if (ix->timing) CK(cudaEventRecord(tl.e1, stream));
        // from here on the query counts as launched: the completion chain expects its stage 2
        ix->stream_seq = seq;
        ix->pool_epoch = p.pool_epoch;
        if (take_prev) pq.have = false;  // (merged by this launch's team)
        if (!self_merge) {
            pq.have = true;
            pq.p = p;
            if (pq.p.debug_ts) pq.p.debug_ts += (uint64_t)grid * kDebugStamps;
            pq.mode = mode;
            pq.reduce_grid = reduce_grid;
            pq.stream = stream;
            pq.ticket = hits ? hits->ticket : 0;
            if (!defer)
                if (int rc = flush_pending(ix)) return rc;
        }
        if (ix->timing) {
            CK(cudaEventRecord(tl.e2, stream));
            ix->timed_used.push_back(tl);
        }
        if (hits && hits->published) *hits->published = will_publish;
    } else {
        if (int rc = flush_pending(ix)) return rc;
        // batches in COUNTS mode: queries that lie inside one slice are finished by the CTA that counts them; the hit
        // counters are zeroed here (the kernel's CTAs start at different times, so it cannot do that itself)
        if (mode == BIGSI_B200_MODE_COUNTS && n_queries > 1 && grid > 0 && ix->opt_direct != 0) {
            p.direct_complete = 1;
            if (hits) CK(cudaMemsetAsync(hits->n, 0, n_queries * sizeof(unsigned long long), stream));
        }
        const uint64_t need = query_partial_bytes(p);
        if (need > ix->partial.cap) {
            CK(cudaStreamSynchronize(stream));
            cudaError_t e = ix->partial.reserve(need);
            if (e != cudaSuccess) return fail_cuda(e, "partial-plane workspace");
        }
        p.partial = static_cast<uint8_t *>(ix->partial.p);
        TimedLaunch tl{};
        if (ix->timing) {
            if (ix->timed_free.empty()) {
                CK(cudaEventCreate(&tl.e0));
                CK(cudaEventCreate(&tl.e1));
                CK(cudaEventCreate(&tl.e2));
            } else {
                tl = ix->timed_free.back();
                ix->timed_free.pop_back();
            }
            CK(cudaEventRecord(tl.e0, stream));
        }
        if (grid > 0) {
            // the host-side targets of the monotonic device counters advance only with a successful launch
            if (p.fuse_merge) {
                p.barrier = static_cast<unsigned long long *>(ix->d_barrier.p);
                p.barrier_target = ix->barrier_target + (uint64_t)grid;
            }
            if (will_publish) {
                p.done_counter = static_cast<unsigned long long *>(ix->d_barrier.p) + 16;
                p.done_target = ix->done_target + (uint64_t)grid;
            }
            cudaError_t e = launch_query(p, mode, grid, stream);
            if (e != cudaSuccess) return fail_cuda(e, "fused_query launch");
            if (p.fuse_merge) ix->barrier_target = p.barrier_target;
            if (will_publish) ix->done_target = p.done_target;
            ix->kernel_launches++;
        } else {
            will_publish = false;
        }
        if (hits && hits->published) *hits->published = will_publish;
        if (ix->timing) CK(cudaEventRecord(tl.e1, stream));
        if (!p.fuse_merge) {
            CK(launch_merge(p, mode, stream));
            ix->kernel_launches++;
        }
        if (ix->timing) {
            CK(cudaEventRecord(tl.e2, stream));
            ix->timed_used.push_back(tl);
        }
    }

    bigsi_b200_info &s = ix->stats;
    s.last_kmers = total_kmers;
    s.last_algorithmic_bytes = total_kmers * (uint64_t)h * row_bytes;
    s.last_grid = (uint32_t)grid;
    s.last_block = query_block_threads(p);
    s.last_smem_bytes = query_smem_bytes(p);
    s.last_tile_bytes = p.tile_bytes;
    s.last_n_tiles = p.n_tiles;
    s.last_kmers_per_stage = p.kmers_per_stage;
    s.last_n_stages = p.n_stages;
    s.last_n_slices = p.n_slices;
    s.last_fused = (p.fuse_merge ? 1u : 0u) | (p.prehash ? 2u : 0u) | (p.solo ? 4u : 0u) | (p.stream ? 8u : 0u);
    s.last_reduce_grid = (uint32_t)reduce_grid;
    s.last_unique_kmers = reuse_unique;
    return 0;
}

int check_index(const bigsi_b200_index *ix)
{
    if (!ix) return fail(BIGSI_B200_ERR_INVALID, "null index handle");
    return 0;
}

// q_offsets sanity on the host (host-buffer entry points only)
int check_offsets(const int64_t *qoff, uint64_t n_queries, uint64_t *total_out, uint64_t *longest_out)
{
    if (n_queries == 0) {
        *total_out = 0;
        *longest_out = 0;
        return 0;
    }
    if (!qoff) return fail(BIGSI_B200_ERR_INVALID, "null q_offsets");
    if (qoff[0] != 0) return fail(BIGSI_B200_ERR_INVALID, "q_offsets[0] must be 0");
    uint64_t longest = 0;
    for (uint64_t q = 0; q < n_queries; ++q) {
        if (qoff[q + 1] < qoff[q]) return fail(BIGSI_B200_ERR_INVALID, "q_offsets must be non-decreasing");
        const uint64_t len = (uint64_t)(qoff[q + 1] - qoff[q]);
        if (len > longest) longest = len;
    }
    *total_out = (uint64_t)qoff[n_queries];
    *longest_out = longest;
    return 0;
}

}  // namespace

// ============================================================================================
// library
// ============================================================================================
extern "C" {

int bigsi_b200_abi_version(void) { return BIGSI_B200_ABI_VERSION; }
const char *bigsi_b200_last_error(void) { return g_err.c_str(); }

int bigsi_b200_device_count(int *count_out)
{
    if (!count_out) return fail(BIGSI_B200_ERR_INVALID, "null count_out");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count_out = 0;
        return fail_cuda(e, "cudaGetDeviceCount");
    }
    *count_out = n;
    return 0;
}

int bigsi_b200_host_alloc(uint64_t bytes, void **ptr_out)
{
    if (!ptr_out) return fail(BIGSI_B200_ERR_INVALID, "null ptr_out");
    *ptr_out = nullptr;
    // mapped + portable: the query kernel can read k-mers straight out of such a buffer (zero copy)
    cudaError_t e = cudaHostAlloc(ptr_out, bytes ? bytes : 1, cudaHostAllocMapped | cudaHostAllocPortable);
    if (e != cudaSuccess) return fail_cuda(e, "cudaHostAlloc");
    return 0;
}
int bigsi_b200_host_free(void *ptr)
{
    if (!ptr) return 0;
    CK(cudaFreeHost(ptr));
    return 0;
}

// ============================================================================================
// index lifecycle
// ============================================================================================
int bigsi_b200_index_create(int device, uint64_t num_rows, uint64_t num_cols, uint64_t col_capacity,
                            uint64_t col_offset, bigsi_b200_index **index_out)
{
    if (!index_out) return fail(BIGSI_B200_ERR_INVALID, "null index_out");
    *index_out = nullptr;
    if (num_rows == 0 || num_rows > 0x7fffffffull)
        return fail(BIGSI_B200_ERR_INVALID, "num_rows must be in [1, 2^31-1] (row ids are int32)");
    if (col_capacity < num_cols) col_capacity = num_cols;
    if (col_capacity == 0) col_capacity = 1;
    if (col_capacity > 0xffffffffull - 1024) return fail(BIGSI_B200_ERR_INVALID, "too many columns for one shard");
    if (col_offset % 8) return fail(BIGSI_B200_ERR_INVALID, "col_offset must be a multiple of 8");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail_cuda(e, "cudaGetDeviceCount");
    if (n == 0) return fail(BIGSI_B200_ERR_NO_DEVICE, "no CUDA device");
    if (device < 0 || device >= n) return fail(BIGSI_B200_ERR_INVALID, "device %d out of range (have %d)", device, n);
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail_cuda(guard.err, "cudaSetDevice");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(BIGSI_B200_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    bigsi_b200_index *ix = new bigsi_b200_index();
    ix->device = device;
    ix->sm_count = prop.multiProcessorCount;
    ix->num_rows = num_rows;
    ix->num_cols = num_cols;
    ix->pitch = round_up((col_capacity + 7) / 8, 128);
    ix->col_capacity = ix->pitch * 8;
    ix->col_offset = col_offset;
    const uint64_t bytes = ix->num_rows * ix->pitch;
    e = cudaMalloc(reinterpret_cast<void **>(&ix->matrix), bytes);
    if (e != cudaSuccess) {
        delete ix;
        return fail_cuda(e, "cudaMalloc(matrix)");
    }
    e = cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = ix->d_barrier.reserve(256);
    if (e == cudaSuccess) e = cudaMemsetAsync(ix->d_barrier.p, 0, 256, ix->stream);
    if (e == cudaSuccess) e = ix->d_stream.reserve(kStreamStateBytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(ix->d_stream.p, 0, kStreamStateBytes, ix->stream);
    if (e == cudaSuccess) e = ix->h_status.reserve(64);
    if (e == cudaSuccess) memset(ix->h_status.p, 0, 64);
    if (e == cudaSuccess) e = cudaMemsetAsync(ix->matrix, 0, bytes, ix->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
    if (e != cudaSuccess) {
        cudaFree(ix->matrix);
        ix->d_barrier.release();
        ix->d_stream.release();
        ix->h_status.release();
        if (ix->stream) cudaStreamDestroy(ix->stream);
        delete ix;
        return fail_cuda(e, "index initialisation");
    }
    *index_out = ix;
    return 0;
}

int bigsi_b200_index_destroy(bigsi_b200_index *ix)
{
    if (!ix) return 0;
    bigsi_b200_exchange_destroy(ix);
    DeviceGuard guard(ix->device);
    cudaDeviceSynchronize();
    for (auto &t : ix->timed_free) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); cudaEventDestroy(t.e2); }
    for (auto &t : ix->timed_used) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); cudaEventDestroy(t.e2); }
    for (auto &b : ix->h_seq) b.release();
    ix->h_tsink.release();
    DevBuf *bufs[] = {&ix->d_seq_tables, &ix->stream_partial, &ix->d_stream, &ix->d_hits_ring, &ix->d_seq, &ix->d_table, &ix->d_pool, &ix->d_barrier, &ix->debug_ts, &ix->partial, &ix->d_kmers, &ix->d_rows, &ix->d_qoff, &ix->d_out, &ix->d_min,
                      &ix->d_nhits, &ix->d_bloom, &ix->d_planted, &ix->d_reuse, &ix->d_reuse_rows};
    for (DevBuf *b : bufs) b->release();
    ix->h_small.release();
    ix->h_status.release();
    ix->h_sink.release();
    ix->h_kmers.release();
    if (ix->matrix) cudaFree(ix->matrix);
    if (ix->stream) cudaStreamDestroy(ix->stream);
    delete ix;
    return 0;
}

int bigsi_b200_index_get_info(const bigsi_b200_index *ix, bigsi_b200_info *out)
{
    if (int rc = check_index(ix)) return rc;
    if (!out) return fail(BIGSI_B200_ERR_INVALID, "null info_out");
    *out = ix->stats;
    out->num_rows = ix->num_rows;
    out->num_cols = ix->num_cols;
    out->col_capacity = ix->col_capacity;
    out->col_offset = ix->col_offset;
    out->row_bytes = (ix->num_cols + 7) / 8;
    out->row_pitch_bytes = ix->pitch;
    out->matrix_bytes = ix->num_rows * ix->pitch;
    out->device = ix->device;
    out->sm_count = ix->sm_count;
    out->kernel_launches = ix->kernel_launches;
    out->scratch_bytes = ix->partial.cap;
    return 0;
}

int bigsi_b200_index_set_option(bigsi_b200_index *ix, const char *key, int64_t value)
{
    if (int rc = check_index(ix)) return rc;
    if (!key) return fail(BIGSI_B200_ERR_INVALID, "null option key");
    if (value < 0) return fail(BIGSI_B200_ERR_INVALID, "option %s: negative value", key);
    if (!strcmp(key, "tile_bytes")) ix->opt_tile_bytes = value;
    else if (!strcmp(key, "grid")) ix->opt_grid = value;
    else if (!strcmp(key, "kmers_per_stage")) ix->opt_kmers_per_stage = value;
    else if (!strcmp(key, "n_stages")) ix->opt_n_stages = value;
    else if (!strcmp(key, "ctas_per_sm")) ix->opt_ctas_per_sm = value;
    else if (!strcmp(key, "timing")) ix->timing = value != 0;
    else if (!strcmp(key, "debug_flags")) ix->opt_debug_flags = value;
    else if (!strcmp(key, "prehash")) ix->opt_prehash = value;
    else if (!strcmp(key, "fuse_merge")) ix->opt_fuse_merge = value;
    else if (!strcmp(key, "merge_chunk_bytes")) ix->opt_merge_chunk_bytes = value;
    else if (!strcmp(key, "solo")) ix->opt_solo = value;
    else if (!strcmp(key, "zero_copy")) ix->opt_zero_copy = value;
    else if (!strcmp(key, "cooperative")) ix->opt_cooperative = value;
    else if (!strcmp(key, "inputs_ready")) ix->opt_inputs_ready = value;
    else if (!strcmp(key, "self_merge")) ix->opt_self_merge = value;  // 1: a synchronous single query is merged by its own
                                                                      // kernel's team (cooperative launch) instead of the flush kernel
    else if (!strcmp(key, "push_repeat")) ix->opt_push_repeat = value > 16 ? 16 : value;  // diagnostics
    else if (!strcmp(key, "push_all_warps")) ix->opt_push_all_warps = value;             // diagnostics
    else if (!strcmp(key, "batch_reuse")) ix->opt_batch_reuse = value;
    else if (!strcmp(key, "direct")) ix->opt_direct = value;  // 0: batches merge every query (no direct finish)
    else if (!strcmp(key, "defer")) ix->opt_defer = value;  // 0: every streamed query is flushed at once (diagnostics)
    else if (!strcmp(key, "spin_timeout_ms")) ix->opt_spin_timeout_ms = value < 1 ? 1 : value;
    else if (!strcmp(key, "pool_pct")) ix->opt_pool_pct = value > 100 ? -1 : value;  // > 100 = automatic
    else return fail(BIGSI_B200_ERR_INVALID, "unknown option '%s'", key);
    return 0;
}

int bigsi_b200_index_debug_read(bigsi_b200_index *ix, uint64_t *out, uint64_t n_words)
{
    if (int rc = check_index(ix)) return rc;
    if (!out) return fail(BIGSI_B200_ERR_INVALID, "null out");
    if (n_words * 8 > ix->debug_ts.cap) return fail(BIGSI_B200_ERR_RANGE, "debug buffer holds %llu bytes",
                                                     (unsigned long long)ix->debug_ts.cap);
    DeviceGuard guard(ix->device);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, ix->debug_ts.p, n_words * 8, cudaMemcpyDeviceToHost));
    return 0;
}

int bigsi_b200_index_timing_collect(bigsi_b200_index *ix, double *fused_ms_out, double *merge_ms_out, uint64_t *n_out)
{
    if (int rc = check_index(ix)) return rc;
    DeviceGuard guard(ix->device);
    double fused = 0, merge = 0;
    uint64_t n = 0;
    for (auto &t : ix->timed_used) {
        CK(cudaEventSynchronize(t.e2));
        float a = 0, b = 0;
        CK(cudaEventElapsedTime(&a, t.e0, t.e1));
        CK(cudaEventElapsedTime(&b, t.e1, t.e2));
        fused += a;
        merge += b;
        ++n;
        ix->timed_free.push_back(t);
    }
    ix->timed_used.clear();
    if (fused_ms_out) *fused_ms_out = fused;
    if (merge_ms_out) *merge_ms_out = merge;
    if (n_out) *n_out = n;
    return 0;
}

int bigsi_b200_index_upload_rows(bigsi_b200_index *ix, uint64_t row0, uint64_t n_rows, const uint8_t *rows,
                                 uint64_t src_stride, uint64_t src_byte_offset)
{
    if (int rc = check_index(ix)) return rc;
    if (n_rows == 0) return 0;
    if (!rows) return fail(BIGSI_B200_ERR_INVALID, "null rows");
    if (row0 + n_rows > ix->num_rows) return fail(BIGSI_B200_ERR_RANGE, "rows [%llu,%llu) exceed m=%llu",
                                                   (unsigned long long)row0, (unsigned long long)(row0 + n_rows),
                                                   (unsigned long long)ix->num_rows);
    const uint64_t row_bytes = (ix->num_cols + 7) / 8;
    if (row_bytes == 0) return 0;
    if (src_stride < row_bytes) return fail(BIGSI_B200_ERR_INVALID, "src_stride smaller than the row");
    DeviceGuard guard(ix->device);
    CK(cudaMemcpy2DAsync(ix->matrix + row0 * ix->pitch, ix->pitch, rows + src_byte_offset, src_stride, row_bytes, n_rows,
                         cudaMemcpyHostToDevice, ix->stream));
    CK(cudaStreamSynchronize(ix->stream));
    // a partial last byte may carry columns >= num_cols of a wider source row: keep padding zero
    if (ix->num_cols & 7) {
        // handled on the host side of the plugin (rows are uploaded at their own width); enforce here
        // by re-masking the last byte of every uploaded row.
        const uint8_t mask = (uint8_t)(0xff00u >> (ix->num_cols & 7));
        std::vector<uint8_t> last(n_rows);
        for (uint64_t i = 0; i < n_rows; ++i) last[i] = rows[src_byte_offset + i * src_stride + row_bytes - 1] & mask;
        CK(cudaMemcpy2DAsync(ix->matrix + row0 * ix->pitch + row_bytes - 1, ix->pitch, last.data(), 1, 1, n_rows,
                             cudaMemcpyHostToDevice, ix->stream));
        CK(cudaStreamSynchronize(ix->stream));
    }
    return 0;
}

int bigsi_b200_index_download_rows(const bigsi_b200_index *ix, uint64_t row0, uint64_t n_rows, uint8_t *out,
                                   uint64_t dst_stride)
{
    if (int rc = check_index(ix)) return rc;
    if (n_rows == 0) return 0;
    if (!out) return fail(BIGSI_B200_ERR_INVALID, "null out");
    if (row0 + n_rows > ix->num_rows) return fail(BIGSI_B200_ERR_RANGE, "rows out of range");
    const uint64_t row_bytes = (ix->num_cols + 7) / 8;
    if (row_bytes == 0) return 0;
    if (dst_stride < row_bytes) return fail(BIGSI_B200_ERR_INVALID, "dst_stride smaller than the row");
    DeviceGuard guard(ix->device);
    CK(cudaMemcpy2DAsync(out, dst_stride, ix->matrix + row0 * ix->pitch, ix->pitch, row_bytes, n_rows,
                         cudaMemcpyDeviceToHost, ix->stream));
    CK(cudaStreamSynchronize(ix->stream));
    return 0;
}

int bigsi_b200_index_set_column(bigsi_b200_index *ix, uint64_t col, const uint8_t *bloom, uint64_t n_bits)
{
    if (int rc = check_index(ix)) return rc;
    if (!bloom && n_bits) return fail(BIGSI_B200_ERR_INVALID, "null bloom filter");
    if (n_bits > ix->num_rows) return fail(BIGSI_B200_ERR_RANGE, "bloom filter longer than m");
    if (col > ix->num_cols) return fail(BIGSI_B200_ERR_RANGE, "column %llu beyond num_cols=%llu",
                                        (unsigned long long)col, (unsigned long long)ix->num_cols);
    if (col >= ix->col_capacity) return fail(BIGSI_B200_ERR_RANGE, "column capacity %llu exhausted",
                                             (unsigned long long)ix->col_capacity);
    DeviceGuard guard(ix->device);
    const uint64_t nbytes = (n_bits + 7) / 8;
    cudaError_t e = ix->d_bloom.reserve(nbytes ? nbytes : 1);
    if (e != cudaSuccess) return fail_cuda(e, "bloom staging");
    if (nbytes) CK(cudaMemcpyAsync(ix->d_bloom.p, bloom, nbytes, cudaMemcpyHostToDevice, ix->stream));
    CK(launch_set_column(ix->matrix, ix->pitch, ix->num_rows, col, static_cast<const uint8_t *>(ix->d_bloom.p), n_bits,
                         ix->stream));
    ix->kernel_launches++;
    CK(cudaStreamSynchronize(ix->stream));
    if (col == ix->num_cols) ix->num_cols++;
    return 0;
}

int bigsi_b200_index_fill_synthetic(bigsi_b200_index *ix, uint64_t seed, int and_draws, const uint64_t *planted_cols,
                                    const uint32_t *planted_thr, int n_planted)
{
    if (int rc = check_index(ix)) return rc;
    if (and_draws < 0 || and_draws > 8) return fail(BIGSI_B200_ERR_INVALID, "and_draws must be in [0,8]");
    if (n_planted < 0 || (n_planted && (!planted_cols || !planted_thr)))
        return fail(BIGSI_B200_ERR_INVALID, "bad planted-column list");
    DeviceGuard guard(ix->device);
    const uint64_t pc_bytes = round_up((uint64_t)n_planted * 8, 256);
    uint64_t *d_cols = nullptr;
    uint32_t *d_thr = nullptr;
    if (n_planted) {
        cudaError_t e = ix->d_planted.reserve(pc_bytes + (uint64_t)n_planted * 4);
        if (e != cudaSuccess) return fail_cuda(e, "planted staging");
        d_cols = static_cast<uint64_t *>(ix->d_planted.p);
        d_thr = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(ix->d_planted.p) + pc_bytes);
        CK(cudaMemcpyAsync(d_cols, planted_cols, (uint64_t)n_planted * 8, cudaMemcpyHostToDevice, ix->stream));
        CK(cudaMemcpyAsync(d_thr, planted_thr, (uint64_t)n_planted * 4, cudaMemcpyHostToDevice, ix->stream));
    }
    CK(launch_fill_synthetic(ix->matrix, ix->pitch, ix->num_rows, ix->num_cols, ix->col_offset, seed, and_draws, d_cols,
                             d_thr, n_planted, ix->stream));
    ix->kernel_launches += n_planted ? 2 : 1;
    CK(cudaStreamSynchronize(ix->stream));
    return 0;
}

// ============================================================================================
// hashing
// ============================================================================================
int bigsi_b200_hash_kmers_dev(const char *d_kmers, uint64_t n, int k, int h, uint64_t m, int canonical,
                              int32_t *d_rows_out, void *stream)
{
    if (k < 1) return fail(BIGSI_B200_ERR_INVALID, "k must be >= 1");
    if (h < 1) return fail(BIGSI_B200_ERR_INVALID, "h must be >= 1");
    if (m == 0 || m > 0x7fffffffull) return fail(BIGSI_B200_ERR_INVALID, "m must be in [1, 2^31-1]");
    if (n == 0) return 0;
    if (!d_kmers || !d_rows_out) return fail(BIGSI_B200_ERR_INVALID, "null device pointer");
    CK(launch_hash_kmers(d_kmers, n, k, h, m, canonical, d_rows_out, static_cast<cudaStream_t>(stream)));
    return 0;
}

int bigsi_b200_hash_kmers(int device, const char *kmers, uint64_t n, int k, int h, uint64_t m, int canonical,
                          int32_t *rows_out)
{
    if (k < 1 || h < 1) return fail(BIGSI_B200_ERR_INVALID, "k and h must be >= 1");
    if (n == 0) return 0;
    if (!kmers || !rows_out) return fail(BIGSI_B200_ERR_INVALID, "null pointer");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess) return fail_cuda(e, "cudaGetDeviceCount");
    if (ndev == 0) return fail(BIGSI_B200_ERR_NO_DEVICE, "no CUDA device");
    if (device < 0 || device >= ndev) return fail(BIGSI_B200_ERR_INVALID, "device out of range");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail_cuda(guard.err, "cudaSetDevice");
    char *d_k = nullptr;
    int32_t *d_r = nullptr;
    CK(cudaMalloc(reinterpret_cast<void **>(&d_k), n * (uint64_t)k));
    e = cudaMalloc(reinterpret_cast<void **>(&d_r), n * (uint64_t)h * 4);
    if (e != cudaSuccess) {
        cudaFree(d_k);
        return fail_cuda(e, "cudaMalloc");
    }
    int rc = 0;
    e = cudaMemcpy(d_k, kmers, n * (uint64_t)k, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        rc = bigsi_b200_hash_kmers_dev(d_k, n, k, h, m, canonical, d_r, nullptr);
        if (rc == 0) e = cudaMemcpy(rows_out, d_r, n * (uint64_t)h * 4, cudaMemcpyDeviceToHost);
    }
    cudaFree(d_k);
    cudaFree(d_r);
    if (rc) return rc;
    if (e != cudaSuccess) return fail_cuda(e, "hash_kmers copy");
    return 0;
}

// ============================================================================================
// device-pointer query entry points
// ============================================================================================
int bigsi_b200_query_dev(bigsi_b200_index *ix, int mode, const int32_t *d_rows, const int64_t *d_q_offsets,
                         uint64_t n_queries, uint64_t total_kmers, uint64_t max_query_kmers, int h, void *d_out,
                         uint64_t out_stride, void *stream)
{
    if (int rc = check_index(ix)) return rc;
    if (n_queries && (!d_q_offsets || !d_out)) return fail(BIGSI_B200_ERR_INVALID, "null device pointer");
    if (total_kmers && !d_rows) return fail(BIGSI_B200_ERR_INVALID, "null d_rows");
    DeviceGuard guard(ix->device);
    return run_query(ix, mode, d_rows, nullptr, 0, d_q_offsets, n_queries, total_kmers, max_query_kmers, h, d_out,
                     out_stride, static_cast<cudaStream_t>(stream));
}

int bigsi_b200_query_hits_dev(bigsi_b200_index *ix, const int32_t *d_rows, const int64_t *d_q_offsets,
                              uint64_t n_queries, uint64_t total_kmers, uint64_t max_query_kmers, int h,
                              const uint32_t *d_min_kmers, int32_t *d_cols_out, uint32_t *d_counts_out, uint64_t cap,
                              uint64_t *d_n_out, uint32_t *d_counts_full, uint64_t counts_stride, void *stream)
{
    if (int rc = check_index(ix)) return rc;
    if (n_queries == 0) return 0;
    if (!d_q_offsets || !d_min_kmers || !d_n_out || (cap && (!d_cols_out || !d_counts_out)))
        return fail(BIGSI_B200_ERR_INVALID, "null device pointer");
    if (total_kmers && !d_rows) return fail(BIGSI_B200_ERR_INVALID, "null d_rows");
    DeviceGuard guard(ix->device);
    HitsOut ho;
    ho.min_kmers = d_min_kmers;
    ho.cols = d_cols_out;
    ho.counts = d_counts_out;
    ho.n = reinterpret_cast<unsigned long long *>(d_n_out);
    ho.cap = cap;
    return run_query(ix, BIGSI_B200_MODE_COUNTS, d_rows, nullptr, 0, d_q_offsets, n_queries, total_kmers, max_query_kmers,
                     h, d_counts_full, counts_stride, static_cast<cudaStream_t>(stream), &ho);
}

int bigsi_b200_query_kmers_hits_dev(bigsi_b200_index *ix, const char *d_kmers, int k, const int64_t *d_q_offsets,
                                    uint64_t n_queries, uint64_t total_kmers, uint64_t max_query_kmers, int h,
                                    const uint32_t *d_min_kmers, int32_t *d_cols_out, uint32_t *d_counts_out,
                                    uint64_t cap, uint64_t *d_n_out, uint32_t *d_counts_full, uint64_t counts_stride,
                                    void *stream)
{
    if (int rc = check_index(ix)) return rc;
    if (n_queries == 0) return 0;
    if (!d_q_offsets || !d_min_kmers || !d_n_out || (cap && (!d_cols_out || !d_counts_out)))
        return fail(BIGSI_B200_ERR_INVALID, "null device pointer");
    if (total_kmers && !d_kmers) return fail(BIGSI_B200_ERR_INVALID, "null d_kmers");
    DeviceGuard guard(ix->device);
    HitsOut ho;
    ho.min_kmers = d_min_kmers;
    ho.cols = d_cols_out;
    ho.counts = d_counts_out;
    ho.n = reinterpret_cast<unsigned long long *>(d_n_out);
    ho.cap = cap;
    return run_query(ix, BIGSI_B200_MODE_COUNTS, nullptr, total_kmers ? d_kmers : nullptr, k, d_q_offsets, n_queries,
                     total_kmers, max_query_kmers, h, d_counts_full, counts_stride, static_cast<cudaStream_t>(stream),
                     &ho);
}

int bigsi_b200_query_kmers_hits_stream_dev(bigsi_b200_index *ix, const char *d_kmers, int k, uint64_t n_kmers, int h,
                                           uint32_t min_kmers, int32_t *d_cols_out, uint32_t *d_counts_out, uint64_t cap,
                                           uint64_t *d_n_out, void *stream)
{
    if (int rc = check_index(ix)) return rc;
    if (!d_n_out || (cap && (!d_cols_out || !d_counts_out))) return fail(BIGSI_B200_ERR_INVALID, "null device pointer");
    if (n_kmers == 0 || !d_kmers) return fail(BIGSI_B200_ERR_INVALID, "a streamed query needs at least one k-mer");
    DeviceGuard guard(ix->device);
    HitsOut ho;
    ho.by_value = true;
    ho.min_value = min_kmers;
    ho.cols = d_cols_out;
    ho.counts = d_counts_out;
    ho.n = reinterpret_cast<unsigned long long *>(d_n_out);
    ho.cap = cap;
    ho.deferred = true;
    // (a plan that is not the streamed one runs to completion in stream order like bigsi_b200_query_kmers_hits_dev)
    return run_query(ix, BIGSI_B200_MODE_COUNTS, nullptr, d_kmers, k, nullptr, 1, n_kmers, n_kmers, h, nullptr, 0,
                     static_cast<cudaStream_t>(stream), &ho);
}

int bigsi_b200_index_flush(bigsi_b200_index *ix)
{
    if (int rc = check_index(ix)) return rc;
    DeviceGuard guard(ix->device);
    return flush_pending(ix);
}

int bigsi_b200_lookup_dev(bigsi_b200_index *ix, const int32_t *d_rows, uint64_t n_kmers, int h, uint8_t *d_out,
                          uint64_t out_stride, void *stream)
{
    if (int rc = check_index(ix)) return rc;
    if (h < 1) return fail(BIGSI_B200_ERR_INVALID, "h must be >= 1");
    const uint64_t row_bytes = (ix->num_cols + 7) / 8;
    if (out_stride < row_bytes) return fail(BIGSI_B200_ERR_INVALID, "out_stride too small");
    if (n_kmers == 0 || row_bytes == 0) return 0;
    if (!d_rows || !d_out) return fail(BIGSI_B200_ERR_INVALID, "null device pointer");
    DeviceGuard guard(ix->device);
    CK(launch_lookup(ix->matrix, ix->pitch, (uint32_t)row_bytes, d_rows, n_kmers, h, d_out, out_stride,
                     static_cast<cudaStream_t>(stream)));
    ix->kernel_launches++;
    return 0;
}

int bigsi_b200_threshold_dev(const uint32_t *d_counts, uint64_t counts_stride, uint64_t n_queries, uint64_t num_cols,
                             const uint32_t *d_min_kmers, int32_t *d_cols_out, uint32_t *d_counts_out, uint64_t cap,
                             uint64_t *d_n_out, void *stream)
{
    if (n_queries == 0) return 0;
    if (!d_counts || !d_min_kmers || !d_n_out || (cap && (!d_cols_out || !d_counts_out)))
        return fail(BIGSI_B200_ERR_INVALID, "null device pointer");
    if (num_cols > 0xffffffffull) return fail(BIGSI_B200_ERR_INVALID, "too many columns");
    // grid.y carries the query index: batches beyond 65535 queries are split here
    for (uint64_t q0 = 0; q0 < n_queries; q0 += 65535) {
        const uint64_t nq = n_queries - q0 < 65535 ? n_queries - q0 : 65535;
        CK(launch_threshold(d_counts + q0 * counts_stride, counts_stride, nq, num_cols, d_min_kmers + q0,
                            d_cols_out + q0 * cap, d_counts_out + q0 * cap, cap,
                            reinterpret_cast<unsigned long long *>(d_n_out) + q0, static_cast<cudaStream_t>(stream)));
    }
    return 0;
}

// ============================================================================================
// host-buffer entry points
// ============================================================================================
static int search_common(bigsi_b200_index *ix, int mode, const char *kmers, const int32_t *rows, const int64_t *qoff,
                         uint64_t n_queries, int k, int h, uint64_t *total_out, uint64_t *longest_out,
                         uint64_t out_stride, const HitsOut *hits = nullptr)
{
    // stages inputs, hashes if needed and runs the query into ix->d_out (device); no D2H here
    uint64_t total = 0, longest = 0;
    if (int rc = check_offsets(qoff, n_queries, &total, &longest)) return rc;
    *total_out = total;
    *longest_out = longest;
    if (h < 1) return fail(BIGSI_B200_ERR_INVALID, "h must be >= 1");
    if (total && !kmers && !rows) return fail(BIGSI_B200_ERR_INVALID, "null query input");
    cudaError_t e;
    if ((e = ix->d_qoff.reserve((n_queries + 1) * 8)) != cudaSuccess) return fail_cuda(e, "staging");
    const uint64_t out_bytes = hits ? 0 : n_queries * out_stride * (mode == BIGSI_B200_MODE_COUNTS ? 4 : 1);
    if (!hits && (e = ix->d_out.reserve(out_bytes + 16)) != cudaSuccess) return fail_cuda(e, "staging");
    CK(cudaMemcpyAsync(ix->d_qoff.p, qoff, (n_queries + 1) * 8, cudaMemcpyHostToDevice, ix->stream));
    const char *d_kmers = nullptr;
    const int32_t *d_rows = nullptr;
    if (total) {
        if (kmers) {
            if (k < 1) return fail(BIGSI_B200_ERR_INVALID, "k must be >= 1");
            if ((e = ix->d_kmers.reserve(total * (uint64_t)k + 32)) != cudaSuccess) return fail_cuda(e, "staging");
            if ((e = ix->d_rows.reserve(total * (uint64_t)h * 4 + 4)) != cudaSuccess) return fail_cuda(e, "staging");
            CK(cudaMemcpyAsync(ix->d_kmers.p, kmers, total * (uint64_t)k, cudaMemcpyHostToDevice, ix->stream));
            d_kmers = static_cast<const char *>(ix->d_kmers.p);  // hashed in-kernel when the plan allows
        } else {
            for (uint64_t i = 0; i < total * (uint64_t)h; ++i)
                if (rows[i] < 0 || (uint64_t)rows[i] >= ix->num_rows)
                    return fail(BIGSI_B200_ERR_RANGE, "row id %d out of range at %llu", rows[i], (unsigned long long)i);
            if ((e = ix->d_rows.reserve(total * (uint64_t)h * 4 + 4)) != cudaSuccess) return fail_cuda(e, "staging");
            CK(cudaMemcpyAsync(ix->d_rows.p, rows, total * (uint64_t)h * 4, cudaMemcpyHostToDevice, ix->stream));
            d_rows = static_cast<const int32_t *>(ix->d_rows.p);
        }
    }
    return run_query(ix, mode, d_rows, d_kmers, k, static_cast<const int64_t *>(ix->d_qoff.p), n_queries, total, longest, h,
                     hits ? nullptr : ix->d_out.p, out_stride, ix->stream, hits);
}

static int search_full(bigsi_b200_index *ix, int mode, const char *kmers, const int32_t *rows, const int64_t *qoff,
                       uint64_t n_queries, int k, int h, void *out, uint64_t out_stride)
{
    if (int rc = check_index(ix)) return rc;
    if (n_queries == 0) return 0;
    if (!out) return fail(BIGSI_B200_ERR_INVALID, "null out");
    if (mode != BIGSI_B200_MODE_COUNTS && mode != BIGSI_B200_MODE_AND)
        return fail(BIGSI_B200_ERR_INVALID, "unknown query mode %d", mode);
    DeviceGuard guard(ix->device);
    uint64_t total = 0, longest = 0;
    if (int rc = search_common(ix, mode, kmers, rows, qoff, n_queries, k, h, &total, &longest, out_stride)) return rc;
    const uint64_t width = mode == BIGSI_B200_MODE_COUNTS ? ix->num_cols * 4 : (ix->num_cols + 7) / 8;
    const uint64_t stride_b = out_stride * (mode == BIGSI_B200_MODE_COUNTS ? 4 : 1);
    if (width)
        CK(cudaMemcpy2DAsync(out, stride_b, ix->d_out.p, stride_b, width, n_queries, cudaMemcpyDeviceToHost, ix->stream));
    CK(cudaStreamSynchronize(ix->stream));
    return 0;
}

int bigsi_b200_search_kmers(bigsi_b200_index *ix, int mode, const char *kmers, const int64_t *q_offsets,
                            uint64_t n_queries, int k, int h, void *out, uint64_t out_stride)
{
    return search_full(ix, mode, kmers, nullptr, q_offsets, n_queries, k, h, out, out_stride);
}

int bigsi_b200_search_rows(bigsi_b200_index *ix, int mode, const int32_t *rows, const int64_t *q_offsets,
                           uint64_t n_queries, int h, void *out, uint64_t out_stride)
{
    return search_full(ix, mode, nullptr, rows, q_offsets, n_queries, 0, h, out, out_stride);
}

// BIGSI.search for ONE query with no staging copies: the kernel reads the raw k-mers straight out of
// (mapped, pinned) host memory, takes the threshold by value and publishes the hit list itself into a
// mapped host block; the host polls that block's sequence word instead of synchronising the stream.
// Returns 1 when the path does not apply (the caller falls back to the staged path).
static int search_one_published(bigsi_b200_index *ix, const char *d_kmers, uint64_t total, int k, int h, uint32_t min_kmers,
                                int32_t *cols_out, uint32_t *counts_out, uint64_t cap, uint64_t *n_out,
                                const unsigned long long *total_dev = nullptr, const uint32_t *d_min = nullptr,
                                uint64_t *total_out = nullptr, unsigned long long *scrub = nullptr, uint64_t scrub_words = 0);

static int search_one_zero_copy(bigsi_b200_index *ix, const char *kmers, const int64_t *qoff, int k, int h,
                                uint32_t min_kmers, int32_t *cols_out, uint32_t *counts_out, uint64_t cap, uint64_t *n_out)
{
    if (!qoff || qoff[0] != 0 || qoff[1] <= 0 || !kmers || k < 1 || h < 1 || ix->num_cols == 0) return 1;
    const uint64_t total = (uint64_t)qoff[1];
    cudaError_t e;
    // k-mers: use the caller's buffer when the device can address it, else one memcpy into our pinned buffer
    const char *d_kmers = nullptr;
    cudaPointerAttributes attr;
    e = cudaPointerGetAttributes(&attr, kmers);
    if (e == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer) {
        d_kmers = static_cast<const char *>(attr.devicePointer);
    } else {
        (void)cudaGetLastError();
        if ((e = ix->h_kmers.reserve(total * (uint64_t)k + 64)) != cudaSuccess) return fail_cuda(e, "pinned staging");
        memcpy(ix->h_kmers.p, kmers, total * (uint64_t)k);
        void *dp = nullptr;
        CK(cudaHostGetDevicePointer(&dp, ix->h_kmers.p, 0));
        d_kmers = static_cast<const char *>(dp);
    }
    return search_one_published(ix, d_kmers, total, k, h, min_kmers, cols_out, counts_out, cap, n_out);
}

// One query whose unique raw k-mers are device-addressable: launch with the threshold by value, let the
// kernel publish the hit list into the mapped host block and poll it.
// total_dev / d_min (query front-end): the number of k-mers (<= total) and the threshold are device words a
// preceding kernel in the stream writes; *total_out receives the number the kernel saw.  Returns 1 without
// launching anything when the launch plan cannot follow a device-side count.
static int search_one_published(bigsi_b200_index *ix, const char *d_kmers, uint64_t total, int k, int h, uint32_t min_kmers,
                                int32_t *cols_out, uint32_t *counts_out, uint64_t cap, uint64_t *n_out,
                                const unsigned long long *total_dev, const uint32_t *d_min, uint64_t *total_out,
                                unsigned long long *scrub, uint64_t scrub_words)
{
    cudaError_t e;
    const uint64_t spec = cap < 1024 ? cap : 1024;  // hits the host block holds; longer lists are fetched afterwards
    if ((e = ix->h_sink.reserve(16 + 2 * spec * 4 + 64)) != cudaSuccess) return fail_cuda(e, "pinned result block");
    if ((e = ix->d_nhits.reserve(8 + 2 * cap * 4 + 16)) != cudaSuccess) return fail_cuda(e, "staging");
    volatile unsigned long long *blk = static_cast<volatile unsigned long long *>(ix->h_sink.p);
    void *d_blk = nullptr;
    CK(cudaHostGetDevicePointer(&d_blk, ix->h_sink.p, 0));
    uint8_t *dev = static_cast<uint8_t *>(ix->d_nhits.p);
    bool published = false;
    HitsOut ho;
    ho.n = reinterpret_cast<unsigned long long *>(dev);
    ho.cols = reinterpret_cast<int32_t *>(dev + 8);
    ho.counts = reinterpret_cast<uint32_t *>(dev + 8 + cap * 4);
    ho.cap = cap;
    if (d_min) {
        ho.min_kmers = d_min;
    } else {
        ho.by_value = true;
        ho.min_value = min_kmers;
    }
    ho.total_dev = total_dev;
    ho.scrub = scrub;
    ho.scrub_words = scrub_words;
    ho.isolated = true;
    ho.inputs_ready = total_dev == nullptr;  // host-written (or staged before this call) k-mers: nothing in the stream produces them
    ho.n_sinks = 1;
    ho.sinks[0] = static_cast<unsigned long long *>(d_blk);
    ho.sink_spec = (uint32_t)spec;
    ho.sink_seq = ix->sink_seq + 1;
    ho.published = &published;
    if (int rc = run_query(ix, BIGSI_B200_MODE_COUNTS, nullptr, d_kmers, k, nullptr, 1, total, total, h, nullptr, 0, ix->stream,
                           &ho))
        return rc;
    ++ix->sink_seq;
    uint64_t n = 0;
    if (total_dev && !published) return fail(BIGSI_B200_ERR_CUDA, "internal: front-end launch without publication");
    if (published) {
        // poll the sequence word; look at the stream now and then so that a failed launch cannot hang us
        uint64_t spins = 0;
        while (__atomic_load_n(&blk[0], __ATOMIC_ACQUIRE) != ho.sink_seq) {
            if ((++spins & 0x3fff) == 0) {
                if (const unsigned long long av = abort_state(ix)) return fail_aborted(av);
                e = cudaStreamQuery(ix->stream);
                if (e != cudaErrorNotReady) {
                    if (e != cudaSuccess) return fail_cuda(e, "query kernel");
                    if (__atomic_load_n(&blk[0], __ATOMIC_ACQUIRE) != ho.sink_seq)
                        return fail(BIGSI_B200_ERR_CUDA, "query kernel finished without publishing its result");
                }
            }
        }
        // (acquire load above: the payload behind the sequence word is read after it, also on weakly ordered hosts)
        n = blk[1];
        if (total_out) *total_out = total_dev ? blk[2 + spec] : total;
        const uint64_t m = n < cap ? n : cap;
        const uint64_t ms = m < spec ? m : spec;
        const int32_t *hc = reinterpret_cast<const int32_t *>(const_cast<unsigned long long *>(blk) + 2);
        memcpy(cols_out, hc, ms * 4);
        memcpy(counts_out, hc + spec, ms * 4);
        if (m > spec) {  // a long hit list: the device buffers are complete once the block was published
            CK(cudaMemcpyAsync(cols_out, ho.cols, m * 4, cudaMemcpyDeviceToHost, ix->stream));
            CK(cudaMemcpyAsync(counts_out, ho.counts, m * 4, cudaMemcpyDeviceToHost, ix->stream));
            CK(cudaStreamSynchronize(ix->stream));
        }
    } else {
        CK(cudaMemcpyAsync(&n, ho.n, 8, cudaMemcpyDeviceToHost, ix->stream));
        CK(cudaStreamSynchronize(ix->stream));
        const uint64_t m = n < cap ? n : cap;
        if (m) {
            CK(cudaMemcpyAsync(cols_out, ho.cols, m * 4, cudaMemcpyDeviceToHost, ix->stream));
            CK(cudaMemcpyAsync(counts_out, ho.counts, m * 4, cudaMemcpyDeviceToHost, ix->stream));
            CK(cudaStreamSynchronize(ix->stream));
        }
    }
    n_out[0] = n;
    return 0;
}

int bigsi_b200_search_kmers_hits(bigsi_b200_index *ix, const char *kmers, const int64_t *q_offsets, uint64_t n_queries,
                                 int k, int h, const uint32_t *min_kmers, int32_t *cols_out, uint32_t *counts_out,
                                 uint64_t cap, uint64_t *n_out)
{
    if (int rc = check_index(ix)) return rc;
    if (n_queries == 0) return 0;
    if (!min_kmers || !n_out || (cap && (!cols_out || !counts_out))) return fail(BIGSI_B200_ERR_INVALID, "null pointer");
    DeviceGuard guard(ix->device);
    cudaError_t e;
    if (n_queries == 1 && ix->opt_zero_copy != 0) {
        int rc = search_one_zero_copy(ix, kmers, q_offsets, k, h, min_kmers[0], cols_out, counts_out, cap, n_out);
        if (rc != 1) return rc;  // 1 = not applicable, take the staged path below
    }
    // one device block [n_hits: Q x u64][cols: Q x cap][counts: Q x cap] so that small results come
    // back in ONE device-to-host copy
    const uint64_t n_bytes = n_queries * 8, list_bytes = n_queries * cap * 4;
    if ((e = ix->d_min.reserve(n_queries * 4)) != cudaSuccess) return fail_cuda(e, "staging");
    if ((e = ix->d_nhits.reserve(n_bytes + 2 * list_bytes + 16)) != cudaSuccess) return fail_cuda(e, "staging");
    uint8_t *blk = static_cast<uint8_t *>(ix->d_nhits.p);
    HitsOut ho;
    ho.min_kmers = static_cast<const uint32_t *>(ix->d_min.p);
    ho.n = reinterpret_cast<unsigned long long *>(blk);
    ho.cols = reinterpret_cast<int32_t *>(blk + n_bytes);
    ho.counts = reinterpret_cast<uint32_t *>(blk + n_bytes + list_bytes);
    ho.cap = cap;
    CK(cudaMemcpyAsync(ix->d_min.p, min_kmers, n_queries * 4, cudaMemcpyHostToDevice, ix->stream));
    uint64_t total = 0, longest = 0;
    if (int rc = search_common(ix, BIGSI_B200_MODE_COUNTS, kmers, nullptr, q_offsets, n_queries, k, h, &total, &longest, 0,
                               &ho))
        return rc;
    // Single query (the BIGSI.search case): speculatively fetch the count and the first hits in one copy.
    const uint64_t spec = cap < 64 ? cap : 64;
    if (n_queries == 1) {
        if ((e = ix->h_small.reserve(8 + 2 * spec * 4)) != cudaSuccess) return fail_cuda(e, "pinned staging");
        uint8_t *hs = static_cast<uint8_t *>(ix->h_small.p);
        CK(cudaMemcpyAsync(hs, blk, 8, cudaMemcpyDeviceToHost, ix->stream));
        if (spec) {
            CK(cudaMemcpyAsync(hs + 8, ho.cols, spec * 4, cudaMemcpyDeviceToHost, ix->stream));
            CK(cudaMemcpyAsync(hs + 8 + spec * 4, ho.counts, spec * 4, cudaMemcpyDeviceToHost, ix->stream));
        }
        CK(cudaStreamSynchronize(ix->stream));
        memcpy(n_out, hs, 8);
        const uint64_t n = n_out[0] < cap ? n_out[0] : cap;
        if (n <= spec) {
            memcpy(cols_out, hs + 8, n * 4);
            memcpy(counts_out, hs + 8 + spec * 4, n * 4);
        } else {
            CK(cudaMemcpyAsync(cols_out, ho.cols, n * 4, cudaMemcpyDeviceToHost, ix->stream));
            CK(cudaMemcpyAsync(counts_out, ho.counts, n * 4, cudaMemcpyDeviceToHost, ix->stream));
            CK(cudaStreamSynchronize(ix->stream));
        }
        return 0;
    }
    CK(cudaMemcpyAsync(n_out, blk, n_bytes, cudaMemcpyDeviceToHost, ix->stream));
    CK(cudaStreamSynchronize(ix->stream));
    if (cap) {
        uint64_t max_n = 0;
        for (uint64_t q = 0; q < n_queries; ++q) {
            const uint64_t n = n_out[q] < cap ? n_out[q] : cap;
            if (n > max_n) max_n = n;
        }
        if (max_n) {
            CK(cudaMemcpy2DAsync(cols_out, cap * 4, ho.cols, cap * 4, max_n * 4, n_queries, cudaMemcpyDeviceToHost,
                                 ix->stream));
            CK(cudaMemcpy2DAsync(counts_out, cap * 4, ho.counts, cap * 4, max_n * 4, n_queries, cudaMemcpyDeviceToHost,
                                 ix->stream));
            CK(cudaStreamSynchronize(ix->stream));
        }
    }
    return 0;
}

int bigsi_b200_lookup_kmers(bigsi_b200_index *ix, const char *kmers, uint64_t n, int k, int h, uint8_t *out,
                            uint64_t out_stride)
{
    if (int rc = check_index(ix)) return rc;
    if (n == 0) return 0;
    if (!kmers || !out) return fail(BIGSI_B200_ERR_INVALID, "null pointer");
    if (k < 1 || h < 1) return fail(BIGSI_B200_ERR_INVALID, "k and h must be >= 1");
    const uint64_t row_bytes = (ix->num_cols + 7) / 8;
    if (out_stride < row_bytes) return fail(BIGSI_B200_ERR_INVALID, "out_stride too small");
    if (row_bytes == 0) return 0;
    DeviceGuard guard(ix->device);
    cudaError_t e;
    const uint64_t dstride = round_up(row_bytes, 16);
    if ((e = ix->d_kmers.reserve(n * (uint64_t)k)) != cudaSuccess) return fail_cuda(e, "staging");
    if ((e = ix->d_rows.reserve(n * (uint64_t)h * 4)) != cudaSuccess) return fail_cuda(e, "staging");
    if ((e = ix->d_out.reserve(n * dstride)) != cudaSuccess) return fail_cuda(e, "staging");
    CK(cudaMemcpyAsync(ix->d_kmers.p, kmers, n * (uint64_t)k, cudaMemcpyHostToDevice, ix->stream));
    if (int rc = bigsi_b200_hash_kmers_dev(static_cast<const char *>(ix->d_kmers.p), n, k, h, ix->num_rows, 1,
                                           static_cast<int32_t *>(ix->d_rows.p), ix->stream))
        return rc;
    ix->kernel_launches++;
    if (int rc = bigsi_b200_lookup_dev(ix, static_cast<const int32_t *>(ix->d_rows.p), n, h,
                                       static_cast<uint8_t *>(ix->d_out.p), dstride, ix->stream))
        return rc;
    CK(cudaMemcpy2DAsync(out, out_stride, ix->d_out.p, dstride, row_bytes, n, cudaMemcpyDeviceToHost, ix->stream));
    CK(cudaStreamSynchronize(ix->stream));
    return 0;
}

// ============================================================================================
// query front-end + search: BIGSI.search's filter stage for one sequence (graph/bigsi.py:174-230)
// ============================================================================================
// The general path: a front-end kernel (aux_kernels.cu:dedup_windows_kernel) compacts the unique windows, the search
// follows -- streamed without a host round trip when the plan allows, else after fetching U.  Used for k > 32, for
// sequences too long for the streamed plan and for rows wider than one column tile.
static int search_sequence_general(bigsi_b200_index *ix, const char *seq, uint64_t len, int k, int h, double threshold,
                                   int32_t *cols_out, uint32_t *counts_out, uint64_t cap, uint64_t *n_hits_out,
                                   uint64_t *num_kmers_out)
{
    const uint64_t n = len - (uint64_t)k + 1;
    cudaError_t e;
    uint64_t T = 1024;
    while (T < 2 * n) T <<= 1;
    // [table: T x u64][U: u64][ticket: u32][min_kmers: u32]; the block is cleared behind every query, so it is
    // clean on entry unless it has just been (re)allocated
    const uint64_t tail = T * 8, clear_bytes = tail + 16;
    const uint64_t old_cap = ix->d_table.cap;
    if ((e = ix->h_kmers.reserve(len + 64)) != cudaSuccess) return fail_cuda(e, "pinned sequence staging");
    if ((e = ix->d_table.reserve(clear_bytes + 64)) != cudaSuccess) return fail_cuda(e, "de-duplication table");
    if ((e = ix->d_kmers.reserve(n * (uint64_t)k + 64)) != cudaSuccess) return fail_cuda(e, "k-mer staging");
    if ((e = ix->h_small.reserve(64)) != cudaSuccess) return fail_cuda(e, "pinned staging");
    if (ix->d_table.cap != old_cap || ix->table_clean_bytes < clear_bytes) {
        CK(cudaMemsetAsync(ix->d_table.p, 0, ix->d_table.cap, ix->stream));
        ix->table_clean_bytes = ix->d_table.cap;
    }
    uint8_t *tb = static_cast<uint8_t *>(ix->d_table.p);
    unsigned long long *d_counter = reinterpret_cast<unsigned long long *>(tb + tail);
    unsigned int *d_ticket = reinterpret_cast<unsigned int *>(tb + tail + 8);
    uint32_t *d_min = reinterpret_cast<uint32_t *>(tb + tail + 12);
    // the sequence stays in (mapped, pinned) host memory: the front-end kernel stages its spans itself
    memcpy(ix->h_kmers.p, seq, len);
    void *d_seq = nullptr;
    CK(cudaHostGetDevicePointer(&d_seq, ix->h_kmers.p, 0));
    CK(launch_dedup_windows(static_cast<const uint8_t *>(d_seq), n, k, static_cast<unsigned long long *>(ix->d_table.p), T,
                            static_cast<uint8_t *>(ix->d_kmers.p), d_counter, d_ticket, threshold, d_min, ix->stream));
    ix->kernel_launches++;
    ix->table_clean_bytes = 0;
    int rc = 1;
    uint64_t U = 0;
    bool scrubbed = false;
    if (ix->num_cols && ix->opt_zero_copy != 0) {
        // no host round trip: the search kernel reads U and min_kmers = ceil(U * threshold) from the device; it also
        // clears the table (T words) and the count / ticket / threshold words behind it for the next query
        rc = search_one_published(ix, static_cast<const char *>(ix->d_kmers.p), n, k, h, 0, cols_out, counts_out, cap, n_hits_out,
                                  d_counter, d_min, &U, static_cast<unsigned long long *>(ix->d_table.p), T);
        scrubbed = rc == 0;
    }
    if (rc == 1) {  // the plan cannot follow a device-side count (or the shard is empty): fetch U, then search
        CK(cudaMemcpyAsync(ix->h_small.p, d_counter, 8, cudaMemcpyDeviceToHost, ix->stream));
        CK(cudaStreamSynchronize(ix->stream));
        U = *static_cast<const unsigned long long *>(ix->h_small.p);
        rc = 0;
        if (ix->num_cols) {
            // min_kmers = math.ceil(U * threshold) in IEEE double (graph/bigsi.py:179); <= 0 keeps every sample
            const double need = ceil((double)U * threshold);
            const uint32_t min_kmers = need <= 0.0 ? 0u : need >= 4294967295.0 ? 0xffffffffu : (uint32_t)need;
            rc = search_one_published(ix, static_cast<const char *>(ix->d_kmers.p), U, k, h, min_kmers, cols_out, counts_out, cap,
                                      n_hits_out);
        }
    }
    if (rc) return rc;
    *num_kmers_out = U;
    // clear the table for the next query now, off its critical path (the fast path's kernel has done it itself)
    if (!scrubbed) CK(cudaMemsetAsync(ix->d_table.p, 0, clear_bytes, ix->stream));
    ix->table_clean_bytes = clear_bytes;
    return 0;
}


// --------------------------------------------------------------------------------------------
// sequence searches with the front-end inside the gather kernel: submit / wait
// --------------------------------------------------------------------------------------------
static constexpr uint64_t kTicketSpec = 1024;  // hits one mapped host result block holds
static constexpr uint64_t kTicketBlock = (16 + 8 * kTicketSpec + 8 + 127) / 128 * 128;

static int seq_submit(bigsi_b200_index *ix, const char *seq, uint64_t len, int k, int h, double threshold, uint64_t cap,
                      bool isolated, uint64_t *ticket_out)
{
    if (k < 1 || h < 1) return fail(BIGSI_B200_ERR_INVALID, "k and h must be >= 1");
    if (len && !seq) return fail(BIGSI_B200_ERR_INVALID, "null sequence");
    if (const unsigned long long av = abort_state(ix)) return fail_aborted(av);
    const uint64_t id = ix->next_ticket;
    const int slot = (int)(id % bigsi_b200_index::kSeqTickets);
    bigsi_b200_index::SeqTicket &t = ix->tickets[slot];
    if (t.pending) return fail(BIGSI_B200_ERR_INVALID, "too many sequence searches in flight (at most %d): wait for ticket %llu first",
                               bigsi_b200_index::kSeqTickets, (unsigned long long)t.id);
    t = bigsi_b200_index::SeqTicket();
    t.id = id;
    t.k = k;
    t.h = h;
    t.threshold = threshold;
    t.cap = cap;
    t.spec = cap < kTicketSpec ? cap : kTicketSpec;
    auto defer = [&]() {  // not streamable: searched synchronously when the caller waits for it
        t.deferred = true;
        t.seq.assign(seq ? seq : "", len);
        t.pending = true;
        ++ix->next_ticket;
        *ticket_out = id;
        return 0;
    };
    if (len < (uint64_t)k || ix->num_cols == 0 || k > 32 || ix->opt_zero_copy == 0 || cap == 0) return defer();
    const uint64_t n = len - (uint64_t)k + 1;
    if (n > 0xfffffff0ull) return fail(BIGSI_B200_ERR_RANGE, "sequence too long");
    cudaError_t e;
    // tables: one per ring slot, 2 entries per window (a power of two), zeroed when (re)allocated
    uint64_t T = 1024;
    while (T < 2 * n) T <<= 1;
    if (T > (1ull << 26)) return defer();  // (longer sequences never get the streamed plan anyway)
    if (T > ix->seq_table_entries) {
        CK(cudaStreamSynchronize(ix->stream));
        if ((e = ix->d_seq_tables.reserve(kStreamRing * T * 8)) != cudaSuccess) return fail_cuda(e, "de-duplication tables");
        CK(cudaMemsetAsync(ix->d_seq_tables.p, 0, kStreamRing * T * 8, ix->stream));
        ix->seq_table_entries = T;
        for (auto &u : ix->seq_table_uses) u = 0;
    }
    const uint64_t hit_slot = round_up(8 + 8ull * cap + 16, 256);
    if (hit_slot * kStreamStates > ix->d_hits_ring.cap) {
        if (int rc = flush_pending(ix)) return rc;  // its hit buffers live in the ring that is about to go
        CK(cudaStreamSynchronize(ix->stream));
        if ((e = ix->d_hits_ring.reserve(hit_slot * kStreamStates)) != cudaSuccess) return fail_cuda(e, "staging");
    }
    if ((e = ix->h_tsink.reserve(kTicketBlock * bigsi_b200_index::kSeqTickets)) != cudaSuccess) return fail_cuda(e, "pinned result blocks");
    if ((e = ix->h_seq[slot].reserve(len + 64)) != cudaSuccess) return fail_cuda(e, "pinned sequence staging");
    // the sequence stays in (mapped, pinned) host memory: every gather CTA stages its own span (one PCIe round trip)
    memcpy(ix->h_seq[slot].p, seq, len);
    memset(static_cast<uint8_t *>(ix->h_seq[slot].p) + len, 0, 64);
    void *d_seq = nullptr, *d_blk = nullptr;
    CK(cudaHostGetDevicePointer(&d_seq, ix->h_seq[slot].p, 0));
    CK(cudaHostGetDevicePointer(&d_blk, static_cast<uint8_t *>(ix->h_tsink.p) + (uint64_t)slot * kTicketBlock, 0));
    uint8_t *dev = static_cast<uint8_t *>(ix->d_hits_ring.p) + ((ix->stream_seq + 1) % kStreamStates) * hit_slot;
    bool published = false;
    HitsOut ho;
    ho.n = reinterpret_cast<unsigned long long *>(dev);
    ho.cols = reinterpret_cast<int32_t *>(dev + 8);
    ho.counts = reinterpret_cast<uint32_t *>(dev + 8 + cap * 4);
    ho.cap = cap;
    ho.seq_mode = true;
    ho.seq_threshold = threshold;
    ho.isolated = isolated;
    ho.deferred = !isolated;  // bulk searches: stage 2 rides in the next sequence's gather kernel; seq_wait flushes the last one
    ho.ticket = id;
    ho.inputs_ready = true;
    ho.n_sinks = 1;
    ho.sinks[0] = static_cast<unsigned long long *>(d_blk);
    ho.sink_spec = (uint32_t)t.spec;
    ho.sink_seq = id;
    ho.published = &published;
    const int rc = run_query(ix, BIGSI_B200_MODE_COUNTS, nullptr, static_cast<const char *>(d_seq), k, nullptr, 1, n, n, h, nullptr, 0,
                             ix->stream, &ho);
    if (rc == 1) return defer();  // the launch plan is not the streamed one (wide rows, very long sequence)
    if (rc) return rc;
    if (!published) return fail(BIGSI_B200_ERR_CUDA, "internal: sequence search launched without publication");
    t.d_hits = dev;
    t.pending = true;
    ++ix->next_ticket;
    *ticket_out = id;
    return 0;
}

static int seq_wait(bigsi_b200_index *ix, uint64_t ticket, int32_t *cols_out, uint32_t *counts_out, uint64_t cap,
                    uint64_t *n_hits_out, uint64_t *num_kmers_out)
{
    const int slot = (int)(ticket % bigsi_b200_index::kSeqTickets);
    bigsi_b200_index::SeqTicket &t = ix->tickets[slot];
    if (!t.pending || t.id != ticket) return fail(BIGSI_B200_ERR_INVALID, "unknown or already collected ticket %llu", (unsigned long long)ticket);
    if (cap < t.cap) return fail(BIGSI_B200_ERR_INVALID, "output capacity %llu smaller than the one submitted (%llu)",
                                 (unsigned long long)cap, (unsigned long long)t.cap);
    t.pending = false;
    *n_hits_out = 0;
    *num_kmers_out = 0;
    if (t.deferred) {
        if (t.seq.size() < (size_t)t.k) return 0;  // no window: the caller reproduces the reference's TypeError
        std::string seq;
        seq.swap(t.seq);
        return search_sequence_general(ix, seq.data(), seq.size(), t.k, t.h, t.threshold, cols_out, counts_out, t.cap, n_hits_out,
                                       num_kmers_out);
    }
    // the newest search has nobody behind it to run its stage 2: flush it
    if (ix->pending.have && ix->pending.ticket == ticket)
        if (int rc = flush_pending(ix)) return rc;
    volatile unsigned long long *blk =
        reinterpret_cast<volatile unsigned long long *>(static_cast<uint8_t *>(ix->h_tsink.p) + (uint64_t)slot * kTicketBlock);
    uint64_t spins = 0;
    while (__atomic_load_n(&blk[0], __ATOMIC_ACQUIRE) != ticket) {
        if ((++spins & 0x3fff) == 0) {
            if (const unsigned long long av = abort_state(ix)) return fail_aborted(av);
            cudaError_t e = cudaStreamQuery(ix->stream);
            if (e != cudaErrorNotReady) {
                if (e != cudaSuccess) return fail_cuda(e, "query kernel");
                if (__atomic_load_n(&blk[0], __ATOMIC_ACQUIRE) != ticket)
                    return fail(BIGSI_B200_ERR_CUDA, "query kernels finished without publishing their result");
            }
        }
    }
    const uint64_t n = blk[1];
    *num_kmers_out = blk[2 + t.spec];
    const uint64_t m = n < t.cap ? n : t.cap;
    const uint64_t ms = m < t.spec ? m : t.spec;
    const int32_t *hc = reinterpret_cast<const int32_t *>(const_cast<unsigned long long *>(blk) + 2);
    memcpy(cols_out, hc, ms * 4);
    memcpy(counts_out, hc + t.spec, ms * 4);
    if (m > t.spec) {  // a long hit list: the device buffers are complete once the block was published
        CK(cudaMemcpyAsync(cols_out, t.d_hits + 8, m * 4, cudaMemcpyDeviceToHost, ix->stream));
        CK(cudaMemcpyAsync(counts_out, t.d_hits + 8 + t.cap * 4, m * 4, cudaMemcpyDeviceToHost, ix->stream));
        CK(cudaStreamSynchronize(ix->stream));
    }
    *n_hits_out = n;
    return 0;
}

int bigsi_b200_search_sequence_submit(bigsi_b200_index *ix, const char *seq, uint64_t len, int k, int h, double threshold,
                                      uint64_t cap, uint64_t *ticket_out)
{
    if (int rc = check_index(ix)) return rc;
    if (!ticket_out) return fail(BIGSI_B200_ERR_INVALID, "null pointer");
    DeviceGuard guard(ix->device);
    return seq_submit(ix, seq, len, k, h, threshold, cap, false, ticket_out);
}

int bigsi_b200_search_sequence_wait(bigsi_b200_index *ix, uint64_t ticket, int32_t *cols_out, uint32_t *counts_out, uint64_t cap,
                                    uint64_t *n_hits_out, uint64_t *num_kmers_out)
{
    if (int rc = check_index(ix)) return rc;
    if (!n_hits_out || !num_kmers_out || (cap && (!cols_out || !counts_out))) return fail(BIGSI_B200_ERR_INVALID, "null pointer");
    DeviceGuard guard(ix->device);
    return seq_wait(ix, ticket, cols_out, counts_out, cap, n_hits_out, num_kmers_out);
}

int bigsi_b200_search_sequence(bigsi_b200_index *ix, const char *seq, uint64_t len, int k, int h, double threshold,
                               int32_t *cols_out, uint32_t *counts_out, uint64_t cap, uint64_t *n_hits_out,
                               uint64_t *num_kmers_out)
{
    if (int rc = check_index(ix)) return rc;
    if (!n_hits_out || !num_kmers_out || (cap && (!cols_out || !counts_out))) return fail(BIGSI_B200_ERR_INVALID, "null pointer");
    DeviceGuard guard(ix->device);
    uint64_t ticket = 0;
    if (int rc = seq_submit(ix, seq, len, k, h, threshold, cap, /*isolated=*/true, &ticket)) return rc;
    return seq_wait(ix, ticket, cols_out, counts_out, cap, n_hits_out, num_kmers_out);
}

// bulk_search (bigsi/__main__.py:261-314: every record of a FASTA file through BIGSI.search): n_seqs sequences in one
// call, up to kSeqTickets - 1 searches in flight -- the gather kernel of one sequence overlaps the reduce kernel of
// the one before, the host stages the next sequence meanwhile.
int bigsi_b200_search_sequences(bigsi_b200_index *ix, const char *seqs, const uint64_t *offsets, uint64_t n_seqs, int k, int h,
                                double threshold, int32_t *cols_out, uint32_t *counts_out, uint64_t cap, uint64_t *n_hits_out,
                                uint64_t *num_kmers_out)
{
    if (int rc = check_index(ix)) return rc;
    if (n_seqs == 0) return 0;
    if (!offsets || !n_hits_out || !num_kmers_out || (cap && (!cols_out || !counts_out))) return fail(BIGSI_B200_ERR_INVALID, "null pointer");
    for (uint64_t q = 0; q < n_seqs; ++q)
        if (offsets[q + 1] < offsets[q]) return fail(BIGSI_B200_ERR_INVALID, "offsets must be non-decreasing");
    if (offsets[n_seqs] > offsets[0] && !seqs) return fail(BIGSI_B200_ERR_INVALID, "null sequences");
    DeviceGuard guard(ix->device);
    const uint64_t window = bigsi_b200_index::kSeqTickets - 1;
    std::vector<uint64_t> tickets(n_seqs, 0);
    uint64_t submitted = 0, collected = 0;
    int rc = 0;
    while (collected < n_seqs && !rc) {
        while (submitted < n_seqs && submitted - collected < window && !rc) {
            rc = seq_submit(ix, seqs + offsets[submitted], offsets[submitted + 1] - offsets[submitted], k, h, threshold, cap, false,
                            &tickets[submitted]);
            if (!rc) ++submitted;
        }
        if (rc) break;
        rc = seq_wait(ix, tickets[collected], cols_out + collected * cap, counts_out + collected * cap, cap, n_hits_out + collected,
                      num_kmers_out + collected);
        ++collected;
    }
    // on an error the searches still in flight are abandoned: their tickets are released once the stream is idle
    if (rc) {
        const std::string msg = g_err;
        cudaStreamSynchronize(ix->stream);
        for (auto &t : ix->tickets) t.pending = false;
        g_err = msg;
    }
    return rc;
}

// ============================================================================================
// build path: Bloom filters on the device, N x m -> m x N transpose
// ============================================================================================
int bigsi_b200_bloom_kmers(int device, const char *kmers, uint64_t n, int k, int h, uint64_t m, int canonical,
                           uint8_t *bloom_out)
{
    if (k < 1 || h < 1) return fail(BIGSI_B200_ERR_INVALID, "k and h must be >= 1");
    if (m == 0 || m > 0x7fffffffull) return fail(BIGSI_B200_ERR_INVALID, "m must be in [1, 2^31-1]");
    if (!bloom_out || (n && !kmers)) return fail(BIGSI_B200_ERR_INVALID, "null pointer");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess) return fail_cuda(e, "cudaGetDeviceCount");
    if (ndev == 0) return fail(BIGSI_B200_ERR_NO_DEVICE, "no CUDA device");
    if (device < 0 || device >= ndev) return fail(BIGSI_B200_ERR_INVALID, "device out of range");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail_cuda(guard.err, "cudaSetDevice");
    const uint64_t nbytes = (m + 7) / 8, words_bytes = round_up(nbytes, 4);
    DevBuf d_k, d_r, d_b;
    int rc = 0;
    auto done = [&](int code) {
        d_k.release();
        d_r.release();
        d_b.release();
        return code;
    };
    if ((e = d_b.reserve(words_bytes)) != cudaSuccess) return done(fail_cuda(e, "bloom buffer"));
    if ((e = cudaMemsetAsync(d_b.p, 0, words_bytes, nullptr)) != cudaSuccess) return done(fail_cuda(e, "memset"));
    if (n) {
        if ((e = d_k.reserve(n * (uint64_t)k)) != cudaSuccess) return done(fail_cuda(e, "k-mer staging"));
        if ((e = d_r.reserve(n * (uint64_t)h * 4)) != cudaSuccess) return done(fail_cuda(e, "row-id staging"));
        if ((e = cudaMemcpyAsync(d_k.p, kmers, n * (uint64_t)k, cudaMemcpyHostToDevice, nullptr)) != cudaSuccess)
            return done(fail_cuda(e, "k-mer copy"));
        if ((rc = bigsi_b200_hash_kmers_dev(static_cast<const char *>(d_k.p), n, k, h, m, canonical, static_cast<int32_t *>(d_r.p),
                                            nullptr)))
            return done(rc);
        if ((e = launch_bloom_set_bits(static_cast<const int32_t *>(d_r.p), n * (uint64_t)h, static_cast<uint8_t *>(d_b.p),
                                       nullptr)) != cudaSuccess)
            return done(fail_cuda(e, "bloom_set_bits"));
    }
    e = cudaMemcpy(bloom_out, d_b.p, nbytes, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return done(fail_cuda(e, "bloom copy"));
    return done(0);
}

static int build_columns_check(bigsi_b200_index *ix, uint64_t col0, uint64_t n_blooms, uint64_t n_bits)
{
    if (n_bits > ix->num_rows) return fail(BIGSI_B200_ERR_RANGE, "bloom filters longer than m");
    if (col0 > ix->num_cols) return fail(BIGSI_B200_ERR_RANGE, "first column %llu beyond num_cols=%llu", (unsigned long long)col0,
                                         (unsigned long long)ix->num_cols);
    if (col0 + n_blooms > ix->col_capacity)
        return fail(BIGSI_B200_ERR_RANGE, "columns [%llu,%llu) exceed the column capacity %llu", (unsigned long long)col0,
                    (unsigned long long)(col0 + n_blooms), (unsigned long long)ix->col_capacity);
    return 0;
}

int bigsi_b200_index_build_columns_dev(bigsi_b200_index *ix, uint64_t col0, uint64_t n_blooms, const uint8_t *d_blooms,
                                       uint64_t bloom_stride, uint64_t n_bits, void *stream)
{
    if (int rc = check_index(ix)) return rc;
    if (n_blooms == 0) return 0;
    if (!d_blooms) return fail(BIGSI_B200_ERR_INVALID, "null bloom filters");
    if (int rc = build_columns_check(ix, col0, n_blooms, n_bits)) return rc;
    if ((bloom_stride & 31) || bloom_stride < (ix->num_rows + 255) / 256 * 32 || (reinterpret_cast<uintptr_t>(d_blooms) & 15))
        return fail(BIGSI_B200_ERR_INVALID, "device bloom filters need a 16-byte aligned base and a stride that is a multiple "
                                            "of 32 bytes and >= ceil(m/256)*32");
    DeviceGuard guard(ix->device);
    CK(launch_transpose_blooms(ix->matrix, ix->pitch, ix->num_rows, col0, n_blooms, d_blooms, bloom_stride, n_bits,
                               static_cast<cudaStream_t>(stream)));
    ix->kernel_launches++;
    if (col0 + n_blooms > ix->num_cols) ix->num_cols = col0 + n_blooms;
    return 0;
}

int bigsi_b200_index_build_columns(bigsi_b200_index *ix, uint64_t col0, uint64_t n_blooms, const uint8_t *blooms,
                                   uint64_t bloom_stride, uint64_t n_bits)
{
    if (int rc = check_index(ix)) return rc;
    if (n_blooms == 0) return 0;
    if (!blooms) return fail(BIGSI_B200_ERR_INVALID, "null bloom filters");
    const uint64_t nbytes = (n_bits + 7) / 8;
    if (bloom_stride < nbytes) return fail(BIGSI_B200_ERR_INVALID, "bloom_stride smaller than a filter");
    if (int rc = build_columns_check(ix, col0, n_blooms, n_bits)) return rc;
    DeviceGuard guard(ix->device);
    // stage the filters in chunks of whole 32-column words (<= ~1 GiB of HBM), transpose chunk by chunk
    const uint64_t dstride = (ix->num_rows + 255) / 256 * 32;
    uint64_t chunk = (1ull << 30) / dstride;
    chunk = chunk < 32 ? 32 : chunk / 32 * 32;
    cudaError_t e;
    uint64_t done = 0;
    while (done < n_blooms) {
        const uint64_t c0 = col0 + done;
        uint64_t nb = n_blooms - done;
        const uint64_t to_word_end = chunk - (c0 & 31);  // later chunks start on a word boundary
        if (nb > to_word_end) nb = to_word_end;
        if ((e = ix->d_bloom.reserve(nb * dstride)) != cudaSuccess) return fail_cuda(e, "bloom staging");
        if (dstride > nbytes) CK(cudaMemsetAsync(ix->d_bloom.p, 0, nb * dstride, ix->stream));
        if (nbytes)
            CK(cudaMemcpy2DAsync(ix->d_bloom.p, dstride, blooms + done * bloom_stride, bloom_stride, nbytes, nb,
                                 cudaMemcpyHostToDevice, ix->stream));
        CK(launch_transpose_blooms(ix->matrix, ix->pitch, ix->num_rows, c0, nb, static_cast<const uint8_t *>(ix->d_bloom.p),
                                   dstride, n_bits, ix->stream));
        ix->kernel_launches++;
        CK(cudaStreamSynchronize(ix->stream));
        done += nb;
    }
    if (col0 + n_blooms > ix->num_cols) ix->num_cols = col0 + n_blooms;
    return 0;
}

// ============================================================================================
// score=True support: per-window presence of the hit columns
// ============================================================================================
int bigsi_b200_sequence_presence(bigsi_b200_index *ix, const char *seq, uint64_t len, int k, int h, const int32_t *cols,
                                 uint64_t n_cols, uint8_t *out)
{
    if (int rc = check_index(ix)) return rc;
    if (k < 1 || h < 1) return fail(BIGSI_B200_ERR_INVALID, "k and h must be >= 1");
    if (len < (uint64_t)k || n_cols == 0) return 0;
    if (!seq || !cols || !out) return fail(BIGSI_B200_ERR_INVALID, "null pointer");
    for (uint64_t i = 0; i < n_cols; ++i)
        if (cols[i] < 0 || (uint64_t)cols[i] >= ix->col_capacity)
            return fail(BIGSI_B200_ERR_RANGE, "column %d outside the matrix", cols[i]);
    const uint64_t n = len - (uint64_t)k + 1;
    DeviceGuard guard(ix->device);
    cudaError_t e;
    if ((e = ix->d_seq.reserve(len + 64)) != cudaSuccess) return fail_cuda(e, "sequence staging");
    if ((e = ix->d_rows.reserve(n * (uint64_t)h * 4)) != cudaSuccess) return fail_cuda(e, "row-id staging");
    if ((e = ix->d_min.reserve(n_cols * 4)) != cudaSuccess) return fail_cuda(e, "column staging");
    if ((e = ix->d_out.reserve(n * n_cols)) != cudaSuccess) return fail_cuda(e, "presence staging");
    CK(cudaMemcpyAsync(ix->d_seq.p, seq, len, cudaMemcpyHostToDevice, ix->stream));
    CK(cudaMemcpyAsync(ix->d_min.p, cols, n_cols * 4, cudaMemcpyHostToDevice, ix->stream));
    CK(launch_hash_windows(static_cast<const uint8_t *>(ix->d_seq.p), n, k, h, ix->num_rows, static_cast<int32_t *>(ix->d_rows.p),
                           ix->stream));
    CK(launch_presence(ix->matrix, ix->pitch, static_cast<const int32_t *>(ix->d_rows.p), n, h,
                       static_cast<const int32_t *>(ix->d_min.p), n_cols, static_cast<uint8_t *>(ix->d_out.p), ix->stream));
    ix->kernel_launches += 2;
    CK(cudaMemcpyAsync(out, ix->d_out.p, n * n_cols, cudaMemcpyDeviceToHost, ix->stream));
    CK(cudaStreamSynchronize(ix->stream));
    return 0;
}

// ============================================================================================
// persistence: flat index file <-> HBM through two pinned buffers
// ============================================================================================
namespace {
constexpr char kFileMagic[8] = {'B', 'I', 'G', 'S', 'I', 'B', '2', '\n'};
constexpr uint64_t kFileAlign = 4096;
constexpr uint64_t kIoChunkBytes = 32ull << 20;

struct FileCloser {
    FILE *f;
    ~FileCloser()
    {
        if (f) fclose(f);
    }
};
}  // namespace

int bigsi_b200_index_save(bigsi_b200_index *ix, const char *path, const void *meta, uint64_t meta_bytes)
{
    if (int rc = check_index(ix)) return rc;
    if (!path || (meta_bytes && !meta)) return fail(BIGSI_B200_ERR_INVALID, "null pointer");
    const uint64_t row_bytes = (ix->num_cols + 7) / 8;
    bigsi_b200_file_header hd;
    memset(&hd, 0, sizeof hd);
    memcpy(hd.magic, kFileMagic, 8);
    hd.version = 1;
    hd.header_bytes = (uint32_t)sizeof hd;
    hd.num_rows = ix->num_rows;
    hd.num_cols = ix->num_cols;
    hd.col_offset = ix->col_offset;
    hd.row_bytes = row_bytes;
    hd.meta_bytes = meta_bytes;
    hd.rows_offset = round_up(sizeof hd + meta_bytes, kFileAlign);
    FileCloser fc{fopen(path, "wb")};
    if (!fc.f) return fail(BIGSI_B200_ERR_INVALID, "cannot open %s for writing: %s", path, strerror(errno));
    if (fwrite(&hd, sizeof hd, 1, fc.f) != 1 || (meta_bytes && fwrite(meta, 1, meta_bytes, fc.f) != meta_bytes))
        return fail(BIGSI_B200_ERR_INVALID, "write to %s failed: %s", path, strerror(errno));
    std::vector<uint8_t> pad(hd.rows_offset - sizeof hd - meta_bytes, 0);
    if (!pad.empty() && fwrite(pad.data(), 1, pad.size(), fc.f) != pad.size())
        return fail(BIGSI_B200_ERR_INVALID, "write to %s failed: %s", path, strerror(errno));
    if (row_bytes == 0) return 0;
    DeviceGuard guard(ix->device);
    // device -> pinned buffer A while buffer B is being written to the file
    const uint64_t rows_per_chunk = kIoChunkBytes / row_bytes ? kIoChunkBytes / row_bytes : 1;
    PinnedBuf buf[2];
    cudaEvent_t ev[2] = {nullptr, nullptr};
    int rc = 0;
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = buf[i].reserve(rows_per_chunk * row_bytes);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) rc = fail_cuda(e, "pinned staging");
    const uint64_t n_chunks = (ix->num_rows + rows_per_chunk - 1) / rows_per_chunk;
    auto issue = [&](uint64_t c) {
        const uint64_t r0 = c * rows_per_chunk, nr = std::min(rows_per_chunk, ix->num_rows - r0);
        cudaError_t ee = cudaMemcpy2DAsync(buf[c & 1].p, row_bytes, ix->matrix + r0 * ix->pitch, ix->pitch, row_bytes, nr,
                                           cudaMemcpyDeviceToHost, ix->stream);
        if (ee == cudaSuccess) ee = cudaEventRecord(ev[c & 1], ix->stream);
        return ee;
    };
    if (!rc && (e = issue(0)) != cudaSuccess) rc = fail_cuda(e, "device-to-host copy");
    for (uint64_t c = 0; c < n_chunks && !rc; ++c) {
        if (c + 1 < n_chunks && (e = issue(c + 1)) != cudaSuccess) {
            rc = fail_cuda(e, "device-to-host copy");
            break;
        }
        if ((e = cudaEventSynchronize(ev[c & 1])) != cudaSuccess) {
            rc = fail_cuda(e, "device-to-host copy");
            break;
        }
        const uint64_t r0 = c * rows_per_chunk, nr = std::min(rows_per_chunk, ix->num_rows - r0);
        if (fwrite(buf[c & 1].p, 1, nr * row_bytes, fc.f) != nr * row_bytes)
            rc = fail(BIGSI_B200_ERR_INVALID, "write to %s failed: %s", path, strerror(errno));
    }
    cudaStreamSynchronize(ix->stream);
    for (int i = 0; i < 2; ++i) {
        buf[i].release();
        if (ev[i]) cudaEventDestroy(ev[i]);
    }
    if (!rc && fflush(fc.f) != 0) rc = fail(BIGSI_B200_ERR_INVALID, "flush of %s failed: %s", path, strerror(errno));
    return rc;
}

int bigsi_b200_file_info(const char *path, bigsi_b200_file_header *header_out, void *meta_out, uint64_t meta_cap)
{
    if (!path || !header_out) return fail(BIGSI_B200_ERR_INVALID, "null pointer");
    FileCloser fc{fopen(path, "rb")};
    if (!fc.f) return fail(BIGSI_B200_ERR_INVALID, "cannot open %s: %s", path, strerror(errno));
    bigsi_b200_file_header hd;
    if (fread(&hd, sizeof hd, 1, fc.f) != 1 || memcmp(hd.magic, kFileMagic, 8) != 0)
        return fail(BIGSI_B200_ERR_INVALID, "%s is not a bigsi_b200 index file", path);
    if (hd.version != 1 || hd.header_bytes != sizeof hd || hd.row_bytes != (hd.num_cols + 7) / 8 ||
        hd.rows_offset < sizeof hd + hd.meta_bytes)
        return fail(BIGSI_B200_ERR_INVALID, "%s: unsupported or corrupt header (version %u)", path, hd.version);
    *header_out = hd;
    if (meta_out && meta_cap) {
        const uint64_t n = hd.meta_bytes < meta_cap ? hd.meta_bytes : meta_cap;
        if (n && fread(meta_out, 1, n, fc.f) != n) return fail(BIGSI_B200_ERR_INVALID, "%s: truncated metadata", path);
    }
    return 0;
}

int bigsi_b200_index_load_rows(bigsi_b200_index *ix, const char *path, uint64_t file_offset, uint64_t file_stride,
                               uint64_t src_byte_offset, uint64_t row0, uint64_t n_rows)
{
    if (int rc = check_index(ix)) return rc;
    if (!path) return fail(BIGSI_B200_ERR_INVALID, "null path");
    if (n_rows == 0) return 0;
    if (row0 + n_rows > ix->num_rows) return fail(BIGSI_B200_ERR_RANGE, "rows out of range");
    const uint64_t row_bytes = (ix->num_cols + 7) / 8;
    if (row_bytes == 0) return 0;
    if (file_stride < src_byte_offset + row_bytes) return fail(BIGSI_B200_ERR_INVALID, "file rows narrower than this shard");
    FileCloser fc{fopen(path, "rb")};
    if (!fc.f) return fail(BIGSI_B200_ERR_INVALID, "cannot open %s: %s", path, strerror(errno));
    if (fseeko(fc.f, (off_t)file_offset, SEEK_SET) != 0) return fail(BIGSI_B200_ERR_INVALID, "seek in %s failed", path);
    DeviceGuard guard(ix->device);
    // file -> pinned buffer B while buffer A travels to the device (cudaMemcpy2DAsync from pinned memory)
    const uint64_t rows_per_chunk = kIoChunkBytes / file_stride ? kIoChunkBytes / file_stride : 1;
    PinnedBuf buf[2];
    cudaEvent_t ev[2] = {nullptr, nullptr};
    int rc = 0;
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = buf[i].reserve(rows_per_chunk * file_stride);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) rc = fail_cuda(e, "pinned staging");
    const uint8_t last_mask = (ix->num_cols & 7) ? (uint8_t)(0xff00u >> (ix->num_cols & 7)) : (uint8_t)0xff;
    const uint64_t n_chunks = (n_rows + rows_per_chunk - 1) / rows_per_chunk;
    for (uint64_t c = 0; c < n_chunks && !rc; ++c) {
        const uint64_t r0 = c * rows_per_chunk, nr = std::min(rows_per_chunk, n_rows - r0);
        uint8_t *b = static_cast<uint8_t *>(buf[c & 1].p);
        if (c >= 2 && (e = cudaEventSynchronize(ev[c & 1])) != cudaSuccess) {  // the copy that used this buffer is done
            rc = fail_cuda(e, "host-to-device copy");
            break;
        }
        if (fread(b, 1, nr * file_stride, fc.f) != nr * file_stride) {
            rc = fail(BIGSI_B200_ERR_INVALID, "%s: truncated row data", path);
            break;
        }
        if (last_mask != 0xff)  // columns >= num_cols of a wider file row must not leak into the padding
            for (uint64_t i = 0; i < nr; ++i) b[i * file_stride + src_byte_offset + row_bytes - 1] &= last_mask;
        e = cudaMemcpy2DAsync(ix->matrix + (row0 + r0) * ix->pitch, ix->pitch, b + src_byte_offset, file_stride, row_bytes, nr,
                              cudaMemcpyHostToDevice, ix->stream);
        if (e == cudaSuccess) e = cudaEventRecord(ev[c & 1], ix->stream);
        if (e != cudaSuccess) rc = fail_cuda(e, "host-to-device copy");
    }
    e = cudaStreamSynchronize(ix->stream);
    if (!rc && e != cudaSuccess) rc = fail_cuda(e, "host-to-device copy");
    for (int i = 0; i < 2; ++i) {
        buf[i].release();
        if (ev[i]) cudaEventDestroy(ev[i]);
    }
    return rc;
}

// ============================================================================================
// column-sharded exchange (multi-GPU, one handle per GPU)
// ============================================================================================
int bigsi_b200_exchange_create(bigsi_b200_index *ix, int world, int rank, uint64_t max_kmer_bytes, uint32_t spec,
                               uint8_t *ipc_handle_out)
{
    if (int rc = check_index(ix)) return rc;
    if (world < 1 || world > kMaxSinks - 1 || rank < 0 || rank >= world)
        return fail(BIGSI_B200_ERR_INVALID, "bad world/rank %d/%d (at most %d shards)", world, rank, kMaxSinks - 1);
    if (spec == 0 || max_kmer_bytes == 0) return fail(BIGSI_B200_ERR_INVALID, "spec and max_kmer_bytes must be positive");
    if (ix->ex.local) return fail(BIGSI_B200_ERR_INVALID, "exchange already created");
    DeviceGuard guard(ix->device);
    Exchange &ex = ix->ex;
    ex.world = world;
    ex.rank = rank;
    ex.spec = spec;
    ex.max_kmer_bytes = max_kmer_bytes;
    ex.kmers_stride = round_up(max_kmer_bytes + 64, 256);
    ex.ll_off = 0;  // LL inboxes: every data byte takes two (hash.cuh:ll_store_line)
    ex.sinks_off = ex.ll_off + kExInboxes * 2 * ex.kmers_stride;
    ex.block_bytes = round_up(16 + 8ull * spec + 8, 128);
    ex.total_bytes = ex.sinks_off + kExGenerations * world * ex.block_bytes;
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&ex.local), ex.total_bytes);
    if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc(exchange)");
    CK(cudaMemset(ex.local, 0, ex.total_bytes));
    ex.peer[rank] = ex.local;
    if (ipc_handle_out) {
        cudaIpcMemHandle_t hdl;
        CK(cudaIpcGetMemHandle(&hdl, ex.local));
        static_assert(sizeof(hdl) == 64, "cudaIpcMemHandle_t is 64 bytes");
        memcpy(ipc_handle_out, &hdl, 64);
    }
    ex.ready = world == 1;
    return 0;
}

int bigsi_b200_exchange_open(bigsi_b200_index *ix, const uint8_t *ipc_handles)
{
    if (int rc = check_index(ix)) return rc;
    Exchange &ex = ix->ex;
    if (!ex.local) return fail(BIGSI_B200_ERR_INVALID, "exchange not created");
    if (!ipc_handles) return fail(BIGSI_B200_ERR_INVALID, "null handles");
    DeviceGuard guard(ix->device);
    for (int r = 0; r < ex.world; ++r) {
        if (r == ex.rank || ex.peer[r]) continue;
        cudaIpcMemHandle_t hdl;
        memcpy(&hdl, ipc_handles + 64 * r, 64);
        void *ptr = nullptr;
        CK(cudaIpcOpenMemHandle(&ptr, hdl, cudaIpcMemLazyEnablePeerAccess));
        ex.peer[r] = static_cast<uint8_t *>(ptr);
        ex.ipc_opened[r] = true;
    }
    ex.ready = true;
    return 0;
}

int bigsi_b200_exchange_open_local(bigsi_b200_index *ix, bigsi_b200_index *const *peers)
{
    if (int rc = check_index(ix)) return rc;
    Exchange &ex = ix->ex;
    if (!ex.local) return fail(BIGSI_B200_ERR_INVALID, "exchange not created");
    if (!peers) return fail(BIGSI_B200_ERR_INVALID, "null peers");
    DeviceGuard guard(ix->device);
    for (int r = 0; r < ex.world; ++r) {
        if (r == ex.rank) continue;
        const bigsi_b200_index *pr = peers[r];
        if (!pr || !pr->ex.local || pr->ex.world != ex.world || pr->ex.rank != r || pr->ex.total_bytes != ex.total_bytes)
            return fail(BIGSI_B200_ERR_INVALID, "peer %d has no matching exchange block", r);
        if (pr->device != ix->device) {
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, ix->device, pr->device));
            if (!can) return fail(BIGSI_B200_ERR_CUDA, "device %d cannot access device %d", ix->device, pr->device);
            cudaError_t e = cudaDeviceEnablePeerAccess(pr->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail_cuda(e, "cudaDeviceEnablePeerAccess");
            (void)cudaGetLastError();
        }
        ex.peer[r] = pr->ex.local;
    }
    ex.ready = true;
    return 0;
}

int bigsi_b200_exchange_destroy(bigsi_b200_index *ix)
{
    if (!ix) return 0;
    Exchange &ex = ix->ex;
    if (!ex.local) return 0;
    DeviceGuard guard(ix->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < ex.world; ++r)
        if (ex.ipc_opened[r] && ex.peer[r]) cudaIpcCloseMemHandle(ex.peer[r]);
    cudaFree(ex.local);
    ex.h_gather.release();
    ex = Exchange();
    return 0;
}

// slot `rank` of generation seq % kExGenerations in shard r's result blocks: where this shard's hits of query `seq` go
static unsigned long long *exchange_slot(const Exchange &ex, int r, uint64_t seq)
{
    return reinterpret_cast<unsigned long long *>(ex.peer[r] + ex.sinks_off +
                                                  ((seq % kExGenerations) * ex.world + ex.rank) * ex.block_bytes);
}
static const uint8_t *exchange_blocks(const Exchange &ex, uint64_t seq)
{
    return ex.local + ex.sinks_off + (seq % kExGenerations) * ex.world * ex.block_bytes;
}

// One query of a column-sharded search: a streamed launch on every rank.  Rank 0's gather kernel pushes the k-mer
// bytes into the peers' LL inboxes, the peers' gather kernels hash out of theirs; every rank's reduce kernel
// publishes its hits into slot `rank` of every rank's result blocks and finishes when all slots of its own copy
// carry this query's number.
static int exchange_search(bigsi_b200_index *ix, const char *d_kmers, uint64_t n_kmers, int k, int h, uint32_t min_kmers,
                           cudaStream_t stream)
{
    Exchange &ex = ix->ex;
    if (!ex.local || !ex.ready) return fail(BIGSI_B200_ERR_INVALID, "exchange not created / peers not opened");
    if (k < 1 || h < 1 || n_kmers == 0) return fail(BIGSI_B200_ERR_INVALID, "k, h and the number of k-mers must be positive");
    if (n_kmers * (uint64_t)k > ex.max_kmer_bytes) return fail(BIGSI_B200_ERR_RANGE, "query larger than the exchange inbox");
    if (ex.rank == 0 && (!d_kmers || (reinterpret_cast<uintptr_t>(d_kmers) & 15)))
        return fail(BIGSI_B200_ERR_INVALID, "rank 0 needs the k-mers in a 16-byte aligned device-addressable buffer");
    if (ix->num_cols == 0) return fail(BIGSI_B200_ERR_INVALID, "empty shard");
    cudaError_t e;
    const uint64_t hit_slot = round_up(8 + 8ull * ex.spec + 16, 256);
    if (hit_slot * kStreamStates > ix->d_hits_ring.cap) {
        if (int rc = flush_pending(ix)) return rc;
        CK(cudaStreamSynchronize(stream));
        if ((e = ix->d_hits_ring.reserve(hit_slot * kStreamStates)) != cudaSuccess) return fail_cuda(e, "staging");
    }
    // the exchange counts its own queries (the same number on every rank: it picks the inbox, the result generation
    // and the LL flag); ring slots of the handle follow the handle's streamed launch number, which advances at least
    // as fast, so the entry gate of the streamed path also covers the reuse of inboxes and generations
    const uint64_t seq = ex.seq + 1;
    const uint64_t inbox = seq % kExInboxes;
    uint8_t *dev = static_cast<uint8_t *>(ix->d_hits_ring.p) + ((ix->stream_seq + 1) % kStreamStates) * hit_slot;
    HitsOut ho;
    ho.n = reinterpret_cast<unsigned long long *>(dev);
    ho.cols = reinterpret_cast<int32_t *>(dev + 8);
    ho.counts = reinterpret_cast<uint32_t *>(dev + 8 + 4ull * ex.spec);
    ho.cap = ex.spec;
    ho.by_value = true;
    ho.min_value = min_kmers;
    ho.require_stream = true;
    ho.deferred = true;  // published by the next search's merge team, or by bigsi_b200_index_flush
    ho.inputs_ready = ix->opt_inputs_ready != 0;
    ho.sink_spec = ex.spec;
    ho.sink_seq = seq;
    ho.ll.flag = (uint32_t)seq ? (uint32_t)seq : 0x80000000u;  // an LL inbox is reused every kExInboxes queries: never 0, never the old value
    auto ll_inbox = [&](int r) { return reinterpret_cast<uint4 *>(ex.peer[r] + ex.ll_off + inbox * 2 * ex.kmers_stride); };
    ho.n_sinks = (uint32_t)ex.world;
    ho.gather_seq = seq;
    ho.n_gather = (uint32_t)ex.world;
    const uint8_t *blocks = exchange_blocks(ex, seq);
    for (int r = 0; r < ex.world; ++r) {
        ho.sinks[r] = exchange_slot(ex, r, seq);
        ho.gather_blocks[r] = reinterpret_cast<const unsigned long long *>(blocks + r * ex.block_bytes);
    }
    if (ex.h_gather.p && ex.h_gather_on) {
        void *dp = nullptr;
        CK(cudaHostGetDevicePointer(&dp, static_cast<uint8_t *>(ex.h_gather.p) + (seq % kStreamStates) * ex.h_gather_block, 0));
        ho.host_gather = static_cast<unsigned long long *>(dp);
        ho.host_block_words = (uint32_t)(ex.block_bytes / 8);
        ex.h_gather_seq[seq % kStreamStates] = seq;
    }
    const char *kmers = d_kmers;
    if (ex.rank == 0) {
        for (int r = 1; r < ex.world; ++r) ho.ll.out[ho.n_push++] = ll_inbox(r);
    } else {
        kmers = nullptr;  // hashed out of the inbox; the address only fixes the line numbering (see below)
        ho.ll.in = ll_inbox(ex.rank);
    }
    // a peer's "k-mer array" starts at line 0 of its inbox: any 16-byte aligned non-null base does (never dereferenced)
    if (!kmers) kmers = reinterpret_cast<const char *>(ex.local);
    bool published = false;
    ho.published = &published;
    if (int rc = run_query(ix, BIGSI_B200_MODE_COUNTS, nullptr, kmers, k, nullptr, 1, n_kmers, n_kmers, h, nullptr, 0, stream, &ho))
        return rc;
    if (!published) return fail(BIGSI_B200_ERR_INVALID, "the launch could not publish its result");
    ex.seq = seq;
    return 0;
}

int bigsi_b200_exchange_reserve(bigsi_b200_index *ix, uint64_t max_kmers, int k, int h)
{
    if (int rc = check_index(ix)) return rc;
    Exchange &ex = ix->ex;
    if (!ex.local) return fail(BIGSI_B200_ERR_INVALID, "exchange not created");
    if (k < 1 || h < 1 || max_kmers == 0) return fail(BIGSI_B200_ERR_INVALID, "k, h and max_kmers must be positive");
    if (ix->num_cols == 0) return 0;
    DeviceGuard guard(ix->device);
    if (int rc = ensure_kernels()) return rc;
    // scratch of a streamed query grows with the number of k-mers per CTA (partial planes) and with the grid (pool,
    // none for back-to-back queries): the largest query and the first one that fills the grid cover both
    for (uint64_t n : {max_kmers, std::min<uint64_t>(max_kmers, (uint64_t)ix->sm_count)}) {
        QueryParams p;
        int grid = 0;
        if (int rc = plan_query(ix, BIGSI_B200_MODE_COUNTS, 1, n, n, h, true, k, p, grid)) return rc;
        if (!p.stream) return fail(BIGSI_B200_ERR_INVALID, "a query of %llu k-mers cannot run as a streamed launch", (unsigned long long)n);
        if (int rc = reserve_stream_scratch(ix, p, grid, ix->stream)) return rc;
    }
    const uint64_t hit_slot = round_up(8 + 8ull * ex.spec + 16, 256);
    if (hit_slot * kStreamStates > ix->d_hits_ring.cap) {
        cudaError_t e = ix->d_hits_ring.reserve(hit_slot * kStreamStates);
        if (e != cudaSuccess) return fail_cuda(e, "staging");
    }
    CK(cudaStreamSynchronize(ix->stream));
    return 0;
}

int bigsi_b200_exchange_search_dev(bigsi_b200_index *ix, const char *d_kmers, uint64_t n_kmers, int k, int h,
                                   uint32_t min_kmers, void *stream, const void **d_blocks_out, uint64_t *block_bytes_out)
{
    if (int rc = check_index(ix)) return rc;
    DeviceGuard guard(ix->device);
    if (int rc = exchange_search(ix, d_kmers, n_kmers, k, h, min_kmers, static_cast<cudaStream_t>(stream))) return rc;
    if (d_blocks_out) *d_blocks_out = exchange_blocks(ix->ex, ix->ex.seq);
    if (block_bytes_out) *block_bytes_out = ix->ex.block_bytes;
    return 0;
}

int bigsi_b200_exchange_last_seq(bigsi_b200_index *ix, uint64_t *seq_out)
{
    if (int rc = check_index(ix)) return rc;
    if (!seq_out) return fail(BIGSI_B200_ERR_INVALID, "null pointer");
    *seq_out = ix->ex.seq;
    return 0;
}

int bigsi_b200_exchange_host_results(bigsi_b200_index *ix, int enable)
{
    if (int rc = check_index(ix)) return rc;
    Exchange &ex = ix->ex;
    if (!ex.local) return fail(BIGSI_B200_ERR_INVALID, "exchange not created");
    DeviceGuard guard(ix->device);
    if (enable && !ex.h_gather.p) {
        ex.h_gather_block = round_up(16 + (uint64_t)ex.world * ex.block_bytes, 128);
        cudaError_t e = ex.h_gather.reserve(ex.h_gather_block * kStreamStates);
        if (e != cudaSuccess) return fail_cuda(e, "pinned result blocks");
    }
    ex.h_gather_on = enable != 0;
    return 0;
}

int bigsi_b200_exchange_wait_host(bigsi_b200_index *ix, uint64_t seq, const void **blocks_out, uint64_t *block_bytes_out)
{
    if (int rc = check_index(ix)) return rc;
    Exchange &ex = ix->ex;
    if (!ex.h_gather.p) return fail(BIGSI_B200_ERR_INVALID, "host results are not enabled on this exchange");
    if (seq == 0 || seq > ex.seq || seq + kStreamStates <= ex.seq)
        return fail(BIGSI_B200_ERR_INVALID, "query %llu is not among the last %d searches (last: %llu)", (unsigned long long)seq,
                    kStreamStates, (unsigned long long)ex.seq);
    if (ex.h_gather_seq[seq % kStreamStates] != seq)
        return fail(BIGSI_B200_ERR_INVALID, "search %llu was launched with host results off", (unsigned long long)seq);
    DeviceGuard guard(ix->device);
    // the newest search has nobody behind it to run its stage 2: flush it (SPMD: every rank does, or searches on)
    if (ix->pending.have && ix->pending.p.gather_seq == seq && ix->pending.p.host_gather)
        if (int rc = flush_pending(ix)) return rc;
    volatile unsigned long long *blk = reinterpret_cast<volatile unsigned long long *>(static_cast<uint8_t *>(ex.h_gather.p) +
                                                                                    (seq % kStreamStates) * ex.h_gather_block);
    uint64_t spins = 0;
    while (__atomic_load_n(&blk[0], __ATOMIC_ACQUIRE) != seq) {
        if ((++spins & 0x3fff) == 0) {
            // (every device-side wait is bounded: a search that cannot complete raises the abort word)
            if (const unsigned long long av = abort_state(ix)) return fail_aborted(av);
        }
    }
    if (blocks_out) *blocks_out = const_cast<unsigned long long *>(blk) + 2;
    if (block_bytes_out) *block_bytes_out = ex.block_bytes;
    return 0;
}

int bigsi_b200_exchange_wait_ns(bigsi_b200_index *ix, uint64_t *wait_ns_out, uint64_t *queries_out)
{
    if (int rc = check_index(ix)) return rc;
    if (!wait_ns_out) return fail(BIGSI_B200_ERR_INVALID, "null pointer");
    DeviceGuard guard(ix->device);
    CK(cudaDeviceSynchronize());
    unsigned long long v = 0;
    CK(cudaMemcpy(&v, static_cast<uint8_t *>(ix->d_stream.p) + 128, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemset(static_cast<uint8_t *>(ix->d_stream.p) + 128, 0, 8));
    *wait_ns_out = v;
    if (queries_out) *queries_out = ix->stream_seq;
    return 0;
}

int bigsi_b200_index_status(bigsi_b200_index *ix)
{
    if (int rc = check_index(ix)) return rc;
    if (const unsigned long long av = abort_state(ix)) return fail_aborted(av);
    return 0;
}

}  // extern "C"
