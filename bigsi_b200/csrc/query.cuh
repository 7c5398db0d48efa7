// Launch interface of the fused gather-AND-{count|AND} kernel and its merge kernel
// (query_kernels.cu) and of the small helper kernels (aux_kernels.cu).
// See DESIGN.md "Kernels" for the work decomposition.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "hash.cuh"

namespace bigsi {

constexpr int kMaxConsumerWarps = 13;                           // 416 consumer threads, one 16-byte unit each
constexpr int kMaxBlockThreads = (kMaxConsumerWarps + 1) * 32;  // + one TMA producer warp
constexpr int kMaxTileBytes = kMaxConsumerWarps * 32 * 16;      // 6656 B = 53 248 sample columns per tile
constexpr int kMaxStages = 32;
constexpr int kSmemHeaderBytes = 1024;                          // mbarriers
constexpr int kSmemBudget = 227 * 1024;
constexpr uint32_t kMaxSliceItems = 65535;                      // a segment's count fits 16 bit planes
constexpr int kSegPlanes = 16;
constexpr int kMaxH = 1024;
constexpr int kPoolMaxH = 8;   // the solo path pools k-mers only for h <= 8 (row-id stash in the smem header)
constexpr int kMaxSinks = 9;   // host + up to 8 GPUs of one box
// "streamed" single-query launches: the gather kernel of query s+1 overlaps the reduce
// kernel of query s, so everything a query owns rotates: big buffers (partial planes, pooled row ids, the
// front-end's table) over kStreamRing queries, the small state blocks over kStreamStates
constexpr int kStreamRing = 4;
constexpr int kStreamStates = 8;
// Stage 2 (merge + threshold + publication) of a streamed query s normally runs INSIDE the gather kernel of query
// s + 1: kMergeTeamWarps extra warps per gather CTA (the "merge team") wait for the preceding grid to complete and
// then share the merge items of query s, while the other warps of the same CTA stream the rows of query s + 1.  One
// CTA per SM as before -- nothing has to become co-resident with anything.  The last query of a burst (and every
// synchronous call) is flushed by reduce_kernel (merge_kernels.cu), kReduceThreads threads per CTA.
constexpr int kMergeTeamWarps = 4;
constexpr int kMergeTeamThreads = kMergeTeamWarps * 32;
constexpr int kReduceThreads = 256;
// merge scratch of a streamed query (team and flush kernel alike): large enough to stage all 148 slots of a
// 10 000-k-mer query's merge item (7 planes x 48 B each) in ONE batch -- staging is latency bound (one bulk-copy
// round trip per batch), and with five batches the team was still merging when its CTA's gather ended.  The ring
// keeps 8 stages beside it, which is as fast as 11 (measured).
constexpr int kReduceSmemBytes = 56 * 1024;
constexpr int kReduceFatSmemBytes = 64 * 1024;   // isolated queries: the flush is on the critical path, bigger batches

enum { kModeCounts = 0, kModeAnd = 1 };

// Per-query state block of a streamed launch (64 bytes).  All zero when its query starts: stage 2 of
// query s clears the block of query s + kStreamRing before it advances the completion word.
struct QState {
    unsigned int pool_claims;       // claim counter of the k-mer pool (gather kernel)
    unsigned int reduce_arrivals;   // reduce-kernel CTAs that have finished their merge items
    unsigned long long n_unique;    // sequence front-end: unique windows found by the gather kernel's CTAs
    unsigned long long wait_ns;     // diagnostics: time the reduce kernel's last CTA waited for the other shards
    unsigned int gather_arrivals;   // self-merging launches: gather CTAs of this query that have flushed their planes
    unsigned int pad0;
    unsigned long long pad[4];
};
static_assert(sizeof(QState) == 64, "QState is one 64-byte block");
// error codes a kernel leaves in the abort word when a bounded wait times out (sticky; see bounded_wait)
enum { kAbortGate = 1, kAbortPool = 2, kAbortInbox = 3, kAbortPeers = 4, kAbortChain = 5, kAbortExit = 6 };

// Work space of one launch: items = (column tile, global k-mer index), k-mer fastest.  It is cut
// into n_slices equal slices of items_per_slice (<= 65535); CTA b owns the contiguous slices
// [b*slices_per_cta, (b+1)*slices_per_cta); a slice is cut into SEGMENTS at (tile, query)
// boundaries.  A segment accumulates in registers and is written once, as bit planes, to partial
// slot  slice + tile*n_queries + query  (unique per segment, monotone in item order).
struct QueryParams {
    const uint8_t *matrix;    // m rows x pitch bytes, MSB-first sample columns
    uint64_t pitch;           // multiple of 128
    const int32_t *rows;      // [total_kmers * h] row ids
    const int64_t *qoff;      // [n_queries + 1] k-mer offsets
    uint32_t n_queries;
    uint32_t h;
    uint64_t total_kmers;
    uint32_t num_cols;        // valid columns of this shard
    uint32_t row_bytes16;     // ceil(num_cols/8) rounded up to 16 (padding bits are zero in HBM)
    uint32_t tile_bytes;      // column-tile width in bytes, multiple of 16, <= kMaxTileBytes
    uint32_t n_tiles;         // ceil(row_bytes16 / tile_bytes)
    uint32_t kmers_per_stage; // G: k-mers staged per ring slot
    uint32_t n_stages;        // ring depth
    uint64_t total_items;     // n_tiles * total_kmers
    uint32_t items_per_slice; // <= kMaxSliceItems
    uint32_t n_slices;
    uint32_t slices_per_cta;
    uint8_t *partial;         // [merge_cpt chunks][n_slots_total][planes_per_slot][merge_cb bytes], see partial_offset
    uint32_t planes_per_slot; // COUNTS: bits(min(items_per_slice, longest query)); AND: 1
    uint32_t total_planes;    // bits(longest query) <= 32 (merge accumulator width)
    uint64_t max_query_kmers; // upper bound on the longest query (sizes the merge's slot loop)
    void *out;                // COUNTS: uint32 [n_queries][out_stride]; AND: uint8 [n_queries][out_stride]; may be
                              // null in COUNTS mode when only the thresholded hits are wanted
    uint64_t out_stride;      // elements (COUNTS) or bytes (AND)
    // optional fused threshold (COUNTS mode): keep columns with count >= min_kmers[q]
    const uint32_t *min_kmers;       // [n_queries] or null
    int32_t *hit_cols;               // [n_queries][hit_cap] LOCAL column ids, order unspecified
    uint32_t *hit_counts;            // [n_queries][hit_cap]
    unsigned long long *n_hits;      // [n_queries] number of hits (may exceed hit_cap); zeroed by stage 1
    uint64_t hit_cap;
    // optional in-kernel hashing: the CTA hashes its own (contiguous) k-mers in the prologue instead
    // of reading row ids (needs n_tiles == 1 and slices_per_cta == 1)
    const uint8_t *kmers;     // [total_kmers * k] raw ASCII k-mers, or null (then `rows` is used)
    uint32_t k;
    uint32_t num_rows;        // m
    uint64_t mod_magic;       // hash.cuh:mod_magic(m), the fast-modulo constant of the in-kernel hashing
    uint32_t prehash;         // 1: hash in the prologue into the shared-memory id table
    uint32_t ids_bytes;       // size of that table + the hashing scratch (between the mbarriers and the ring)
    uint32_t ids_table_bytes; // the table alone; the hashing scratch follows it
    // "solo" path (one query, one tile, one slice per CTA, in-kernel hashing and merge): the producer warp
    // hashes the first ring-full of k-mers itself and starts gathering while the consumer warps hash the
    // rest; the last pool_share k-mers of every CTA's range go to a shared POOL that the CTAs drain
    // dynamically (atomic claims), so fast CTAs take over work of slow ones (tail balance)
    uint32_t solo;
    // the query front-end (dedup_windows_kernel) determines the number of unique k-mers on the device: the
    // kernel then reads it here (total_kmers is the upper bound the launch was planned for) and reports it in
    // word 2 + sink_spec of every sink block
    const unsigned long long *total_dev;
    unsigned long long *scrub;   // [scrub_words] cleared behind the grid barrier (the front-end's table), or null; the two
                                 // words behind it ({U}, {ticket, threshold}) are cleared by the last CTA after publishing
    uint64_t scrub_words;
    uint32_t pool_share;      // k-mers per CTA that go to the pool (0 = static split only)
    uint32_t solo_max_kmers;  // most k-mers one CTA may count (2^planes_per_slot - 1)
    int32_t *pool_ids;        // [grid][pool_share][h] row ids of the pooled k-mers, written by their owner CTA
    unsigned long long *pool_ready;  // [grid] = pool_epoch once the owner's ids are visible
    unsigned int *pool_counter;      // claim counter, zero at launch (reset behind the grid barrier)
    unsigned long long pool_epoch;
    // optional in-kernel merge: stage 2 runs inside stage 1 after a grid-wide barrier
    uint32_t fuse_merge;
    // merge geometry (merge.cuh:plan_merge); it also fixes the layout of `partial`
    uint32_t merge_cb;        // column chunk of one merge item in bytes (multiple of 16)
    uint32_t merge_cpt;       // chunks per tile
    uint32_t merge_smem;      // shared-memory scratch one merge item may use
    uint64_t n_slots_total;   // partial slots of this launch (n_slices + n_tiles * n_queries)
    uint64_t merge_items;
    unsigned long long *barrier;        // monotonic arrival counter shared by all launches of a handle
    unsigned long long barrier_target;  // value the counter reaches when every CTA of THIS launch arrived
    // single-query extras ------------------------------------------------------------------------
    uint32_t min_by_value;    // 1: threshold = min_kmers_value (no device array to read)
    uint32_t min_kmers_value;
    // query broadcast of a column-sharded search (streamed launches only): rank 0's producer warp stores its CTA's
    // slice of the k-mer bytes into every peer's LL inbox ("low-latency" lines, hash.cuh:ll_store_line: no fence,
    // no flag); the peers' hashing reads its k-mer bytes from its own LL inbox and spins per 16-byte line
    uint32_t n_push;          // rank 0: number of peers (ll.out[0 .. n_push))
    uint32_t push_repeat;     // diagnostics (option "push_repeat"): every line is stored this many times (emulates more peers)
    uint32_t push_all_warps;  // diagnostics (option "push_all_warps"): 1 = the team's warp 0 pushes too (see gather_solo)
    LlRoute ll;               // hash.cuh: who sends the k-mer bytes to whom
    // result publication: after the merge the LAST CTA (of the reduce kernel, or of the generic kernel's merge
    // phase) copies the hit list of query 0 to every sink -- a block [0] = sequence flag, [1] = number of hits,
    // then int32 cols[sink_spec], uint32 counts[sink_spec] -- in this GPU's, a peer GPU's (NVLink) or the
    // host's (mapped pinned) memory
    uint32_t n_sinks;
    uint32_t sink_spec;
    unsigned long long sink_seq;
    unsigned long long *sinks[kMaxSinks];
    unsigned long long *done_counter;   // generic kernel: monotonic; the CTA that brings it to done_target is the last
    unsigned long long done_target;
    // all-gather completion: after publishing, the last CTA of the reduce kernel waits (bounded) until these LOCAL
    // blocks (the slots the peers publish into) carry gather_seq too, so that stream order implies "all shards
    // have reported".  The next query's gather kernel is already running by then: the wait costs no bandwidth.
    uint32_t n_gather;
    unsigned long long gather_seq;
    // optional: once all blocks are here, copy them (header + the hits each one holds) into a block of mapped HOST
    // memory, [0] = gather_seq (written last, behind a system fence), [1] unused, then n_gather blocks of
    // host_block_words words in the layout of the device blocks -- the host reads every shard's hits without a copy
    // operation in the stream
    unsigned long long *host_gather;
    uint32_t host_block_words;
    const unsigned long long *gather_blocks[kMaxSinks];
    // streamed launch (solo geometry, fuse_merge == 0): no grid barrier and no wait for the preceding kernel.  The
    // gather kernel flushes its planes and exits; stage 2 (merge team of the next launch / reduce_kernel) merges, thresholds and publishes
    // while the NEXT query's gather kernel already runs on the same SMs.
    uint32_t stream;
    uint32_t merge_team;                 // threads of the gather CTA's merge team (kMergeTeamThreads, or 0: a variant without)
    uint32_t merge_prev;                 // 1: the kernel's second argument describes the previous query, to be merged by the team
    uint32_t self_merge;                 // 1 (isolated query, COOPERATIVE launch: all CTAs co-resident): the team merges THIS
                                         // query once every gather CTA of the grid has flushed (arrival counter) -- no flush kernel
    uint32_t stream_wait_inputs;         // 1: the k-mers may be produced by the preceding kernel of the stream: wait for it
    unsigned long long stream_seq;       // number of this query among the handle's streamed launches (1-based)
    unsigned long long *stream_done;     // device word: every streamed query <= *stream_done is completely reduced
    QState *qstate;                      // this query's state block
    QState *qstate_next;                 // the block of query stream_seq + kStreamRing (cleared by the reduce kernel)
    unsigned long long *abort_word;      // device word, non-zero once a bounded wait has timed out (sticky)
    unsigned long long *host_abort;      // its mirror in mapped host memory
    unsigned long long spin_timeout_ns;  // bound of every device-side wait
    unsigned long long *wait_ns_out;     // diagnostics: the reduce kernel adds the time it waited for the other shards
    double seq_threshold;                // sequence front-end: min_kmers = ceil(n_unique * seq_threshold)
    uint32_t seq_mode;                   // 1: `kmers` is a SEQUENCE of total_kmers + k - 1 bytes; windows are de-duplicated in the kernel
    unsigned long long *seq_table;       // its open-addressing table (seq_table_entries x u64, a power of two)
    uint64_t seq_table_entries;
    uint32_t seq_epoch;                  // 16-bit epoch that marks live table entries (stale ones count as empty)
    // batches (generic path, COUNTS): a segment that covers a WHOLE query (the query lies inside one slice) is finished
    // by the CTA that counted it, straight from its registers -- threshold by a bit-sliced comparison, the few hits
    // extracted individually, optionally the full count vector -- and neither writes partial planes nor takes part in
    // the merge phase (merge_item skips queries with a single slot).  Only queries cut by a slice boundary are merged.
    // The hit counters are then zeroed by the host before the launch (the kernel's CTAs start at different times).
    uint32_t direct_complete;
    uint32_t plain_launch;    // 1: launch WITHOUT the cooperative attribute (option "cooperative" = 0): the grid barrier
                              // then relies on grid <= resident CTA capacity alone; lets PDL start the next grid's CTAs early
    uint32_t debug_flags;     // bit 0: consumers skip the AND/count work (pure-gather ceiling measurement)
    unsigned long long *debug_ts;  // optional [grid][kDebugStamps] timeline stamps (globaltimer ns), see fused_query
};

// timeline stamps per CTA (debug_flags bit 1): 0 entry, 1 producer first issue, 2 first slot landed,
// 3 last slot consumed, 4 after flush, 5 producer last issue, 6 past the grid barrier, 7 merge phase
// done, 8 past the PDL wait, 9 prologue hashing done, 10 merge: partial planes loaded + counted,
// 11 merge: counters in shared memory, 12 merge: expansion done, 13 merge phase re-run done (bit 2)
constexpr int kDebugStamps = 16;
#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long debug_gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define BIGSI_TS(slot)                                                                          \
    do {                                                                                        \
        if (P.debug_ts) P.debug_ts[(size_t)blockIdx.x * kDebugStamps + (slot)] = debug_gtime(); \
    } while (0)
#endif

inline uint32_t query_consumer_warps(const QueryParams &p) { return (p.tile_bytes + 511) / 512; }
inline uint32_t query_block_threads(const QueryParams &p) { return (query_consumer_warps(p) + 1) * 32 + p.merge_team; }
constexpr int kMergeScratchBytes = 16 * 16 * 32 * 4 + 256;  // == kMergeSmemBytes (merge.cuh)
inline uint32_t query_ring_bytes(const QueryParams &p) { return p.n_stages * p.kmers_per_stage * p.h * p.tile_bytes; }
inline uint32_t query_smem_bytes(const QueryParams &p)
{
    uint32_t ring = query_ring_bytes(p);
    if (p.fuse_merge && ring < (uint32_t)kMergeScratchBytes) ring = kMergeScratchBytes;  // the merge phase reuses the ring
    return kSmemHeaderBytes + p.ids_bytes + ring + (p.merge_team ? (uint32_t)kReduceSmemBytes : 0u);  // + the team's scratch
}
inline uint64_t query_n_slots(const QueryParams &p)
{
    return (uint64_t)p.n_slices + (uint64_t)p.n_tiles * p.n_queries;
}
// byte offset of (chunk of the tile, partial slot, plane) in `partial`: chunk-major, so that all slots
// and planes of one merge item are contiguous
__host__ __device__ inline uint64_t partial_offset(const QueryParams &p, uint32_t chunk, uint64_t slot, uint32_t plane)
{
    return (((uint64_t)chunk * p.n_slots_total + slot) * p.planes_per_slot + plane) * p.merge_cb;
}
inline uint64_t query_partial_bytes(const QueryParams &p)
{
    return (uint64_t)p.merge_cpt * p.n_slots_total * p.planes_per_slot * p.merge_cb + 256;
}

// Stage 1: gather + AND + vertical count, per-segment bit planes -> p.partial.  prev (streamed launches with
// p.merge_prev only): the previous streamed query of the handle, merged by this launch's merge team.
cudaError_t launch_query(const QueryParams &p, int mode, int grid, cudaStream_t stream, const QueryParams *prev = nullptr);
// true when the streamed launch of p runs the kernel variant that carries a merge team
inline bool query_has_merge_team(const QueryParams &p, int mode) { return p.stream && mode == kModeCounts && p.planes_per_slot <= 8; }
// shared-memory scratch the in-kernel hashing needs per k-mer (raw bytes + flag + canonical words)
inline uint64_t prehash_bytes_per_kmer(uint32_t k) { return (uint64_t)k + 1 + 4ull * ((((uint64_t)k + 3) >> 2) | 1); }
// Stage 2: sum (or AND) the partial planes of every (query, column), expand to integers -> p.out.
cudaError_t launch_merge(const QueryParams &p, int mode, cudaStream_t stream);
// Stage 2 of a streamed launch (p.stream) as its own kernel (the flush): merge + threshold + publication + completion chain, `grid` CTAs.
cudaError_t launch_reduce(const QueryParams &p, int mode, int grid, cudaStream_t stream);
cudaError_t query_kernels_init();  // opt-in to large dynamic shared memory
// drain of the pipelined exchange (uses n_hits/hit_*/sink_spec, n_pub/pub_*, n_gather/gather_* of p)
cudaError_t launch_exchange_drain(const QueryParams &p, cudaStream_t stream);
cudaError_t merge_kernels_init();

// ---- helper kernels (aux_kernels.cu) -------------------------------------------------------
cudaError_t launch_hash_kmers(const char *d_kmers, uint64_t n, int k, int h, uint64_t m, int canonical,
                              int32_t *d_rows_out, cudaStream_t stream);
cudaError_t launch_lookup(const uint8_t *matrix, uint64_t pitch, uint32_t row_bytes, const int32_t *d_rows,
                          uint64_t n_kmers, int h, uint8_t *d_out, uint64_t out_stride, cudaStream_t stream);
// batch reuse: classes of equal row-id tuples (table: a power of two >= 2 n entries, zeroed; *d_counter zeroed, receives
// the number of classes U'); d_ids_out[t] = class of k-mer t, d_unique_rows[u] = the row ids of class u
cudaError_t launch_dedup_rows(const int32_t *d_rows, uint64_t n, int h, unsigned long long *d_table, uint64_t table_entries,
                              uint32_t *d_rep, uint32_t *d_uid_of, unsigned int *d_counter, int32_t *d_ids_out,
                              int32_t *d_unique_rows, cudaStream_t stream);
cudaError_t launch_threshold(const uint32_t *d_counts, uint64_t counts_stride, uint64_t n_queries,
                             uint64_t num_cols, const uint32_t *d_min_kmers, int32_t *d_cols_out,
                             uint32_t *d_counts_out, uint64_t cap, unsigned long long *d_n_out,
                             cudaStream_t stream);
// query front-end: unique raw k-mers of a sequence (table_entries: a power of two >= 2 * n_windows, zeroed;
// *d_counter zeroed, receives U)
// d_ticket (zeroed) / d_min_out: the last block to finish stores min_kmers = ceil(U * threshold) (IEEE double,
// graph/bigsi.py:179; <= 0 -> 0) to *d_min_out, so that the search kernel can follow without a host round trip
cudaError_t launch_dedup_windows(const uint8_t *d_seq, uint64_t n_windows, int k, unsigned long long *d_table,
                                 uint64_t table_entries, uint8_t *d_out_kmers, unsigned long long *d_counter,
                                 unsigned int *d_ticket, double threshold, uint32_t *d_min_out, cudaStream_t stream);
cudaError_t launch_set_column(uint8_t *matrix, uint64_t pitch, uint64_t num_rows, uint64_t col,
                              const uint8_t *d_bloom, uint64_t n_bits, cudaStream_t stream);
cudaError_t launch_fill_synthetic(uint8_t *matrix, uint64_t pitch, uint64_t num_rows, uint64_t num_cols,
                                  uint64_t col_offset, uint64_t seed, int and_draws,
                                  const uint64_t *d_planted_cols, const uint32_t *d_planted_thr, int n_planted,
                                  cudaStream_t stream);

// ---- build path / scoring support (build_kernels.cu) -----------------------------------------
cudaError_t launch_bloom_set_bits(const int32_t *d_rows, uint64_t n, uint8_t *d_bloom, cudaStream_t stream);
cudaError_t launch_transpose_blooms(uint8_t *matrix, uint64_t pitch, uint64_t num_rows, uint64_t col0, uint64_t n_blooms,
                                    const uint8_t *d_blooms, uint64_t bloom_stride, uint64_t n_bits, cudaStream_t stream);
cudaError_t launch_hash_windows(const uint8_t *d_seq, uint64_t n_windows, int k, int h, uint64_t m, int32_t *d_rows_out,
                                cudaStream_t stream);
cudaError_t launch_presence(const uint8_t *matrix, uint64_t pitch, const int32_t *d_rows, uint64_t n_windows, int h,
                            const int32_t *d_cols, uint64_t n_cols, uint8_t *d_out, cudaStream_t stream);

}  // namespace bigsi
