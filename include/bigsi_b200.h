/*
 * bigsi_b200.h -- C ABI of the B200-native BIGSI query engine (libbigsi_b200.so).
 *
 * Drop-in boundary for the reference's search hot path (Phelimb/BIGSI v0.3.8; citations are
 * relative to the reference tree, bigsi/...).  The m x N bit-sliced signature matrix that the
 * reference keeps one row per key in RocksDB/BerkeleyDB/Redis (storage/base.py:86-109) lives
 * here as ONE packed array in HBM: row r starts at r * row_pitch_bytes, bytes are exactly the
 * reference's `bitarray.tobytes()` (column c -> byte c>>3, mask 0x80>>(c&7); storage/base.py:86-99),
 * the pitch is padded to a multiple of 128 bytes and all padding bits are zero.
 *
 * Conventions: plain C types only; every function returns 0 (BIGSI_B200_OK) or a negative
 * error code and leaves a message for bigsi_b200_last_error() (thread local); the caller owns
 * every host buffer; `_dev` entry points take DEVICE pointers plus a cudaStream_t (passed as
 * void*) and are stream-ordered/asynchronous, all other entry points synchronise before they
 * return.  One index handle = one column shard on one GPU; multi-GPU deployments run one
 * process per GPU (column sharding).  A handle owns scratch memory, so calls on
 * one handle must not overlap (one stream / one host thread at a time).  There is NO CPU
 * fallback: without a CUDA device every compute entry point fails with BIGSI_B200_ERR_NO_DEVICE
 * or BIGSI_B200_ERR_CUDA.
 */
#ifndef BIGSI_B200_H
#define BIGSI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BIGSI_B200_ABI_VERSION 7

enum {
    BIGSI_B200_OK = 0,
    BIGSI_B200_ERR_INVALID = -1,   /* bad argument */
    BIGSI_B200_ERR_CUDA = -2,      /* CUDA runtime error (message has the cudaError string) */
    BIGSI_B200_ERR_OOM = -3,       /* device or pinned-host allocation failed */
    BIGSI_B200_ERR_NO_DEVICE = -4, /* no usable CUDA device */
    BIGSI_B200_ERR_RANGE = -5,     /* row / column / capacity out of range */
    BIGSI_B200_ERR_TIMEOUT = -6    /* a bounded device-side wait timed out (a peer shard never launched, ...): the
                                      handle is unusable afterwards; every later call on it returns this code */
};

/* query modes */
enum {
    BIGSI_B200_MODE_COUNTS = 0, /* per-column k-mer counts: graph/bigsi.py:35-44,211-219 (unpack_and_sum) */
    BIGSI_B200_MODE_AND = 1     /* AND over all k-mers:     graph/bigsi.py:192-195 (exact_filter)         */
};

typedef struct bigsi_b200_index bigsi_b200_index; /* opaque */

typedef struct {
    uint64_t num_rows;        /* m  (Bloom filter size; "number_of_rows", matrix/bitmatrix.py:3)     */
    uint64_t num_cols;        /* N of this shard ("number_of_cols", matrix/bitmatrix.py:4)            */
    uint64_t col_capacity;    /* columns the pitch can hold (insert grows num_cols up to this)        */
    uint64_t col_offset;      /* global colour of local column 0 (multiple of 8)                      */
    uint64_t row_bytes;       /* ceil(num_cols / 8)                                                   */
    uint64_t row_pitch_bytes; /* multiple of 128                                                      */
    uint64_t matrix_bytes;    /* num_rows * row_pitch_bytes resident in HBM                           */
    int32_t device;
    int32_t sm_count;
    /* statistics of the most recent query launch on this handle */
    uint64_t last_kmers;             /* k-mer lookups in that launch                                  */
    uint64_t last_algorithmic_bytes; /* kmers * h * row_bytes                 */
    uint32_t last_grid, last_block, last_smem_bytes;
    uint32_t last_tile_bytes, last_n_tiles, last_kmers_per_stage, last_n_stages, last_n_slices;
    uint64_t kernel_launches;        /* cumulative count of kernels this handle has launched          */
    uint64_t scratch_bytes;          /* partial-plane workspace currently allocated                   */
    uint32_t last_fused;             /* bit 0: merge ran inside the fused kernel, bit 1: k-mers hashed in it,
                                        bit 3: streamed launch (gather kernel; stage 2 deferred)                */
    uint32_t last_reduce_grid;       /* CTAs of the flush kernel of a streamed launch (else 0)                */
    uint64_t last_unique_kmers;      /* batches with shared row-gather reuse: classes of equal row-id tuples whose rows
                                        were gathered once (0: the launch gathered every k-mer's rows itself)   */
} bigsi_b200_info;

/* ---- library ---------------------------------------------------------------------------- */
int bigsi_b200_abi_version(void);
const char *bigsi_b200_last_error(void);
int bigsi_b200_device_count(int *count_out);

/* Pinned host buffers for the host<->device legs of the host-buffer entry points. */
int bigsi_b200_host_alloc(uint64_t bytes, void **ptr_out);
int bigsi_b200_host_free(void *ptr);

/* ---- index lifecycle (replaces get_storage()/BitMatrix.create: storage/__init__.py:18-19,
 *      matrix/bitmatrix.py:19-25) ------------------------------------------------------------ */
int bigsi_b200_index_create(int device, uint64_t num_rows, uint64_t num_cols, uint64_t col_capacity,
                            uint64_t col_offset, bigsi_b200_index **index_out);
int bigsi_b200_index_destroy(bigsi_b200_index *index); /* storage.delete_all()/close(): graph/bigsi.py:249-250 */
int bigsi_b200_index_get_info(const bigsi_b200_index *index, bigsi_b200_info *info_out);

/* Tuning / instrumentation knobs (value 0 = automatic): "tile_bytes", "grid", "kmers_per_stage",
 * "n_stages", "ctas_per_sm", "merge_chunk_bytes", "debug_flags"; "prehash" / "fuse_merge" / "solo" /
 * "zero_copy" (default 1; 0 forces the separate hash / merge kernels, the generic single-query
 * path, the staged host copies); "batch_reuse" (default 1: a batch of >= 2 queries and >= 16 384 k-mers is
 * de-duplicated by row-id tuple first; when at most half of its k-mers are distinct, the distinct ones' AND vectors
 * are gathered ONCE into scratch and the queries count over those -- shared row-gather reuse; 0 = never);
 * "direct" (default 1: batch queries that lie inside one slice are finished by the CTA that counted them; 0 = every
 * query goes through the merge); "defer" (default 1; 0 = deferred entry points flush at once); "self_merge" (default 0;
 * 1 = a synchronous single query is merged by its own kernel's team, cooperative launch, instead of the flush kernel);
 * diagnostics of the multi-GPU query broadcast: "push_repeat" (every line stored 1 + value times), "push_all_warps"; "pool_pct" (0..100: share of a single query's k-mers that the CTAs claim
 * dynamically; default / > 100 = automatic: 12 for an isolated query of the synchronous host calls, 0 for streamed
 * back-to-back queries); "cooperative" (default 1: a generic-path kernel that merges behind its own grid barrier is
 * launched with the cooperative attribute, so the driver verifies that all its CTAs are co-resident; 0 = plain
 * launch, only safe when nothing else runs on the device); "inputs_ready" (default 0; 1 = the k-mer buffers handed
 * to the `_dev` single-query entry points are never produced by the kernel that precedes the call in the stream
 * -- e.g. they are resident, or were copied in -- so a streamed query does not wait for its predecessor and
 * consecutive queries overlap fully); "spin_timeout_ms" (default 10000: bound of every device-side wait, see
 * BIGSI_B200_ERR_TIMEOUT); "timing" (1 = bracket the gather / fused kernel and the reduce / merge kernel of every
 * query launch with CUDA events, read back with bigsi_b200_index_timing_collect; serialises the launches). */
int bigsi_b200_index_set_option(bigsi_b200_index *index, const char *key, int64_t value);
/* Synchronises, sums the event-timed durations recorded since the last collect and resets them.
 * fused_ms = fused gather-AND-count kernel, merge_ms = merge kernel, n = query launches timed. */
int bigsi_b200_index_timing_collect(bigsi_b200_index *index, double *fused_ms_out, double *merge_ms_out,
                                    uint64_t *n_out);

/* 0, or BIGSI_B200_ERR_TIMEOUT (with the reason in bigsi_b200_last_error) once a device-side wait of this handle
 * has timed out.  Host calls that wait for a result check it themselves; callers of the stream-ordered `_dev`
 * entry points use this after synchronising their stream. */
int bigsi_b200_index_status(bigsi_b200_index *index);

/* Debug aid: with option "debug_flags" bit 1 set, every CTA of the fused kernel records 8 uint64
 * globaltimer stamps (ns) of its last launch; this copies the first n_words of them to the host. */
int bigsi_b200_index_debug_read(bigsi_b200_index *index, uint64_t *out, uint64_t n_words);

/* Rows in the reference's byte layout (BitMatrix.set_rows: matrix/bitmatrix.py:42-44 ->
 * storage/base.py:91-94).  Source row i starts at rows + i*src_stride; the shard's bytes are
 * taken from byte offset src_byte_offset (= col_offset/8 when uploading from a full-width row). */
int bigsi_b200_index_upload_rows(bigsi_b200_index *index, uint64_t row0, uint64_t n_rows,
                                 const uint8_t *rows, uint64_t src_stride, uint64_t src_byte_offset);
/* BitMatrix.get_rows (matrix/bitmatrix.py:30-37): row_bytes bytes per row into out + i*dst_stride. */
int bigsi_b200_index_download_rows(const bigsi_b200_index *index, uint64_t row0, uint64_t n_rows,
                                   uint8_t *out, uint64_t dst_stride);
/* BitMatrix.insert_column (matrix/bitmatrix.py:67-75 -> storage/base.py:111-122): write Bloom
 * filter bits (MSB-first, n_bits <= num_rows; rows >= n_bits get 0) into LOCAL column `col`;
 * col == num_cols appends (num_cols grows by one, up to col_capacity). */
int bigsi_b200_index_set_column(bigsi_b200_index *index, uint64_t col, const uint8_t *bloom_msb_first,
                                uint64_t n_bits);
/* Synthetic matrix, a pure function of (seed, row, GLOBAL column) -- DESIGN.md "Synthetic index".
 * density = 2^-and_draws; planted columns are GLOBAL ids with a u32 threshold (0xffffffff = all ones). */
int bigsi_b200_index_fill_synthetic(bigsi_b200_index *index, uint64_t seed, int and_draws,
                                    const uint64_t *planted_cols, const uint32_t *planted_thr, int n_planted);

/* ---- k-mer -> row ids (replaces convert_query_kmer + generate_hashes: utils/fncts.py:47-54,
 *      bloom/bloomfilter.py:5-13, graph/index.py:62-70).  kmers = n*k raw ASCII bytes;
 *      rows_out = n*h int32: signed MurmurHash3_x86_32 of the k bytes, seeds 0..h-1, floor-mod m.
 *      canonical != 0 hashes min(kmer, reverse complement) (the search/bloom paths); canonical == 0
 *      hashes the bytes as given (bare generate_hashes). */
int bigsi_b200_hash_kmers(int device, const char *kmers, uint64_t n, int k, int h, uint64_t m,
                          int canonical, int32_t *rows_out);
int bigsi_b200_hash_kmers_dev(const char *d_kmers, uint64_t n, int k, int h, uint64_t m, int canonical,
                              int32_t *d_rows_out, void *stream);

/* ---- fused gather-AND-{count|AND} (replaces BitMatrix.get_rows + bitwise_and + exact_filter /
 *      unpack_and_sum: matrix/bitmatrix.py:30-37, graph/index.py:72-80, graph/bigsi.py:35-44,192-219).
 *      d_rows: total_kmers*h row ids; d_q_offsets: n_queries+1 k-mer offsets (q_offsets[0]=0,
 *      q_offsets[n_queries]=total_kmers); max_query_kmers: upper bound on the longest query
 *      (0 = unknown, total_kmers is assumed).
 *      COUNTS: d_out = uint32 [n_queries][out_stride] (out_stride in ELEMENTS >= num_cols).
 *      AND:    d_out = uint8  [n_queries][out_stride] (out_stride in BYTES >= row_bytes),
 *              MSB-first presence bits; an empty query yields all-ones in the valid columns.    */
int bigsi_b200_query_dev(bigsi_b200_index *index, int mode, const int32_t *d_rows,
                         const int64_t *d_q_offsets, uint64_t n_queries, uint64_t total_kmers,
                         uint64_t max_query_kmers, int h, void *d_out, uint64_t out_stride, void *stream);

/* Same gather-AND-count with the threshold fused into the merge kernel (graph/bigsi.py:211-230,241-242):
 * query q keeps the columns whose count >= d_min_kmers[q]; hits are laid out as in
 * bigsi_b200_threshold_dev.  d_counts_full may be NULL (hits only) or receive the full uint32 counts
 * [n_queries][counts_stride] as well. */
int bigsi_b200_query_hits_dev(bigsi_b200_index *index, const int32_t *d_rows, const int64_t *d_q_offsets,
                              uint64_t n_queries, uint64_t total_kmers, uint64_t max_query_kmers, int h,
                              const uint32_t *d_min_kmers, int32_t *d_cols_out, uint32_t *d_counts_out,
                              uint64_t cap, uint64_t *d_n_out, uint32_t *d_counts_full, uint64_t counts_stride,
                              void *stream);

/* The whole search path from raw unique k-mers (n*k ASCII, device): canonicalised and hashed in the kernel
 * prologue, rows gathered/ANDed/counted, merged and thresholded.  One query: a streamed launch (see "streamed
 * single-query launches" below).  A batch: one kernel that merges behind a grid-wide barrier (cooperative
 * launch), or hash kernel + fused kernel + merge kernel when the launch geometry does not allow that.
 * Outputs as bigsi_b200_query_hits_dev. */
int bigsi_b200_query_kmers_hits_dev(bigsi_b200_index *index, const char *d_kmers, int k, const int64_t *d_q_offsets,
                                    uint64_t n_queries, uint64_t total_kmers, uint64_t max_query_kmers, int h,
                                    const uint32_t *d_min_kmers, int32_t *d_cols_out, uint32_t *d_counts_out,
                                    uint64_t cap, uint64_t *d_n_out, uint32_t *d_counts_full,
                                    uint64_t counts_stride, void *stream);

/* Per-k-mer AND vectors (KmerSignatureIndex.lookup: graph/index.py:42-49,75-80): d_out =
 * uint8 [n_kmers][out_stride], out_stride >= row_bytes. */
int bigsi_b200_lookup_dev(bigsi_b200_index *index, const int32_t *d_rows, uint64_t n_kmers, int h,
                          uint8_t *d_out, uint64_t out_stride, void *stream);

/* counts >= min_kmers (graph/bigsi.py:241-242) for a batch of queries: query q keeps the columns
 * with d_counts[q*counts_stride + c] >= d_min_kmers[q] and stores up to `cap` (colour, count)
 * pairs at d_cols_out/d_counts_out + q*cap (order unspecified; colours are LOCAL column ids);
 * d_n_out[q] receives the number of hits (may exceed cap; only the first cap are stored). */
int bigsi_b200_threshold_dev(const uint32_t *d_counts, uint64_t counts_stride, uint64_t n_queries,
                             uint64_t num_cols, const uint32_t *d_min_kmers, int32_t *d_cols_out,
                             uint32_t *d_counts_out, uint64_t cap, uint64_t *d_n_out, void *stream);

/* ---- host-buffer calls (the reference-facing plugin path: everything below takes and returns
 *      HOST memory; H2D/D2H copies happen inside). ------------------------------------------- */
/* Queries given as raw k-mers (n*k ASCII, already unique per query -- graph/index.py:45). */
int bigsi_b200_search_kmers(bigsi_b200_index *index, int mode, const char *kmers,
                            const int64_t *q_offsets, uint64_t n_queries, int k, int h,
                            void *out, uint64_t out_stride);
/* Queries given as precomputed row ids. */
int bigsi_b200_search_rows(bigsi_b200_index *index, int mode, const int32_t *rows,
                           const int64_t *q_offsets, uint64_t n_queries, int h,
                           void *out, uint64_t out_stride);
/* Fused search + threshold (BIGSI.search's inexact_filter, graph/bigsi.py:211-230): raw unique
 * k-mers in, compact (colour, count) hits out; layout as bigsi_b200_threshold_dev. */
int bigsi_b200_search_kmers_hits(bigsi_b200_index *index, const char *kmers, const int64_t *q_offsets,
                                 uint64_t n_queries, int k, int h, const uint32_t *min_kmers,
                                 int32_t *cols_out, uint32_t *counts_out, uint64_t cap, uint64_t *n_out);
/* BIGSI.search's whole filter stage for ONE sequence (graph/bigsi.py:174-230) on the device: every
 * window of length k (utils/fncts.py:63-65), de-duplicated as RAW byte strings (graph/index.py:45) -- for k <= 32
 * inside the gather kernel itself: every CTA stages its span of the sequence (read straight out of pinned host
 * memory) and claims its windows in an exact hash table; else by a front-end kernel --,
 * *num_kmers_out = U = number of unique k-mers, min_kmers = ceil(U * threshold) in IEEE double
 * (<= 0 keeps every column), then canonical + hash + gather-AND-count + threshold.  Hits as in
 * bigsi_b200_search_kmers_hits for one query (order unspecified, *n_hits_out may exceed cap).
 * len < k: no window, both outputs 0 (the reference raises TypeError there; the caller decides). */
int bigsi_b200_search_sequence(bigsi_b200_index *index, const char *seq, uint64_t len, int k, int h, double threshold,
                               int32_t *cols_out, uint32_t *counts_out, uint64_t cap, uint64_t *n_hits_out,
                               uint64_t *num_kmers_out);
/* The same search split in two halves, so that several searches are in flight on one handle (and one host thread
 * can drive several handles, i.e. several column shards, at once): submit stages the sequence and launches the
 * kernels, wait returns the result of that ticket.  Up to 8 tickets may be outstanding per handle; tickets are
 * collected in any order.  Consecutive searches overlap on the device (see "streamed single-query launches"). */
int bigsi_b200_search_sequence_submit(bigsi_b200_index *index, const char *seq, uint64_t len, int k, int h, double threshold,
                                      uint64_t cap, uint64_t *ticket_out);
int bigsi_b200_search_sequence_wait(bigsi_b200_index *index, uint64_t ticket, int32_t *cols_out, uint32_t *counts_out,
                                    uint64_t cap, uint64_t *n_hits_out, uint64_t *num_kmers_out);
/* bulk_search (bigsi/__main__.py:261-314: every record of a FASTA file through BIGSI.search with one threshold):
 * sequence q = seqs[offsets[q] .. offsets[q+1]); outputs of sequence q at cols_out / counts_out + q*cap,
 * n_hits_out[q], num_kmers_out[q], each exactly as bigsi_b200_search_sequence would return them.  The searches are
 * pipelined (up to 7 in flight). */
int bigsi_b200_search_sequences(bigsi_b200_index *index, const char *seqs, const uint64_t *offsets, uint64_t n_seqs, int k,
                                int h, double threshold, int32_t *cols_out, uint32_t *counts_out, uint64_t cap,
                                uint64_t *n_hits_out, uint64_t *num_kmers_out);
/* lookup(): per-k-mer AND vectors to host, out = uint8 [n][out_stride]. */
int bigsi_b200_lookup_kmers(bigsi_b200_index *index, const char *kmers, uint64_t n, int k, int h,
                            uint8_t *out, uint64_t out_stride);

/* ---- build path ------------------------------------------------
 * BIGSI.bloom (graph/bigsi.py:150-155 -> bloom/bloomfilter.py:16-32): the m-bit Bloom filter of n
 * k-mers (n*k raw ASCII bytes), hashed like bigsi_b200_hash_kmers and written as the reference's
 * `bitarray.tobytes()` / .bloom file bytes (ceil(m/8) bytes, MSB first, cmds/bloom.py:26-27). */
int bigsi_b200_bloom_kmers(int device, const char *kmers, uint64_t n, int k, int h, uint64_t m, int canonical,
                           uint8_t *bloom_out);
/* BIGSI.build's transpose and bulk insert (matrix/transpose.py:33-50 -> graph/index.py:27-40 ->
 * matrix/bitmatrix.py:19-25,67-75) as one bit-transpose kernel: n_blooms Bloom filters of n_bits bits
 * (MSB-first bytes, filter i at blooms + i*bloom_stride; rows >= n_bits get 0) become the LOCAL
 * columns [col0, col0+n_blooms).  col0 <= num_cols; num_cols grows to col0+n_blooms (<= col_capacity);
 * other columns keep their bits.  The _dev variant takes filters already in device memory (16-byte
 * aligned, stride a multiple of 32 bytes and >= ceil(m/256)*32) and is stream-ordered. */
int bigsi_b200_index_build_columns(bigsi_b200_index *index, uint64_t col0, uint64_t n_blooms, const uint8_t *blooms,
                                   uint64_t bloom_stride, uint64_t n_bits);
int bigsi_b200_index_build_columns_dev(bigsi_b200_index *index, uint64_t col0, uint64_t n_blooms, const uint8_t *d_blooms,
                                       uint64_t bloom_stride, uint64_t n_bits, void *stream);

/* ---- score=True support (graph/bigsi.py:232-239 with unpack_and_cat,
 * graph/bigsi.py:47-56): for EVERY window of length k of seq (duplicates included, sequence order) and
 * each of the n_cols LOCAL columns: out[c*n_windows + w] = '1' if the window's canonical k-mer is
 * present in column cols[c] (AND of its h rows), else '0' -- row c of `out` is the reference's
 * "kmer-presence" string of that hit.  n_windows = len-k+1; nothing is written when len < k. */
int bigsi_b200_sequence_presence(bigsi_b200_index *index, const char *seq, uint64_t len, int k, int h,
                                 const int32_t *cols, uint64_t n_cols, uint8_t *out);

/* ---- persistence (replaces the durability of storage/*.py) ---------
 * Flat file: this header, meta_bytes of caller-defined metadata (the Python layer stores JSON: k, h,
 * the metadata:* keys of graph/metadata.py), zero padding up to rows_offset, then num_rows rows of
 * row_bytes bytes each -- the concatenation of the reference's "<row>:bitarray" values
 * (storage/base.py:86-94) in row order.  Little-endian. */
typedef struct {
    char magic[8];        /* "BIGSIB2\n" */
    uint32_t version;     /* 1 */
    uint32_t header_bytes;
    uint64_t num_rows, num_cols, col_offset, row_bytes, meta_bytes, rows_offset;
} bigsi_b200_file_header;
/* HBM -> file through two pinned buffers (the copy of chunk i+1 overlaps the write of chunk i). */
int bigsi_b200_index_save(bigsi_b200_index *index, const char *path, const void *meta, uint64_t meta_bytes);
/* Header (+ up to meta_cap bytes of metadata) of an index file. */
int bigsi_b200_file_info(const char *path, bigsi_b200_file_header *header_out, void *meta_out, uint64_t meta_cap);
/* file -> HBM through two pinned buffers (the read of chunk i+1 overlaps the upload of chunk i): rows
 * [row0, row0+n_rows) of the index are taken from file_offset + i*file_stride + src_byte_offset (a
 * column shard loads its byte range of a full-width file: src_byte_offset = col_offset/8).  Works on
 * any file of fixed-stride rows in the reference's byte layout, not only on bigsi_b200_index_save's. */
int bigsi_b200_index_load_rows(bigsi_b200_index *index, const char *path, uint64_t file_offset, uint64_t file_stride,
                               uint64_t src_byte_offset, uint64_t row0, uint64_t n_rows);

/* ---- streamed single-query launches -----------------------------------------------------------
 * A single query whose k-mers are hashed in the kernel (one column tile, up to ~200 000 k-mers) runs in two
 * stages.  Stage 1, the gather kernel: hash, row gather, AND, vertical count; one CTA per SM; it carries the
 * programmatic-dependent-launch attribute and its gather warps do not wait for the preceding kernel, so the gather
 * CTA of query s+1 takes over an SM the moment the gather CTA of query s leaves it.  Stage 2 (merge of the CTAs'
 * bit planes, threshold, publication) of query s is executed by the MERGE TEAM of the gather kernel of query s+1 --
 * four extra warps per CTA that wait for query s's kernel to complete and work in the shadow of the row stream --
 * or, when nothing follows, by a flush kernel.  No grid-wide barrier, no cooperative launch, one CTA per SM at all
 * times.  Scratch rotates over 4 queries.
 *   - The ordinary entry points (bigsi_b200_query_kmers_hits_dev, the host-buffer calls) launch the flush kernel
 *     right behind the gather kernel: results are complete in stream order after the call, as usual.
 *   - The DEFERRED entry points (bigsi_b200_query_kmers_hits_stream_dev, bigsi_b200_exchange_search_dev,
 *     bigsi_b200_search_sequence_submit, bigsi_b200_search_sequences) leave stage 2 of a query to the next streamed
 *     query of the handle: the result of call s is complete in stream order after call s+1 (any streamed query on
 *     the same stream) or after bigsi_b200_index_flush.  Any other query call on the handle flushes first.
 *   - Output buffers handed to one single-query call must not be handed to any of the next 3 single-query calls
 *     on the same handle.
 *   - See option "inputs_ready".
 *
 * query_kmers_hits_stream_dev: ONE query of n_kmers raw unique k-mers (device), threshold by value; outputs as
 * bigsi_b200_query_hits_dev for one query (d_n_out: one u64).  Falls back to an ordinary, complete-in-stream-order
 * launch when the launch plan is not the streamed one (rows wider than one tile, very long queries).
 * index_flush: launches stage 2 of the handle's pending deferred query, if any, on the stream that query was
 * launched on (no host synchronisation). */
int bigsi_b200_query_kmers_hits_stream_dev(bigsi_b200_index *index, const char *d_kmers, int k, uint64_t n_kmers, int h,
                                           uint32_t min_kmers, int32_t *d_cols_out, uint32_t *d_counts_out, uint64_t cap,
                                           uint64_t *d_n_out, void *stream);
int bigsi_b200_index_flush(bigsi_b200_index *index);

/*
 * ---- column-sharded search over several GPUs WITHOUT per-query collectives ------------------
 * The reference has no distributed path; sample columns are independent (graph/index.py:42-80,
 * graph/bigsi.py:192-230), so shard g holds all rows of its column range on its own GPU and a
 * query needs two exchanges: the query itself to every shard, the per-shard hits back.  Both are
 * fused into the query kernels: rank 0's gather kernel stores the k-mer bytes into its peers' inboxes over
 * NVLink as "low-latency lines" (every 8 bytes carry 4 data bytes and the query's 32-bit sequence
 * number, so the receiving kernel spins per 16-byte line and no fence or separate flag is needed);
 * every rank's stage 2 (the merge team of the next query's gather kernel, or the flush kernel) publishes its
 * hit list into slot `rank` of every rank's result blocks and finishes only when all slots of its own copy have
 * arrived (all-gather semantics in stream order) -- while the next query's rows are already streaming, so
 * the shards do not wait for each other on the critical path.  Every device-side wait is bounded ("spin_timeout_ms"): a rank whose peers
 * never launch gets BIGSI_B200_ERR_TIMEOUT instead of a hung GPU.  One handle per GPU; handles may live in
 * different processes (CUDA IPC) or in one (open_local; they may even share a device).  Calls are SPMD:
 * every rank calls exchange_search_dev once per query with the same n_kmers / k / h / min_kmers.
 *
 * create: allocates this rank's block; ipc_handle_out (64 bytes, may be NULL) is what the other
 *         processes pass to open.  spec = hits one result block holds (longer hit lists are cut,
 *         the count stays exact); max_kmer_bytes = largest query (n_kmers * k).
 * open / open_local: map the peers' blocks (handles: world x 64 bytes in rank order; peers: world
 *         handles of the same process).
 * search_dev: d_kmers = n_kmers * k raw unique k-mers, 16-byte aligned, addressable by rank 0's device
 *         (device memory or mapped pinned host memory; ignored on other ranks).  *d_blocks_out = device
 *         pointer to `world` result blocks of *block_bytes_out bytes each, block r = { u64 seq; u64
 *         n_hits; int32 cols[spec]; uint32 counts[spec] } with LOCAL column ids of shard r; a DEFERRED
 *         call (see above): complete in stream order after the NEXT search_dev call or bigsi_b200_index_flush
 *         (every rank flushes), valid until 4 more searches have been issued on this handle (eight generations
 *         of blocks rotate).
 * reserve: optional; see below.
 * wait_ns: synchronises the device; returns (and resets) the sum over the queries since the last call of
 *         the time this rank's stage 2 waited for the other shards' hit lists (diagnostics). */
int bigsi_b200_exchange_create(bigsi_b200_index *index, int world, int rank, uint64_t max_kmer_bytes, uint32_t spec,
                               uint8_t *ipc_handle_out);
int bigsi_b200_exchange_open(bigsi_b200_index *index, const uint8_t *ipc_handles);
int bigsi_b200_exchange_open_local(bigsi_b200_index *index, bigsi_b200_index *const *peers);
int bigsi_b200_exchange_search_dev(bigsi_b200_index *index, const char *d_kmers, uint64_t n_kmers, int k, int h,
                                   uint32_t min_kmers, void *stream, const void **d_blocks_out,
                                   uint64_t *block_bytes_out);
/* Allocates the scratch a search of up to max_kmers k-mers needs now, so that no later search has to grow a buffer
 * (growing frees device memory, which waits for every kernel on the device -- also for a kernel of ANOTHER shard of
 * this process on the same GPU that is itself waiting for this shard's launch). */
int bigsi_b200_exchange_reserve(bigsi_b200_index *index, uint64_t max_kmers, int k, int h);
/* Host results (optional; enable = 1 / 0 at any time, per rank): the all-gathered hit lists of the searches launched
 * while it is on are ALSO written by the stage-2 code into mapped host memory -- no copy operation in the stream, so
 * back-to-back searches keep overlapping while the host consumes earlier results.  It costs the last stage-2 CTA of
 * every query a system-scope fence over PCIe (~3 us): leave it off on ranks that do not consume results on the host.  last_seq: the number of the latest search_dev call on this handle
 * (1, 2, ...).  wait_host(seq): flushes the search if it is the newest one (SPMD: then every rank must flush or
 * search on), polls until its block is complete and returns a HOST pointer to `world` blocks in the layout
 * search_dev documents; valid until 8 more searches have been issued. */
int bigsi_b200_exchange_host_results(bigsi_b200_index *index, int enable);
int bigsi_b200_exchange_last_seq(bigsi_b200_index *index, uint64_t *seq_out);
int bigsi_b200_exchange_wait_host(bigsi_b200_index *index, uint64_t seq, const void **blocks_out, uint64_t *block_bytes_out);
int bigsi_b200_exchange_wait_ns(bigsi_b200_index *index, uint64_t *wait_ns_out, uint64_t *queries_out);
int bigsi_b200_exchange_destroy(bigsi_b200_index *index);

#ifdef __cplusplus
}
#endif
#endif /* BIGSI_B200_H */
