// Helper kernels around the fused query kernel: k-mer hashing, per-k-mer lookup vectors,
// threshold/compaction, column insert and the synthetic index generator.
#include "hash.cuh"
#include "launch.cuh"
#include "ptx.cuh"
#include "query.cuh"

namespace bigsi {

// ------------------------------------------------------------------------------------------
// K1: canonical k-mer + MurmurHash3_x86_32 (device code in hash.cuh), one block per group of k-mers.
// ------------------------------------------------------------------------------------------
constexpr int kHashThreads = 128;
constexpr int kHashSmemBytes = 40 * 1024;

__global__ void __launch_bounds__(kHashThreads) hash_kmers_kernel(const uint8_t *__restrict__ kmers, uint64_t n, int k,
                                                                 int h, uint32_t m, int canonical, uint32_t kpb,
                                                                 int use_smem, int32_t *__restrict__ rows_out, uint64_t magic)
{
    extern __shared__ __align__(16) uint8_t sk[];
    grid_dependency_wait();  // the previous query's fused kernel may still be reading rows_out
    grid_launch_dependents();
    const uint64_t base = (uint64_t)blockIdx.x * kpb;
    const uint32_t cnt = (uint32_t)min((uint64_t)kpb, n - base);
    const uint8_t *g0 = kmers + base * (uint64_t)k;
    const int nblocks = k >> 2, rem = k & 3;
    if (use_smem) {
        hash_kmers_cooperative(g0, cnt, k, h, m, canonical, sk, rows_out + base * (uint64_t)h, magic);
        return;
    }
    // very long "k-mers" (k > the staging buffer): one thread per (k-mer, seed) straight from global
    for (uint32_t w = threadIdx.x; w < cnt * (uint32_t)h; w += kHashThreads) {
        const uint32_t km = w / (uint32_t)h, seed = w % (uint32_t)h;
        const uint8_t *s = g0 + (size_t)km * k;
        bool fwd = true;
        if (canonical) {
            for (int j = 0; j < k; ++j) {
                const uint32_t a = s[j], b = comp_base(s[k - 1 - j]);
                if (a != b) {
                    fwd = a < b;
                    break;
                }
            }
        }
        auto byte_at = [&](int j) -> uint32_t { return fwd ? (uint32_t)s[j] : comp_base(s[k - 1 - j]); };
        uint32_t h1 = seed;
        for (int b = 0; b < nblocks; ++b)
            h1 = murmur_block(h1, byte_at(4 * b) | (byte_at(4 * b + 1) << 8) | (byte_at(4 * b + 2) << 16) |
                                      (byte_at(4 * b + 3) << 24));
        if (rem) {
            uint32_t k1 = 0;
            for (int u = 0; u < rem; ++u) k1 |= byte_at(4 * nblocks + u) << (8 * u);
            h1 = murmur_tail(h1, k1);
        }
        rows_out[(base + km) * (uint64_t)h + seed] = murmur_finish_mod(h1, (uint32_t)k, m);
    }
}

cudaError_t launch_hash_kmers(const char *d_kmers, uint64_t n, int k, int h, uint64_t m, int canonical,
                              int32_t *d_rows_out, cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    // k-mers per block: about two (k-mer, seed) items per thread, bounded by the staging buffer
    uint32_t kpb = (uint32_t)((2 * kHashThreads + h - 1) / h);
    const uint64_t per_kmer = prehash_bytes_per_kmer((uint32_t)k);  // raw + flag + words
    int use_smem = 1;
    if (per_kmer + 96 > (uint64_t)kHashSmemBytes) {
        use_smem = 0;
    } else if ((uint64_t)kpb * per_kmer + 96 > (uint64_t)kHashSmemBytes) {
        kpb = (uint32_t)(((uint64_t)kHashSmemBytes - 96) / per_kmer);
    }
    if (kpb < 1) kpb = 1;
    const uint64_t blocks = (n + kpb - 1) / kpb;
    if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    const size_t smem = use_smem ? (size_t)kpb * per_kmer + 96 : 0;
    return launch_pdl(hash_kmers_kernel, dim3((unsigned)blocks), dim3(kHashThreads), smem, stream,
                      reinterpret_cast<const uint8_t *>(d_kmers), n, k, h, (uint32_t)m, canonical, kpb, use_smem,
                      d_rows_out, mod_magic((uint32_t)m));
}

// ------------------------------------------------------------------------------------------
// K5: per-k-mer AND vectors (KmerSignatureIndex.lookup, bigsi/graph/index.py:42-49,75-80).
// One thread per (k-mer, 16-byte unit); output keeps the reference's MSB-first bytes.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lookup_kernel(const uint8_t *__restrict__ matrix, uint64_t pitch,
                                                     uint32_t row_bytes, uint32_t units, const int32_t *__restrict__ rows,
                                                     uint64_t n_kmers, int h, uint8_t *__restrict__ out,
                                                     uint64_t out_stride)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t km = tid / units;
    const uint32_t u = (uint32_t)(tid % units);
    if (km >= n_kmers) return;
    uint4 acc = make_uint4(~0u, ~0u, ~0u, ~0u);
    for (int j = 0; j < h; ++j) {
        const uint4 v = ldg128_stream(matrix + (uint64_t)(uint32_t)__ldg(rows + km * h + j) * pitch + u * 16);
        acc.x &= v.x; acc.y &= v.y; acc.z &= v.z; acc.w &= v.w;
    }
    uint8_t *dst = out + km * out_stride + u * 16;
    const uint32_t remain = row_bytes - u * 16;
    if (remain >= 16 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
        *reinterpret_cast<uint4 *>(dst) = acc;
    } else {
        const uint32_t w[4] = {acc.x, acc.y, acc.z, acc.w};
        for (uint32_t b = 0; b < min(remain, 16u); ++b) dst[b] = (uint8_t)(w[b >> 2] >> (8 * (b & 3)));
    }
}

cudaError_t launch_lookup(const uint8_t *matrix, uint64_t pitch, uint32_t row_bytes, const int32_t *d_rows,
                          uint64_t n_kmers, int h, uint8_t *d_out, uint64_t out_stride, cudaStream_t stream)
{
    if (n_kmers == 0 || row_bytes == 0) return cudaSuccess;
    const uint32_t units = (row_bytes + 15) / 16;
    const uint64_t threads = n_kmers * units;
    const uint64_t blocks = (threads + 255) / 256;
    if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    lookup_kernel<<<(unsigned)blocks, 256, 0, stream>>>(matrix, pitch, row_bytes, units, d_rows, n_kmers, h, d_out,
                                                        out_stride);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K12: shared row-gather reuse for batches (BASELINE configs[4]; the reference gathers every distinct row of ONE
// query once, graph/index.py:45-48 -- here across the queries of a batch).  Two k-mers gather the same bytes iff
// their h row ids agree, so the batch's k-mers are de-duplicated by their row-id tuples: an open-addressing table
// keyed by a fingerprint of the tuple (equal tags are confirmed by comparing the ids, so the classes are exact),
// the first k-mer to claim an entry represents its class and takes the next free unique id.  Pass 2 gives every
// k-mer the id of its class and copies the representatives' row ids.  Table entry: tag (32) | k-mer index + 1 (32).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t row_tuple_fingerprint(const int32_t *r, int h)
{
    uint64_t fp = 0x9e3779b97f4a7c15ull ^ (uint64_t)h;
    for (int j = 0; j < h; ++j) {
        fp = (fp ^ (uint64_t)(uint32_t)__ldg(r + j)) * 0xff51afd7ed558ccdull;
        fp ^= fp >> 29;
    }
    fp *= 0xc4ceb9fe1a85ec53ull;
    return fp ^ (fp >> 32);
}

__global__ void __launch_bounds__(256) dedup_rows_claim_kernel(const int32_t *__restrict__ rows, uint64_t n, int h,
                                                              unsigned long long *__restrict__ table, uint64_t mask,
                                                              uint32_t *__restrict__ rep, uint32_t *__restrict__ uid_of,
                                                              unsigned int *__restrict__ counter)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool claimed = false;
    if (t < n) {
        const int32_t *mine = rows + t * (uint64_t)h;
        const uint64_t fp = row_tuple_fingerprint(mine, h);
        const unsigned long long entry = ((fp >> 32) << 32) | (unsigned long long)(uint32_t)(t + 1);
        uint64_t slot = fp & mask;
        for (;;) {
            unsigned long long cur = ld_volatile_u64(table + slot);
            if (cur == 0ull) {
                const unsigned long long old = atomicCAS(table + slot, 0ull, entry);
                if (old == 0ull) {  // this k-mer represents its class
                    rep[t] = (uint32_t)t;
                    claimed = true;
                    break;
                }
                cur = old;
            }
            if ((cur >> 32) == (fp >> 32)) {
                const uint64_t o = (cur & 0xffffffffull) - 1;
                const int32_t *other = rows + o * (uint64_t)h;
                bool eq = true;
                for (int j = 0; j < h && eq; ++j) eq = __ldg(other + j) == __ldg(mine + j);
                if (eq) {
                    rep[t] = (uint32_t)o;
                    break;
                }
            }
            slot = (slot + 1) & mask;
        }
    }
    // unique ids: one atomic per warp (a batch of all-distinct k-mers would otherwise serialise a million atomics on one word)
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t m = __ballot_sync(0xffffffffu, claimed);
    if (m) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(counter, (unsigned int)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (claimed) uid_of[t] = base + __popc(m & ((1u << lane) - 1u));
    }
}

__global__ void __launch_bounds__(256) dedup_rows_assign_kernel(const int32_t *__restrict__ rows, uint64_t n, int h,
                                                               const uint32_t *__restrict__ rep, const uint32_t *__restrict__ uid_of,
                                                               int32_t *__restrict__ ids_out, int32_t *__restrict__ unique_rows)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint32_t r = rep[t];
    const uint32_t u = uid_of[r];
    ids_out[t] = (int32_t)u;
    if (r == (uint32_t)t)
        for (int j = 0; j < h; ++j) unique_rows[(uint64_t)u * h + j] = __ldg(rows + t * (uint64_t)h + j);
}

cudaError_t launch_dedup_rows(const int32_t *d_rows, uint64_t n, int h, unsigned long long *d_table, uint64_t table_entries,
                              uint32_t *d_rep, uint32_t *d_uid_of, unsigned int *d_counter, int32_t *d_ids_out,
                              int32_t *d_unique_rows, cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    if (n > 0xfffffff0ull) return cudaErrorInvalidConfiguration;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    dedup_rows_claim_kernel<<<blocks, 256, 0, stream>>>(d_rows, n, h, d_table, table_entries - 1, d_rep, d_uid_of, d_counter);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    dedup_rows_assign_kernel<<<blocks, 256, 0, stream>>>(d_rows, n, h, d_rep, d_uid_of, d_ids_out, d_unique_rows);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// counts >= min_kmers -> compact (colour, count) pairs (bigsi/graph/bigsi.py:241-242).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) threshold_kernel(const uint32_t *__restrict__ counts, uint64_t counts_stride,
                                                        uint32_t num_cols, const uint32_t *__restrict__ min_kmers,
                                                        int32_t *__restrict__ cols_out, uint32_t *__restrict__ counts_out,
                                                        uint64_t cap, unsigned long long *__restrict__ n_out)
{
    __shared__ uint32_t warp_cnt[8];
    __shared__ unsigned long long block_base;
    const uint32_t q = blockIdx.y;
    const uint32_t col = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t thr = __ldg(min_kmers + q);
    uint32_t c = 0;
    bool hit = false;
    if (col < num_cols) {
        c = __ldg(counts + (uint64_t)q * counts_stride + col);
        hit = c >= thr;
    }
    const uint32_t ballot = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) warp_cnt[warp] = __popc(ballot);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
        for (int w = 0; w < 8; ++w) {
            const uint32_t t = warp_cnt[w];
            warp_cnt[w] = tot;
            tot += t;
        }
        block_base = tot ? atomicAdd(n_out + q, (unsigned long long)tot) : 0ull;
    }
    __syncthreads();
    if (hit) {
        const uint64_t pos = block_base + warp_cnt[warp] + __popc(ballot & ((1u << lane) - 1));
        if (pos < cap) {
            cols_out[(uint64_t)q * cap + pos] = (int32_t)col;
            counts_out[(uint64_t)q * cap + pos] = c;
        }
    }
}

cudaError_t launch_threshold(const uint32_t *d_counts, uint64_t counts_stride, uint64_t n_queries, uint64_t num_cols,
                             const uint32_t *d_min_kmers, int32_t *d_cols_out, uint32_t *d_counts_out, uint64_t cap,
                             unsigned long long *d_n_out, cudaStream_t stream)
{
    if (n_queries == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(d_n_out, 0, n_queries * sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    if (num_cols == 0) return cudaSuccess;
    if (n_queries > 65535) return cudaErrorInvalidConfiguration;
    dim3 grid((unsigned)((num_cols + 255) / 256), (unsigned)n_queries);
    threshold_kernel<<<grid, 256, 0, stream>>>(d_counts, counts_stride, (uint32_t)num_cols, d_min_kmers, d_cols_out,
                                               d_counts_out, cap, d_n_out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K6: insert_column (bigsi/matrix/bitmatrix.py:67-75 -> storage/base.py:111-122): one thread per row.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) set_column_kernel(uint8_t *__restrict__ matrix, uint64_t pitch, uint64_t num_rows,
                                                         uint64_t col, const uint8_t *__restrict__ bloom, uint64_t n_bits)
{
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= num_rows) return;
    const uint32_t bit = r < n_bits ? (bloom[r >> 3] >> (7 - (r & 7))) & 1u : 0u;
    uint8_t *p = matrix + r * pitch + (col >> 3);
    const uint8_t mask = (uint8_t)(0x80u >> (col & 7));
    const uint8_t v = *p;
    *p = bit ? (uint8_t)(v | mask) : (uint8_t)(v & ~mask);
}

cudaError_t launch_set_column(uint8_t *matrix, uint64_t pitch, uint64_t num_rows, uint64_t col, const uint8_t *d_bloom,
                              uint64_t n_bits, cudaStream_t stream)
{
    if (num_rows == 0) return cudaSuccess;
    const uint64_t blocks = (num_rows + 255) / 256;
    set_column_kernel<<<(unsigned)blocks, 256, 0, stream>>>(matrix, pitch, num_rows, col, d_bloom, n_bits);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K8: query front-end.  seq_to_kmers + set(kmers) (bigsi/utils/fncts.py:63-65, graph/index.py:45,
// graph/bigsi.py:177-179): every window of length k of the sequence, de-duplicated as RAW byte
// strings.  One thread per window inserts (fingerprint tag, window index) into an open-addressing
// table; a hit with the same tag is confirmed by comparing the k bytes, so the result is exact, not
// probabilistic.  First occurrences are compacted (order unspecified: the query is a set) into a
// dense k-mer array the search kernels take; *counter ends up as U = number of unique k-mers.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dedup_windows_kernel(const uint8_t *__restrict__ seq, uint64_t n, int k,
                                                            unsigned long long *table, uint64_t mask,
                                                            uint8_t *__restrict__ out_kmers, unsigned long long *counter,
                                                            unsigned int *ticket, double threshold, uint32_t *min_out)
{
    // The block's span of the sequence (its 256 windows = 256 + k - 1 bytes) is staged in shared memory with one
    // round trip of 16-byte loads: `seq` may be mapped pinned HOST memory (zero-copy over PCIe), where the k
    // dependent byte loads per window below would each cost a bus round trip.  The aligned window reaches at
    // most 15 bytes past either end of the sequence, inside the allocation's padding.
    extern __shared__ __align__(16) uint8_t span[];
    grid_dependency_wait();   // PDL: the previous query may still be reading out_kmers / the counter
    grid_launch_dependents();
    const uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x;
    const uint64_t i = i0 + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t span_bytes = min((uint64_t)blockDim.x, n - i0) + (uint64_t)k - 1;
    const uint32_t skew = (uint32_t)(reinterpret_cast<uintptr_t>(seq + i0) & 15);
    const uint4 *a0 = reinterpret_cast<const uint4 *>(seq + i0 - skew);
    for (uint32_t v = threadIdx.x; v < (uint32_t)((skew + span_bytes + 15) >> 4); v += blockDim.x)
        reinterpret_cast<uint4 *>(span)[v] = ldg128_stream(a0 + v);
    __syncthreads();
    bool first = false;
    if (i < n) {
        const uint8_t *w = span + skew + threadIdx.x;
        uint64_t fp = 0xcbf29ce484222325ull;  // FNV-1a over the bytes, then a 64-bit finaliser
        for (int j = 0; j < k; ++j) fp = (fp ^ w[j]) * 0x100000001b3ull;
        fp ^= fp >> 33;
        fp *= 0xff51afd7ed558ccdull;
        fp ^= fp >> 33;
        const unsigned long long tag = fp >> 32;
        const unsigned long long entry = (tag << 32) | (unsigned long long)(i + 1);
        uint64_t slot = fp & mask;
        for (;;) {
            unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(table + slot);
            if (cur == 0ull) {
                cur = atomicCAS(table + slot, 0ull, entry);
                if (cur == 0ull) {
                    first = true;
                    break;
                }
            }
            if ((cur >> 32) == tag) {  // same tag: the same k-mer seen earlier, or a true collision
                const uint8_t *o = seq + ((cur & 0xffffffffull) - 1);
                bool same = true;
                for (int j = 0; j < k; ++j)
                    if (o[j] != w[j]) {
                        same = false;
                        break;
                    }
                if (same) break;
            }
            slot = (slot + 1) & mask;
        }
    }
    const uint32_t ballot = __ballot_sync(0xffffffffu, first);
    if (ballot) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counter, (unsigned long long)__popc(ballot));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (first) {
            uint8_t *dst = out_kmers + (base + __popc(ballot & ((1u << lane) - 1))) * (uint64_t)k;
            const uint8_t *w = span + skew + threadIdx.x;
            for (int j = 0; j < k; ++j) dst[j] = w[j];
        }
    }
    if (min_out != nullptr) {
        // the last block to finish knows U: min_kmers = math.ceil(U * threshold) in IEEE double (graph/bigsi.py:179)
        __shared__ bool last;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            last = atomicAdd(ticket, 1u) + 1u == gridDim.x;
        }
        __syncthreads();
        if (last && threadIdx.x == 0) {
            __threadfence();
            const unsigned long long U = *reinterpret_cast<volatile unsigned long long *>(counter);
            const double need = ceil(__dmul_rn((double)U, threshold));
            *min_out = need <= 0.0 ? 0u : need >= 4294967295.0 ? 0xffffffffu : (uint32_t)need;
        }
    }
}

cudaError_t launch_dedup_windows(const uint8_t *d_seq, uint64_t n_windows, int k, unsigned long long *d_table,
                                 uint64_t table_entries, uint8_t *d_out_kmers, unsigned long long *d_counter,
                                 unsigned int *d_ticket, double threshold, uint32_t *d_min_out, cudaStream_t stream)
{
    if (n_windows == 0) return cudaSuccess;
    const uint64_t blocks = (n_windows + 255) / 256;
    if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    const size_t smem = 256 + (size_t)k - 1 + 32;
    if (smem > 48 * 1024) return cudaErrorInvalidConfiguration;
    return launch_pdl(dedup_windows_kernel, dim3((unsigned)blocks), dim3(256), smem, stream, d_seq, n_windows, k, d_table,
                      table_entries - 1, d_out_kmers, d_counter, d_ticket, threshold, d_min_out);
}

// ------------------------------------------------------------------------------------------
// K7: synthetic index, a pure function of (seed, row, GLOBAL column).  Same arithmetic as
// oracle/bigsi_oracle.c oracle_synth_row (the spec is in DESIGN.md "Synthetic index").
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t synth_row_key(uint64_t seed, uint64_t row)
{
    return mix64(seed ^ (row * 0xd1b54a32d192ed03ull));
}
__device__ __forceinline__ uint64_t synth_word(uint64_t rk, uint64_t W, int and_draws)
{
    uint64_t w = ~0ull;
    for (int i = 0; i < and_draws; ++i) w &= mix64(rk + (W * 4 + (uint64_t)i) * 0x9e3779b97f4a7c15ull);
    return w;
}
__device__ __forceinline__ uint32_t synth_plant_u32(uint64_t rk, uint64_t col)
{
    return (uint32_t)(mix64(rk ^ (col * 0xc2b2ae3d27d4eb4full + 0x165667b19e3779f9ull)) >> 32);
}

// one thread per (row, 8 local bytes); the whole pitch is written (padding = 0)
__global__ void __launch_bounds__(256) fill_synthetic_kernel(uint8_t *__restrict__ matrix, uint64_t pitch,
                                                             uint64_t num_rows, uint64_t num_cols, uint64_t col_offset,
                                                             uint64_t seed, int and_draws)
{
    const uint32_t wpr = (uint32_t)(pitch >> 3);  // 8-byte words per row
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t row = tid / wpr;
    const uint32_t j = (uint32_t)(tid % wpr);
    if (row >= num_rows) return;
    const uint64_t rk = synth_row_key(seed, row);
    const uint64_t gb = (col_offset >> 3) + (uint64_t)j * 8;  // global byte of local byte 8j
    const uint64_t W = gb >> 3;
    const uint32_t sh = (uint32_t)(gb & 7) * 8;
    uint64_t v = synth_word(rk, W, and_draws) >> sh;
    if (sh) v |= synth_word(rk, W + 1, and_draws) << (64 - sh);
    // zero the columns >= num_cols: local byte b holds columns 8b..8b+7, MSB first
    const uint64_t first_col = (uint64_t)j * 64;
    if (first_col >= num_cols) {
        v = 0;
    } else if (first_col + 64 > num_cols) {
        const uint32_t keep = (uint32_t)(num_cols - first_col);  // 1..63 valid columns
        const uint32_t full_bytes = keep >> 3, rem = keep & 7;
        uint64_t mask = full_bytes ? (~0ull >> (64 - 8 * full_bytes)) : 0ull;
        if (rem) mask |= (uint64_t)(0xff00u >> rem & 0xffu) << (8 * full_bytes);
        v &= mask;
    }
    *reinterpret_cast<uint64_t *>(matrix + row * pitch + (uint64_t)j * 8) = v;
}

// planted columns: one thread per row walks the (short) planted list, so there are no races
__global__ void __launch_bounds__(256) plant_columns_kernel(uint8_t *__restrict__ matrix, uint64_t pitch,
                                                            uint64_t num_rows, uint64_t num_cols, uint64_t col_offset,
                                                            uint64_t seed, const uint64_t *__restrict__ planted_cols,
                                                            const uint32_t *__restrict__ planted_thr, int n_planted)
{
    const uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= num_rows) return;
    const uint64_t rk = synth_row_key(seed, row);
    uint8_t *base = matrix + row * pitch;
    for (int p = 0; p < n_planted; ++p) {
        const uint64_t c = planted_cols[p];
        if (c < col_offset || c >= col_offset + num_cols) continue;
        const uint64_t lc = c - col_offset;
        const uint8_t mask = (uint8_t)(0x80u >> (lc & 7));
        const uint32_t thr = planted_thr[p];
        const bool bit = thr == 0xffffffffu ? true : synth_plant_u32(rk, c) < thr;
        const uint8_t v = base[lc >> 3];
        base[lc >> 3] = bit ? (uint8_t)(v | mask) : (uint8_t)(v & ~mask);
    }
}

cudaError_t launch_fill_synthetic(uint8_t *matrix, uint64_t pitch, uint64_t num_rows, uint64_t num_cols,
                                  uint64_t col_offset, uint64_t seed, int and_draws, const uint64_t *d_planted_cols,
                                  const uint32_t *d_planted_thr, int n_planted, cudaStream_t stream)
{
    if (num_rows == 0) return cudaSuccess;
    const uint64_t threads = num_rows * (pitch >> 3);
    const uint64_t blocks = (threads + 255) / 256;
    if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    fill_synthetic_kernel<<<(unsigned)blocks, 256, 0, stream>>>(matrix, pitch, num_rows, num_cols, col_offset, seed,
                                                                and_draws);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || n_planted == 0) return e;
    plant_columns_kernel<<<(unsigned)((num_rows + 255) / 256), 256, 0, stream>>>(
        matrix, pitch, num_rows, num_cols, col_offset, seed, d_planted_cols, d_planted_thr, n_planted);
    return cudaGetLastError();
}

}  // namespace bigsi
