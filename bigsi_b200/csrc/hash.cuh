// Canonical k-mer + MurmurHash3_x86_32 device code shared by the stand-alone hash kernel
// (aux_kernels.cu) and by the prologue of the fused query kernel (query_kernels.cu).
// Replaces convert_query_kmer/canonical (bigsi/utils/fncts.py:38-54) and _hash/generate_hashes
// (bigsi/bloom/bloomfilter.py:5-13; third-party mmh3 2.5.1 = MurmurHash3_x86_32): seeds 0..h-1,
// SIGNED 32-bit result, Python floor-mod m.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace bigsi {

__device__ __forceinline__ uint32_t comp_base(uint32_t b)
{
    // only A<->T and C<->G are complemented (utils/fncts.py:12); anything else passes through
    return b == 'A' ? 'T' : b == 'T' ? 'A' : b == 'C' ? 'G' : b == 'G' ? 'C' : b;
}
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return __funnelshift_l(x, x, r); }

__device__ __forceinline__ int32_t murmur_finish_mod(uint32_t h1, uint32_t len, uint32_t m)
{
    h1 ^= len;
    h1 ^= h1 >> 16;
    h1 *= 0x85ebca6bu;
    h1 ^= h1 >> 13;
    h1 *= 0xc2b2ae35u;
    h1 ^= h1 >> 16;
    // Python floor-mod of the SIGNED 32-bit hash (bloom/bloomfilter.py:5-6), in 32-bit arithmetic
    if ((int32_t)h1 >= 0) return (int32_t)(h1 % m);
    const uint32_t r = (0u - h1) % m;  // |s| mod m
    return (int32_t)(r ? m - r : 0u);
}
__device__ __forceinline__ uint32_t murmur_block(uint32_t h1, uint32_t k1)
{
    k1 *= 0xcc9e2d51u;
    k1 = rotl32(k1, 15);
    k1 *= 0x1b873593u;
    h1 ^= k1;
    h1 = rotl32(h1, 13);
    return h1 * 5u + 0xe6546b64u;
}
__device__ __forceinline__ uint32_t murmur_tail(uint32_t h1, uint32_t k1)
{
    k1 *= 0xcc9e2d51u;
    k1 = rotl32(k1, 15);
    k1 *= 0x1b873593u;
    return h1 ^ k1;
}

// shared-memory scratch hash_kmers_cooperative needs for cnt k-mers of length k
__host__ __device__ inline uint64_t hash_scratch_bytes(uint64_t cnt, uint32_t k)
{
    return cnt * ((uint64_t)k + 1 + 4ull * ((((uint64_t)k + 3) >> 2) | 1)) + 96;
}

// Lemire's fastmod: a mod m for 32-bit a with magic = 2^64 / m + 1 (precomputed on the host; 0 for m == 1)
__host__ __device__ inline uint64_t mod_magic(uint32_t m) { return m ? 0xffffffffffffffffull / m + 1ull : 0ull; }
__device__ __forceinline__ uint32_t fastmod_u32(uint32_t a, uint64_t magic, uint32_t m)
{
    return (uint32_t)__umul64hi(magic * (uint64_t)a, (uint64_t)m);
}
// fmix32 + Python floor-mod of the SIGNED hash, with the precomputed magic
__device__ __forceinline__ int32_t murmur_finish_fastmod(uint32_t h1, uint32_t len, uint32_t m, uint64_t magic)
{
    h1 ^= len;
    h1 ^= h1 >> 16;
    h1 *= 0x85ebca6bu;
    h1 ^= h1 >> 13;
    h1 *= 0xc2b2ae35u;
    h1 ^= h1 >> 16;
    if ((int32_t)h1 >= 0) return (int32_t)fastmod_u32(h1, magic, m);
    const uint32_t r = fastmod_u32(0u - h1, magic, m);  // |s| mod m
    return (int32_t)(r ? m - r : 0u);
}
// complement of four packed ASCII bases: A<->T, C<->G, anything else (and zero padding) unchanged
__device__ __forceinline__ uint32_t comp_base4(uint32_t w)
{
    const uint32_t at = __vcmpeq4(w, 0x41414141u) | __vcmpeq4(w, 0x54545454u);
    const uint32_t cg = __vcmpeq4(w, 0x43434343u) | __vcmpeq4(w, 0x47474747u);
    return w ^ (at & 0x15151515u) ^ (cg & 0x04040404u);
}

// One k-mer of length k <= 32 entirely in registers: s = its bytes in SHARED memory.  Forward and
// reverse-complement strings are built as eight little-endian words (zero padded), compared as
// big-endian integers (= lexicographic byte order, utils/fncts.py:38-39), and the smaller one goes
// through MurmurHash3_x86_32 for seeds 0..h-1 (the seed-independent k1 mixing is shared).
__device__ __forceinline__ void hash_one_kmer_regs(const uint8_t *s, int k, int h, uint32_t m, uint64_t magic, int canonical,
                                                   int32_t *ids)
{
    // The k bytes as eight little-endian words, zero padded: nine aligned 32-bit loads around s (the staging buffers
    // are padded, so the words before / behind the k-mer exist) and one funnel shift per word, instead of k byte
    // loads -- this code runs once per CTA on the critical path of the first bulk copy.
    uint32_t F[8], R[8];
    {
        const uint32_t addr = smem_u32(s);
        const uint32_t base = addr & ~3u, sh = (addr & 3u) * 8u;
        uint32_t W[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(W[j]) : "r"(base + 4u * j));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t w = __funnelshift_r(W[j], W[j + 1], sh);
            const int nv = k - 4 * j;  // valid bytes of this word
            F[j] = nv >= 4 ? w : nv <= 0 ? 0u : (w & ((1u << (8 * nv)) - 1u));
        }
    }
    // reverse string: byte i = s[k-1-i] = the fully reversed 32-byte block shifted down by 32-k bytes
    {
        uint32_t V[9];
#pragma unroll
        for (int j = 0; j < 8; ++j) V[j] = __byte_perm(F[7 - j], 0, 0x0123);
        V[8] = 0;
        const uint32_t d = 32u - (uint32_t)k;  // 0 .. 31 bytes
        if (d & 16u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) V[j] = j + 4 < 8 ? V[j + 4] : 0u;
        }
        if (d & 8u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) V[j] = j + 2 < 8 ? V[j + 2] : 0u;
        }
        if (d & 4u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) V[j] = j + 1 < 8 ? V[j + 1] : 0u;
        }
        const uint32_t db = (d & 3u) * 8u;
#pragma unroll
        for (int j = 0; j < 8; ++j) R[j] = __funnelshift_r(V[j], V[j + 1], db);
    }
    bool fwd = true;
    if (canonical) {
#pragma unroll
        for (int j = 7; j >= 0; --j) {  // the lowest differing word (earliest bytes) decides last
            R[j] = comp_base4(R[j]);
            const uint32_t fb = __byte_perm(F[j], 0, 0x0123), rb = __byte_perm(R[j], 0, 0x0123);
            if (fb != rb) fwd = fb < rb;
        }
    }
    const int nblocks = k >> 2, rem = k & 3;
#pragma unroll
    for (int j = 0; j < 8; ++j) {  // seed-independent part of every block / of the tail
        uint32_t k1 = fwd ? F[j] : R[j];
        k1 *= 0xcc9e2d51u;
        k1 = rotl32(k1, 15);
        F[j] = k1 * 0x1b873593u;
    }
    // four seeds at a time: the block chain of one seed is 8 dependent steps, four independent chains fill the pipeline
    for (int seed0 = 0; seed0 < h; seed0 += 4) {
        uint32_t h1[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) h1[u] = (uint32_t)(seed0 + u);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j < nblocks) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    h1[u] ^= F[j];
                    h1[u] = rotl32(h1[u], 13);
                    h1[u] = h1[u] * 5u + 0xe6546b64u;
                }
            } else if (j == nblocks && rem) {
#pragma unroll
                for (int u = 0; u < 4; ++u) h1[u] ^= F[j];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (seed0 + u < h) ids[seed0 + u] = murmur_finish_fastmod(h1[u], (uint32_t)k, m, magic);
    }
}

// "Low-latency" transfer of 16 data bytes between GPUs without a fence or a separate ready flag: two 16-byte
// stores {d0, flag, d1, flag}, {d2, flag, d3, flag}; 8-byte halves are written atomically, so the receiver
// re-reads a line until all four flags match the value it expects (the scheme of NCCL's LL protocol).
__device__ __forceinline__ void ll_store_line(uint4 *dst, const uint4 &v, uint32_t flag)
{
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(v.x), "r"(flag), "r"(v.y), "r"(flag) : "memory");
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst + 1), "r"(v.z), "r"(flag), "r"(v.w), "r"(flag) : "memory");
}
struct LlRoute;
__device__ __forceinline__ uint4 ll_load_line(const uint4 *src, const LlRoute &route);

// Route of a query's k-mer bytes in a column-sharded search: rank 0 stores every 16-byte line of the k-mer
// array into every peer's LL inbox (out[0..n_push)); a peer reads its k-mers from its own inbox `in`.
// (A scatter + relay route -- every shard forwarding one segment -- balances the NVLink ports but makes the
// peers wait for each other by one hop in EVERY query: measured 2.5 us per query on 4 GPUs; the star costs
// rank 0's port (world-1) x 2 x the query = 4.3 MB per 10 000-k-mer query on 8 GPUs, a few per cent of it.)
struct LlRoute {
    const uint4 *in;           // this shard's LL inbox (null on rank 0 and outside a sharded search)
    const uint8_t *kmers_base; // address line 0 corresponds to (16-byte aligned)
    uint32_t flag;             // low 32 bits of the query's sequence number (never 0)
    uint4 *out[8];             // rank 0: the peers' LL inboxes
    // the wait for a line is bounded (ptx.cuh:bounded_wait): a sender that never launches must not hang this GPU
    unsigned long long *abort_word, *host_abort;
    unsigned long long timeout_ns, seq;
};
// A line is re-read until all four flags match; gives up (abort word raised, data undefined) after timeout_ns.
__device__ __forceinline__ uint4 ll_load_line(const uint4 *src, const LlRoute &route)
{
    uint4 p, q;
    const uint32_t flag = route.flag;
    bounded_wait(route.abort_word, route.host_abort, route.timeout_ns, /*kAbortInbox*/ 3u, route.seq, [&]() {
        asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(p.x), "=r"(p.y), "=r"(p.z), "=r"(p.w) : "l"(src) : "memory");
        asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(src + 1) : "memory");
        return p.y == flag && p.w == flag && q.y == flag && q.w == flag;
    });
    return make_uint4(p.x, p.z, q.x, q.z);
}

// Synchronisation of a hashing group: the whole CTA (id 0), a subset of warps on a named barrier
// (id > 0, nthreads a multiple of 32), or one warp (nthreads == 32).  A runtime choice on purpose: the
// hashing code exists ONCE per kernel (it runs once per launch, so its cost is instruction fetch).
struct GroupSync {
    int id, nthreads;  // id 0 = the CTA barrier, any other id = hardware barrier 1 (ids are immediates: see ptx.cuh:named_bar_sync)
    __device__ __forceinline__ void operator()() const
    {
        if (nthreads == 32) __syncwarp();
        else if (id == 0) asm volatile("bar.sync 0, %0;" ::"r"(nthreads) : "memory");
        else asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
    }
};
typedef GroupSync SyncNamed;
__device__ __forceinline__ GroupSync SyncBlock() { return GroupSync{0, (int)blockDim.x}; }
__device__ __forceinline__ GroupSync SyncWarp() { return GroupSync{0, 32}; }

// Group-cooperative hashing of cnt CONTIGUOUS k-mers starting at g0: thread `tid` of a group of
// `nthreads` threads that synchronise with `sync` (every thread of the group must call; contains
// group barriers).  (1) the bytes are staged in shared memory with one round trip of 16-byte loads
// over the enclosing aligned window (the window leaves the k-mer array only inside its first / last
// 16-byte line, which every CUDA allocation covers); (2) one thread per k-mer decides the
// orientation; (3) one thread per (k-mer, 4-byte block) writes the canonical bytes as little-endian
// words (zero padded, so the last word IS murmur's tail); (4) one thread per (k-mer, seed) runs
// MurmurHash3 over those words and stores ids[km * h + seed].  `scratch` (16-byte aligned,
// hash_scratch_bytes(cnt, k) bytes) and `ids` may be shared or global memory.  magic = mod_magic(m).
//
// route->in != nullptr: the k-mer bytes are not read from g0 but from a "low-latency" inbox other GPUs write
// over NVLink (LlRoute): 16-byte data line j of the k-mer array that starts at kmers_base (16-byte aligned)
// lives in the line pair in[2j], in[2j+1]; every 8-byte half carries 4 data bytes and the 32-bit flag of the
// query, so a line is simply re-read until all its flags match -- no separate ready flag, no fence on the
// sender's side, and the wait is per 16-byte line.
static __device__ __noinline__ void hash_kmers_group(const uint8_t *g0, uint32_t cnt, int k, int h, uint32_t m, int canonical,
                                             uint8_t *scratch, int32_t *ids, uint32_t tid, uint32_t nthreads,
                                             const GroupSync sync, uint64_t magic, const LlRoute *route = nullptr)
{
    const int nblocks = k >> 2, rem = k & 3;
    const uint32_t wpk = (uint32_t)(k + 3) >> 2;
    const uint32_t wstride = wpk | 1;  // odd stride: conflict-free LDS across k-mers
    const uint32_t raw_bytes = ((cnt * (uint32_t)k + 32) + 15) & ~15u;
    uint8_t *fwd = scratch + raw_bytes;
    uint32_t *cw = reinterpret_cast<uint32_t *>(scratch + raw_bytes + ((cnt + 15) & ~15u));
    const uint32_t nbytes = cnt * (uint32_t)k;
    const uint32_t skew = (uint32_t)(reinterpret_cast<uintptr_t>(g0) & 15);
    const uint4 *a0 = reinterpret_cast<const uint4 *>(g0 - skew);
    const uint32_t nvec = (skew + nbytes + 15) >> 4;
    uint4 *sv = reinterpret_cast<uint4 *>(scratch);
    if (route == nullptr || route->in == nullptr) {
        for (uint32_t i = tid; i < nvec; i += nthreads) sv[i] = __ldg(a0 + i);
    } else {
        const uint64_t line0 = (uint64_t)(reinterpret_cast<const uint8_t *>(a0) - route->kmers_base) >> 4;
        for (uint32_t i = tid; i < nvec; i += nthreads) {
            const uint64_t line = line0 + i;
            sv[i] = ll_load_line(route->in + 2 * line, *route);
        }
    }
    sync();
    const uint8_t *src = scratch + skew;
    if (k <= 32) {  // the common case (k = 31): one thread per k-mer, no further barriers
        for (uint32_t km = tid; km < cnt; km += nthreads)
            hash_one_kmer_regs(src + (size_t)km * k, k, h, m, magic, canonical, ids + (size_t)km * h);
        return;
    }
    for (uint32_t km = tid; km < cnt; km += nthreads) {
        const uint8_t *s = src + (size_t)km * k;
        bool f = true;  // forward unless the reverse complement is lexicographically smaller
        if (canonical) {
            for (int j = 0; j < k; ++j) {
                const uint32_t a = s[j], b = comp_base(s[k - 1 - j]);
                if (a != b) {
                    f = a < b;
                    break;
                }
            }
        }
        fwd[km] = f ? 1 : 0;
    }
    sync();
    for (uint32_t i = tid; i < cnt * wpk; i += nthreads) {
        const uint32_t km = i / wpk, wi = i % wpk;
        const uint8_t *s = src + (size_t)km * k;
        const bool f = fwd[km] != 0;
        uint32_t word = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = (int)wi * 4 + u;
            if (j < k) word |= (f ? (uint32_t)s[j] : comp_base(s[k - 1 - j])) << (8 * u);
        }
        cw[km * wstride + wi] = word;
    }
    sync();
    for (uint32_t w = tid; w < cnt * (uint32_t)h; w += nthreads) {
        const uint32_t km = w / (uint32_t)h, seed = w % (uint32_t)h;
        const uint32_t *wp = cw + km * wstride;
        uint32_t h1 = seed;
        for (int b = 0; b < nblocks; ++b) h1 = murmur_block(h1, wp[b]);
        if (rem) h1 = murmur_tail(h1, wp[nblocks]);
        ids[w] = murmur_finish_mod(h1, (uint32_t)k, m);
    }
}

// ------------------------------------------------------------------------------------------------
// Sequence front-end inside the query kernel (k <= 32): BIGSI.search's `set(kmers)` over RAW window strings
// (bigsi/utils/fncts.py:63-65, graph/index.py:45, graph/bigsi.py:177-179) without a separate kernel.  Every CTA
// stages its span of the sequence, every window is inserted into a global open-addressing table with atomicCAS;
// the window that wins its entry represents its string (any representative will do for a set), equal tags are
// confirmed by comparing the k bytes, so the set is exact.  Entries carry a 16-bit epoch: entries of other epochs
// count as empty, the table is never cleared between queries.
//   entry = epoch (16) | tag (16) | window index (32)
// ------------------------------------------------------------------------------------------------
struct SeqTable {
    unsigned long long *entries;
    uint64_t mask;        // entries - 1 (a power of two minus one)
    uint32_t epoch;       // 1 .. 65535
    const uint8_t *seq;   // the whole sequence (device-addressable; may be mapped host memory)
};

// 8 bytes of a window held in shared memory (unaligned), zero padded past k
__device__ __forceinline__ unsigned long long window_word(const uint8_t *s, int k, int j)
{
    unsigned long long v = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b)
        if (j * 8 + b < k) v |= (unsigned long long)s[j * 8 + b] << (8 * b);
    return v;
}
__device__ __forceinline__ unsigned long long window_fingerprint(const uint8_t *s, int k)
{
    unsigned long long fp = 0x9e3779b97f4a7c15ull ^ (unsigned long long)k;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j * 8 < k) {
            fp = (fp ^ window_word(s, k, j)) * 0xff51afd7ed558ccdull;
            fp ^= fp >> 29;
        }
    }
    fp *= 0xc4ceb9fe1a85ec53ull;
    return fp ^ (fp >> 32);
}
// window `s` (shared memory) == window `other` of the sequence?  The other window is read from the CTA's own span
// when it lies inside it, else from the sequence in global memory as aligned 8-byte words (one round trip).
__device__ __forceinline__ bool window_equals(const SeqTable &T, const uint8_t *s, int k, uint64_t other, const uint8_t *span,
                                              uint64_t span_w0, uint32_t span_cnt)
{
    if (other >= span_w0 && other < span_w0 + span_cnt) {
        const uint8_t *o = span + (other - span_w0);
        bool eq = true;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j * 8 < k) eq = eq && window_word(s, k, j) == window_word(o, k, j);
        return eq;
    }
    const uint8_t *g = T.seq + other;
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(g) & 7) * 8;
    const unsigned long long *a = reinterpret_cast<const unsigned long long *>(g - (sh >> 3));
    unsigned long long w[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) w[j] = (j * 8 < k + 8) ? __ldg(a + j) : 0ull;  // the staging buffer is padded: never out of bounds
    bool eq = true;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j * 8 < k) {
            unsigned long long v = sh ? (w[j] >> sh) | (w[j + 1] << (64 - sh)) : w[j];
            const int rem = k - j * 8;
            if (rem < 8) v &= (1ull << (8 * rem)) - 1ull;
            eq = eq && v == window_word(s, k, j);
        }
    }
    return eq;
}
// true: this window is the representative of its string (first to claim the entry)
__device__ __forceinline__ bool seq_table_insert(const SeqTable &T, const uint8_t *s, int k, uint64_t widx, const uint8_t *span,
                                                 uint64_t span_w0, uint32_t span_cnt)
{
    const unsigned long long fp = window_fingerprint(s, k);
    const unsigned long long tag = (fp >> 44) & 0xffffull;
    const unsigned long long mine = ((unsigned long long)T.epoch << 48) | (tag << 32) | (unsigned long long)(uint32_t)widx;
    uint64_t slot = fp & T.mask;
    for (;;) {
        unsigned long long cur = ld_volatile_u64(T.entries + slot);
        if ((cur >> 48) != T.epoch) {
            const unsigned long long old = atomicCAS(T.entries + slot, cur, mine);
            if (old == cur) return true;
            cur = old;
            if ((cur >> 48) != T.epoch) continue;  // (cannot happen: only this query writes this table) look again
        }
        if (((cur >> 32) & 0xffffull) == tag && window_equals(T, s, k, cur & 0xffffffffull, span, span_w0, span_cnt)) return false;
        slot = (slot + 1) & T.mask;
    }
}

// Hash `cnt` windows of a staged span: window j starts at span + list[j]; one thread per window (k <= 32).
__device__ __forceinline__ void hash_window_list(const uint8_t *span, const uint16_t *list, uint32_t cnt, int k, int h, uint32_t m,
                                                 uint64_t magic, int32_t *ids, uint32_t tid, uint32_t nthreads)
{
    for (uint32_t j = tid; j < cnt; j += nthreads) hash_one_kmer_regs(span + list[j], k, h, m, magic, 1, ids + (size_t)j * h);
}

// the whole CTA as one group
__device__ __forceinline__ void hash_kmers_cooperative(const uint8_t *g0, uint32_t cnt, int k, int h, uint32_t m,
                                                       int canonical, uint8_t *scratch, int32_t *ids, uint64_t magic)
{
    hash_kmers_group(g0, cnt, k, h, m, canonical, scratch, ids, threadIdx.x, blockDim.x, SyncBlock(), magic);
}

}  // namespace bigsi
