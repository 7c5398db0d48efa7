// Stage 2 of a query launch: merge the per-segment partial planes written by stage 1.
// Device functions shared by the stand-alone merge kernels (merge_kernels.cu) and by the merge
// phase that fused_query runs itself after a grid-wide barrier (query_kernels.cu).
//
// COUNTS.  For one (query, tile) the partial buffer holds S slots (one per slice the query spans)
// of `pps` bit planes each: slot s, plane b, bit c  =  bit b of the number of k-mers of segment s
// whose AND vector has column c set.  The count of column c is  sum_b 2^b * (number of slots whose
// plane b has bit c set), so every plane is an independent VERTICAL POPCOUNT over the S slots --
// the same Harley-Seal carry-save counting stage 1 does over k-mers.  Work item = (query, tile,
// chunk of `gpi` word groups); a word group is 32/NG consecutive 32-bit words.  Thread layout:
// lane = (word w of the group, slot group g) with 32/NG words x NG slot groups; warp = (plane b,
// word group) pair.  Each thread counts its slots for its (word, plane) from double-buffered
// batches of 16 independent loads, the NG slot groups are added with warp shuffles (bit-sliced
// full adders), the per-plane counters go to shared memory and every thread then expands whole
// columns: count = sum_{b,j} bit(cnt[b][j]) << (b + j).  The threshold (graph/bigsi.py:241-242)
// is applied while the count is in a register.
//
// AND.  One plane per slot; AND over the slots (graph/bigsi.py:192-195).
#pragma once
#include "ptx.cuh"
#include "query.cuh"

namespace bigsi {

constexpr int kCntPlanes = 16;        // counter planes per (word, plane): up to 65 535 slots per (tile, query)
constexpr int kMergeMaxWarps = 16;
constexpr int kMergeMaxItemWords = 32;  // words (of 32 columns) one work item covers at most
// shared memory one merge item needs: [plane b][counter plane j][word] + compaction scratch
constexpr int kMergeSmemBytes = kSegPlanes * kCntPlanes * kMergeMaxItemWords * 4 + 256;

struct MergeGeom {
    uint32_t q, t, tb0, tw, cb;
    uint64_t s_first, n_slots;  // slices spanned by (tile, query); n_slots == 0 when the query is empty
};

// item -> (query, tile, chunk); false when the chunk lies past the end of a narrow last tile
__device__ __forceinline__ bool merge_geometry(const QueryParams &P, uint64_t item, uint32_t words_per_item, MergeGeom &g)
{
    const uint32_t chunk_bytes = words_per_item * 4;
    const uint32_t cpt = (P.tile_bytes + chunk_bytes - 1) / chunk_bytes;
    const uint32_t chunk = (uint32_t)(item % cpt);
    const uint64_t tq = item / cpt;
    g.t = (uint32_t)(tq % P.n_tiles);
    g.q = (uint32_t)(tq / P.n_tiles);
    g.tb0 = g.t * P.tile_bytes;
    g.tw = min(P.tile_bytes, P.row_bytes16 - g.tb0);
    g.cb = chunk * chunk_bytes;
    if (g.cb >= g.tw) return false;
    // a single query spans [0, total_kmers) by contract: no dependent load in front of the planes
    const bool one = P.n_queries == 1;
    const uint64_t k0 = one ? 0ull : (uint64_t)__ldg(P.qoff + g.q);
    const uint64_t k1 = one ? P.total_kmers : (uint64_t)__ldg(P.qoff + g.q + 1);
    if (k1 > k0) {
        const uint64_t I0 = (uint64_t)g.t * P.total_kmers + k0, I1 = (uint64_t)g.t * P.total_kmers + k1;
        g.s_first = I0 / P.items_per_slice;
        g.n_slots = (I1 - 1) / P.items_per_slice - g.s_first + 1;
    } else {
        g.s_first = 0;
        g.n_slots = 0;
    }
    return true;
}
inline uint64_t merge_item_count(const QueryParams &p, uint32_t words_per_item)
{
    const uint64_t cpt = (p.tile_bytes + words_per_item * 4 - 1) / (words_per_item * 4);
    return cpt * p.n_tiles * p.n_queries;
}

// full-adder step of a bit-sliced add: acc += x (one plane), carry chained by the caller
__device__ __forceinline__ void fa(uint32_t &acc, uint32_t x, uint32_t &carry)
{
    const uint32_t o = acc;
    acc = xor3(o, x, carry);
    carry = maj3(o, x, carry);
}

// One COUNTS work item, executed by the whole CTA (blockDim.x threads, a multiple of 32, at most
// kMergeMaxWarps warps).  gpi = word groups per item; gpi * (32/NG) <= kMergeMaxItemWords.
// All threads must call it (it contains __syncthreads); smem = kMergeSmemBytes of scratch.
template <int NG>
__device__ __forceinline__ void merge_counts_item(const QueryParams &P, uint64_t item, uint32_t gpi, uint8_t *smem)
{
    constexpr int WPG = 32 / NG;  // words per word group
    uint32_t *sm = reinterpret_cast<uint32_t *>(smem);
    uint32_t *warp_hits = sm + kSegPlanes * kCntPlanes * kMergeMaxItemWords;
    unsigned long long *hit_base = reinterpret_cast<unsigned long long *>(warp_hits + kMergeMaxWarps + 2);
    const uint32_t wpi = WPG * gpi;  // words per item

    MergeGeom G;
    const bool in_range = merge_geometry(P, item, wpi, G);  // block-uniform
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t pps = P.planes_per_slot;
    const uint32_t J = in_range ? 32 - __clz((uint32_t)G.n_slots) : 0;  // counter planes needed: bits(S)
    const uint64_t slot_stride = (uint64_t)pps * P.tile_bytes;

    if (in_range) {
        // warp-level work units: (plane b, word group wg), pps * gpi of them
        for (uint32_t unit = warp; unit < pps * gpi; unit += nwarps) {
            const uint32_t b = unit % pps, wg = unit / pps;
            const uint32_t w = wg * WPG + lane % WPG, g = lane / WPG;
            const bool valid = G.cb + w * 4 < G.tw;
            uint32_t c[kCntPlanes];
#pragma unroll
            for (int j = 0; j < kCntPlanes; ++j) c[j] = 0;
            if (valid && g < G.n_slots) {
                const uint32_t n_mine = (uint32_t)((G.n_slots - g + NG - 1) / NG);  // slots s_first + g, + NG, ...
                const uint64_t step = (uint64_t)NG * slot_stride;
                const uint8_t *src = P.partial + (G.s_first + g + (uint64_t)G.t * P.n_queries + G.q) * slot_stride +
                                     (uint64_t)b * P.tile_bytes + G.cb + w * 4;
                const uint32_t nhi = J > 3 ? J - 3 : 0;
                // Batches of kB independent loads, double buffered: while one batch is counted the
                // next is in flight (clamped index + select keeps the loads branch-free).  Plain
                // weak loads: the planes were published before the kernel boundary / grid barrier.
                constexpr int kB = 16;
                uint32_t xn[kB];
                auto load_batch = [&](uint32_t i0) {
#pragma unroll
                    for (int u = 0; u < kB; ++u)
                        xn[u] = *reinterpret_cast<const uint32_t *>(src + (uint64_t)min(i0 + u, n_mine - 1) * step);
                };
                load_batch(0);
                for (uint32_t i = 0; i < n_mine; i += kB) {
                    uint32_t x[kB];
#pragma unroll
                    for (int u = 0; u < kB; ++u) x[u] = i + u < n_mine ? xn[u] : 0u;
                    if (i + kB < n_mine) load_batch(i + kB);
#pragma unroll
                    for (int v = 0; v < kB; v += 8) {
                        // Harley-Seal block: 8 inputs of weight 1 -> ones/twos/fours + a carry of weight 8
                        uint32_t t0 = maj3(c[0], x[v + 0], x[v + 1]);
                        c[0] = xor3(c[0], x[v + 0], x[v + 1]);
                        uint32_t t1 = maj3(c[0], x[v + 2], x[v + 3]);
                        c[0] = xor3(c[0], x[v + 2], x[v + 3]);
                        const uint32_t f0 = maj3(c[1], t0, t1);
                        c[1] = xor3(c[1], t0, t1);
                        t0 = maj3(c[0], x[v + 4], x[v + 5]);
                        c[0] = xor3(c[0], x[v + 4], x[v + 5]);
                        t1 = maj3(c[0], x[v + 6], x[v + 7]);
                        c[0] = xor3(c[0], x[v + 6], x[v + 7]);
                        const uint32_t f1 = maj3(c[1], t0, t1);
                        c[1] = xor3(c[1], t0, t1);
                        uint32_t carry = maj3(c[2], f0, f1);
                        c[2] = xor3(c[2], f0, f1);
#pragma unroll
                        for (int j = 3; j < kCntPlanes; ++j) {
                            if (j - 3 < (int)nhi) {
                                const uint32_t o = c[j];
                                c[j] = o ^ carry;
                                carry = o & carry;
                            }
                        }
                    }
                }
            }
            // add the NG slot groups: lanes l and l ^ (WPG * 2^k) hold the same word
            if (NG > 1) {
#pragma unroll
                for (int d = WPG; d < 32; d <<= 1) {
                    uint32_t carry = 0;
#pragma unroll
                    for (int j = 0; j < kCntPlanes; ++j) {
                        const uint32_t o = __shfl_xor_sync(0xffffffffu, c[j], d);
                        fa(c[j], o, carry);
                    }
                }
            }
            if (g == 0) {
#pragma unroll
                for (int j = 0; j < kCntPlanes; ++j)
                    if (j < (int)J) sm[(b * kCntPlanes + j) * kMergeMaxItemWords + w] = c[j];
            }
        }
    }
    __syncthreads();

    if (in_range) {
        // expansion: one column per thread and pass
        uint32_t *out = P.out ? reinterpret_cast<uint32_t *>(P.out) + (uint64_t)G.q * P.out_stride : nullptr;
        const bool thresholding = P.min_kmers != nullptr || P.min_by_value;
        const uint32_t thr = P.min_by_value ? P.min_kmers_value : (thresholding ? __ldg(P.min_kmers + G.q) : 0u);
        const uint32_t col_base = (G.tb0 + G.cb) * 8;
        const uint32_t ncols_here = min(wpi * 32u, (G.tw - G.cb) * 8u);  // never past this tile
        for (uint32_t c0 = 0; c0 < ncols_here; c0 += blockDim.x) {
            const uint32_t cc = c0 + threadIdx.x;
            const uint32_t col = col_base + cc;
            const bool live = cc < ncols_here && col < P.num_cols;
            uint32_t cnt = 0;
            if (live) {
                // bit i of a little-endian 32-bit word of MSB-first bytes is column (i ^ 7) of that word
                const uint32_t word = cc >> 5, bit = (cc & 31) ^ 7;
                for (uint32_t b = 0; b < pps; ++b) {
                    const uint32_t *row = sm + (b * kCntPlanes) * kMergeMaxItemWords + word;
#pragma unroll
                    for (int j0 = 0; j0 < kCntPlanes; j0 += 4) {  // four independent LDS per step
                        if (j0 < (int)J) {
                            uint32_t v[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) v[u] = j0 + u < (int)J ? row[(j0 + u) * kMergeMaxItemWords] : 0u;
#pragma unroll
                            for (int u = 0; u < 4; ++u) cnt += ((v[u] >> bit) & 1u) << (b + j0 + u);
                        }
                    }
                }
                if (out) out[col] = cnt;
            }
            if (thresholding) {  // counts >= min_kmers (graph/bigsi.py:241-242), block-level compaction
                const bool hit = live && cnt >= thr;
                const uint32_t ballot = __ballot_sync(0xffffffffu, hit);
                if (lane == 0) warp_hits[warp] = __popc(ballot);
                __syncthreads();
                if (threadIdx.x == 0) {
                    uint32_t tot = 0;
                    for (uint32_t i = 0; i < nwarps; ++i) {
                        const uint32_t v = warp_hits[i];
                        warp_hits[i] = tot;
                        tot += v;
                    }
                    *hit_base = tot ? atomicAdd(P.n_hits + G.q, (unsigned long long)tot) : 0ull;
                }
                __syncthreads();
                if (hit) {
                    const uint64_t pos = *hit_base + warp_hits[warp] + __popc(ballot & ((1u << lane) - 1));
                    if (pos < P.hit_cap) {
                        P.hit_cols[(uint64_t)G.q * P.hit_cap + pos] = (int32_t)col;
                        P.hit_counts[(uint64_t)G.q * P.hit_cap + pos] = cnt;
                    }
                }
                __syncthreads();
            }
        }
    }
    __syncthreads();  // smem is reused by the next item
}

// One AND work item: lanes = 8 words x 4 slot groups, every warp its own 8 words.
constexpr int kAndNG = 4, kAndWPW = 32 / kAndNG;
__device__ __forceinline__ void merge_and_item(const QueryParams &P, uint64_t item, uint8_t *smem)
{
    uint32_t *sm = reinterpret_cast<uint32_t *>(smem);
    const uint32_t nwarps = blockDim.x >> 5;
    const uint32_t wpi = kAndWPW * nwarps;  // words per item
    MergeGeom G;
    const bool in_range = merge_geometry(P, item, wpi, G);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t w = warp * kAndWPW + lane % kAndWPW, g = lane / kAndWPW;
    if (in_range) {
        const bool valid = G.cb + w * 4 < G.tw;
        const uint64_t slot_stride = (uint64_t)P.planes_per_slot * P.tile_bytes;
        uint32_t acc = 0xffffffffu;
        if (valid && g < G.n_slots) {
            const uint32_t n_mine = (uint32_t)((G.n_slots - g + kAndNG - 1) / kAndNG);
            const uint64_t step = (uint64_t)kAndNG * slot_stride;
            const uint8_t *src =
                P.partial + (G.s_first + g + (uint64_t)G.t * P.n_queries + G.q) * slot_stride + G.cb + w * 4;
            for (uint32_t i = 0; i < n_mine; i += 16) {
                uint32_t x[16];
#pragma unroll
                for (int u = 0; u < 16; ++u)
                    x[u] = *reinterpret_cast<const uint32_t *>(src + (uint64_t)min(i + u, n_mine - 1) * step);
#pragma unroll
                for (int u = 0; u < 16; ++u) acc &= x[u];  // the clamped duplicates are harmless under AND
            }
        }
#pragma unroll
        for (int d = kAndWPW; d < 32; d <<= 1) acc &= __shfl_xor_sync(0xffffffffu, acc, d);
        if (g == 0) sm[w] = valid ? acc : 0u;
    }
    __syncthreads();
    if (in_range) {
        uint8_t *out = reinterpret_cast<uint8_t *>(P.out) + (uint64_t)G.q * P.out_stride;
        const uint32_t row_bytes = (P.num_cols + 7) >> 3;
        const uint32_t nbytes_here = min(wpi * 4u, G.tw - G.cb);  // never past this tile
        for (uint32_t c = threadIdx.x; c < nbytes_here; c += blockDim.x) {
            const uint32_t byte = G.tb0 + G.cb + c;
            if (byte >= row_bytes) break;
            uint32_t v = (sm[c >> 2] >> (8 * (c & 3))) & 0xffu;
            if (byte == row_bytes - 1 && (P.num_cols & 7)) v &= 0xff00u >> (P.num_cols & 7);
            out[byte] = (uint8_t)v;
        }
    }
    __syncthreads();
}

// launch geometry of the merge work (shared by both launch styles)
struct MergePlan {
    int ng;             // slot groups per warp (1 or 4)
    uint32_t gpi;       // word groups per item (COUNTS)
    uint32_t wpi;       // words per item
    uint64_t n_items;
};
inline MergePlan plan_merge(const QueryParams &p, int mode, uint32_t block_threads)
{
    MergePlan m;
    const uint32_t nwarps = block_threads / 32;
    if (mode == kModeAnd) {
        m.ng = kAndNG;
        m.gpi = nwarps;
        m.wpi = kAndWPW * nwarps;
    } else {
        // slots one (tile, query) can span decide how many lanes share a word's slot loop
        const uint64_t max_slots = p.max_query_kmers / p.items_per_slice + 2;
        m.ng = max_slots <= 4 ? 1 : 4;
        const uint32_t wpg = 32 / m.ng;
        uint32_t gpi = nwarps / (p.planes_per_slot ? p.planes_per_slot : 1);
        if (gpi < 1) gpi = 1;
        if (gpi * wpg > (uint32_t)kMergeMaxItemWords) gpi = kMergeMaxItemWords / wpg;
        m.gpi = gpi;
        m.wpi = gpi * wpg;
    }
    m.n_items = merge_item_count(p, m.wpi);
    return m;
}

}  // namespace bigsi
