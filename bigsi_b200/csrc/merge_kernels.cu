// Stage 2 of a query launch: merge the per-segment partial planes written by fused_query.
//
// COUNTS.  For one (query, tile) the partial buffer holds S slots (one per slice the query spans)
// of `pps` bit planes each: slot s, plane b, bit c  =  bit b of the number of k-mers of segment s
// whose AND vector has column c set.  The count of column c is  sum_b 2^b * (number of slots whose
// plane b has bit c set), so every plane is an independent VERTICAL POPCOUNT over the S slots --
// the same Harley-Seal carry-save counting the fused kernel does over k-mers.  Thread layout of a
// block of 256: lane = (word w, slot group g) with 32/NG words x NG groups, warp = plane lane
// (8 planes per pass).  Each thread counts its slots for its (word, plane) from batches of 8
// independent loads, the NG groups are added with warp shuffles (bit-sliced full adders), the
// per-plane counters go to shared memory and every thread then expands whole columns:
// count = sum_{b,j} bit(cnt[b][j]) << (b + j).  No ripple over all planes per slot, no tree of
// __syncthreads, and the threshold (graph/bigsi.py:241-242) is applied while the count is in a
// register.
//
// AND.  One plane per slot; AND over the slots (graph/bigsi.py:192-195).
#include "launch.cuh"
#include "ptx.cuh"
#include "query.cuh"

namespace bigsi {

constexpr int kMergeThreads = 256;
constexpr int kMergeWarps = kMergeThreads / 32;
constexpr int kCntPlanes = 16;  // counter planes per (word, plane): up to 65 535 slots per (tile, query)

struct MergeGeom {
    uint32_t q, t, tb0, tw, cb;
    uint64_t s_first, s_last;  // slices spanned by (tile, query); s_first > s_last when the query is empty
    bool empty;
};

__device__ __forceinline__ bool merge_geometry(const QueryParams &P, uint32_t words_per_block, MergeGeom &g)
{
    const uint32_t chunk_bytes = words_per_block * 4;
    const uint32_t cpt = (P.tile_bytes + chunk_bytes - 1) / chunk_bytes;
    const uint64_t bid = blockIdx.x;
    const uint32_t chunk = (uint32_t)(bid % cpt);
    const uint64_t tq = bid / cpt;
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
    g.empty = k1 <= k0;
    if (!g.empty) {
        const uint64_t I0 = (uint64_t)g.t * P.total_kmers + k0, I1 = (uint64_t)g.t * P.total_kmers + k1;
        g.s_first = I0 / P.items_per_slice;
        g.s_last = (I1 - 1) / P.items_per_slice;
    } else {
        g.s_first = 1;
        g.s_last = 0;
    }
    return true;
}

// full-adder step of a bit-sliced add: acc += x (one plane), carry chained by the caller
__device__ __forceinline__ void fa(uint32_t &acc, uint32_t x, uint32_t &carry)
{
    const uint32_t o = acc;
    acc = xor3(o, x, carry);
    carry = maj3(o, x, carry);
}

template <int NG>
__global__ void __launch_bounds__(kMergeThreads) merge_counts_kernel(const __grid_constant__ QueryParams P)
{
    constexpr int WPB = 32 / NG;  // words per block
    __shared__ uint32_t sm[kSegPlanes * kCntPlanes * WPB];  // [plane b][counter plane j][word]
    __shared__ uint32_t warp_hits[kMergeWarps];
    __shared__ unsigned long long hit_base;

    grid_dependency_wait();  // stage 1 must have completed (PDL launch)
    MergeGeom G;
    if (!merge_geometry(P, WPB, G)) return;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t w = lane % WPB, g = lane / WPB;
    const bool valid = G.cb + w * 4 < G.tw;
    const uint32_t pps = P.planes_per_slot;
    const uint64_t n_slots = G.empty ? 0 : G.s_last - G.s_first + 1;
    const uint32_t J = 32 - __clz((uint32_t)n_slots);  // counter planes needed: bits(S)
    const uint64_t slot_stride = (uint64_t)pps * P.tile_bytes;

    for (uint32_t b = warp; b < pps; b += kMergeWarps) {
        uint32_t c[kCntPlanes];
#pragma unroll
        for (int j = 0; j < kCntPlanes; ++j) c[j] = 0;
        if (valid && g < n_slots) {
            const uint32_t n_mine = (uint32_t)((n_slots - g + NG - 1) / NG);  // my slots: s_first + g, + NG, ...
            const uint64_t step = (uint64_t)NG * slot_stride;
            const uint8_t *src = P.partial + (G.s_first + g + (uint64_t)G.t * P.n_queries + G.q) * slot_stride +
                                 (uint64_t)b * P.tile_bytes + G.cb + w * 4;
            const uint32_t nhi = J > 3 ? J - 3 : 0;
            // Batches of kB independent loads, double buffered: while one batch is counted the next
            // is in flight (clamped index + select keeps the loads branch-free; the planes were
            // written by the previous kernel, so the read-only path is legal).
            constexpr int kB = 16;
            uint32_t xn[kB];
            auto load_batch = [&](uint32_t i0) {
#pragma unroll
                for (int u = 0; u < kB; ++u)
                    xn[u] = __ldg(reinterpret_cast<const uint32_t *>(src + (uint64_t)min(i0 + u, n_mine - 1) * step));
            };
            load_batch(0);
            for (uint32_t i = 0; i < n_mine; i += kB) {
                uint32_t x[kB];
#pragma unroll
                for (int u = 0; u < kB; ++u) x[u] = i + u < n_mine ? xn[u] : 0u;
                if (i + kB < n_mine) load_batch(i + kB);
#pragma unroll
                for (int v = 0; v < kB; v += 8) {
                    // Harley-Seal block: 8 inputs of weight 1 -> ones/twos/fours + one carry of weight 8
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
        // add the NG slot groups: lanes l and l ^ (WPB * 2^k) hold the same word
        if (NG > 1) {
#pragma unroll
            for (int d = WPB; d < 32; d <<= 1) {
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
                if (j < (int)J) sm[(b * kCntPlanes + j) * WPB + w] = c[j];
        }
    }
    __syncthreads();

    // expansion: one column per thread and pass
    uint32_t *out = P.out ? reinterpret_cast<uint32_t *>(P.out) + (uint64_t)G.q * P.out_stride : nullptr;
    const bool thresholding = P.min_kmers != nullptr;
    const uint32_t thr = thresholding ? __ldg(P.min_kmers + G.q) : 0u;
    const uint32_t col_base = (G.tb0 + G.cb) * 8;
    const uint32_t ncols_here = min((uint32_t)WPB * 32u, (G.tw - G.cb) * 8u);  // never past this tile
    for (uint32_t c0 = 0; c0 < ncols_here; c0 += kMergeThreads) {
        const uint32_t cc = c0 + threadIdx.x;
        const uint32_t col = col_base + cc;
        const bool live = cc < ncols_here && col < P.num_cols;
        uint32_t cnt = 0;
        if (live) {
            // bit i of a little-endian 32-bit word of MSB-first bytes is column (i ^ 7) of that word
            const uint32_t word = cc >> 5, bit = (cc & 31) ^ 7;
            for (uint32_t b = 0; b < pps; ++b) {
                const uint32_t *row = sm + (b * kCntPlanes) * WPB + word;
#pragma unroll
                for (int j0 = 0; j0 < kCntPlanes; j0 += 4) {  // four independent LDS per step
                    if (j0 < (int)J) {
                        uint32_t v[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) v[u] = j0 + u < (int)J ? row[(j0 + u) * WPB] : 0u;
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
                for (int i = 0; i < kMergeWarps; ++i) {
                    const uint32_t v = warp_hits[i];
                    warp_hits[i] = tot;
                    tot += v;
                }
                hit_base = tot ? atomicAdd(P.n_hits + G.q, (unsigned long long)tot) : 0ull;
            }
            __syncthreads();
            if (hit) {
                const uint64_t pos = hit_base + warp_hits[warp] + __popc(ballot & ((1u << lane) - 1));
                if (pos < P.hit_cap) {
                    P.hit_cols[(uint64_t)G.q * P.hit_cap + pos] = (int32_t)col;
                    P.hit_counts[(uint64_t)G.q * P.hit_cap + pos] = cnt;
                }
            }
            __syncthreads();
        }
    }
}

// AND mode: lanes = 8 words x 4 slot groups, 8 warps -> 64 words (256 bytes) per block
__global__ void __launch_bounds__(kMergeThreads) merge_and_kernel(const __grid_constant__ QueryParams P)
{
    constexpr int NG = 4, WPW = 32 / NG, WPB = WPW * kMergeWarps;
    __shared__ uint32_t sm[WPB];
    grid_dependency_wait();  // stage 1 must have completed (PDL launch)
    MergeGeom G;
    if (!merge_geometry(P, WPB, G)) return;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t w = warp * WPW + lane % WPW, g = lane / WPW;
    const bool valid = G.cb + w * 4 < G.tw;
    const uint64_t n_slots = G.empty ? 0 : G.s_last - G.s_first + 1;
    const uint64_t slot_stride = (uint64_t)P.planes_per_slot * P.tile_bytes;
    uint32_t acc = 0xffffffffu;
    if (valid && g < n_slots) {
        const uint32_t n_mine = (uint32_t)((n_slots - g + NG - 1) / NG);
        const uint64_t step = (uint64_t)NG * slot_stride;
        const uint8_t *src = P.partial + (G.s_first + g + (uint64_t)G.t * P.n_queries + G.q) * slot_stride + G.cb + w * 4;
        for (uint32_t i = 0; i < n_mine; i += 8) {
            uint32_t x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                x[u] = __ldg(reinterpret_cast<const uint32_t *>(src + (uint64_t)min(i + u, n_mine - 1) * step));
#pragma unroll
            for (int u = 0; u < 8; ++u) acc &= x[u];  // the clamped duplicates are harmless under AND
        }
    }
#pragma unroll
    for (int d = WPW; d < 32; d <<= 1) acc &= __shfl_xor_sync(0xffffffffu, acc, d);
    if (g == 0) sm[w] = valid ? acc : 0u;
    __syncthreads();
    uint8_t *out = reinterpret_cast<uint8_t *>(P.out) + (uint64_t)G.q * P.out_stride;
    const uint32_t row_bytes = (P.num_cols + 7) >> 3;
    const uint32_t nbytes_here = min((uint32_t)WPB * 4u, G.tw - G.cb);  // never past this tile
    for (uint32_t c = threadIdx.x; c < nbytes_here; c += kMergeThreads) {
        const uint32_t byte = G.tb0 + G.cb + c;
        if (byte >= row_bytes) break;
        uint32_t v = (sm[c >> 2] >> (8 * (c & 3))) & 0xffu;
        if (byte == row_bytes - 1 && (P.num_cols & 7)) v &= 0xff00u >> (P.num_cols & 7);
        out[byte] = (uint8_t)v;
    }
}

cudaError_t launch_merge(const QueryParams &p, int mode, cudaStream_t stream)
{
    if (p.n_queries == 0) return cudaSuccess;
    if (mode == kModeAnd) {
        const uint64_t cpt = (p.tile_bytes + 255) / 256;
        const uint64_t blocks = cpt * p.n_tiles * p.n_queries;
        if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
        return launch_pdl(merge_and_kernel, dim3((unsigned)blocks), dim3(kMergeThreads), 0, stream, p);
    }
    // slots one (tile, query) can span decide how many lanes share a word's slot loop
    const uint64_t max_slots = p.max_query_kmers / p.items_per_slice + 2;
    const int ng = max_slots <= 4 ? 1 : 4;
    const uint32_t chunk_bytes = (32 / ng) * 4;
    const uint64_t cpt = (p.tile_bytes + chunk_bytes - 1) / chunk_bytes;
    const uint64_t blocks = cpt * p.n_tiles * p.n_queries;
    if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    if (ng == 1) return launch_pdl(merge_counts_kernel<1>, dim3((unsigned)blocks), dim3(kMergeThreads), 0, stream, p);
    return launch_pdl(merge_counts_kernel<4>, dim3((unsigned)blocks), dim3(kMergeThreads), 0, stream, p);
}

}  // namespace bigsi
