// Stage 2 of a query launch: merge the per-segment partial planes written by stage 1.
// Device functions shared by the stand-alone merge kernels (merge_kernels.cu) and by the merge
// phase that fused_query runs itself after a grid-wide barrier (query_kernels.cu).
//
// Partial layout (written by stage 1's flush, query.cuh:partial_offset): CHUNK-major,
//   [chunk of the tile][slot][plane][merge_cb bytes]
// so the input of one merge work item = (query, tile, chunk) -- all slots the (tile, query) pair
// spans, all planes, merge_cb bytes of columns -- is ONE contiguous region.  It is staged into
// shared memory with bulk async copies (the same TMA path stage 1 gathers rows with) instead of
// thousands of strided 4-byte loads.
//
// COUNTS.  Slot s, plane b, bit c = bit b of the number of k-mers of segment s whose AND vector has
// column c set, so count(c) = sum_b 2^b * |{s : plane b of s has bit c}|: every plane is a VERTICAL
// POPCOUNT over the slots.  Threads = (word w, plane b) pairs x Gs slot groups (Gs consecutive
// lanes); each thread counts its slots with Harley-Seal carry-save blocks, the Gs groups are added
// with warp shuffles (bit-sliced full adders) and the J = bits(n_slots) counter planes of every
// (w, b) land in shared memory.  Expansion: a warp takes one word; per plane b lane j holds counter
// plane j, a 32x32 bit transpose across the warp (5 shuffle steps) turns that into "lane c holds
// the count of column c", shifted by b and summed.  The threshold (graph/bigsi.py:241-242) is
// applied while the count is in a register; hits are compacted per warp.
//
// AND.  One plane per slot; AND over the slots (graph/bigsi.py:192-195).
#pragma once
#include "ptx.cuh"
#include "query.cuh"

namespace bigsi {

constexpr uint32_t kMaxSlotGroups = 8;  // lanes that share one (unit, plane) pair; each level of the shuffle tree costs ~100 instructions
constexpr int kCntPlanes = 16;  // counter planes per (word, plane): up to 65 535 slots per (tile, query)

struct MergeGeom {
    uint32_t q, t, chunk;
    uint32_t col0;      // first column of the chunk (local column id)
    uint32_t vw;        // valid 32-bit words in this chunk (the last chunk of a tile may be narrow)
    uint64_t s_first;   // first partial slot of (tile, query)
    uint64_t n_slots;   // slices spanned by (tile, query); 0 when the query is empty
};

// item -> (query, tile, chunk); false when the chunk lies past the end of a narrow last tile
__device__ __forceinline__ bool merge_geometry(const QueryParams &P, uint64_t item, MergeGeom &g)
{
    if (P.n_queries == 1 && P.n_tiles == 1) {
        // the common single-query case without any 64-bit division: every slice belongs to (tile 0, query 0)
        g.chunk = (uint32_t)item;
        g.t = g.q = 0;
        const uint32_t cb0 = g.chunk * P.merge_cb;
        if (cb0 >= P.row_bytes16) return false;
        g.col0 = cb0 * 8;
        g.vw = min(P.merge_cb, P.row_bytes16 - cb0) >> 2;
        g.s_first = 0;
        g.n_slots = P.total_kmers ? P.n_slices : 0;
        return true;
    }
    const uint32_t cpt = P.merge_cpt;
    g.chunk = (uint32_t)(item % cpt);
    const uint64_t tq = item / cpt;
    g.t = (uint32_t)(tq % P.n_tiles);
    g.q = (uint32_t)(tq / P.n_tiles);
    const uint32_t tb0 = g.t * P.tile_bytes;
    const uint32_t tw = min(P.tile_bytes, P.row_bytes16 - tb0);
    const uint32_t cb0 = g.chunk * P.merge_cb;
    if (cb0 >= tw) return false;
    g.col0 = (tb0 + cb0) * 8;
    g.vw = min(P.merge_cb, tw - cb0) >> 2;
    // a single query spans [0, total_kmers) by contract: no dependent load in front of the planes
    const bool one = P.n_queries == 1;
    const uint64_t k0 = one ? 0ull : (uint64_t)__ldg(P.qoff + g.q);
    const uint64_t k1 = one ? P.total_kmers : (uint64_t)__ldg(P.qoff + g.q + 1);
    if (k1 > k0) {
        const uint64_t I0 = (uint64_t)g.t * P.total_kmers + k0, I1 = (uint64_t)g.t * P.total_kmers + k1;
        const uint64_t sl0 = I0 / P.items_per_slice;
        g.n_slots = (I1 - 1) / P.items_per_slice - sl0 + 1;
        g.s_first = sl0 + (uint64_t)g.t * P.n_queries + g.q;
    } else {
        g.s_first = 0;
        g.n_slots = 0;
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

// 32x32 bit-matrix transpose across a warp: in = row `lane`, out = column `lane` (bit r = in[r] bit lane)
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, uint32_t lane)
{
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const uint32_t m = s == 16 ? 0x0000ffffu : s == 8 ? 0x00ff00ffu : s == 4 ? 0x0f0f0f0fu : s == 2 ? 0x33333333u : 0x55555555u;
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, s);
        x = (lane & s) ? ((x & ~m) | ((y >> s) & m)) : ((x & m) | ((y & m) << s));
    }
    return x;
}

// The threads that execute the merge together: a whole CTA (stand-alone merge / reduce kernels, the merge phase of the
// generic kernel), or a group of warps of a gather CTA on a named barrier (the deferred merge of the previous query,
// query_kernels.cu:gather_solo).  tid() / size() replace threadIdx.x / blockDim.x, sync() replaces __syncthreads().
struct CtaTeam {
    __device__ __forceinline__ uint32_t tid() const { return threadIdx.x; }
    __device__ __forceinline__ uint32_t size() const { return blockDim.x; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
};
template <int BAR>
struct WarpGroupTeam {
    uint32_t first_thread, n;  // the group = threads [first_thread, first_thread + n), n a multiple of 32
    __device__ __forceinline__ uint32_t tid() const { return threadIdx.x - first_thread; }
    __device__ __forceinline__ uint32_t size() const { return n; }
    __device__ __forceinline__ void sync() const { named_bar_sync<BAR>((int)n); }
};

// The hit list of query 0 (n_hits, hit_cols, hit_counts) -> every sink block: payload, system-scope fence, then
// {sequence word, number of hits} as ONE 16-byte release store, so the block header is never seen torn.
// Every thread of the team must call it.
template <class Team>
__device__ __forceinline__ void publish_hits(const QueryParams &P, unsigned long long *const *sinks, uint32_t n_sinks,
                                             unsigned long long seq, const Team &T)
{
    const unsigned long long n = *reinterpret_cast<volatile unsigned long long *>(P.n_hits);
    unsigned long long m = n < P.hit_cap ? n : P.hit_cap;
    if (m > P.sink_spec) m = P.sink_spec;
    for (uint32_t sidx = 0; sidx < n_sinks; ++sidx) {
        int32_t *dc = reinterpret_cast<int32_t *>(sinks[sidx] + 2);
        uint32_t *dv = reinterpret_cast<uint32_t *>(dc + P.sink_spec);
        for (uint32_t i = T.tid(); i < (uint32_t)m; i += T.size()) {
            dc[i] = __ldcg(P.hit_cols + i);
            dv[i] = __ldcg(P.hit_counts + i);
        }
    }
    if (P.total_dev && T.tid() < n_sinks)  // the query's k-mer count was determined on the device: report it
        sinks[T.tid()][2 + P.sink_spec] = __ldcg(P.total_dev);
    // ONE system-scope fence per sink on the critical path: the CTA barrier orders every thread's payload
    // stores before the publishing thread's fence (fences are cumulative -- the same pattern grid-wide barriers
    // rely on), and the header store behind the fence can then be a plain one
    T.sync();
    if (T.tid() < n_sinks) {
        __threadfence_system();
        asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(sinks[T.tid()]), "l"(seq), "l"(n) : "memory");
    }
}

// Stage `bytes` (a multiple of 16) from global to shared memory with bulk async copies issued by warp 0
// and wait for them.  Every thread of the CTA must call it; `phase` is the CTA-uniform parity of `mbar`.
template <class Team>
__device__ __forceinline__ void merge_stage_region(const QueryParams &P, uint8_t *dst, const uint8_t *src, uint32_t bytes,
                                                   uint64_t *mbar, uint32_t &phase, const Team &T)
{
    T.sync();  // every reader of the previous contents is done
    if (T.tid() < 32) {
        const uint32_t lane = T.tid();
        if (lane == 0) BIGSI_TS(13);
        if (!(P.debug_flags & 8u)) fence_proxy_async_all();  // partial planes were written through the generic proxy (by other SMs)
        if (lane == 0) mbar_arrive_expect_tx(mbar, bytes);
        __syncwarp();
        uint32_t piece = ((bytes + 31) / 32 + 15) & ~15u;
        if (piece < 2048) piece = 2048;
        const uint32_t off = lane * piece;
        if (off < bytes) bulk_g2s(dst + off, src + off, min(piece, bytes - off), mbar);
    }
    mbar_wait(mbar, phase);
    phase ^= 1;
}

// One merge work item, executed by the whole CTA (all threads must call it: it contains
// __syncthreads).  smem: 128-byte aligned scratch of P.merge_smem bytes; mbar: an initialised
// (count 1) mbarrier owned by the merge phase; phase: its CTA-uniform parity.
template <int MODE, class Team>
__device__ __forceinline__ void merge_item(const QueryParams &P, uint64_t item, uint8_t *smem, uint64_t *mbar, uint32_t &phase,
                                           const Team &T)
{
    MergeGeom G;
    if (!merge_geometry(P, item, G)) return;  // block-uniform
    // a query inside one slice was finished by the CTA that counted it (query_kernels.cu:direct_output)
    if (MODE == kModeCounts && P.direct_complete && G.n_slots == 1) return;
    const uint32_t lane = T.tid() & 31, warp = T.tid() >> 5, nwarps = T.size() >> 5;
    const uint32_t pps = MODE == kModeCounts ? P.planes_per_slot : 1;
    const uint32_t cb = P.merge_cb;
    const uint32_t wpi = cb >> 2;
    const uint32_t J = MODE == kModeCounts ? 32 - __clz((uint32_t)G.n_slots) : 1;  // counter planes: bits(n_slots)
    uint32_t *cnt = reinterpret_cast<uint32_t *>(smem);  // COUNTS: [w][b][J]; AND: [w]
    const uint32_t cnt_bytes = (((MODE == kModeCounts ? wpi * pps * J : wpi) * 4) + 127) & ~127u;
    uint8_t *stage = smem + cnt_bytes;
    const uint32_t slot_bytes = pps * cb;
    // slots staged per batch: >= 1 by plan; <= 255 so that a batch's counters fit kLocalPlanes bit planes
    const uint32_t SB = min((P.merge_smem - cnt_bytes) / slot_bytes, 255u);
    const uint8_t *region = P.partial + partial_offset(P, G.chunk, G.s_first, 0);

    // thread layout: Gs consecutive lanes share a (16-byte unit, plane) pair and split its slots
    constexpr int kLocalPlanes = 8;
    const uint32_t pairs = (G.vw >> 2) * pps;
    uint32_t Gs = 1;
    {
        const uint32_t lim = (uint32_t)min((uint64_t)SB, G.n_slots);
        while (Gs < kMaxSlotGroups && Gs < lim && pairs * (Gs * 2) <= T.size()) Gs <<= 1;
    }
    const uint32_t g = lane & (Gs - 1);
    const uint32_t pairs_per_pass = T.size() / Gs;
    const uint32_t stage_s = smem_u32(stage);

    T.sync();  // the previous item's expansion is done with `cnt`
    if (MODE == kModeAnd && G.n_slots == 0)  // an empty query is all-ones (reduce over nothing is the identity)
        for (uint32_t w = T.tid(); w < G.vw; w += T.size()) cnt[w] = 0xffffffffu;

    for (uint64_t s0 = 0; s0 < G.n_slots; s0 += SB) {
        const uint32_t nb = (uint32_t)min((uint64_t)SB, G.n_slots - s0);
        merge_stage_region(P, stage, region + s0 * slot_bytes, nb * slot_bytes, mbar, phase, T);
        if (T.tid() == 0 && s0 == 0) BIGSI_TS(10);
        const bool first = s0 == 0;
        for (uint32_t p0 = (warp * 32) / Gs; p0 < pairs; p0 += pairs_per_pass) {  // warp-uniform trip count
            const uint32_t p = p0 + lane / Gs;
            const bool valid = p < pairs;
            const uint32_t u = valid ? p / pps : 0, b = valid ? p - u * pps : 0;
            const uint32_t n_mine = valid && g < nb ? (nb - g + Gs - 1) / Gs : 0;  // slots g, g + Gs, ...
            const uint32_t src = stage_s + g * slot_bytes + b * cb + u * 16;
            const uint32_t step = Gs * slot_bytes;
            if (MODE == kModeCounts) {
                // straight-line code on purpose: no data-dependent branches around loads / shuffles (a
                // branch per plane costs more than the plane); all kLocalPlanes planes are always processed
                uint32_t c[kLocalPlanes][4];
#pragma unroll
                for (int j = 0; j < kLocalPlanes; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) c[j][i] = 0;
                for (uint32_t s = 0; s < n_mine; s += 8) {
                    uint32_t x[8][4];
#pragma unroll
                    for (int v = 0; v < 8; ++v) {
                        const uint32_t sv = min(s + v, n_mine - 1);  // clamped: always a valid slot
                        const uint32_t keep = s + v < n_mine ? 0xffffffffu : 0u;
                        uint32_t y0, y1, y2, y3;
                        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                                     : "=r"(y0), "=r"(y1), "=r"(y2), "=r"(y3)
                                     : "r"(src + sv * step));
                        x[v][0] = y0 & keep;
                        x[v][1] = y1 & keep;
                        x[v][2] = y2 & keep;
                        x[v][3] = y3 & keep;
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        // Harley-Seal block: 8 inputs of weight 1 -> ones/twos/fours + a carry of weight 8
                        uint32_t t0 = maj3(c[0][i], x[0][i], x[1][i]);
                        c[0][i] = xor3(c[0][i], x[0][i], x[1][i]);
                        uint32_t t1 = maj3(c[0][i], x[2][i], x[3][i]);
                        c[0][i] = xor3(c[0][i], x[2][i], x[3][i]);
                        const uint32_t f0 = maj3(c[1][i], t0, t1);
                        c[1][i] = xor3(c[1][i], t0, t1);
                        t0 = maj3(c[0][i], x[4][i], x[5][i]);
                        c[0][i] = xor3(c[0][i], x[4][i], x[5][i]);
                        t1 = maj3(c[0][i], x[6][i], x[7][i]);
                        c[0][i] = xor3(c[0][i], x[6][i], x[7][i]);
                        const uint32_t f1 = maj3(c[1][i], t0, t1);
                        c[1][i] = xor3(c[1][i], t0, t1);
                        uint32_t carry = maj3(c[2][i], f0, f1);
                        c[2][i] = xor3(c[2][i], f0, f1);
#pragma unroll
                        for (int j = 3; j < kLocalPlanes; ++j) {
                            const uint32_t o = c[j][i];
                            c[j][i] = o ^ carry;
                            carry = o & carry;
                        }
                    }
                }
                // add the Gs slot groups (lanes l and l ^ d hold the same pair); the sum is <= nb <= 255
                for (uint32_t d = 1; d < Gs; d <<= 1) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint32_t carry = 0;
#pragma unroll
                        for (int j = 0; j < kLocalPlanes; ++j) {
                            const uint32_t o = __shfl_xor_sync(0xffffffffu, c[j][i], d);
                            fa(c[j][i], o, carry);
                        }
                    }
                }
                if (valid && g == 0) {
                    if (first) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            uint32_t *dst = cnt + ((size_t)(u * 4 + i) * pps + b) * J;
#pragma unroll
                            for (int j = 0; j < kLocalPlanes; ++j)
                                if (j < (int)J) dst[j] = c[j][i];
                            for (uint32_t j = kLocalPlanes; j < J; ++j) dst[j] = 0u;
                        }
                    } else {  // later slot batches are added to the counters of the earlier ones
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            uint32_t *dst = cnt + ((size_t)(u * 4 + i) * pps + b) * J;
                            uint32_t carry = 0;
#pragma unroll
                            for (int j = 0; j < kLocalPlanes; ++j) {
                                if (j < (int)J) {
                                    uint32_t a = dst[j];
                                    fa(a, c[j][i], carry);
                                    dst[j] = a;
                                }
                            }
                            for (uint32_t j = kLocalPlanes; j < J; ++j) {
                                const uint32_t a = dst[j];
                                dst[j] = a ^ carry;
                                carry = a & carry;
                            }
                        }
                    }
                }
            } else {
                uint32_t acc[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
                for (uint32_t s = 0; s < n_mine; ++s) {
                    uint32_t x[4];
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3])
                                 : "r"(src + s * step));
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[i] &= x[i];
                }
                for (uint32_t d = 1; d < Gs; d <<= 1)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[i] &= __shfl_xor_sync(0xffffffffu, acc[i], d);
                if (valid && g == 0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) cnt[u * 4 + i] = first ? acc[i] : (cnt[u * 4 + i] & acc[i]);
                }
            }
        }
    }
    T.sync();
    if (T.tid() == 0) BIGSI_TS(11);

    if (MODE == kModeCounts) {
        uint32_t *out = P.out ? reinterpret_cast<uint32_t *>(P.out) + (uint64_t)G.q * P.out_stride : nullptr;
        const bool thresholding = P.min_kmers != nullptr || P.min_by_value || P.seq_mode;
        uint32_t thr = P.min_by_value ? P.min_kmers_value : (P.min_kmers != nullptr ? __ldg(P.min_kmers + G.q) : 0u);
        if (P.seq_mode) {
            // min_kmers = math.ceil(U * threshold) in IEEE double (graph/bigsi.py:179) with the U the gather kernel's
            // front-end counted; <= 0 keeps every column
            const double need = ceil(__dmul_rn((double)__ldcg(&P.qstate->n_unique), P.seq_threshold));
            thr = need <= 0.0 ? 0u : need >= 4294967295.0 ? 0xffffffffu : (uint32_t)need;
        }
        // Word-parallel expansion: thread = one 32-column word.  Its pps x J planes (plane (b, j) has weight 2^(b+j))
        // are added into a BINARY bit-sliced accumulator of np = bits(longest query) planes -- a ripple-carry add per
        // plane, 32 columns at a time -- kept in the (now free) staging area, plane-major so that the warp's accesses
        // never conflict.  The threshold is a bit-sliced >= comparison against the constant; only hits (and, with a
        // count buffer, all columns) are turned into integers.  ~20 instructions per word and warp instead of ~130
        // for the column-per-lane expansion below, which remains for narrow items and for scratch areas too small for
        // the accumulators.
        const uint32_t np = P.total_planes < 1 ? 1 : P.total_planes;
        const uint32_t stage_bytes_avail = P.merge_smem - cnt_bytes;
        // (few words per item -- a single query cut into ~one item per CTA -- leave most threads idle here: the
        // column-per-lane expansion below is faster then: 2.3 us against 8.5 us for the 12 words of a config-2 item)
        if ((uint64_t)T.size() * np * 4 <= stage_bytes_avail && G.vw * 4 >= T.size()) {
            uint32_t *acc = reinterpret_cast<uint32_t *>(stage) + T.tid();
            const uint32_t ts = T.size();
            for (uint32_t w0 = 0; w0 < G.vw; w0 += ts) {  // team-uniform trip count (warp shuffles inside)
                const uint32_t w = w0 + T.tid();
                const bool have = w < G.vw;
                uint32_t ge = 0, valid = 0;
                const uint32_t wcol0 = G.col0 + w * 32;
                if (have) {
                    for (uint32_t p = 0; p < np; ++p) acc[p * ts] = 0u;
                    const uint32_t *cw = cnt + (size_t)w * pps * J;
                    for (uint32_t b = 0; b < pps; ++b)
                        for (uint32_t j = 0; j < J; ++j) {
                            uint32_t x = cw[b * J + j];
                            for (uint32_t p = b + j; x && p < np; ++p) {
                                const uint32_t a = acc[p * ts];
                                acc[p * ts] = a ^ x;
                                x &= a;
                            }
                        }
                    if (wcol0 < P.num_cols) {  // columns of this word that exist (bit i <-> column wcol0 + (i ^ 7))
                        const uint32_t rem = P.num_cols - wcol0;
                        if (rem >= 32) valid = 0xffffffffu;
                        else
                            for (uint32_t k8 = 0; k8 < 4; ++k8) {
                                const uint32_t r = rem > 8 * k8 ? rem - 8 * k8 : 0;
                                valid |= (r >= 8 ? 0xffu : r ? ((0xffu << (8 - r)) & 0xffu) : 0u) << (8 * k8);
                            }
                    }
                    if (out)
                        for (uint32_t i = 0; i < 32; ++i)
                            if ((valid >> i) & 1u) {
                                uint32_t v = 0;
                                for (uint32_t p = 0; p < np; ++p) v |= ((acc[p * ts] >> i) & 1u) << p;
                                out[wcol0 + (i ^ 7u)] = v;
                            }
                    if (thresholding) {
                        uint32_t g = 0xffffffffu;  // "equal so far" counts as >=
                        for (uint32_t p = 0; p < np; ++p) {
                            const uint32_t a = acc[p * ts];
                            g = ((thr >> p) & 1u) ? (a & g) : (a | g);
                        }
                        ge = (np < 32 && (thr >> np)) ? 0u : (g & valid);  // beyond the counters' range: never reached
                    }
                }
                if (thresholding) {  // counts >= min_kmers (graph/bigsi.py:241-242), warp-level compaction
                    const uint32_t mine = __popc(ge);
                    uint32_t incl = mine;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= (uint32_t)d) incl += o;
                    }
                    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                    if (total) {  // warp-uniform
                        unsigned long long base = 0;
                        if (lane == 31) base = atomicAdd(P.n_hits + G.q, (unsigned long long)total);
                        base = __shfl_sync(0xffffffffu, base, 31);
                        uint64_t pos = base + (incl - mine);
                        while (ge) {
                            const uint32_t i = __ffs(ge) - 1;
                            ge &= ge - 1;
                            if (pos < P.hit_cap) {
                                uint32_t v = 0;
                                for (uint32_t p = 0; p < np; ++p) v |= ((acc[p * ts] >> i) & 1u) << p;
                                P.hit_cols[(uint64_t)G.q * P.hit_cap + pos] = (int32_t)(wcol0 + (i ^ 7u));
                                P.hit_counts[(uint64_t)G.q * P.hit_cap + pos] = v;
                            }
                            ++pos;
                        }
                    }
                }
            }
        } else
        for (uint32_t w = warp; w < G.vw; w += nwarps) {
            uint32_t total = 0;
            const uint32_t *cw = cnt + (size_t)w * pps * J;
            if (J <= 4) {
                for (uint32_t b = 0; b < pps; ++b)
                    for (uint32_t j = 0; j < J; ++j) total += ((cw[b * J + j] >> lane) & 1u) << (b + j);
            } else {
#pragma unroll 2
                for (uint32_t b = 0; b < pps; ++b) {
                    const uint32_t x = lane < J ? cw[b * J + lane] : 0u;
                    total += warp_transpose32(x, lane) << b;
                }
            }
            // bit i of a little-endian 32-bit word of MSB-first bytes is column (i ^ 7) of that word
            const uint32_t col = G.col0 + w * 32 + (lane ^ 7);
            const bool live = col < P.num_cols;
            if (live && out) out[col] = total;
            if (thresholding) {  // counts >= min_kmers (graph/bigsi.py:241-242), warp-level compaction
                const bool hit = live && total >= thr;
                const uint32_t ballot = __ballot_sync(0xffffffffu, hit);
                if (ballot) {
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(P.n_hits + G.q, (unsigned long long)__popc(ballot));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (hit) {
                        const uint64_t pos = base + __popc(ballot & ((1u << lane) - 1));
                        if (pos < P.hit_cap) {
                            P.hit_cols[(uint64_t)G.q * P.hit_cap + pos] = (int32_t)col;
                            P.hit_counts[(uint64_t)G.q * P.hit_cap + pos] = total;
                        }
                    }
                }
            }
        }
    } else {
        uint8_t *out = reinterpret_cast<uint8_t *>(P.out) + (uint64_t)G.q * P.out_stride;
        const uint32_t row_bytes = (P.num_cols + 7) >> 3;
        const uint32_t byte0 = G.col0 >> 3;
        for (uint32_t c = T.tid(); c < G.vw * 4; c += T.size()) {
            const uint32_t byte = byte0 + c;
            if (byte >= row_bytes) break;
            uint32_t v = (cnt[c >> 2] >> (8 * (c & 3))) & 0xffu;
            if (byte == row_bytes - 1 && (P.num_cols & 7)) v &= 0xff00u >> (P.num_cols & 7);
            out[byte] = (uint8_t)v;
        }
    }
    if (T.tid() == 0) BIGSI_TS(12);
    // the next item's staging starts with __syncthreads, which also protects `cnt`
}

// Stage 2 of a STREAMED query (query_kernels.cu:gather_solo wrote its planes and is complete): merge + threshold +
// publication + completion chain, executed by `n_ctas` teams (one per CTA of the executing grid; `cta` = this team's
// index).  Two executors: the merge warps of the NEXT query's gather kernel (the normal case for back-to-back
// queries: the merge of query s overlaps the gather of query s+1 inside the same CTAs, so nothing has to become
// co-resident with anything) and reduce_kernel (merge_kernels.cu: the flush behind the last query of a burst and
// behind every synchronous call).
// The last team to finish publishes the hit list (host block and / or every shard's result blocks), waits -- bounded
// -- for the other shards' blocks of the same query, clears the state block of query seq + kStreamRing and advances
// the handle's completion word in query order.  s_flag: a shared-memory word owned by the team.
template <int MODE, class Team>
__device__ __forceinline__ void reduce_query(const QueryParams &P, uint8_t *merge_smem, uint64_t *bar, volatile int *s_flag,
                                             const Team &T, uint32_t cta, uint32_t n_ctas)
{
    uint32_t phase = 0;
    for (uint64_t item = cta; item < P.merge_items; item += n_ctas) merge_item<MODE>(P, item, merge_smem, bar, phase, T);
    if (P.scrub_words) {
        // query front-end (general path): its de-duplication table is dead once the gather kernel has read the
        // k-mers; clear it for the next query here, off every critical path
        for (uint64_t i = (uint64_t)cta * T.size() + T.tid(); i < P.scrub_words; i += (uint64_t)n_ctas * T.size()) P.scrub[i] = 0ull;
    }
    T.sync();
    if (T.tid() == 0) {
        BIGSI_TS(7);
        __threadfence();
        *s_flag = atomicAdd(&P.qstate->reduce_arrivals, 1u) + 1u == n_ctas;
    }
    T.sync();
    if (!*s_flag) return;
    // ---- the last team of the query ------------------------------------------------------------------------
    __threadfence();
    if (P.n_sinks) publish_hits(P, P.sinks, P.n_sinks, P.sink_seq, T);
    // front-end words behind its table ({U}, {ticket, threshold}): every reader is done, re-arm them
    if (P.scrub_words && T.tid() < 2) P.scrub[P.scrub_words + T.tid()] = 0ull;
    if (T.tid() < P.n_gather) {  // all-gather: every shard's block of this query has arrived here
        const unsigned long long t0 = globaltimer_ns();
        const unsigned long long *blk = P.gather_blocks[T.tid()];
        bounded_wait(P.abort_word, P.host_abort, P.spin_timeout_ns, kAbortPeers, P.stream_seq,
                     [&]() { return ld_acquire_sys_u64(blk) == P.gather_seq; });
        atomicMax(&P.qstate->wait_ns, globaltimer_ns() - t0);
    }
    T.sync();
    if (P.host_gather && P.n_gather) {
        // every shard's hit list -> mapped host memory (the waiting threads' acquire loads + the barrier above order
        // these reads behind the peers' publications; volatile loads: nothing stale from L1)
        for (uint32_t r = 0; r < P.n_gather; ++r) {
            const unsigned long long *src = P.gather_blocks[r];
            unsigned long long *dst = P.host_gather + 2 + (size_t)r * P.host_block_words;
            const unsigned long long n = ld_volatile_u64(src + 1);
            const uint32_t m = (uint32_t)(n < P.sink_spec ? n : P.sink_spec);
            if (T.tid() == 0) {
                dst[0] = P.gather_seq;
                dst[1] = n;
            }
            const volatile uint32_t *sc = reinterpret_cast<const volatile uint32_t *>(src + 2);
            uint32_t *dc = reinterpret_cast<uint32_t *>(dst + 2);
            for (uint32_t i = T.tid(); i < m; i += T.size()) {
                dc[i] = sc[i];
                dc[P.sink_spec + i] = sc[P.sink_spec + i];
            }
        }
        T.sync();
        if (T.tid() == 0) {
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long *>(P.host_gather) = P.gather_seq;
        }
    }
    if (T.tid() == 0) {
        BIGSI_TS(6);
        if (P.wait_ns_out) atomicAdd(P.wait_ns_out, P.qstate->wait_ns);
    }
    T.sync();
    if (T.tid() < sizeof(QState) / 8) reinterpret_cast<unsigned long long *>(P.qstate_next)[T.tid()] = 0ull;
    T.sync();
    if (T.tid() == 0) {
        // completion in query order: "done >= s" implies every query <= s is reduced and its ring slots are free
        bounded_wait(P.abort_word, P.host_abort, P.spin_timeout_ns, kAbortChain, P.stream_seq,
                     [&]() { return ld_acquire_gpu_u64(P.stream_done) + 1ull >= P.stream_seq; });
        __threadfence();
        st_release_gpu_u64(P.stream_done, P.stream_seq);
    }
}

// ------------------------------------------------------------------------------------------
// launch geometry of the merge work (fixes the partial layout, so stage 1 needs it too)
// ------------------------------------------------------------------------------------------
constexpr uint32_t kMergeKernelThreads = 256;
constexpr uint32_t kMergeKernelSmem = 64 * 1024;  // dynamic shared memory of the stand-alone merge kernels

// Fills p.merge_cb / merge_cpt / merge_items / merge_smem.  `smem` = scratch one item may use,
// `ctas` = CTAs that will share the items.  Needs tile geometry, slices and planes_per_slot.
inline void plan_merge(QueryParams &p, int mode, uint32_t smem, uint64_t ctas, uint32_t cb_override = 0)
{
    const uint32_t pps = mode == kModeCounts ? (p.planes_per_slot ? p.planes_per_slot : 1) : 1;
    const uint64_t max_slots = p.items_per_slice ? p.max_query_kmers / p.items_per_slice + 2 : 2;
    uint32_t jmax = 0;
    for (uint64_t v = max_slots; v; v >>= 1) ++jmax;
    if (mode != kModeCounts) jmax = 1;
    // parallelism: about one item per CTA when there are few (tile, query) pairs
    const uint64_t tq = (uint64_t)p.n_tiles * (p.n_queries ? p.n_queries : 1);
    uint64_t cb = p.tile_bytes;
    if (tq < ctas) {
        const uint64_t per = (ctas + tq - 1) / tq;  // chunks per tile wanted
        cb = (p.tile_bytes + per - 1) / per;
    }
    cb = (cb + 15) / 16 * 16;
    // memory: counters [wpi][pps][jmax] + at least min(max_slots, 16) staged slots must fit
    const uint64_t batch = max_slots < 16 ? max_slots : 16;
    uint64_t cb_mem = (uint64_t)(smem - 256) / ((uint64_t)pps * (jmax + batch)) / 16 * 16;
    if (cb_mem < 16) cb_mem = 16;
    if (cb > cb_mem) cb = cb_mem;
    if (cb > p.tile_bytes) cb = p.tile_bytes;
    if (cb < 16) cb = 16;
    uint64_t cpt = (p.tile_bytes + cb - 1) / cb;
    cb = ((p.tile_bytes + cpt - 1) / cpt + 15) / 16 * 16;  // balance the chunks of a tile
    if (cb_override) {
        cb = ((uint64_t)cb_override + 15) / 16 * 16;
        if (cb > cb_mem) cb = cb_mem;
        if (cb > p.tile_bytes) cb = p.tile_bytes;
    }
    cpt = (p.tile_bytes + cb - 1) / cb;
    p.merge_cb = (uint32_t)cb;
    p.merge_cpt = (uint32_t)cpt;
    p.merge_items = cpt * p.n_tiles * p.n_queries;
    p.merge_smem = smem;
}

}  // namespace bigsi
