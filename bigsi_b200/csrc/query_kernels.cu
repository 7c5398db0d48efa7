// Fused gather-AND-{popcount|AND} kernel and its merge kernel for sm_100a.
//
// Replaces, for a whole batch of queries in one launch, the reference's
//   BitMatrix.get_rows            (bigsi/matrix/bitmatrix.py:30-37, storage/base.py:106-109)
//   per-k-mer bitwise_and         (bigsi/graph/index.py:75-80, utils/fncts.py:24-25)
//   exact_filter AND over k-mers  (bigsi/graph/bigsi.py:192-195)
//   unpack_and_sum column counts  (bigsi/graph/bigsi.py:35-44, 211-219)
//
// Work space: items = (column tile, global k-mer index), k-mer fastest.  Every CTA owns an equal,
// contiguous run of slices of it; a slice is cut into SEGMENTS at (tile, query) boundaries; one
// segment accumulates in registers and is written once, as bit planes, to its partial slot.
//
// Stage 1 (fused_query): persistent, warp specialised, one CTA per SM.  The last warp is the
// producer: each lane issues one 1-D bulk async copy (cp.async.bulk, SASS UBLKCP) of one row
// segment (tile_bytes of one gathered row) into a shared-memory ring slot; slot = G k-mers x h
// rows, completion counted in bytes on an mbarrier.  The other warps are consumers: thread t owns
// the 16-byte unit t of the tile (128 sample columns), reads the h row segments of each staged
// k-mer as conflict-free LDS.128, ANDs them (LOP3) and feeds the 128-bit vector into a bit-sliced
// carry-save counter (Harley-Seal: ones/twos/fours + ripple planes => up to 16-bit vertical
// counters), so the kernel stays HBM-bound instead of ALU-bound.
//
// Stage 2 (merge_kernels.cu): per (query, tile, column chunk) counts the partial planes of all
// segments vertically once more and expands them to uint32 counts (or ANDs the presence planes).
#include "hash.cuh"
#include "launch.cuh"
#include "merge.cuh"
#include "ptx.cuh"
#include "query.cuh"

namespace bigsi {

struct W4 {
    uint32_t v[4];
};
__device__ __forceinline__ W4 to_w4(const uint4 &a) { return W4{{a.x, a.y, a.z, a.w}}; }

// ------------------------------------------------------------------------------------------
// segment iteration (identical on producer and consumer side)
// ------------------------------------------------------------------------------------------
struct Seg {
    uint64_t kg0;  // first global k-mer index
    uint32_t nk;   // k-mers in this segment (<= items_per_slice)
    uint32_t tile;
    uint32_t q;
    uint32_t slice;
};

struct SegIter {
    uint64_t p, end, total;
    const int64_t *qoff;
    uint32_t nq, ips, q, tile;
    bool have_tile;

    __device__ SegIter(const QueryParams &P, uint64_t b, uint64_t e)
        : p(b), end(e), total(P.total_kmers), qoff(P.qoff), nq(P.n_queries), ips(P.items_per_slice), q(0),
          tile(0), have_tile(false)
    {
    }
    // end offset of query q; a single query ends at total_kmers by contract (no global load)
    __device__ __forceinline__ uint64_t qend(uint32_t qq) const
    {
        return nq == 1 ? total : (uint64_t)__ldg(qoff + qq + 1);
    }
    __device__ uint32_t find_query(uint64_t kg) const
    {
        // q with qoff[q] <= kg < qoff[q+1]  (skips empty queries)
        uint32_t lo = 0, hi = nq;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if ((uint64_t)__ldg(qoff + mid) <= kg) lo = mid; else hi = mid;
        }
        return lo;
    }
    __device__ bool next(Seg &s)
    {
        if (p >= end) return false;
        const uint32_t t = (uint32_t)(p / total);
        const uint64_t kg = p - (uint64_t)t * total;
        if (!have_tile || t != tile) {
            tile = t;
            have_tile = true;
            q = find_query(kg);
        } else {
            while (kg >= qend(q)) ++q;
        }
        uint64_t lim = qend(q) - kg;  // to the end of the query (<= end of tile)
        if (end - p < lim) lim = end - p;
        const uint64_t sl = p / ips;
        const uint64_t to_slice_end = (sl + 1) * ips - p;
        if (to_slice_end < lim) lim = to_slice_end;
        s.kg0 = kg;
        s.nk = (uint32_t)lim;
        s.tile = t;
        s.q = q;
        s.slice = (uint32_t)sl;
        p += lim;
        return true;
    }
};

// ------------------------------------------------------------------------------------------
// bit-sliced vertical counter: 128 columns x 16 bits per thread
// ------------------------------------------------------------------------------------------
// NP = most planes a segment may need (registers: 4 x (NP + 3) words of state per thread)
template <int NP>
struct VCounter {
    static constexpr int kHiPlanes = NP - 3;  // planes of weight 8 .. 2^(NP-1)
    W4 ones, twos, fours;  // accumulators of weight 1, 2, 4
    W4 p0, p1, p2;         // pending operands waiting for a partner at each level
    W4 hi[kHiPlanes];
    uint32_t n;            // inputs so far (uniform across the CTA)

    __device__ __forceinline__ void reset()
    {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            ones.v[j] = twos.v[j] = fours.v[j] = 0;
            p0.v[j] = p1.v[j] = p2.v[j] = 0;
#pragma unroll
            for (int b = 0; b < kHiPlanes; ++b) hi[b].v[j] = 0;
        }
        n = 0;
    }
    // acc <- acc ^ a ^ b ; carry <- maj(acc_old, a, b)
    static __device__ __forceinline__ void csa(W4 &carry, W4 &acc, const W4 &a, const W4 &b)
    {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t o = acc.v[j];
            acc.v[j] = xor3(o, a.v[j], b.v[j]);
            carry.v[j] = maj3(o, a.v[j], b.v[j]);
        }
    }
    // nhi = number of live hi planes for this segment (uniform)
    __device__ __forceinline__ void add(const W4 &x, uint32_t nhi)
    {
        const uint32_t i = n++;
        if (!(i & 1)) { p0 = x; return; }
        W4 c1;
        csa(c1, ones, p0, x);
        if (!(i & 2)) { p1 = c1; return; }
        W4 c2;
        csa(c2, twos, p1, c1);
        if (!(i & 4)) { p2 = c2; return; }
        W4 c;
        csa(c, fours, p2, c2);
#pragma unroll
        for (int b = 0; b < kHiPlanes; ++b) {
            if (b < (int)nhi) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t o = hi[b].v[j];
                    hi[b].v[j] = o ^ c.v[j];
                    c.v[j] = o & c.v[j];
                }
            }
        }
    }
    // fold pending operands: pad the input stream with zero vectors to a multiple of 8
    __device__ __forceinline__ void finish(uint32_t nhi)
    {
        const W4 z = {{0, 0, 0, 0}};
        while (n & 7) add(z, nhi);
    }
};

__device__ __forceinline__ void stg128(void *p, const W4 &x)
{
    *reinterpret_cast<uint4 *>(p) = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
}

// write the planes of one segment to its partial slot: plane b of this thread's 16-byte unit goes to
// slot_unit + b * plane_stride (plane_stride = merge_cb, see partial_offset).  All planes_per_slot
// planes are written (planes the segment cannot reach are zero in the counter), so the merge loads
// them without looking at segment lengths.
template <int NP>
__device__ __forceinline__ void flush_planes(const VCounter<NP> &c, uint32_t pps, uint8_t *slot_unit, uint32_t plane_stride)
{
    stg128(slot_unit, c.ones);
    if (pps > 1) stg128(slot_unit + plane_stride, c.twos);
    if (pps > 2) stg128(slot_unit + 2 * (size_t)plane_stride, c.fours);
#pragma unroll
    for (int b = 0; b < VCounter<NP>::kHiPlanes; ++b)
        if (b + 3 < (int)pps) stg128(slot_unit + (size_t)(b + 3) * plane_stride, c.hi[b]);
}

// A segment that covers a whole query: its counters ARE the query's counts.  Every thread holds the bit planes of
// its 128 columns (word j of the unit, bit i <-> column 128*unit + 32*j + (i ^ 7): MSB-first bytes in little-endian
// words).  Threshold by a bit-sliced >= comparison against the constant (2 operations per plane and word), hits
// compacted per warp with one atomic, their counts extracted bit by bit; with a count buffer every column is
// extracted.  Called by all threads of the consumer warps (warp shuffles inside); `active` = the unit lies in the tile.
template <int NP>
__device__ __forceinline__ void direct_output(const QueryParams &P, const VCounter<NP> &c, uint32_t q, uint32_t unit_col0, bool active)
{
    const uint32_t lane = threadIdx.x & 31;
    auto plane = [&](int b, int j) -> uint32_t {
        return b == 0 ? c.ones.v[j] : b == 1 ? c.twos.v[j] : b == 2 ? c.fours.v[j] : c.hi[b - 3].v[j];
    };
    auto count_of = [&](int j, uint32_t i) -> uint32_t {
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < NP; ++b) v |= ((plane(b, j) >> i) & 1u) << b;
        return v;
    };
    // columns of word j that exist: padding bits of the last byte(s) never count
    uint32_t valid[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t w0 = unit_col0 + 32u * j;
        uint32_t m = 0;
        if (active && w0 < P.num_cols) {
            const uint32_t rem = P.num_cols - w0;
            if (rem >= 32) m = 0xffffffffu;
            else
                for (uint32_t k8 = 0; k8 < 4; ++k8) {
                    const uint32_t r = rem > 8 * k8 ? rem - 8 * k8 : 0;
                    m |= (r >= 8 ? 0xffu : r ? ((0xffu << (8 - r)) & 0xffu) : 0u) << (8 * k8);
                }
        }
        valid[j] = m;
    }
    if (P.out) {
        uint32_t *out = reinterpret_cast<uint32_t *>(P.out) + (uint64_t)q * P.out_stride;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            for (uint32_t i = 0; i < 32; ++i)
                if ((valid[j] >> i) & 1u) out[unit_col0 + 32u * j + (i ^ 7u)] = count_of(j, i);
    }
    if (P.min_kmers == nullptr && !P.min_by_value) return;
    const uint32_t thr = P.min_by_value ? P.min_kmers_value : __ldg(P.min_kmers + q);
    uint32_t ge[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t g = 0xffffffffu;  // "equal so far" counts as >=
#pragma unroll
        for (int b = 0; b < NP; ++b) g = ((thr >> b) & 1u) ? (plane(b, j) & g) : (plane(b, j) | g);
        ge[j] = (thr >> NP) ? 0u : (g & valid[j]);  // a threshold beyond the counter's range is never reached
    }
    const uint32_t mine = __popc(ge[0]) + __popc(ge[1]) + __popc(ge[2]) + __popc(ge[3]);
    uint32_t incl = mine;  // inclusive prefix sum over the warp
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += o;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;  // warp-uniform
    unsigned long long base = 0;
    if (lane == 31) base = atomicAdd(P.n_hits + q, (unsigned long long)total);
    base = __shfl_sync(0xffffffffu, base, 31);
    uint64_t pos = base + (incl - mine);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t g = ge[j];
        while (g) {
            const uint32_t i = __ffs(g) - 1;
            g &= g - 1;
            if (pos < P.hit_cap) {
                P.hit_cols[(uint64_t)q * P.hit_cap + pos] = (int32_t)(unit_col0 + 32u * j + (i ^ 7u));
                P.hit_counts[(uint64_t)q * P.hit_cap + pos] = count_of(j, i);
            }
            ++pos;
        }
    }
}

// ------------------------------------------------------------------------------------------
// stage 1: producer warp
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_producer(const QueryParams &P, uint8_t *ring, const int32_t *ids, uint64_t *full,
                                             uint64_t *empty, uint64_t begin, uint64_t end)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t h = P.h;
    const uint32_t G = P.kmers_per_stage;
    const uint32_t seg_stride = P.tile_bytes;
    const uint32_t stage_bytes = G * h * seg_stride;
    const uint64_t policy = policy_evict_first();
    const uint64_t total = P.total_kmers;
    const uint64_t n_copies = (end - begin) * h;  // row copies this CTA issues, in item order
    uint32_t stage = 0, parity = 0;

    // Row ids are read through a two-deep window of 32 consecutive copy slots per warp register,
    // so the ids of the next >= 32 copies are always in registers before their ring slot frees up.
    auto load_id = [&](uint64_t c) -> int32_t {
        if (c >= n_copies) return 0;
        if (P.prehash) return ids[c];  // hashed by this CTA in the prologue (shared memory)
        const uint64_t item = begin + c / h;
        const uint32_t j = (uint32_t)(c % h);
        const uint64_t kg = item % total;
        return __ldg(P.rows + kg * h + j);
    };
    uint64_t win_base = 0;
    int32_t cur = load_id(lane), nxt = load_id(32 + lane);
    uint64_t c0 = 0;  // copy index of the first row of the current ring slot

    SegIter it(P, begin, end);
    Seg s;
    bool first = true;
    while (it.next(s)) {
        const uint32_t tb0 = s.tile * P.tile_bytes;
        const uint32_t tw = min(P.tile_bytes, P.row_bytes16 - tb0);
        const uint8_t *src0 = P.matrix + tb0;
        if (first && lane == 0) BIGSI_TS(1);
        first = false;
        for (uint32_t k0 = 0; k0 < s.nk; k0 += G) {
            const uint32_t n_rows = min(G, s.nk - k0) * h;
            uint8_t *dst0 = ring + (size_t)stage * stage_bytes;
            mbar_wait(&empty[stage], parity ^ 1);  // slot drained by all consumer warps
            if (lane == 0) mbar_arrive_expect_tx(&full[stage], n_rows * tw);
            __syncwarp();
            for (uint32_t i0 = 0; i0 < n_rows; i0 += 32) {
                const uint64_t cb = c0 + i0;
                while (cb >= win_base + 32) {
                    cur = nxt;
                    win_base += 32;
                    nxt = load_id(win_base + 32 + lane);
                }
                const uint32_t off = (uint32_t)(cb - win_base) + lane;  // 0..62
                const int32_t r0 = __shfl_sync(0xffffffffu, cur, off & 31);
                const int32_t r1 = __shfl_sync(0xffffffffu, nxt, off & 31);
                const int32_t row = off < 32 ? r0 : r1;
                const uint32_t i = i0 + lane;
                if (i < n_rows)
                    bulk_g2s_hint(dst0 + (size_t)i * seg_stride, src0 + (uint64_t)(uint32_t)row * P.pitch, tw,
                                  &full[stage], policy);
            }
            c0 += n_rows;
            if (++stage == P.n_stages) {
                stage = 0;
                parity ^= 1;
            }
        }
    }
    if (lane == 0) BIGSI_TS(5);
}

// ------------------------------------------------------------------------------------------
// stage 1: consumer warps
// ------------------------------------------------------------------------------------------
template <int MODE, int HC>
__device__ __forceinline__ void tma_consumer(const QueryParams &P, const uint8_t *ring, uint64_t *full,
                                             uint64_t *empty, uint64_t begin, uint64_t end)
{
    const uint32_t unit = threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t h = HC ? HC : P.h;
    const uint32_t G = P.kmers_per_stage;
    const uint32_t seg_stride = P.tile_bytes;
    const uint32_t stage_bytes = G * h * seg_stride;
    uint32_t stage = 0, parity = 0;

    VCounter<kSegPlanes> ctr;
    W4 acc;
    SegIter it(P, begin, end);
    Seg s;
    bool first_slot = true;
    while (it.next(s)) {
        const uint32_t tb0 = s.tile * P.tile_bytes;
        const uint32_t tw = min(P.tile_bytes, P.row_bytes16 - tb0);
        const bool active = unit * 16 < tw;
        const uint32_t nplanes = 32 - __clz(s.nk);
        const uint32_t nhi = nplanes > 3 ? nplanes - 3 : 0;
        if (MODE == kModeCounts) ctr.reset();
        else acc = W4{{0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu}};

        for (uint32_t k0 = 0; k0 < s.nk; k0 += G) {
            const uint32_t gn = min(G, s.nk - k0);
            mbar_wait(&full[stage], parity);  // all row segments of this slot have landed
            if (first_slot && threadIdx.x == 0) BIGSI_TS(2);
            first_slot = false;
            if (active && !(P.debug_flags & 1u)) {
                const uint8_t *base = ring + (size_t)stage * stage_bytes + unit * 16;
                for (uint32_t g = 0; g < gn; ++g) {
                    const uint8_t *b = base + g * h * seg_stride;
                    W4 x = to_w4(lds128(b));
                    if (HC == 3) {
                        const W4 y = to_w4(lds128(b + seg_stride));
                        const W4 z = to_w4(lds128(b + 2 * seg_stride));
#pragma unroll
                        for (int j = 0; j < 4; ++j) x.v[j] = and3(x.v[j], y.v[j], z.v[j]);
                    } else {
                        for (uint32_t r = 1; r < h; ++r) {
                            const W4 y = to_w4(lds128(b + r * seg_stride));
#pragma unroll
                            for (int j = 0; j < 4; ++j) x.v[j] &= y.v[j];
                        }
                    }
                    if (MODE == kModeCounts) {
                        ctr.add(x, nhi);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc.v[j] &= x.v[j];
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);  // this warp is done reading the slot
            if (++stage == P.n_stages) {
                stage = 0;
                parity ^= 1;
            }
        }
        if (threadIdx.x == 0) BIGSI_TS(3);
        // the whole query in this one segment: finished here, no partial planes, no merge (QueryParams::direct_complete)
        const bool complete = MODE == kModeCounts && P.direct_complete &&
                              (uint64_t)s.nk == it.qend(s.q) - (P.n_queries == 1 ? 0ull : (uint64_t)__ldg(P.qoff + s.q));
        if (complete) {
            ctr.finish(nhi);
            direct_output(P, ctr, s.q, (tb0 + unit * 16) * 8, active);
        } else if (active) {
            const uint64_t slot = (uint64_t)s.slice + (uint64_t)s.tile * P.n_queries + s.q;
            const uint32_t chunk = unit * 16 / P.merge_cb;  // merge chunk this unit's columns belong to
            uint8_t *dst = P.partial + partial_offset(P, chunk, slot, 0) + (unit * 16 - chunk * P.merge_cb);
            if (MODE == kModeCounts) {
                ctr.finish(nhi);
                flush_planes(ctr, P.planes_per_slot, dst, P.merge_cb);
            } else {
                stg128(dst, acc);
            }
        }
        if (threadIdx.x == 0) BIGSI_TS(4);
    }
}

// ------------------------------------------------------------------------------------------
// "solo" path = STREAMED single-query launch: one query, one tile, one slice per CTA, k-mers hashed in the kernel.
// gather_solo hashes, gathers, ANDs, counts, writes its CTA's bit planes and exits -- no grid barrier, no merge phase
// of its own.  It is launched with programmatic dependent launch and its gather warps do NOT wait for the preceding
// kernel, so the gather CTA of query s+1 takes over each SM the moment the gather CTA of query s leaves it.  Stage 2
// of query s (merge.cuh:reduce_query) is executed by the MERGE TEAM of the gather kernel of query s+1: four extra
// warps per CTA that wait for the preceding grid (query s's gather kernel) to complete and then share the merge
// items, in the shadow of the row stream of query s+1.  Behind the last query of a burst, reduce_kernel
// (merge_kernels.cu) does the same as a kernel of its own.  Everything a query owns rotates (query.cuh:kStreamRing);
// the gate at kernel entry makes query s wait until query s - kStreamRing is completely reduced (normally long ago:
// one L2 read).
// ------------------------------------------------------------------------------------------
constexpr int kBarHashGroup = 1;   // named barrier of the consumer warps while they hash (== the id GroupSync uses)
constexpr int kBarIdsReady = 2;    // consumers -> producer: the id table is complete
constexpr int kBarGather = 3;      // consumer warps + producer warp (the sequence front-end)
constexpr int kBarMergeTeam = 4;   // the merge team
constexpr uint32_t kTeamBarOffset = 544;   // uint64: staging mbarrier of the merge team
constexpr uint32_t kTeamFlagOffset = 552;  // int: the team's "last CTA" flag
constexpr uint32_t kStageCntOffset = 640;  // uint32 [kMaxStages] k-mers per ring slot, in the shared-memory header
constexpr uint32_t kPoolStashOffset = 768; // int32 [kPoolBatch][kPoolMaxH] row ids of one claimed pool batch
constexpr uint32_t kSeqCountOffset = 536;  // uint32: unique windows of this CTA (sequence front-end)
constexpr uint32_t kGateOffset = 528;      // int: 1 = the entry gate passed (behind the 2 x kMaxStages mbarriers)
constexpr int kPoolBatch = 8;

// threads of a streamed gather CTA that gather: the consumer warps (one 16-byte unit of the tile per thread) + the
// producer warp; the merge team (if the variant has one) follows them
__device__ __forceinline__ uint32_t gather_threads(const QueryParams &P) { return (((P.tile_bytes + 511u) >> 9) + 1u) * 32u; }

struct SoloGeom {
    uint64_t begin;     // first k-mer of this CTA's contiguous range
    uint32_t cnt;       // k-mers in the range (hashed by this CTA)
    uint32_t n_static;  // the first n_static are gathered by this CTA itself
    uint32_t n_pool;    // the last n_pool go to the shared pool
    uint32_t n_first;   // static k-mers the producer warp hashes itself (one ring-full)
    uint64_t total;     // k-mers of the query: P.total_kmers, or *P.total_dev when a preceding kernel determines it
    // sequence front-end: the CTA's staged span of the sequence and the list of its unique windows (offsets into it)
    const uint8_t *span;
    const uint16_t *ulist;
};
__device__ __forceinline__ uint32_t solo_range_cnt(const QueryParams &P, uint32_t cta, uint64_t total)
{
    const uint64_t b = (uint64_t)cta * P.items_per_slice;
    return b >= total ? 0u : (uint32_t)min((uint64_t)P.items_per_slice, total - b);
}
// seq_cnt: sequence front-end -- the number of unique windows this CTA represents (its k-mers are listed in shared memory)
__device__ __forceinline__ SoloGeom solo_geometry(const QueryParams &P, uint32_t seq_cnt)
{
    SoloGeom g;
    // the geometry was planned for P.total_kmers (an upper bound when total_dev is set): fewer k-mers only
    // leave the ranges of the last CTAs short or empty
    g.total = (P.total_dev && !P.seq_mode) ? min((uint64_t)__ldcg(P.total_dev), P.total_kmers) : P.total_kmers;
    g.begin = (uint64_t)blockIdx.x * P.items_per_slice;
    g.cnt = P.seq_mode ? seq_cnt : solo_range_cnt(P, blockIdx.x, g.total);
    g.n_pool = min(P.pool_share, g.cnt);
    g.n_static = g.cnt - g.n_pool;
    g.n_first = min(g.n_static, P.n_stages * P.kmers_per_stage);
    return g;
}

// Query broadcast (rank 0 of a column-sharded search): this CTA's slice of the k-mer bytes (the 16-byte lines that cover
// it; neighbouring CTAs write identical lines at the boundaries) goes into every peer's LL inbox over NVLink -- plain
// stores with the flag embedded: no fence (which would also wait for the bulk copies in flight), nothing to wait
// for.  The inbox was last used by query seq - kStreamRing, which every shard has finished (the entry gate).
// Executed by `nthreads` threads (tid = 0 .. nthreads-1): the merge team, which is idle at this point of the
// CTA's life, or -- in the kernel variants without one -- the producer warp once its first ring-full is in flight.
__device__ __forceinline__ void push_query_slice(const QueryParams &P, uint64_t begin, uint32_t cnt, uint32_t tid, uint32_t nthreads)
{
    const uint64_t b0 = begin * P.k, b1 = (begin + cnt) * P.k;
    const uint64_t l0 = b0 >> 4, nvec = ((b1 + 15) >> 4) - l0;
    const uint4 *src = reinterpret_cast<const uint4 *>(P.kmers) + l0;
    constexpr int kBatch = 6;  // loads in flight per thread: a 68-k-mer slice (133 lines) is one batch of a warp
    for (uint64_t i0 = tid; i0 < nvec; i0 += (uint64_t)nthreads * kBatch) {
        uint4 v[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u)
            if (i0 + (uint64_t)nthreads * u < nvec) v[u] = __ldg(src + i0 + (uint64_t)nthreads * u);
#pragma unroll
        for (int u = 0; u < kBatch; ++u)
            if (i0 + (uint64_t)nthreads * u < nvec)
                for (uint32_t rep = 0; rep <= P.push_repeat; ++rep)  // (push_repeat = 0 outside diagnostics)
                    for (uint32_t r = 0; r < P.n_push; ++r)
                        ll_store_line(P.ll.out[r] + 2 * (l0 + i0 + (uint64_t)nthreads * u), v[u], P.ll.flag);
    }
}

template <bool PUSH>
__device__ __forceinline__ void solo_producer(const QueryParams &P, const SoloGeom &sg, uint8_t *smem, uint8_t *ring,
                                              int32_t *ids, uint8_t *scratch, uint64_t *full, uint64_t *empty)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t h = P.h, G = P.kmers_per_stage;
    const uint32_t tw = min(P.tile_bytes, P.row_bytes16);
    const uint32_t seg_stride = P.tile_bytes;
    const uint32_t stage_bytes = G * h * seg_stride;
    const uint64_t policy = policy_evict_first();
    volatile uint32_t *stage_cnt = reinterpret_cast<volatile uint32_t *>(smem + kStageCntOffset);
    uint32_t stage = 0, parity = 0;
    auto advance = [&]() {
        if (++stage == P.n_stages) {
            stage = 0;
            parity ^= 1;
        }
    };
    // one ring slot: nk k-mers whose row ids are id[0 .. nk*h), read through `load`
    auto issue = [&](uint32_t nk, auto load) {
        const uint32_t n_rows = nk * h;
        uint8_t *dst0 = ring + (size_t)stage * stage_bytes;
        mbar_wait(&empty[stage], parity ^ 1);  // slot drained by all consumer warps
        if (lane == 0) {
            stage_cnt[stage] = nk;
            mbar_arrive_expect_tx(&full[stage], n_rows * tw);
        }
        __syncwarp();
        for (uint32_t i = lane; i < n_rows; i += 32)
            bulk_g2s_hint(dst0 + (size_t)i * seg_stride, P.matrix + (uint64_t)(uint32_t)load(i) * P.pitch, tw, &full[stage],
                          policy);
        advance();
    };

    // the first ring-full of k-mers is hashed by this warp alone, so the gather starts at once
    if (P.seq_mode)
        hash_window_list(sg.span, sg.ulist, sg.n_first, (int)P.k, (int)h, P.num_rows, P.mod_magic, ids, lane, 32u);
    else
        hash_kmers_group(P.kmers + sg.begin * P.k, sg.n_first, (int)P.k, (int)h, P.num_rows, 1, scratch, ids, lane, 32u,
                         SyncWarp(), P.mod_magic, &P.ll);
    __syncwarp();
    if (lane == 0) BIGSI_TS(1);
    uint32_t k0 = 0;
    for (; k0 < sg.n_first; k0 += G) {
        const int32_t *id = ids + (size_t)k0 * h;
        issue(min(G, sg.n_first - k0), [&](uint32_t i) { return id[i]; });
    }
    if (PUSH && P.n_push && sg.cnt) push_query_slice(P, sg.begin, sg.cnt, lane, 32u);  // (variants without a merge team)
    named_bar_sync<kBarIdsReady>(gather_threads(P));  // the consumer warps have hashed the rest of the range
    for (; k0 < sg.n_static; k0 += G) {
        const int32_t *id = ids + (size_t)k0 * h;
        issue(min(G, sg.n_static - k0), [&](uint32_t i) { return id[i]; });
    }

    // pool: claim kPoolBatch indices with one atomic (index -> owner CTA round-robin), resolve them lane
    // parallel (owner's ready flag, row ids), then gather them one k-mer per ring slot.  At most `cap`
    // pooled k-mers per CTA so that the segment counter cannot overflow; every claimed index below
    // pool_total IS processed (a claim is only made with capacity reserved for all of it).
    const uint32_t pool_total = P.pool_share * gridDim.x;
    const uint32_t cap = P.solo_max_kmers - sg.n_static;
    int32_t *stash = reinterpret_cast<int32_t *>(smem + kPoolStashOffset);  // [kPoolBatch][h <= kPoolMaxH]
    uint32_t taken = 0;
    bool exhausted = pool_total == 0;
    while (!exhausted && taken < cap) {
        const uint32_t n = min((uint32_t)kPoolBatch, cap - taken);
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(P.pool_counter, n);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= pool_total) break;
        const uint32_t idx = base + lane;
        uint32_t owner = 0, j = 0;
        bool real = false;
        if (lane < n && idx < pool_total) {
            owner = idx % gridDim.x;
            j = idx / gridDim.x;
            // the owner's ready word = launch epoch (high half) | number of k-mers it pooled (short ranges pool fewer)
            unsigned long long word = 0;
            bounded_wait(P.abort_word, P.host_abort, P.spin_timeout_ns, kAbortPool, P.stream_seq, [&]() {
                word = ld_acquire_gpu_u64(P.pool_ready + owner);
                return (uint32_t)(word >> 32) == (uint32_t)P.pool_epoch;
            });
            real = j < (uint32_t)word;
        }
        if (base + n >= pool_total) exhausted = true;
        const uint32_t real_mask = __ballot_sync(0xffffffffu, real);
        for (uint32_t t0 = 0; t0 < n * h; t0 += 32) {  // warp-uniform trip count (shuffles inside)
            const uint32_t t = t0 + lane;
            const bool in = t < n * h;
            const uint32_t e = in ? t / h : 0, r = t - e * h;
            const uint32_t oe = __shfl_sync(0xffffffffu, owner, e), je = __shfl_sync(0xffffffffu, j, e);
            if (in && ((real_mask >> e) & 1u)) stash[t] = __ldcg(P.pool_ids + ((size_t)oe * P.pool_share + je) * h + r);
        }
        __syncwarp();
        for (uint32_t e = 0; e < n; ++e) {
            if (!((real_mask >> e) & 1u)) continue;
            const int32_t *id = stash + e * h;
            issue(1, [&](uint32_t i) { return id[i]; });
        }
        __syncwarp();  // the stash is reused by the next batch
        taken += __popc(real_mask);
    }
    // end marker: an empty slot
    mbar_wait(&empty[stage], parity ^ 1);
    if (lane == 0) {
        stage_cnt[stage] = 0;
        mbar_arrive(&full[stage]);
        BIGSI_TS(5);
    }
}

template <int MODE, int HC, int NP>
__device__ __forceinline__ void solo_consumer(const QueryParams &P, const SoloGeom &sg, uint8_t *smem, const uint8_t *ring,
                                              int32_t *ids, uint8_t *scratch, uint64_t *full, uint64_t *empty)
{
    const uint32_t unit = threadIdx.x, lane = threadIdx.x & 31;
    const uint32_t consumer_threads = gather_threads(P) - 32;
    const uint32_t h = HC ? HC : P.h, G = P.kmers_per_stage;
    const uint32_t seg_stride = P.tile_bytes;
    const uint32_t stage_bytes = G * h * seg_stride;
    const uint32_t tw = min(P.tile_bytes, P.row_bytes16);
    volatile uint32_t *stage_cnt = reinterpret_cast<volatile uint32_t *>(smem + kStageCntOffset);

    // hash the part of the range the producer did not take, publish the pooled ids, release the producer
    const uint32_t rest = sg.cnt - sg.n_first;
    if (P.seq_mode)
        hash_window_list(sg.span, sg.ulist + sg.n_first, rest, (int)P.k, (int)P.h, P.num_rows, P.mod_magic,
                         ids + (size_t)sg.n_first * P.h, unit, consumer_threads);
    else
        hash_kmers_group(P.kmers + (sg.begin + sg.n_first) * P.k, rest, (int)P.k, (int)P.h, P.num_rows, 1,
                         scratch + ((hash_scratch_bytes(sg.n_first, P.k) + 127) & ~127ull), ids + (size_t)sg.n_first * P.h, unit,
                         consumer_threads, GroupSync{kBarHashGroup, (int)consumer_threads}, P.mod_magic, &P.ll);
    named_bar_sync<kBarHashGroup>(consumer_threads);
    if (P.pool_share) {
        int32_t *dst = P.pool_ids + (size_t)blockIdx.x * P.pool_share * P.h;
        const int32_t *src = ids + (size_t)sg.n_static * P.h;
        for (uint32_t i = unit; i < sg.n_pool * P.h; i += consumer_threads) dst[i] = src[i];
        named_bar_sync<kBarHashGroup>(consumer_threads);
        if (unit == 0) {  // one cumulative fence behind the barrier publishes every thread's ids
            __threadfence();
            st_release_gpu_u64(P.pool_ready + blockIdx.x, (P.pool_epoch << 32) | (unsigned long long)sg.n_pool);
        }
    }
    named_bar_arrive<kBarIdsReady>(gather_threads(P));
    if (unit == 0) BIGSI_TS(9);

    const bool active = unit * 16 < tw;
    const uint32_t nhi = P.planes_per_slot > 3 ? P.planes_per_slot - 3 : 0;
    VCounter<NP> ctr;
    W4 acc;
    if (MODE == kModeCounts) ctr.reset();
    else acc = W4{{0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu}};
    uint32_t stage = 0, parity = 0;
    bool first_slot = true;
    for (;;) {
        mbar_wait(&full[stage], parity);
        if (first_slot && unit == 0) BIGSI_TS(2);
        first_slot = false;
        const uint32_t gn = stage_cnt[stage];
        if (gn == 0) break;  // end marker
        if (active && !(P.debug_flags & 1u)) {
            const uint8_t *base = ring + (size_t)stage * stage_bytes + unit * 16;
            for (uint32_t g = 0; g < gn; ++g) {
                const uint8_t *b = base + g * h * seg_stride;
                W4 x = to_w4(lds128(b));
                if (HC == 3) {
                    const W4 y = to_w4(lds128(b + seg_stride));
                    const W4 z = to_w4(lds128(b + 2 * seg_stride));
#pragma unroll
                    for (int j = 0; j < 4; ++j) x.v[j] = and3(x.v[j], y.v[j], z.v[j]);
                } else {
                    for (uint32_t r = 1; r < h; ++r) {
                        const W4 y = to_w4(lds128(b + r * seg_stride));
#pragma unroll
                        for (int j = 0; j < 4; ++j) x.v[j] &= y.v[j];
                    }
                }
                if (MODE == kModeCounts) {
                    ctr.add(x, nhi);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc.v[j] &= x.v[j];
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == P.n_stages) {
            stage = 0;
            parity ^= 1;
        }
    }
    if (unit == 0) BIGSI_TS(3);
    if (active) {
        const uint32_t chunk = unit * 16 / P.merge_cb;
        uint8_t *dst = P.partial + partial_offset(P, chunk, blockIdx.x, 0) + (unit * 16 - chunk * P.merge_cb);
        if (MODE == kModeCounts) {
            ctr.finish(nhi);
            flush_planes(ctr, P.planes_per_slot, dst, P.merge_cb);
        } else {
            stg128(dst, acc);
        }
    }
    if (unit == 0) BIGSI_TS(4);
}

template <int MODE, int HC, int NP, bool TEAM>
__global__ void __launch_bounds__(kMaxBlockThreads + (TEAM ? kMergeTeamThreads : 0), 1)
gather_solo(const __grid_constant__ QueryParams P, const __grid_constant__ QueryParams PV)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);
    uint64_t *empty = full + kMaxStages;
    int32_t *ids = reinterpret_cast<int32_t *>(smem + kSmemHeaderBytes);
    uint8_t *ring = smem + kSmemHeaderBytes + P.ids_bytes;
    volatile int *s_gate = reinterpret_cast<volatile int *>(smem + kGateOffset);
    const uint32_t n_gather = gather_threads(P);
    const uint32_t consumer_warps = (n_gather >> 5) - 1;

    // the dependents (the next query's gather kernel, or the flush) may be scheduled as soon as every CTA of this grid
    // has started: they only become resident where an SM is free
    grid_launch_dependents();
    if (threadIdx.x == 0) {
        BIGSI_TS(0);
        if (P.debug_ts) {
            unsigned int smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            P.debug_ts[(size_t)blockIdx.x * kDebugStamps + 15] = smid;
        }
        for (uint32_t s = 0; s < P.n_stages; ++s) {
            mbar_init(&full[s], 1);                // one arrive.expect_tx by the producer + tx bytes
            mbar_init(&empty[s], consumer_warps);  // one arrive per consumer warp
        }
        if (TEAM) mbar_init(reinterpret_cast<uint64_t *>(smem + kTeamBarOffset), 1);
        fence_barrier_init();
    }
    if (threadIdx.x == 32) {
        // entry gate: the buffers of this query's ring slot were last used by query seq - kStreamRing, whose
        // stage 2 must have finished (it also cleared our state block); a handle that has aborted stays dead.
        // Both words are requested before either is looked at: one L2 round trip, not two.
        const bool gated = P.stream_seq > (unsigned long long)kStreamRing;
        const unsigned long long done = gated ? ld_acquire_gpu_u64(P.stream_done) : 0ull;
        bool ok = ld_volatile_u64(P.abort_word) == 0ull;
        if (ok && gated && done + kStreamRing < P.stream_seq)
            ok = bounded_wait(P.abort_word, P.host_abort, P.spin_timeout_ns, kAbortGate, P.stream_seq,
                              [&]() { return ld_acquire_gpu_u64(P.stream_done) + kStreamRing >= P.stream_seq; });
        *s_gate = ok ? 1 : 0;
    } else if (threadIdx.x >= 64 && threadIdx.x < 96 && !P.stream_wait_inputs && !P.seq_mode && P.ll.in == nullptr) {
        // this CTA's k-mer bytes on their way into L2 while the gate is read (they are host- or copy-written: in HBM)
        const uint32_t cnt = solo_range_cnt(P, blockIdx.x, P.total_kmers);
        const uint8_t *b0 = P.kmers + (uint64_t)blockIdx.x * P.items_per_slice * P.k;
        for (uint32_t off = (threadIdx.x - 64) * 128u; off < cnt * P.k; off += 32u * 128u)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(b0 + off));
    }
    __syncthreads();
    if (!*s_gate) return;

    if (TEAM && threadIdx.x >= n_gather) {
        // ---- merge team.  First the query broadcast of a column-sharded search (rank 0): the team has nothing else to
        // do yet, and the peers' hashing waits for these lines.  Then stage 2 of the PREVIOUS streamed query, once its
        // gather kernel (the preceding grid) is complete and its planes are visible.  Its scratch lies behind the ring.
        if (P.n_push) {
            if (P.stream_wait_inputs) grid_dependency_wait();
            // The team's warp 0 does NOT push: it executes the fences of the merge (staging, arrival, publication), and a
            // fence waits for the warp's outstanding stores -- with 7 peers ~500 remote stores per warp over NVLink, which
            // held rank 0's teams back by 8 us per query on 8 GPUs.  Warps 1-3 carry the push; they never fence.
            const uint32_t cnt = solo_range_cnt(P, blockIdx.x, P.total_kmers);
            const uint32_t tt = threadIdx.x - n_gather;
            if (cnt) {
                if (P.push_all_warps) push_query_slice(P, (uint64_t)blockIdx.x * P.items_per_slice, cnt, tt, (uint32_t)kMergeTeamThreads);
                else if (tt >= 32) push_query_slice(P, (uint64_t)blockIdx.x * P.items_per_slice, cnt, tt - 32, (uint32_t)kMergeTeamThreads - 32);
            }
            if (threadIdx.x == n_gather + 32) BIGSI_TS(6);
        }
        const WarpGroupTeam<kBarMergeTeam> T{n_gather, (uint32_t)kMergeTeamThreads};
        uint8_t *team_smem = ring + (size_t)P.n_stages * P.kmers_per_stage * P.h * P.tile_bytes;
        uint64_t *team_bar = reinterpret_cast<uint64_t *>(smem + kTeamBarOffset);
        volatile int *team_flag = reinterpret_cast<volatile int *>(smem + kTeamFlagOffset);
        if (P.merge_prev) {
            grid_dependency_wait();
            if (T.tid() == 0 && PV.debug_ts) PV.debug_ts[(size_t)blockIdx.x * kDebugStamps + 8] = debug_gtime();
            reduce_query<kModeCounts>(PV, team_smem, team_bar, team_flag, T, blockIdx.x, gridDim.x);
            if (threadIdx.x == n_gather) BIGSI_TS(7);
        }
        if (P.self_merge) {
            // isolated query, cooperative launch (every CTA of the grid is resident): stage 2 of THIS query as soon as all
            // gather CTAs have flushed their planes -- no flush kernel, no second launch on the critical path
            if (T.tid() == 0)
                *team_flag = bounded_wait(P.abort_word, P.host_abort, P.spin_timeout_ns, kAbortExit, P.stream_seq,
                                          [&]() { return ld_acquire_gpu_u32(&P.qstate->gather_arrivals) >= gridDim.x; })
                                 ? 1 : 0;
            T.sync();
            if (!*team_flag) return;
            __threadfence();
            T.sync();
            reduce_query<kModeCounts>(P, team_smem, team_bar, team_flag, T, blockIdx.x, gridDim.x);
        }
        return;
    }

    // the k-mers may come out of the preceding kernel of the stream (query front-end, a caller's kernel): wait for
    // it.  Otherwise nothing the gather warps read or write depends on their predecessor.
    if (P.stream_wait_inputs) grid_dependency_wait();
    if (threadIdx.x == 0) BIGSI_TS(8);
    if (blockIdx.x == 0 && threadIdx.x == 0 && P.n_hits != nullptr) P.n_hits[0] = 0ull;  // stage 2 adds to it

    uint8_t *scratch = smem + kSmemHeaderBytes + P.ids_table_bytes;
    uint32_t seq_cnt = 0;
    const uint8_t *span = nullptr;
    const uint16_t *ulist = nullptr;
    if (P.seq_mode) {
        // sequence front-end (hash.cuh): this CTA's windows [w0, w0 + cw) of the sequence; all gather threads stage the
        // span with one round trip of 16-byte loads (the sequence may live in mapped host memory), every window
        // tries to claim its string in the table, the winners are listed -- they are this CTA's k-mers
        volatile uint32_t *s_cnt = reinterpret_cast<volatile uint32_t *>(smem + kSeqCountOffset);
        const uint64_t n = P.total_kmers;
        const uint64_t w0 = (uint64_t)blockIdx.x * P.items_per_slice;
        const uint32_t cw = w0 >= n ? 0u : (uint32_t)min((uint64_t)P.items_per_slice, n - w0);
        if (threadIdx.x == 0) *s_cnt = 0;
        const uint8_t *g0 = P.kmers + w0;
        const uint32_t skew = (uint32_t)(reinterpret_cast<uintptr_t>(g0) & 15);
        const uint32_t nvec = cw ? (skew + cw + P.k - 1 + 15) >> 4 : 0u;
        const uint4 *a0 = reinterpret_cast<const uint4 *>(g0 - skew);
        uint4 *sv = reinterpret_cast<uint4 *>(scratch);
        for (uint32_t i = threadIdx.x; i < nvec; i += n_gather) sv[i] = __ldg(a0 + i);
        named_bar_sync<kBarGather>(n_gather);
        span = scratch + skew;
        uint16_t *list = reinterpret_cast<uint16_t *>(scratch + ((P.items_per_slice + P.k + 47u) & ~15u));
        ulist = list;
        const SeqTable T{P.seq_table, P.seq_table_entries - 1, P.seq_epoch, P.kmers};
        const uint32_t lane = threadIdx.x & 31;
        for (uint32_t base = 0; base < cw; base += n_gather) {  // uniform trip count (ballots inside)
            const uint32_t w = base + threadIdx.x;
            const bool win = w < cw && seq_table_insert(T, span + w, (int)P.k, w0 + w, span, w0, cw);
            const uint32_t mask = __ballot_sync(0xffffffffu, win);
            uint32_t pos = 0;
            if (mask) {
                if (lane == 0) pos = atomicAdd(const_cast<uint32_t *>(s_cnt), (uint32_t)__popc(mask));
                pos = __shfl_sync(0xffffffffu, pos, 0);
            }
            if (win) list[pos + __popc(mask & ((1u << lane) - 1u))] = (uint16_t)w;
        }
        named_bar_sync<kBarGather>(n_gather);
        seq_cnt = *s_cnt;
        if (threadIdx.x == 0) {
            if (seq_cnt) atomicAdd(&P.qstate->n_unique, (unsigned long long)seq_cnt);
            BIGSI_TS(10);
        }
    }
    SoloGeom sg = solo_geometry(P, seq_cnt);
    sg.span = span;
    sg.ulist = ulist;
    if ((threadIdx.x >> 5) == consumer_warps)
        solo_producer<!TEAM>(P, sg, smem, ring, ids, scratch, full, empty);
    else
        solo_consumer<MODE, HC, NP>(P, sg, smem, ring, ids, scratch, full, empty);
    if (TEAM && P.self_merge) {
        // every plane store of this CTA before its arrival: barrier over the gather threads, then one release
        named_bar_sync<kBarGather>(n_gather);
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(&P.qstate->gather_arrivals, 1u);
        }
    }
}

// grid-wide barrier over a monotonic arrival counter (all CTAs of the launch are co-resident: the generic kernel
// is launched cooperatively with at most one CTA per SM whenever it merges in the kernel)
__device__ __forceinline__ void grid_barrier(unsigned long long *counter, unsigned long long target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();  // publish this CTA's partial planes
        atomicAdd(counter, 1ull);
        unsigned long long seen;
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(counter) : "memory");
        } while (seen < target);
        __threadfence();
    }
    __syncthreads();
}

// generic path: any number of queries / tiles / slices; row ids or k-mers hashed in the prologue
template <int MODE, int HC>
__global__ void __launch_bounds__(kMaxBlockThreads, 1) fused_query(const __grid_constant__ QueryParams P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);
    uint64_t *empty = full + kMaxStages;
    uint64_t *merge_bar = empty + kMaxStages;  // staging barrier of the merge phase
    int32_t *ids = reinterpret_cast<int32_t *>(smem + kSmemHeaderBytes);
    uint8_t *ring = smem + kSmemHeaderBytes + P.ids_bytes;
    const uint32_t consumer_warps = (blockDim.x >> 5) - 1;

    if (threadIdx.x == 0) {
        BIGSI_TS(0);
        for (uint32_t s = 0; s < P.n_stages; ++s) {
            mbar_init(&full[s], 1);                // one arrive.expect_tx by the producer + tx bytes
            mbar_init(&empty[s], consumer_warps);  // one arrive per consumer warp
        }
        mbar_init(merge_bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    // PDL: everything above overlapped the previous kernel's tail; from here on we read what it
    // produced (k-mers / row ids) and overwrite what the previous query's merge may still read
    grid_launch_dependents();
    grid_dependency_wait();
    if (threadIdx.x == 0) BIGSI_TS(8);

    if (P.n_hits != nullptr && !P.direct_complete) {  // hit counters of the fused threshold (stage 2 adds to them)
        for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < P.n_queries; q += gridDim.x * blockDim.x)
            P.n_hits[q] = 0ull;
    }

    const uint64_t span = (uint64_t)P.slices_per_cta * P.items_per_slice;
    const uint64_t begin = (uint64_t)blockIdx.x * span;
    uint64_t end = begin + span;
    if (end > P.total_items) end = P.total_items;
    const bool have_work = begin < end;

    if (P.prehash && have_work) {
        // n_tiles == 1 and one slice per CTA: this CTA's k-mers are [begin, end), contiguous in memory
        hash_kmers_cooperative(P.kmers + begin * P.k, (uint32_t)(end - begin), (int)P.k, (int)P.h, P.num_rows, 1,
                               smem + kSmemHeaderBytes + P.ids_table_bytes, ids, P.mod_magic);
        __syncthreads();
        if (threadIdx.x == 0) BIGSI_TS(9);
    }
    if (have_work) {
        if ((threadIdx.x >> 5) == consumer_warps)
            tma_producer(P, ring, ids, full, empty, begin, end);
        else
            tma_consumer<MODE, HC>(P, ring, full, empty, begin, end);
    }

    if (P.fuse_merge) {
        // stage 2 in the same kernel: once every CTA has published its planes, the CTAs share the
        // merge work items; the drained ring is the scratch
        grid_barrier(P.barrier, P.barrier_target);
        if (threadIdx.x == 0) BIGSI_TS(6);
        uint32_t merge_phase = 0;
        for (uint64_t item = blockIdx.x; item < P.merge_items; item += gridDim.x)
            merge_item<MODE>(P, item, ring, merge_bar, merge_phase, CtaTeam());
        if (threadIdx.x == 0) BIGSI_TS(7);
        if (P.n_sinks) {
            // the last CTA to finish its merge items publishes query 0's hit list to every sink
            int *s_last = reinterpret_cast<int *>(smem + 2 * kMaxStages * 8 + 64);  // header space behind the mbarriers
            if (threadIdx.x == 0) {
                __threadfence();
                *s_last = atomicAdd(P.done_counter, 1ull) + 1ull == P.done_target;
            }
            __syncthreads();
            if (*s_last) {
                __threadfence();
                publish_hits(P, P.sinks, P.n_sinks, P.sink_seq, CtaTeam());
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <int MODE, int HC>
static cudaError_t launch_one(const QueryParams &p, int grid, cudaStream_t stream, const QueryParams *prev)
{
    if (p.solo) {
        // streamed: plain launch with the programmatic-stream-serialization attribute (the kernel decides itself
        // what it waits for).  COUNTS with planes_per_slot <= 8 (up to 255 k-mers per CTA) takes the variant with
        // the small counter and the merge team; the others leave stage 2 to the flush kernel.
        const dim3 g(grid), b(query_block_threads(p));
        const QueryParams &pv = prev ? *prev : p;
        constexpr int SHC = HC == 3 ? 3 : 0;  // (the streamed kernels exist for h = 3 and for any h)
        if constexpr (MODE == kModeCounts) {
            if (p.planes_per_slot > 8)
                return launch_ex(gather_solo<MODE, SHC, kSegPlanes, false>, g, b, query_smem_bytes(p), stream, /*pdl=*/true,
                                 /*cooperative=*/false, p, pv);
            return launch_ex(gather_solo<MODE, SHC, 8, true>, g, b, query_smem_bytes(p), stream, true, /*cooperative=*/p.self_merge != 0, p, pv);
        } else {
            return launch_ex(gather_solo<MODE, SHC, 8, false>, g, b, query_smem_bytes(p), stream, true, false, p, pv);
        }
    }
    return launch_ex(fused_query<MODE, HC>, dim3(grid), dim3(query_block_threads(p)), query_smem_bytes(p), stream,
                     /*pdl=*/true, /*cooperative=*/p.fuse_merge != 0 && !p.plain_launch, p);
}

cudaError_t launch_query(const QueryParams &p, int mode, int grid, cudaStream_t stream, const QueryParams *prev)
{
    if (p.merge_prev && !(prev && query_has_merge_team(p, mode))) return cudaErrorInvalidValue;
    if ((p.merge_team != 0) != (p.solo && query_has_merge_team(p, mode))) return cudaErrorInvalidValue;
    // h = 3 (the reference's default) and h = 1 (batches counting over once-gathered AND vectors) are compiled in
    if (mode == kModeCounts)
        return p.h == 3   ? launch_one<kModeCounts, 3>(p, grid, stream, prev)
               : p.h == 1 ? launch_one<kModeCounts, 1>(p, grid, stream, prev)
                          : launch_one<kModeCounts, 0>(p, grid, stream, prev);
    return p.h == 3   ? launch_one<kModeAnd, 3>(p, grid, stream, prev)
           : p.h == 1 ? launch_one<kModeAnd, 1>(p, grid, stream, prev)
                      : launch_one<kModeAnd, 0>(p, grid, stream, prev);
}

cudaError_t query_kernels_init()
{
    cudaError_t e;
#define BIGSI_SET_SMEM(K)                                                                   \
    e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);  \
    if (e != cudaSuccess) return e;                                                         \
    e = cudaFuncSetAttribute(K, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
    if (e != cudaSuccess) return e;
    BIGSI_SET_SMEM((fused_query<kModeCounts, 3>))
    BIGSI_SET_SMEM((fused_query<kModeCounts, 0>))
    BIGSI_SET_SMEM((fused_query<kModeAnd, 3>))
    BIGSI_SET_SMEM((fused_query<kModeAnd, 0>))
    BIGSI_SET_SMEM((fused_query<kModeCounts, 1>))
    BIGSI_SET_SMEM((fused_query<kModeAnd, 1>))
    BIGSI_SET_SMEM((gather_solo<kModeCounts, 3, 8, true>))
    BIGSI_SET_SMEM((gather_solo<kModeCounts, 3, kSegPlanes, false>))
    BIGSI_SET_SMEM((gather_solo<kModeCounts, 0, 8, true>))
    BIGSI_SET_SMEM((gather_solo<kModeCounts, 0, kSegPlanes, false>))
    BIGSI_SET_SMEM((gather_solo<kModeAnd, 3, 8, false>))
    BIGSI_SET_SMEM((gather_solo<kModeAnd, 0, 8, false>))
#undef BIGSI_SET_SMEM
    return merge_kernels_init();
}

}  // namespace bigsi
