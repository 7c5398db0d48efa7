/*
 * oracle/bigsi_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the BIGSI search hot path of the reference
 * (Phelimb/BIGSI v0.3.8, /root/reference).  It exists to CHECK the CUDA
 * product path (tests/, __graft_entry__.smoke()) and to serve as the CPU
 * baseline leg of bench.py.  Nothing under bigsi_b200/ may import, link or
 * execute it; the product path has no CPU fallback.
 *
 * Parity pin: every function below is checked against outputs of the
 * UNMODIFIED reference package (imported through oracle/ref_shims) by
 * tests/golden/make_golden.py -> tests/golden/ (JSON), and against the
 * reference's own known-answer tests
 *   bigsi/tests/bloom/test_create_bloomfilter.py:5-8
 *   bigsi/tests/graph/test_index.py:14-105
 *   bigsi/tests/graph/test_end_to_end.py:69-131.
 *
 * Third-party arithmetic restated here because its source is not vendored in
 * the reference: mmh3 2.5.1 (.conda/mmh3/meta.yaml:1-5) = MurmurHash3_x86_32
 * (Austin Appleby, public domain algorithm), called at
 * bigsi/bloom/bloomfilter.py:5-6 as `mmh3.hash(element, seed) % m`, i.e.
 * SIGNED 32-bit result with Python floor-mod.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* MurmurHash3_x86_32 (mmh3.hash)                                     */
/* ------------------------------------------------------------------ */
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

uint32_t oracle_murmur3_x86_32(const uint8_t *key, int len, uint32_t seed)
{
    const int nblocks = len / 4;
    uint32_t h1 = seed;
    const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
    for (int i = 0; i < nblocks; i++) {
        uint32_t k1;
        memcpy(&k1, key + 4 * i, 4); /* little-endian block read */
        k1 *= c1; k1 = rotl32(k1, 15); k1 *= c2;
        h1 ^= k1; h1 = rotl32(h1, 13); h1 = h1 * 5 + 0xe6546b64u;
    }
    const uint8_t *tail = key + 4 * nblocks;
    uint32_t k1 = 0;
    switch (len & 3) {
    case 3: k1 ^= (uint32_t)tail[2] << 16; /* fallthrough */
    case 2: k1 ^= (uint32_t)tail[1] << 8;  /* fallthrough */
    case 1: k1 ^= tail[0];
        k1 *= c1; k1 = rotl32(k1, 15); k1 *= c2; h1 ^= k1;
    }
    h1 ^= (uint32_t)len;
    h1 ^= h1 >> 16; h1 *= 0x85ebca6bu; h1 ^= h1 >> 13; h1 *= 0xc2b2ae35u; h1 ^= h1 >> 16;
    return h1;
}

/* bigsi/bloom/bloomfilter.py:5-6  _hash(): mmh3.hash(element, seed) % m
 * (signed int32, Python floor-mod -> result in [0, m)). */
int64_t oracle_hash_row(const uint8_t *key, int len, uint32_t seed, int64_t m)
{
    int64_t s = (int32_t)oracle_murmur3_x86_32(key, len, seed);
    int64_t r = s % m;
    if (r < 0) r += m;
    return r;
}

/* ------------------------------------------------------------------ */
/* canonical k-mer: bigsi/utils/fncts.py:12,38-39,51-54               */
/* reverse_comp maps only A<->T, C<->G, every other byte passes       */
/* through; canonical = sorted([k, revcomp(k)])[0] (byte order for    */
/* ASCII input).                                                      */
/* ------------------------------------------------------------------ */
static inline uint8_t comp_base(uint8_t b)
{
    switch (b) {
    case 'A': return 'T';
    case 'T': return 'A';
    case 'C': return 'G';
    case 'G': return 'C';
    default:  return b;
    }
}

void oracle_canonical(const uint8_t *kmer, int k, uint8_t *out)
{
    /* compare kmer with its reverse complement lexicographically */
    int cmp = 0;
    for (int i = 0; i < k && cmp == 0; i++) {
        uint8_t a = kmer[i], b = comp_base(kmer[k - 1 - i]);
        cmp = (a < b) ? -1 : (a > b) ? 1 : 0;
    }
    if (cmp <= 0) {
        memcpy(out, kmer, (size_t)k);
    } else {
        for (int i = 0; i < k; i++) out[i] = comp_base(kmer[k - 1 - i]);
    }
}

/* bigsi/graph/index.py:62-70 __kmers_to_hashes: canonical k-mer ->
 * generate_hashes (bloomfilter.py:9-13), seeds 0..h-1.  Output keeps all h
 * slots (duplicates inside a k-mer are harmless: AND is idempotent).
 * kmers: n*k raw ASCII bytes. rows_out: n*h. */
void oracle_hash_kmers(const uint8_t *kmers, int64_t n, int k, int h, int64_t m,
                       int32_t *rows_out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        uint8_t canon[256];
        uint8_t *buf = canon;
        uint8_t *heap = NULL;
        if (k > 256) { heap = (uint8_t *)malloc((size_t)k); buf = heap; }
        oracle_canonical(kmers + i * (int64_t)k, k, buf);
        for (int s = 0; s < h; s++)
            rows_out[i * h + s] = (int32_t)oracle_hash_row(buf, k, (uint32_t)s, m);
        free(heap);
    }
}

/* ------------------------------------------------------------------ */
/* Row store view: `store` holds R rows of `stride` bytes each in the  */
/* reference's byte layout (bitarray.tobytes(): column c -> byte c>>3, */
/* mask 0x80 >> (c&7); storage/base.py:86-99).  slot[i*h+j] = which   */
/* stored row k-mer i's j-th hash refers to.                          */
/* ------------------------------------------------------------------ */

/* bigsi/graph/index.py:75-80 __bitwise_and_kmers: per unique k-mer, AND of
 * its (<=h) rows (utils/fncts.py:24-25).  out: n * row_bytes. */
void oracle_and_per_kmer(const uint8_t *store, int64_t stride, int64_t row_bytes,
                         const int64_t *slot, int64_t n, int h, uint8_t *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        uint8_t *o = out + i * row_bytes;
        memcpy(o, store + slot[i * h] * stride, (size_t)row_bytes);
        for (int j = 1; j < h; j++) {
            const uint8_t *r = store + slot[i * h + j] * stride;
            for (int64_t b = 0; b < row_bytes; b++) o[b] &= r[b];
        }
    }
}

/* bigsi/graph/bigsi.py:192-195 exact_filter: AND over all per-k-mer vectors.
 * presence_out: row_bytes (MSB-first).  n must be >= 1 (the reference raises
 * TypeError on an empty reduce). */
void oracle_exact(const uint8_t *store, int64_t stride, int64_t row_bytes,
                  const int64_t *slot, int64_t n, int h, uint8_t *presence_out)
{
    memset(presence_out, 0xff, (size_t)row_bytes);
    const int64_t CH = 1024; /* column-chunk parallel: race free, row order kept */
    const int64_t nch = (row_bytes + CH - 1) / CH;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t ch = 0; ch < nch; ch++) {
        const int64_t b0 = ch * CH;
        const int64_t b1 = (b0 + CH < row_bytes) ? b0 + CH : row_bytes;
        for (int64_t i = 0; i < n * h; i++) {
            const uint8_t *r = store + slot[i] * stride;
            for (int64_t b = b0; b < b1; b++) presence_out[b] &= r[b];
        }
    }
}

/* bigsi/graph/bigsi.py:35-44 unpack_and_sum: per-column count over the n
 * per-k-mer AND vectors.  counts_out: 8*row_bytes int32 (the reference's
 * vector has 8*ceil(N/8) entries; zip() at bigsi.py:215-216 truncates to N).
 * Column-range parallel so the threaded version is race free. */
void oracle_counts(const uint8_t *store, int64_t stride, int64_t row_bytes,
                   const int64_t *slot, int64_t n, int h, int32_t *counts_out)
{
    memset(counts_out, 0, (size_t)row_bytes * 8 * sizeof(int32_t));
    const int64_t CH = 512; /* bytes per column chunk */
    const int64_t nch = (row_bytes + CH - 1) / CH;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t ch = 0; ch < nch; ch++) {
        const int64_t b0 = ch * CH;
        const int64_t b1 = (b0 + CH < row_bytes) ? b0 + CH : row_bytes;
        uint8_t acc[512];
        for (int64_t i = 0; i < n; i++) {
            const uint8_t *r0 = store + slot[i * h] * stride;
            for (int64_t b = b0; b < b1; b++) acc[b - b0] = r0[b];
            for (int j = 1; j < h; j++) {
                const uint8_t *r = store + slot[i * h + j] * stride;
                for (int64_t b = b0; b < b1; b++) acc[b - b0] &= r[b];
            }
            for (int64_t b = b0; b < b1; b++) {
                uint8_t v = acc[b - b0];
                int32_t *c = counts_out + b * 8;
                c[0] += (v >> 7) & 1; c[1] += (v >> 6) & 1;
                c[2] += (v >> 5) & 1; c[3] += (v >> 4) & 1;
                c[4] += (v >> 3) & 1; c[5] += (v >> 2) & 1;
                c[6] += (v >> 1) & 1; c[7] += v & 1;
            }
        }
    }
}

/* ------------------------------------------------------------------ */
/* Synthetic index generator (no reference analogue -- SURVEY.md K7). */
/* The device fill kernel (bigsi_b200/csrc) implements the SAME pure   */
/* function of (seed,row,global column); this copy lets the CPU       */
/* regenerate only the rows a query touches.  Spec in DESIGN.md.       */
/* ------------------------------------------------------------------ */
static inline uint64_t mix64(uint64_t z)
{
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static inline uint64_t synth_row_key(uint64_t seed, uint64_t row)
{
    return mix64(seed ^ (row * 0xd1b54a32d192ed03ull));
}
/* 64-bit word covering global bytes [8W, 8W+8), little-endian byte order */
static inline uint64_t synth_word(uint64_t rk, uint64_t W, int and_draws)
{
    uint64_t w = ~0ull;
    for (int i = 0; i < and_draws; i++)
        w &= mix64(rk + (W * 4 + (uint64_t)i) * 0x9e3779b97f4a7c15ull);
    return w;
}
static inline uint32_t synth_plant_u32(uint64_t rk, uint64_t col)
{
    return (uint32_t)(mix64(rk ^ (col * 0xc2b2ae3d27d4eb4full + 0x165667b19e3779f9ull)) >> 32);
}

/* One row slice: global columns [col_offset, col_offset + 8*row_bytes),
 * col_offset % 8 == 0; columns >= col_offset + num_cols are zero padding.
 * planted columns are GLOBAL column ids with a u32 threshold: bit = 1 when
 * thr == 0xffffffff, else (hash32(row,col) < thr); the planted value REPLACES
 * the base bit. */
void oracle_synth_row(uint64_t seed, int and_draws, uint64_t row,
                      uint64_t col_offset, uint64_t num_cols, int64_t row_bytes,
                      const uint64_t *planted_cols, const uint32_t *planted_thr,
                      int n_planted, uint8_t *out)
{
    const uint64_t rk = synth_row_key(seed, row);
    const uint64_t gb0 = col_offset >> 3;
    uint64_t curW = ~0ull, cur = 0;
    for (int64_t b = 0; b < row_bytes; b++) {
        uint64_t gb = gb0 + (uint64_t)b;
        uint64_t W = gb >> 3;
        if (W != curW) { curW = W; cur = synth_word(rk, W, and_draws); }
        out[b] = (uint8_t)(cur >> (8 * (gb & 7)));
    }
    for (int p = 0; p < n_planted; p++) {
        uint64_t c = planted_cols[p];
        if (c < col_offset || c >= col_offset + num_cols) continue;
        uint64_t lc = c - col_offset;
        uint8_t mask = (uint8_t)(0x80u >> (lc & 7));
        int bit = (planted_thr[p] == 0xffffffffu) ? 1 : (synth_plant_u32(rk, c) < planted_thr[p]);
        if (bit) out[lc >> 3] |= mask; else out[lc >> 3] &= (uint8_t)~mask;
    }
    /* zero the padding columns */
    for (uint64_t lc = num_cols; lc < (uint64_t)row_bytes * 8; lc++)
        out[lc >> 3] &= (uint8_t)~(0x80u >> (lc & 7));
}

void oracle_synth_rows(uint64_t seed, int and_draws, const int64_t *rows, int64_t n_rows,
                       uint64_t col_offset, uint64_t num_cols, int64_t row_bytes, int64_t stride,
                       const uint64_t *planted_cols, const uint32_t *planted_thr,
                       int n_planted, uint8_t *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_rows; i++) {
        oracle_synth_row(seed, and_draws, (uint64_t)rows[i], col_offset, num_cols, row_bytes,
                         planted_cols, planted_thr, n_planted, out + i * stride);
        if (stride > row_bytes) memset(out + i * stride + row_bytes, 0, (size_t)(stride - row_bytes));
    }
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
