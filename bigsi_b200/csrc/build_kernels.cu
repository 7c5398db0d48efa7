// Kernels of the rows next to the search path: Bloom-filter construction,
// the N x m -> m x N bit transpose of BIGSI.build, and the per-window presence bits of score=True.
#include "hash.cuh"
#include "merge.cuh"  // warp_transpose32
#include "ptx.cuh"
#include "query.cuh"

namespace bigsi {

// storage order <-> index order inside a 32-bit word loaded little-endian from MSB-first bytes
// (storage/base.py:86-99): element 8b+i lives in byte b at bit 7-i.  The map is an involution.
__device__ __forceinline__ uint32_t msb_first_swizzle(uint32_t w) { return __byte_perm(__brev(w), 0, 0x0123); }

// ------------------------------------------------------------------------------------------
// K9: BloomFilter.update (bigsi/bloom/bloomfilter.py:16-32): set bit `row` of an m-bit filter kept in
// the reference's MSB-first byte layout, for every row id produced by the hash kernel.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bloom_set_bits_kernel(const int32_t *__restrict__ rows, uint64_t n,
                                                             uint32_t *__restrict__ bloom_words)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t r = (uint32_t)__ldg(rows + i);
    atomicOr(bloom_words + (r >> 5), 1u << (8 * ((r >> 3) & 3) + 7 - (r & 7)));
}

cudaError_t launch_bloom_set_bits(const int32_t *d_rows, uint64_t n, uint8_t *d_bloom, cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    const uint64_t blocks = (n + 255) / 256;
    if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    bloom_set_bits_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_rows, n, reinterpret_cast<uint32_t *>(d_bloom));
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K10: BIGSI.build's transpose (bigsi/matrix/transpose.py:33-50 -> graph/index.py:27-40 ->
// matrix/bitmatrix.py:19-25) and bulk insert (matrix/bitmatrix.py:67-75): n Bloom filters of n_bits
// bits (MSB-first bytes, filter c at blooms + c * bloom_stride, stride a multiple of 32 bytes and
// at least ceil(num_rows / 256) * 32 so that every 32-byte line read below is inside the buffer)
// become the columns [col0, col0 + n) of the row-major matrix.
//
// One CTA = 256 rows x 256 columns, one warp = 256 rows x one aligned 32-column word of the matrix.
// Lane l reads 32 bytes (256 rows) of filter l as two 16-byte loads, each of the eight 32x32 bit
// blocks is transposed across the warp (5 shuffle steps) and lane j stores the 32-bit column word
// of row j; the eight warps of a CTA write neighbouring words of the same rows, so L2 merges them
// into full 32-byte sectors.  Words that are only partly inside [col0, col0 + n) keep their other
// columns (read-modify-write by the one lane that owns the word).  Rows >= n_bits get 0.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_blooms_kernel(uint8_t *__restrict__ matrix, uint64_t pitch, uint64_t num_rows,
                                                               uint64_t col0, uint64_t n_blooms,
                                                               const uint8_t *__restrict__ blooms, uint64_t bloom_stride,
                                                               uint64_t n_bits)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t word = (col0 >> 5) + (uint64_t)blockIdx.y * 8 + warp;  // 32-column word of the row
    const uint64_t col = word * 32 + lane;
    const bool valid = col >= col0 && col < col0 + n_blooms;
    const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
    if (vmask == 0) return;  // warp-uniform
    const uint64_t r0 = (uint64_t)blockIdx.x * 256;
    uint32_t w[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) w[s] = 0;
    if (valid) {
        const uint8_t *src = blooms + (col - col0) * bloom_stride + (r0 >> 3);
        const uint4 a = ldg128_stream(src), b = ldg128_stream(src + 16);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
        w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    }
    const uint32_t smask = msb_first_swizzle(vmask);  // the valid columns in storage order
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        const uint64_t rb = r0 + 32 * s;  // first row of this 32 x 32 block
        uint32_t x = msb_first_swizzle(w[s]);  // bit j <-> row rb + j of this lane's filter
        if (rb >= n_bits) x = 0;
        else if (n_bits - rb < 32) x &= (1u << (uint32_t)(n_bits - rb)) - 1u;
        const uint32_t y = msb_first_swizzle(warp_transpose32(x, lane));  // lane j: columns of row rb + j
        const uint64_t row = rb + lane;
        if (row < num_rows) {
            uint32_t *dst = reinterpret_cast<uint32_t *>(matrix + row * pitch + word * 4);
            *dst = vmask == 0xffffffffu ? y : ((*dst & ~smask) | (y & smask));
        }
    }
}

// Wide variant: one warp = 256 rows x EIGHT aligned 32-column words (256 columns), one CTA = 256 rows x 2 048
// columns.  Lane l holds 32 bytes (one 256-bit load: a whole sector) of the eight filters l, l+32, ..., l+224 of its
// warp's column range; per 32-row block the eight 32x32 transposes leave lane j with the 256 columns of row j,
// written as ONE 32-byte store: a full sector.  (With four words per warp the 16-byte stores were partial-sector
// writes and L2 filled every sector from DRAM first: ncu showed 3.4 GB read for 1.28 GB of filters.)  The pitch is a
// multiple of 128 bytes and the first word a multiple of 8, so the address is 32-byte aligned; the filters' stride
// is a multiple of 32.  Words only partly inside [col0, col0 + n) fall back to the per-word read-modify-write.
constexpr int kTWords = 8;

// 32x32 bit-matrix transpose across a warp in 3 instructions per step (rotate, shuffle, select): the SENDER rotates its
// word so that the bits its partner wants sit where they will be stored (partner = lane ^ s: a lane with bit s clear
// sends its word rotated right by s, the other one rotated left), the receiver selects with a per-lane mask.  Rotate
// amounts and masks are per-lane constants of the five steps.  Same result as merge.cuh:warp_transpose32.
struct Transpose32 {
    uint32_t rot[5], msk[5];
    __device__ __forceinline__ explicit Transpose32(uint32_t lane)
    {
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const uint32_t s = 16u >> i;
            const uint32_t m = i == 0 ? 0x0000ffffu : i == 1 ? 0x00ff00ffu : i == 2 ? 0x0f0f0f0fu : i == 3 ? 0x33333333u : 0x55555555u;
            const bool hi = (lane & s) != 0;
            rot[i] = hi ? s : 32u - s;  // rotate-left amount of what this lane SENDS
            msk[i] = hi ? m : ~m;       // bits this lane TAKES from what it receives
        }
    }
    __device__ __forceinline__ uint32_t operator()(uint32_t x) const
    {
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const uint32_t z = __shfl_xor_sync(0xffffffffu, __funnelshift_l(x, x, rot[i]), 16 >> i);
            x = lop3<0xCA>(msk[i], z, x);  // msk ? z : x
        }
        return x;
    }
};
__global__ void __launch_bounds__(256, 3) transpose_blooms_wide_kernel(uint8_t *__restrict__ matrix, uint64_t pitch,
                                                                    uint64_t num_rows, uint64_t col0, uint64_t n_blooms,
                                                                    const uint8_t *__restrict__ blooms, uint64_t bloom_stride,
                                                                    uint64_t n_bits)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t word0 = ((col0 >> 5) & ~(uint64_t)(kTWords - 1)) + ((uint64_t)blockIdx.y * 8 + warp) * kTWords;  // first of this warp's words
    const uint64_t r0 = (uint64_t)blockIdx.x * 256;
    const Transpose32 transpose(lane);
    uint32_t w[kTWords][8], vmask[kTWords];
    bool any = false, all = true;
#pragma unroll
    for (int g = 0; g < kTWords; ++g) {
        const uint64_t col = (word0 + g) * 32 + lane;
        const bool valid = col >= col0 && col < col0 + n_blooms;
        vmask[g] = __ballot_sync(0xffffffffu, valid);
        any |= vmask[g] != 0;
        all &= vmask[g] == 0xffffffffu;
#pragma unroll
        for (int s = 0; s < 8; ++s) w[g][s] = 0;
        if (valid) ldg256_stream(blooms + (col - col0) * bloom_stride + (r0 >> 3), w[g]);
    }
    if (!any) return;  // warp-uniform
    const bool aligned = all && ((reinterpret_cast<uintptr_t>(matrix) + word0 * 4) & 31) == 0 && (pitch & 31) == 0;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        const uint64_t rb = r0 + 32 * s;
        uint32_t rowmask = 0xffffffffu;  // rows of this block below n_bits
        if (rb >= n_bits) rowmask = 0;
        else if (n_bits - rb < 32) rowmask = (1u << (uint32_t)(n_bits - rb)) - 1u;
        uint32_t y[kTWords];
#pragma unroll
        for (int g = 0; g < kTWords; ++g) y[g] = msb_first_swizzle(transpose(msb_first_swizzle(w[g][s]) & rowmask));
        const uint64_t row = rb + lane;
        if (row < num_rows) {
            uint32_t *dst = reinterpret_cast<uint32_t *>(matrix + row * pitch + word0 * 4);
            if (aligned) {
                stg256(dst, y);
            } else {
#pragma unroll
                for (int g = 0; g < kTWords; ++g) {
                    if (vmask[g] == 0) continue;
                    const uint32_t smask = msb_first_swizzle(vmask[g]);
                    dst[g] = vmask[g] == 0xffffffffu ? y[g] : ((dst[g] & ~smask) | (y[g] & smask));
                }
            }
        }
    }
}

constexpr bool kTransposeWide = true;  // false: the one-word-per-warp kernel above

cudaError_t launch_transpose_blooms(uint8_t *matrix, uint64_t pitch, uint64_t num_rows, uint64_t col0, uint64_t n_blooms,
                                    const uint8_t *d_blooms, uint64_t bloom_stride, uint64_t n_bits, cudaStream_t stream)
{
    if (n_blooms == 0 || num_rows == 0) return cudaSuccess;
    const uint64_t row_blocks = (num_rows + 255) / 256;
    if ((bloom_stride & 31) || bloom_stride < row_blocks * 32) return cudaErrorInvalidValue;
    if (kTransposeWide && (reinterpret_cast<uintptr_t>(d_blooms) & 31) == 0) {  // (256-bit loads need 32-byte aligned filters)
        const uint64_t w_first = (col0 >> 5) & ~(uint64_t)(kTWords - 1), w_end = (col0 + n_blooms + 31) >> 5;  // words [w_first, w_end)
        const uint64_t gy = (w_end - w_first + 8 * kTWords - 1) / (8 * kTWords);  // 8 warps x kTWords words per CTA
        if (row_blocks > 0x7fffffffull || gy > 65535) return cudaErrorInvalidConfiguration;
        transpose_blooms_wide_kernel<<<dim3((unsigned)row_blocks, (unsigned)gy), 256, 0, stream>>>(
            matrix, pitch, num_rows, col0, n_blooms, d_blooms, bloom_stride, n_bits);
        return cudaGetLastError();
    }
    const uint64_t words = ((col0 + n_blooms + 31) >> 5) - (col0 >> 5);
    const uint64_t gy = (words + 7) / 8;
    if (row_blocks > 0x7fffffffull || gy > 65535) return cudaErrorInvalidConfiguration;
    transpose_blooms_kernel<<<dim3((unsigned)row_blocks, (unsigned)gy), 256, 0, stream>>>(
        matrix, pitch, num_rows, col0, n_blooms, d_blooms, bloom_stride, n_bits);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K11: score=True (bigsi/graph/bigsi.py:232-239 with unpack_and_cat, 47-56): for EVERY window of the
// query (duplicates included, in sequence order) and every hit column, is the k-mer present?
// hash_windows_kernel: one thread per (window, seed) hashes the canonical k-mer straight from the
// sequence (windows overlap, stride 1).  presence_kernel: one thread per (window, hit column) ANDs
// the h single bits; out[c * n_windows + w] is the character '0' or '1' -- row c IS the reference's
// "kmer-presence" string of that hit.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hash_windows_kernel(const uint8_t *__restrict__ seq, uint64_t n_windows, int k, int h,
                                                           uint32_t m, int32_t *__restrict__ rows_out)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_windows * (uint64_t)h) return;
    const uint64_t wdw = t / (uint32_t)h;
    const uint32_t seed = (uint32_t)(t % (uint32_t)h);
    const uint8_t *s = seq + wdw;
    bool fwd = true;
    for (int j = 0; j < k; ++j) {
        const uint32_t a = s[j], b = comp_base(s[k - 1 - j]);
        if (a != b) {
            fwd = a < b;
            break;
        }
    }
    auto byte_at = [&](int j) -> uint32_t { return fwd ? (uint32_t)s[j] : comp_base(s[k - 1 - j]); };
    const int nblocks = k >> 2, rem = k & 3;
    uint32_t h1 = seed;
    for (int b = 0; b < nblocks; ++b)
        h1 = murmur_block(h1, byte_at(4 * b) | (byte_at(4 * b + 1) << 8) | (byte_at(4 * b + 2) << 16) | (byte_at(4 * b + 3) << 24));
    if (rem) {
        uint32_t k1 = 0;
        for (int u = 0; u < rem; ++u) k1 |= byte_at(4 * nblocks + u) << (8 * u);
        h1 = murmur_tail(h1, k1);
    }
    rows_out[t] = murmur_finish_mod(h1, (uint32_t)k, m);
}

__global__ void __launch_bounds__(256) presence_kernel(const uint8_t *__restrict__ matrix, uint64_t pitch,
                                                       const int32_t *__restrict__ rows, uint64_t n_windows, int h,
                                                       const int32_t *__restrict__ cols, uint32_t n_cols,
                                                       uint8_t *__restrict__ out)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_windows * n_cols) return;
    const uint64_t wdw = t / n_cols;
    const uint32_t ci = (uint32_t)(t % n_cols);
    const uint32_t c = (uint32_t)__ldg(cols + ci);
    const uint32_t mask = 0x80u >> (c & 7);
    uint32_t bit = 1;
    for (int j = 0; j < h; ++j)
        bit &= (__ldg(matrix + (uint64_t)(uint32_t)__ldg(rows + wdw * h + j) * pitch + (c >> 3)) & mask) ? 1u : 0u;
    out[(uint64_t)ci * n_windows + wdw] = (uint8_t)('0' + bit);
}

cudaError_t launch_hash_windows(const uint8_t *d_seq, uint64_t n_windows, int k, int h, uint64_t m, int32_t *d_rows_out,
                                cudaStream_t stream)
{
    if (n_windows == 0) return cudaSuccess;
    const uint64_t blocks = (n_windows * (uint64_t)h + 255) / 256;
    if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    hash_windows_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_seq, n_windows, k, h, (uint32_t)m, d_rows_out);
    return cudaGetLastError();
}

cudaError_t launch_presence(const uint8_t *matrix, uint64_t pitch, const int32_t *d_rows, uint64_t n_windows, int h,
                            const int32_t *d_cols, uint64_t n_cols, uint8_t *d_out, cudaStream_t stream)
{
    if (n_windows == 0 || n_cols == 0) return cudaSuccess;
    const uint64_t blocks = (n_windows * n_cols + 255) / 256;
    if (blocks > 0x7fffffffull || n_cols > 0xffffffffull) return cudaErrorInvalidConfiguration;
    presence_kernel<<<(unsigned)blocks, 256, 0, stream>>>(matrix, pitch, d_rows, n_windows, h, d_cols, (uint32_t)n_cols, d_out);
    return cudaGetLastError();
}

}  // namespace bigsi
