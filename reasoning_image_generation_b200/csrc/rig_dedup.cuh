// rig_dedup.cuh -- near-duplicate filter: packed perceptual hash + all-pairs popcount Hamming scan.
//
// The reference has NO cross-sample duplicate filter; this stage is new
// functionality defined by this repository (oracle/dedup_oracle.py is its CPU statement).  It is
// integer work on the INT/POPC pipes -- no tensor cores:
//   hash  : Y = (77 R + 150 G + 29 B) >> 8; S_b = sum of Y over block b of a 16 x 16 grid;
//           T = sum_b S_b; bit b = (256 * S_b > T)  ->  256 bits = 4 x uint64 per image
//   filter: sample j is dropped iff some i < j has popc(h_i xor h_j) <= threshold
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rig {

constexpr int HASH_GRID = 16;                       // 16 x 16 blocks
constexpr int HASH_WORDS = HASH_GRID * HASH_GRID / 64;   // uint64 words per hash

// One CTA per (image, block-row), 8 warps; a warp sweeps rows of the block-row.  A lane reads 16 pixels
// (3 aligned 128-bit words = 4 quads of 3 words), luma per pixel is one byte-permute + one dp4a, and a
// 16-pixel group always lies inside one block (W / 16 is a multiple of 16), so block sums are shuffles
// plus one shared atomic per block and sweep.
// coef = 77 | 150 << 8 | 29 << 16 for RGB byte order, 29 | 150 << 8 | 77 << 16 for BGR.
__device__ __forceinline__ uint32_t luma4(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t coef)
{
    return (__dp4a(w0 & 0x00ffffffu, coef, 0u) >> 8) + (__dp4a(__byte_perm(w0, w1, 0x4543) & 0x00ffffffu, coef, 0u) >> 8) +
           (__dp4a(__byte_perm(w1, w2, 0x4432) & 0x00ffffffu, coef, 0u) >> 8) + (__dp4a(w2 >> 8, coef, 0u) >> 8);
}

__global__ void __launch_bounds__(256)
k_block_sums(const uint8_t *__restrict__ imgs, int W, int H, uint32_t coef, uint32_t *__restrict__ sums)
{
    __shared__ uint32_t acc[HASH_GRID];
    const int img = blockIdx.x / HASH_GRID, by = blockIdx.x % HASH_GRID;
    const int bh = H / HASH_GRID, bw = W / HASH_GRID;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < HASH_GRID) acc[threadIdx.x] = 0;
    __syncthreads();
    const uint4 *base = (const uint4 *)(imgs + ((size_t)img * H + (size_t)by * bh) * (size_t)W * 3);
    const int vec_per_row = W * 3 / 16, groups_per_row = W / 16;    // a group = 16 pixels = 3 uint4
    const int lanes_per_block = bw / 16;                            // a power of two (checked on the host)
    for (int row = warp; row < bh; row += 8) {
        const uint4 *rp = base + (size_t)row * vec_per_row;
        for (int g0 = 0; g0 < groups_per_row; g0 += 32) {
            const int g = g0 + lane;
            uint32_t s = 0;
            if (g < groups_per_row) {
                const uint4 a = __ldg(rp + 3 * g), b = __ldg(rp + 3 * g + 1), c = __ldg(rp + 3 * g + 2);
                s = luma4(a.x, a.y, a.z, coef) + luma4(a.w, b.x, b.y, coef) + luma4(b.z, b.w, c.x, coef) + luma4(c.y, c.z, c.w, coef);
            }
            // lanes of one block are contiguous: segmented sum over groups of lanes_per_block lanes
            for (int o = 1; o < lanes_per_block && o < 32; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const int group = lanes_per_block < 32 ? lanes_per_block : 32;
            if ((lane & (group - 1)) == 0 && g < groups_per_row) atomicAdd(&acc[(g * 16) / bw], s);
        }
    }
    __syncthreads();
    if (threadIdx.x < HASH_GRID) sums[(size_t)img * HASH_GRID * HASH_GRID + by * HASH_GRID + threadIdx.x] = acc[threadIdx.x];
}

// One warp per image: sums -> 256-bit hash
__global__ void k_hash_bits(const uint32_t *__restrict__ sums, int n, uint64_t *__restrict__ hashes)
{
    const int img = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (img >= n) return;
    const uint32_t *s = sums + (size_t)img * 256;
    uint32_t v[8];
    uint64_t t = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { v[k] = s[k * 32 + lane]; t += v[k]; }
#pragma unroll
    for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        uint32_t bits = __ballot_sync(0xffffffffu, 256ull * v[k] > t);
        if (lane == 0) {
            // block index b = 32k + lane -> word b / 64, bit b % 64
            uint32_t *w = (uint32_t *)(hashes + (size_t)img * HASH_WORDS);
            w[k] = bits;
        }
    }
}

// Same, for sums produced by k_raster<true>: those are deltas against the all-template panel, whose block sums are
// constants: bh rows x (bw background pixels minus the panel's vertical grid lines crossing the block) x luma(bg).
__global__ void k_hash_bits_fused(const uint32_t *__restrict__ sums, const uint8_t *__restrict__ panel_flags, int n, int W, int H,
                                  uint32_t luma_bg, int n_grid, const int16_t *__restrict__ grid_x, uint64_t *__restrict__ hashes)
{
    const int img = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (img >= n) return;
    const uint32_t *s = sums + (size_t)img * 256;
    const bool grid = panel_flags[img] & 1;
    const int bw = W / 16, bh = H / 16;
    uint32_t v[8];
    uint64_t t = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int b = k * 32 + lane, bx = b & 15;
        int lines = 0;
        if (grid)
            for (int g = 0; g < n_grid; g++) lines += grid_x[g] >= bx * bw && grid_x[g] < (bx + 1) * bw && grid_x[g] < W;
        v[k] = s[b] + (uint32_t)(bh * (bw - lines)) * luma_bg;
        t += v[k];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        uint32_t bits = __ballot_sync(0xffffffffu, 256ull * v[k] > t);
        if (lane == 0) ((uint32_t *)(hashes + (size_t)img * HASH_WORDS))[k] = bits;
    }
}

// All-pairs scan.
//
// Work unit: a macro tile = DD_B rows x DD_B columns of the pair matrix (only tiles that reach above the diagonal exist; the
// 1-D grid enumerates them row-major).  The CTA stages the FIRST 64 BITS of the DD_B column hashes in shared memory (32 KB) and
// walks them once per pass of DD_R x 256 rows; a thread keeps the first 64 bits of its DD_R rows in registers.
//
// Per pair the inner loop costs 1 POPC:   f = (a.lo ^ t.lo) | (a.hi ^ t.hi)      2 x LOP3
//                                         p = popc(f)                            1 x POPC
//                                         m = min(m, p, p')                      1/2 x VIMNMX3
// popc(f) <= popc of the 64-bit xor <= the 256-bit distance, so `popc(f) > thr` rejects exactly (no false negatives); for
// unrelated hashes f has ~24 of 32 bits set, so the filter passes ~1e-8 of the pairs at thr = 10.  Only when the minimum over
// a group of 8 columns x DD_R rows is <= thr does the thread re-walk the group, fetch the full 256-bit hashes (L2) and apply
// the definition `dist <= thr && i < j`.  The column word pair is one broadcast LDS.128 per 2 columns and DD_R rows.
// Bound: the POPC pipe (a quarter-rate pipe: 16 lanes / clk / SM measured by k_int_pipe_peak<0>, see bench.py `dedup.roofline`).
constexpr int DD_B = 4096;          // macro tile edge
constexpr int DD_R = 4;             // rows per thread and pass
constexpr int DD_THREADS = 256;
constexpr int DD_G = 8;             // columns per group (one min-test per group)

__device__ __forceinline__ int dd_pre(uint2 a, uint2 t) { return __popc((a.x ^ t.x) | (a.y ^ t.y)); }

// The rare path: re-walk one group (operands re-read from shared memory / L2, so that nothing of the hot loop's registers has
// its address taken) and apply the definition on the full 256 bits.  Returns the hits of this thread.
__device__ __noinline__ unsigned dd_confirm(const uint64_t *__restrict__ h, unsigned long long n, unsigned long long i_end, int thr,
                                            unsigned long long i_first, const uint2 *cols, unsigned long long j_first, uint8_t *__restrict__ drop)
{
    unsigned found = 0;
#pragma unroll 1
    for (int r = 0; r < DD_R; r++) {
        const unsigned long long i = i_first + (unsigned long long)r * DD_THREADS;
        if (i >= i_end) break;
        const ulonglong2 *hi = (const ulonglong2 *)(h + i * HASH_WORDS);
        const ulonglong2 x0 = __ldg(hi), x1 = __ldg(hi + 1);
        const uint2 a = make_uint2((uint32_t)x0.x, (uint32_t)(x0.x >> 32));
#pragma unroll 1
        for (int c = 0; c < DD_G; c++) {
            const unsigned long long j = j_first + c;
            if (j >= n) break;
            if (j <= i || dd_pre(a, cols[c]) > thr) continue;
            const ulonglong2 *hj = (const ulonglong2 *)(h + j * HASH_WORDS);
            const ulonglong2 y0 = __ldg(hj), y1 = __ldg(hj + 1);
            const int d = __popcll(x0.x ^ y0.x) + __popcll(x0.y ^ y0.y) + __popcll(x1.x ^ y1.x) + __popcll(x1.y ^ y1.y);
            if (d <= thr) { drop[j] = 1; found++; }
        }
    }
    return found;
}

// tile id -> (row block I, column block J): row block I holds the column blocks J >= I + c0, `per_row0` of them for I = 0
__device__ __forceinline__ void dd_tile(unsigned long long t, long long per_row0, unsigned &I, unsigned &J_off)
{
    // P(I) = I * per_row0 - I (I - 1) / 2 tiles precede row block I; largest I with P(I) <= t
    const double b = 2.0 * (double)per_row0 + 1.0;
    long long i = (long long)((b - sqrt(fmax(b * b - 8.0 * (double)t, 0.0))) * 0.5);
    if (i < 0) i = 0;
    while (i > 0 && (unsigned long long)(i * per_row0 - i * (i - 1) / 2) > t) i--;
    while ((unsigned long long)((i + 1) * per_row0 - (i + 1) * i / 2) <= t) i++;
    I = (unsigned)i;
    J_off = (unsigned)(t - (unsigned long long)(i * per_row0 - i * (i - 1) / 2));
}

__global__ void __launch_bounds__(DD_THREADS)
k_hamming(const uint64_t *__restrict__ h, unsigned long long n, unsigned long long row_begin, unsigned long long row_end,
          int thr, long long per_row0, unsigned c0, unsigned long long tile0, uint8_t *__restrict__ drop, unsigned long long *__restrict__ hits)
{
    __shared__ __align__(16) uint2 tile[DD_B];
    unsigned I, Joff;
    dd_tile(tile0 + blockIdx.x, per_row0, I, Joff);
    const unsigned long long i0 = row_begin + (unsigned long long)I * DD_B;
    const unsigned long long j0 = ((unsigned long long)I + c0 + Joff) * DD_B;
    const unsigned long long jn = (n - j0 < (unsigned long long)DD_B) ? n - j0 : DD_B;
    for (unsigned k = threadIdx.x; k < DD_B; k += DD_THREADS)
        tile[k] = k < jn ? *(const uint2 *)(h + (j0 + k) * HASH_WORDS) : make_uint2(0u, 0u);
    __syncthreads();
    const unsigned long long i_end = (i0 + DD_B < row_end) ? i0 + DD_B : row_end;
    const int ngroups = (int)((jn + DD_G - 1) / DD_G);
    unsigned long long found = 0;
#pragma unroll 1
    for (unsigned long long ip = i0; ip < i_end; ip += DD_R * DD_THREADS) {
        // columns at or below this pass's first row cannot pair with it: start at the group holding column ip + 1
        const int g_begin = (ip + 1 > j0) ? (int)((ip + 1 - j0) / DD_G) : 0;
        if (g_begin >= ngroups) break;
        uint2 a[DD_R];
        const unsigned long long i_first = ip + threadIdx.x;
#pragma unroll
        for (int r = 0; r < DD_R; r++) {
            const unsigned long long i = i_first + (unsigned long long)r * DD_THREADS;
            // rows past the range keep the previous row's bits: the confirm step ignores them
            a[r] = i < i_end ? *(const uint2 *)(h + i * HASH_WORDS) : (r ? a[r - 1] : make_uint2(0x55555555u, 0xaaaaaaaau));
        }
#pragma unroll 1
        for (int g = g_begin; g < ngroups; g++) {
            const uint4 *q = (const uint4 *)(tile + g * DD_G);
            int m = 64;
#pragma unroll
            for (int k = 0; k < DD_G / 2; k++) {
                const uint4 w = q[k];
#pragma unroll
                for (int r = 0; r < DD_R; r++) m = __vimin3_s32(m, dd_pre(a[r], make_uint2(w.x, w.y)), dd_pre(a[r], make_uint2(w.z, w.w)));
            }
            if (m <= thr) found += dd_confirm(h, n, i_end, thr, i_first, tile + g * DD_G, j0 + (unsigned long long)g * DD_G, drop);
        }
    }
    if (hits && found) atomicAdd(hits, found);
}

// Integer-pipe peaks for the scan's roofline ("measure with a microbenchmark; do not assume").
// WHICH = 0: POPC; 1: LOP3 (the ALU pipe the filter's xor/or and the 3-input minimum share).  8 independent chains per thread.
template <int WHICH>
__global__ void __launch_bounds__(256)
k_int_pipe_peak(uint32_t *__restrict__ out, uint32_t seed, int iters)
{
    uint32_t x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = seed * (threadIdx.x + 1u) + 0x9e3779b9u * (k + blockIdx.x);
    const uint32_t y = seed ^ 0x5bd1e995u, z = seed * 3u + 1u;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (WHICH == 0) x[k] = __popc(x[k]);                     // a pure dependent POPC chain (its timing does not depend on the data)
                else x[k] = (x[k] ^ y) | (x[(k + 1) & 7] & z);            // two LOP3 (four runtime inputs)
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= x[k];
    if (s == 0x12345678u + y) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace rig
