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

// All-pairs scan.  CTA = 256 rows x one column tile of TJ hashes staged in shared memory; a thread
// keeps its row hash in registers and walks the tile (every lane reads the same shared address:
// broadcast).  The first 64 bits reject almost every pair with 2 POPCs; the remaining 192 bits are
// only examined when some lane of the warp is still under the threshold.
constexpr int DD_TI = 256;
constexpr int DD_TJ = 1024;

__global__ void __launch_bounds__(DD_TI)
k_hamming(const uint64_t *__restrict__ h, unsigned long long n, unsigned long long row_begin, unsigned long long row_end,
          int thr, uint8_t *__restrict__ drop, unsigned long long *__restrict__ hits)
{
    __shared__ uint64_t tile[DD_TJ * HASH_WORDS];
    const unsigned long long i0 = row_begin + (unsigned long long)blockIdx.y * DD_TI;
    const unsigned long long j0 = (unsigned long long)blockIdx.x * DD_TJ;
    if (j0 + DD_TJ <= i0 + 1 || i0 >= row_end) return;   // tile entirely at or below the diagonal
    const unsigned long long jn = (n - j0 < (unsigned long long)DD_TJ) ? n - j0 : DD_TJ;
    for (unsigned long long k = threadIdx.x; k < jn * HASH_WORDS; k += blockDim.x) tile[k] = h[j0 * HASH_WORDS + k];
    __syncthreads();
    const unsigned long long i = i0 + threadIdx.x;
    const bool live = i < row_end && i < n;
    uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    if (live) { a0 = h[i * 4]; a1 = h[i * 4 + 1]; a2 = h[i * 4 + 2]; a3 = h[i * 4 + 3]; }
    unsigned long long found = 0;
    for (unsigned long long jj = 0; jj < jn; jj++) {
        const uint64_t *t = tile + jj * HASH_WORDS;
        int d = __popcll(a0 ^ t[0]);
        bool cand = live && d <= thr && (j0 + jj) > i;
        if (__any_sync(0xffffffffu, cand)) {
            if (cand) {
                d += __popcll(a1 ^ t[1]) + __popcll(a2 ^ t[2]) + __popcll(a3 ^ t[3]);
                if (d <= thr) { drop[j0 + jj] = 1; found++; }
            }
        }
    }
    if (hits && found) atomicAdd(hits, found);
}

}  // namespace rig
