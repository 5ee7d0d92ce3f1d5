// rig_png.cuh -- device-side PNG writer: filter + deflate + checksums in CUDA.
//
// Replaces cv2.imwrite / fig.savefig's libpng on the host (src/utils.py:18-21; multigraph_generation/generator.py:256).
// A file holds ONE dynamic-Huffman deflate block whose code is fixed per context (rig_png_code.h: built on the host from a
// frequency model of this writer's token stream, so that "258 zeros", the literals 0 / 255 and the filter byte cost 2-4 bits;
// files are half the size fixed-Huffman ones were and about libpng's).  An image is cut into pieces of PNG_PIECE_ROWS rows; one
// warp encodes one piece into a BIT string
//     per row: the Sub-filtered bytes as literals, zero runs as a literal 0 + (length, distance 1) matches
// with no framing of its own, so pieces are produced independently and k_png_assemble concatenates them at bit offsets behind
// the block header and closes the block.  The puzzle images are flat colour with thin strokes: after the Sub filter
// (x - left pixel) almost every byte is zero, and restricting LZ77 to zero runs loses little.
//   k_png_pieces    lanes own 48-byte chunks (16 pixels) of a row; zero runs are merged across lanes with a ballot; bits are
//                   counted, offsets come from a warp scan, then the same walk emits them with shared-memory ORs
//   k_png_sizes / k_png_scan   bytes per file and their exclusive prefix sum (file offsets)
//   k_png_assemble  one warp per file: PNG / zlib framing, the deflate stream gathered word by word (block header, the pieces'
//                   bits at their prefix-sum offsets, end of block), Adler-32 (combined from per-piece partial sums), CRC-32
//                   (per-lane table-driven CRC of a contiguous segment, combined by polynomial multiplication), IEND
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rig_png_code.h"

namespace rig {

#ifndef PNG_PIECE_ROWS
#define PNG_PIECE_ROWS 2
#endif
constexpr int PNG_ROWS = PNG_PIECE_ROWS;
constexpr int PNG_CHUNK_PX = 16;              // pixels per lane per pass: 48 filtered bytes (RGB files) or 64 (RGBA files)
#ifndef PNG_CLAIM_PIECES
#define PNG_CLAIM_PIECES 4
#endif
constexpr uint32_t PNG_CLAIM = PNG_CLAIM_PIECES;   // consecutive pieces a warp claims at a time
constexpr uint32_t PNG_SLOT = 512;            // bytes reserved per piece; larger pieces go to the overflow arena
constexpr uint32_t PNG_FILE_OVERHEAD = 63;    // signature 8 + IHDR 25 + IDAT 12 + zlib header 2 + Adler 4 + IEND 12
constexpr uint32_t ADLER_MOD = 65521u;
constexpr uint32_t CRC_POLY = 0xedb88320u;

struct PngPiece {
    uint32_t nbits;      // deflate bits of the piece
    uint32_t where;      // byte offset of the piece in the scratch area
    uint32_t a, b;       // Adler partial sums over the piece's uncompressed bytes: sum d, sum d * (len - index), mod 65521
};

struct PngTables {
    uint32_t crc[256];
    uint32_t x2n[32];    // x^(2^n) mod P (reflected), as zlib's crc32_combine uses
    uint8_t header[33];  // signature + IHDR chunk of this context's W x H
    // the context's Huffman code (rig_png_code.h): token tables (code bits LSB-first | number of bits << 24) and the block header
    uint32_t lit[256];
    uint32_t match[259];
    uint32_t eob;
    uint32_t hdr_bits;
    uint32_t hdr[PNG_HDR_WORDS];
};

__host__ __device__ inline size_t png_piece_buf_bytes(int W, int ch)   // worst case of one piece, + tail, rounded to 16
{
    const size_t bits = (size_t)PNG_ROWS * (1 + (size_t)ch * W) * PNG_MAX_CODE_BITS + 64;     // every byte a literal of the longest code
    return ((bits + 7) / 8 + 8 + 15) & ~(size_t)15;
}

// single-walk sink: the lane's bits go to its own staging words; they are moved to the piece buffer once the warp scan has
// placed them
constexpr int PNG_STAGE_WORDS = 28;      // >= (10 + 64 * 10 + 9 * 16 + 2 * 10 + 31) / 32 words of bits a chunk can need (PNG_MAX_CODE_BITS = 10)
static_assert(PNG_MAX_CODE_BITS <= 10, "PNG_STAGE_WORDS is sized for codes of at most 10 bits");
struct PngStage {
    uint32_t *stage;
    uint32_t acc = 0;            // the bits not yet written (fewer than 32)
    uint32_t nacc = 0, k = 0;
    __device__ __forceinline__ void put(uint32_t code, int n)       // n <= 32 bits of `code`, the rest of it zero
    {
        acc |= code << nacc;
        const uint32_t tot = nacc + (uint32_t)n;
        if (tot >= 32u) {
            stage[k++] = acc;
            acc = __funnelshift_rc(code, 0u, 32u - nacc);           // the bits of `code` that did not fit (a clamped shift: 0 when nacc is 0)
            nacc = tot - 32u;
        } else nacc = tot;
    }
    __device__ __forceinline__ uint32_t finish()      // flushes the tail, returns the number of bits
    {
        stage[k] = acc;
        return 32u * k + nacc;
    }
};
struct PngEmit {
    uint32_t *buf;
    uint32_t pos;
    __device__ __forceinline__ void put(uint32_t code, int n)
    {
        const uint32_t w = pos >> 5, sh = pos & 31u;
        atomicOr(buf + w, code << sh);
        if (sh + (uint32_t)n > 32u) atomicOr(buf + w + 1, code >> (32u - sh));
        pos += (uint32_t)n;
    }
};
// token tables (static shared memory, filled once per CTA): code bits | number of bits << 24
struct PngLut {
    uint32_t lit[256];      // literal byte
    uint32_t match[259];    // match of length 3..258 at distance 1
};
__device__ __forceinline__ void png_fill_lut(PngLut &T, const PngTables *tab)
{
    for (uint32_t i = threadIdx.x; i < 256u; i += blockDim.x) T.lit[i] = tab->lit[i];
    for (uint32_t i = 3u + threadIdx.x; i <= 258u; i += blockDim.x) T.match[i] = tab->match[i];
}
// r zeros whose predecessor in the stream is a zero byte
template <class Sink>
__device__ __forceinline__ void png_zero_matches(Sink &s, uint32_t r, const PngLut &T)
{
    const uint32_t full = T.match[258];
    while (r >= 258u) { s.put(full & 0xffffffu, (int)(full >> 24)); r -= 258u; }
    if (r >= 3u) { const uint32_t t = T.match[r]; s.put(t & 0xffffffu, (int)(t >> 24)); r = 0; }
    while (r) { s.put(T.lit[0] & 0xffffffu, (int)(T.lit[0] >> 24)); r--; }
}
// The tokens of one lane's chunk.  nz: bit i set = filtered byte i is non-zero; fb: the chunk's filtered bytes (shared memory);
// skip: leading zeros that belong to a run an earlier lane emits; tail_extra: zeros after this chunk that continue its trailing
// run (all-zero lanes that follow, plus the leading zeros of the lane after them).
template <int CHUNK, class Sink>
__device__ __forceinline__ void png_chunk_tokens(Sink &s, unsigned long long nz, const uint8_t *fb, int skip, uint32_t tail_extra, bool filter_byte,
                                                 const PngLut &T)
{
    if (filter_byte) s.put(T.lit[1] & 0xffffffu, (int)(T.lit[1] >> 24));          // filter type 1 (Sub) opens the row
    int pos = skip;
    while (pos < CHUNK) {
        const unsigned long long rem = nz >> pos;
        if (rem & 1ull) {
            const uint32_t t = T.lit[fb[pos]];
            if ((rem & 2ull) && pos + 1 < CHUNK) {             // two literals in one go (at most 30 bits)
                const uint32_t u = T.lit[fb[pos + 1]];
                s.put((t & 0xffffffu) | ((u & 0xffffffu) << (t >> 24)), (int)((t >> 24) + (u >> 24)));
                pos += 2;
            } else {
                s.put(t & 0xffffffu, (int)(t >> 24));
                pos++;
            }
        } else if ((rem & 6ull) != 0ull) {                     // one or two zeros between literals (the alpha delta of an RGBA
            const int gap = (rem & 2ull) ? 1 : 2;              // pixel): too short for a match, literal zeros
            const uint32_t z = T.lit[0] & 0xffffffu, zb = T.lit[0] >> 24;
            s.put(gap == 2 ? (z | (z << zb)) : z, (int)(zb * gap));
            pos += gap;
        } else {
            const int gap = rem ? (__ffsll((long long)rem) - 1) : (CHUNK - pos);
            uint32_t r = (uint32_t)gap;
            if (!rem) r += tail_extra;
            s.put(T.lit[0] & 0xffffffu, (int)(T.lit[0] >> 24));                    // the run's first zero is a literal
            png_zero_matches(s, r - 1u, T);
            pos += gap;
        }
    }
}

struct PngParams {
    const uint8_t *imgs;
    int n, W, H, bgr;               // bgr: channel order of the input images
    uint32_t pieces_per_image;
    PngPiece *pieces;
    uint8_t *scratch;               // [n * pieces_per_image * PNG_SLOT][overflow arena]
    unsigned long long over_begin, over_cap;
    unsigned long long *counters;   // [0] overflow arena cursor, [1] error flag, [2] next unclaimed piece
    const PngTables *tables;
    const uint8_t *blank;           // optional hint, [n][blank_bands]: 1 = rows [b * blank_rows, (b + 1) * blank_rows) are all white
    int blank_rows;
    uint32_t blank_bands;
};

// shared memory per warp: bit buffer + 32 chunks of filtered bytes + staging words + the bits of an all-white row
__host__ __device__ inline size_t png_warp_smem(int W, int ch) { return png_piece_buf_bytes(W, ch) + 32 * PNG_CHUNK_PX * ch + 32 * 4 * PNG_STAGE_WORDS + 64; }

// ORs the nbits bits that start at bit src_start of src into dst at bit dst_pos (dst is zero there); lanes share the words
__device__ __forceinline__ void png_copy_bits(uint32_t *dst, uint32_t dst_pos, const uint32_t *src, uint32_t src_start, uint32_t nbits, int lane)
{
    for (uint32_t o = (uint32_t)lane * 32u; o < nbits; o += 1024u) {
        const uint32_t sp = src_start + o, sw = sp >> 5, ss = sp & 31u;
        uint32_t v = __funnelshift_r(src[sw], src[sw + 1], ss);           // 32 source bits from bit sp
        const uint32_t left = nbits - o;
        if (left < 32u) v &= (1u << left) - 1u;
        const uint32_t dp = dst_pos + o, dw = dp >> 5, ds = dp & 31u;
        atomicOr(dst + dw, v << ds);
        if (ds) atomicOr(dst + dw + 1, v >> (32u - ds));
    }
}

// CH = 3: RGB files (what cv2.imwrite writes for the panels); CH = 4: RGBA files with alpha 255 (what fig.savefig writes: Agg's
// canvas is RGBA and opaque).  The input images are 3 bytes per pixel either way.
template <int CH>
__global__ void __launch_bounds__(256) k_png_pieces(const PngParams P)
{
    constexpr int CHUNK = PNG_CHUNK_PX * CH, FW = CHUNK / 4;      // filtered bytes / words per lane chunk
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ PngLut lut;
    png_fill_lut(lut, P.tables);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per_warp = png_warp_smem(P.W, CH), bufb = png_piece_buf_bytes(P.W, CH);
    uint32_t *buf = (uint32_t *)(smem + (size_t)warp * per_warp);
    uint8_t *fbytes = (uint8_t *)buf + bufb + lane * CHUNK;
    uint32_t *stage = (uint32_t *)((uint8_t *)buf + bufb + 32 * CHUNK) + lane * PNG_STAGE_WORDS;
    uint32_t *wrow = (uint32_t *)((uint8_t *)buf + bufb + 32 * CHUNK + 32 * 4 * PNG_STAGE_WORDS);
    const int row_bytes = 3 * P.W, chunks = P.W / 16;            // input rows: packed RGB
    const uint32_t RL = 1u + (uint32_t)(CH * P.W);                // filtered bytes per row of the file
    const uint32_t buf_words = (uint32_t)(bufb / 4);
    const bool one_pass = chunks <= 32;        // rows of up to 512 pixels: a row that repeats the row above copies its bits
    // the bits of an all-white row (what most rows of this data set are): the virtual "row above" of every piece
    uint32_t white_bits = 0;
    if (lane < 16) wrow[lane] = 0u;
    for (uint32_t i = lane; i < buf_words; i += 32) buf[i] = 0u;
    __syncwarp();
    if (one_pass) {
        if (lane == 0) {
            PngEmit em;
            em.buf = wrow;
            em.pos = 0;
            em.put(lut.lit[1] & 0xffffffu, (int)(lut.lit[1] >> 24));
            for (int q = 0; q < CH; q++) em.put(lut.lit[255] & 0xffffffu, (int)(lut.lit[255] >> 24));
            em.put(lut.lit[0] & 0xffffffu, (int)(lut.lit[0] >> 24));
            png_zero_matches(em, RL - 1u - (uint32_t)CH - 1u, lut);
            white_bits = em.pos;
        }
        white_bits = __shfl_sync(0xffffffffu, white_bits, 0);
    }
    __syncwarp();
    const unsigned long long total_pieces = (unsigned long long)P.n * P.pieces_per_image;
    uint32_t used_words = 0;                   // words of buf the previous piece dirtied
    PngPiece canned;                           // this warp's first blank piece (nbits 0: none yet)
    canned.nbits = 0u; canned.where = 0u; canned.a = 0u; canned.b = 0u;
    // pieces are claimed a few at a time from a global cursor: blank pieces cost a fraction of drawn ones, and CTAs that become
    // resident late (behind a k_figure launch on another stream) must not be left with a full static share
    for (;;) {
      unsigned long long gp0 = 0;
      if (lane == 0) gp0 = atomicAdd(P.counters + 2, (unsigned long long)PNG_CLAIM);
      gp0 = __shfl_sync(0xffffffffu, gp0, 0);
      if (gp0 >= total_pieces) break;
      const unsigned long long gp1 = gp0 + PNG_CLAIM < total_pieces ? gp0 + PNG_CLAIM : total_pieces;
      uint32_t image = (uint32_t)(gp0 / P.pieces_per_image), piece = (uint32_t)(gp0 % P.pieces_per_image) - 1u;
      for (unsigned long long gp = gp0; gp < gp1; gp++) {
        if (++piece == P.pieces_per_image) { piece = 0u; image++; }
        const int y0 = (int)piece * PNG_ROWS, y1 = min(P.H, y0 + PNG_ROWS);
        const uint32_t L = (uint32_t)(y1 - y0) * RL;              // uncompressed bytes in this piece
        const uint8_t *img = P.imgs + (size_t)image * P.H * row_bytes;
        // Blank pieces (every pixel white: three quarters of this data set) all have the same bytes and Adler sums: the first one
        // this warp meets is encoded the long way, the others just point at its bytes in the scratch area.
        bool blank = false;
        if (y1 - y0 == PNG_ROWS) {
            uint32_t acc = 0xffffffffu;
            bool hinted = false;
            if (P.blank) {                       // the renderer's word for it: no need to read the pixels
                const uint8_t *fl = P.blank + (size_t)image * P.blank_bands;
                if (P.blank_rows == PNG_ROWS) hinted = __ldg(fl + piece) != 0;          // the renderer's bands are the pieces
                else hinted = (__ldg(fl + y0 / P.blank_rows) & __ldg(fl + (y1 - 1) / P.blank_rows)) != 0;
            }
            if (!hinted || !canned.nbits)       // (the copy every other blank piece will point at is made from pixels that were read)
            for (int y = y0; y < y1; y++)
                for (int c = lane; c < chunks; c += 32) {
                    const uint4 *src = (const uint4 *)(img + (size_t)y * row_bytes + (size_t)c * 48);
                    const uint4 a = __ldg(src), b = __ldg(src + 1), d = __ldg(src + 2);
                    acc &= a.x & a.y & a.z & a.w & b.x & b.y & b.z & b.w & d.x & d.y & d.z & d.w;
                }
            blank = __all_sync(0xffffffffu, acc == 0xffffffffu);
            if (blank && canned.nbits) {
                if (lane == 0) P.pieces[gp] = canned;
                continue;
            }
        }
        for (uint32_t i = lane; i < used_words; i += 32) buf[i] = 0u;
        __syncwarp();
        uint32_t bitpos = 0;
        unsigned long long ad_a = 0, ad_b = 0;
        // the row above, as this lane saw it: raw words, Adler sums of its filtered chunk, where its bits are
        uint32_t pw[12];
#pragma unroll
        for (int i = 0; i < 12; i++) pw[i] = 0xffffffffu;
        uint32_t p_s0 = lane == 0 ? 255u * CH : 0u, p_s1 = lane == 0 ? 255u * (CH * (CH - 1) / 2) : 0u;
        const uint32_t *p_src = wrow;
        uint32_t p_start = 0, p_bits = white_bits;
        for (int y = y0; y < y1; y++) {
            const uint32_t row_pos = (uint32_t)(y - y0) * RL;        // position of the row's filter byte in the piece
            if (lane == 0) { ad_a += 1u; ad_b += (unsigned long long)(L - row_pos); }
            uint32_t pass_prev = 0u;           // last word of the previous pass's last chunk (rows wider than 512 pixels)
            const uint32_t row_start = bitpos;
            for (int c0 = 0; c0 < chunks; c0 += 32) {
                const int c = c0 + lane;
                const bool live = c < chunks;
                uint32_t w[12];
#pragma unroll
                for (int i = 0; i < 12; i++) w[i] = 0xffffffffu;
                if (live) {
                    const uint4 *src = (const uint4 *)(img + (size_t)y * row_bytes + (size_t)c * 48);
                    const uint4 a = __ldg(src), b = __ldg(src + 1), d = __ldg(src + 2);
                    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
                    w[8] = d.x; w[9] = d.y; w[10] = d.z; w[11] = d.w;
                    if (P.bgr) {
#pragma unroll
                        for (int g = 0; g < 12; g += 3) {
                            const uint32_t t = __byte_perm(w[g], w[g + 2], 0x0043);
                            const uint32_t o0 = __byte_perm(w[g], w[g + 1], 0x5012), o1 = __byte_perm(w[g + 1], t, 0x3540),
                                           o2 = __byte_perm(w[g + 1], w[g + 2], 0x5672);
                            w[g] = o0; w[g + 1] = o1; w[g + 2] = o2;
                        }
                    }
                }
                const uint32_t base = row_pos + 1u + (uint32_t)c * CHUNK;
                if (one_pass) {
                    uint32_t diff = 0u;
#pragma unroll
                    for (int i = 0; i < 12; i++) diff |= w[i] ^ pw[i];
                    if (__all_sync(0xffffffffu, diff == 0u)) {       // same row as the one above: same filtered bytes, same bits
                        png_copy_bits(buf, bitpos, p_src, p_start, p_bits, lane);
                        ad_a += p_s0;
                        ad_b += (unsigned long long)(L - base) * p_s0 - p_s1;
                        p_src = buf;
                        p_start = bitpos;
                        bitpos += p_bits;
                        __syncwarp();
                        continue;
                    }
#pragma unroll
                    for (int i = 0; i < 12; i++) pw[i] = w[i];
                }
                // Sub filter: byte i minus byte i - 3 (the row's first pixel keeps its value)
                uint32_t prev = __shfl_up_sync(0xffffffffu, w[11], 1);
                if (lane == 0) prev = pass_prev;   // 0 at the start of the row: the first pixel keeps its value
                pass_prev = __shfl_sync(0xffffffffu, w[11], 31);
                uint32_t f[12];
                uint32_t any = 0u;
#pragma unroll
                for (int i = 0; i < 12; i++) {
                    const uint32_t left = __funnelshift_r(i ? w[i - 1] : prev, w[i], 8);
                    f[i] = __vsub4(w[i], left);
                    any |= f[i];
                }
                // the file's filtered bytes: the RGB deltas, for RGBA files with a zero alpha delta per pixel (255 for the row's
                // first pixel, whose left neighbour counts as 0)
                uint32_t g[FW];
                if (CH == 3) {
#pragma unroll
                    for (int i = 0; i < 12; i++) g[i] = f[i];
                } else {
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const uint32_t f0 = f[3 * q], f1 = f[3 * q + 1], f2 = f[3 * q + 2];
                        g[(4 * q) % FW] = f0 & 0x00ffffffu;
                        g[(4 * q + 1) % FW] = __byte_perm(f0, f1, 0x0543) & 0x00ffffffu;
                        g[(4 * q + 2) % FW] = __byte_perm(f1, f2, 0x0432) & 0x00ffffffu;
                        g[(4 * q + 3) % FW] = f2 >> 8;
                    }
                    if (lane == 0 && c0 == 0) g[0] |= 0xff000000u;
                    any = 0u;
#pragma unroll
                    for (int i = 0; i < FW; i++) any |= g[i];
                }
                // zero structure of the chunk
                unsigned long long nz = 0ull;
                uint32_t s0 = 0, s1 = 0;
                if (any && live) {                                  // (a dead lane's chunk is all zeros but its left neighbour is not)
#pragma unroll
                    for (int i = 0; i < FW; i++) {
                        const uint32_t m = __vsetne4(g[i], 0u);       // 1 per non-zero byte, at bits 0, 8, 16, 24
                        const uint32_t nib = (m * 0x01020408u) >> 24; // ... gathered into bits 0..3 (the partial products share no bit)
                        nz |= (unsigned long long)nib << (4 * i);
                    }
                    uint4 *dst = (uint4 *)fbytes;
#pragma unroll
                    for (int i = 0; i < FW / 4; i++) dst[i] = make_uint4(g[4 * i], g[4 * i + 1], g[4 * i + 2], g[4 * i + 3]);
                    // Adler partial sums of this chunk: sum of the bytes and sum of byte * index, four bytes per dot product
#pragma unroll
                    for (int i = 0; i < FW; i++) {
                        const uint32_t t = __dp4a(g[i], 0x01010101u, 0u);
                        s0 += t;
                        s1 += __dp4a(g[i], 0x03020100u, 0u) + (uint32_t)(4 * i) * t;
                    }
                    ad_a += s0;
                    ad_b += (unsigned long long)(L - base) * s0 - s1;
                }
                p_s0 = s0; p_s1 = s1;
                const int lead = live ? (nz ? (__ffsll((long long)nz) - 1) : CHUNK) : 0;      // leading zeros
                const bool all_zero = live && nz == 0ull;
                const bool last_zero = live && !((nz >> (CHUNK - 1)) & 1ull);
                const uint32_t zmask = __ballot_sync(0xffffffffu, all_zero);
                const uint32_t prev_last_zero = __shfl_up_sync(0xffffffffu, (uint32_t)last_zero, 1);
                const int skip = (lane > 0 && prev_last_zero) ? lead : 0;
                // zeros that continue this chunk's trailing run: the all-zero lanes after it and the next lane's leading zeros
                uint32_t tail_extra = 0;
                {
                    const uint32_t after = lane == 31 ? 0u : (zmask >> (lane + 1));
                    const int run = lane == 31 ? 0 : (__ffs((int)~after) - 1);          // consecutive all-zero lanes that follow
                    const int nxt = lane + 1 + run;
                    const int nxt_lead = __shfl_sync(0xffffffffu, lead, nxt & 31);
                    tail_extra = (uint32_t)run * CHUNK + ((nxt < 32) ? (uint32_t)nxt_lead : 0u);
                }
                __syncwarp();
                const bool mine = live && skip < CHUNK;          // a fully skipped chunk emits nothing
                const bool fbyte = lane == 0 && c0 == 0;
                // one walk: tokens into the lane's staging words, a warp scan of the bit counts, then a shifted move
                uint32_t nbits_mine = 0;
                if (mine || fbyte) {
                    PngStage sg;
                    sg.stage = stage;
                    png_chunk_tokens<CHUNK>(sg, nz, fbytes, mine ? skip : CHUNK, tail_extra, fbyte, lut);
                    nbits_mine = sg.finish();
                }
                uint32_t incl = nbits_mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                {
                    const uint32_t at = bitpos + incl - nbits_mine;
                    for (uint32_t o = 0; o < nbits_mine; o += 32u) {
                        uint32_t v = stage[o >> 5];
                        const uint32_t left = nbits_mine - o;
                        if (left < 32u) v &= (1u << left) - 1u;
                        const uint32_t dp = at + o, dw = dp >> 5, ds = dp & 31u;
                        atomicOr(buf + dw, v << ds);
                        if (ds) atomicOr(buf + dw + 1, v >> (32u - ds));
                    }
                }
                bitpos += __shfl_sync(0xffffffffu, incl, 31);
                __syncwarp();
            }
            if (one_pass) {                             // the next row's "row above" (a copied row already points at itself)
                p_src = buf;
                p_start = row_start;
                p_bits = bitpos - row_start;
            }
        }
        // the piece is a bare bit string: k_png_assemble puts it behind its predecessor and closes the block after the last one
        __syncwarp();
        const uint32_t nbits = bitpos, nbytes = (nbits + 7u) / 8u;
        used_words = (nbytes + 3u) / 4u + 2u;
        // Adler partial sums of the piece
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            ad_a += __shfl_xor_sync(0xffffffffu, ad_a, o);
            ad_b += __shfl_xor_sync(0xffffffffu, ad_b, o);
        }
        unsigned long long where = gp * PNG_SLOT;
        bool failed = false;
        if (nbytes > PNG_SLOT) {
            unsigned long long off = 0;
            if (lane == 0) off = atomicAdd(P.counters, (unsigned long long)((nbytes + 15u) & ~15u));
            off = __shfl_sync(0xffffffffu, off, 0);
            if (off + nbytes + 16u > P.over_cap) {
                failed = true;
                if (lane == 0) atomicExch(P.counters + 1, 1ull);
            }
            where = P.over_begin + off;
        }
        __syncwarp();
        if (!failed) {
            uint4 *dst = (uint4 *)(P.scratch + where);
            const uint4 *src = (const uint4 *)buf;
            for (uint32_t i = lane; i < (nbytes + 15u) / 16u; i += 32) dst[i] = src[i];
        }
        if (lane == 0) {
            PngPiece pc;
            pc.nbits = failed ? 0u : nbits;
            pc.where = failed ? 0u : (uint32_t)(where / 16);      // scratch offsets are multiples of 16
            pc.a = (uint32_t)(ad_a % ADLER_MOD);
            pc.b = (uint32_t)(ad_b % ADLER_MOD);
            P.pieces[gp] = pc;
        }
        if (blank && !failed && nbytes <= PNG_SLOT) {
            canned.nbits = nbits;
            canned.where = (uint32_t)(where / 16);
            canned.a = (uint32_t)(ad_a % ADLER_MOD);
            canned.b = (uint32_t)(ad_b % ADLER_MOD);
        }
        __syncwarp();
      }
    }
}

// bytes of the deflate stream of an image: block header, the pieces' bits, end of block, padded to a byte
__host__ __device__ inline unsigned long long png_deflate_bytes(unsigned long long piece_bits, uint32_t hdr_bits, uint32_t eob_bits)
{
    return (piece_bits + hdr_bits + eob_bits + 7ull) / 8ull;
}

__global__ void k_png_sizes(const PngPiece *__restrict__ pieces, int n, uint32_t ppi, const PngTables *__restrict__ tables,
                            unsigned long long *__restrict__ sizes)
{
    const int img = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (img >= n) return;
    unsigned long long s = 0;
    for (uint32_t p = lane; p < ppi; p += 32) s += pieces[(size_t)img * ppi + p].nbits;
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) sizes[img] = png_deflate_bytes(s, tables->hdr_bits, tables->eob >> 24) + PNG_FILE_OVERHEAD;
}

// exclusive prefix sum of sizes[0..n) into offsets[0..n]; one CTA of 1024 threads.  offsets_host / counters_host are page-locked
// HOST memory (unified addressing): the host waits for these numbers, and a copy-engine transfer of them would queue behind the
// previous batch's files.
__global__ void __launch_bounds__(1024) k_png_scan(const unsigned long long *__restrict__ sizes, int n, unsigned long long *__restrict__ offsets,
                                                   unsigned long long *__restrict__ offsets_host, const unsigned long long *__restrict__ counters,
                                                   unsigned long long *__restrict__ counters_host)
{
    __shared__ unsigned long long part[1024];
    const int t = threadIdx.x;
    const int per = (n + 1023) / 1024;
    const int b = min(n, t * per), e = min(n, b + per);
    unsigned long long s = 0;
    for (int i = b; i < e; i++) s += sizes[i];
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const unsigned long long v = t >= o ? part[t - o] : 0ull;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    unsigned long long run = t ? part[t - 1] : 0ull;
    for (int i = b; i < e; i++) {
        offsets[i] = run;
        offsets_host[i] = run;
        run += sizes[i];
    }
    if (t == 1023) { offsets[n] = part[1023]; offsets_host[n] = part[1023]; }
    if (t < 2) counters_host[t] = counters[t];
}

// a * b mod P over GF(2), reflected bit order (zlib crc32.c multmodp)
__device__ __forceinline__ uint32_t crc_multmodp(uint32_t a, uint32_t b)
{
    uint32_t p = 0;
#pragma unroll 4
    for (int i = 0; i < 32; i++) {
        if (a & (0x80000000u >> i)) p ^= b;
        b = (b & 1u) ? (b >> 1) ^ CRC_POLY : b >> 1;
    }
    return p;
}
// x^(8 * nbytes) mod P
__device__ __forceinline__ uint32_t crc_x8n(const uint32_t *x2n, unsigned long long nbytes)
{
    uint32_t p = 0x80000000u;
    unsigned k = 3;
    while (nbytes) {
        if (nbytes & 1ull) p = crc_multmodp(x2n[k & 31], p);
        nbytes >>= 1;
        k++;
    }
    return p;
}

struct PngAssembleParams {
    const PngPiece *pieces;
    const uint8_t *scratch;
    const unsigned long long *offsets;
    const PngTables *tables;
    uint8_t *out;
    int n, W, H, ch;
    uint32_t pieces_per_image;
};

__device__ __forceinline__ void put_be32(uint8_t *p, uint32_t v)
{
    p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v;
}

// dynamic shared memory: per warp, (pieces_per_image + 1) bit offsets and pieces_per_image scratch positions
__host__ __device__ inline size_t png_assemble_smem(uint32_t ppi, int warps) { return (size_t)warps * (2 * (size_t)ppi + 1) * sizeof(uint32_t); }

__global__ void __launch_bounds__(256) k_png_assemble(const PngAssembleParams P)
{
    __shared__ uint32_t crc_t[256];
    __shared__ uint32_t x2n[32];
    extern __shared__ uint32_t asm_smem[];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) crc_t[i] = P.tables->crc[i];
    if (threadIdx.x < 32) x2n[threadIdx.x] = P.tables->x2n[threadIdx.x];
    __syncthreads();
    const int warp = threadIdx.x >> 5, img = blockIdx.x * (blockDim.x >> 5) + warp, lane = threadIdx.x & 31;
    if (img >= P.n) return;
    const uint32_t ppi = P.pieces_per_image;
    uint32_t *pre = asm_smem + (size_t)warp * (2 * (size_t)ppi + 1);      // pre[p]: bits of the pieces before p; pre[ppi]: all
    uint32_t *where = pre + ppi + 1;
    const PngPiece *pcs = P.pieces + (size_t)img * ppi;
    uint8_t *file = P.out + P.offsets[img];
    const unsigned long long total = P.offsets[img + 1] - P.offsets[img];
    const uint32_t deflate_bytes = (uint32_t)(total - PNG_FILE_OVERHEAD);
    const uint32_t data_len = deflate_bytes + 6u;                               // zlib header + deflate stream + Adler
    const uint32_t RL = 1u + (uint32_t)(P.ch * P.W);
    const unsigned long long ulen = (unsigned long long)P.H * RL;               // uncompressed (filtered) bytes
    // ---- bit offsets of the pieces, Adler-32 of the whole stream (from the pieces' partial sums)
    unsigned long long sa = 0, sb = 0;
    uint32_t run = 0;
    for (uint32_t base = 0; base < ppi; base += 32u) {
        const uint32_t p = base + (uint32_t)lane;
        PngPiece pc;
        pc.nbits = 0; pc.where = 0; pc.a = 0; pc.b = 0;
        if (p < ppi) {
            pc = pcs[p];
            const unsigned long long rows_before = (unsigned long long)p * PNG_ROWS, rows_p = min((unsigned long long)PNG_ROWS, (unsigned long long)P.H - rows_before);
            const unsigned long long after = ulen - (rows_before + rows_p) * RL;
            sa += pc.a;
            sb += pc.b + (unsigned long long)pc.a * (after % ADLER_MOD);
            sb %= ADLER_MOD;
        }
        uint32_t incl = pc.nbits;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (p < ppi) { pre[p] = run + incl - pc.nbits; where[p] = pc.where; }
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) pre[ppi] = run;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        sa += __shfl_xor_sync(0xffffffffu, sa, o);
        sb += __shfl_xor_sync(0xffffffffu, sb, o);
    }
    const uint32_t adler = ((uint32_t)((ulen + sb) % ADLER_MOD) << 16) | (uint32_t)((1ull + sa) % ADLER_MOD);
    __syncwarp();
    // ---- fixed parts
    for (int i = lane; i < 33; i += 32) file[i] = P.tables->header[i];
    uint8_t *idat = file + 33;
    if (lane == 0) {
        put_be32(idat, data_len);
        idat[4] = 'I'; idat[5] = 'D'; idat[6] = 'A'; idat[7] = 'T';
        idat[8] = 0x78; idat[9] = 0x01;
    }
    uint8_t *body = idat + 10;                                   // first byte of the deflate stream
    uint8_t *tail = body + deflate_bytes;
    if (lane == 31) {
        put_be32(tail, adler);
        const uint8_t iend[12] = {0, 0, 0, 0, 'I', 'E', 'N', 'D', 0xae, 0x42, 0x60, 0x82};
        for (int i = 0; i < 12; i++) tail[8 + i] = iend[i];
    }
    // ---- pass 1: the deflate stream, 32 bits per lane and step: [block header][piece 0][piece 1] ... [end of block][padding]
    {
        const uint32_t H0 = P.tables->hdr_bits, PB = run, eob = P.tables->eob & 0xffffffu, eob_bits = P.tables->eob >> 24;
        const uint32_t n_words = (deflate_bytes + 3u) / 4u;
        uint32_t p = 0;                                           // a piece at or before the lane's position (positions only grow)
        for (uint32_t k = (uint32_t)lane; k < n_words; k += 32u) {
            const uint32_t b0 = 32u * k;
            uint32_t word = 0u;
            if (b0 < H0) word = P.tables->hdr[k];                 // (the words behind the header's last bit are zero-padded)
            uint32_t lo = (b0 > H0 ? b0 : H0) - H0;               // in piece space: bit 0 = first bit of piece 0
            const uint32_t hi_abs = b0 + 32u < H0 + PB ? b0 + 32u : H0 + PB;
            if (hi_abs > H0 && lo < hi_abs - H0) {
                const uint32_t hi = hi_abs - H0;
                // largest p with pre[p] <= lo
                {
                    uint32_t a = p, b = ppi;                      // pre[a] <= lo < pre[b] = PB ... find by bisection
                    while (b - a > 1u) {
                        const uint32_t m = (a + b) >> 1;
                        if (pre[m] <= lo) a = m; else b = m;
                    }
                    p = a;
                }
                uint32_t q = p;
                while (lo < hi) {
                    const uint32_t pe = pre[q + 1];
                    if (pe <= lo) { q++; continue; }              // (an empty piece)
                    const uint32_t take = (hi < pe ? hi : pe) - lo;
                    const uint32_t sbit = lo - pre[q];
                    const uint32_t *src = (const uint32_t *)(P.scratch + (size_t)where[q] * 16);
                    const uint32_t sw = sbit >> 5, ss = sbit & 31u;
                    uint32_t v = __funnelshift_r(src[sw], src[sw + 1], ss);
                    if (take < 32u) v &= (1u << take) - 1u;
                    word |= v << (lo + H0 - b0);
                    lo += take;
                }
            }
            // end of block
            {
                const uint32_t e0 = H0 + PB;
                if (e0 >= b0 && e0 < b0 + 32u) word |= eob << (e0 - b0);
                else if (e0 < b0 && e0 + eob_bits > b0) word |= eob >> (b0 - e0);
            }
            uint8_t *dst = body + 4u * k;
            const uint32_t left = deflate_bytes - 4u * k;
            dst[0] = (uint8_t)word;
            if (left > 1u) dst[1] = (uint8_t)(word >> 8);
            if (left > 2u) dst[2] = (uint8_t)(word >> 16);
            if (left > 3u) dst[3] = (uint8_t)(word >> 24);
        }
    }
    __syncwarp();
    // ---- pass 2: CRC-32 of "IDAT" + data: every lane takes a contiguous segment of the file (lane 0 the odd-sized first one),
    //      then the 32 CRCs are combined in order: crc(A || B) = crc(A) * x^(8 len B) + crc(B)   (zlib crc32_combine)
    const uint32_t crc_len = 4u + data_len;
    const uint32_t seg = crc_len / 32u, first = crc_len - 31u * seg;
    const uint8_t *cp = idat + 4 + (lane == 0 ? 0u : first + (uint32_t)(lane - 1) * seg);
    const uint32_t my_len = lane == 0 ? first : seg;
    uint32_t crc = 0xffffffffu;
    for (uint32_t i = 0; i < my_len; i++) crc = crc_t[(crc ^ cp[i]) & 0xffu] ^ (crc >> 8);
    crc ^= 0xffffffffu;
    uint32_t xpow = crc_x8n(x2n, seg);                       // x^(8 * seg): the right operand of level o is o segments long
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t rc = __shfl_down_sync(0xffffffffu, crc, o);
        if ((lane & (2 * o - 1)) == 0) crc = crc_multmodp(xpow, crc) ^ rc;
        xpow = crc_multmodp(xpow, xpow);
    }
    if (lane == 0) put_be32(tail + 4, crc);
}

}  // namespace rig
