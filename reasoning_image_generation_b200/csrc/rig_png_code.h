// rig_png_code.h -- the Huffman code of the device-side PNG writer (host code, no CUDA).
//
// libpng's files are half the size of fixed-Huffman ones on this data because zlib builds a Huffman code per block: on puzzle
// images nearly every token is "258 zeros at distance 1", a literal 0 / 255, or the row's filter byte, and those get 2-4 bit
// codes.  The device encoder cannot afford a histogram pass and a code per image, but it does not need one: the token statistics
// of these images are stable, so ONE code, built here when the context is created from a frequency model of the writer's own
// token stream (Sub-filtered rows; zero runs as literal 0 + matches at distance 1), is shipped in every file as the header of a
// single dynamic-Huffman block (RFC 1951, 3.2.7).  Any inflate reads it; a poor match between model and image only costs bits.
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace rig {

constexpr int PNG_NLIT = 286;          // literal/length symbols 0..285
// Longest code the builder hands out.  The encoder's shared-memory buffers are sized for the worst case (every byte a literal of
// this length); on the puzzle images a 10-bit limit costs 0.1 % of the file size against the deflate maximum of 15
// (tools/png_token_stats.py) and keeps the buffers at the size the fixed-Huffman encoder (9 bits) had.
constexpr int PNG_MAX_CODE_BITS = 10;
constexpr int PNG_HDR_WORDS = 96;      // room for the block header (3 + 14 + 19 * 3 + 287 * 7 bits at most, far less in practice)

struct PngCode {
    uint8_t len[PNG_NLIT];             // code length of every literal/length symbol (1..PNG_MAX_CODE_BITS)
    uint16_t code[PNG_NLIT];           // canonical code, MSB first (as RFC 1951 assigns them)
    uint32_t hdr[PNG_HDR_WORDS];       // BFINAL = 1, BTYPE = 10, the code description; LSB-first bit string
    uint32_t hdr_bits;
};

// Huffman code lengths for freq[0..n) (all > 0), none longer than maxbits, Kraft sum exactly 1.
inline void png_huffman_lengths(const uint64_t *freq, int n, int maxbits, uint8_t *len)
{
    struct Node { uint64_t f; int l, r; };
    std::vector<Node> nodes;
    nodes.reserve(2 * n);
    std::vector<int> alive;
    for (int i = 0; i < n; i++) { nodes.push_back(Node{freq[i], -1, -1}); alive.push_back(i); }
    // O(n^2) pairing is fine for 286 symbols, once per context
    while (alive.size() > 1) {
        std::sort(alive.begin(), alive.end(), [&](int a, int b) { return nodes[a].f != nodes[b].f ? nodes[a].f > nodes[b].f : a > b; });
        const int a = alive.back(); alive.pop_back();
        const int b = alive.back(); alive.pop_back();
        nodes.push_back(Node{nodes[a].f + nodes[b].f, a, b});
        alive.push_back((int)nodes.size() - 1);
    }
    std::vector<int> depth(nodes.size(), 0);
    for (int i = (int)nodes.size() - 1; i >= 0; i--)
        if (nodes[i].l >= 0) { depth[nodes[i].l] = depth[i] + 1; depth[nodes[i].r] = depth[i] + 1; }
    for (int i = 0; i < n; i++) len[i] = (uint8_t)std::max(1, depth[i]);
    // length limit: clamp, then repair the Kraft sum (in units of 2^-maxbits)
    const uint64_t full = 1ull << maxbits;
    uint64_t kraft = 0;
    for (int i = 0; i < n; i++) { if (len[i] > maxbits) len[i] = (uint8_t)maxbits; kraft += full >> len[i]; }
    std::vector<int> by_freq(n);
    for (int i = 0; i < n; i++) by_freq[i] = i;
    std::sort(by_freq.begin(), by_freq.end(), [&](int a, int b) { return freq[a] != freq[b] ? freq[a] < freq[b] : a < b; });
    while (kraft > full) {                       // over-subscribed: lengthen the rarest symbol that still can be
        bool done = false;
        for (int k = 0; k < n && !done; k++) {
            const int s = by_freq[k];
            if (len[s] < maxbits) { kraft -= full >> (len[s] + 1); len[s]++; done = true; }
        }
        if (!done) break;
    }
    while (kraft < full) {                       // incomplete: shorten the most frequent symbol whose gain still fits
        bool done = false;
        for (int k = n - 1; k >= 0 && !done; k--) {
            const int s = by_freq[k];
            if (len[s] > 1 && kraft + (full >> len[s]) <= full) { kraft += full >> len[s]; len[s]--; done = true; }
        }
        if (!done) break;
    }
}

// canonical codes of RFC 1951 3.2.2
inline void png_canonical_codes(const uint8_t *len, int n, uint16_t *code)
{
    int count[16] = {0};
    for (int i = 0; i < n; i++) count[len[i]]++;
    count[0] = 0;
    uint32_t next[16] = {0}, c = 0;
    for (int b = 1; b < 16; b++) { c = (c + count[b - 1]) << 1; next[b] = c; }
    for (int i = 0; i < n; i++) code[i] = len[i] ? (uint16_t)next[len[i]]++ : 0;
}

inline uint32_t png_bitrev(uint32_t v, int n)
{
    uint32_t r = 0;
    for (int i = 0; i < n; i++) r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
}

struct PngBitWriter {
    uint32_t *w;
    uint32_t cap_bits, pos = 0;
    bool overflow = false;
    void put(uint32_t v, int n)      // n <= 16 bits, LSB first
    {
        if (pos + (uint32_t)n > cap_bits) { overflow = true; return; }
        for (int i = 0; i < n; i++, pos++)
            if ((v >> i) & 1u) w[pos >> 5] |= 1u << (pos & 31);
    }
    void put_code(uint32_t code, int n) { put(png_bitrev(code, n), n); }      // Huffman codes go MSB first
};

// The frequency model: token counts of the writer's own stream over a few hundred synthetic single-shape figures (RGBA files,
// 14 156 tokens per image: tools/png_token_stats.py), smoothed so that every symbol has a code.  ch only moves weight between the
// literals 255 (white pixels at the start of a row: 3 or 4 per row) -- the code is not sensitive to it.
inline void png_frequency_model(int ch, uint64_t *f)
{
    for (int s = 0; s < 256; s++) f[s] = 300;                 // anti-aliased stroke pixels: any delta, roughly uniform
    for (int s = 1; s <= 12; s++) { f[s] = 1200; f[256 - s] = 1200; }         // small deltas are commoner
    f[0] = 268000;                                             // the literal that opens a zero run, short zero gaps
    f[255] = ch == 4 ? 257000 : 200000;                        // the row's first pixel (white), alpha of RGBA
    f[1] = 72000;                                              // filter type 1 (Sub) opens every row
    f[256] = 120;                                              // end of block: once per image
    for (int s = 257; s <= 264; s++) f[s] = 2500;              // short matches (3..10 zeros)
    for (int s = 265; s <= 280; s++) f[s] = 4500;
    f[281] = 5300; f[282] = 9000; f[283] = 9700;
    f[284] = 52600;                                            // the remainder of a row after the 258s
    f[285] = 411000;                                           // 258 zeros
}

// false if the header does not fit PNG_HDR_WORDS (cannot happen with the model above)
inline bool png_build_code(int ch, PngCode &C)
{
    uint64_t f[PNG_NLIT];
    png_frequency_model(ch, f);
    png_huffman_lengths(f, PNG_NLIT, PNG_MAX_CODE_BITS, C.len);
    png_canonical_codes(C.len, PNG_NLIT, C.code);
    // ---- the block header.  One distance code (distance 1), length 1: "one distance code of one bit" (RFC 1951 3.2.7).
    uint8_t all[PNG_NLIT + 1];
    memcpy(all, C.len, PNG_NLIT);
    all[PNG_NLIT] = 1;
    // code-length alphabet: lengths 0..15 sent literally (symbols 16-18, the run codes, are not used: 287 short codes cost
    // about 130 bytes per file, 1 % of it)
    uint64_t clf[19] = {0};
    for (int i = 0; i <= PNG_NLIT; i++) clf[all[i]]++;
    uint8_t cll[19] = {0};
    {
        std::vector<uint64_t> fr;
        std::vector<int> sym;
        for (int s = 0; s < 19; s++) if (clf[s]) { fr.push_back(clf[s]); sym.push_back(s); }
        std::vector<uint8_t> l(fr.size());
        if (fr.size() == 1) l[0] = 1;
        else png_huffman_lengths(fr.data(), (int)fr.size(), 7, l.data());
        for (size_t k = 0; k < sym.size(); k++) cll[sym[k]] = l[k];
    }
    uint16_t clc[19];
    png_canonical_codes(cll, 19, clc);
    static const int order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    int hclen = 19;
    while (hclen > 4 && cll[order[hclen - 1]] == 0) hclen--;
    memset(C.hdr, 0, sizeof C.hdr);
    PngBitWriter bw{C.hdr, PNG_HDR_WORDS * 32u};
    bw.put(1, 1);                       // BFINAL
    bw.put(2, 2);                       // BTYPE = 10: dynamic Huffman
    bw.put(PNG_NLIT - 257, 5);          // HLIT
    bw.put(0, 5);                       // HDIST: 1 distance code
    bw.put((uint32_t)(hclen - 4), 4);   // HCLEN
    for (int i = 0; i < hclen; i++) bw.put(cll[order[i]], 3);
    for (int i = 0; i <= PNG_NLIT; i++) bw.put_code(clc[all[i]], cll[all[i]]);
    C.hdr_bits = bw.pos;
    return !bw.overflow;
}

// token -> (bits, count) as the device tables hold them: code | nbits << 24, LSB-first
inline uint32_t png_lut_literal(const PngCode &C, uint32_t v) { return png_bitrev(C.code[v], C.len[v]) | ((uint32_t)C.len[v] << 24); }
inline uint32_t png_lut_match(const PngCode &C, uint32_t length)    // 3..258 at distance 1: length code, extra bits, 1-bit distance code 0
{
    uint32_t sym, eb = 0, extra = 0;
    if (length == 258u) sym = 285u;
    else {
        const uint32_t l = length - 3u;
        if (l < 8u) sym = 257u + l;
        else {
            uint32_t hb = 31u;
            while (!((l >> hb) & 1u)) hb--;
            eb = hb - 2u;
            sym = 257u + 4u * eb + 4u + ((l >> eb) & 3u);
            extra = l & ((1u << eb) - 1u);
        }
    }
    uint32_t bits = png_bitrev(C.code[sym], C.len[sym]), n = C.len[sym];
    bits |= extra << n;
    n += eb;
    n += 1;                              // the distance code: one zero bit
    return bits | (n << 24);
}

}  // namespace rig
