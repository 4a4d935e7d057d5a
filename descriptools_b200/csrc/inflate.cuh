// inflate.cuh -- zlib / Deflate (RFC 1950, RFC 1951) and PackBits decoding for the device tile decoder (tiffcodec.cu)
// and its CPU replay.  GeoTIFFs written with COMPRESS=DEFLATE hold one zlib stream per tile or strip.
//
// Same execution model as the LZW decoder in lzw.cuh: a whole warp runs the decoder in lock step (identical state in
// every lane, uniform control flow, broadcast loads, table stores of the same value from every lane), and the only
// work that is spread over the lanes is the copy of a match -- `byte i = source[i mod distance]`, which also covers
// the overlapping matches Deflate uses for runs.  Huffman codes of up to nine bits are decoded with one table look-up,
// longer ones canonically, one bit per step, from the per-length counts and the symbols sorted by code (about 3 KB of
// tables per warp, in the warp's table area).
// The Adler-32 trailer is not checked: a damaged chunk shows up as a bad code, a bad distance or a short chunk.
// `lane0, lane1, nlanes` as in lzw.cuh.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "lzw.cuh"  // DTB_LZW_HD, DTB_LZW_WARP_SYNC

namespace dtb {

constexpr int kInflateFastBits = 9;  // codes up to this length are decoded with one table look-up

struct InflateScratch {
    uint16_t len_count[16], len_symbol[288];   // literal / length code: codes per length, symbols sorted by code
    uint16_t dist_count[16], dist_symbol[32];  // distance code
    uint16_t len_fast[1 << kInflateFastBits];  // next 9 stream bits -> (symbol << 4) | code length, 0 = a longer code
    uint16_t dist_fast[1 << kInflateFastBits];
    uint16_t offs[16];
    uint8_t lengths[320];                      // code lengths while a dynamic block's header is read
};

struct InflateBits {
    const uint8_t *in;
    size_t n, ip;
    uint64_t buf;
    int cnt;       // bits in buf
    int fake;      // of which zero padding past the end of the input (always the topmost ones)
    bool starved;  // padding bits were consumed: the stream ended early
};

// at least k <= 16 bits in the buffer (least significant first); past the end of the input the buffer is padded with
// zeros, which only counts as an error once such bits are consumed -- so looking ahead near the end is harmless
DTB_LZW_HD void inflate_fill(InflateBits &b, int k)
{
    while (b.cnt < k) {
        if (b.ip < b.n) b.buf |= (uint64_t)b.in[b.ip++] << b.cnt;
        else b.fake += 8;
        b.cnt += 8;
    }
}

DTB_LZW_HD void inflate_drop(InflateBits &b, int k)
{
    b.buf >>= k;
    b.cnt -= k;
    if (b.cnt < b.fake) {
        b.starved = true;
        b.fake = b.cnt;
    }
}

DTB_LZW_HD uint32_t inflate_bits(InflateBits &b, int k)
{
    inflate_fill(b, k);
    const uint32_t v = (uint32_t)(b.buf & ((1u << k) - 1u));
    inflate_drop(b, k);
    return v;
}

// canonical Huffman decoding: codes of one length are consecutive, and the first code of each length follows from the
// counts of the shorter ones.  Returns the symbol or -1 (no such code).
DTB_LZW_HD int inflate_symbol(InflateBits &b, const uint16_t *count, const uint16_t *symbol)
{
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; ++len) {
        code |= (int)inflate_bits(b, 1);
        const int c = count[len];
        if (code - c < first) return symbol[index + (code - first)];
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

// Tables for the code whose symbol i has length[i] bits (0 = unused).  Returns 0 for a complete code, a negative
// number for an over-subscribed one, a positive number for an incomplete one.
DTB_LZW_HD int inflate_build(uint16_t *count, uint16_t *symbol, uint16_t *offs, const uint8_t *length, int n)
{
    for (int len = 0; len <= 15; ++len) count[len] = 0;
    for (int i = 0; i < n; ++i) count[length[i]] = (uint16_t)(count[length[i]] + 1);
    if (count[0] == n) return 0;  // no codes at all: complete, and decoding anything with it fails
    int left = 1;
    for (int len = 1; len <= 15; ++len) {
        left <<= 1;
        left -= count[len];
        if (left < 0) return left;
    }
    offs[1] = 0;
    for (int len = 1; len < 15; ++len) offs[len + 1] = (uint16_t)(offs[len] + count[len]);
    for (int i = 0; i < n; ++i)
        if (length[i] != 0) {
            symbol[offs[length[i]]] = (uint16_t)i;
            offs[length[i]] = (uint16_t)(offs[length[i]] + 1);
        }
    return left;
}

// One-look-up table for the codes of at most kInflateFastBits bits, from the counts and sorted symbols inflate_build
// left: the j-th symbol of length L has the code first(L) + j (RFC 1951, 3.2.2), Huffman codes enter the stream most
// significant bit first, so the index is the bit-reversed code, repeated for every value of the bits that follow it.
DTB_LZW_HD void inflate_build_fast(uint16_t *fast, const uint16_t *count, const uint16_t *symbol)
{
    for (int i = 0; i < (1 << kInflateFastBits); ++i) fast[i] = 0;
    int code = 0, index = 0;
    for (int len = 1; len <= kInflateFastBits; ++len) {
        code <<= 1;  // first code of this length
        for (int j = 0; j < (int)count[len]; ++j) {
            int c = code + j, r = 0;
            for (int k = 0; k < len; ++k) {
                r = (r << 1) | (c & 1);
                c >>= 1;
            }
            const uint16_t e = (uint16_t)((symbol[index + j] << 4) | len);
            for (int i = r; i < (1 << kInflateFastBits); i += 1 << len) fast[i] = e;
        }
        code += count[len];
        index += count[len];
    }
}

// a symbol of the literal / length or distance code: table look-up, canonical decoding for the rare long codes
DTB_LZW_HD int inflate_symbol_fast(InflateBits &b, const uint16_t *fast, const uint16_t *count, const uint16_t *symbol)
{
    inflate_fill(b, kInflateFastBits);
    const uint32_t e = fast[b.buf & ((1u << kInflateFastBits) - 1u)];
    if (e != 0) {
        inflate_drop(b, (int)(e & 15u));
        return (int)(e >> 4);
    }
    return inflate_symbol(b, count, symbol);
}

// order in which the code-length code's own lengths are stored (RFC 1951, 3.2.7), five bits per entry
DTB_LZW_HD int inflate_order(int i)
{
    // 16 17 18 0 8 7 9 6 10 5 11 4 | 12 3 13 2 14 1 15
    const uint64_t lo = 16ull | 17ull << 5 | 18ull << 10 | 0ull << 15 | 8ull << 20 | 7ull << 25 | 9ull << 30 | 6ull << 35 | 10ull << 40 |
                        5ull << 45 | 11ull << 50 | 4ull << 55;
    const uint64_t hi = 12ull | 3ull << 5 | 13ull << 10 | 2ull << 15 | 14ull << 20 | 1ull << 25 | 15ull << 30;
    return i < 12 ? (int)((lo >> (5 * i)) & 31u) : (int)((hi >> (5 * (i - 12))) & 31u);
}

// Decodes at most `cap` bytes of the zlib stream `in` into `out`.  Returns the number of bytes produced or -1 for a
// damaged stream.
DTB_LZW_HD int64_t zlib_inflate(const uint8_t *in, size_t n, uint8_t *out, size_t cap, InflateScratch *t, int lane0 = 0, int lane1 = 1,
                                int nlanes = 1)
{
    if (n < 2) return -1;
    const uint32_t cmf = in[0], flg = in[1];
    if ((cmf & 0x0Fu) != 8u || ((cmf << 8) | flg) % 31u != 0u || (flg & 0x20u)) return -1;  // not Deflate, bad check, preset dictionary
    InflateBits b{in, n, 2, 0, 0, 0, false};
    size_t op = 0;
    bool last = false;
    while (!last && op < cap) {
        // every lane rebuilds the same Huffman tables for a new block: no lane may still be decoding with the old ones
        DTB_LZW_WARP_SYNC();
        last = inflate_bits(b, 1) != 0;
        const uint32_t type = inflate_bits(b, 2);
        if (b.starved) return -1;
        if (type == 0) {  // stored: the rest of the byte is skipped, LEN, ~LEN, then LEN bytes
            b.ip -= (size_t)((b.cnt - b.fake) >> 3);  // whole bytes looked at ahead of time go back to the input
            b.buf = 0;
            b.cnt = 0;
            b.fake = 0;
            if (b.ip + 4 > n) return -1;
            const uint32_t len = in[b.ip] | ((uint32_t)in[b.ip + 1] << 8), nlen = in[b.ip + 2] | ((uint32_t)in[b.ip + 3] << 8);
            if (len != (~nlen & 0xFFFFu)) return -1;
            b.ip += 4;
            if (b.ip + len > n) return -1;
            const size_t k = len < cap - op ? len : cap - op;
            for (int l = lane0; l < lane1; ++l)
                for (size_t i = (size_t)l; i < k; i += (size_t)nlanes) out[op + i] = in[b.ip + i];
            b.ip += len;
            op += k;
            continue;
        }
        if (type == 3) return -1;
        if (type == 1) {  // fixed code
            for (int i = 0; i < 144; ++i) t->lengths[i] = 8;
            for (int i = 144; i < 256; ++i) t->lengths[i] = 9;
            for (int i = 256; i < 280; ++i) t->lengths[i] = 7;
            for (int i = 280; i < 288; ++i) t->lengths[i] = 8;
            inflate_build(t->len_count, t->len_symbol, t->offs, t->lengths, 288);
            for (int i = 0; i < 30; ++i) t->lengths[i] = 5;
            inflate_build(t->dist_count, t->dist_symbol, t->offs, t->lengths, 30);
        } else {  // dynamic code: the code lengths are themselves Huffman coded
            const int nlen = (int)inflate_bits(b, 5) + 257, ndist = (int)inflate_bits(b, 5) + 1, ncode = (int)inflate_bits(b, 4) + 4;
            if (nlen > 286 || ndist > 30) return -1;
            for (int i = 0; i < 19; ++i) t->lengths[i] = 0;
            for (int i = 0; i < ncode; ++i) t->lengths[inflate_order(i)] = (uint8_t)inflate_bits(b, 3);
            if (inflate_build(t->len_count, t->len_symbol, t->offs, t->lengths, 19) != 0) return -1;  // must be complete
            int index = 0;
            while (index < nlen + ndist) {
                const int sym = inflate_symbol(b, t->len_count, t->len_symbol);
                if (sym < 0 || b.starved) return -1;
                if (sym < 16) {
                    t->lengths[index++] = (uint8_t)sym;
                } else {
                    int prev = 0, rep;
                    if (sym == 16) {
                        if (index == 0) return -1;
                        prev = t->lengths[index - 1];
                        rep = 3 + (int)inflate_bits(b, 2);
                    } else if (sym == 17) {
                        rep = 3 + (int)inflate_bits(b, 3);
                    } else {
                        rep = 11 + (int)inflate_bits(b, 7);
                    }
                    if (index + rep > nlen + ndist) return -1;
                    while (rep--) t->lengths[index++] = (uint8_t)prev;
                }
            }
            if (t->lengths[256] == 0) return -1;  // no end-of-block code
            // the distance lengths follow the literal / length ones; build that table first, its lengths are read in place
            int err = inflate_build(t->dist_count, t->dist_symbol, t->offs, t->lengths + nlen, ndist);
            if (err != 0 && (err < 0 || ndist != t->dist_count[0] + t->dist_count[1])) return -1;  // incomplete only with a single code
            err = inflate_build(t->len_count, t->len_symbol, t->offs, t->lengths, nlen);
            if (err != 0 && (err < 0 || nlen != t->len_count[0] + t->len_count[1])) return -1;
        }
        inflate_build_fast(t->len_fast, t->len_count, t->len_symbol);
        inflate_build_fast(t->dist_fast, t->dist_count, t->dist_symbol);
        // the block's symbols
        for (;;) {
            int sym = inflate_symbol_fast(b, t->len_fast, t->len_count, t->len_symbol);
            if (sym < 0 || b.starved) return -1;
            if (sym < 256) {
                out[op++] = (uint8_t)sym;  // every lane stores the same byte
                if (op == cap) return (int64_t)op;
                continue;
            }
            if (sym == 256) break;
            sym -= 257;
            if (sym >= 29) return -1;
            // length: 3..10 directly, then groups of four per extra bit, 258 for the last symbol
            const int lext = sym < 8 || sym == 28 ? 0 : (sym - 4) >> 2;
            const uint32_t len = (sym < 8 ? 3u + (uint32_t)sym : sym == 28 ? 258u : 3u + ((4u + ((uint32_t)sym & 3u)) << lext)) + inflate_bits(b, lext);
            const int ds = inflate_symbol_fast(b, t->dist_fast, t->dist_count, t->dist_symbol);
            if (ds < 0 || ds >= 30) return -1;
            const int dext = ds < 4 ? 0 : (ds >> 1) - 1;
            const size_t dist = (ds < 4 ? 1u + (uint32_t)ds : 1u + ((2u + ((uint32_t)ds & 1u)) << dext)) + inflate_bits(b, dext);
            if (b.starved || dist > op) return -1;
            const size_t k = len < cap - op ? len : cap - op;
            DTB_LZW_WARP_SYNC();  // the source bytes were stored by other lanes
            const uint8_t *s = out + (op - dist);
            uint8_t *d = out + op;
            if (dist >= k) {
                for (int l = lane0; l < lane1; ++l)
                    for (size_t i = (size_t)l; i < k; i += (size_t)nlanes) d[i] = s[i];
            } else {
                for (int l = lane0; l < lane1; ++l)
                    for (size_t i = (size_t)l; i < k; i += (size_t)nlanes) d[i] = s[i % dist];
            }
            op += k;
            if (op == cap) return (int64_t)op;
        }
    }
    return (int64_t)op;
}

// PackBits (TIFF 6.0 section 9; compression 32773): a header byte h, then h + 1 literal bytes (0 <= h <= 127) or one
// byte repeated 1 - h times (-127 <= h <= -1); -128 is a no-op.  Both kinds of run are written by the lanes together.
// Returns the number of bytes produced (input that ends inside a run just ends the output).
DTB_LZW_HD int64_t packbits_decode_lanes(const uint8_t *in, size_t n, uint8_t *out, size_t cap, int lane0 = 0, int lane1 = 1,
                                         int nlanes = 1)
{
    size_t ip = 0, op = 0;
    while (ip < n && op < cap) {
        const int h = (int8_t)in[ip++];
        if (h >= 0) {
            size_t k = (size_t)h + 1;
            if (k > n - ip) k = n - ip;
            const size_t w = k < cap - op ? k : cap - op;
            for (int l = lane0; l < lane1; ++l)
                for (size_t i = (size_t)l; i < w; i += (size_t)nlanes) out[op + i] = in[ip + i];
            ip += k;
            op += w;
        } else if (h != -128) {
            if (ip >= n) break;
            const uint8_t v = in[ip++];
            const size_t k = (size_t)(1 - h);
            const size_t w = k < cap - op ? k : cap - op;
            for (int l = lane0; l < lane1; ++l)
                for (size_t i = (size_t)l; i < w; i += (size_t)nlanes) out[op + i] = v;
            op += w;
        }
    }
    return (int64_t)op;
}

}  // namespace dtb
