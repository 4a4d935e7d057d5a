// lzw.cuh -- TIFF-flavoured LZW decoding, one source for the host codec (geotiff.cpp) and the device tile
// decoder (tiffdecode.cu).
//
// TIFF 6.0 section 13: MSB-first codes of 9..12 bits, Clear = 256, EndOfInformation = 257, the code width
// grows one code early.  The string table does not store strings: every string the coder can name has
// already been written to the output, so an entry is (offset of an earlier occurrence, length) and emitting
// a code is a copy within the output buffer.  The entry made after reading a code is "previous string + first
// byte of this one", which is exactly the bytes that start where the previous string was emitted -- one byte
// longer.  That also covers the code-not-yet-in-table case (the copy then overlaps its own output, so bytes
// are moved front to back).
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define DTB_LZW_HD __host__ __device__ __forceinline__
#else
#define DTB_LZW_HD inline
#endif

namespace dtb {

struct LzwSlot {
    uint32_t off;  // where in the output an occurrence of the string starts
    uint32_t len;
};

constexpr int kLzwTableSlots = 4096;

// Decodes at most `cap` bytes into `out`.  `tab` holds kLzwTableSlots slots of scratch.
// Returns the number of bytes produced, -1 for a corrupt stream, -2 for the pre-6.0 LSB-first variant.
DTB_LZW_HD int64_t lzw_decode(const uint8_t *in, size_t n, uint8_t *out, size_t cap, LzwSlot *tab)
{
    if (n >= 2 && in[0] == 0 && (in[1] & 1)) return -2;
    uint64_t acc = 0;
    int have = 0;
    size_t ip = 0, op = 0;
    int nbits = 9, next = 258;
    bool fresh = true;             // no previous string (start, or just after a Clear)
    size_t prev_off = 0, prev_len = 0;
    while (op < cap) {
        while (have < nbits) {
            if (ip >= n) return (int64_t)op;  // input exhausted: hand back what there is
            acc = (acc << 8) | in[ip++];
            have += 8;
        }
        const int code = (int)((acc >> (have - nbits)) & ((1u << nbits) - 1u));
        have -= nbits;
        if (code == 257) break;
        if (code == 256) {
            nbits = 9;
            next = 258;
            fresh = true;
            continue;
        }
        if (fresh) {
            if (code > 255) return -1;
            out[op] = (uint8_t)code;
            prev_off = op;
            prev_len = 1;
            ++op;
            fresh = false;
            continue;
        }
        size_t src, len;
        if (code < 256) {
            src = op;  // literal, written below
            len = 1;
        } else if (code < next) {
            src = tab[code].off;
            len = tab[code].len;
        } else if (code == next && next < kLzwTableSlots) {
            src = prev_off;
            len = prev_len + 1;
        } else {
            return -1;
        }
        if (next < kLzwTableSlots) {
            tab[next].off = (uint32_t)prev_off;
            tab[next].len = (uint32_t)(prev_len + 1);
            ++next;
        }
        if (next >= (1 << nbits) - 1 && nbits < 12) ++nbits;
        size_t k = len < cap - op ? len : cap - op;
        if (code < 256) {
            out[op] = (uint8_t)code;
        } else {
            const uint8_t *s = out + src;
            uint8_t *d = out + op;
            for (size_t i = 0; i < k; ++i) d[i] = s[i];  // front to back: the ranges may overlap
        }
        prev_off = op;
        prev_len = len;
        op += k;
    }
    return (int64_t)op;
}

}  // namespace dtb
