// lzw.cuh -- TIFF-flavoured LZW decoding, one source for the host codec (geotiff.cpp) and the device tile
// decoder (tiffcodec.cu).
//
// TIFF 6.0 section 13: MSB-first codes of 9..12 bits, Clear = 256, EndOfInformation = 257, the code width
// grows one code early.  The string table does not store strings: every string the coder can name has
// already been written to the output, so an entry is (offset of an earlier occurrence, length) and emitting
// a code is a copy within the output buffer.  The entry made after reading a code is "previous string + first
// byte of this one", which is exactly the bytes that start where the previous string was emitted -- one byte
// longer.  That also covers the code-not-yet-in-table case: the string then overlaps its own output with
// period p = (output position - source), and byte i of it is source byte i mod p.
//
// On the device a whole warp runs the decoder in lock step (identical state in every lane, so control flow is
// uniform and the loads are broadcasts): the recurrence is serial, but the copy of a string is spread over the
// lanes, and the packed table slot carries the first four bytes of its string so that short strings -- the
// common case in terrain data -- cost one table read and no read of the output.  `lane0, lane1, nlanes` name the
// lanes the caller plays: (lane, lane + 1, 32) in the kernel, (0, 1, 1) in the host codec, (0, 32, 32) in the
// CPU self-test that replays the kernel's lane code.
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define DTB_LZW_HD __host__ __device__ __forceinline__
#else
#define DTB_LZW_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define DTB_LZW_WARP_SYNC() __syncwarp()
#else
#define DTB_LZW_WARP_SYNC() ((void)0)
#endif

namespace dtb {

constexpr int kLzwTableSlots = 4096;

// host codec: chunks of any size, strings always copied from the output
struct LzwWideSlot {
    uint32_t off, len;
    static constexpr bool kInline = false;
    DTB_LZW_HD void set(uint32_t o, uint32_t l, uint32_t) { off = o; len = l; }
    DTB_LZW_HD uint32_t offset() const { return off; }
    DTB_LZW_HD uint32_t length() const { return len; }
    DTB_LZW_HD uint32_t first4() const { return 0; }
};

// device: 8 bytes, [31..12 offset | 11..0 length] + the string's first four bytes (little-endian);
// chunks of at most 1 MiB decoded (a string is at most 3 839 bytes long)
struct alignas(8) LzwPackedSlot {
    uint32_t ol, f4;
    static constexpr bool kInline = true;
    static constexpr size_t kMaxChunkBytes = (size_t)1 << 20;
    DTB_LZW_HD void set(uint32_t o, uint32_t l, uint32_t f) { ol = (o << 12) | l; f4 = f; }
    DTB_LZW_HD uint32_t offset() const { return ol >> 12; }
    DTB_LZW_HD uint32_t length() const { return ol & 0xFFFu; }
    DTB_LZW_HD uint32_t first4() const { return f4; }
};

DTB_LZW_HD uint32_t lzw_bswap32(uint32_t v)
{
    return (v >> 24) | ((v >> 8) & 0xFF00u) | ((v << 8) & 0xFF0000u) | (v << 24);
}

// Decodes at most `cap` bytes into `out`.  `tab` holds kLzwTableSlots slots of scratch.
// Returns the number of bytes produced, -1 for a corrupt stream, -2 for the pre-6.0 LSB-first variant.
template <class Slot>
DTB_LZW_HD int64_t lzw_decode(const uint8_t *in, size_t n, uint8_t *out, size_t cap, Slot *tab, int lane0 = 0, int lane1 = 1,
                              int nlanes = 1)
{
    if (n >= 2 && in[0] == 0 && (in[1] & 1)) return -2;
    uint64_t acc = 0;
    int have = 0;
    size_t ip = 0, op = 0;
    int nbits = 9, next = 258;
    bool fresh = true;  // no previous string (start, or just after a Clear)
    size_t prev_off = 0, prev_len = 0;
    uint32_t prev4 = 0;  // first four bytes of the previous string, byte 0 in bits 7..0
    while (op < cap) {
        if (have < nbits) {
            if (ip + 4 <= n && (((uintptr_t)(in + ip)) & 3u) == 0) {  // have < 12: 32 more bits fit
                acc = (acc << 32) | lzw_bswap32(*reinterpret_cast<const uint32_t *>(in + ip));
                ip += 4;
                have += 32;
            } else {
                while (have < nbits) {
                    if (ip >= n) return (int64_t)op;  // input exhausted: hand back what there is
                    acc = (acc << 8) | in[ip++];
                    have += 8;
                }
            }
        }
        const int code = (int)((acc >> (have - nbits)) & ((1u << nbits) - 1u));
        have -= nbits;
        if (code == 257) break;
        if (code == 256) {
            // every lane rewrites the same table slots after a Clear: no lane may still be reading the old generation
            DTB_LZW_WARP_SYNC();
            nbits = 9;
            next = 258;
            fresh = true;
            continue;
        }
        if (fresh) {
            if (code > 255) return -1;
            out[op] = (uint8_t)code;  // every lane stores the same byte
            prev_off = op;
            prev_len = 1;
            prev4 = (uint32_t)code;
            ++op;
            fresh = false;
            continue;
        }
        size_t src, len;
        uint32_t cur4;
        if (code < 256) {
            src = op;
            len = 1;
            cur4 = (uint32_t)code;
        } else if (code < next) {
            const Slot e = tab[code];
            src = e.offset();
            len = e.length();
            cur4 = e.first4();
        } else if (code == next && next < kLzwTableSlots) {
            src = prev_off;
            len = prev_len + 1;
            cur4 = prev_len < 4 ? (prev4 | ((prev4 & 0xFFu) << (8 * prev_len))) : prev4;
        } else {
            return -1;
        }
        if (next < kLzwTableSlots) {
            const uint32_t new4 = prev_len < 4 ? (prev4 | ((cur4 & 0xFFu) << (8 * prev_len))) : prev4;
            tab[next].set((uint32_t)prev_off, (uint32_t)(prev_len + 1), new4);  // every lane stores the same slot
            ++next;
        }
        if (next >= (1 << nbits) - 1 && nbits < 12) ++nbits;
        const size_t k = len < cap - op ? len : cap - op;
        uint8_t *d = out + op;
        if (code < 256) {
            d[0] = (uint8_t)code;
        } else if (Slot::kInline && len <= 4) {
            for (int l = lane0; l < lane1; ++l)
                for (size_t i = (size_t)l; i < k; i += (size_t)nlanes) d[i] = (uint8_t)(cur4 >> (8 * i));
        } else {
            DTB_LZW_WARP_SYNC();  // the source bytes were stored by other lanes
            const uint8_t *s = out + src;
            const size_t p = op - src;  // >= 1; a string longer than p repeats with that period
            if (p >= k) {
                for (int l = lane0; l < lane1; ++l)
                    for (size_t i = (size_t)l; i < k; i += (size_t)nlanes) d[i] = s[i];
            } else {
                for (int l = lane0; l < lane1; ++l)
                    for (size_t i = (size_t)l; i < k; i += (size_t)nlanes) d[i] = s[i % p];
            }
        }
        prev_off = op;
        prev_len = len;
        prev4 = cur4;
        op += k;
    }
    return (int64_t)op;
}

// ---------------------------------------------------------------------------------------------------------
// Encoder for the device (and its CPU replay).  Any parse into table strings is a valid stream as long as the
// code numbering stays in step with the decoder's (one new entry per code), so the dictionary may forget:
// it is a 2-way bucketed hash of (prefix code, byte) -> code that overwrites on conflict.  A slot is 32 bits,
// [31..12 prefix << 8 | byte | 11..0 code], all ones = empty (code 4095 is never assigned); a bucket is one aligned
// 64-bit word, so a lookup is one load.  `tab` holds kLzwHashSlots slots (32 KB: small enough to live in shared
// memory, one table per warp) and is wiped by the lanes together at every table reset.
// ---------------------------------------------------------------------------------------------------------
constexpr int kLzwHashSlots = 8192;

// worst case: every byte its own 12-bit code, plus Clear / EOI codes and the final flush
DTB_LZW_HD size_t lzw_encode_bound(size_t n) { return n + n / 2 + n / 2048 + 16; }

DTB_LZW_HD void lzw_hash_wipe(uint64_t *buckets, int lane0, int lane1, int nlanes)
{
    DTB_LZW_WARP_SYNC();  // every earlier store to the table is ordered before the wipe
    for (int l = lane0; l < lane1; ++l)
        for (int i = l; i < kLzwHashSlots / 2; i += nlanes) buckets[i] = ~(uint64_t)0;
    DTB_LZW_WARP_SYNC();  // and the wipe before the lookups that follow
}

// Returns the number of bytes written to `out`, or -1 if `cap` is too small.  `buckets`: kLzwHashSlots / 2 words of
// 64 bits (two slots each).  [lane0, lane1) of nlanes as for lzw_decode.
DTB_LZW_HD int64_t lzw_encode(const uint8_t *in, size_t n, uint8_t *out, size_t cap, uint64_t *buckets, int lane0 = 0, int lane1 = 1,
                              int nlanes = 1)
{
    uint64_t acc = 0;
    int have = 0;
    size_t op = 0;
    bool overflow = false;
    int nbits = 9, next = 258, maxcode = 511;
    auto put = [&](int code) {
        acc = (acc << nbits) | (uint64_t)code;
        have += nbits;
        while (have >= 8) {
            if (op < cap) out[op] = (uint8_t)(acc >> (have - 8));  // every lane stores the same byte
            else overflow = true;
            ++op;
            have -= 8;
        }
    };
    lzw_hash_wipe(buckets, lane0, lane1, nlanes);
    put(256);
    if (n > 0) {
        int ent = in[0];
        for (size_t i = 1; i < n; ++i) {
            const uint32_t c = in[i];
            const uint32_t key = ((uint32_t)ent << 8) | c;
            const uint32_t b = ((key * 2654435761u) >> 20) & (uint32_t)(kLzwHashSlots / 2 - 1);
            const uint64_t pair = buckets[b];
            const uint32_t s0 = (uint32_t)pair, s1 = (uint32_t)(pair >> 32);
            if ((s0 >> 12) == key && s0 != ~0u) { ent = (int)(s0 & 0xFFFu); continue; }
            if ((s1 >> 12) == key && s1 != ~0u) { ent = (int)(s1 & 0xFFFu); continue; }
            put(ent);
            // an empty way first, else the one the byte picks (the older entry is forgotten)
            const uint32_t way = s0 == ~0u ? 0u : s1 == ~0u ? 1u : (c & 1u);
            const uint64_t slot = (uint64_t)((key << 12) | (uint32_t)next);
            buckets[b] = way ? ((pair & 0xFFFFFFFFull) | (slot << 32)) : ((pair & ~0xFFFFFFFFull) | slot);  // same word from every lane
            ent = (int)c;
            ++next;
            if (next == 4094) {  // table full: Clear, and forget everything
                put(256);
                lzw_hash_wipe(buckets, lane0, lane1, nlanes);
                nbits = 9;
                next = 258;
                maxcode = 511;
            } else if (next > maxcode) {
                ++nbits;
                maxcode = (1 << nbits) - 1;
            }
        }
        put(ent);
        // the decoder makes one more entry for that code before it reads EOI
        ++next;
        if (next == 4094) {
            put(256);
            nbits = 9;
        } else if (next > maxcode) {
            ++nbits;
        }
    }
    put(257);
    if (have > 0) {
        if (op < cap) out[op] = (uint8_t)((acc << (8 - have)) & 0xFFu);
        else overflow = true;
        ++op;
    }
    return overflow ? -1 : (int64_t)op;
}

}  // namespace dtb
