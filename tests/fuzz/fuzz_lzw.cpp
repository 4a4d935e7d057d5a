// Built and run by tests/test_raster_io.py::test_decoders_under_sanitizers (g++ -fsanitize=address,undefined).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include "../../descriptools_b200/csrc/lzw.cuh"
using namespace dtb;
int main(int argc, char **argv) {
    std::mt19937 rng(11);
    size_t total = 0, bad = 0;
    const int iters = argc > 1 ? atoi(argv[1]) : 3000;
    for (int it = 0; it < iters; ++it) {
        size_t n = rng() % 5 == 0 ? rng() % 4 : 1 + rng() % 30000;
        std::vector<uint8_t> raw(n);
        int kind = rng() % 3;
        for (size_t i = 0; i < n; ++i) raw[i] = kind == 0 ? (uint8_t)rng() : kind == 1 ? (uint8_t)((i / 41) & 1) : (uint8_t)("terrain"[i % 7] + (rng() % 8 == 0));
        size_t bound = lzw_encode_bound(n);
        uint8_t *rawp = (uint8_t *)malloc(n ? n : 1);
        if (n) memcpy(rawp, raw.data(), n);
        uint8_t *enc = (uint8_t *)malloc(bound);
        uint64_t *buckets = (uint64_t *)malloc(sizeof(uint32_t) * kLzwHashSlots);
        int64_t m = lzw_encode(rawp, n, enc, bound, buckets, 0, 32, 32);
        if (m <= 0 || (size_t)m > bound) { ++bad; printf("encode size it=%d m=%ld\n", it, (long)m); }
        // too small a slot must be reported, not overrun
        if (n > 100) {
            uint8_t *small = (uint8_t *)malloc(n / 8);
            int64_t m2 = lzw_encode(rawp, n, small, n / 8, buckets, 0, 32, 32);
            if (m2 != -1 && m2 > (int64_t)(n / 8)) { ++bad; printf("encode overrun\n"); }
            free(small);
        }
        // intact round trip, both slot types, also clipped
        for (int wide = 0; wide < 2; ++wide) {
            size_t cap = rng() % 3 == 0 ? n / 2 + 1 : n + (rng() % 2);
            uint8_t *in = (uint8_t *)malloc(m);
            memcpy(in, enc, m);
            uint8_t *out = (uint8_t *)malloc(cap ? cap : 1);
            int64_t got;
            if (wide) { LzwWideSlot *t = (LzwWideSlot *)malloc(sizeof(LzwWideSlot) * kLzwTableSlots); got = lzw_decode(in, m, out, cap, t, 0, 1, 1); free(t); }
            else { LzwPackedSlot *t = (LzwPackedSlot *)malloc(sizeof(LzwPackedSlot) * kLzwTableSlots); got = lzw_decode(in, m, out, cap, t, 0, 32, 32); free(t); }
            size_t want = cap < n ? cap : n;
            if (got != (int64_t)want || (want && memcmp(out, raw.data(), want))) { ++bad; printf("decode mismatch it=%d wide=%d got=%ld want=%zu\n", it, wide, (long)got, want); }
            free(in); free(out);
        }
        // corrupted streams: any outcome but a memory error
        for (int c = 0; c < 3; ++c) {
            size_t len = m;
            uint8_t *in = (uint8_t *)malloc(len);
            memcpy(in, enc, len);
            int flips = 1 + rng() % 4;
            for (int f = 0; f < flips; ++f) in[rng() % len] ^= (uint8_t)(1u << (rng() % 8));
            if (rng() % 4 == 0) len = 1 + rng() % len;
            size_t cap = n ? n : 1;
            uint8_t *out = (uint8_t *)malloc(cap);
            LzwPackedSlot *t = (LzwPackedSlot *)malloc(sizeof(LzwPackedSlot) * kLzwTableSlots);
            int64_t got = lzw_decode(in, len, out, cap, t, 0, 32, 32);
            if (got > (int64_t)cap) { ++bad; printf("overrun report\n"); }
            free(t); free(out); free(in);
        }
        free(rawp); free(enc); free(buckets);
        ++total;
    }
    printf("lzw fuzz: %zu inputs, %zu problems\n", total, bad);
    return bad != 0;
}
