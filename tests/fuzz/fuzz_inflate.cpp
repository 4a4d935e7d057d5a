// Built and run by tests/test_raster_io.py::test_decoders_under_sanitizers (g++ -fsanitize=address,undefined).
// ASAN/UBSAN fuzz of the shared device/host decoders with corrupted streams (exact-size heap buffers)
#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include "../../descriptools_b200/csrc/inflate.cuh"
using namespace dtb;
int main(int argc, char **argv) {
    std::mt19937 rng(7);
    size_t total = 0, bad = 0;
    const int iters = argc > 1 ? atoi(argv[1]) : 3000;
    for (int it = 0; it < iters; ++it) {
        size_t n = 1 + rng() % 20000;
        std::vector<uint8_t> raw(n);
        int kind = rng() % 3;
        for (size_t i = 0; i < n; ++i) raw[i] = kind == 0 ? (uint8_t)rng() : kind == 1 ? (uint8_t)((i / 37) & 3) : (uint8_t)("terrain"[i % 7] + (rng() % 8 == 0));
        // ---- deflate
        uLongf cl = compressBound(n);
        std::vector<uint8_t> comp(cl);
        compress2(comp.data(), &cl, raw.data(), n, 1 + rng() % 9);
        comp.resize(cl);
        int flips = rng() % 4;  // 0 = intact
        for (int f = 0; f < flips; ++f) comp[rng() % comp.size()] ^= (uint8_t)(1u << (rng() % 8));
        if (rng() % 5 == 0) comp.resize(1 + rng() % comp.size());
        size_t cap = rng() % 3 == 0 ? n / 2 + 1 : n;
        uint8_t *in = (uint8_t *)malloc(comp.size());  // exact-size heap blocks: ASAN sees any overrun
        memcpy(in, comp.data(), comp.size());
        uint8_t *out = (uint8_t *)malloc(cap);
        InflateScratch *t = (InflateScratch *)malloc(sizeof(InflateScratch));
        int64_t got = zlib_inflate(in, comp.size(), out, cap, t, 0, 32, 32);
        if (flips == 0 && comp.size() == cl) {
            if (got != (int64_t)cap || memcmp(out, raw.data(), cap)) { ++bad; printf("inflate mismatch it=%d got=%ld cap=%zu\n", it, (long)got, cap); }
        }
        if (got > (int64_t)cap) { ++bad; printf("inflate overrun report\n"); }
        free(in); free(out); free(t);
        ++total;
    }
    printf("inflate fuzz: %zu streams, %zu problems\n", total, bad);
    return bad != 0;
}
