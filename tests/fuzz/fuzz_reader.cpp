// Built and run by tests/test_raster_io.py::test_reader_under_sanitizers (g++ -fsanitize=address,undefined):
// small valid TIFFs (every layout / codec the writer produces) with random bytes flipped anywhere -- header, IFD, tag
// payloads, chunk data -- are opened and read; any outcome but a memory error or an uncaught exception is fine.
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../descriptools_b200/csrc/geotiff.cpp"

int main(int argc, char **argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 2000;
    const std::string dir = argc > 2 ? argv[2] : "/tmp";
    std::mt19937 rng(5);
    const std::string good = dir + "/fz_good.tif", bad = dir + "/fz_bad.tif";
    size_t opened = 0, read_ok = 0;
    for (int it = 0; it < iters; ++it) {
        dtbio_info info{};
        info.rows = 1 + rng() % 70;
        info.cols = 1 + rng() % 70;
        info.dtype = rng() % 10;
        const int comps[3] = {DTBIO_COMP_NONE, DTBIO_COMP_LZW, DTBIO_COMP_DEFLATE};
        info.compression = comps[rng() % 3];
        info.predictor = 1 + rng() % 2;
        if ((info.dtype == DTBIO_F32 || info.dtype == DTBIO_F64) && rng() % 2) info.predictor = 3;
        if (rng() % 2) info.tile_rows = info.tile_cols = 16 << (rng() % 2);
        else info.rows_per_strip = rng() % 9;
        info.bigtiff = rng() % 4 == 0;
        dtbio_writer *w = nullptr;
        if (dtbio_create(good.c_str(), &info, &w) != DTBIO_OK) { printf("create failed: %s\n", dtbio_last_error()); return 1; }
        const size_t bps = (size_t)dtbio_dtype_size(info.dtype);
        std::vector<uint8_t> px((size_t)info.rows * info.cols * bps);
        for (auto &b : px) b = (uint8_t)(rng() % 7);
        dtbio_set_tag(w, 42113, 2, 5, "-100");
        if (dtbio_write_rows(w, 0, info.rows, px.data(), info.cols * bps, 1 + rng() % 3) != DTBIO_OK ||
            dtbio_close_writer(w) != DTBIO_OK) { printf("write failed: %s\n", dtbio_last_error()); return 1; }
        // read the file back, damage it, write the damaged copy
        FILE *f = fopen(good.c_str(), "rb");
        std::vector<uint8_t> bytes;
        uint8_t buf[4096];
        size_t k;
        while ((k = fread(buf, 1, sizeof buf, f)) > 0) bytes.insert(bytes.end(), buf, buf + k);
        fclose(f);
        const int flips = rng() % 6;  // 0 = intact: must read back exactly
        for (int i = 0; i < flips; ++i) {
            // favour the structural parts: header at the front, IFD and tag payloads at the back
            size_t pos = rng() % 3 == 0 ? rng() % std::min<size_t>(bytes.size(), 16) : rng() % 3 == 0 ? bytes.size() - 1 - rng() % std::min<size_t>(bytes.size(), 400) : rng() % bytes.size();
            bytes[pos] = rng() % 2 ? (uint8_t)rng() : (uint8_t)(bytes[pos] ^ (1u << (rng() % 8)));
        }
        const bool truncated = rng() % 8 == 0;
        if (truncated) bytes.resize(rng() % bytes.size());
        const bool intact = flips == 0 && !truncated;
        f = fopen(bad.c_str(), "wb");
        fwrite(bytes.data(), 1, bytes.size(), f);
        fclose(f);
        dtbio_reader *r = nullptr;
        if (dtbio_open(bad.c_str(), &r) != DTBIO_OK) {
            if (intact) { printf("intact file failed to open (it=%d): %s\n", it, dtbio_last_error()); return 1; }
            continue;
        }
        ++opened;
        dtbio_info got{};
        dtbio_get_info(r, &got);
        const size_t gb = (size_t)dtbio_dtype_size(got.dtype);
        if (got.rows > 0 && got.cols > 0 && (double)got.rows * got.cols * gb < 64e6) {
            // exact-size heap block: the sanitizer sees any write outside it
            uint8_t *out = (uint8_t *)malloc((size_t)got.rows * got.cols * gb);
            const int rc = dtbio_read_rows(r, 0, got.rows, out, got.cols * gb, 1 + rng() % 3);
            if (rc == DTBIO_OK) {
                ++read_ok;
                if (intact && (got.rows != info.rows || got.cols != info.cols || memcmp(out, px.data(), px.size()))) {
                    printf("intact file read back wrong (it=%d)\n", it);
                    return 1;
                }
            } else if (intact) {
                printf("intact file failed to read (it=%d): %s\n", it, dtbio_last_error());
                return 1;
            }
            free(out);
        }
        dtbio_close_reader(r);
    }
    unlink(good.c_str());
    unlink(bad.c_str());
    printf("reader fuzz: %d files, %zu opened, %zu read, 0 problems\n", iters, opened, read_ok);
    return 0;
}
