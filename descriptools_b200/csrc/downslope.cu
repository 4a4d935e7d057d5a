// downslope.cu -- downslope index: walk the D8 path from each cell until the drop reaches
// `delta`; index = drop / path length (not in percent).
//
// Semantics: the COMPOSITE the reference's public downsloper() produces -- downslope_gpu
// (downslope.py:458-532) followed by the whole-raster CPU pass over the cells it flagged -50
// (downslope_sequential_jit, downslope.py:194-312, 373-374): walks that leave the raster or
// would step onto nodata return the partial slope to the last valid cell (0 if no move was
// made), walks that reach max_moves (5000, downslope.py:303) return the partial slope.  The
// reference's GPU kernel punts ~10 % of the cells of its own example to that single-threaded
// CPU pass; here one kernel finishes every cell.  The stop condition depends on the start
// cell's elevation (downslope.py:468), so the walk cannot be pointer-jumped; the distance is
// summed move by move in f64 exactly as the reference does (downslope.py:490-513), which
// makes the result bit-identical to oracle/dt_oracle.c.
#include <math.h>

#include "common.cuh"

namespace dtb {
namespace {

constexpr int DS_THREADS = 256;

template <typename T> struct Diff;
template <> struct Diff<float> {
    static __device__ __forceinline__ double sub(float a, float b) { return (double)(a - b); }  // f32 - f32 -> f32
};
template <> struct Diff<int16_t> {
    static __device__ __forceinline__ double sub(int16_t a, int16_t b) { return (double)((int)a - (int)b); }  // i64 in Numba
};

// open_above / open_below: the buffer is a window of a larger raster (a row band plus halo rows) and its first / last row
// is NOT the raster's edge: a walk that would leave through it cannot be finished here -- the cell gets the reference's
// own "redo me" marker -50 (downslope.py:526-529) and *escaped counts it, for the band driver to widen the window.
template <typename T>
__global__ void __launch_bounds__(DS_THREADS)
downslope_kernel(const T *__restrict__ dem, const uint8_t *__restrict__ fdr, int64_t rows, int64_t cols, int64_t row_begin,
                 int64_t row_end, double px, double pd, double delta, int64_t max_moves, float *__restrict__ out, int open_above,
                 int open_below, unsigned long long *__restrict__ escaped)
{
    // cells of rows [row_begin, row_end) are computed (walks may wander over the whole raster); out starts at row_begin
    const int64_t n = (row_end - row_begin) * cols;
    const int64_t o = (int64_t)blockIdx.x * DS_THREADS + threadIdx.x;
    if (o >= n) return;
    const int64_t i = o + row_begin * cols;
    const T z0 = dem[i];
    if (z0 <= (T)ND_I) {  // downslope.py:459
        out[o] = ND_F;
        return;
    }
    int64_t y = i / cols, x = i - y * cols, pos = i, loop = 0;
    double dist = 0.0;
    T zc = z0;
    while (Diff<T>::sub(z0, zc) < delta) {  // downslope.py:211 / :468
        const unsigned f = fdr[pos];
        int dr, dc;
        if (d8_offset(f, dr, dc)) {
            const int64_t yy = y + dr, xx = x + dc;
            if ((yy < 0 && open_above) || (yy >= rows && open_below)) {  // leaves the window, not the raster
                out[o] = -50.0f;
                atomicAdd(escaped, 1ull);
                return;
            }
            if (yy < 0 || yy >= rows || xx < 0 || xx >= cols) break;  // downslope.py:212-231
            const int64_t q = yy * cols + xx;
            const T zq = dem[q];
            if (zq == (T)ND_I) break;  // downslope.py:234-276: do not step onto nodata
            y = yy; x = xx; pos = q; zc = zq;
            dist += d8_is_diag(f) ? pd : px;
        }
        if (++loop == max_moves) break;  // downslope.py:300-304
    }
    // downslope.py:306-312 (dist == 0 also covers the reference's 0/0 ZeroDivisionError case)
    out[o] = (dist == 0.0) ? 0.0f : (float)(Diff<T>::sub(z0, zc) / dist);
}

}  // namespace
}  // namespace dtb

extern "C" int dtb_downslope_window(const void *dem, int dem_dtype, const uint8_t *fdr, int64_t rows, int64_t cols, int64_t row_begin,
                                    int64_t row_end, double px, double delta, int64_t max_moves, float *out, int open_above,
                                    int open_below, unsigned long long *escaped, void *stream)
{
    using namespace dtb;
    if (!dem || !fdr || !out || rows <= 0 || cols <= 0 || !(px > 0.0) || row_begin < 0 || row_end > rows || row_begin > row_end)
        return DTB_ERR_INVALID;
    if ((open_above || open_below) && !escaped) return DTB_ERR_INVALID;
    if (row_begin == row_end) return DTB_OK;
    if (max_moves <= 0) max_moves = 5000;
    const int64_t n = (row_end - row_begin) * cols;
    const unsigned blocks = (unsigned)((n + DS_THREADS - 1) / DS_THREADS);
    cudaStream_t st = as_stream(stream);
    const double pd = px * sqrt(2.0);
    if (dem_dtype == DTB_F32)
        DTB_KERNEL("downslope_kernel<f32>", st, downslope_kernel<float><<<blocks, DS_THREADS, 0, st>>>((const float *)dem, fdr, rows, cols, row_begin, row_end, px, pd, delta, max_moves, out, open_above, open_below, escaped));
    else if (dem_dtype == DTB_I16)
        DTB_KERNEL("downslope_kernel<i16>", st, downslope_kernel<int16_t><<<blocks, DS_THREADS, 0, st>>>((const int16_t *)dem, fdr, rows, cols, row_begin, row_end, px, pd, delta, max_moves, out, open_above, open_below, escaped));
    else
        return DTB_ERR_INVALID;
    return DTB_OK;
}

extern "C" int dtb_downslope_rows(const void *dem, int dem_dtype, const uint8_t *fdr, int64_t rows, int64_t cols, int64_t row_begin,
                                  int64_t row_end, double px, double delta, int64_t max_moves, float *out, void *stream)
{
    return dtb_downslope_window(dem, dem_dtype, fdr, rows, cols, row_begin, row_end, px, delta, max_moves, out, 0, 0, nullptr, stream);
}

extern "C" int dtb_downslope(const void *dem, int dem_dtype, const uint8_t *fdr, int64_t rows, int64_t cols, double px,
                             double delta, int64_t max_moves, float *out, void *stream)
{
    return dtb_downslope_window(dem, dem_dtype, fdr, rows, cols, 0, rows, px, delta, max_moves, out, 0, 0, nullptr, stream);
}
