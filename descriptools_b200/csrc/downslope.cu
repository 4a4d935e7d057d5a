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
    // "the drop is still below delta" (downslope.py:211 / :468: (double)(a - b) < delta) without a conversion per move:
    // for an f32 d, d < delta  <=>  d < the smallest f32 that is >= delta (host: drop_limit_f32)
    typedef float Limit;
    static __device__ __forceinline__ bool below(float a, float b, float lim) { return a - b < lim; }
};
template <> struct Diff<int16_t> {
    static __device__ __forceinline__ double sub(int16_t a, int16_t b) { return (double)((int)a - (int)b); }  // i64 in Numba
    typedef int Limit;  // for an integer d, d < delta  <=>  d < ceil(delta) (host: drop_limit_i16)
    static __device__ __forceinline__ bool below(int16_t a, int16_t b, int lim) { return (int)a - (int)b < lim; }
};

// open_above / open_below: the buffer is a window of a larger raster (a row band plus halo rows) and its first / last row
// is NOT the raster's edge: a walk that would leave through it cannot be finished here -- the cell gets the reference's
// own "redo me" marker -50 (downslope.py:526-529) and *escaped counts it, for the band driver to widen the window.
//
// The kernel is bound by instruction issue (ncu, profiles/r2zz_extras_10k_summary.txt: 80 % of the issue slots, ~100
// instructions per move in its first form), not by the two dependent gathers of a move -- walking two or four cells per
// thread in step, to have more gathers in flight, made it slower (63 / 75 / 112 ms for 1 / 2 / 4 at 10k x 10k against
// 46.5 for this loop).  So the move is kept short: 32-bit row / column / position arithmetic (IX = int32 whenever the
// window has fewer than 2^31 cells), what a code does read from an 8-entry shared table, the drop test as one compare in
// the DEM's own type; the only f64 left in the loop is the distance sum the reference makes (downslope.py:490-513).
// 10k x 10k, delta = 5 m: 72.9 -> 46.5 ms.
struct alignas(16) Move {
    int off;
    short dr, dc;
    double step;
};

template <typename T, typename IX>
__global__ void __launch_bounds__(DS_THREADS)
downslope_kernel(const T *__restrict__ dem, const uint8_t *__restrict__ fdr, int rows, int cols, int64_t row_begin,
                 int64_t row_end, double px, double pd, typename Diff<T>::Limit limit, int max_moves, float *__restrict__ out,
                 int open_above, int open_below, unsigned long long *__restrict__ escaped)
{
    // what a move does, by the bit of its D8 code: position offset, row / column step, path length (one 16-byte shared load)
    __shared__ Move mv[8];
    if (threadIdx.x < 8) {
        int dr, dc;
        d8_offset(1u << threadIdx.x, dr, dc);
        Move m;
        m.off = dr * cols + dc;
        m.dr = (short)dr;
        m.dc = (short)dc;
        m.step = (dr != 0 && dc != 0) ? pd : px;
        mv[threadIdx.x] = m;
    }
    __syncthreads();
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
    int y = (int)(i / cols), x = (int)(i - (int64_t)y * cols), loop = 0;
    IX pos = (IX)i;
    double dist = 0.0;
    T zc = z0;
    while (Diff<T>::below(z0, zc, limit)) {  // downslope.py:211 / :468
        const unsigned f = fdr[pos];
        // a cell without a direction code never moves: the reference spins here until max_moves (downslope.py:300-304)
        // and returns what it has -- so does leaving the loop at once
        const int b = 31 - __clz((int)f);  // -1 for 0
        if (f != (1u << (b & 31))) break;   // not one bit of 1..128 (f is a byte)
        const Move m = mv[b];               // 0=E 1=SE 2=S 3=SW 4=W 5=NW 6=N 7=NE
        const int yy = y + m.dr, xx = x + m.dc;
        if ((unsigned)yy >= (unsigned)rows) {
            if (yy < 0 ? open_above : open_below) {  // leaves the window, not the raster
                out[o] = -50.0f;
                atomicAdd(escaped, 1ull);
                return;
            }
            break;  // downslope.py:212-231
        }
        if ((unsigned)xx >= (unsigned)cols) break;
        const IX q = pos + (IX)m.off;
        const T zq = dem[q];
        if (zq == (T)ND_I) break;  // downslope.py:234-276: do not step onto nodata
        y = yy; x = xx; pos = q; zc = zq;
        dist += m.step;
        if (++loop == max_moves) break;  // downslope.py:300-304
    }
    // downslope.py:306-312 (dist == 0 also covers the reference's 0/0 ZeroDivisionError case)
    out[o] = (dist == 0.0) ? 0.0f : (float)(Diff<T>::sub(z0, zc) / dist);
}

// smallest float that is >= delta (NaN stays NaN: the walk never starts, as with the f64 comparison)
inline float drop_limit_f32(double delta)
{
    float t = (float)delta;
    if ((double)t < delta) t = nextafterf(t, INFINITY);
    return t;
}
// ceil(delta) clamped to what an i16 difference can reach (|d| <= 65535); NaN: never below
inline int drop_limit_i16(double delta)
{
    if (delta != delta) return -(1 << 30);
    if (delta > 70000.0) return 1 << 30;
    if (delta < -70000.0) return -(1 << 30);
    return (int)ceil(delta);
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
    if (rows >= (int64_t)1 << 31 || cols >= (int64_t)1 << 31) return DTB_ERR_UNSUPPORTED;
    if (max_moves > 0x7fffffffLL) max_moves = 0x7fffffffLL;
    const int64_t n = (row_end - row_begin) * cols;
    const unsigned blocks = (unsigned)((n + DS_THREADS - 1) / DS_THREADS);
    cudaStream_t st = as_stream(stream);
    const double pd = px * sqrt(2.0);
    const bool small = rows * cols < ((int64_t)1 << 31);  // positions fit 32 bits
#define DTB_DS(NAME, T, IX, LIM)                                                                                                \
    DTB_KERNEL(NAME, st, (downslope_kernel<T, IX><<<blocks, DS_THREADS, 0, st>>>((const T *)dem, fdr, (int)rows, (int)cols, row_begin, \
                                                                                row_end, px, pd, LIM, (int)max_moves, out,          \
                                                                                open_above, open_below, escaped)))
    if (dem_dtype == DTB_F32) {
        const float lim = drop_limit_f32(delta);
        if (small) DTB_DS("downslope_kernel<f32>", float, int32_t, lim);
        else DTB_DS("downslope_kernel<f32>", float, int64_t, lim);
    } else if (dem_dtype == DTB_I16) {
        const int lim = drop_limit_i16(delta);
        if (small) DTB_DS("downslope_kernel<i16>", int16_t, int32_t, lim);
        else DTB_DS("downslope_kernel<i16>", int16_t, int64_t, lim);
    } else {
        return DTB_ERR_INVALID;
    }
#undef DTB_DS
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
