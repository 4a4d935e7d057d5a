/*
 * dt_oracle.c -- CPU restatement of the descriptools terrain-descriptor path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is product code: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker / reported CPU baseline.  The product
 * path (descriptools_b200/) never calls it and has no CPU fallback.
 *
 * Every function restates the semantics of the reference's *_gpu Numba kernel with
 * division_* = 0 (the unpartitioned path) and cites the reference
 * file:line it follows (paths relative to the reference checkout, descriptools/...).
 * Parity pin: tests/test_oracle_golden.py checks these functions against vectors
 * produced by the reference itself (tests/golden/make_golden.py: compiled CPU-jit
 * twins on the full bundled example, the unmodified @cuda.jit kernels under
 * NUMBA_ENABLE_CUDASIM=1 on crops) and against the bundled KAT files.
 *
 * Two stages have NO reference implementation (D8 flow direction, D8 flow
 * accumulation): the reference only consumes them (flowhand.py:801-824,
 * gfi.py:432, topoindexes.py:252-255).  For those this file IS the specification
 * ("parity unpinned by reference code"); the accumulation convention is pinned by
 * the bundled 12_fdr.tif -> 12_fac.tif pair (KAT-2).
 *
 * Typing follows compiled Numba: Python float/int scalars are f64/i64, f32 array
 * element (op) f64 scalar promotes to f64, f32 - f32 stays f32, i16 - i16 is i64.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ND (-100)

/* D8 offsets in the slope loop's scan order (slope.py:244-246: y outer, x inner):
 * NW, N, NE, W, E, SW, S, SE.  Codes per flowhand.py:801-824. */
static const int SCAN_DR[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
static const int SCAN_DC[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
static const uint8_t SCAN_CODE[8] = {32, 64, 128, 16, 1, 8, 4, 2};

/* code -> (dr, dc); returns 0 for unknown codes (flowhand.py:801-824) */
static inline int code_offset(uint8_t code, int *dr, int *dc)
{
    switch (code) {
    case 1:   *dr = 0;  *dc = 1;  return 1;
    case 2:   *dr = 1;  *dc = 1;  return 1;
    case 4:   *dr = 1;  *dc = 0;  return 1;
    case 8:   *dr = 1;  *dc = -1; return 1;
    case 16:  *dr = 0;  *dc = -1; return 1;
    case 32:  *dr = -1; *dc = -1; return 1;
    case 64:  *dr = -1; *dc = 0;  return 1;
    case 128: *dr = -1; *dc = 1;  return 1;
    default:  *dr = 0;  *dc = 0;  return 0;
    }
}

/* ------------------------------------------------------------------------- */
/* A1 slope + A2 D8.  slope.py:228-259 (kernel) with the -100 padding ring of
 * slope.py:175-182 expressed as bounds checks.  D8 is new:
 * the neighbour that last updated the running maximum (strict '<' keeps the
 * first maximum in scan order); nodata -> 0; a valid cell with no strictly
 * lower valid neighbour points at its first neighbour in scan order whose gradient
 * is undefined (off-raster, == -100, or a NaN difference), else 0.
 * Rows [row_begin,row_end) of a raster with `rows` rows are computed; outputs are
 * indexed from row_begin (used by the band tests). */
#define SLOPE_D8_BODY(T, DIFF_T)                                                         \
    const double dcard = px;                                                             \
    const double ddiag = px * sqrt(2.0);                                                 \
    _Pragma("omp parallel for schedule(static)")                                         \
    for (int64_t r = row_begin; r < row_end; ++r) {                                      \
        for (int64_t c = 0; c < cols; ++c) {                                             \
            const int64_t o = (r - row_begin) * cols + c;                                \
            const T zc = dem[r * cols + c];                                              \
            if (zc <= (T)ND) { /* slope.py:231 */                                        \
                if (slope) slope[o] = (float)ND;                                         \
                if (d8) d8[o] = 0;                                                       \
                continue;                                                                \
            }                                                                            \
            double m = 0.0;                                                              \
            uint8_t code = 0;                                                            \
            for (int k = 0; k < 8; ++k) {                                                \
                const int64_t rr = r + SCAN_DR[k], cc = c + SCAN_DC[k];                  \
                if (rr < 0 || rr >= rows || cc < 0 || cc >= cols) continue;              \
                const T zq = dem[rr * cols + cc];                                        \
                if (zq == (T)ND) continue; /* slope.py:247 */                            \
                const DIFF_T diff = (DIFF_T)zc - (DIFF_T)zq;                             \
                const double g = (double)diff / ((SCAN_DR[k] == 0 || SCAN_DC[k] == 0)    \
                                                     ? dcard : ddiag);                   \
                if (m < g) { m = g; code = SCAN_CODE[k]; } /* slope.py:250,255 */        \
            }                                                                            \
            if (slope) slope[o] = (float)(m * 100.0); /* slope.py:259 */                 \
            if (d8) {                                                                    \
                if (code == 0) {                                                         \
                    for (int k = 0; k < 8; ++k) {                                        \
                        const int64_t rr = r + SCAN_DR[k], cc = c + SCAN_DC[k];          \
                        int skipped = (rr < 0 || rr >= rows || cc < 0 || cc >= cols);    \
                        if (!skipped) {                                                  \
                            const T zq = dem[rr * cols + cc];                            \
                            const DIFF_T df = (DIFF_T)zc - (DIFF_T)zq;                   \
                            skipped = (zq == (T)ND) || (df != df);                       \
                        }                                                                \
                        if (skipped) { code = SCAN_CODE[k]; break; }                     \
                    }                                                                    \
                }                                                                        \
                d8[o] = code;                                                            \
            }                                                                            \
        }                                                                                \
    }

void orc_slope_d8_f32(const float *dem, int64_t rows, int64_t cols, int64_t row_begin,
                      int64_t row_end, double px, float *slope, uint8_t *d8)
{
    SLOPE_D8_BODY(float, float)
}

void orc_slope_d8_i16(const int16_t *dem, int64_t rows, int64_t cols, int64_t row_begin,
                      int64_t row_end, double px, float *slope, uint8_t *d8)
{
    SLOPE_D8_BODY(int16_t, int64_t)
}

/* ------------------------------------------------------------------------- */
/* A3 D8 flow accumulation (new; convention pinned by KAT-2, consumers gfi.py:432,
 * topoindexes.py:252-255): A[p] = number of cells strictly upstream of p.  Kahn
 * sweep.  A move that leaves the raster or lands on a code-0 cell is dropped.
 * Cells on a D8 cycle are never finalised and keep their partial count.
 * Cells with code 0 receive `nodata_fill`.  Returns the number of valid cells
 * that were NOT finalised (0 on a cycle-free grid). */
int64_t orc_flowacc(const uint8_t *d8, int64_t rows, int64_t cols, int64_t *acc,
                    int64_t nodata_fill)
{
    const int64_t n = rows * cols;
    uint8_t *indeg = (uint8_t *)calloc((size_t)n, 1);
    int64_t *next = (int64_t *)malloc((size_t)n * sizeof(int64_t));
    int64_t *stack = (int64_t *)malloc((size_t)n * sizeof(int64_t));
    int64_t valid = 0, done = 0, top = 0;
    for (int64_t r = 0; r < rows; ++r)
        for (int64_t c = 0; c < cols; ++c) {
            const int64_t p = r * cols + c;
            int dr, dc;
            acc[p] = 0;
            next[p] = -1;
            if (d8[p] == 0) continue;
            ++valid;
            if (!code_offset(d8[p], &dr, &dc)) continue;
            const int64_t rr = r + dr, cc = c + dc;
            if (rr < 0 || rr >= rows || cc < 0 || cc >= cols) continue;
            const int64_t q = rr * cols + cc;
            if (d8[q] == 0) continue;
            next[p] = q;
            indeg[q]++;
        }
    for (int64_t p = 0; p < n; ++p)
        if (d8[p] != 0 && indeg[p] == 0) stack[top++] = p;
    while (top > 0) {
        const int64_t p = stack[--top];
        ++done;
        const int64_t q = next[p];
        if (q >= 0) {
            acc[q] += acc[p] + 1;
            if (--indeg[q] == 0) stack[top++] = q;
        }
    }
    for (int64_t p = 0; p < n; ++p)
        if (d8[p] == 0) acc[p] = nodata_fill;
    free(indeg);
    free(next);
    free(stack);
    return valid - done;
}

/* ------------------------------------------------------------------------- */
/* A4 flow distance + river-cell index.  flowhand.py:599-846 with out[*] = 0,
 * row_start = col_start = 0, matrix_columns = col (the unpartitioned call,
 * flowhand.py:392-402).  `max_moves` is the reference's 20000 (flowhand.py:835);
 * it is a parameter only so the cap can be exercised on small rasters. */
void orc_flow_distance_index(const uint8_t *fdr, const int8_t *river, int64_t rows,
                             int64_t cols, double px, int64_t max_moves, float *fdist,
                             int64_t *idx)
{
    const int64_t n = rows * cols;
    const double ldiag = px * sqrt(2.0);
    #pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t i = 0; i < n; ++i) {
        if (fdr[i] <= 0) { /* flowhand.py:601-603 */
            fdist[i] = (float)ND;
            idx[i] = ND;
            continue;
        }
        if (river[i] == 1) { /* flowhand.py:609-612 */
            fdist[i] = 0.0f;
            idx[i] = i;
            continue;
        }
        int64_t pos = i, loop = 0, loop1 = -10, loop2 = -20, loop3 = -30;
        int isnan_ = 0;
        double dist = 0.0;
        while (river[pos] != 1) { /* flowhand.py:622 */
            const uint8_t f = fdr[pos];
            /* border tests, flowhand.py:623-628, 671-677, 715-721, 759-764 */
            if (pos < cols && (f == 32 || f == 64 || f == 128)) { isnan_ = 1; break; }
            else if (pos % cols == 0 && (f == 8 || f == 16 || f == 32)) { isnan_ = 1; break; }
            else if (pos % cols == cols - 1 && (f == 128 || f == 1 || f == 2)) { isnan_ = 1; break; }
            else if (pos >= (rows - 1) * cols && (f == 2 || f == 4 || f == 8)) { isnan_ = 1; break; }
            loop3 = loop2; loop2 = loop1; loop1 = pos; /* flowhand.py:797-799 */
            switch (f) { /* flowhand.py:801-824 */
            case 1:   pos += 1;         dist += px;    break;
            case 2:   pos += 1 + cols;  dist += ldiag; break;
            case 4:   pos += cols;      dist += px;    break;
            case 8:   pos += cols - 1;  dist += ldiag; break;
            case 16:  pos += -1;        dist += px;    break;
            case 32:  pos += -1 - cols; dist += ldiag; break;
            case 64:  pos += -cols;     dist += px;    break;
            case 128: pos += -cols + 1; dist += ldiag; break;
            default: break;
            }
            if (fdr[pos] == 0) { isnan_ = 1; break; } /* flowhand.py:826 */
            if (pos == loop1 || pos == loop2 || pos == loop3) { isnan_ = 1; break; } /* :830 */
            loop += 1;
            if (loop > max_moves) { isnan_ = 1; break; } /* flowhand.py:835 */
        }
        if (isnan_) {
            fdist[i] = (float)ND;
            idx[i] = ND;
        } else {
            fdist[i] = (float)dist; /* f64 -> f32 store, flowhand.py:540,843 */
            idx[i] = pos;           /* flowhand.py:845 */
        }
    }
}

/* A5 HAND.  flowhand.py:431-442: dem - dem[idx] in the DEM's dtype where both
 * are valid, else -100; negatives other than -100 clamp to 0. */
void orc_hand_f32(const float *dem, const int64_t *idx, int64_t n, float *hand)
{
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float h = (float)ND;
        if (dem[i] != (float)ND && idx[i] != ND) h = dem[i] - dem[idx[i]];
        if (h < 0.0f && h != (float)ND) h = 0.0f;
        hand[i] = h;
    }
}

void orc_hand_i16(const int16_t *dem, const int64_t *idx, int64_t n, int16_t *hand)
{
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        int16_t h = ND;
        if (dem[i] != ND && idx[i] != ND) h = (int16_t)(dem[i] - dem[idx[i]]); /* i16 wrap */
        if (h < 0 && h != ND) h = 0;
        hand[i] = h;
    }
}

/* ------------------------------------------------------------------------- */
/* A6 downslope index.  Composite of downslope_gpu (downslope.py:458-532) and the
 * CPU pass over its -50 flags (downslope.py:194-312, 373-374), which equals the
 * CPU-jit semantics on every cell (SURVEY.md addendum), plus the kernel's
 * `dem <= -100 -> -100` (downslope.py:459).  `max_moves` is the reference's 5000
 * (downslope.py:303).  A valid cell that never moves (code 0 / unknown) makes the
 * reference raise ZeroDivisionError; defined as 0 here. */
#define DOWNSLOPE_BODY(T, DIFF_T)                                                        \
    const double ldiag = px * sqrt(2.0);                                                 \
    _Pragma("omp parallel for schedule(dynamic, 4096)")                                  \
    for (int64_t i = 0; i < rows * cols; ++i) {                                          \
        const T z0 = dem[i];                                                             \
        if (z0 <= (T)ND) { out[i] = (float)ND; continue; }                               \
        int64_t y = i / cols, x = i % cols, loop = 0;                                    \
        int is_nan = 0; (void)is_nan;                                                               \
        double dist = 0.0;                                                               \
        while ((double)((DIFF_T)z0 - (DIFF_T)dem[y * cols + x]) < delta) {               \
            const uint8_t f = fdr[y * cols + x];                                         \
            /* downslope.py:212-231 */                                                   \
            if (y == 0 && (f == 32 || f == 64 || f == 128)) { is_nan = 1; break; }       \
            else if (y == rows - 1 && (f == 2 || f == 4 || f == 8)) { is_nan = 1; break; } \
            else if (x == 0 && (f == 32 || f == 16 || f == 8)) { is_nan = 1; break; }    \
            else if (x == cols - 1 && (f == 128 || f == 1 || f == 2)) { is_nan = 1; break; } \
            int dr, dc;                                                                  \
            if (code_offset(f, &dr, &dc)) { /* downslope.py:233-280 */                   \
                if (dem[(y + dr) * cols + (x + dc)] == (T)ND) { is_nan = 1; break; }     \
                y += dr; x += dc;                                                        \
                dist += (dr == 0 || dc == 0) ? px : ldiag;                               \
            }                                                                            \
            loop += 1;                                                                   \
            if (loop == max_moves) break; /* downslope.py:302-304 */                     \
        }                                                                                \
        if (dist == 0.0) out[i] = 0.0f; /* :306-307; also the 0/0 case */                \
        else out[i] = (float)((double)((DIFF_T)z0 - (DIFF_T)dem[y * cols + x]) / dist);  \
    }

void orc_downslope_f32(const float *dem, const uint8_t *fdr, int64_t rows, int64_t cols,
                       double px, double delta, int64_t max_moves, float *out)
{
    DOWNSLOPE_BODY(float, float)
}

void orc_downslope_i16(const int16_t *dem, const uint8_t *fdr, int64_t rows, int64_t cols,
                       double px, double delta, int64_t max_moves, float *out)
{
    DOWNSLOPE_BODY(int16_t, int64_t)
}

/* ------------------------------------------------------------------------- */
/* A7 river-cell accumulation gather.  gfi.py:136-147. */
void orc_river_accumulation(const int64_t *fac, const int64_t *idx, int64_t n, int64_t *out)
{
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) out[i] = (idx[i] != ND) ? fac[idx[i]] : fac[0];
}

/* A8 GFI.  gfi.py:287-294: log(b * pow(A_r * size^2, n) / (H + 0.01)) in f64,
 * stored f32; H <= -100 -> -100. */
#define GFI_BODY(T)                                                                      \
    const double s2 = size * size;                                                       \
    _Pragma("omp parallel for schedule(static)")                                         \
    for (int64_t i = 0; i < n; ++i) {                                                    \
        if (hand[i] <= (T)ND) { out[i] = (float)ND; continue; }                          \
        out[i] = (float)log(b * pow((double)racc[i] * s2, expo) / ((double)hand[i] + 0.01)); \
    }

void orc_gfi_f32(const float *hand, const int64_t *racc, int64_t n, double expo, double b,
                 double size, float *out)
{
    GFI_BODY(float)
}

void orc_gfi_i16(const int16_t *hand, const int64_t *racc, int64_t n, double expo, double b,
                 double size, float *out)
{
    GFI_BODY(int16_t)
}

/* A9 ln(hl/H).  gfi.py:427-440: the cell's own accumulation, A == 0 -> 1. */
#define LNHLH_BODY(T)                                                                    \
    const double s2 = size * size;                                                       \
    _Pragma("omp parallel for schedule(static)")                                         \
    for (int64_t i = 0; i < n; ++i) {                                                    \
        if (hand[i] <= (T)ND) { out[i] = (float)ND; continue; }                          \
        const double a = (fac[i] == 0) ? 1.0 : (double)fac[i];                           \
        out[i] = (float)log((b * pow(a * s2, expo)) / ((double)hand[i] + 0.01));         \
    }

void orc_lnhlh_f32(const float *hand, const int64_t *fac, int64_t n, double expo, double b,
                   double size, float *out)
{
    LNHLH_BODY(float)
}

void orc_lnhlh_i16(const int16_t *hand, const int64_t *fac, int64_t n, double expo, double b,
                   double size, float *out)
{
    LNHLH_BODY(int16_t)
}

/* A10 TI / MTI.  topoindexes.py:250-261, 284-295 (the *_gpu formula:
 * tan(beta + 0.01), nodata test on A <= -100, A == 0 -> 1). */
void orc_ti_mti(const int64_t *fac, const float *slope_rad, int64_t n, double px, double expo,
                float *ti, float *mti)
{
    const double p2 = px * px;
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        if (fac[i] <= ND) {
            if (ti) ti[i] = (float)ND;
            if (mti) mti[i] = (float)ND;
            continue;
        }
        const double a = (fac[i] == 0) ? 1.0 : (double)fac[i];
        const double t = tan((double)slope_rad[i] + 0.01);
        if (ti) ti[i] = (float)log((a * p2) / t);
        if (mti) mti[i] = (float)log(pow(a * p2, expo) / t);
    }
}

int orc_abi_version(void) { return 1; }

/* OpenMP team size of this library's loops (bench.py's CPU legs: torchrun exports OMP_NUM_THREADS=1, and the
 * process may hold more than one OpenMP runtime, so the setting has to go through the one linked here). */
#ifdef _OPENMP
#include <omp.h>
void orc_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int orc_max_threads(void) { return omp_get_max_threads(); }
#else
void orc_set_threads(int n) { (void)n; }
int orc_max_threads(void) { return 1; }
#endif
