// slope_d8.cu -- fused slope (%) + D8 flow direction, one 3x3 stencil pass.
//
// Semantics: slope.py:228-259 (slope_gpu) with the -100 padding ring of slope.py:175-182
// replaced by "off-raster neighbours are skipped"; D8 per SURVEY.md App. A2 (direction of
// the neighbour that last raised the running maximum in scan order NW,N,NE,W,E,SW,S,SE;
// outlets point at their first undefined neighbour).  Bit-exact against oracle/dt_oracle.c.
//
// B200 design: persistent CTAs (3 per SM) walk 128x32-cell tiles in row-major order; each
// tile plus its halo (a 136x34 f32 box starting 4 columns left of the tile: the TMA unit
// faults unless the box's innermost start coordinate is 16-byte aligned -- measured on B200,
// scripts/tma_probe.cu) is staged in shared memory by TMA
// (cp.async.bulk.tensor.2d, out-of-bounds elements filled with NaN, which stands in for
// the reference's -100 padding ring) through a 3-stage mbarrier ring, so the next two
// tiles are in flight while one is computed.  A thread owns 4 columns x 4 rows and slides
// a 3-row register window down its strip: LDS.128+LDS.64 per row, every elevation
// difference computed once and used by both endpoints (19 FSUB per 4 cells), results
// leave as one STG.128 (slope) and one STG.32 (four D8 codes) per row.
//
// Exactness without f64 divides in the hot loop (the reference divides in f64 inside the
// 8-neighbour loop, slope.py:249-257):
//   * within a class (cardinal / diagonal) the divisor is a common positive constant, so
//     comparing the f32 differences is equivalent to comparing the f64 gradients;
//   * cardinal-vs-diagonal is decided in f32 with a 1e-6 guard band and falls back to the
//     two f64 divisions only inside the band;
//   * slope = f32(f64(diff)/d*100) is computed as f64(diff)*(100/d) and falls back to the
//     exact expression when the product sits within 16 ulp of an f32 rounding boundary.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace dtb {
namespace {

constexpr int TW = 128;           // tile width  (cells)
constexpr int TH = 32;            // tile height (cells)
constexpr int HALO_L = 4;         // staged columns start at c0-4 (TMA: 16-byte aligned box start)
constexpr int BOXW = TW + 8;      // staged columns c0-4 .. c0+TW+3
constexpr int BOXH = TH + 2;      // staged rows    r0-1 .. r0+TH
constexpr int NTHREADS = 256;     // 32 column groups x 8 row groups
constexpr int RPT = TH / 8;       // rows per thread
constexpr int STAGES = 3;
constexpr int STAGE_BYTES = ((BOXW * BOXH * 4 + 127) / 128) * 128;
constexpr int TMA_BYTES = BOXW * BOXH * 4;
constexpr size_t SMEM_TMA = (size_t)STAGES * STAGE_BYTES + 256;

struct SlopeConsts {
    double px;   // cardinal step            (slope.py:250)
    double pd;   // px * sqrt(2.0)           (slope.py:255)
    double kc;   // 100 / px
    double kd;   // 100 / pd
    float r32;   // (float)(px / pd)
};

// ---- PTX wrappers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}

// ---- per-cell finish: class decision, exact slope, outlet rule --------------------------------
// ac/cc: first-maximum positive cardinal difference and its code; ad/cd: same for diagonals.
__device__ __forceinline__ int scan_pos(int code)
{
    // NW(32)=0 N(64)=1 NE(128)=2 W(16)=3 E(1)=4 SW(8)=5 S(4)=6 SE(2)=7
    switch (code) {
    case 32: return 0; case 64: return 1; case 128: return 2; case 16: return 3;
    case 1: return 4; case 8: return 5; case 4: return 6; default: return 7;
    }
}

__device__ __noinline__ bool cardinal_wins_exact(float ac, int cc, float ad, int cd, double px, double pd)
{
    const double gc = (double)ac / px, gd = (double)ad / pd;
    if (gc > gd) return true;
    if (gc < gd) return false;
    return scan_pos(cc) < scan_pos(cd);
}

__device__ __noinline__ float slope_exact(float a, double d) { return (float)(((double)a / d) * 100.0); }

__device__ __forceinline__ void finish_cell(float ac, int cc, float ad, int cd, const SlopeConsts &k, float &slope,
                                            int &code)
{
    bool use_c;
    if (!(ad > 0.0f)) use_c = true;
    else if (!(ac > 0.0f)) use_c = false;
    else {
        const float t = ad * k.r32;
        if (!(t > 1e-30f)) use_c = cardinal_wins_exact(ac, cc, ad, cd, k.px, k.pd);  // subnormal products
        else if (ac > t * 1.000001f) use_c = true;
        else if (ac < t * 0.999999f) use_c = false;
        else use_c = cardinal_wins_exact(ac, cc, ad, cd, k.px, k.pd);
    }
    const float a = use_c ? ac : ad;
    code = use_c ? cc : cd;
    if (a > 0.0f) {
        const double y = (double)a * (use_c ? k.kc : k.kd);
        const uint32_t frac = (uint32_t)__double2loint(y) & 0x1FFFFFFFu;  // mantissa bits f32 drops
        const int dist = (int)frac - 0x10000000;
        if ((dist < 17 && dist > -17) || !(y > 1e-30 && y < 1e38)) slope = slope_exact(a, use_c ? k.px : k.pd);
        else slope = (float)y;
    } else {
        slope = 0.0f;
    }
}

// load one staged row (6 values around the thread's 4 cells), NaN-ify -100, report centre nodata
__device__ __forceinline__ void load_row(const float *p, float (&w)[6], unsigned &ndmask)
{
    // p -> staged column of the thread's first cell (16-byte aligned); halos at p[-1], p[4]
    const float4 a = *reinterpret_cast<const float4 *>(p);
    float raw[6] = {p[-1], a.x, a.y, a.z, a.w, p[4]};
    ndmask = 0;
#pragma unroll
    for (int j = 1; j <= 4; ++j) ndmask |= (raw[j] <= ND_F) ? (1u << (j - 1)) : 0u;  // slope.py:231
    const float qnan = __int_as_float(0x7fc00000);
#pragma unroll
    for (int j = 0; j < 6; ++j) w[j] = (raw[j] == ND_F) ? qnan : raw[j];  // slope.py:247
}

// The stencil over one thread strip: 4 columns x nrows rows of the staged tile.
//   tile: smem, row pitch BOXW; smem (row j, col k) <-> raster (r0-1+j, c0-HALO_L+k)
//   gx: column group (cells c0+4gx..+3), ry0: first tile row of the strip
template <bool VEC>
__device__ __forceinline__ void stencil_strip(const float *tile, int gx, int ry0, const SlopeConsts &k,
                                              int64_t out_row0 /* output row of tile row 0 */, int64_t out_rows,
                                              int64_t c_first /* raster col of the first cell */, int64_t cols,
                                              float *__restrict__ slope, uint8_t *__restrict__ d8)
{
    const float *base = tile + 4 * gx + HALO_L;
    float up[6], mid[6], dn[6];
    unsigned nd_mid, nd_dn, nd_unused;
    load_row(base + (ry0)*BOXW, up, nd_unused);
    load_row(base + (ry0 + 1) * BOXW, mid, nd_mid);
    float vSp[6], sEp[6], sWp[6];  // differences (upper row) - (lower row) of the previous row pair
#pragma unroll
    for (int j = 1; j <= 4; ++j) vSp[j] = up[j] - mid[j];
#pragma unroll
    for (int j = 0; j <= 3; ++j) sEp[j] = up[j] - mid[j + 1];
#pragma unroll
    for (int j = 2; j <= 5; ++j) sWp[j] = up[j] - mid[j - 1];

#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        load_row(base + (ry0 + 2 + i) * BOXW, dn, nd_dn);
        float hE[5], vS[6], sE[5], sW[6];
#pragma unroll
        for (int j = 0; j <= 4; ++j) hE[j] = mid[j] - mid[j + 1];
#pragma unroll
        for (int j = 1; j <= 4; ++j) vS[j] = mid[j] - dn[j];
#pragma unroll
        for (int j = 0; j <= 4; ++j) sE[j] = mid[j] - dn[j + 1];
#pragma unroll
        for (int j = 1; j <= 5; ++j) sW[j] = mid[j] - dn[j - 1];

        float s_out[4];
        uint32_t codes = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = c + 1;
            const float dN = -vSp[j], dW = -hE[j - 1], dE = hE[j], dS = vS[j];
            const float dNW = -sEp[j - 1], dNE = -sWp[j + 1], dSW = sW[j], dSE = sE[j];
            float ac = 0.0f, ad = 0.0f;
            int cc = 0, cd = 0;
            if (dN > ac) { ac = dN; cc = 64; }
            if (dW > ac) { ac = dW; cc = 16; }
            if (dE > ac) { ac = dE; cc = 1; }
            if (dS > ac) { ac = dS; cc = 4; }
            if (dNW > ad) { ad = dNW; cd = 32; }
            if (dNE > ad) { ad = dNE; cd = 128; }
            if (dSW > ad) { ad = dSW; cd = 8; }
            if (dSE > ad) { ad = dSE; cd = 2; }
            float s;
            int code;
            finish_cell(ac, cc, ad, cd, k, s, code);
            if (code == 0) {
                // outlet / flat: first neighbour in scan order whose difference is undefined
                // (off-raster, nodata or NaN) -- SURVEY.md App. A2
                if (dNW != dNW) code = 32;
                else if (dN != dN) code = 64;
                else if (dNE != dNE) code = 128;
                else if (dW != dW) code = 16;
                else if (dE != dE) code = 1;
                else if (dSW != dSW) code = 8;
                else if (dS != dS) code = 4;
                else if (dSE != dSE) code = 2;
            }
            if (nd_mid & (1u << c)) { s = ND_F; code = 0; }
            s_out[c] = s;
            codes |= (uint32_t)code << (8 * c);
        }
        const int64_t orow = out_row0 + ry0 + i;
        if (orow >= 0 && orow < out_rows) {
            const int64_t o = orow * cols + c_first;
            if (VEC) {
                if (c_first < cols) {
                    if (slope) *reinterpret_cast<float4 *>(slope + o) = make_float4(s_out[0], s_out[1], s_out[2], s_out[3]);
                    if (d8) *reinterpret_cast<uint32_t *>(d8 + o) = codes;
                }
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c_first + c < cols) {
                        if (slope) slope[o + c] = s_out[c];
                        if (d8) d8[o + c] = (uint8_t)(codes >> (8 * c));
                    }
            }
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) { up[j] = mid[j]; mid[j] = dn[j]; }
#pragma unroll
        for (int j = 1; j <= 4; ++j) vSp[j] = vS[j];
#pragma unroll
        for (int j = 0; j <= 3; ++j) sEp[j] = sE[j];
#pragma unroll
        for (int j = 2; j <= 5; ++j) sWp[j] = sW[j];
        nd_mid = nd_dn;
    }
}

// ---- fast strip for the TMA kernel ------------------------------------------------------------
// Same arithmetic as stencil_strip, organised for the B200 issue ports: the compare/select chains
// (ALU port, half rate) are replaced by 3-input max reductions, a sign-bit mask of (difference - class
// maximum) built with funnel shifts, and one priority encode.  Everything that is rare is delegated, one
// row of four cells at a time, to slow_row(): the literal per-neighbour loop of slope.py:244-258 over the
// staged raw values.  Rare = the 3x6 window holds an undefined value (off-raster NaN or anything <= -100:
// nodata centres slope.py:231, skipped neighbours slope.py:247), no positive gradient (outlet rule),
// cardinal/diagonal near-tie, a product next to an f32 rounding boundary, sub-/super-normal range.
__device__ __noinline__ void slow_row(const float *tile, int trow, int tcol0, double px, double pd, float4 &slope4, uint32_t &codes)
{
    // trow/tcol0: position of the first of the four cells inside the staged box
    const int DR[8] = {-1, -1, -1, 0, 0, 1, 1, 1}, DC[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
    const int CODE[8] = {32, 64, 128, 16, 1, 8, 4, 2};
    float out[4];
    codes = 0;
    for (int c = 0; c < 4; ++c) {
        const int tcol = tcol0 + c;
        const float zc = tile[trow * BOXW + tcol];
        if (zc <= ND_F) { out[c] = ND_F; continue; }  // slope.py:231 (a NaN centre is off-raster: never stored)
        double m = 0.0;
        int code = 0, first_undef = 0;
        for (int k = 0; k < 8; ++k) {
            const float zq = tile[(trow + DR[k]) * BOXW + tcol + DC[k]];
            const float diff = zc - zq;
            if (zq == ND_F || diff != diff) {  // skipped neighbour (slope.py:247) / off-raster
                if (!first_undef) first_undef = CODE[k];
                continue;
            }
            const double g = (double)diff / ((DR[k] == 0 || DC[k] == 0) ? px : pd);
            if (m < g) { m = g; code = CODE[k]; }  // slope.py:250,255
        }
        out[c] = (float)(m * 100.0);  // slope.py:259
        if (code == 0) code = first_undef;  // outlet rule, SURVEY.md App. A2
        codes |= (uint32_t)code << (8 * c);
    }
    slope4 = make_float4(out[0], out[1], out[2], out[3]);
}

// one staged row: 6 values around the thread's 4 cells; `bad` = some value is NaN or <= -100
__device__ __forceinline__ void load_row_fast(const float *p, float (&w)[6], bool &bad)
{
    const float4 a = *reinterpret_cast<const float4 *>(p);
    w[0] = p[-1]; w[1] = a.x; w[2] = a.y; w[3] = a.z; w[4] = a.w; w[5] = p[4];
    bad = false;
#pragma unroll
    for (int j = 0; j < 6; ++j) bad |= !(w[j] > ND_F);
}

__device__ __forceinline__ void stencil_strip_fast(const float *tile, int gx, int ry0, const SlopeConsts &k, int64_t out_row0,
                                                   int64_t out_rows, int64_t c_first, int64_t cols, float *__restrict__ slope,
                                                   uint8_t *__restrict__ d8)
{
    const float *base = tile + 4 * gx + HALO_L;
    float up[6], mid[6], dn[6];
    bool bad_up, bad_mid, bad_dn;
    load_row_fast(base + (ry0)*BOXW, up, bad_up);
    load_row_fast(base + (ry0 + 1) * BOXW, mid, bad_mid);
    float vSp[6], sEp[6], sWp[6];
#pragma unroll
    for (int j = 1; j <= 4; ++j) vSp[j] = up[j] - mid[j];
#pragma unroll
    for (int j = 0; j <= 3; ++j) sEp[j] = up[j] - mid[j + 1];
#pragma unroll
    for (int j = 2; j <= 5; ++j) sWp[j] = up[j] - mid[j - 1];

#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        load_row_fast(base + (ry0 + 2 + i) * BOXW, dn, bad_dn);
        float hE[5], vS[6], sE[5], sW[6];
#pragma unroll
        for (int j = 0; j <= 4; ++j) hE[j] = mid[j] - mid[j + 1];
#pragma unroll
        for (int j = 1; j <= 4; ++j) vS[j] = mid[j] - dn[j];
#pragma unroll
        for (int j = 0; j <= 4; ++j) sE[j] = mid[j] - dn[j + 1];
#pragma unroll
        for (int j = 1; j <= 5; ++j) sW[j] = mid[j] - dn[j - 1];

        float s_out[4];
        uint32_t bsel = 0;
        bool slow = bad_up | bad_mid | bad_dn;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = c + 1;
            const float dN = -vSp[j], dW = -hE[j - 1], dE = hE[j], dS = vS[j];
            const float dNW = -sEp[j - 1], dNE = -sWp[j + 1], dSW = sW[j], dSE = sE[j];
            const float ac = fmaxf(fmaxf(fmaxf(fmaxf(dN, dW), dE), dS), 0.0f);
            const float ad = fmaxf(fmaxf(fmaxf(fmaxf(dNW, dNE), dSW), dSE), 0.0f);
            const float t = ad * k.r32;
            const bool use_c = ac >= t;
            const float a = use_c ? ac : ad;
            // neighbours below their class maximum: sign bit of (d - max); scan order NW,N,NE,W,E,SW,S,SE from bit 7 down
            unsigned lose = 0;
            lose = __funnelshift_l(__float_as_uint(dNW - ad), lose, 1);
            lose = __funnelshift_l(__float_as_uint(dN - ac), lose, 1);
            lose = __funnelshift_l(__float_as_uint(dNE - ad), lose, 1);
            lose = __funnelshift_l(__float_as_uint(dW - ac), lose, 1);
            lose = __funnelshift_l(__float_as_uint(dE - ac), lose, 1);
            lose = __funnelshift_l(__float_as_uint(dSW - ad), lose, 1);
            lose = __funnelshift_l(__float_as_uint(dS - ac), lose, 1);
            lose = __funnelshift_l(__float_as_uint(dSE - ad), lose, 1);
            const unsigned win = ~lose & (use_c ? 0x5Au : 0xA5u);
            // highest set bit = first maximum in scan order (strict '<', slope.py:250,255); one nibble per cell, turned
            // into the four code bytes by a single byte permute below (no winner: the row takes the slow path)
            unsigned hb;
            asm("bfind.u32 %0, %1;" : "=r"(hb) : "r"(win));
            bsel += hb << (4 * c);
            const double y = (double)a * (use_c ? k.kc : k.kd);
            s_out[c] = (float)y;
            // range (no positive gradient, subnormal / huge), cardinal-diagonal near-tie, product within 16 ulp(f64)
            // of an f32 rounding boundary
            slow |= (__float_as_uint(a) - 0x0D800000u >= 0x7E000000u - 0x0D800000u) | (fabsf(ac - t) <= 1e-6f * t) |
                    ((((uint32_t)__double2loint(y) + 0x10u - 0x10000000u) & 0x1FFFFFE0u) == 0u);
        }
        float4 s4 = make_float4(s_out[0], s_out[1], s_out[2], s_out[3]);
        uint32_t codes = __byte_perm(0x01080402u, 0x20408010u, bsel);  // SE,S,SW,E | W,NE,N,NW
        if (slow) slow_row(tile, ry0 + 1 + i, 4 * gx + HALO_L, k.px, k.pd, s4, codes);
        const int64_t orow = out_row0 + ry0 + i;
        if (orow >= 0 && orow < out_rows && c_first < cols) {
            const int64_t o = orow * cols + c_first;
            if (slope) *reinterpret_cast<float4 *>(slope + o) = s4;
            if (d8) *reinterpret_cast<uint32_t *>(d8 + o) = codes;
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) { up[j] = mid[j]; mid[j] = dn[j]; }
#pragma unroll
        for (int j = 1; j <= 4; ++j) vSp[j] = vS[j];
#pragma unroll
        for (int j = 0; j <= 3; ++j) sEp[j] = sE[j];
#pragma unroll
        for (int j = 2; j <= 5; ++j) sWp[j] = sW[j];
        bad_up = bad_mid;
        bad_mid = bad_dn;
    }
}

// ---- TMA persistent kernel (f32, cols % 4 == 0, 16-byte aligned bases) ----------------------
__global__ void __launch_bounds__(NTHREADS, 3)
slope_d8_tma_kernel(const __grid_constant__ CUtensorMap dem_map, int64_t row_begin, int64_t row_end, int64_t cols,
                    int tiles_x, int ntiles, SlopeConsts k, float *__restrict__ slope, uint8_t *__restrict__ d8)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    if ((smem_u32(smem) & 127u) != 0u) __trap();  // TMA destination alignment
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)STAGES * STAGE_BYTES);
    const int tid = threadIdx.x;

    auto issue = [&](int tile, int stage) {
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        mbar_expect_tx(&full[stage], TMA_BYTES);
        tma_load_2d(smem + (size_t)stage * STAGE_BYTES, &dem_map, tx * TW - HALO_L, (int)(row_begin + (int64_t)ty * TH - 1),
                    &full[stage]);
    };

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int s = 0; s < STAGES; ++s) {
            const int tile = blockIdx.x + s * gridDim.x;
            if (tile < ntiles) issue(tile, s);
        }
    }
    __syncthreads();

    const int gx = tid & 31, gy = tid >> 5;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int stage = it % STAGES;
        const uint32_t parity = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(&full[stage], parity);
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const float *tbuf = reinterpret_cast<const float *>(smem + (size_t)stage * STAGE_BYTES);
        stencil_strip_fast(tbuf, gx, gy * RPT, k, (int64_t)ty * TH, row_end - row_begin, (int64_t)tx * TW + 4 * gx, cols,
                           slope, d8);
        __syncthreads();  // every thread is done reading this stage
        if (tid == 0) {
            const int next = tile + STAGES * gridDim.x;
            if (next < ntiles) issue(next, stage);
        }
    }
}

// ---- generic kernel (any width / alignment, f32 or i16 DEM) --------------------------------
template <typename T>
__global__ void __launch_bounds__(NTHREADS)
slope_d8_generic_kernel(const T *__restrict__ dem, int64_t buf_rows, int64_t row_begin, int64_t row_end, int64_t cols,
                        int tiles_x, SlopeConsts k, float *__restrict__ slope, uint8_t *__restrict__ d8)
{
    __shared__ __align__(16) float tile[BOXH * BOXW];
    const int tid = threadIdx.x;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int64_t r0 = row_begin + (int64_t)ty * TH - 1, c0 = (int64_t)tx * TW - HALO_L;
    const float qnan = __int_as_float(0x7fc00000);
    for (int idx = tid; idx < BOXH * BOXW; idx += NTHREADS) {
        const int j = idx / BOXW, kk = idx - j * BOXW;
        const int64_t r = r0 + j, c = c0 + kk;
        float v = qnan;
        if (r >= 0 && r < buf_rows && c >= 0 && c < cols) v = (float)dem[r * cols + c];
        tile[idx] = v;
    }
    __syncthreads();
    const int gx = tid & 31, gy = tid >> 5;
    stencil_strip<false>(tile, gx, gy * RPT, k, (int64_t)ty * TH, row_end - row_begin, (int64_t)tx * TW + 4 * gx, cols,
                         slope, d8);
}

// ---- tensor map (driver entry point fetched at run time: libdtb200 does not link libcuda) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

}  // namespace
}  // namespace dtb

extern "C" int dtb_slope_d8(const void *dem, int dem_dtype, int64_t buf_rows, int64_t cols, int64_t row_begin,
                            int64_t row_end, double px, float *slope, uint8_t *d8, void *stream)
{
    using namespace dtb;
    if (!dem || buf_rows <= 0 || cols <= 0 || row_begin < 0 || row_end > buf_rows || row_begin > row_end || !(px > 0.0))
        return DTB_ERR_INVALID;
    if (dem_dtype != DTB_F32 && dem_dtype != DTB_I16) return DTB_ERR_INVALID;
    if (row_begin == row_end || (!slope && !d8)) return DTB_OK;
    if (cols > (int64_t)1 << 30 || buf_rows > (int64_t)1 << 30) return DTB_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);

    SlopeConsts k;
    k.px = px;
    k.pd = px * sqrt(2.0);
    k.kc = 100.0 / k.px;
    k.kd = 100.0 / k.pd;
    k.r32 = (float)(k.px / k.pd);

    const int64_t nrows = row_end - row_begin;
    const int tiles_x = (int)((cols + TW - 1) / TW);
    const int64_t tiles_y = (nrows + TH - 1) / TH;
    const int64_t ntiles64 = tiles_y * tiles_x;
    if (ntiles64 > 0x7fffffff) return DTB_ERR_UNSUPPORTED;
    const int ntiles = (int)ntiles64;

    static const bool tma_disabled = getenv("DTB_DISABLE_TMA") != nullptr;  // debugging aid only
    const bool aligned = !tma_disabled && dem_dtype == DTB_F32 && (cols % 4 == 0) && (((uintptr_t)dem) % 16 == 0) &&
                         (!slope || ((uintptr_t)slope) % 16 == 0) && (!d8 || ((uintptr_t)d8) % 4 == 0);
    if (aligned) {
        EncodeTiledFn enc = get_encode_fn();
        if (!enc) return cuda_fail_msg("cuTensorMapEncodeTiled entry point not available");
        CUtensorMap map;
        const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)buf_rows};
        const cuuint64_t gstride[1] = {(cuuint64_t)cols * 4};
        const cuuint32_t box[2] = {BOXW, BOXH};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(dem), gdim, gstride, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA);
        if (r != CUDA_SUCCESS) return cuda_fail_msg("cuTensorMapEncodeTiled failed");
        static bool attr_set = false;
        if (!attr_set) {
            DTB_CUDA(cudaFuncSetAttribute(slope_d8_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TMA));
            attr_set = true;
        }
        const int grid = ntiles < 3 * kNumSMs ? ntiles : 3 * kNumSMs;
        DTB_KERNEL("slope_d8_tma_kernel", st, slope_d8_tma_kernel<<<grid, NTHREADS, SMEM_TMA, st>>>(map, row_begin, row_end, cols, tiles_x, ntiles, k, slope, d8));
    } else if (dem_dtype == DTB_F32) {
        DTB_KERNEL("slope_d8_generic_kernel<f32>", st, slope_d8_generic_kernel<float><<<ntiles, NTHREADS, 0, st>>>((const float *)dem, buf_rows, row_begin, row_end, cols,
                                                                    tiles_x, k, slope, d8));
    } else {
        DTB_KERNEL("slope_d8_generic_kernel<i16>", st, slope_d8_generic_kernel<int16_t><<<ntiles, NTHREADS, 0, st>>>((const int16_t *)dem, buf_rows, row_begin, row_end,
                                                                      cols, tiles_x, k, slope, d8));
    }
    return DTB_OK;
}
