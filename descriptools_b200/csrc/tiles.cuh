// tiles.cuh -- 64x64 tile geometry shared by the tiled D8-forest kernels (flowacc.cu, hand.cu).
//
// A raster (or row band) is cut into 64x64 tiles; the 252 perimeter cells of a tile own a "slot"
// and the pair (tile, slot) is a node id.  Cells that receive flow from outside their tile are the
// entry nodes of the second-level forest both stages resolve.
#pragma once
#include "common.cuh"

namespace dtb {

constexpr int T = 64;                 // tile edge (cells)
constexpr int TCELLS = T * T;
constexpr int SLOTS = 256;            // perimeter slots per tile (252 used)
constexpr int FT_THREADS = 256;
constexpr int CPT = TCELLS / FT_THREADS;  // cells per thread (16 consecutive columns of one row)
constexpr int CP = 80;                // shared-memory pitch of a code row: col -1 at 15, col 0 at 16
constexpr uint32_t NXT_EXIT = 0xFFFEu, NXT_TERM = 0xFFFFu;
constexpr uint32_t LINK_NONE = 0xFFFFFFFFu;   // in-tile path ends inside the tile
constexpr uint32_t LINK_OUT = 0x80000000u;    // leaves the band: | (below ? 0x40000000 : 0) | column
constexpr uint32_t LINK_BELOW = 0x40000000u;
constexpr uint32_t TERM_NONE = 255u;

__host__ __device__ __forceinline__ int slot_of(int lr, int lc)
{
    if (lr == 0) return lc;
    if (lr == T - 1) return T + lc;
    if (lc == 0) return 2 * T + lr - 1;
    return 2 * T + (T - 2) + lr - 1;  // lc == T-1
}
__host__ __device__ __forceinline__ void slot_cell(int s, int &lr, int &lc)
{
    if (s < T) { lr = 0; lc = s; }
    else if (s < 2 * T) { lr = T - 1; lc = s - T; }
    else if (s < 2 * T + (T - 2)) { lr = s - 2 * T + 1; lc = 0; }
    else { lr = s - (2 * T + (T - 2)) + 1; lc = T - 1; }
}
constexpr int USED_SLOTS = 4 * T - 4;

// neighbour positions in scan order NW,N,NE,W,E,SW,S,SE (bit k of the in-masks)
__device__ __forceinline__ void nbr_offset(int k, int &dr, int &dc)
{
    const int kk = k < 4 ? k : k + 1;  // skip the centre of the 3x3
    dr = kk / 3 - 1;
    dc = kk % 3 - 1;
}

struct TileView {
    const uint8_t *d8, *halo_above, *halo_below;
    int64_t rows, cols;
    int tiles_x;
};

__device__ __forceinline__ unsigned fetch_code(const TileView &v, int64_t gr, int64_t gc)
{
    if (gc < 0 || gc >= v.cols) return 0;
    if (gr >= 0 && gr < v.rows) return v.d8[gr * v.cols + gc];
    if (gr == -1 && v.halo_above) return v.halo_above[gc];
    if (gr == v.rows && v.halo_below) return v.halo_below[gc];
    return 0;
}


// Stage the tile's codes plus a 1-cell halo ring into shared memory.  Layout: row j (-1..T) at
// (j+1)*CP, column k (-1..T) at 16+k; `codes` must hold (T+2)*CP+16 bytes, 16-byte aligned.
// Returns true if the vector path was used (16-byte aligned full-width rows).
__device__ __forceinline__ bool stage_codes(const TileView &v, int64_t r0, int64_t c0, uint8_t *codes, int tid, int nthreads)
{
    const bool fast = (v.cols % 16 == 0) && ((reinterpret_cast<uintptr_t>(v.d8) & 15u) == 0) && (c0 + T <= v.cols);
    if (fast) {
        for (int t = tid; t < T * 4; t += nthreads) {
            const int lr = t >> 2, ch = t & 3;
            uint4 w = make_uint4(0, 0, 0, 0);
            if (r0 + lr < v.rows) w = __ldg(reinterpret_cast<const uint4 *>(v.d8 + (r0 + lr) * v.cols + c0 + ch * 16));
            *reinterpret_cast<uint4 *>(&codes[(lr + 1) * CP + 16 + ch * 16]) = w;
        }
        for (int h = tid; h < 4 * T + 4; h += nthreads) {
            int lr2, lc2;
            if (h < T + 2) { lr2 = -1; lc2 = h - 1; }
            else if (h < 2 * T + 4) { lr2 = T; lc2 = h - (T + 2) - 1; }
            else if (h < 3 * T + 4) { lr2 = h - (2 * T + 4); lc2 = -1; }
            else { lr2 = h - (3 * T + 4); lc2 = T; }
            codes[(lr2 + 1) * CP + 16 + lc2] = (uint8_t)fetch_code(v, r0 + lr2, c0 + lc2);
        }
    } else {
        for (int h = tid; h < (T + 2) * (T + 2); h += nthreads) {
            const int lr2 = h / (T + 2) - 1, lc2 = h % (T + 2) - 1;
            codes[(lr2 + 1) * CP + 16 + lc2] = (uint8_t)fetch_code(v, r0 + lr2, c0 + lc2);
        }
    }
    return fast;
}

// D8 code -> displacement of the move inside the staged tile.  `code` must be one of the eight one-hot
// direction codes (bit b: 0=E 1=SE 2=S 3=SW 4=W 5=NW 6=N 7=NE); returns false for 0 / unknown codes.
//   dloc  = dr*T  + dc  (local cell index)      dcode = dr*CP + dc  (byte offset in the staged codes)
__device__ __forceinline__ bool d8_delta(unsigned code, int &dloc, int &dcode)
{
    if (code == 0u || (code & (code - 1u)) != 0u || code > 128u) { dloc = 0; dcode = 0; return false; }
    const unsigned b = (unsigned)__ffs((int)code) - 1u;
    // signed bytes, LSB first: E, SE, S, SW | W, NW, N, NE
    constexpr unsigned L0 = (1u & 0xFF) | ((unsigned)(T + 1) << 8) | ((unsigned)T << 16) | ((unsigned)(T - 1) << 24);
    constexpr unsigned L1 = ((unsigned)(-1) & 0xFF) | (((unsigned)(-T - 1) & 0xFF) << 8) | (((unsigned)(-T) & 0xFF) << 16) |
                            (((unsigned)(-T + 1) & 0xFF) << 24);
    constexpr unsigned C0 = (1u & 0xFF) | ((unsigned)(CP + 1) << 8) | ((unsigned)CP << 16) | ((unsigned)(CP - 1) << 24);
    constexpr unsigned C1 = ((unsigned)(-1) & 0xFF) | (((unsigned)(-CP - 1) & 0xFF) << 8) | (((unsigned)(-CP) & 0xFF) << 16) |
                            (((unsigned)(-CP + 1) & 0xFF) << 24);
    dloc = (int)(int8_t)__byte_perm(L0, L1, b);
    dcode = (int)(int8_t)__byte_perm(C0, C1, b);
    return true;
}
// Shared-memory layout of per-cell arrays.  Thread t owns the 16 cells 16t..16t+15 (row t/4, columns
// 16(t%4)..+15), so a row-major array would put every lane of a warp in the same one or two banks.
// Cell p therefore lives at slot (p % 16) * 256 + p / 16: for a fixed cell number i the 256 threads touch
// consecutive words.  Pointers between cells are stored as slots.
__host__ __device__ __forceinline__ uint32_t phys_of(uint32_t p) { return ((p & 15u) << 8) | (p >> 4); }
__host__ __device__ __forceinline__ uint32_t logical_of(uint32_t q) { return ((q & 255u) << 4) | (q >> 8); }

// one-hot codes whose move leaves the tile from local cell (lr, lc)
__device__ __forceinline__ unsigned exit_codes(int lr, int lc)
{
    return (lr == 0 ? 0xE0u : 0u) | (lr == T - 1 ? 0x0Eu : 0u) | (lc == 0 ? 0x38u : 0u) | (lc == T - 1 ? 0x83u : 0u);
}

// in-mask of a cell: bit k set iff the neighbour at scan position k (NW,N,NE,W,E,SW,S,SE) points at it
__device__ __forceinline__ unsigned in_mask(const uint8_t *codes, int lr, int lc)
{
    const uint8_t *c = codes + (lr + 1) * CP + 16 + lc;
    unsigned inm = 0;
    inm |= (c[-CP - 1] == 2u) << 0;
    inm |= (c[-CP] == 4u) << 1;
    inm |= (c[-CP + 1] == 8u) << 2;
    inm |= (c[-1] == 1u) << 3;
    inm |= (c[1] == 16u) << 4;
    inm |= (c[CP - 1] == 128u) << 5;
    inm |= (c[CP] == 64u) << 6;
    inm |= (c[CP + 1] == 32u) << 7;
    return inm;
}
// neighbour positions that lie outside the tile
__device__ __forceinline__ unsigned out_mask(int lr, int lc)
{
    return (lr == 0 ? 0x07u : 0u) | (lr == T - 1 ? 0xE0u : 0u) | (lc == 0 ? 0x29u : 0u) | (lc == T - 1 ? 0x94u : 0u);
}
__device__ __forceinline__ int64_t node_of_cell(int64_t gr, int64_t gc, int tiles_x)
{
    return ((gr / T) * tiles_x + gc / T) * SLOTS + slot_of((int)(gr % T), (int)(gc % T));
}

// ---- packed path state of the HAND stage (hand.cu; also written by the fused finish pass of flowacc.cu) ----
//     [63..62 kind | 61..47 diagonal moves | 46..32 cardinal moves | 31..0 target]
constexpr uint64_t KIND_ACTIVE = 0, KIND_RIVER = 1, KIND_FAIL = 2, KIND_EXIT = 3;
constexpr uint32_t CNT_SAT = 32767;
constexpr int JUMP_BLOCKS = kNumSMs * 8;

__device__ __forceinline__ uint64_t pack(uint64_t kind, uint32_t nd, uint32_t nc, uint32_t ptr)
{
    return (kind << 62) | ((uint64_t)nd << 47) | ((uint64_t)nc << 32) | (uint64_t)ptr;
}
__device__ __forceinline__ uint64_t kind_of(uint64_t s) { return s >> 62; }
__device__ __forceinline__ uint32_t nd_of(uint64_t s) { return (uint32_t)(s >> 47) & 0x7FFFu; }
__device__ __forceinline__ uint32_t nc_of(uint64_t s) { return (uint32_t)(s >> 32) & 0x7FFFu; }
__device__ __forceinline__ uint32_t ptr_of(uint64_t s) { return (uint32_t)s; }
__device__ __forceinline__ uint32_t sat_add(uint32_t a, uint32_t b) { return min(a + b, CNT_SAT); }
// state of a path that continues with `t` after the moves recorded in `s`
__device__ __forceinline__ uint64_t compose(uint64_t s, uint64_t t)
{
    const uint64_t kt = kind_of(t);
    if (kt == KIND_FAIL) return pack(KIND_FAIL, 0, 0, 0);
    return pack(kt, sat_add(nd_of(s), nd_of(t)), sat_add(nc_of(s), nc_of(t)), ptr_of(t));
}


}  // namespace dtb
