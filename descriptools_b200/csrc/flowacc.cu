// flowacc.cu -- D8 flow accumulation: acc[p] = number of cells strictly upstream of p.
//
// New stage (the reference only consumes accumulation rasters: gfi.py:432,
// topoindexes.py:252-255, example.py:52); semantics per SURVEY.md App. A3, restated in
// oracle/dt_oracle.c:orc_flowacc and pinned by the bundled 12_fdr.tif -> 12_fac.tif pair.
//
// B200 design: a two-level sweep over the D8 forest (DESIGN.md "flow accumulation").
//   T1  one CTA per 64x64 tile, the tile's codes (+1-cell halo) and one packed word per cell in
//       shared memory.  Every source cell (no in-tile tributary) walks downstream; a step is one
//       shared-memory atomicAdd that adds (count+1) and decrements the `pending` field of the next
//       cell, and only the LAST tributary to arrive continues -- every cell is visited once, no
//       thread waits, no barrier per level.  The tile emits a summary for its 252 perimeter
//       cells only: weight leaving the tile (exitw), where the in-tile path of every cell that
//       receives flow from outside ends (term/link), how many such paths end in each exit (nterm).
//   N   the perimeter cells that receive outside flow ("entry nodes", ~3 % of the cells) form a
//       forest of their own: node q -> the entry node its in-tile path drains into.  fa_node_init
//       gathers each node's base inflow and in-degree from the neighbouring tiles' summaries,
//       fa_node_sweep runs the same last-arriver walk on that forest with 64-bit global atomics.
//   T2  the tile kernel again, now seeded with the resolved inflow of its entry nodes, writes acc.
// d8 is read twice, acc written once; the node arrays are ~1 byte per cell.
// Row bands (multi-GPU): the band's neighbours are represented by one halo row of codes above and
// below plus the inflow carried by those rows (dtb_flowacc_args); the band-level summary
// (exit_* / term_*) is the same construction one level up and is solved by the band driver.
//
// D8 cycles (possible only in caller-supplied grids, never in dtb_slope_d8 output) are detected
// on the device (a tile or node that cannot be finalised) and handled by the single-level sweep
// fa_flat_* over the whole raster, which reproduces the oracle's Kahn-order partial counts.
#include <type_traits>

#include "tiles.cuh"

namespace dtb {
namespace {

// node state: [63 source | 62 active | 61..48 pending | 47..0 inflow count]
constexpr uint64_t N_SRC = 1ull << 63, N_ACTIVE = 1ull << 62, N_PEND_ONE = 1ull << 48;
constexpr uint64_t N_CNT = N_PEND_ONE - 1ull, N_PEND = 0x3FFFull;
constexpr int N_PEND_SHIFT = 48;

// ---- T1: tile-local accumulation + perimeter summary ---------------------------------------------
// Shared word per cell: [31..28 pending | 27..14 local count | 13..0 successor slot] -- one atomicAdd per
// step both accumulates and returns the successor of the cell just finalised.
constexpr uint32_t W_EXIT = 0x3FFEu, W_TERM = 0x3FFFu, W_NXT = 0x3FFFu, W_PEND_ONE = 1u << 28;
constexpr int W_CNT_SHIFT = 14;
constexpr uint32_t W_CNT_ONE = 1u << W_CNT_SHIFT, W_CNT_MASK = 0x3FFFu << W_CNT_SHIFT;
constexpr int WALK_CAP = 4;
// successor table, 16 bits per cell in cell order: [15 river cell | 14 diagonal move | 13..0 successor slot / W_EXIT / W_TERM]
constexpr uint32_t NX_DIAG = 0x4000u, NX_RIVER = 0x8000u, NX_CYCLE = 0xFFFFFFFFu, NX_NODATA = W_TERM | NX_DIAG;

template <typename ACC>
__global__ void __launch_bounds__(FT_THREADS, 8)
fa_tile_kernel(TileView v, uint32_t *__restrict__ exitw, uint32_t *__restrict__ link, uint32_t *__restrict__ meta,
               ACC *__restrict__ acc, ACC nodata_fill, unsigned long long *__restrict__ counters,
               uint16_t *__restrict__ table)
{
    __shared__ __align__(16) uint8_t codes[(T + 2) * CP + 16];  // col T of the last row sits at (T+2)*CP
    __shared__ uint32_t word[TCELLS];
    __shared__ uint32_t nterm_s[SLOTS];
    __shared__ uint32_t queue[TCELLS / 4];  // a parked walk has finalised >= WALK_CAP cells: at most TCELLS / WALK_CAP of them
    __shared__ uint32_t qn;

    const int tid = threadIdx.x;
    const int tile = blockIdx.x;
    const int ty = tile / v.tiles_x, tx = tile - ty * v.tiles_x;
    const int64_t r0 = (int64_t)ty * T, c0 = (int64_t)tx * T;
    const bool fast = stage_codes(v, r0, c0, codes, tid, FT_THREADS);
    if (tid < SLOTS) nterm_s[tid] = 0;
    if (tid == 0) qn = 0;
    __syncthreads();

    // ---- per-cell set-up: successor slot; in-tile in-degree by scatter (one shared RED per cell) ----
    const int lr = tid >> 2, lcb = (tid & 3) * CPT;
    unsigned validmask = 0;
    uint32_t tab[CPT / 2];
    const unsigned rowexit = (lr == 0 ? 0xE0u : 0u) | (lr == T - 1 ? 0x0Eu : 0u);
    {
        const uint8_t *crow = codes + (lr + 1) * CP + 16 + lcb;
        const uint4 cw = *reinterpret_cast<const uint4 *>(crow);
        const uint32_t cws[4] = {cw.x, cw.y, cw.z, cw.w};
        // displacement of the move of one-hot code bit b (0=E 1=SE 2=S 3=SW 4=W 5=NW 6=N 7=NE), as unsigned bytes:
        // cell index + 65 and staged-code offset + 81 (the NW entries are 0: they also fill the upper result bytes)
        constexpr uint32_t LB0 = 66u | (130u << 8) | (129u << 16) | (128u << 24), LB1 = 64u | (0u << 8) | (1u << 16) | (2u << 24);
        constexpr uint32_t CB0 = (uint32_t)(81 + 1) | ((uint32_t)(81 + CP + 1) << 8) | ((uint32_t)(81 + CP) << 16) | ((uint32_t)(81 + CP - 1) << 24);
        constexpr uint32_t CB1 = (uint32_t)(81 - 1) | ((uint32_t)(81 - CP - 1) << 8) | ((uint32_t)(81 - CP) << 16) | ((uint32_t)(81 - CP + 1) << 24);
        static_assert(CP == 80 && T == 64, "biased displacement tables");
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // cells with a code: one bit per non-zero byte
            const uint32_t nz = (((cws[k] & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | cws[k]) & 0x80808080u;
            validmask |= ((((nz >> 7) * 0x00204081u) >> 21) & 0xFu) << (4 * k);
        }
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
            const int p = lr * T + lcb + i;
            const uint32_t code = (cws[i >> 2] >> (8 * (i & 3))) & 0xFFu;
            // one-hot codes whose move leaves the tile from this cell: the row part is per thread, the column part
            // only concerns the first / last cell of the row
            const unsigned xm = rowexit | ((i == 0 && lcb == 0) ? 0x38u : 0u) | ((i == CPT - 1 && lcb == T - CPT) ? 0x83u : 0u);
            uint32_t hb;  // index of the highest set bit (all ones for 0)
            asm("bfind.u32 %0, %1;" : "=r"(hb) : "r"(code));
            const uint32_t sel = (hb & 7u) | 0x5550u;
            const uint32_t oc = __byte_perm(CB0, CB1, sel), ol = __byte_perm(LB0, LB1, sel);
            // a move: one-hot code whose landing cell has a code (the staged halo ring makes the read safe for any b)
            const bool moves = __popc(code) == 1 && crow[i + (int)oc - 81] != 0;
            const uint32_t mv = ((code & xm) ? W_EXIT : phys_of((uint32_t)(p - 65) + ol)) | ((code & 0xAAu) ? NX_DIAG : 0u);
            // successor table entry (T2 and HAND's tile pass reuse it): successor + diagonal flag of the move; a cell
            // without a direction code is marked by the flag on a terminal entry (NX_NODATA)
            const uint32_t t16 = moves ? mv : (code ? W_TERM : NX_NODATA);
            word[i * FT_THREADS + tid] = t16 & W_NXT;
            if (i & 1) tab[i >> 1] |= t16 << 16; else tab[i >> 1] = t16;
        }
    }
    {   // cell order = thread order: thread t owns cells 16t..16t+15
        uint4 *dst = reinterpret_cast<uint4 *>(table + (size_t)tile * TCELLS + tid * CPT);
        dst[0] = make_uint4(tab[0], tab[1], tab[2], tab[3]);
        dst[1] = make_uint4(tab[4], tab[5], tab[6], tab[7]);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
        const uint32_t n = word[i * FT_THREADS + tid] & W_NXT;  // (other threads only touch the pending field)
        if (n < W_EXIT) atomicAdd(&word[n], W_PEND_ONE);
    }
    __syncthreads();
    // Sources are spatially clustered (ridges), so the walks are dealt out across the CTA instead of by owner: for its
    // k-th walk slot thread t takes cell k of thread (t + 37 k) mod 256 (still consecutive words for consecutive lanes).
    // A cell without a direction code looks like a source with no successor: its walk is empty.
    unsigned srcmask = 0;
#pragma unroll
    for (int i = 0; i < CPT; ++i) srcmask |= ((word[i * FT_THREADS + ((tid + 37 * i) & (FT_THREADS - 1))] >> 28) == 0u ? 1u : 0u) << i;
    __syncthreads();  // every thread has read its pending fields before the sweep starts changing them

    // ---- sweep: every source walks until it is not the last tributary to arrive.  A walk that is still
    // going after WALK_CAP steps (a main stem) is parked in a queue; the parked walks are then spread over
    // the lanes of as few warps as possible, so the long serial chains do not idle 31 lanes each ----
    while (srcmask) {
        const int i = __ffs((int)srcmask) - 1;
        srcmask &= srcmask - 1;
        // `carry` = cells finalised so far, kept in the position of the count field (<< 14)
        uint32_t n = word[i * FT_THREADS + ((tid + 37 * i) & (FT_THREADS - 1))] & W_NXT, carry = 0;
        int left = WALK_CAP;
        while (n < W_EXIT) {
            if (left-- == 0) {
                queue[atomicAdd(&qn, 1u)] = carry | n;  // count < 4096: carry < 2^26
                break;
            }
            const uint32_t old = atomicAdd(&word[n], carry + (W_CNT_ONE - W_PEND_ONE));
            if (old >= 2u * W_PEND_ONE) break;  // other tributaries still pending (the field was >= 1: mine)
            carry += (old & W_CNT_MASK) + W_CNT_ONE;
            n = old & W_NXT;
        }
    }
    __syncthreads();
    for (uint32_t k = tid; k < qn; k += FT_THREADS) {
        uint32_t n = queue[k] & W_NXT, carry = queue[k] & ~W_NXT;
        while (n < W_EXIT) {
            const uint32_t old = atomicAdd(&word[n], carry + (W_CNT_ONE - W_PEND_ONE));
            if (old >= 2u * W_PEND_ONE) break;
            carry += (old & W_CNT_MASK) + W_CNT_ONE;
            n = old & W_NXT;
        }
    }
    __syncthreads();
    // a cell that was never finalised (pending left) means the grid has a D8 cycle
    uint32_t my[CPT];
    {
        uint32_t pend = 0;
        unsigned bad = 0;
#pragma unroll
        for (int i = 0; i < CPT; ++i) { my[i] = word[i * FT_THREADS + tid]; pend |= my[i]; }
        if (pend >> 28) {  // (a cell without a code has no tributary in the tile: its field is 0)
#pragma unroll
            for (int i = 0; i < CPT; ++i) bad += (my[i] >> 28) != 0u;
        }
        bad = __reduce_add_sync(0xffffffffu, bad);
        if ((tid & 31) == 0 && bad) atomicAdd(&counters[0], (unsigned long long)bad);
    }

    // ---- tile-local counts -> acc (the finish pass adds the inflow along the entry paths) ----
    const int64_t gr = r0 + lr;
    if (acc && gr < v.rows) {
        alignas(16) ACC out[CPT];
#pragma unroll
        for (int i = 0; i < CPT; ++i)
            out[i] = ((validmask >> i) & 1u) ? (ACC)((my[i] >> W_CNT_SHIFT) & 0x3FFFu) : nodata_fill;
        ACC *dst = acc + gr * v.cols + c0 + lcb;
        if (fast && ((reinterpret_cast<uintptr_t>(acc) & 15u) == 0)) {
            constexpr int V = 16 / sizeof(ACC);
#pragma unroll
            for (int i = 0; i < CPT; i += V) *reinterpret_cast<uint4 *>(dst + i) = *reinterpret_cast<const uint4 *>(&out[i]);
        } else {
#pragma unroll
            for (int i = 0; i < CPT; ++i)
                if (c0 + lcb + i < v.cols) dst[i] = out[i];
        }
    }

    // ---- summary of the perimeter ----
    uint32_t my_link = LINK_NONE, my_term = TERM_NONE, my_exitw = 0, my_inmask = 0;
    if (tid < USED_SLOTS) {
        int plr, plc;
        slot_cell(tid, plr, plc);
        uint32_t q = phys_of((uint32_t)(plr * T + plc));
        uint32_t w = word[q];
        if ((w & W_NXT) == W_EXIT) my_exitw = ((w >> W_CNT_SHIFT) & 0x3FFFu) + 1u;
        my_inmask = codes[(plr + 1) * CP + 16 + plc] != 0 ? (in_mask(codes, plr, plc) & out_mask(plr, plc)) : 0u;
        if (my_inmask) {
            int steps = 0;
            while ((w & W_NXT) < W_EXIT && steps < TCELLS) { q = w & W_NXT; w = word[q]; ++steps; }
            if ((w & W_NXT) == W_EXIT) {
                const uint32_t ql = logical_of(q);
                const int qr = (int)(ql / T), qc = (int)(ql % T);
                my_term = (uint32_t)slot_of(qr, qc);
                atomicAdd(&nterm_s[my_term], 1u);
                int dr, dc;
                d8_offset(codes[(qr + 1) * CP + 16 + qc], dr, dc);
                const int64_t tr = r0 + qr + dr, tc = c0 + qc + dc;
                // leaving the band: remember the column of the band cell the path leaves through
                if (tr < 0) my_link = LINK_OUT | (uint32_t)(c0 + qc);
                else if (tr >= v.rows) my_link = LINK_OUT | LINK_BELOW | (uint32_t)(c0 + qc);
                else my_link = (uint32_t)node_of_cell(tr, tc, v.tiles_x);
            }
        }
    }
    __syncthreads();
    if (tid < SLOTS) {
        const size_t node = (size_t)tile * SLOTS + tid;
        exitw[node] = my_exitw;
        link[node] = my_link;
        meta[node] = my_term | (nterm_s[tid] << 8) | (my_inmask << 16);
    }
}

// ---- T2: add the resolved inflow of the entry nodes along their in-tile paths --------------------------
// acc already holds the tile-local counts; only the 64-byte runs that an entry path touches are rewritten.
// The walks run over the successor table T1 wrote for the tile (16 bits per cell, see NX_*).
// HAND = true fuses the first pass of the HAND stage (hand.cu, H1) into the same tile visit: with the river
// mask defined as acc > threshold (example.py:52) every river cell lies on an entry path or has a local
// count above the threshold, so once the inflow is added the tile knows its river cells and the entry nodes
// can walk to their first river cell / failure / next entry node right away.

// REDO = true is the second run of the HAND part on a grid with D8 cycles: the flat sweep has rewritten acc since, so
// the river bits of the table and the entry states are derived again from the final counts (nothing is added to acc).
template <typename ACC, bool HAND, bool REDO>
__device__ __forceinline__ void
fa_finish_tile(const int tile, const TileView &v, const uint32_t *__restrict__ meta, const uint32_t *__restrict__ link,
               const unsigned long long *__restrict__ nstate, ACC *__restrict__ acc,
               unsigned long long *__restrict__ counters, int64_t thr, unsigned long long *__restrict__ hand_nstate,
               unsigned *__restrict__ hand_active, uint16_t *__restrict__ table)
{
    typedef typename std::conditional<sizeof(ACC) == 8, unsigned long long, uint32_t>::type EXT;
    __shared__ uint16_t nxt[TCELLS];
    // inflow added to each cell.  64-bit counts keep the two halves in separate words: a 64-bit shared-memory atomicAdd is
    // a compare-and-swap loop (3.5x the time of the 32-bit form at 100 000 x 100 000), two native 32-bit adds with a
    // carry are not
    constexpr bool WIDE = sizeof(ACC) == 8;
    __shared__ uint32_t ext[TCELLS], ext_hi[WIDE ? TCELLS : 1];
    const int tid = threadIdx.x;
    const int ty = tile / v.tiles_x, tx = tile - ty * v.tiles_x;
    const int64_t r0 = (int64_t)ty * T, c0 = (int64_t)tx * T;
    const bool fast = (v.cols % 16 == 0) && (c0 + T <= v.cols);
    // 32-bit counts: the tile's acc rows are fetched into shared memory asynchronously, right now -- the walks below hide
    // the latency, and the pass no longer holds four 16-byte loads per thread in flight at its end (the registers for
    // those set the occupancy).  Thread t fetches its own 64-byte run, so no barrier is involved; the 16-byte chunks of a
    // run are rotated by t / 2 to keep the 128-bit shared accesses of a quarter-warp on distinct banks.
    constexpr bool PRE = !WIDE;
    __shared__ __align__(16) uint4 accs[PRE ? TCELLS / 4 : 1];
    const bool pre = PRE && fast && ((reinterpret_cast<uintptr_t>(acc) & 15u) == 0) && (r0 + (tid >> 2) < v.rows);
    if (PRE && pre) {
        const ACC *src = acc + (r0 + (tid >> 2)) * v.cols + c0 + (tid & 3) * CPT;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t sa = (uint32_t)__cvta_generic_to_shared(&accs[tid * 4 + ((k + (tid >> 1)) & 3)]);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(reinterpret_cast<const uint4 *>(src) + k) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

    // Everything this tile needs from the node arrays is requested up front, next to the table: the kernel is a chain
    // of dependent latencies (table -> walk -> acc), these loads would add three more links to it.
    const size_t node = (size_t)tile * SLOTS + tid;
    const uint32_t my_meta = tid < USED_SLOTS ? meta[node] : 0u;
    const uint64_t my_ns = tid < USED_SLOTS ? nstate[node] : 0ull;
    const uint32_t my_link = (HAND && tid < USED_SLOTS) ? link[node] : LINK_NONE;

    // the successor table T1 left for this tile -> shared memory (slot layout)
    const int lr = tid >> 2, lcb = (tid & 3) * CPT;
    unsigned validmask = 0;
    {
        const uint4 *tp = reinterpret_cast<const uint4 *>(table + (size_t)tile * TCELLS + tid * CPT);
        const uint4 ta = __ldcg(tp), tb = __ldcg(tp + 1);
        const uint32_t tw[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
            uint32_t e = (tw[i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
            if (REDO) e &= ~NX_RIVER;  // river bits of the first run
            nxt[i * FT_THREADS + tid] = (uint16_t)e;
            ext[i * FT_THREADS + tid] = 0;
            if (WIDE) ext_hi[i * FT_THREADS + tid] = 0;
            validmask |= (e != NX_NODATA ? 1u : 0u) << i;
        }
    }
    __syncthreads();

    // ---- entry nodes push their resolved inflow down their in-tile path ----
    // The same walk records where the path ends and how many cardinal / diagonal moves it takes: in a tile
    // without river cells that already is the entry node's HAND state.
    const uint32_t my_inmask = (my_meta >> 16) & 0xFFu;
    uint32_t my_slot = 0, path_moves = 0, path_end = W_TERM, path_last = 0;  // moves: n_card | n_diag << 16
    unsigned unresolved = 0;
    if (my_inmask) {
        int plr, plc;
        slot_cell(tid, plr, plc);
        my_slot = phys_of((uint32_t)(plr * T + plc));
        const uint64_t ns = my_ns;
        if (!REDO && ((ns >> N_PEND_SHIFT) & N_PEND)) ++unresolved;  // never finalised: node-level cycle
        const EXT w = REDO ? (EXT)0 : (EXT)(ns & N_CNT);
        uint32_t q = my_slot, n16 = 0, ndiag = 0;
        int steps = 0;
        for (; steps < TCELLS; ++steps) {
            if (!WIDE) {
                atomicAdd(&ext[q], (uint32_t)w);
            } else {
                const uint32_t lo = (uint32_t)w, old = atomicAdd(&ext[q], lo);
                const uint32_t hi = (uint32_t)((uint64_t)w >> 32) + ((uint32_t)(old + lo) < old ? 1u : 0u);
                if (hi) atomicAdd(&ext_hi[q], hi);
            }
            n16 = nxt[q];
            if ((n16 & W_NXT) >= W_EXIT) break;
            ndiag += n16 >> 14;  // no river bits yet: bit 14 = diagonal move
            q = n16 & W_NXT;
        }
        path_end = steps == TCELLS ? NX_CYCLE : (n16 & W_NXT);
        if (path_end == W_EXIT) { ++steps; ndiag += (n16 >> 14) & 1u; path_last = q; }  // the exit move itself
        path_moves = (uint32_t)(steps - (int)ndiag) | (ndiag << 16);
    }
    unresolved = __reduce_add_sync(0xffffffffu, unresolved);
    if ((tid & 31) == 0 && unresolved) atomicAdd(&counters[0], (unsigned long long)unresolved);
    __syncthreads();

    // ---- acc += inflow on the touched runs; river bits of my 16 cells ----
    const int64_t gr = r0 + lr;
    unsigned riv = 0;
    if (gr < v.rows) {
        EXT e[CPT];
        bool any = false;
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
            e[i] = (EXT)ext[i * FT_THREADS + tid];
            if (WIDE) e[i] |= (EXT)((uint64_t)ext_hi[i * FT_THREADS + tid] << 32);
            any |= e[i] != 0;
        }
        // cells off every entry path keep their tile-local count (< 4096): they can only be river cells for tiny thresholds
        const bool need = any || (HAND && (REDO || thr < (int64_t)TCELLS));
        if (need) {
            ACC *dst = acc + gr * v.cols + c0 + lcb;
            if (fast && ((reinterpret_cast<uintptr_t>(acc) & 15u) == 0)) {
                constexpr int V = 16 / sizeof(ACC), NV = CPT / V;
                uint4 w[NV];
                if (PRE) {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
                    for (int k = 0; k < NV; ++k) w[k] = accs[tid * 4 + ((k + (tid >> 1)) & 3)];
                } else {
#pragma unroll
                    for (int k = 0; k < NV; ++k) w[k] = __ldcg(reinterpret_cast<const uint4 *>(dst) + k);
                }
                // (the kernel is bound by the latency of these loads: __launch_bounds__(.., 4) gives the scheduler the
                // registers to keep all of them in flight instead of interleaving them with the arithmetic)
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                    ACC *a = reinterpret_cast<ACC *>(&w[k]);
#pragma unroll
                    for (int j = 0; j < V; ++j) {
                        a[j] += (ACC)e[k * V + j];
                        if (HAND) riv |= (unsigned)((int64_t)a[j] > thr) << (k * V + j);
                    }
                }
                if (any) {
#pragma unroll
                    for (int k = 0; k < NV; ++k) reinterpret_cast<uint4 *>(dst)[k] = w[k];
                }
            } else {
#pragma unroll
                for (int i = 0; i < CPT; ++i)
                    if (c0 + lcb + i < v.cols) {
                        const ACC a = dst[i] + (ACC)e[i];
                        if (e[i] != 0) dst[i] = a;
                        if (HAND) riv |= (unsigned)((int64_t)a > thr) << i;
                    }
            }
        }
    }
    if (!HAND) return;
    riv &= validmask;  // a cell without a direction code is nodata for HAND (flowhand.py:601)
    // river cells get bit 15 of their successor entry; a tile without any keeps the states of the first walk
    if (riv) {
#pragma unroll
        for (int i = 0; i < CPT; ++i)
            if ((riv >> i) & 1u) nxt[i * FT_THREADS + tid] |= NX_RIVER;
    }
    const bool tile_has_river = __syncthreads_or(riv != 0);
    if (riv || REDO) {  // patch the river bits into the persistent table (HAND's tile pass reads it)
        uint32_t w[CPT / 2];
#pragma unroll
        for (int i = 0; i < CPT; i += 2)
            w[i / 2] = (uint32_t)nxt[i * FT_THREADS + tid] | ((uint32_t)nxt[(i + 1) * FT_THREADS + tid] << 16);
        uint4 *dst = reinterpret_cast<uint4 *>(table + (size_t)tile * TCELLS + tid * CPT);
        dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
        dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }

    // ---- HAND first pass: entry nodes walk to their first river cell / failure / next entry node ----
    uint64_t hs = 0ull;  // inactive slot (never referenced)
    int is_active = 0;
    if (my_inmask) {
        uint32_t q = my_slot, moves = 0, end = NX_CYCLE;
        if (!tile_has_river) {
            moves = path_moves; end = path_end; q = path_last;
        } else {
            for (int steps = 0; steps <= TCELLS; ++steps) {
                const uint32_t n16 = nxt[q], n = n16 & W_NXT;
                if (n16 & NX_RIVER) { end = NX_RIVER; break; }  // flowhand.py:622
                if (n == W_TERM) { end = W_TERM; break; }     // unknown code, off-raster or code-0 landing (flowhand.py:623-764, 826, 830)
                moves += 1u + ((n16 >> 14) & 1u) * 0xFFFFu;
                if (n == W_EXIT) { end = W_EXIT; break; }
                q = n;
            }
        }
        const uint32_t nc = moves & 0xFFFFu, nd = moves >> 16;
        hs = pack(KIND_FAIL, 0, 0, 0);  // W_TERM, or an in-tile cycle
        if (end == NX_RIVER) {
            const uint32_t pl = logical_of(q);
            hs = pack(KIND_RIVER, nd, nc, (uint32_t)((r0 + (pl >> 6)) * v.cols + c0 + (pl & (T - 1))));
        } else if (end == W_EXIT) {
            // the path leaves through the exit T1 found for this entry: link = the entry node it lands on
            const uint32_t l = my_link;
            if (!(l & LINK_OUT)) { hs = pack(KIND_ACTIVE, nd, nc, l); is_active = 1; }
            else if (l != LINK_NONE) {
                // out of the band: HAND wants the column of the halo cell it lands on (T1 stored the exit cell's column)
                const int64_t xc = (int64_t)(l & 0x3FFFFFFFu);
                const int64_t xr = (l & LINK_BELOW) ? v.rows - 1 : 0;
                int dr, dc;
                d8_offset(v.d8[xr * v.cols + xc], dr, dc);
                hs = pack(KIND_EXIT, nd, nc, (l & (LINK_OUT | LINK_BELOW)) | (uint32_t)(xc + dc));
            }
        }
    }
    hand_nstate[(size_t)tile * SLOTS + tid] = hs;
    const unsigned ballot = __ballot_sync(0xffffffffu, is_active);
    if ((tid & 31) == 0 && ballot) atomicAdd(&hand_active[0], (unsigned)__popc(ballot));
}

// 5 CTAs per SM for 32-bit counts (47 registers, 40 KB of shared memory each): with the acc rows prefetched the pass has no
// long-latency loads left at its end to keep in flight, and what it needs is more tiles in flight -- a tile spends most of
// its time waiting for its longest entry walk (40k x 40k: 6.01 -> 5.71 ms; 4 CTAs with the prefetch alone: no change).
#ifndef FT2_MINB
#define FT2_MINB 5
#endif
template <typename ACC, bool HAND>
__global__ void __launch_bounds__(FT_THREADS, sizeof(ACC) == 8 ? 4 : FT2_MINB)
fa_tile_finish_kernel(TileView v, const uint32_t *__restrict__ meta, const uint32_t *__restrict__ link,
                      const unsigned long long *__restrict__ nstate, ACC *__restrict__ acc,
                      unsigned long long *__restrict__ counters, int64_t thr, unsigned long long *__restrict__ hand_nstate,
                      unsigned *__restrict__ hand_active, uint16_t *__restrict__ table)
{
    fa_finish_tile<ACC, HAND, false>((int)blockIdx.x, v, meta, link, nstate, acc, counters, thr, hand_nstate, hand_active, table);
}

// cyclic grids only: returns at once when counters[0] == 0 (the count of unfinalised cells / nodes)
template <typename ACC>
__global__ void __launch_bounds__(FT_THREADS, 4)
fa_tile_rehand_kernel(TileView v, int64_t tiles, const uint32_t *__restrict__ meta, const uint32_t *__restrict__ link,
                      const unsigned long long *__restrict__ nstate, ACC *__restrict__ acc,
                      unsigned long long *__restrict__ counters, int64_t thr, unsigned long long *__restrict__ hand_nstate,
                      unsigned *__restrict__ hand_active, uint16_t *__restrict__ table)
{
    if (counters[0] == 0ull) return;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        fa_finish_tile<ACC, true, true>((int)tile, v, meta, link, nstate, acc, counters, thr, hand_nstate, hand_active, table);
        __syncthreads();
    }
}

// ---- N: entry-node forest ---------------------------------------------------------------------
// One CTA per tile (its 256 slots); 32-bit arithmetic throughout -- the first form (flat 64-bit node numbers, a 64-bit
// division per node) kept the issue slots 64 % busy for what is a gather of two words per feeding neighbour.
__global__ void __launch_bounds__(SLOTS)
fa_node_init_kernel(int rows, int cols, int tiles_x, const uint32_t *__restrict__ exitw, const uint32_t *__restrict__ meta,
                    const int64_t *__restrict__ inflow_above, const int64_t *__restrict__ inflow_below,
                    unsigned long long *__restrict__ nstate)
{
    const int tile = blockIdx.x, s = threadIdx.x;
    const size_t node = (size_t)tile * SLOTS + s;
    const uint32_t m = meta[node];
    const unsigned inmask = (m >> 16) & 0xFFu;
    if (!inmask) { nstate[node] = 0ull; return; }
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    int lr, lc;
    slot_cell(s, lr, lc);
    const int gr = ty * T + lr, gc = tx * T + lc;
    uint64_t base = 0, pend = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (!((inmask >> k) & 1u)) continue;
        int dr, dc;
        nbr_offset(k, dr, dc);
        const int ur = gr + dr, uc = gc + dc;
        if (ur < 0) base += inflow_above ? (uint64_t)inflow_above[uc] : 0ull;
        else if (ur >= rows) base += inflow_below ? (uint64_t)inflow_below[uc] : 0ull;
        else {
            const uint32_t un = (uint32_t)(((ur >> 6) * tiles_x + (uc >> 6)) * SLOTS + slot_of(ur & (T - 1), uc & (T - 1)));
            base += exitw[un];
            pend += (meta[un] >> 8) & 0xFFu;
        }
    }
    nstate[node] = N_ACTIVE | (pend == 0 ? N_SRC : 0ull) | (pend << N_PEND_SHIFT) | (base & N_CNT);
}

__global__ void __launch_bounds__(256)
fa_node_sweep_kernel(int64_t nnodes, const uint32_t *__restrict__ link, unsigned long long *nstate)
{
    const int64_t node = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (node >= nnodes) return;
    const uint64_t s = nstate[node];
    if (!(s & N_SRC)) return;
    uint64_t w = s & N_CNT;
    // The walk is a chain of dependent round trips to L2 (the main stem: ~2 000 nodes at 40k x 40k).  The link of the
    // NEXT node is requested together with the atomic on it, not after its answer: one round trip per node, not two.
    uint32_t l = __ldg(&link[node]);
    while (!(l & LINK_OUT)) {  // LINK_NONE or out of the band
        const uint32_t lnext = __ldg(&link[l]);
        const uint64_t old = atomicAdd(&nstate[l], (unsigned long long)(w - N_PEND_ONE));
        if (((old >> N_PEND_SHIFT) & N_PEND) != 1ull) break;
        w += old & N_CNT;
        l = lnext;
    }
}

// Band FINISH: the SUMMARY call has already swept the node forest with zero inflow from the halo rows.  What the
// boundary solve adds is a constant per seam entry: every node downstream of the entry gains exactly that much, so the
// forest is not swept again -- each entry node of the first / last tile row adds its halo inflow to itself and to the
// nodes along its chain (plain atomic adds: they commute, no pending counts involved).
__global__ void __launch_bounds__(256)
fa_node_inflow_kernel(int64_t tile0, int64_t nslots, int64_t rows, int64_t cols, int tiles_x, const uint32_t *__restrict__ meta,
                      const uint32_t *__restrict__ link, const int64_t *__restrict__ inflow_above,
                      const int64_t *__restrict__ inflow_below, unsigned long long *nstate)
{
    const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (k >= nslots) return;
    const int64_t node = tile0 * SLOTS + k;
    const unsigned inmask = (meta[node] >> 16) & 0xFFu;
    if (!inmask) return;
    const int64_t tile = node / SLOTS;
    int lr, lc;
    slot_cell((int)(node % SLOTS), lr, lc);
    const int64_t gr = (tile / tiles_x) * T + lr, gc = (tile % tiles_x) * T + lc;
    uint64_t delta = 0;
    for (int b = 0; b < 8; ++b) {
        if (!((inmask >> b) & 1u)) continue;
        int dr, dc;
        nbr_offset(b, dr, dc);
        const int64_t ur = gr + dr, uc = gc + dc;
        if (ur < 0) delta += inflow_above ? (uint64_t)inflow_above[uc] : 0ull;
        else if (ur >= rows) delta += inflow_below ? (uint64_t)inflow_below[uc] : 0ull;
    }
    if (!delta) return;
    delta &= N_CNT;
    atomicAdd(&nstate[node], (unsigned long long)delta);
    uint32_t l = __ldg(&link[node]);
    while (!(l & LINK_OUT)) {  // LINK_NONE or out of the band
        atomicAdd(&nstate[l], (unsigned long long)delta);  // no return value needed: the loads alone form the chain
        l = __ldg(&link[l]);
    }
}

// ---- cyclic grids: single-level sweep over the whole raster (gated on counters[0] != 0) -----------
constexpr uint64_t F_CNT = (1ull << 44) - 1ull, F_PEND_ONE = 1ull << 44, F_SRC = 1ull << 63;
constexpr int FLAT_BLOCKS = kNumSMs * 8;

// counters: [0] cyclic evidence, [1] valid cells, [2] finalised cells
__global__ void __launch_bounds__(256)
fa_flat_init_kernel(TileView v, const int64_t *__restrict__ inflow_above, const int64_t *__restrict__ inflow_below,
                    unsigned long long *__restrict__ state, unsigned long long *__restrict__ counters)
{
    if (counters[0] == 0) return;
    const int64_t n = v.rows * v.cols;
    unsigned long long valid = 0;
    for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < n; p += (int64_t)gridDim.x * 256) {
        const unsigned code = v.d8[p];
        if (code == 0) { state[p] = 0ull; continue; }
        ++valid;
        const int64_t r = p / v.cols, c = p - r * v.cols;
        uint64_t pending = 0, seed = 0;
        const unsigned want[8] = {2, 4, 8, 1, 16, 128, 64, 32};
        for (int k = 0; k < 8; ++k) {
            int dr, dc;
            nbr_offset(k, dr, dc);
            if (fetch_code(v, r + dr, c + dc) != want[k]) continue;
            if (r + dr < 0) seed += inflow_above ? (uint64_t)inflow_above[c + dc] : 0ull;
            else if (r + dr >= v.rows) seed += inflow_below ? (uint64_t)inflow_below[c + dc] : 0ull;
            else ++pending;
        }
        state[p] = (pending == 0 ? F_SRC : 0ull) | (pending << 44) | (seed & F_CNT);
    }
    valid = __reduce_add_sync(0xffffffffu, (unsigned)valid);
    if ((threadIdx.x & 31) == 0 && valid) atomicAdd(&counters[1], valid);
}

template <typename ACC>
__global__ void __launch_bounds__(256)
fa_flat_sweep_kernel(TileView v, ACC *__restrict__ acc, unsigned long long *state, unsigned long long *__restrict__ counters)
{
    if (counters[0] == 0) return;
    const int64_t n = v.rows * v.cols;
    unsigned finalised = 0;
    for (int64_t p0 = (int64_t)blockIdx.x * 256 + threadIdx.x; p0 < n; p0 += (int64_t)gridDim.x * 256) {
        const uint64_t w = state[p0];
        if (!(w & F_SRC)) continue;
        uint64_t carry = w & F_CNT;
        int64_t p = p0;
        unsigned code = v.d8[p];
        acc[p] = (ACC)carry;
        ++finalised;
        int64_t r = p / v.cols, c = p - r * v.cols;
        for (;;) {
            int dr, dc;
            if (!d8_offset(code, dr, dc)) break;
            r += dr;
            c += dc;
            if (r < 0 || r >= v.rows || c < 0 || c >= v.cols) break;
            p = r * v.cols + c;
            code = v.d8[p];
            if (code == 0) break;
            const uint64_t old = atomicAdd(&state[p], (unsigned long long)((carry + 1ull) - F_PEND_ONE));
            if (((old >> 44) & 0xFull) != 1ull) break;
            carry = (old & F_CNT) + carry + 1ull;
            acc[p] = (ACC)carry;
            ++finalised;
        }
    }
    finalised = __reduce_add_sync(0xffffffffu, finalised);
    if ((threadIdx.x & 31) == 0 && finalised) atomicAdd(&counters[2], (unsigned long long)finalised);
}

template <typename ACC>
__global__ void __launch_bounds__(256)
fa_flat_fix_kernel(TileView v, ACC *__restrict__ acc, const unsigned long long *__restrict__ state,
                   const unsigned long long *__restrict__ counters)
{
    if (counters[0] == 0) return;
    const int64_t n = v.rows * v.cols;
    for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < n; p += (int64_t)gridDim.x * 256) {
        if (v.d8[p] == 0) continue;
        const uint64_t w = state[p];
        if ((w >> 44) & 0xFull) acc[p] = (ACC)(w & F_CNT);  // never finalised: partial count (oracle: Kahn leftovers)
    }
}

// ---- band summary (multi-GPU): the same construction one level up ------------------------------
// For the band's first (side 0) or last (side 1) row:
//   exit_out[c] = (band-local acc)+1 of a cell that drains into the halo row, else 0.  The band-local
//                 count of such a perimeter cell is its tile count (exitw-1) plus the resolved inflow of
//                 every entry node of its tile that ends there (scattered by fa_band_exit_scatter_kernel);
//   term_out[c] = -2 if the cell takes no flow from that halo row, else where the in-band path that
//                 starts there leaves the band: (side << 30) | column of the band cell it leaves through, or -1.
__device__ __forceinline__ bool drains_across(const TileView &v, int side, int64_t r, int64_t c)
{
    int dr, dc;
    if (!d8_offset(v.d8[r * v.cols + c], dr, dc)) return false;
    if (side ? !(dr > 0 && r + 1 == v.rows) : !(dr < 0 && r == 0)) return false;
    return fetch_code(v, r + dr, c + dc) != 0;
}

__global__ void __launch_bounds__(256)
fa_band_summary_kernel(TileView v, int side, const uint32_t *__restrict__ exitw, const uint32_t *__restrict__ link,
                       const uint32_t *__restrict__ meta, long long *__restrict__ exit_out, int32_t *__restrict__ term_out)
{
    const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (c >= v.cols) return;
    const int64_t r = side ? v.rows - 1 : 0;
    const int64_t node = node_of_cell(r, c, v.tiles_x);
    exit_out[c] = drains_across(v, side, r, c) ? (long long)exitw[node] : 0ll;
    const unsigned inmask = (meta[node] >> 16) & 0xFFu;
    const bool fed = (inmask & (side ? 0xE0u : 0x07u)) != 0;  // tributaries in the halo row (band seams are tile seams)
    int32_t t = -2;
    if (fed) {
        t = -1;
        uint32_t q = (uint32_t)node;
        for (;;) {
            const uint32_t l = link[q];
            if (l == LINK_NONE) break;
            if (l & LINK_OUT) { t = (int32_t)(((l & LINK_BELOW) ? 1u << 30 : 0u) | (l & 0x3FFFFFFFu)); break; }
            q = l;
        }
    }
    term_out[c] = t;
}

// one thread per node of the band's first / last tile row
__global__ void __launch_bounds__(256)
fa_band_exit_scatter_kernel(TileView v, int side, int64_t tile0, const uint32_t *__restrict__ meta,
                            const unsigned long long *__restrict__ nstate, long long *__restrict__ exit_out)
{
    const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (k >= (int64_t)v.tiles_x * SLOTS) return;
    const int64_t node = tile0 * SLOTS + k;
    const uint32_t m = meta[node];
    if (!((m >> 16) & 0xFFu)) return;           // not an entry node
    const unsigned ts = m & 0xFFu;
    if (ts == TERM_NONE) return;
    int lr, lc;
    slot_cell((int)ts, lr, lc);
    const int64_t tile = node / SLOTS;
    const int64_t r = (tile / v.tiles_x) * T + lr, c = (tile % v.tiles_x) * T + lc;
    if (r != (side ? v.rows - 1 : 0) || !drains_across(v, side, r, c)) return;
    atomicAdd(reinterpret_cast<unsigned long long *>(&exit_out[c]), (unsigned long long)(nstate[node] & N_CNT));
}

// ---- generic forest accumulation (band boundary graph, bands.py) -------------------------------------
// out[i] = base[i] + sum of out[j] over j with next[j] == i (next < 0: none): the same last-arriver sweep on
// a caller-supplied forest.  state: [63 source | 62..44 pending | 43..0 count].
__global__ void __launch_bounds__(256)
forest_init_kernel(int64_t n, const long long *__restrict__ next, const long long *__restrict__ base, unsigned long long *state)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    atomicAdd(&state[i], (unsigned long long)base[i] & F_CNT);  // state was zeroed; tributaries may already have counted in
    const long long t = next[i];
    if (t >= 0 && t < n) atomicAdd(&state[t], F_PEND_ONE);
}
__global__ void __launch_bounds__(256)
forest_sweep_kernel(int64_t n, const long long *__restrict__ next, unsigned long long *state, const unsigned long long *__restrict__ pend0)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    if ((pend0[i] >> 44) != 0ull) return;  // not a source (pending as counted before the sweep started)
    uint64_t w = pend0[i] & F_CNT;
    long long t = next[i];
    while (t >= 0 && t < n) {
        const long long tnext = next[t];  // requested together with the atomic (see fa_node_sweep_kernel)
        const uint64_t old = atomicAdd(&state[t], (unsigned long long)(w - F_PEND_ONE));
        if (((old >> 44) & 0x7FFFFull) != 1ull) break;
        w += old & F_CNT;
        t = tnext;
    }
}
__global__ void __launch_bounds__(256)
forest_out_kernel(int64_t n, const unsigned long long *__restrict__ state, long long *__restrict__ out, int *__restrict__ unresolved)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const uint64_t s = state[i];
    out[i] = (long long)(s & F_CNT);
    if ((s >> 44) & 0x7FFFFull) atomicOr(unresolved, 1);  // never finalised: the forest has a cycle
}

// ---- boundary graph between row bands (multi-GPU): the whole solve in the library ------------------------------
// summ int64 [N][6][cols]: exit_above, exit_below, term_above, term_below, d8 first row, d8 last row of every band
// (dtb_flowacc_band, DTB_FA_SUMMARY).  Node (b, side, c) = cell c of the first (side 0) / last (side 1) row of band b.
__global__ void __launch_bounds__(256)
fb_prep_kernel(const long long *__restrict__ summ, int64_t nbands, int64_t cols, long long *__restrict__ next, long long *__restrict__ base)
{
    const int64_t node = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (node >= 2 * nbands * cols) return;
    const int64_t b = node / (2 * cols), rem = node - b * 2 * cols, side = rem / cols, c = rem - side * cols;
    const long long w = summ[(b * 6 + side) * cols + c];
    const long long code = summ[(b * 6 + 4 + side) * cols + c];
    // column shift of the move across the seam: NW/SW -1, N/S 0, NE/SE +1 (flowhand.py:801-824)
    const int64_t c2 = c + ((code == 128 || code == 2) ? 1 : 0) - ((code == 32 || code == 8) ? 1 : 0);
    const int64_t b2 = side == 0 ? b - 1 : b + 1;
    long long nx = -1;
    if (w > 0 && b2 >= 0 && b2 < nbands && c2 >= 0 && c2 < cols) {
        const long long t = summ[(b2 * 6 + 2 + (1 - side)) * cols + c2];  // where the landing cell's in-band path leaves
        if (t >= 0) nx = (b2 * 2 + ((t >> 30) & 1)) * cols + (t & 0x3FFFFFFF);
    }
    next[node] = nx;
    base[node] = w;
}
// inflow [N][2][cols]: (acc + 1) carried by the halo row above / below each band
__global__ void __launch_bounds__(256)
fb_inflow_kernel(const long long *__restrict__ f, int64_t nbands, int64_t cols, long long *__restrict__ inflow)
{
    const int64_t node = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (node >= 2 * nbands * cols) return;
    const int64_t b = node / (2 * cols), rem = node - b * 2 * cols, side = rem / cols, c = rem - side * cols;
    const int64_t b2 = side == 0 ? b - 1 : b + 1;  // above band b = last row of band b-1; below = first row of band b+1
    inflow[node] = (b2 >= 0 && b2 < nbands) ? f[(b2 * 2 + (1 - side)) * cols + c] : 0;
}

struct NodeLayout {
    int64_t tiles, nnodes;
    size_t off_counters, off_exitw, off_link, off_meta, off_nstate, off_flat, total;
};

NodeLayout layout(int64_t rows, int64_t cols)
{
    NodeLayout L;
    const int64_t tx = (cols + T - 1) / T, ty = (rows + T - 1) / T;
    L.tiles = tx * ty;
    L.nnodes = L.tiles * SLOTS;
    size_t o = 0;
    L.off_counters = o; o += 256;
    L.off_exitw = o; o += (size_t)L.nnodes * 4;
    L.off_link = o; o += (size_t)L.nnodes * 4;
    L.off_meta = o; o += (size_t)L.nnodes * 4;
    L.off_nstate = o; o += (size_t)L.nnodes * 8;
    o = (o + 255) & ~(size_t)255;
    // flat per-cell state, only touched for cyclic grids; until then the region carries the 16-bit successor
    // table between the two tile passes when the caller gave no HAND workspace to keep it in
    const size_t flat_bytes = (size_t)rows * (size_t)cols * 8, table_bytes = (size_t)L.tiles * TCELLS * 2;
    L.off_flat = o; o += flat_bytes > table_bytes ? flat_bytes : table_bytes;
    L.total = o;
    return L;
}

template <typename ACC>
int run(const dtb_flowacc_args *a, void *ws, cudaStream_t st)
{
    const NodeLayout L = layout(a->rows, a->cols);
    char *base = reinterpret_cast<char *>(ws);
    unsigned long long *counters = reinterpret_cast<unsigned long long *>(base + L.off_counters);
    uint32_t *exitw = reinterpret_cast<uint32_t *>(base + L.off_exitw);
    uint32_t *link = reinterpret_cast<uint32_t *>(base + L.off_link);
    uint32_t *meta = reinterpret_cast<uint32_t *>(base + L.off_meta);
    unsigned long long *nstate = reinterpret_cast<unsigned long long *>(base + L.off_nstate);
    unsigned long long *flat = reinterpret_cast<unsigned long long *>(base + L.off_flat);
    TileView v{a->d8, a->halo_above, a->halo_below, a->rows, a->cols, (int)((a->cols + T - 1) / T)};
    ACC *acc = reinterpret_cast<ACC *>(a->acc);
    const unsigned nb_nodes = (unsigned)((L.nnodes + 255) / 256);
    // successor table: behind the node states of the HAND workspace if there is one (HAND's tile pass reads it), else
    // in the flat region of this workspace
    unsigned *hactive = reinterpret_cast<unsigned *>(a->hand_ws);
    unsigned long long *hstate = a->hand_ws ? reinterpret_cast<unsigned long long *>((char *)a->hand_ws + 256) : nullptr;
    uint16_t *table = a->hand_ws ? reinterpret_cast<uint16_t *>(hstate + L.nnodes) : reinterpret_cast<uint16_t *>(flat);

    if (a->mode != DTB_FA_FINISH) {
        DTB_CUDA(cudaMemsetAsync(counters, 0, 256, st));
        DTB_KERNEL("fa_tile_kernel", st, fa_tile_kernel<ACC><<<(unsigned)L.tiles, FT_THREADS, 0, st>>>(v, exitw, link, meta, acc, (ACC)a->nodata_fill, counters, table));
    }
    if (a->mode != DTB_FA_FINISH) {
        static_assert(T == 64 && SLOTS == 256, "fa_node_init_kernel: one CTA per tile, shifts by 6");
        DTB_KERNEL("fa_node_init_kernel", st, fa_node_init_kernel<<<(unsigned)L.tiles, SLOTS, 0, st>>>((int)a->rows, (int)a->cols, v.tiles_x, exitw, meta,
                                                     a->inflow_above, a->inflow_below, nstate));
        DTB_KERNEL("fa_node_sweep_kernel", st, fa_node_sweep_kernel<<<nb_nodes, 256, 0, st>>>(L.nnodes, link, nstate));
    } else {
        // the node states of the SUMMARY call (zero halo inflow) + what each seam entry receives, pushed down its chain
        const int64_t tiles_y = (a->rows + T - 1) / T;
        const int64_t nslots = (int64_t)v.tiles_x * SLOTS;
        const unsigned nbs = (unsigned)((nslots + 255) / 256);
        if (a->inflow_above)
            DTB_KERNEL("fa_node_inflow_kernel", st, fa_node_inflow_kernel<<<nbs, 256, 0, st>>>(0, nslots, a->rows, a->cols, v.tiles_x, meta, link,
                                                               a->inflow_above, tiles_y == 1 ? a->inflow_below : nullptr, nstate));
        if (a->inflow_below && tiles_y > 1)
            DTB_KERNEL("fa_node_inflow_kernel", st, fa_node_inflow_kernel<<<nbs, 256, 0, st>>>((tiles_y - 1) * v.tiles_x, nslots, a->rows, a->cols, v.tiles_x,
                                                               meta, link, nullptr, a->inflow_below, nstate));
        else if (a->inflow_below && !a->inflow_above)
            DTB_KERNEL("fa_node_inflow_kernel", st, fa_node_inflow_kernel<<<nbs, 256, 0, st>>>(0, nslots, a->rows, a->cols, v.tiles_x, meta, link, nullptr,
                                                               a->inflow_below, nstate));
    }

    if (a->mode == DTB_FA_SUMMARY) {
        const unsigned nbc = (unsigned)((a->cols + 255) / 256);
        const unsigned nbs = (unsigned)(((int64_t)v.tiles_x * SLOTS + 255) / 256);
        const int64_t tiles_y = (a->rows + T - 1) / T;
        for (int side = 0; side < 2; ++side) {
            long long *ex = reinterpret_cast<long long *>(side ? a->exit_below : a->exit_above);
            int32_t *tm = side ? a->term_below : a->term_above;
            if (!ex || !tm) continue;
            DTB_KERNEL("fa_band_summary_kernel", st, fa_band_summary_kernel<<<nbc, 256, 0, st>>>(v, side, exitw, link, meta, ex, tm));
            DTB_KERNEL("fa_band_exit_scatter_kernel", st, fa_band_exit_scatter_kernel<<<nbs, 256, 0, st>>>(v, side, side ? (tiles_y - 1) * v.tiles_x : 0, meta, nstate, ex));
        }
        return DTB_OK;
    }

    if (a->hand_ws) {
        // fused HAND first pass: node states and the ACTIVE counter live at the head of the HAND workspace (hand.cu)
        DTB_CUDA(cudaMemsetAsync(hactive, 0, 256, st));
        DTB_KERNEL("fa_tile_finish_kernel<hand>", st, fa_tile_finish_kernel<ACC, true><<<(unsigned)L.tiles, FT_THREADS, 0, st>>>(
                       v, meta, link, nstate, acc, counters, a->hand_river_threshold, hstate, hactive, table));
    } else {
        DTB_KERNEL("fa_tile_finish_kernel", st, fa_tile_finish_kernel<ACC, false><<<(unsigned)L.tiles, FT_THREADS, 0, st>>>(
                       v, meta, link, nstate, acc, counters, 0, nullptr, nullptr, table));
    }
    // cyclic grids only (each kernel returns at once when counters[0] == 0)
    DTB_KERNEL("fa_flat_init_kernel", st, fa_flat_init_kernel<<<FLAT_BLOCKS, 256, 0, st>>>(v, a->inflow_above, a->inflow_below, flat, counters));
    DTB_KERNEL("fa_flat_sweep_kernel", st, fa_flat_sweep_kernel<ACC><<<FLAT_BLOCKS, 256, 0, st>>>(v, acc, flat, counters));
    DTB_KERNEL("fa_flat_fix_kernel", st, fa_flat_fix_kernel<ACC><<<FLAT_BLOCKS, 256, 0, st>>>(v, acc, flat, counters));
    if (a->hand_ws)
        DTB_KERNEL("fa_tile_rehand_kernel", st, fa_tile_rehand_kernel<ACC><<<FLAT_BLOCKS, FT_THREADS, 0, st>>>(
                       v, L.tiles, meta, link, nstate, acc, counters, a->hand_river_threshold, hstate, hactive, table));
    if (a->unfinalised_host) {
        unsigned long long h[3];
        DTB_CUDA(cudaMemcpyAsync(h, counters, sizeof(h), cudaMemcpyDeviceToHost, st));
        DTB_CUDA(cudaStreamSynchronize(st));
        *a->unfinalised_host = h[0] ? (int64_t)(h[1] - h[2]) : 0;
    }
    return DTB_OK;
}

}  // namespace
}  // namespace dtb

extern "C" size_t dtb_flowacc_workspace_bytes(int64_t rows, int64_t cols)
{
    if (rows <= 0 || cols <= 0) return 0;
    return dtb::layout(rows, cols).total;
}

extern "C" int dtb_flowacc_band(const dtb_flowacc_args *a, void *ws, size_t ws_bytes, void *stream)
{
    using namespace dtb;
    if (!a || !a->d8 || !ws || a->rows <= 0 || a->cols <= 0) return DTB_ERR_INVALID;
    if (a->mode < DTB_FA_FULL || a->mode > DTB_FA_FINISH) return DTB_ERR_INVALID;
    if (!a->acc) return DTB_ERR_INVALID;  // the tile pass leaves its local counts in acc
    if (a->acc_dtype != DTB_I32 && a->acc_dtype != DTB_I64) return DTB_ERR_INVALID;
    if (ws_bytes < dtb_flowacc_workspace_bytes(a->rows, a->cols)) return DTB_ERR_WORKSPACE;
    if (a->cols >= (int64_t)1 << 30) return DTB_ERR_UNSUPPORTED;
    if (a->halo_below && a->rows % T != 0) return DTB_ERR_INVALID;  // band seams sit on tile seams
    if (a->hand_ws && (a->hand_ws_bytes < dtb_hand_workspace_bytes(a->rows, a->cols) || a->rows * a->cols > 0xffffffffLL))
        return DTB_ERR_INVALID;
    const NodeLayout L = layout(a->rows, a->cols);
    if (L.nnodes >= (int64_t)1 << 30) return DTB_ERR_UNSUPPORTED;  // node ids share a word with the LINK_OUT flags
    if (a->rows * a->cols >= (int64_t)1 << 43) return DTB_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    return a->acc_dtype == DTB_I32 ? run<int32_t>(a, ws, st) : run<int64_t>(a, ws, st);
}

extern "C" int dtb_flowacc(const uint8_t *d8, int64_t rows, int64_t cols, void *acc, int acc_dtype, int64_t nodata_fill,
                           void *ws, size_t ws_bytes, int64_t *unfinalised_host, void *stream)
{
    dtb_flowacc_args a = {};
    a.d8 = d8;
    a.rows = rows;
    a.cols = cols;
    a.acc = acc;
    a.acc_dtype = acc_dtype;
    a.nodata_fill = nodata_fill;
    a.unfinalised_host = unfinalised_host;
    if (acc_dtype == DTB_I32 && rows * cols > 0x7fffffffLL) return DTB_ERR_UNSUPPORTED;
    return dtb_flowacc_band(&a, ws, ws_bytes, stream);
}

extern "C" size_t dtb_forest_workspace_bytes(int64_t n) { return n > 0 ? (size_t)n * 16 : 0; }

extern "C" int dtb_forest_accumulate(const int64_t *next, const int64_t *base, int64_t n, int64_t *out, int *unresolved,
                                     void *ws, size_t ws_bytes, void *stream)
{
    using namespace dtb;
    if (!next || !base || !out || !unresolved || !ws || n <= 0) return DTB_ERR_INVALID;
    if (ws_bytes < dtb_forest_workspace_bytes(n)) return DTB_ERR_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    unsigned long long *state = reinterpret_cast<unsigned long long *>(ws), *pend0 = state + n;
    const unsigned nb = (unsigned)((n + 255) / 256);
    DTB_CUDA(cudaMemsetAsync(state, 0, (size_t)n * 8, st));
    DTB_CUDA(cudaMemsetAsync(unresolved, 0, sizeof(int), st));
    DTB_KERNEL("forest_init_kernel", st, forest_init_kernel<<<nb, 256, 0, st>>>(n, (const long long *)next, (const long long *)base, state));
    DTB_CUDA(cudaMemcpyAsync(pend0, state, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));  // snapshot: who is a source
    DTB_KERNEL("forest_sweep_kernel", st, forest_sweep_kernel<<<nb, 256, 0, st>>>(n, (const long long *)next, state, pend0));
    DTB_KERNEL("forest_out_kernel", st, forest_out_kernel<<<nb, 256, 0, st>>>(n, state, (long long *)out, unresolved));
    return DTB_OK;
}

extern "C" size_t dtb_flowacc_boundary_workspace_bytes(int64_t nbands, int64_t cols)
{
    if (nbands <= 0 || cols <= 0) return 0;
    const int64_t n = 2 * nbands * cols;
    return (size_t)n * 24 + dtb_forest_workspace_bytes(n);
}

extern "C" int dtb_flowacc_boundary_solve(const int64_t *summ, int64_t nbands, int64_t cols, int64_t *inflow, int *unresolved,
                                          void *ws, size_t ws_bytes, void *stream)
{
    using namespace dtb;
    if (!summ || !inflow || !unresolved || !ws || nbands <= 0 || cols <= 0) return DTB_ERR_INVALID;
    if (cols >= (int64_t)1 << 30) return DTB_ERR_UNSUPPORTED;
    if (ws_bytes < dtb_flowacc_boundary_workspace_bytes(nbands, cols)) return DTB_ERR_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    const int64_t n = 2 * nbands * cols;
    long long *next = reinterpret_cast<long long *>(ws), *base = next + n, *f = base + n;
    const unsigned nb = (unsigned)((n + 255) / 256);
    DTB_KERNEL("fb_prep_kernel", st, (fb_prep_kernel<<<nb, 256, 0, st>>>((const long long *)summ, nbands, cols, next, base)));
    const int rc = dtb_forest_accumulate((const int64_t *)next, (const int64_t *)base, n, (int64_t *)f, unresolved, f + n,
                                         dtb_forest_workspace_bytes(n), stream);
    if (rc != DTB_OK) return rc;
    DTB_KERNEL("fb_inflow_kernel", st, (fb_inflow_kernel<<<nb, 256, 0, st>>>(f, nbands, cols, (long long *)inflow)));
    return DTB_OK;
}
