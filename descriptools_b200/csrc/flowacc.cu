// flowacc.cu -- D8 flow accumulation: acc[p] = number of cells strictly upstream of p.
//
// New stage (the reference only consumes accumulation rasters: gfi.py:432,
// topoindexes.py:252-255, example.py:52); semantics per SURVEY.md App. A3, restated in
// oracle/dt_oracle.c:orc_flowacc and pinned by the bundled 12_fdr.tif -> 12_fac.tif pair.
//
// Algorithm (single sweep, no level synchronisation, every cell visited once):
//   1. init:  per cell, pending = number of valid neighbours whose D8 code points at it
//             (a gather over the 3x3 neighbourhood, no atomics); one packed 64-bit word per
//             cell: [63 source flag | 47..44 pending | 43..0 running count (+ seed)].
//   2. sweep: one thread per source cell walks downstream.  Each step is ONE 64-bit
//             atomicAdd on the next cell's word that adds (count+1) to the low field and
//             subtracts 1 from `pending`; the returned old value tells the walker whether it
//             was the last tributary to arrive -- only then it owns the cell's final count,
//             stores it and continues.  All other walkers retire.  No thread ever waits.
//   3. fix:   only if the grid has D8 cycles (count of finalised cells != valid cells):
//             cycle cells keep their partial count, like the oracle's Kahn sweep.
#include "common.cuh"

namespace dtb {
namespace {

constexpr int FA_THREADS = 256;
constexpr uint64_t CNT_MASK = (1ull << 44) - 1ull;
constexpr uint64_t PEND_ONE = 1ull << 44;
constexpr uint64_t SRC_FLAG = 1ull << 63;

// counters: [0] valid cells, [1] finalised cells
__global__ void __launch_bounds__(FA_THREADS)
fa_init_kernel(const uint8_t *__restrict__ d8, int64_t rows, int64_t cols, const int64_t *__restrict__ seeds,
               unsigned long long *__restrict__ state, unsigned long long *__restrict__ counters)
{
    const int64_t n = rows * cols;
    const int64_t p = (int64_t)blockIdx.x * FA_THREADS + threadIdx.x;
    int valid = 0;
    if (p < n) {
        const unsigned code = d8[p];
        if (code != 0) {
            valid = 1;
            const int64_t r = p / cols, c = p - r * cols;
            int pending = 0;
            // neighbour at (r+dr, c+dc) points at p iff its code is the opposite direction
            // E(1)<->W(16), SE(2)<->NW(32), S(4)<->N(64), SW(8)<->NE(128)
            const bool up = r > 0, dn = r + 1 < rows, lf = c > 0, rt = c + 1 < cols;
            if (up && lf) pending += d8[p - cols - 1] == 2;
            if (up) pending += d8[p - cols] == 4;
            if (up && rt) pending += d8[p - cols + 1] == 8;
            if (lf) pending += d8[p - 1] == 1;
            if (rt) pending += d8[p + 1] == 16;
            if (dn && lf) pending += d8[p + cols - 1] == 128;
            if (dn) pending += d8[p + cols] == 64;
            if (dn && rt) pending += d8[p + cols + 1] == 32;
            uint64_t w = seeds ? (uint64_t)seeds[p] & CNT_MASK : 0ull;
            w |= (uint64_t)pending << 44;
            if (pending == 0) w |= SRC_FLAG;
            state[p] = w;
        } else {
            state[p] = 0ull;
        }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, valid);
    if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(&counters[0], (unsigned long long)__popc(ballot));
}

template <typename ACC>
__global__ void __launch_bounds__(FA_THREADS)
fa_sweep_kernel(const uint8_t *__restrict__ d8, int64_t rows, int64_t cols, ACC *__restrict__ acc, ACC nodata_fill,
                unsigned long long *__restrict__ state, unsigned long long *__restrict__ counters)
{
    const int64_t n = rows * cols;
    int64_t p = (int64_t)blockIdx.x * FA_THREADS + threadIdx.x;
    unsigned finalised = 0;
    if (p < n) {
        unsigned code = d8[p];
        if (code == 0) {
            acc[p] = nodata_fill;
        } else {
            const uint64_t w = state[p];
            if (w & SRC_FLAG) {
                uint64_t carry = w & CNT_MASK;
                acc[p] = (ACC)carry;
                ++finalised;
                int64_t r = p / cols, c = p - r * cols;
                for (;;) {
                    int dr, dc;
                    if (!d8_offset(code, dr, dc)) break;
                    r += dr;
                    c += dc;
                    if (r < 0 || r >= rows || c < 0 || c >= cols) break;
                    p = r * cols + c;
                    code = d8[p];
                    if (code == 0) break;
                    const uint64_t old = atomicAdd(&state[p], (unsigned long long)((carry + 1ull) - PEND_ONE));
                    if (((old >> 44) & 0xFull) != 1ull) break;  // other tributaries still pending
                    carry = (old & CNT_MASK) + carry + 1ull;
                    acc[p] = (ACC)carry;
                    ++finalised;
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) finalised += __shfl_xor_sync(0xffffffffu, finalised, o);
    if ((threadIdx.x & 31) == 0 && finalised) atomicAdd(&counters[1], (unsigned long long)finalised);
}

template <typename ACC>
__global__ void __launch_bounds__(FA_THREADS)
fa_fix_kernel(const uint8_t *__restrict__ d8, int64_t n, ACC *__restrict__ acc, const unsigned long long *__restrict__ state,
              const unsigned long long *__restrict__ counters)
{
    if (counters[0] == counters[1]) return;  // no cycles: nothing to do
    const int64_t p = (int64_t)blockIdx.x * FA_THREADS + threadIdx.x;
    if (p >= n || d8[p] == 0) return;
    const uint64_t w = state[p];
    if ((w >> 44) & 0xFull) acc[p] = (ACC)(w & CNT_MASK);  // never finalised: partial count
}

}  // namespace
}  // namespace dtb

extern "C" size_t dtb_flowacc_workspace_bytes(int64_t rows, int64_t cols)
{
    if (rows <= 0 || cols <= 0) return 0;
    return (size_t)rows * (size_t)cols * 8 + 256;
}

extern "C" int dtb_flowacc(const uint8_t *d8, int64_t rows, int64_t cols, void *acc, int acc_dtype, int64_t nodata_fill,
                           const int64_t *seeds, void *ws, size_t ws_bytes, int64_t *unfinalised_host, void *stream)
{
    using namespace dtb;
    if (!d8 || !acc || !ws || rows <= 0 || cols <= 0) return DTB_ERR_INVALID;
    if (acc_dtype != DTB_I32 && acc_dtype != DTB_I64) return DTB_ERR_INVALID;
    if (ws_bytes < dtb_flowacc_workspace_bytes(rows, cols)) return DTB_ERR_WORKSPACE;
    const int64_t n = rows * cols;
    if (acc_dtype == DTB_I32 && n > 0x7fffffffLL) return DTB_ERR_UNSUPPORTED;
    if (n >= (int64_t)1 << 43) return DTB_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    unsigned long long *counters = reinterpret_cast<unsigned long long *>(ws);
    unsigned long long *state = reinterpret_cast<unsigned long long *>((char *)ws + 256);
    const unsigned blocks = (unsigned)((n + FA_THREADS - 1) / FA_THREADS);
    DTB_CUDA(cudaMemsetAsync(counters, 0, 256, st));
    fa_init_kernel<<<blocks, FA_THREADS, 0, st>>>(d8, rows, cols, seeds, state, counters);
    DTB_LAUNCH_CHECK("fa_init_kernel");
    if (acc_dtype == DTB_I32) {
        fa_sweep_kernel<int32_t><<<blocks, FA_THREADS, 0, st>>>(d8, rows, cols, (int32_t *)acc, (int32_t)nodata_fill, state, counters);
        DTB_LAUNCH_CHECK("fa_sweep_kernel<i32>");
        fa_fix_kernel<int32_t><<<blocks, FA_THREADS, 0, st>>>(d8, n, (int32_t *)acc, state, counters);
        DTB_LAUNCH_CHECK("fa_fix_kernel<i32>");
    } else {
        fa_sweep_kernel<int64_t><<<blocks, FA_THREADS, 0, st>>>(d8, rows, cols, (int64_t *)acc, (int64_t)nodata_fill, state, counters);
        DTB_LAUNCH_CHECK("fa_sweep_kernel<i64>");
        fa_fix_kernel<int64_t><<<blocks, FA_THREADS, 0, st>>>(d8, n, (int64_t *)acc, state, counters);
        DTB_LAUNCH_CHECK("fa_fix_kernel<i64>");
    }
    if (unfinalised_host) {
        unsigned long long h[2];
        DTB_CUDA(cudaMemcpyAsync(h, counters, sizeof(h), cudaMemcpyDeviceToHost, st));
        DTB_CUDA(cudaStreamSynchronize(st));
        *unfinalised_host = (int64_t)(h[0] - h[1]);
    }
    return DTB_OK;
}
