// verify.cu -- size-independent identities of a finished chain (slope -> D8 -> accumulation -> HAND), evaluated on the
// device: what bench.py checks after its timed region at sizes the CPU oracle cannot reach (40 000 x 40 000 and the
// 100 000 x 100 000 band run), and what tests/ cross-checks against the oracle at small sizes.
//
// For a band of rows [row0, row0 + rows) of a raster with total_rows rows (a whole raster: row0 = 0, rows = total_rows):
//   out[0] = sum over roots of (acc + 1)      a root is a valid cell (code != 0) whose move leaves the raster, lands on a
//   out[1] = number of valid cells            cell without a code, or has no move; summed over all bands, out[0] == out[1]
//                                             (every valid cell is counted once at the root of its path)
//   out[2] = cells whose count is not the sum over their tributaries of (count + 1) -- the definition, checked
//            for every cell whose eight neighbours lie in this band or off the raster (by induction over the forest this
//            pins every count, given the codes)
//   out[3] = cells whose river index (inside this band) points at a cell that is not a river cell (acc <= threshold)
//   out[4] = river cells (acc > threshold) whose index is not their own, or whose HAND is not 0 (flowhand.py:609-612)
//   out[5] = resolved cells whose HAND is not max(z - z[idx], 0) (flowhand.py:436-438), index inside this band
//   out[6] = valid cells without a river index (paths that leave the raster before meeting a river cell)
//   out[7] = resolved cells whose index lies outside this band (not checked by 3 and 5)
#include "common.cuh"

namespace dtb {
namespace {

template <typename ACC, typename IDX>
__global__ void __launch_bounds__(256)
chain_check_kernel(const uint8_t *__restrict__ d8, const ACC *__restrict__ acc, const IDX *__restrict__ idx,
                   const float *__restrict__ dem, const float *__restrict__ hand, int64_t rows, int64_t cols, int64_t row0,
                   int64_t total_rows, int64_t thr, unsigned long long *__restrict__ out)
{
    unsigned long long c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int64_t n = rows * cols;
    for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < n; p += (int64_t)gridDim.x * 256) {
        const unsigned code = d8[p];
        if (code == 0) continue;
        ++c[1];
        const int64_t r = p / cols, col = p - r * cols;
        int dr, dc;
        const bool moves = d8_offset(code, dr, dc);
        const int64_t gr = row0 + r + dr, gc = col + dc;
        const bool inside = moves && gr >= 0 && gr < total_rows && gc >= 0 && gc < cols;
        const int64_t lr = r + dr;
        const bool local = inside && lr >= 0 && lr < rows;
        const int64_t a = (int64_t)acc[p];
        bool root = !inside;
        if (local && d8[lr * cols + gc] == 0) root = true;
        {
            // tributaries: neighbour k (NW,N,NE,W,E,SW,S,SE) flows here iff it carries the code pointing back
            const int DR[8] = {-1, -1, -1, 0, 0, 1, 1, 1}, DC[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
            const unsigned WANT[8] = {2, 4, 8, 1, 16, 128, 64, 32};
            int64_t s = 0;
            bool complete = true;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int64_t qr = r + DR[k], qc = col + DC[k];
                if (qc < 0 || qc >= cols || row0 + qr < 0 || row0 + qr >= total_rows) continue;  // off the raster
                if (qr < 0 || qr >= rows) { complete = false; continue; }                        // in another band
                if (d8[qr * cols + qc] == WANT[k]) s += (int64_t)acc[qr * cols + qc] + 1;
            }
            if (complete && s != a) ++c[2];
        }
        if (root) c[0] += (unsigned long long)(a + 1);
        if (idx) {
            const int64_t j = (int64_t)idx[p];
            const bool river = a > thr;
            const int64_t self = (row0 + r) * cols + col;
            if (river && (j != self || (hand && hand[p] != 0.0f))) ++c[4];
            if (j < 0) ++c[6];
            else {
                const int64_t jl = j - row0 * cols;
                if (jl < 0 || jl >= n) ++c[7];
                else {
                    if (!((int64_t)acc[jl] > thr)) ++c[3];
                    if (hand && dem) {
                        const float h = dem[p] - dem[jl];
                        if (hand[p] != (h < 0.0f ? 0.0f : h)) ++c[5];
                    }
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        unsigned long long v = c[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&out[k], v);
    }
}

}  // namespace
}  // namespace dtb

extern "C" int dtb_chain_check(const uint8_t *d8, const void *acc, int acc_dtype, const void *idx, int idx_dtype, const float *dem,
                               const float *hand, int64_t rows, int64_t cols, int64_t row0, int64_t total_rows,
                               int64_t river_threshold, unsigned long long *out, void *stream)
{
    using namespace dtb;
    if (!d8 || !acc || !out || rows <= 0 || cols <= 0 || row0 < 0 || row0 + rows > total_rows) return DTB_ERR_INVALID;
    cudaStream_t st = as_stream(stream);
    DTB_CUDA(cudaMemsetAsync(out, 0, 8 * sizeof(unsigned long long), st));
    const int64_t n = rows * cols;
    const int64_t want = (n + 255) / 256;
    const unsigned blocks = (unsigned)(want < (int64_t)kNumSMs * 16 ? want : (int64_t)kNumSMs * 16);
#define DTB_CC(A, I)                                                                                                               \
    DTB_KERNEL("chain_check_kernel", st,                                                                                          \
               (chain_check_kernel<A, I><<<blocks, 256, 0, st>>>(d8, (const A *)acc, (const I *)idx, dem, hand, rows, cols, row0, \
                                                                  total_rows, river_threshold, out)))
    if (acc_dtype == DTB_I32 && (idx_dtype == DTB_I32 || !idx)) DTB_CC(int32_t, int32_t);
    else if (acc_dtype == DTB_I32 && idx_dtype == DTB_I64) DTB_CC(int32_t, int64_t);
    else if (acc_dtype == DTB_I64 && (idx_dtype == DTB_I64 || !idx)) DTB_CC(int64_t, int64_t);
    else if (acc_dtype == DTB_I64 && idx_dtype == DTB_I32) DTB_CC(int64_t, int32_t);
    else return DTB_ERR_INVALID;
#undef DTB_CC
    return DTB_OK;
}
