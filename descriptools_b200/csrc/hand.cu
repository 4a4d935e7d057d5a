// hand.cu -- flow distance to the nearest drainage, river-cell index, HAND (+ fused GFI).
//
// Semantics: flow_distance_index_gpu (flowhand.py:599-846, unpartitioned: out[*] = 0) and
// hand_calculator (flowhand.py:431-442); optional fused river_accumulation + GFI
// (gfi.py:136-147, 287-294).  Restated in oracle/dt_oracle.c.
//
// The reference walks every cell down the D8 pointers (O(path length) dependent loads per
// cell, mean 214 on its example).  Here the D8 forest is resolved by pointer jumping over a
// packed 64-bit state per cell:
//     [63..62 kind | 61..47 diagonal moves | 46..32 cardinal moves | 31..0 target cell]
// kind: ACTIVE (target = cell reached so far), RIVER (target = river cell), FAIL.
// One round = every ACTIVE cell adopts its target's state and adds the move counts.  Rounds
// update in place (asynchronously): any state read is a valid description of that cell's
// path, and every launch at least doubles the moves covered by each ACTIVE cell, so
// ceil(log2(max_moves+1)) launches decide every cell: whatever is still ACTIVE needs more
// than max_moves moves or sits on / drains into a cycle -> FAIL, exactly the reference's
// outcomes (flowhand.py:826 code-0 landing, :830 cycle detector, :835 move cap, border
// exits :623-764).  A launch whose predecessor left nothing ACTIVE exits immediately.
// Distance = n_card*px + n_diag*(px*sqrt(2)) in f64 (the reference adds the same terms one
// move at a time, flowhand.py:803-824; the two agree far below f32 resolution).
#include <math.h>

#include "common.cuh"

namespace dtb {
namespace {

constexpr int H_THREADS = 256;
constexpr uint64_t KIND_ACTIVE = 0, KIND_RIVER = 1, KIND_FAIL = 2;
constexpr uint32_t CNT_SAT = 32767;

__device__ __forceinline__ uint64_t pack(uint64_t kind, uint32_t nd, uint32_t nc, uint32_t ptr)
{
    return (kind << 62) | ((uint64_t)nd << 47) | ((uint64_t)nc << 32) | (uint64_t)ptr;
}
__device__ __forceinline__ uint64_t kind_of(uint64_t s) { return s >> 62; }
__device__ __forceinline__ uint32_t nd_of(uint64_t s) { return (uint32_t)(s >> 47) & 0x7FFFu; }
__device__ __forceinline__ uint32_t nc_of(uint64_t s) { return (uint32_t)(s >> 32) & 0x7FFFu; }
__device__ __forceinline__ uint32_t ptr_of(uint64_t s) { return (uint32_t)s; }
__device__ __forceinline__ uint32_t sat_add(uint32_t a, uint32_t b) { return min(a + b, CNT_SAT); }

template <typename ACC>
__device__ __forceinline__ bool is_river(const int8_t *river, const ACC *acc, int64_t thr, int64_t p)
{
    return river ? (river[p] == 1) : ((int64_t)acc[p] > thr);  // flowhand.py:609 / example.py:52
}

template <typename ACC>
__global__ void __launch_bounds__(H_THREADS)
hand_init_kernel(const uint8_t *__restrict__ fdr, const int8_t *__restrict__ river, const ACC *__restrict__ acc,
                 int64_t thr, int64_t rows, int64_t cols, unsigned long long *__restrict__ state,
                 unsigned *__restrict__ active)
{
    const int64_t n = rows * cols;
    const int64_t p = (int64_t)blockIdx.x * H_THREADS + threadIdx.x;
    int is_active = 0;
    if (p < n) {
        const unsigned code = fdr[p];
        uint64_t s;
        if (code == 0) {
            s = pack(KIND_FAIL, 0, 0, 0);  // flowhand.py:601
        } else if (is_river<ACC>(river, acc, thr, p)) {
            s = pack(KIND_RIVER, 0, 0, (uint32_t)p);  // flowhand.py:609-612
        } else {
            int dr, dc;
            const int64_t r = p / cols, c = p - r * cols;
            if (!d8_offset(code, dr, dc)) {
                s = pack(KIND_FAIL, 0, 0, 0);  // unknown code: "did not move", flowhand.py:830
            } else {
                const int64_t rr = r + dr, cc = c + dc;
                if (rr < 0 || rr >= rows || cc < 0 || cc >= cols) {
                    s = pack(KIND_FAIL, 0, 0, 0);  // leaves the raster, flowhand.py:623-764
                } else {
                    const int64_t q = rr * cols + cc;
                    if (fdr[q] == 0) s = pack(KIND_FAIL, 0, 0, 0);  // flowhand.py:826
                    else {
                        const bool diag = d8_is_diag(code);
                        s = pack(KIND_ACTIVE, diag ? 1u : 0u, diag ? 0u : 1u, (uint32_t)q);
                        is_active = 1;
                    }
                }
            }
        }
        state[p] = s;
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, is_active);
    if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(&active[0], (unsigned)__popc(ballot));
}

// round `rnd` (1-based): reads active[rnd-1], writes the number of cells still ACTIVE to active[rnd]
__global__ void __launch_bounds__(H_THREADS)
hand_jump_kernel(int64_t n, unsigned long long *state, unsigned *__restrict__ active, int rnd, int jumps)
{
    if (active[rnd - 1] == 0) return;
    const int64_t p = (int64_t)blockIdx.x * H_THREADS + threadIdx.x;
    int still = 0;
    if (p < n) {
        uint64_t s = __ldcg(&state[p]);
        if (kind_of(s) == KIND_ACTIVE) {
            for (int j = 0; j < jumps; ++j) {
                const uint64_t t = __ldcg(&state[ptr_of(s)]);
                const uint64_t kt = kind_of(t);
                if (kt == KIND_FAIL) { s = pack(KIND_FAIL, 0, 0, 0); break; }
                s = pack(kt, sat_add(nd_of(s), nd_of(t)), sat_add(nc_of(s), nc_of(t)), ptr_of(t));
                if (kt != KIND_ACTIVE) break;
            }
            __stcg(&state[p], s);
            still = kind_of(s) == KIND_ACTIVE;
        }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, still);
    if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(&active[rnd], (unsigned)__popc(ballot));
}

template <typename T> struct HandOps;
template <> struct HandOps<float> {
    static __device__ __forceinline__ bool is_nd(float z) { return z == ND_F; }
    static __device__ __forceinline__ float sub(float a, float b) { return a - b; }
    static __device__ __forceinline__ float nd() { return ND_F; }
};
template <> struct HandOps<int16_t> {
    static __device__ __forceinline__ bool is_nd(int16_t z) { return z == ND_I; }
    static __device__ __forceinline__ int16_t sub(int16_t a, int16_t b) { return (int16_t)(a - b); }  // numpy i16 wrap
    static __device__ __forceinline__ int16_t nd() { return (int16_t)ND_I; }
};

// flowhand.py:436-438
template <typename T>
__device__ __forceinline__ T hand_value(T z, bool resolved, const T *__restrict__ dem, int64_t idx)
{
    T h = HandOps<T>::nd();
    if (!HandOps<T>::is_nd(z) && resolved) h = HandOps<T>::sub(z, dem[idx]);
    if (h < (T)0 && h != HandOps<T>::nd()) h = (T)0;
    return h;
}

// gfi.py:287-294
template <typename T>
__device__ __forceinline__ float gfi_value(T h, double racc, double n, double b, double s2)
{
    if (h <= HandOps<T>::nd()) return ND_F;
    return (float)log(b * pow(racc * s2, n) / ((double)h + 0.01));
}

template <typename T, typename IDX, typename ACC>
__global__ void __launch_bounds__(H_THREADS)
hand_final_kernel(int64_t n, const unsigned long long *__restrict__ state, const T *__restrict__ dem,
                  const ACC *__restrict__ acc, double px, double pd, uint32_t max_moves, float *__restrict__ fdist,
                  IDX *__restrict__ idx_out, T *__restrict__ hand, float *__restrict__ gfi, double gn, double gb, double gs2)
{
    const int64_t p = (int64_t)blockIdx.x * H_THREADS + threadIdx.x;
    if (p >= n) return;
    const uint64_t s = state[p];
    const uint32_t nc = nc_of(s), nd = nd_of(s);
    const bool ok = kind_of(s) == KIND_RIVER && nc + nd <= max_moves;  // flowhand.py:835
    const int64_t idx = ok ? (int64_t)ptr_of(s) : (int64_t)ND_I;
    if (fdist) fdist[p] = ok ? (float)((double)nc * px + (double)nd * pd) : ND_F;  // flowhand.py:840-843
    if (idx_out) idx_out[p] = (IDX)idx;
    if (hand || gfi) {
        const T h = hand_value<T>(dem[p], ok, dem, idx);
        if (hand) hand[p] = h;
        if (gfi) {
            // river_accumulation: idx == -100 -> fac.flat[0] (gfi.py:141-143); irrelevant: H is -100 then
            const double racc = (double)acc[ok ? idx : 0];
            gfi[p] = gfi_value<T>(h, racc, gn, gb, gs2);
        }
    }
}

template <typename T, typename IDX>
__global__ void __launch_bounds__(H_THREADS)
hand_from_index_kernel(const T *__restrict__ dem, const IDX *__restrict__ idx, int64_t n, T *__restrict__ hand)
{
    const int64_t p = (int64_t)blockIdx.x * H_THREADS + threadIdx.x;
    if (p >= n) return;
    const int64_t i = (int64_t)idx[p];
    // numpy fancy indexing wraps negative indices; -100 is masked out anyway (flowhand.py:436)
    hand[p] = hand_value<T>(dem[p], i != ND_I, dem, i != ND_I ? i : 0);
}

inline int rounds_for(int64_t max_moves)
{
    int r = 0;
    while (((int64_t)1 << r) < max_moves + 1) ++r;
    return r;
}

template <typename T, typename IDX, typename ACC>
int run_final(const dtb_hand_args *a, const unsigned long long *state, uint32_t max_moves, cudaStream_t st)
{
    const int64_t n = a->rows * a->cols;
    const unsigned blocks = (unsigned)((n + H_THREADS - 1) / H_THREADS);
    hand_final_kernel<T, IDX, ACC><<<blocks, H_THREADS, 0, st>>>(
        n, state, (const T *)a->dem, (const ACC *)a->acc, a->px, a->px * sqrt(2.0), max_moves, a->fdist, (IDX *)a->idx,
        (T *)a->hand, a->gfi, a->gfi_n, a->gfi_b, a->gfi_size * a->gfi_size);
    DTB_LAUNCH_CHECK("hand_final_kernel");
    return DTB_OK;
}

template <typename T, typename IDX>
int run_final_acc(const dtb_hand_args *a, const unsigned long long *state, uint32_t mm, cudaStream_t st)
{
    return a->acc_dtype == DTB_I64 ? run_final<T, IDX, int64_t>(a, state, mm, st) : run_final<T, IDX, int32_t>(a, state, mm, st);
}

template <typename T>
int run_final_idx(const dtb_hand_args *a, const unsigned long long *state, uint32_t mm, cudaStream_t st)
{
    return a->idx_dtype == DTB_I64 ? run_final_acc<T, int64_t>(a, state, mm, st) : run_final_acc<T, int32_t>(a, state, mm, st);
}

}  // namespace
}  // namespace dtb

extern "C" size_t dtb_hand_workspace_bytes(int64_t rows, int64_t cols)
{
    if (rows <= 0 || cols <= 0) return 0;
    return (size_t)rows * (size_t)cols * 8 + 256;
}

extern "C" int dtb_hand(const dtb_hand_args *a, void *ws, size_t ws_bytes, void *stream)
{
    using namespace dtb;
    if (!a || !a->fdr || !ws || a->rows <= 0 || a->cols <= 0 || !(a->px > 0.0)) return DTB_ERR_INVALID;
    if (!a->river && !a->acc) return DTB_ERR_INVALID;
    if ((a->hand || a->gfi) && !a->dem) return DTB_ERR_INVALID;
    if (a->gfi && !a->acc) return DTB_ERR_INVALID;
    if (a->dem_dtype != DTB_F32 && a->dem_dtype != DTB_I16) return DTB_ERR_INVALID;
    if (ws_bytes < dtb_hand_workspace_bytes(a->rows, a->cols)) return DTB_ERR_WORKSPACE;
    const int64_t n = a->rows * a->cols;
    if (n > 0xffffffffLL) return DTB_ERR_UNSUPPORTED;  // 32-bit targets: shard larger rasters into bands
    if (a->idx && a->idx_dtype == DTB_I32 && n > 0x7fffffffLL) return DTB_ERR_UNSUPPORTED;
    const int64_t max_moves = a->max_moves > 0 ? a->max_moves : 20000;  // flowhand.py:835
    if (max_moves > 32000) return DTB_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    unsigned *active = reinterpret_cast<unsigned *>(ws);
    unsigned long long *state = reinterpret_cast<unsigned long long *>((char *)ws + 256);
    const unsigned blocks = (unsigned)((n + H_THREADS - 1) / H_THREADS);

    DTB_CUDA(cudaMemsetAsync(active, 0, 256, st));
    if (a->acc_dtype == DTB_I64)
        hand_init_kernel<int64_t><<<blocks, H_THREADS, 0, st>>>(a->fdr, a->river, (const int64_t *)a->acc, a->river_threshold,
                                                                a->rows, a->cols, state, active);
    else
        hand_init_kernel<int32_t><<<blocks, H_THREADS, 0, st>>>(a->fdr, a->river, (const int32_t *)a->acc, a->river_threshold,
                                                                a->rows, a->cols, state, active);
    DTB_LAUNCH_CHECK("hand_init_kernel");
    const int rounds = rounds_for(max_moves);
    for (int r = 1; r <= rounds; ++r) {
        hand_jump_kernel<<<blocks, H_THREADS, 0, st>>>(n, state, active, r, 2);
        DTB_LAUNCH_CHECK("hand_jump_kernel");
    }
    if (!a->fdist && !a->idx && !a->hand && !a->gfi) return DTB_OK;
    return a->dem_dtype == DTB_I16 ? run_final_idx<int16_t>(a, state, (uint32_t)max_moves, st)
                                   : run_final_idx<float>(a, state, (uint32_t)max_moves, st);
}

extern "C" int dtb_hand_from_index(const void *dem, int dem_dtype, const void *idx, int idx_dtype, int64_t n, void *hand,
                                   void *stream)
{
    using namespace dtb;
    if (!dem || !idx || !hand || n < 0) return DTB_ERR_INVALID;
    if (n == 0) return DTB_OK;
    cudaStream_t st = as_stream(stream);
    const unsigned blocks = (unsigned)((n + H_THREADS - 1) / H_THREADS);
    if (dem_dtype == DTB_F32 && idx_dtype == DTB_I64)
        hand_from_index_kernel<float, int64_t><<<blocks, H_THREADS, 0, st>>>((const float *)dem, (const int64_t *)idx, n, (float *)hand);
    else if (dem_dtype == DTB_F32 && idx_dtype == DTB_I32)
        hand_from_index_kernel<float, int32_t><<<blocks, H_THREADS, 0, st>>>((const float *)dem, (const int32_t *)idx, n, (float *)hand);
    else if (dem_dtype == DTB_I16 && idx_dtype == DTB_I64)
        hand_from_index_kernel<int16_t, int64_t><<<blocks, H_THREADS, 0, st>>>((const int16_t *)dem, (const int64_t *)idx, n, (int16_t *)hand);
    else if (dem_dtype == DTB_I16 && idx_dtype == DTB_I32)
        hand_from_index_kernel<int16_t, int32_t><<<blocks, H_THREADS, 0, st>>>((const int16_t *)dem, (const int32_t *)idx, n, (int16_t *)hand);
    else
        return DTB_ERR_INVALID;
    DTB_LAUNCH_CHECK("hand_from_index_kernel");
    return DTB_OK;
}
