// evaluation.cu -- threshold calibration of a descriptor against a benchmark flood map
// (evaluation.py:5-9 minMaxScale, :12-87 calibration, :90-123 binary_map, :126-171 avaliacao).
//
// The reference evaluates one threshold per full-raster NumPy pass (np.where + np.unique, 61 passes per
// calibration).  Here one pass serves a whole stage of the search: every cell finds the first threshold of the
// ascending stage list that classifies it as flooded, a per-benchmark-class histogram over that position is
// accumulated in shared memory, and prefix sums on the host give the confusion counts of every threshold.
#include "common.cuh"

namespace dtb {
namespace {

constexpr int EV_MAXK = 32;
constexpr int EV_THREADS = 256;
constexpr int EV_BLOCKS = kNumSMs * 8;

struct Thresholds {
    double th[EV_MAXK];
    int k;
};

// binary_map (evaluation.py:90-123): cells equal to `nodata` (= descriptor_matrix[0,0]) and NaNs are never flooded
template <typename T>
__device__ __forceinline__ bool usable(T d, double nodata) { return !((double)d != (double)d) && !((double)d == nodata); }

// avaliacao's remapping of the benchmark (evaluation.py:149-150): 1 -> 2, -100 -> 0
__device__ __forceinline__ int remap(int f) { return f == 1 ? 2 : (f == ND_I ? 0 : f); }

// hist[c][j], c = remapped benchmark value 0..3, j = 0..k: number of thresholds of the list that do NOT flag the cell
// before the first one that does (under: first i with d <= th[i], i.e. j = #{th < d}; over: thresholds th[0..j-1] satisfy
// d >= th, i.e. j = #{th <= d}).
// Four cells per thread and pass (one 16 / 32-byte load of the descriptor, one 4-byte load of the benchmark); j by a
// branch-free binary search over the threshold list in shared memory (padded to 32 entries with +inf); the lanes of a warp
// that hit the same bin as the first of them are summed and added by it (eval_count).  CT = float when the descriptor is f32 and every threshold (and the nodata value)
// is a float exactly: the comparison is the same, without a conversion to f64 per cell.
template <typename T> struct alignas(sizeof(T) * 4) EvPack4 { T v[4]; };

// bin of one cell, or -1 for a benchmark value outside {0, 1, 2, 3, -100}.  Branch-free: the class slot is the low three
// bits of the benchmark byte (-100 = 0x9c -> 4; slots become avaliacao's classes when the block adds its counts to the global histogram);
// a search step is one shared load at [offset + constant], one compare and one predicated add of the byte offset.
// `nodata` arrives as NaN when it cannot equal a CT descriptor.
template <typename T, typename CT, bool UNDER>
__device__ __forceinline__ int eval_bin(T d0, int f, CT nodata, const CT *__restrict__ th, int k)
{
    const CT d = (CT)d0;
    unsigned off = 0;
    const char *base = reinterpret_cast<const char *>(th);
#pragma unroll
    for (int step = EV_MAXK / 2; step >= 1; step >>= 1) {
        const CT t = *reinterpret_cast<const CT *>(base + off + (step - 1) * sizeof(CT));
        if (UNDER ? t < d : t <= d) off += step * (unsigned)sizeof(CT);
    }
    const CT t = *reinterpret_cast<const CT *>(base + off);
    if (UNDER ? t < d : t <= d) off += (unsigned)sizeof(CT);
    int j = min((int)(off / sizeof(CT)), k);  // (d = +inf counts the +inf padding as well)
    if (d != d || d == nodata) j = UNDER ? k : 0;  // never flagged
    const bool valid = (unsigned)f <= 3u || f == ND_I;
    return valid ? (f & 7) * (EV_MAXK + 1) + j : -1;
}

// One histogram update for the four cells of every lane: the cells that fall into the bin of the first lane's first valid
// cell -- in a raster nearly all of them -- are counted with one redux and added by one lane; the others add on their own
// (same-address shared atomics serialise, so per-cell atomics cost 32 passes per warp and cell).
__device__ __forceinline__ void eval_count4(unsigned *sh, const int (&key)[4])
{
    const unsigned act = __activemask();
    int first = key[0];
#pragma unroll
    for (int i = 1; i < 4; ++i) first = first >= 0 ? first : key[i];
    const unsigned want = __ballot_sync(act, first >= 0);
    if (!want) return;
    const int lead = __ffs((int)want) - 1;
    const int cand = __shfl_sync(act, first, lead);
    unsigned hits = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) hits += key[i] == cand ? 1u : 0u;
    const unsigned sum = __reduce_add_sync(act, hits);
    if ((int)(threadIdx.x & 31) == lead) atomicAdd(&sh[cand], sum);
    if (hits != 4u) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (key[i] >= 0 && key[i] != cand) atomicAdd(&sh[key[i]], 1u);
    }
}

constexpr int EV_SLOTS = 5;  // benchmark bytes 0, 1, 2, 3 and -100 (& 7 = 4)

template <typename T, typename CT, bool VEC, bool UNDER>
__global__ void __launch_bounds__(EV_THREADS)
eval_hist_kernel(const T *__restrict__ desc, const int8_t *__restrict__ flood, int64_t n, double nodata, Thresholds t,
                 unsigned long long *__restrict__ hist)
{
    __shared__ unsigned sh[EV_SLOTS * (EV_MAXK + 1)];
    __shared__ CT th[EV_MAXK + 1];
    for (int i = threadIdx.x; i < EV_SLOTS * (EV_MAXK + 1); i += EV_THREADS) sh[i] = 0;
    if (threadIdx.x <= EV_MAXK) {
        const int i = threadIdx.x;
        th[i] = i < t.k ? (CT)t.th[i] : (CT)__longlong_as_double(0x7ff0000000000000LL);  // +inf: never below a value
    }
    __syncthreads();
    // (a nodata value that is not a CT never equals a CT descriptor: compare with NaN then)
    const CT nd = (double)(CT)nodata == nodata ? (CT)nodata : (CT)__longlong_as_double(0x7ff8000000000000LL);
    const int64_t t0 = (int64_t)blockIdx.x * EV_THREADS + threadIdx.x, nt = (int64_t)gridDim.x * EV_THREADS;
    const int64_t nv = VEC ? n / 4 : 0;
    for (int64_t q = t0; q < nv; q += nt) {
        const EvPack4<T> d = *reinterpret_cast<const EvPack4<T> *>(desc + 4 * q);
        const char4 f = *reinterpret_cast<const char4 *>(flood + 4 * q);
        const int fv[4] = {f.x, f.y, f.z, f.w};
        int key[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) key[i] = eval_bin<T, CT, UNDER>(d.v[i], fv[i], nd, th, t.k);
        eval_count4(sh, key);
    }
    for (int64_t p = 4 * nv + t0; p < n; p += nt) {
        const int key = eval_bin<T, CT, UNDER>(desc[p], (int)flood[p], nd, th, t.k);
        if (key >= 0) atomicAdd(&sh[key], 1u);
    }
    __syncthreads();
    // slots -> avaliacao's classes (evaluation.py:149-150): 1 -> 2, -100 -> 0
    for (int i = threadIdx.x; i < EV_SLOTS * (EV_MAXK + 1); i += EV_THREADS) {
        if (!sh[i]) continue;
        const int slot = i / (EV_MAXK + 1), j = i - slot * (EV_MAXK + 1);
        const int c = slot == 1 ? 2 : (slot == 4 ? 0 : slot);
        atomicAdd(&hist[c * (EV_MAXK + 1) + j], (unsigned long long)sh[i]);
    }
}

template <typename T>
__global__ void __launch_bounds__(EV_THREADS)
eval_class_kernel(const T *__restrict__ desc, const int8_t *__restrict__ flood, int64_t n, double nodata, double th, int under,
                  int8_t *__restrict__ binary, int8_t *__restrict__ cls)
{
    for (int64_t p = (int64_t)blockIdx.x * EV_THREADS + threadIdx.x; p < n; p += (int64_t)gridDim.x * EV_THREADS) {
        const T d = desc[p];
        const int b = usable(d, nodata) && (under ? (double)d <= th : (double)d >= th);
        if (binary) binary[p] = (int8_t)b;
        if (cls) cls[p] = (int8_t)(b + remap((int)flood[p]));
    }
}

template <typename T>
__global__ void __launch_bounds__(EV_THREADS)
minmax_scale_kernel(const T *__restrict__ mat, int64_t n, double mn, double mx, double nodata, double *__restrict__ out)
{
    const double q = mx - mn;
    for (int64_t p = (int64_t)blockIdx.x * EV_THREADS + threadIdx.x; p < n; p += (int64_t)gridDim.x * EV_THREADS) {
        const double v = (double)mat[p];
        // evaluation.py:6-7: nodata -> NaN; NaN stays NaN; everything else (v - mn) / (mx - mn)
        out[p] = (v == nodata || v != v) ? __longlong_as_double(0x7ff8000000000000LL) : (v - mn) / q;
    }
}

}  // namespace
}  // namespace dtb

extern "C" int dtb_eval_counts(const void *desc, int desc_is_f64, const int8_t *flood, int64_t n, double nodata,
                               const double *thresholds_host, int k, int under, int64_t *counts_host, void *ws, size_t ws_bytes,
                               void *stream)
{
    using namespace dtb;
    if (!desc || !flood || !thresholds_host || !counts_host || !ws || n <= 0 || k <= 0 || k > EV_MAXK) return DTB_ERR_INVALID;
    if (ws_bytes < 4 * (EV_MAXK + 1) * sizeof(unsigned long long)) return DTB_ERR_WORKSPACE;
    for (int i = 1; i < k; ++i)
        if (!(thresholds_host[i - 1] < thresholds_host[i])) return DTB_ERR_INVALID;  // strictly ascending
    cudaStream_t st = as_stream(stream);
    Thresholds t;
    t.k = k;
    for (int i = 0; i < EV_MAXK; ++i) t.th[i] = i < k ? thresholds_host[i] : 0.0;
    unsigned long long *hist = reinterpret_cast<unsigned long long *>(ws);
    DTB_CUDA(cudaMemsetAsync(hist, 0, 4 * (EV_MAXK + 1) * sizeof(unsigned long long), st));
    const bool vec = (reinterpret_cast<uintptr_t>(desc) % (desc_is_f64 ? 32 : 16)) == 0 && (reinterpret_cast<uintptr_t>(flood) % 4) == 0;
    bool as_float = !desc_is_f64 && (double)(float)nodata == nodata;
    for (int i = 0; i < k && as_float; ++i) as_float = (double)(float)thresholds_host[i] == thresholds_host[i];
#define DTB_EVH2(NAME, T, CT, VEC, UNDER) \
    DTB_KERNEL(NAME, st, (eval_hist_kernel<T, CT, VEC, UNDER><<<EV_BLOCKS, EV_THREADS, 0, st>>>((const T *)desc, flood, n, nodata, t, hist)))
#define DTB_EVH(NAME, T, CT)                                   \
    do {                                                       \
        if (vec && under) DTB_EVH2(NAME, T, CT, true, true);   \
        else if (vec) DTB_EVH2(NAME, T, CT, true, false);      \
        else if (under) DTB_EVH2(NAME, T, CT, false, true);    \
        else DTB_EVH2(NAME, T, CT, false, false);              \
    } while (0)
    if (desc_is_f64) DTB_EVH("eval_hist_kernel<f64>", double, double);
    else if (as_float) DTB_EVH("eval_hist_kernel<f32>", float, float);
    else DTB_EVH("eval_hist_kernel<f32,f64>", float, double);
#undef DTB_EVH
#undef DTB_EVH2
    unsigned long long h[4 * (EV_MAXK + 1)];
    DTB_CUDA(cudaMemcpyAsync(h, hist, sizeof(h), cudaMemcpyDeviceToHost, st));
    DTB_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < k; ++i) {
        int64_t cnt[5] = {0, 0, 0, 0, 0};  // classes 0..4 (binary + remapped benchmark)
        for (int c = 0; c < 4; ++c) {
            int64_t flagged = 0, total = 0;
            for (int j = 0; j <= k; ++j) {
                total += (int64_t)h[c * (EV_MAXK + 1) + j];
                if (under ? j <= i : j > i) flagged += (int64_t)h[c * (EV_MAXK + 1) + j];
            }
            cnt[c] += total - flagged;
            cnt[c + 1] += flagged;
        }
        for (int c = 0; c < 4; ++c) counts_host[4 * i + c] = cnt[c];  // tn, fp, fn, tp (evaluation.py:152-166)
    }
    return DTB_OK;
}

extern "C" int dtb_eval_class_map(const void *desc, int desc_is_f64, const int8_t *flood, int64_t n, double nodata, double threshold,
                                  int under, int8_t *binary, int8_t *cls, void *stream)
{
    using namespace dtb;
    if (!desc || n <= 0 || (!binary && !cls) || (cls && !flood)) return DTB_ERR_INVALID;
    cudaStream_t st = as_stream(stream);
    if (desc_is_f64) DTB_KERNEL("eval_class_kernel<f64>", st, eval_class_kernel<double><<<EV_BLOCKS, EV_THREADS, 0, st>>>((const double *)desc, flood, n, nodata, threshold, under, binary, cls));
    else DTB_KERNEL("eval_class_kernel<f32>", st, eval_class_kernel<float><<<EV_BLOCKS, EV_THREADS, 0, st>>>((const float *)desc, flood, n, nodata, threshold, under, binary, cls));
    return DTB_OK;
}

extern "C" int dtb_minmax_scale(const void *mat, int mat_dtype, int64_t n, double mn, double mx, double nodata, double *out, void *stream)
{
    using namespace dtb;
    if (!mat || !out || n <= 0) return DTB_ERR_INVALID;
    cudaStream_t st = as_stream(stream);
    if (mat_dtype == DTB_EV_F64) DTB_KERNEL("minmax_scale_kernel<f64>", st, minmax_scale_kernel<double><<<EV_BLOCKS, EV_THREADS, 0, st>>>((const double *)mat, n, mn, mx, nodata, out));
    else if (mat_dtype == DTB_EV_F32) DTB_KERNEL("minmax_scale_kernel<f32>", st, minmax_scale_kernel<float><<<EV_BLOCKS, EV_THREADS, 0, st>>>((const float *)mat, n, mn, mx, nodata, out));
    else if (mat_dtype == DTB_EV_I16) DTB_KERNEL("minmax_scale_kernel<i16>", st, minmax_scale_kernel<int16_t><<<EV_BLOCKS, EV_THREADS, 0, st>>>((const int16_t *)mat, n, mn, mx, nodata, out));
    else return DTB_ERR_INVALID;
    return DTB_OK;
}
