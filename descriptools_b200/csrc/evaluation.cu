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
// before the first one that does (under: first i with d <= th[i]; over: thresholds th[0..j-1] satisfy d >= th, j = count)
template <typename T>
__global__ void __launch_bounds__(EV_THREADS)
eval_hist_kernel(const T *__restrict__ desc, const int8_t *__restrict__ flood, int64_t n, double nodata, Thresholds t, int under,
                 unsigned long long *__restrict__ hist)
{
    __shared__ unsigned sh[4 * (EV_MAXK + 1)];
    for (int i = threadIdx.x; i < 4 * (EV_MAXK + 1); i += EV_THREADS) sh[i] = 0;
    __syncthreads();
    for (int64_t p = (int64_t)blockIdx.x * EV_THREADS + threadIdx.x; p < n; p += (int64_t)gridDim.x * EV_THREADS) {
        const int c = remap((int)flood[p]);
        if (c < 0 || c > 3) continue;
        const T d = desc[p];
        int j;
        if (!usable(d, nodata)) j = under ? t.k : 0;  // never flagged
        else if (under) { j = 0; while (j < t.k && !((double)d <= t.th[j])) ++j; }   // flagged by thresholds j..k-1
        else { j = 0; while (j < t.k && (double)d >= t.th[j]) ++j; }                  // flagged by thresholds 0..j-1
        atomicAdd(&sh[c * (EV_MAXK + 1) + j], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * (EV_MAXK + 1); i += EV_THREADS)
        if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
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
    if (desc_is_f64) DTB_KERNEL("eval_hist_kernel<f64>", st, eval_hist_kernel<double><<<EV_BLOCKS, EV_THREADS, 0, st>>>((const double *)desc, flood, n, nodata, t, under, hist));
    else DTB_KERNEL("eval_hist_kernel<f32>", st, eval_hist_kernel<float><<<EV_BLOCKS, EV_THREADS, 0, st>>>((const float *)desc, flood, n, nodata, t, under, hist));
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
