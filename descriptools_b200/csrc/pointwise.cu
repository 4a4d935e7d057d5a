// pointwise.cu -- per-cell indices: river accumulation gather, GFI, ln(hl/H), TI + MTI.
//
// Semantics (all arithmetic in f64 as compiled Numba types it, result stored f32):
//   river accumulation  gfi.py:136-147
//   GFI                 gfi.py:287-294    log(b * pow(A_r*size^2, n) / (H + 0.01)), H <= -100 -> -100
//   ln(hl/H)            gfi.py:427-440    same with the cell's own A, A == 0 -> 1
//   TI / MTI            topoindexes.py:250-261, 284-295   log(A*px^2 / tan(beta + 0.01)),
//                       log(pow(A*px^2, n) / tan(beta + 0.01)); A <= -100 -> -100, A == 0 -> 1
// The reference launches TI and MTI back to back on identical inputs
// (topoindexes.py:218-222); here they share one pass (one read of A and beta, one tan()).
// Grid-stride loops sized to the SM count; inputs are streamed once, outputs written once.
#include <math.h>

#include "common.cuh"

namespace dtb {
namespace {

constexpr int PW_THREADS = 256;

inline unsigned pw_blocks(int64_t n)
{
    const int64_t want = (n + PW_THREADS - 1) / PW_THREADS;
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (unsigned)(want < cap ? want : cap);
}

template <typename ACC, typename IDX>
__global__ void __launch_bounds__(PW_THREADS)
river_acc_kernel(const ACC *__restrict__ acc, const IDX *__restrict__ idx, int64_t n, ACC *__restrict__ out, int32_t *__restrict__ oob)
{
    for (int64_t i = (int64_t)blockIdx.x * PW_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * PW_THREADS) {
        int64_t j = (int64_t)idx[i];
        if (j == ND_I) j = 0;       // gfi.py:141-143: idx == -100 reads fac.flat[0]
        else if (j < 0) j += n;     // NumPy / Numba wrap negative flat indices
        if (j < 0 || j >= n) {      // an IndexError in the reference: never dereferenced here
            if (oob) *oob = 1;
            j = 0;
        }
        out[i] = acc[j];
    }
}

// Four consecutive cells per thread and pass: one 8 / 16 / 32-byte load per input, one 16-byte store per output.
template <typename T> struct alignas(sizeof(T) * 4) Pack4 { T v[4]; };
template <typename T> __device__ __forceinline__ Pack4<T> ld4(const T *p) { return *reinterpret_cast<const Pack4<T> *>(p); }
__device__ __forceinline__ void st4(float *p, const float (&v)[4]) { *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]); }
inline bool aligned4(const void *p, size_t elem) { return (reinterpret_cast<uintptr_t>(p) % (4 * elem)) == 0; }

// log(scale * pow(x, expo) / y) = ln scale + expo ln x - ln y: two short logarithms (common.cuh fast_log_pos, |error| < 3e-9)
// instead of pow + log + a division in f64 (~450 -> ~70 instructions per cell; the result is rounded to f32 and compared
// at 1e-5 relative).  Only where every term is an ordinary positive number and the power cannot leave the f64 range;
// everything else (zero / negative / non-finite arguments, extreme exponents) takes the reference's expression as it
// stands, so infinities and NaNs come out as they do there.
struct LogForm {
    double expo, scale, ln_scale;
    int fast;  // scale is positive and finite, |expo| <= 8
};
inline LogForm log_form(double expo, double scale)
{
    LogForm f;
    f.expo = expo;
    f.scale = scale;
    f.fast = (scale > 1e-30 && scale < 1e30 && expo >= -8.0 && expo <= 8.0) ? 1 : 0;
    f.ln_scale = f.fast ? log(scale) : 0.0;
    return f;
}
__device__ __forceinline__ bool ordinary(double x) { return x > 1e-30 && x < 1e30; }
__device__ __forceinline__ float log_ratio(const LogForm &f, double x, double y)
{
    if (f.fast && ordinary(x) && ordinary(y)) return (float)(f.ln_scale + f.expo * fast_log_pos(x) - fast_log_pos(y));
    return (float)log(f.scale * pow(x, f.expo) / y);
}

template <typename H, typename ACC, bool OWN>
__device__ __forceinline__ float gfi_cell(H h, ACC a0, const LogForm &f, double s2)
{
    if (h <= (H)ND_I) return ND_F;  // (a NaN height falls through, as in the reference)
    double a = (double)a0;
    if (OWN && a0 == 0) a = 1.0;  // gfi.py:432-435
    return log_ratio(f, a * s2, (double)h + 0.01);
}

template <typename H, typename ACC, bool OWN, bool VEC>
__global__ void __launch_bounds__(PW_THREADS)
gfi_kernel(const H *__restrict__ hand, const ACC *__restrict__ acc, int64_t n, LogForm f, double s2, float *__restrict__ out)
{
    const int64_t t0 = (int64_t)blockIdx.x * PW_THREADS + threadIdx.x, nt = (int64_t)gridDim.x * PW_THREADS;
    const int64_t nv = VEC ? n / 4 : 0;
    for (int64_t q = t0; q < nv; q += nt) {
        const Pack4<H> h = ld4(hand + 4 * q);
        const Pack4<ACC> a = ld4(acc + 4 * q);
        float r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] = gfi_cell<H, ACC, OWN>(h.v[k], a.v[k], f, s2);
        st4(out + 4 * q, r);
    }
    for (int64_t i = 4 * nv + t0; i < n; i += nt) out[i] = gfi_cell<H, ACC, OWN>(hand[i], acc[i], f, s2);
}

template <typename ACC>
__device__ __forceinline__ void ti_mti_cell(ACC a0, float beta, const LogForm &f, double p2, bool want_ti, bool want_mti, float &t1, float &t2)
{
    t1 = ND_F;
    t2 = ND_F;
    if (a0 <= (ACC)ND_I) return;  // topoindexes.py:252, 286
    const double x = ((a0 == 0) ? 1.0 : (double)a0) * p2;
    const double t = tan((double)beta + 0.01);
    if (f.fast && ordinary(x) && ordinary(t)) {
        const double lx = fast_log_pos(x), lt = fast_log_pos(t);
        t1 = (float)(lx - lt);
        t2 = (float)(f.expo * lx - lt);
    } else {
        if (want_ti) t1 = (float)log(x / t);
        if (want_mti) t2 = (float)log(pow(x, f.expo) / t);
    }
}

template <typename ACC, bool VEC>
__global__ void __launch_bounds__(PW_THREADS)
ti_mti_kernel(const ACC *__restrict__ acc, const float *__restrict__ beta, int64_t n, LogForm f, double p2,
              float *__restrict__ ti, float *__restrict__ mti)
{
    const int64_t t0 = (int64_t)blockIdx.x * PW_THREADS + threadIdx.x, nt = (int64_t)gridDim.x * PW_THREADS;
    const int64_t nv = VEC ? n / 4 : 0;
    for (int64_t q = t0; q < nv; q += nt) {
        const Pack4<ACC> a = ld4(acc + 4 * q);
        const Pack4<float> b = ld4(beta + 4 * q);
        float r1[4], r2[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) ti_mti_cell<ACC>(a.v[k], b.v[k], f, p2, ti != nullptr, mti != nullptr, r1[k], r2[k]);
        if (ti) st4(ti + 4 * q, r1);
        if (mti) st4(mti + 4 * q, r2);
    }
    for (int64_t i = 4 * nv + t0; i < n; i += nt) {
        float t1, t2;
        ti_mti_cell<ACC>(acc[i], beta[i], f, p2, ti != nullptr, mti != nullptr, t1, t2);
        if (ti) ti[i] = t1;
        if (mti) mti[i] = t2;
    }
}

__device__ __forceinline__ float slope_rad_cell(float s) { return (s == ND_F) ? ND_F : atanf(s / 100.0f); }  // example.py:63-64

template <bool VEC>
__global__ void __launch_bounds__(PW_THREADS)
slope_rad_kernel(const float *__restrict__ pct, int64_t n, float *__restrict__ rad)
{
    const int64_t t0 = (int64_t)blockIdx.x * PW_THREADS + threadIdx.x, nt = (int64_t)gridDim.x * PW_THREADS;
    const int64_t nv = VEC ? n / 4 : 0;
    for (int64_t q = t0; q < nv; q += nt) {
        const Pack4<float> s = ld4(pct + 4 * q);
        float r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] = slope_rad_cell(s.v[k]);
        st4(rad + 4 * q, r);
    }
    for (int64_t i = 4 * nv + t0; i < n; i += nt) rad[i] = slope_rad_cell(pct[i]);
}

// example.py:42-43 on the device, in one pass: the file's nodata value (and NaN) becomes the path's sentinel -100
__global__ void __launch_bounds__(PW_THREADS)
nodata_sentinel_kernel(float *__restrict__ dem, int64_t n, float nodata, int has_nodata)
{
    for (int64_t i = (int64_t)blockIdx.x * PW_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * PW_THREADS) {
        const float z = dem[i];
        if (z != z || (has_nodata && z == nodata)) dem[i] = ND_F;
    }
}

}  // namespace
}  // namespace dtb

using namespace dtb;

extern "C" int dtb_nodata_to_sentinel_f32(float *dem, int64_t n, float nodata, int has_nodata, void *stream)
{
    if (!dem || n < 0) return DTB_ERR_INVALID;
    if (n == 0) return DTB_OK;
    cudaStream_t st = as_stream(stream);
    DTB_KERNEL("nodata_sentinel_kernel", st, (nodata_sentinel_kernel<<<pw_blocks(n), PW_THREADS, 0, st>>>(dem, n, nodata, has_nodata)));
    return DTB_OK;
}

extern "C" int dtb_river_accumulation(const void *acc, int acc_dtype, const void *idx, int idx_dtype, int64_t n, void *out,
                                      int32_t *oob, void *stream)
{
    if (!acc || !idx || !out || n < 0) return DTB_ERR_INVALID;
    if (n == 0) return DTB_OK;
    cudaStream_t st = as_stream(stream);
    const unsigned b = pw_blocks(n);
    if (acc_dtype == DTB_I64 && idx_dtype == DTB_I64)
        river_acc_kernel<int64_t, int64_t><<<b, PW_THREADS, 0, st>>>((const int64_t *)acc, (const int64_t *)idx, n, (int64_t *)out, oob);
    else if (acc_dtype == DTB_I64 && idx_dtype == DTB_I32)
        river_acc_kernel<int64_t, int32_t><<<b, PW_THREADS, 0, st>>>((const int64_t *)acc, (const int32_t *)idx, n, (int64_t *)out, oob);
    else if (acc_dtype == DTB_I32 && idx_dtype == DTB_I64)
        river_acc_kernel<int32_t, int64_t><<<b, PW_THREADS, 0, st>>>((const int32_t *)acc, (const int64_t *)idx, n, (int32_t *)out, oob);
    else if (acc_dtype == DTB_I32 && idx_dtype == DTB_I32)
        river_acc_kernel<int32_t, int32_t><<<b, PW_THREADS, 0, st>>>((const int32_t *)acc, (const int32_t *)idx, n, (int32_t *)out, oob);
    else
        return DTB_ERR_INVALID;
    DTB_LAUNCH_CHECK("river_acc_kernel");
    return DTB_OK;
}

template <bool OWN>
static int launch_gfi(const void *hand, int hand_dtype, const void *acc, int acc_dtype, int64_t n, double expo, double scale,
                      double size, float *out, void *stream)
{
    if (!hand || !acc || !out || n < 0) return DTB_ERR_INVALID;
    if (n == 0) return DTB_OK;
    cudaStream_t st = as_stream(stream);
    const LogForm f = log_form(expo, scale);
    const double s2 = size * size;
    const size_t hs = hand_dtype == DTB_F32 ? 4 : 2, as = acc_dtype == DTB_I64 ? 8 : 4;
    const bool vec = aligned4(hand, hs) && aligned4(acc, as) && aligned4(out, 4);
    const unsigned b = pw_blocks(vec ? (n + 3) / 4 : n);
#define DTB_GFI(H, A)                                                                                                  \
    do {                                                                                                               \
        if (vec) gfi_kernel<H, A, OWN, true><<<b, PW_THREADS, 0, st>>>((const H *)hand, (const A *)acc, n, f, s2, out); \
        else gfi_kernel<H, A, OWN, false><<<b, PW_THREADS, 0, st>>>((const H *)hand, (const A *)acc, n, f, s2, out);    \
    } while (0)
    if (hand_dtype == DTB_F32 && acc_dtype == DTB_I64) DTB_GFI(float, int64_t);
    else if (hand_dtype == DTB_F32 && acc_dtype == DTB_I32) DTB_GFI(float, int32_t);
    else if (hand_dtype == DTB_I16 && acc_dtype == DTB_I64) DTB_GFI(int16_t, int64_t);
    else if (hand_dtype == DTB_I16 && acc_dtype == DTB_I32) DTB_GFI(int16_t, int32_t);
    else return DTB_ERR_INVALID;
#undef DTB_GFI
    DTB_LAUNCH_CHECK(OWN ? "lnhlh_kernel" : "gfi_kernel");
    return DTB_OK;
}

extern "C" int dtb_gfi(const void *hand, int hand_dtype, const void *racc, int acc_dtype, int64_t n, double expo, double scale,
                       double size, float *out, void *stream)
{
    return launch_gfi<false>(hand, hand_dtype, racc, acc_dtype, n, expo, scale, size, out, stream);
}

extern "C" int dtb_lnhlh(const void *hand, int hand_dtype, const void *acc, int acc_dtype, int64_t n, double expo, double scale,
                         double size, float *out, void *stream)
{
    return launch_gfi<true>(hand, hand_dtype, acc, acc_dtype, n, expo, scale, size, out, stream);
}

extern "C" int dtb_ti_mti(const void *acc, int acc_dtype, const float *slope_rad, int64_t n, double px, double expo, float *ti,
                          float *mti, void *stream)
{
    if (!acc || !slope_rad || (!ti && !mti) || n < 0) return DTB_ERR_INVALID;
    if (n == 0) return DTB_OK;
    cudaStream_t st = as_stream(stream);
    const LogForm f = log_form(expo, 1.0);
    const size_t as = acc_dtype == DTB_I64 ? 8 : 4;
    const bool vec = aligned4(acc, as) && aligned4(slope_rad, 4) && (!ti || aligned4(ti, 4)) && (!mti || aligned4(mti, 4));
    const unsigned b = pw_blocks(vec ? (n + 3) / 4 : n);
#define DTB_TI(A)                                                                                                               \
    do {                                                                                                                        \
        if (vec) ti_mti_kernel<A, true><<<b, PW_THREADS, 0, st>>>((const A *)acc, slope_rad, n, f, px * px, ti, mti);            \
        else ti_mti_kernel<A, false><<<b, PW_THREADS, 0, st>>>((const A *)acc, slope_rad, n, f, px * px, ti, mti);               \
    } while (0)
    if (acc_dtype == DTB_I64) DTB_TI(int64_t);
    else if (acc_dtype == DTB_I32) DTB_TI(int32_t);
    else return DTB_ERR_INVALID;
#undef DTB_TI
    DTB_LAUNCH_CHECK("ti_mti_kernel");
    return DTB_OK;
}

extern "C" int dtb_slope_to_radians(const float *slope_pct, int64_t n, float *slope_rad, void *stream)
{
    if (!slope_pct || !slope_rad || n < 0) return DTB_ERR_INVALID;
    if (n == 0) return DTB_OK;
    cudaStream_t st = as_stream(stream);
    const bool vec = aligned4(slope_pct, 4) && aligned4(slope_rad, 4);
    if (vec) slope_rad_kernel<true><<<pw_blocks((n + 3) / 4), PW_THREADS, 0, st>>>(slope_pct, n, slope_rad);
    else slope_rad_kernel<false><<<pw_blocks(n), PW_THREADS, 0, st>>>(slope_pct, n, slope_rad);
    DTB_LAUNCH_CHECK("slope_rad_kernel");
    return DTB_OK;
}
