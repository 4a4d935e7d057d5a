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

template <typename H, typename ACC, bool OWN>
__global__ void __launch_bounds__(PW_THREADS)
gfi_kernel(const H *__restrict__ hand, const ACC *__restrict__ acc, int64_t n, double expo, double scale, double s2,
           float *__restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * PW_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * PW_THREADS) {
        const H h = hand[i];
        float r = ND_F;
        if (!(h <= (H)ND_I)) {
            double a = (double)acc[i];
            if (OWN && acc[i] == 0) a = 1.0;  // gfi.py:432-435
            r = (float)log(scale * pow(a * s2, expo) / ((double)h + 0.01));
        }
        out[i] = r;
    }
}

template <typename ACC>
__global__ void __launch_bounds__(PW_THREADS)
ti_mti_kernel(const ACC *__restrict__ acc, const float *__restrict__ beta, int64_t n, double p2, double expo,
              float *__restrict__ ti, float *__restrict__ mti)
{
    for (int64_t i = (int64_t)blockIdx.x * PW_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * PW_THREADS) {
        const ACC a0 = acc[i];
        float t1 = ND_F, t2 = ND_F;
        if (!(a0 <= (ACC)ND_I)) {  // topoindexes.py:252, 286
            const double a = (a0 == 0) ? 1.0 : (double)a0;
            const double t = tan((double)beta[i] + 0.01);
            if (ti) t1 = (float)log((a * p2) / t);
            if (mti) t2 = (float)log(pow(a * p2, expo) / t);
        }
        if (ti) ti[i] = t1;
        if (mti) mti[i] = t2;
    }
}

__global__ void __launch_bounds__(PW_THREADS)
slope_rad_kernel(const float *__restrict__ pct, int64_t n, float *__restrict__ rad)
{
    for (int64_t i = (int64_t)blockIdx.x * PW_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * PW_THREADS) {
        const float s = pct[i];
        rad[i] = (s == ND_F) ? ND_F : atanf(s / 100.0f);  // example.py:63-64
    }
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
    const unsigned b = pw_blocks(n);
    const double s2 = size * size;
    if (hand_dtype == DTB_F32 && acc_dtype == DTB_I64)
        gfi_kernel<float, int64_t, OWN><<<b, PW_THREADS, 0, st>>>((const float *)hand, (const int64_t *)acc, n, expo, scale, s2, out);
    else if (hand_dtype == DTB_F32 && acc_dtype == DTB_I32)
        gfi_kernel<float, int32_t, OWN><<<b, PW_THREADS, 0, st>>>((const float *)hand, (const int32_t *)acc, n, expo, scale, s2, out);
    else if (hand_dtype == DTB_I16 && acc_dtype == DTB_I64)
        gfi_kernel<int16_t, int64_t, OWN><<<b, PW_THREADS, 0, st>>>((const int16_t *)hand, (const int64_t *)acc, n, expo, scale, s2, out);
    else if (hand_dtype == DTB_I16 && acc_dtype == DTB_I32)
        gfi_kernel<int16_t, int32_t, OWN><<<b, PW_THREADS, 0, st>>>((const int16_t *)hand, (const int32_t *)acc, n, expo, scale, s2, out);
    else
        return DTB_ERR_INVALID;
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
    const unsigned b = pw_blocks(n);
    if (acc_dtype == DTB_I64)
        ti_mti_kernel<int64_t><<<b, PW_THREADS, 0, st>>>((const int64_t *)acc, slope_rad, n, px * px, expo, ti, mti);
    else if (acc_dtype == DTB_I32)
        ti_mti_kernel<int32_t><<<b, PW_THREADS, 0, st>>>((const int32_t *)acc, slope_rad, n, px * px, expo, ti, mti);
    else
        return DTB_ERR_INVALID;
    DTB_LAUNCH_CHECK("ti_mti_kernel");
    return DTB_OK;
}

extern "C" int dtb_slope_to_radians(const float *slope_pct, int64_t n, float *slope_rad, void *stream)
{
    if (!slope_pct || !slope_rad || n < 0) return DTB_ERR_INVALID;
    if (n == 0) return DTB_OK;
    slope_rad_kernel<<<pw_blocks(n), PW_THREADS, 0, as_stream(stream)>>>(slope_pct, n, slope_rad);
    DTB_LAUNCH_CHECK("slope_rad_kernel");
    return DTB_OK;
}
