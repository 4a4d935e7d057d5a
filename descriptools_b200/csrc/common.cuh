// common.cuh -- shared host/device helpers for libdtb200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dtb200.h"

namespace dtb {

constexpr float ND_F = -100.0f;
constexpr int ND_I = -100;
constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// --- launch accounting / error plumbing (api.cu) --------------------------------------
void note_launch(int n = 1);
int cuda_fail(cudaError_t e, const char *what);
int cuda_fail_msg(const char *what);

#define DTB_CUDA(call)                                            \
    do {                                                          \
        cudaError_t e__ = (call);                                 \
        if (e__ != cudaSuccess) return dtb::cuda_fail(e__, #call); \
    } while (0)

#define DTB_LAUNCH_CHECK(name)                                      \
    do {                                                            \
        cudaError_t e__ = cudaGetLastError();                       \
        if (e__ != cudaSuccess) return dtb::cuda_fail(e__, name);   \
        dtb::note_launch();                                         \
    } while (0)

// --- optional per-kernel timing (dtb_profile_enable / dtb_profile_collect, api.cu): CUDA events on the
// launching stream around a kernel launch, so bench.py can report the dominant kernel's own duration.
void prof_begin(const char *name, cudaStream_t st);
void prof_end(cudaStream_t st);
struct ProfScope {
    cudaStream_t st;
    ProfScope(const char *name, cudaStream_t s) : st(s) { prof_begin(name, s); }
    ~ProfScope() { prof_end(st); }
};
// launch `...` (a kernel<<<>>>(...) expression) under the name `name`, check it, count it
#define DTB_KERNEL(name, st, ...)                     \
    do {                                              \
        {                                             \
            dtb::ProfScope ps__(name, st);            \
            __VA_ARGS__;                              \
        }                                             \
        DTB_LAUNCH_CHECK(name);                       \
    } while (0)

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// D8 code -> linear offset components (flowhand.py:801-824).  dr/dc = 0 for unknown codes.
__host__ __device__ __forceinline__ bool d8_offset(unsigned code, int &dr, int &dc)
{
    // code is a power of two in 1..128; bit index b: 0=E 1=SE 2=S 3=SW 4=W 5=NW 6=N 7=NE
    if (code == 0u || (code & (code - 1u)) != 0u || code > 128u) { dr = 0; dc = 0; return false; }
#ifdef __CUDA_ARCH__
    const int b = __ffs((int)code) - 1;
#else
    int b = 0; while (!((code >> b) & 1u)) ++b;
#endif
    // dr: E0 SE1 S1 SW1 W0 NW-1 N-1 NE-1 ; dc: E1 SE1 S0 SW-1 W-1 NW-1 N0 NE1
    // packed 2-bit fields (value+1), LSB first
    const unsigned DR = 0x01A9u;  // fields: 1,2,2,2,1,0,0,0
    const unsigned DC = 0x901Au;  // fields: 2,2,1,0,0,0,1,2
    dr = (int)((DR >> (2 * b)) & 3u) - 1;
    dc = (int)((DC >> (2 * b)) & 3u) - 1;
    return true;
}

__host__ __device__ __forceinline__ bool d8_is_diag(unsigned code) { return (code & 0xAAu) != 0u; }

#ifdef __CUDACC__
// Natural logarithm with |error| < 3e-9 (absolute, on results up to ~700) in ~30 instructions:
// x = 2^e * m, m in [sqrt(1/2), sqrt(2)); s = (m-1)/(m+1); ln m = 2 atanh(s) = 2s + s z P(z), z = s^2.
// e ln 2 + 2s is kept in f64 (the quotient uses rcp.approx (~2^-20) + one Newton step); the correction s z P(z)
// <= 3.4e-3 is evaluated in f32 (series through s^9, |s| <= 0.1716 -> truncation 7e-10, rounding ~1e-9).
// Used where the result is rounded to f32 afterwards (GFI: gfi.py:292-294; the test tolerance there is 1e-5
// relative + 1e-6 absolute).  fast_log_pos requires a positive, normal, finite argument; fast_log sends
// everything else to libdevice's log().
__device__ __forceinline__ double fast_log_pos(double x)
{
    const int hi = __double2hiint(x);
    int e = (hi >> 20) - 1023;
    int mhi = (hi & 0x000FFFFF) | 0x3FF00000;  // mantissa in [1, 2)
    if (mhi >= 0x3FF6A09F) { mhi -= 0x00100000; ++e; }  // >= sqrt(2): halve
    const double m = __hiloint2double(mhi, __double2loint(x));
    const double f = m - 1.0, d = m + 1.0;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    r = r * (2.0 - d * r);
    const double s = f * r;
    const float sf = (float)s, zf = sf * sf;
    float p = 2.0f / 9.0f;
    p = fmaf(p, zf, 2.0f / 7.0f);
    p = fmaf(p, zf, 2.0f / 5.0f);
    p = fmaf(p, zf, 2.0f / 3.0f);
    const double lm = fma(2.0, s, (double)(sf * zf * p));
    return fma((double)e, 0.693147180559945309417232, lm);
}
__device__ __forceinline__ double fast_log(double x)
{
    const int hi = __double2hiint(x);
    if (hi < 0x00100000 || hi >= 0x7FF00000) return log(x);
    return fast_log_pos(x);
}
#endif

}  // namespace dtb
