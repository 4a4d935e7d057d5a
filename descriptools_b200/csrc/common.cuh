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

}  // namespace dtb
