// f32x2_check.cu -- does the packed f32x2 pipe round like the scalar one?  Prints the error-free-transformation chain of
// the slope product for one value (debug aid for slope_d8.cu's v2 strip).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long p2;
__device__ __forceinline__ p2 pk(float lo, float hi) { p2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(p2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ p2 sub2(p2 a, p2 b) { p2 r; asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ p2 mul2(p2 a, p2 b) { p2 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ p2 fma2(p2 a, p2 b, p2 c) { p2 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__global__ void k(const float *in, float *out)
{
    const float a = in[0], khi = in[1], klo = in[2], t1 = in[3], t2 = in[4];
    p2 A = pk(a, a), KH = pk(khi, khi), NKH = pk(-khi, -khi), KL = pk(klo, klo), T1 = pk(t1, t1), T2 = pk(t2, t2);
    p2 ny = mul2(A, NKH);
    p2 r = fma2(A, KH, ny);
    p2 c = fma2(A, KL, r);
    p2 ns1 = fma2(c, T1, ny), ns2 = fma2(c, T2, ny);
    p2 acc = fma2(sub2(ns1, ns2), c, pk(0.f, 0.f));
    float x, y;
    upk(ny, x, y); out[0] = x;
    upk(r, x, y); out[1] = x;
    upk(c, x, y); out[2] = x;
    upk(ns1, x, y); out[3] = x;
    upk(ns2, x, y); out[4] = x;
    upk(acc, x, y); out[5] = x;
    // scalar twins
    float sny = __fmul_rn(a, -khi), sr = __fmaf_rn(a, khi, sny), sc = __fmaf_rn(a, klo, sr);
    out[6] = sny; out[7] = sr; out[8] = sc; out[9] = __fmaf_rn(sc, t1, sny); out[10] = __fmaf_rn(sc, t2, sny);
}
int main()
{
    float h[5] = {30.201202f, 100.0f, 0.0f, -(1.0f + 3.814697265625e-06f), -(1.0f - 3.814697265625e-06f)}, o[11];
    float *di, *dout;
    cudaMalloc(&di, sizeof h); cudaMalloc(&dout, sizeof o);
    cudaMemcpy(di, h, sizeof h, cudaMemcpyHostToDevice);
    k<<<1, 1>>>(di, dout);
    cudaMemcpy(o, dout, sizeof o, cudaMemcpyDeviceToHost);
    const char *n[11] = {"ny", "r", "c", "ns1", "ns2", "acc", "s.ny", "s.r", "s.c", "s.ns1", "s.ns2"};
    for (int i = 0; i < 11; ++i) printf("%-6s %.9g  (%a)\n", n[i], o[i], o[i]);
    printf("t1 %a t2 %a\n", h[3], h[4]);
    return 0;
}
