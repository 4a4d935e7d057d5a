// pipe_probe.cu -- issue-rate probe for the instruction mix of the slope/D8 stencil on sm_100a.
// For every op (and a few 1:1 mixes) it runs 8 independent chains per thread, 32 warps per SM on
// all SMs, and prints warp-instructions per clock per SM (SM clock from clock64()).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/pipe_probe scripts/pipe_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define NC 4
typedef unsigned long long u64;

enum Op { FADD, FADD2, FFMA, FFMA2, FMULSAT, FFMASAT, FMNMX, FMNMX3, FMNMX3NAN, FSET, FSETP_SEL, LOP3, SHF, IADD, IMAD, PRMT,
          VIMNMX3, FLO, F2F64, DMUL, DFMA, LDS32, LDS128, VOTE, MIX_FADD2_FMNMX3, MIX_FFMA2_FSET, MIX_FADD_FMNMX3,
          MIX_FFMA2_FFMA, MIX_FADD2_LDS, MIX3, MIX_FADD_FMNMX, MIX_FADD_FSET, MIX_2FADD_FSET, MIX_FMNMX_FSET, MIX_FFMA_FSET, MIX_FFMA2_FMNMX_FSET, FSEL1, MOV1, NOPS };
static const char *names[] = {"FADD", "FADD2", "FFMA", "FFMA2", "FMUL.SAT", "FFMA.SAT", "FMNMX", "FMNMX3", "FMNMX3.NAN", "FSET",
                              "FSETP+FSEL(2)", "LOP3", "SHF", "IADD3", "IMAD", "PRMT", "VIMNMX3.U32", "FLO", "F2F.64.32+back(2)",
                              "DMUL", "DFMA", "LDS.32", "LDS.128", "VOTE.ANY", "FADD2+FMNMX3(2)", "FFMA2+FSET(2)",
                              "FADD+FMNMX3(2)", "FFMA2+FFMA(2)", "FADD2+LDS.128(2)", "FFMA2+FMNMX3+FSET+FFMA(4)", "FADD+FMNMX(2)", "FADD+FSET(2)", "2FADD+FSET(3)", "FMNMX+FSET(2)", "FFMA+FSET(2)", "FFMA2+FMNMX+FSET(3)", "FSEL", "MOV"};

template <int OP>
__global__ void __launch_bounds__(1024, 1) probe(float *out, u64 *cycles, float seed)
{
    __shared__ __align__(16) float sm[1024 + 8];
    for (int i = threadIdx.x; i < 1032; i += 1024) sm[i] = seed * i;
    __syncthreads();
    float a[NC], b[NC];
    u64 p[NC];
    double d[NC];
    unsigned u[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        a[i] = seed * (threadIdx.x + i);
        b[i] = seed + i;
        asm("mov.b64 %0, {%1,%2};" : "=l"(p[i]) : "f"(a[i]), "f"(b[i]));
        d[i] = (double)a[i];
        u[i] = threadIdx.x * 2654435761u + i;
    }
    const float c1 = seed * 1.0001f, c2 = seed * 0.5f;
    u64 pc;
    asm("mov.b64 %0, {%1,%2};" : "=l"(pc) : "f"(c1), "f"(c2));
    const float *lp = sm + (threadIdx.x & 255) * 4;
    u64 t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            if (OP == FADD) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c1));
            if (OP == FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc));
            if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c1), "f"(b[i]));
            if (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(pc));
            if (OP == FMULSAT) asm volatile("mul.sat.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c1));
            if (OP == FFMASAT) asm volatile("fma.rn.sat.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c1), "f"(b[i]));
            if (OP == FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
            if (OP == FMNMX3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(c1));
            if (OP == FMNMX3NAN) asm volatile("max.NaN.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(c1));
            if (OP == FSET) asm volatile("set.neu.f32.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
            if (OP == FSETP_SEL)
                asm volatile("{.reg .pred q; setp.lt.f32 q, %0, %1; selp.f32 %0, %1, %2, q;}" : "+f"(a[i]) : "f"(b[i]), "f"(c1));
            if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) & (NC - 1)]), "r"(0x1234567u + it));
            if (OP == SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, 3;" : "+r"(u[i]) : "r"(u[(i + 1) & (NC - 1)]));
            if (OP == IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & (NC - 1)]));
            if (OP == IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(u[(i + 1) & (NC - 1)]), "r"(it));
            if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(u[(i + 1) & (NC - 1)]), "r"(0x3210u + it));
            if (OP == VIMNMX3) { u[i] = max(max(u[i], u[(i + 1) & (NC - 1)]), (unsigned)it * 77u); asm volatile("" : "+r"(u[i])); }
            if (OP == FLO) asm volatile("bfind.u32 %0, %0;" : "+r"(u[i]));
            if (OP == F2F64) {
                double t;
                asm volatile("cvt.f64.f32 %0, %1;" : "=d"(t) : "f"(a[i]));
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(a[i]) : "d"(t));
            }
            if (OP == DMUL) asm volatile("mul.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(1.0000001));
            if (OP == DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(d[i]) : "d"(1.0000001));
            if (OP == LDS32) {
                float t;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"((unsigned)__cvta_generic_to_shared(lp + i)));
                a[i] += t;
            }
            if (OP == LDS128) {
                float4 t;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                             : "r"((unsigned)__cvta_generic_to_shared(lp)));
                a[i] += t.x;
            }
            if (OP == VOTE) {
                unsigned r;
                asm volatile("{.reg .pred q; setp.lt.f32 q, %1, %2; vote.sync.any.pred q, q, 0xffffffff; selp.u32 %0, 1, 0, q;}"
                             : "=r"(r) : "f"(a[i]), "f"(b[i]));
                u[i] += r;
            }
            if (OP == MIX_FADD2_FMNMX3) {
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc));
                asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(c1));
            }
            if (OP == MIX_FFMA2_FSET) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(pc));
                asm volatile("set.neu.f32.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
            }
            if (OP == MIX_FADD_FMNMX3) {
                asm volatile("add.f32 %0, %0, %1;" : "+f"(b[i]) : "f"(c1));
                asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c2), "f"(c1));
            }
            if (OP == MIX_FFMA2_FFMA) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(pc));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c1), "f"(c2));
            }
            if (OP == MIX_FADD2_LDS) {
                float4 t;
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc));
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                             : "r"((unsigned)__cvta_generic_to_shared(lp)));
                a[i] += t.y;
            }

            if (OP == MIX_FADD_FMNMX) {
                asm volatile("add.f32 %0, %0, %1;" : "+f"(b[i]) : "f"(c1));
                asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c2));
            }
            if (OP == MIX_FADD_FSET) {
                asm volatile("add.f32 %0, %0, %1;" : "+f"(b[i]) : "f"(c1));
                asm volatile("set.neu.f32.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c2));
            }
            if (OP == MIX_2FADD_FSET) {
                asm volatile("add.f32 %0, %0, %1;" : "+f"(b[i]) : "f"(c1));
                float t = __uint_as_float(u[i]);
                asm volatile("add.f32 %0, %0, %1;" : "+f"(t) : "f"(c2));
                u[i] = __float_as_uint(t);
                asm volatile("set.neu.f32.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c2));
            }
            if (OP == MIX_FMNMX_FSET) {
                asm volatile("max.f32 %0, %0, %1;" : "+f"(b[i]) : "f"(c1));
                asm volatile("set.neu.f32.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c2));
            }
            if (OP == MIX_FFMA_FSET) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(b[i]) : "f"(c1), "f"(c2));
                asm volatile("set.neu.f32.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c2));
            }
            if (OP == MIX_FFMA2_FMNMX_FSET) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(pc));
                asm volatile("max.f32 %0, %0, %1;" : "+f"(b[i]) : "f"(c1));
                asm volatile("set.neu.f32.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c2));
            }
            if (OP == FSEL1) asm volatile("{.reg .pred q; setp.ne.u32 q, %2, 0; selp.f32 %0, %0, %1, q;}" : "+f"(a[i]) : "f"(b[i]), "r"(it & (1 << i)));
            if (OP == MOV1) { asm volatile("mov.b32 %0, %1;" : "=f"(a[i]) : "f"(b[i])); asm volatile("mov.b32 %0, %1;" : "=f"(b[i]) : "f"(a[(i+1)&(NC-1)])); }
            if (OP == MIX3) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(pc));
                asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c2), "f"(c1));
                asm volatile("set.neu.f32.f32 %0, %0, %1;" : "+f"(b[i]) : "f"(c2));
                float t = __uint_as_float(u[i]);
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(t) : "f"(c1), "f"(c2));
                u[i] = __float_as_uint(t);
            }
        }
    }
    u64 t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        float x, y;
        asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(p[i]));
        s += a[i] + b[i] + x + y + (float)d[i] + (float)u[i];
    }
    out[blockIdx.x * 1024 + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(float *out, u64 *cyc, int nblk, int per)
{
    probe<OP><<<nblk, 1024>>>(out, cyc, 1.0f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<OP><<<nblk, 1024>>>(out, cyc, 1.0f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    u64 *h = new u64[nblk];
    cudaMemcpy(h, cyc, nblk * sizeof(u64), cudaMemcpyDeviceToHost);
    u64 mx = 0;
    for (int i = 0; i < nblk; ++i) mx = h[i] > mx ? h[i] : mx;
    delete[] h;
    // warp instructions per SM: 4 blocks x 8 warps x ITERS x 8 x per
    const double winst = 32.0 * ITERS * NC * per;
    printf("%-30s %8.3f ms (%7.0f clk @1965MHz) %9llu cyc  %6.3f warp-inst/clk/SM  (%5.2f clk per warp-inst per SMSP)  err=%d\n", names[OP], ms, ms * 1965e3, mx,
           winst / (double)mx, (double)mx * 4 / winst, (int)cudaGetLastError());
}

int main()
{
    const int nblk = 148;
    float *out;
    u64 *cyc;
    cudaMalloc(&out, nblk * 1024 * sizeof(float));
    cudaMalloc(&cyc, nblk * sizeof(u64));
    run<FADD>(out, cyc, nblk, 1);
    run<FADD2>(out, cyc, nblk, 1);
    run<FFMA>(out, cyc, nblk, 1);
    run<FFMA2>(out, cyc, nblk, 1);
    run<FMULSAT>(out, cyc, nblk, 1);
    run<FFMASAT>(out, cyc, nblk, 1);
    run<FMNMX>(out, cyc, nblk, 1);
    run<FMNMX3>(out, cyc, nblk, 1);
    run<FMNMX3NAN>(out, cyc, nblk, 1);
    run<FSET>(out, cyc, nblk, 1);
    run<FSETP_SEL>(out, cyc, nblk, 2);
    run<LOP3>(out, cyc, nblk, 1);
    run<SHF>(out, cyc, nblk, 1);
    run<IADD>(out, cyc, nblk, 1);
    run<IMAD>(out, cyc, nblk, 1);
    run<PRMT>(out, cyc, nblk, 1);
    run<VIMNMX3>(out, cyc, nblk, 1);
    run<FLO>(out, cyc, nblk, 1);
    run<F2F64>(out, cyc, nblk, 2);
    run<DMUL>(out, cyc, nblk, 1);
    run<DFMA>(out, cyc, nblk, 1);
    run<LDS32>(out, cyc, nblk, 2);
    run<LDS128>(out, cyc, nblk, 2);
    run<VOTE>(out, cyc, nblk, 3);
    run<MIX_FADD2_FMNMX3>(out, cyc, nblk, 2);
    run<MIX_FFMA2_FSET>(out, cyc, nblk, 2);
    run<MIX_FADD_FMNMX3>(out, cyc, nblk, 2);
    run<MIX_FFMA2_FFMA>(out, cyc, nblk, 2);
    run<MIX_FADD2_LDS>(out, cyc, nblk, 3);
    run<MIX3>(out, cyc, nblk, 4);
    run<MIX_FADD_FMNMX>(out, cyc, nblk, 2);
    run<MIX_FADD_FSET>(out, cyc, nblk, 2);
    run<MIX_2FADD_FSET>(out, cyc, nblk, 3);
    run<MIX_FMNMX_FSET>(out, cyc, nblk, 2);
    run<MIX_FFMA_FSET>(out, cyc, nblk, 2);
    run<MIX_FFMA2_FMNMX_FSET>(out, cyc, nblk, 3);
    run<FSEL1>(out, cyc, nblk, 2);
    run<MOV1>(out, cyc, nblk, 2);
    return 0;
}
