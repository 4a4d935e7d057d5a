#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include <stdint.h>
#define DTB200_H
#include "../descriptools_b200/csrc/common.cuh"
__global__ void k(const double* x, double* y, int n){ int i=blockIdx.x*blockDim.x+threadIdx.x; if(i<n) y[i]=dtb::fast_log(x[i]); }
int main(){ const int n=1<<22; double *x,*y; cudaMallocManaged(&x,n*8); cudaMallocManaged(&y,n*8);
 for(int i=0;i<n;i++){ double u=(i+0.5)/n; x[i]= (i%3==0)? 0.01+u*700 : (i%3==1? exp(u*45.0) : 1.0 + (u-0.5)*1e-3*(i%7)); }
 k<<<n/256,256>>>(x,y,n); cudaDeviceSynchronize(); double me=0; int mi=0; for(int i=0;i<n;i++){ double e=fabs(y[i]-log(x[i])); if(e>me){me=e;mi=i;} }
 printf("max abs err %.3e at x=%.17g\n", me, x[mi]); return 0; }
