// tma_probe.cu -- standalone probe of cp.async.bulk.tensor.2d variants (debugging aid, not product).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cmath>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BW, int BH>
__global__ void probe(const __grid_constant__ CUtensorMap map, int x, int y, float *out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + ((BW * BH * 4 + 127) / 128) * 128);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(BW * BH * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(smem)), "l"(&map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
    }
    __syncthreads();
    asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WD;\n\tbra WL;\n\tWD:\n\t}"
                 ::"r"(smem_u32(bar)), "r"(0) : "memory");
    const float *t = reinterpret_cast<const float *>(smem);
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = t[i];
}

template <int BW, int BH>
int run(EncodeTiledFn enc, float *d, int rows, int cols, CUtensorMapFloatOOBfill fill, CUtensorMapL2promotion l2, int x, int y, const char *name)
{
    CUtensorMap map;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {BW, BH};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, l2, fill);
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", name, (int)r); return 1; }
    float *out;
    cudaMalloc(&out, BW * BH * 4);
    size_t sm = ((BW * BH * 4 + 127) / 128) * 128 + 64;
    cudaFuncSetAttribute(probe<BW, BH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    probe<BW, BH><<<1, 128, sm>>>(map, x, y, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: FAILED %s\n", name, cudaGetErrorString(e)); return 2; }
    std::vector<float> h(BW * BH);
    cudaMemcpy(h.data(), out, BW * BH * 4, cudaMemcpyDeviceToHost);
    int bad = 0, nan = 0;
    for (int j = 0; j < BH; ++j) for (int k = 0; k < BW; ++k) {
        int rr = y + j, cc = x + k; float v = h[j * BW + k];
        if (rr < 0 || rr >= rows || cc < 0 || cc >= cols) { if (v != v) ++nan; else if (v != 0.f) ++bad; }
        else if (v != (float)(rr * cols + cc)) ++bad;
    }
    printf("%s: ok bad=%d nan=%d\n", name, bad, nan);
    cudaFree(out);
    return 0;
}

int main(int argc, char **argv)
{
    int which = argc > 1 ? atoi(argv[1]) : 0;
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { printf("no entry point\n"); return 1; }
    EncodeTiledFn enc = (EncodeTiledFn)p;
    const int rows = 512, cols = 512;
    std::vector<float> h(rows * cols);
    for (int i = 0; i < rows * cols; ++i) h[i] = (float)i;
    float *d; cudaMalloc(&d, rows * cols * 4); cudaMemcpy(d, h.data(), rows * cols * 4, cudaMemcpyHostToDevice);
    switch (which) {
    case 0: return run<128, 64>(enc, d, rows, cols, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, 0, 0, "128x64 zero-fill in-bounds");
    case 1: return run<132, 66>(enc, d, rows, cols, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, 0, 0, "132x66 zero-fill in-bounds");
    case 2: return run<132, 66>(enc, d, rows, cols, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, -1, -1, "132x66 zero-fill neg coords");
    case 3: return run<132, 66>(enc, d, rows, cols, CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA, CU_TENSOR_MAP_L2_PROMOTION_NONE, -1, -1, "132x66 nan-fill neg coords");
    case 4: return run<132, 66>(enc, d, rows, cols, CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, -1, -1, "132x66 nan-fill neg coords l2-256");
    case 5: return run<132, 66>(enc, d, rows, cols, CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, 383, 447, "132x66 nan-fill bottom-right");
    case 6: return run<64, 66>(enc, d, rows, cols, CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, -1, -1, "64x66 nan-fill neg");
    case 8: return run<136, 66>(enc, d, rows, cols, CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, -4, -1, "136x66 nan-fill x=-4 y=-1");
    case 9: return run<136, 66>(enc, d, rows, cols, CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, 380, 447, "136x66 nan-fill bottom-right aligned");
    case 10: return run<132, 66>(enc, d, rows, cols, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, 1, 1, "132x66 in-bounds x=1 (misaligned)");
    case 11: return run<132, 66>(enc, d, rows, cols, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, 4, 3, "132x66 in-bounds x=4");
    case 7: return run<32, 8>(enc, d, rows, cols, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, 0, 0, "32x8 plain");
    }
    return 0;
}
