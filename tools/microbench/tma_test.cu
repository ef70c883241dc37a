// standalone check of the TMA path used by kernels_advect_tma.cuh: encode a 3-D map, load one box, compare
#include <cstdio>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../smoke-simulation_b200/csrc/kernels_advect_tma.cuh"
using namespace smk;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap m, float* out, int x, int y, int z, int* flag)
{
    extern __shared__ __align__(128) unsigned char smem[];
    float* s = reinterpret_cast<float*>(smem);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + AdvTma::SLOT_BYTES);
    if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(bar, AdvTma::PLANE * 4); tma_load_3d(s, &m, x, y, z, bar); }
    if (!mbar_wait(bar, 0)) { *flag = 1; return; }
    for (int i = threadIdx.x; i < AdvTma::PLANE; i += blockDim.x) out[i] = s[i];
}
int main()
{
    const int P = 72, SY = 41, NZ = 9;
    std::vector<float> h((size_t)P * SY * NZ);
    for (size_t i = 0; i < h.size(); i++) h[i] = (float)i;
    float *d, *out; int* flag;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, AdvTma::PLANE * 4); cudaMalloc(&flag, 4); cudaMemset(flag, 0, 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    printf("entry point: %d %d %p\n", (int)e, (int)q, fn);
    CUtensorMap m;
    const cuuint64_t dims[3] = {P, SY, NZ}, strides[2] = {P * 4, (cuuint64_t)P * SY * 4};
    const cuuint32_t box[3] = {AdvTma::BX, AdvTma::BY, 1}, es[3] = {1, 1, 1};
    CUresult r = ((EncodeTiledFn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    for (int trial = 0; trial < 3; trial++) {
        const int x = trial == 1 ? -4 : trial == 2 ? 60 : 4, y = trial ? -2 : 3, z = trial == 1 ? 8 : trial == 2 ? -1 : 2;
        k<<<1, 128, AdvTma::SLOT_BYTES + 64>>>(m, out, x, y, z, flag);
        e = cudaDeviceSynchronize();
        printf("kernel: %s\n", cudaGetErrorString(e));
        std::vector<float> o(AdvTma::PLANE); int f = 0;
        cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&f, flag, 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int j = 0; j < AdvTma::BY; j++)
            for (int i = 0; i < AdvTma::BX; i++) {
                const int gx = x + i, gy = y + j;
                const float want = (gx < 0 || gy < 0 || gx >= P || gy >= SY || z < 0 || z >= NZ) ? 0.f : h[(size_t)gx + (size_t)gy * P + (size_t)z * P * SY];
                if (o[j * AdvTma::BX + i] != want) bad++;
            }
        printf("trial %d: timeout flag %d, mismatches %d\n", trial, f, bad);
    }
    return 0;
}
