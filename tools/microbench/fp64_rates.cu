// Throughput of the FP64-side instructions the pressure update needs (F2F.F64.F32, DMUL, F2F.F32.F64) versus
// FFMA / FFMA2, measured with clock64 on one full SM load.  Build: nvcc -arch=sm_100a -O3 -o fp64_rates fp64_rates.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, float seed, int iters, long long* cyc)
{
    float a0 = seed + threadIdx.x, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    double d0 = a0, d1 = a1, d2 = a2, d3 = a3;
    float2 p0 = make_float2(a0, a1), p1 = make_float2(a2, a3);
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        if (MODE == 0) { // FFMA
            a0 = __fmaf_rn(a0, 1.0001f, 0.5f); a1 = __fmaf_rn(a1, 1.0001f, 0.5f);
            a2 = __fmaf_rn(a2, 1.0001f, 0.5f); a3 = __fmaf_rn(a3, 1.0001f, 0.5f);
        } else if (MODE == 1) { // DMUL
            d0 = __dmul_rn(d0, 1.0000001); d1 = __dmul_rn(d1, 1.0000001);
            d2 = __dmul_rn(d2, 1.0000001); d3 = __dmul_rn(d3, 1.0000001);
        } else if (MODE == 2) { // F2F.F64.F32 + F2F.F32.F64 round trip (2 conversions per element)
            a0 = __double2float_rn((double)a0 ) ; a1 = __double2float_rn((double)a1);
            a2 = __double2float_rn((double)a2 ) ; a3 = __double2float_rn((double)a3);
            asm volatile("" : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3));
        } else if (MODE == 3) { // the real sequence: cvt, dmul, cvt
            a0 = __double2float_rn(__dmul_rn((double)a0, -1.9)); a1 = __double2float_rn(__dmul_rn((double)a1, -1.9));
            a2 = __double2float_rn(__dmul_rn((double)a2, -1.9)); a3 = __double2float_rn(__dmul_rn((double)a3, -1.9));
        } else if (MODE == 4) { // FFMA2
            p0 = __ffma2_rn(p0, make_float2(1.0001f, 1.0001f), make_float2(0.5f, 0.5f));
            p1 = __ffma2_rn(p1, make_float2(1.0001f, 1.0001f), make_float2(0.5f, 0.5f));
            p0 = __ffma2_rn(p0, make_float2(1.0001f, 1.0001f), make_float2(0.5f, 0.5f));
            p1 = __ffma2_rn(p1, make_float2(1.0001f, 1.0001f), make_float2(0.5f, 0.5f));
        } else if (MODE == 5) { // MUFU.RCP
            a0 = __frcp_rn(a0) ; a1 = __frcp_rn(a1); a2 = __frcp_rn(a2); a3 = __frcp_rn(a3);
        } else if (MODE == 6) { // integer IMAD.WIDE chain
            unsigned long long x0 = __float_as_uint(a0), x1 = __float_as_uint(a1);
            x0 = (unsigned long long)(unsigned)x0 * 0x66666666u + x1; x1 = (unsigned long long)(unsigned)x1 * 0x3ffe6666u + x0;
            a0 = __uint_as_float((unsigned)(x0 >> 7)); a1 = __uint_as_float((unsigned)(x1 >> 9));
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + (float)(d0 + d1 + d2 + d3) + p0.x + p0.y + p1.x + p1.y;
}

template <int MODE>
void run(const char* name, int ops_per_iter)
{
    const int threads = 1024, blocks = 148, iters = 4096;
    float* out; long long* cyc;
    cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, blocks * 8);
    k<MODE><<<blocks, threads>>>(out, 1.0f, iters, cyc);
    k<MODE><<<blocks, threads>>>(out, 1.0f, iters, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < blocks; i++) c += h[i]; c /= blocks;
    // 32 warps per SM, ops_per_iter warp-instructions per iteration per warp
    double per_sm_cycles_per_warp_instr = c / ((double)iters * ops_per_iter * 32);
    printf("%-28s %8.0f cycles  -> %.2f SM-cycles per warp-instruction (%.1f lanes/clk/SM)\n", name, c,
           per_sm_cycles_per_warp_instr, 32.0 / per_sm_cycles_per_warp_instr);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<0>("FFMA", 4);
    run<4>("FFMA2 (f32x2)", 4);
    run<1>("DMUL", 4);
    run<2>("F2F.F64.F32 + F2F.F32.F64", 8);
    run<3>("cvt+DMUL+cvt", 12);
    run<5>("MUFU.RCP (+refine)", 4);
    run<6>("IMAD.WIDE x2 + shifts", 6);
    return 0;
}
