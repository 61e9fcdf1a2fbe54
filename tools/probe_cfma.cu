// Does an fp32 FMA whose weight operand comes from the constant bank (immediate offset) issue faster than one with three register
// operands?  (tools/probe_tmem.cu B0/B1 measured 85 FMA/clock/SM for register operands on B200.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/probe_cfma tools/probe_cfma.cu && tools/_bin/probe_cfma
#include <cstdio>
#include <type_traits>
#include <cuda_runtime.h>
__constant__ float cw[1024];
template <int MODE>
__global__ void __launch_bounds__(512) k(int iters, long long* cyc, float* sink, const float* __restrict__ gw)
{
    __shared__ float sw[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sw[i] = gw[i];
    __syncthreads();
    float x[16], acc[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = 1.f + 0.001f * (threadIdx.x + i);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    const long long t0 = clock64();
    auto body = [&](auto OFFC) {
        constexpr int OFF = decltype(OFFC)::value;
#pragma unroll
        for (int k0 = 0; k0 < 128; k0 += 16) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float w;
                if (MODE == 0) w = cw[OFF * 128 + k0 + i];           // constant bank, immediate offset
                else if (MODE == 1) w = sw[OFF * 128 + k0 + i];      // shared memory broadcast
                else w = x[(i + 5) & 15] * 0.5f;                     // registers
                acc[i & 7] = fmaf(x[i], w, acc[i & 7]);
            }
        }
    };
    for (int it = 0; it < iters; ++it) {
        switch (it & 3) {   // the weights change from tile to tile, at offsets known at compile time
        case 0: body(std::integral_constant<int, 0>()); break;
        case 1: body(std::integral_constant<int, 1>()); break;
        case 2: body(std::integral_constant<int, 2>()); break;
        default: body(std::integral_constant<int, 3>()); break;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] += 1e-7f;
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    if (s == 123.456f) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main()
{
    long long* c; float *sink, *gw;
    cudaMalloc(&c, 8); cudaMalloc(&sink, 4); cudaMalloc(&gw, 4096);
    float h[1024]; for (int i = 0; i < 1024; ++i) h[i] = 1.f / (i + 1);
    cudaMemcpy(gw, h, 4096, cudaMemcpyHostToDevice); cudaMemcpyToSymbol(cw, h, 4096);
    const int iters = 20000;
    for (int warps : {4, 8, 16}) {
        for (int mode = 0; mode < 3; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<148, warps * 32>>>(iters, c, sink, gw);
                else if (mode == 1) k<1><<<148, warps * 32>>>(iters, c, sink, gw);
                else k<2><<<148, warps * 32>>>(iters, c, sink, gw);
                cudaDeviceSynchronize();
            }
            long long hc; cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
            const double fma = (double)iters * 128 * warps * 32;
            printf("%s warps=%2d  %lld cyc  %.1f FMA/clk/SM  (%s)\n", mode == 0 ? "const-bank operand" : mode == 1 ? "LDS operand       " : "register operand  ",
                   warps, hc, fma / hc, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
