// micro-benchmark: dependent-chain latency and multi-warp throughput of DFMA / DMUL / DADD / MUFU.RCP64H on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, long long* cyc, int n, double a, double b) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) x = __fma_rn(x, b, a);
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void lat_ilp4(double* out, long long* cyc, int n, double a, double b) {
    double x0 = a + threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { x0 = __fma_rn(x0, b, a); x1 = __fma_rn(x1, b, a); x2 = __fma_rn(x2, b, a); x3 = __fma_rn(x3, b, a); }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void lat_f32(float* out, long long* cyc, int n, float a, float b) {
    float x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) x = __fmaf_rn(x, b, a);
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 1 << 16);
    long long h[8];
    const int n = 4096;
    for (int warps = 1; warps <= 32; warps *= 2) {
        lat<<<1, 32 * warps>>>(out, cyc, n, 1e-9, 0.999999);
        cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        double c1 = (double)h[0] / (n * 16);
        lat_ilp4<<<1, 32 * warps>>>(out, cyc, n, 1e-9, 0.999999);
        cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        double c4 = (double)h[0] / (n * 16);
        lat_f32<<<1, 32 * warps>>>((float*)out, cyc, n, 1e-9f, 0.999999f);
        cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        double cf = (double)h[0] / (n * 16);
        printf("warps/SM %2d: DFMA chain %.2f cyc/instr | 4 chains %.2f cyc/instr | FFMA chain %.2f cyc/instr\n", warps, c1, c4, cf);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
