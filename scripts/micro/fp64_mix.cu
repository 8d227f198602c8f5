// micro-benchmark: per-SM throughput of DFMA / DMUL / DADD streams and of mixes on sm_100a (16 warps per SM, 8 independent
// chains per thread): is a DMUL or DADD cheaper to issue than a DFMA?
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) mix(double* out, long long* cyc, int n, double a, double b) {
    double x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = a + threadIdx.x + k;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (MODE == 0) x[k] = __fma_rn(x[k], b, a);
                if (MODE == 1) x[k] = __dmul_rn(x[k], b);
                if (MODE == 2) x[k] = __dadd_rn(x[k], a);
                if (MODE == 3) x[k] = (r & 1) ? __dmul_rn(x[k], b) : __dadd_rn(x[k], a);
                if (MODE == 4) x[k] = (r & 1) ? __fma_rn(x[k], b, a) : __dmul_rn(x[k], b);
            }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 26); cudaMalloc(&cyc, 1 << 16);
    const int n = 2048;
    const char* names[5] = {"DFMA", "DMUL", "DADD", "DMUL+DADD", "DFMA+DMUL"};
    for (int blocks_per_sm = 1; blocks_per_sm <= 2; ++blocks_per_sm) {
        for (int m = 0; m < 5; ++m) {
            long long h[1];
            const int grid = 148 * blocks_per_sm;
            for (int rep = 0; rep < 2; ++rep) {
                if (m == 0) mix<0><<<grid, 256>>>(out, cyc, n, 1e-9, 0.999999);
                if (m == 1) mix<1><<<grid, 256>>>(out, cyc, n, 1e-9, 0.999999);
                if (m == 2) mix<2><<<grid, 256>>>(out, cyc, n, 1e-9, 0.999999);
                if (m == 3) mix<3><<<grid, 256>>>(out, cyc, n, 1e-9, 0.999999);
                if (m == 4) mix<4><<<grid, 256>>>(out, cyc, n, 1e-9, 0.999999);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
            const double warp_instr = (double)n * 32 * 8 * blocks_per_sm;       // per SM: 8 warps x n x 32 instr per block
            printf("%d x 8 warps/SM %-10s: %.3f cycles per warp-instruction per SM (%.2f per scheduler)\n", blocks_per_sm, names[m],
                   (double)h[0] / warp_instr, 4.0 * h[0] / warp_instr);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
