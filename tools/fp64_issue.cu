// Issue cadence of independent FP64 instructions from ONE warp per SM sub-partition against two (B200).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/fp64_issue tools/fp64_issue.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_kernel(double *out, long long *cyc, double a, double b, int iters) {
    double acc[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc[i] = threadIdx.x + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) acc[i] = fma(acc[i], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// mixed: per "step" 16 DFMA + 8 LDS.128 (broadcast) like the walk
__global__ void mix_kernel(double *out, long long *cyc, double a, int iters) {
    __shared__ double2 g[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) g[i] = make_double2(1e-3 * i, 2e-3 * i);
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const double2 v = g[(it * 8 + i) & 255];
            acc[2 * i] = fma(-v.x, a, acc[2 * i]);
            acc[2 * i + 1] = fma(-v.y, a, acc[2 * i + 1]);
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    double *out;
    long long *cyc, h[4];
    cudaMalloc(&out, 1 << 20);
    cudaMalloc(&cyc, 64);
    const int iters = 4096;
    for (int threads : {32, 128, 256, 512}) {
        dfma_kernel<16><<<1, threads>>>(out, cyc, 1.0000001, 1e-9, iters);
        cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("16 independent DFMA chains, %3d threads (%d warps per sub-partition): %.2f cycles per warp-DFMA\n", threads, (threads + 127) / 128,
               (double)h[0] / (16.0 * iters));
        dfma_kernel<4><<<1, threads>>>(out, cyc, 1.0000001, 1e-9, iters);
        cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        printf(" 4 independent DFMA chains, %3d threads: %.2f cycles per warp-DFMA\n", threads, (double)h[0] / (4.0 * iters));
        mix_kernel<<<1, threads>>>(out, cyc, 1.0000001, iters);
        cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        printf(" 16 DFMA + 8 LDS.128 per step, %3d threads: %.2f cycles per step\n", threads, (double)h[0] / iters);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
