// fp64_latency.cu -- dependent-issue latencies that bound the serial in-block walk of the GPFQ sweep (not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/fp64_latency tools/fp64_latency.cu && gpurun_out/fp64_latency
// One warp, one CTA: a chain of N dependent operations, timed with clock64().
#include <cuda_runtime.h>
#include <stdio.h>

#define N 2048

__global__ void k_chain(double *out, long long *cyc, double a, double b, int mode) {
    __shared__ double sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = a + threadIdx.x * 1e-3;
    __syncthreads();
    double x = a + threadIdx.x * 1e-6, y = b;
    int idx = threadIdx.x & 31;
    long long t0 = clock64();
    if (mode == 0) {
#pragma unroll 16
        for (int i = 0; i < N; ++i) x = fma(x, a, b);
    } else if (mode == 1) {
#pragma unroll 16
        for (int i = 0; i < N; ++i) x = x + b;
    } else if (mode == 2) {
#pragma unroll 16
        for (int i = 0; i < N; ++i) x = x * a;
    } else if (mode == 3) {  // compare + select chain
#pragma unroll 16
        for (int i = 0; i < N; ++i) { x = (x < y) ? x + b : y; y = y + 1e-9; }
    } else if (mode == 4) {  // LDS.64 pointer chase
#pragma unroll 16
        for (int i = 0; i < N; ++i) { x = sm[idx]; idx = (int)(__double_as_longlong(x) & 31); }
    } else if (mode == 5) {  // shuffle (64-bit = 2 SHFL) chain
#pragma unroll 16
        for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
    } else if (mode == 6) {  // floor + double -> int -> double
#pragma unroll 16
        for (int i = 0; i < N; ++i) { int k = (int)floor(x); x = (double)k + 0.25; }
    } else if (mode == 7) {  // division (IEEE)
#pragma unroll 16
        for (int i = 0; i < N; ++i) x = b / x;
    } else if (mode == 8) {  // reciprocal-multiply with Markstein correction (3 dependent ops)
        const double r = 1.0 / a;
#pragma unroll 16
        for (int i = 0; i < N; ++i) { double q0 = x * r; double e = fma(-q0, a, x); x = fma(e, r, q0) + b; }
    } else if (mode == 9) {  // fp32 FFMA chain for comparison
        float f = (float)x, fa = (float)a, fb = (float)b;
#pragma unroll 16
        for (int i = 0; i < N; ++i) f = fmaf(f, fa, fb);
        x = f;
    } else if (mode == 10) {  // DSETP -> predicate -> DADD
#pragma unroll 16
        for (int i = 0; i < N; ++i) { if (fabs(x) < y) x = x + b; else x = x - b; }
    } else if (mode == 11) {  // __syncthreads round trip (256 threads)
#pragma unroll 16
        for (int i = 0; i < N; ++i) { __syncthreads(); }
    } else if (mode == 12) {  // double -> int conversion alone (F2I.F64) + I2F
#pragma unroll 16
        for (int i = 0; i < N; ++i) { int k = __double2int_rd(x); x = __int2double_rn(k) ; }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[mode] = t1 - t0;
    if (x == 12345.678) out[0] = x + y + idx;
}

int main() {
    const char *names[] = {"DFMA", "DADD", "DMUL", "DSETP+SEL+DADD", "LDS.64 chase", "SHFL.64", "floor+F2I+I2F+DADD", "DDIV",
                           "rcp*x + Markstein (3 ops) + DADD", "FFMA (fp32)", "DSETP+branchless DADD", "__syncthreads (256 thr)",
                           "F2I.F64 + I2F.F64"};
    double *out;
    long long *cyc, h[16];
    cudaMalloc(&out, 8);
    cudaMalloc(&cyc, 16 * 8);
    for (int threads : {32, 256}) {
        printf("---- %d threads (1 CTA)\n", threads);
        for (int mode = 0; mode <= 12; ++mode) {
            k_chain<<<1, threads>>>(out, cyc, 1.0000001, 1e-7, mode);
            k_chain<<<1, threads>>>(out, cyc, 1.0000001, 1e-7, mode);
            cudaDeviceSynchronize();
            cudaMemcpy(h, cyc, 16 * 8, cudaMemcpyDeviceToHost);
            printf("%-36s %8.1f cycles per link\n", names[mode], (double)h[mode] / N);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
