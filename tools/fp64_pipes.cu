// fp64_pipes.cu -- calibration microbenchmarks for the ceilings quoted in DESIGN.md (not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/fp64_pipes tools/fp64_pipes.cu
// Measures on the whole chip: DFMA rate, DMMA (mma.sync f64) rates per shape, F2F.F64.F32 rate, DFMA+DMMA mixed,
// and read-only HBM streaming bandwidth.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double a, double b) {
    double r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = fma(r[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += r[i];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) k_cvt(double *out, int iters, float a) {
    float f[8];
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { f[i] = threadIdx.x + i * a; acc[i] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double d = (double)f[i];
            // keep the conversion live with a cheap integer dependency
            long long bits = __double_as_longlong(d);
            f[i] = __int_as_float((int)(bits >> 29) ^ it);
            acc[i] = d;
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

template <int SHAPE>  // 0: m8n8k4, 1: m16n8k4, 2: m16n8k8, 3: m16n8k16
__global__ void __launch_bounds__(256) k_dmma(double *out, int iters, double a0) {
    constexpr int NC = 4;  // independent accumulator sets
    double c[NC][4];
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a0 + i + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = a0 * 2 + i;
#pragma unroll
    for (int j = 0; j < NC; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) c[j][i] = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            if (SHAPE == 0)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a[0]), "d"(b[0]));
            else if (SHAPE == 1)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                             : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
            else if (SHAPE == 2)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                             : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                             : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                               "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < NC; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) s += c[j][i];
    if (s == 12345.678) out[0] = s;
}

// half the warps of each CTA run DFMA chains, the other half DMMA m8n8k4: do the two share one pipe?
__global__ void __launch_bounds__(256) k_mixed(double *out, int iters, double a0) {
    const int warp = threadIdx.x >> 5;
    double s = 0;
    if (warp & 1) {
        double r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = threadIdx.x * 1e-9 + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = fma(r[i], a0, a0);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) s += r[i];
    } else {
        double c[4][2] = {};
        double a = a0 + threadIdx.x, b = a0 * 2;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) s += c[j][0] + c[j][1];
    }
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(512) k_read(const float4 *__restrict__ p, size_t n4, float *out) {
    float s = 0;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        float4 a = __ldg(p + i), b = __ldg(p + i + stride), c = __ldg(p + i + 2 * stride), d = __ldg(p + i + 3 * stride);
        s += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w + c.x + c.y + c.z + c.w + d.x + d.y + d.z + d.w;
    }
    for (; i < n4; i += stride) { float4 a = __ldg(p + i); s += a.x + a.y + a.z + a.w; }
    if (s == 12345.678f) out[0] = s;
}

template <typename F>
static float time_ms(F f, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, clock %d MHz\n", prop.name, sms, prop.clockRate / 1000);
    double *out; CK(cudaMalloc(&out, 64));
    const int iters = 20000;
    for (int cps : {1, 2, 4, 8}) {
        const int grid = sms * cps;
        float ms = time_ms([&] { k_dfma<<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
        double fma = (double)grid * 256 * 8 * iters;
        printf("DFMA       ctas/SM=%d: %8.3f ms  %7.2f TFMA/s  (%.1f FMA/clk/SM @%d MHz nominal)\n", cps, ms, fma / ms / 1e9,
               fma / ms / 1e3 / sms / (prop.clockRate * 1e3) * 1e6 / 1e6, prop.clockRate / 1000);
    }
    for (int cps : {1, 2, 4}) {
        const int grid = sms * cps;
        float ms = time_ms([&] { k_cvt<<<grid, 256>>>(out, iters, 0.5f); });
        double n = (double)grid * 256 * 8 * iters;
        printf("F2F.f64.f32 ctas/SM=%d: %8.3f ms  %7.2f Tcvt/s\n", cps, ms, n / ms / 1e9);
    }
    for (int cps : {1, 2, 4}) {
        const int grid = sms * cps;
        const double warps = (double)grid * 8;
        float ms;
        ms = time_ms([&] { k_dmma<0><<<grid, 256>>>(out, iters, 1.0); });
        printf("DMMA m8n8k4   ctas/SM=%d: %8.3f ms  %7.2f TFMA/s\n", cps, ms, warps * 4 * iters * 256 / ms / 1e9);
        ms = time_ms([&] { k_dmma<1><<<grid, 256>>>(out, iters, 1.0); });
        printf("DMMA m16n8k4  ctas/SM=%d: %8.3f ms  %7.2f TFMA/s\n", cps, ms, warps * 4 * iters * 512 / ms / 1e9);
        ms = time_ms([&] { k_dmma<2><<<grid, 256>>>(out, iters, 1.0); });
        printf("DMMA m16n8k8  ctas/SM=%d: %8.3f ms  %7.2f TFMA/s\n", cps, ms, warps * 4 * iters * 1024 / ms / 1e9);
        ms = time_ms([&] { k_dmma<3><<<grid, 256>>>(out, iters / 2, 1.0); });
        printf("DMMA m16n8k16 ctas/SM=%d: %8.3f ms  %7.2f TFMA/s\n", cps, ms, warps * 4 * (iters / 2) * 2048 / ms / 1e9);
    }
    {
        const int grid = sms * 2;
        float ms = time_ms([&] { k_mixed<<<grid, 256>>>(out, iters, 1.0000001); });
        double dfma = (double)grid * 128 * 8 * iters, dmma = (double)grid * 4 * 4 * iters * 256;
        printf("mixed (half warps DFMA, half DMMA m8n8k4): %8.3f ms  DFMA %7.2f + DMMA %7.2f TFMA/s\n", ms, dfma / ms / 1e9,
               dmma / ms / 1e9);
    }
    {
        const size_t bytes = (size_t)8 << 30;
        float4 *buf; CK(cudaMalloc(&buf, bytes));
        CK(cudaMemset(buf, 0, bytes));
        for (int cps : {2, 4}) {
            float ms = time_ms([&] { k_read<<<sms * cps, 512>>>(buf, bytes / 16, (float *)out); });
            printf("read-only stream, %d ctas/SM x 512 thr: %8.3f ms  %7.1f GB/s\n", cps, ms, bytes / ms / 1e6);
        }
        CK(cudaFree(buf));
    }
    return 0;
}
