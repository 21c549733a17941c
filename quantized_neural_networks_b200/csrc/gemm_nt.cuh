// gemm_nt.cuh -- fp64-accumulating "NT" contraction  C[i][j] (+)= sum_seg sign * sum_k A[i][k] * B[j][k]
// on the sm_100a DMMA pipe (mma.sync m8n8k4 f64), operands fp32 or fp64, both K-contiguous.
//
// Used for:
//   K1 Gram stage      G1 = Xq X^T, G2 = Xq Xq^T over the m samples (A = Xq rows, B = X rows, fp32 in,
//                      every product exact in fp64, fp64 accumulation) -- lower-triangular tiles only,
//                      split-K over samples with a fixed-order (deterministic) reduction.
//   K2 sweep panels    D^T[j][t] = Wt[j][:p] . G1[t][:p] - Qt[j][:p] . G2[t][:p]  (two segments, fp64 in)
//
// Tiling: CTA tile BM x BN, one warp per 32 x 32 sub-tile (4 x 4 DMMA atoms, 32 fp64 accumulators
// per lane), K staged through a 3-deep cp.async ring in shared memory.  Row pitch BK+4 elements makes
// both fragment loads bank-conflict free (fp32: pitch = 4 mod 32 words; fp64: 4 mod 16 doubles).
#pragma once
#include "common.cuh"

struct GemmSeg {
    const void *A;
    const void *B;
    int64_t lda, ldb, K;
    double sign;
};

struct GemmArgs {
    GemmSeg seg[2];
    int nseg;
    int64_t M, N;
    double *C;
    int64_t ldc;
    int64_t split_stride;  // elements between split-K partial outputs
    int nsplit;
    int lower_only;        // skip tiles that lie entirely above the diagonal (Gram stage)
    int tiles_n;
    int64_t batch_strideA1, batch_strideC;  // blockIdx.z batches (multi-alphabet panels): seg[1].A and C
    int64_t batch_strideA0;                 // ... and seg[0].A (0: shared by the batches)
    int accumulate;                         // C += result instead of C = result (nsplit must be 1)
    int64_t batch_strideB;                  // ... and B of both segments (0: shared): batched block-diagonal Gram tiles
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
template <int BYTES>
__device__ __forceinline__ void cp_async_small(void *smem, const void *gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;\n" ::"r"(s), "l"(gmem), "n"(BYTES), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <typename T, int BM, int BN, int BK>
struct GemmCfg {
    static constexpr int LD = BK + 4;
    static constexpr int STAGES = 3;
    static constexpr int WARPS_M = BM / 32, WARPS_N = BN / 32;
    static constexpr int THREADS = WARPS_M * WARPS_N * 32;
    static constexpr int CH = 16 / (int)sizeof(T);  // elements per 16-byte chunk
    static constexpr size_t SMEM = (size_t)STAGES * (BM + BN) * LD * sizeof(T);
};

// Load one BMxBK (or BNxBK) operand tile into shared memory; rows beyond `rows` and columns beyond
// `kend` are zero-filled by cp.async's src-size operand.
template <typename T, int ROWS, int BK, int THREADS, bool ALIGNED>
__device__ __forceinline__ void load_tile(T *smem, const T *__restrict__ g, int64_t ld, int64_t row0,
                                          int64_t rows, int64_t k0, int64_t kend) {
    constexpr int LD = BK + 4;
    if (ALIGNED) {
        constexpr int CH = 16 / (int)sizeof(T);
        constexpr int CPR = BK / CH;
        for (int idx = threadIdx.x; idx < ROWS * CPR; idx += THREADS) {
            const int r = idx / CPR, c = idx % CPR;
            const int64_t gr = row0 + r, gk = k0 + (int64_t)c * CH;
            int64_t rem = (gr < rows) ? (kend - gk) : 0;
            int bytes = rem <= 0 ? 0 : (rem >= CH ? 16 : (int)rem * (int)sizeof(T));
            const T *src = bytes ? g + gr * ld + gk : g;
            cp_async16(smem + r * LD + c * CH, src, bytes);
        }
    } else {
        for (int idx = threadIdx.x; idx < ROWS * BK; idx += THREADS) {
            const int r = idx / BK, c = idx % BK;
            const int64_t gr = row0 + r, gk = k0 + c;
            const bool ok = (gr < rows) && (gk < kend);
            const T *src = ok ? g + gr * ld + gk : g;
            cp_async_small<(int)sizeof(T)>(smem + r * LD + c, src, ok ? (int)sizeof(T) : 0);
        }
    }
}

template <typename T, int BM, int BN, int BK, bool ALIGNED>
__global__ void __launch_bounds__((BM / 32) * (BN / 32) * 32)
gemm_nt_kernel(const GemmArgs g) {
    using Cfg = GemmCfg<T, BM, BN, BK>;
    constexpr int LD = Cfg::LD, STAGES = Cfg::STAGES, THREADS = Cfg::THREADS;
    extern __shared__ __align__(16) unsigned char gemm_smem[];
    T *As = reinterpret_cast<T *>(gemm_smem);
    T *Bs = As + STAGES * BM * LD;

    const int ti = blockIdx.x / g.tiles_n, tj = blockIdx.x % g.tiles_n;
    const int64_t i0 = (int64_t)ti * BM, j0 = (int64_t)tj * BN;
    if (g.lower_only && j0 > i0 + BM - 1) return;
    const int split = blockIdx.y;
    const int batch = blockIdx.z;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wr = (warp / Cfg::WARPS_N) * 32, wc = (warp % Cfg::WARPS_N) * 32;
    const int grp = lane >> 2, tig = lane & 3;

    // flat list of k-tiles over the (at most two) segments handled by this split
    int64_t kbeg[2], kend[2];
    int ntile[2], total = 0;
    for (int s = 0; s < 2; ++s) {
        ntile[s] = 0; kbeg[s] = kend[s] = 0;
        if (s < g.nseg) {
            int64_t per = ((g.seg[s].K + g.nsplit - 1) / g.nsplit + BK - 1) / BK * BK;
            kbeg[s] = (int64_t)split * per;
            kend[s] = kbeg[s] + per < g.seg[s].K ? kbeg[s] + per : g.seg[s].K;
            if (kend[s] > kbeg[s]) ntile[s] = (int)((kend[s] - kbeg[s] + BK - 1) / BK);
            total += ntile[s];
        }
    }

    auto issue = [&](int it) {
        if (it < total) {
            const int s = (it < ntile[0]) ? 0 : 1;
            const int kt = (s == 0) ? it : it - ntile[0];
            const int64_t k0 = kbeg[s] + (int64_t)kt * BK;
            const T *A = reinterpret_cast<const T *>(g.seg[s].A) + batch * (s == 1 ? g.batch_strideA1 : g.batch_strideA0);
            const T *B = reinterpret_cast<const T *>(g.seg[s].B) + batch * g.batch_strideB;
            const int st = it % STAGES;
            load_tile<T, BM, BK, THREADS, ALIGNED>(As + st * BM * LD, A, g.seg[s].lda, i0, g.M, k0, kend[s]);
            load_tile<T, BN, BK, THREADS, ALIGNED>(Bs + st * BN * LD, B, g.seg[s].ldb, j0, g.N, k0, kend[s]);
        }
        cp_async_commit();
    };

    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    issue(0);
    issue(1);
    for (int it = 0; it < total; ++it) {
        cp_async_wait<1>();
        __syncthreads();            // stage `it` landed for everyone; stage it-1 fully consumed
        issue(it + 2);
        const int st = it % STAGES;
        const T *as = As + st * BM * LD + (wr + grp) * LD + tig;
        const T *bs = Bs + st * BN * LD + (wc + grp) * LD + tig;
        const bool neg = (it >= ntile[0]) ? (g.seg[1].sign < 0.0) : (g.seg[0].sign < 0.0);
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const double v = (double)as[a * 8 * LD + kk];
                af[a] = neg ? -v : v;
            }
#pragma unroll
            for (int b = 0; b < 4; ++b) bf[b] = (double)bs[b * 8 * LD + kk];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) dmma_m8n8k4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
    }
    cp_async_wait<0>();

    double *C = g.C + (int64_t)split * g.split_stride + (int64_t)batch * g.batch_strideC;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int64_t i = i0 + wr + a * 8 + grp;
        if (i >= g.M) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int64_t j = j0 + wc + b * 8 + tig * 2;
            if (j < g.N) C[i * g.ldc + j] = g.accumulate ? C[i * g.ldc + j] + acc[a][b][0] : acc[a][b][0];
            if (j + 1 < g.N) C[i * g.ldc + j + 1] = g.accumulate ? C[i * g.ldc + j + 1] + acc[a][b][1] : acc[a][b][1];
        }
    }
}

// Fixed-order sum of split-K partials (deterministic, independent of launch geometry).
__global__ void reduce_splits_kernel(const double *__restrict__ part, int nsplit, int64_t stride,
                                     double *__restrict__ out, int64_t n) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
         e += (int64_t)gridDim.x * blockDim.x) {
        double s = part[e];
        for (int k = 1; k < nsplit; ++k) s += part[(int64_t)k * stride + e];
        out[e] = s;
    }
}

template <typename T, int BM, int BN, int BK>
static int launch_gemm_nt(gpfq_ctx *ctx, GemmArgs g, int nbatch) {
    using Cfg = GemmCfg<T, BM, BN, BK>;
    bool aligned = true;
    for (int s = 0; s < g.nseg; ++s) {
        aligned = aligned && ((uintptr_t)g.seg[s].A % 16 == 0) && ((uintptr_t)g.seg[s].B % 16 == 0) &&
                  ((g.seg[s].lda * sizeof(T)) % 16 == 0) && ((g.seg[s].ldb * sizeof(T)) % 16 == 0);
    }
    if (nbatch > 1)
        aligned = aligned && ((g.batch_strideA1 * sizeof(T)) % 16 == 0) && ((g.batch_strideA0 * sizeof(T)) % 16 == 0) &&
                  ((g.batch_strideB * sizeof(T)) % 16 == 0);
    const int tiles_m = (int)ceil_div64(g.M, BM);
    g.tiles_n = (int)ceil_div64(g.N, BN);
    dim3 grid((unsigned)(tiles_m * g.tiles_n), (unsigned)g.nsplit, (unsigned)nbatch);
    if (aligned) {
        auto k = gemm_nt_kernel<T, BM, BN, BK, true>;
        CUDA_TRY(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        k<<<grid, Cfg::THREADS, Cfg::SMEM, ctx->stream>>>(g);
    } else {
        auto k = gemm_nt_kernel<T, BM, BN, BK, false>;
        CUDA_TRY(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        k<<<grid, Cfg::THREADS, Cfg::SMEM, ctx->stream>>>(g);
    }
    KERNEL_CHECK(ctx);
    return GPFQ_OK;
}
