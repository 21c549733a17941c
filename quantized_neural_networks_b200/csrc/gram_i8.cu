// gram_i8.cu -- Dense Gram stage on the 5th-generation tensor cores: error-free int8 slicing (Ozaki scheme) +
// tcgen05.mma.kind::i8 with s32 accumulators in TMEM, operands staged by TMA.
//
// The Gram stage G1 = Xq X^T, G2 = Xq Xq^T (quantized_network.py:83-89 in Gram form) is the only dense
// contraction of the path, but its accuracy requirement (<= 1e-9 relative) rules out single-pass
// TF32/BF16 tensor-core products, and there is no fp64 kind of tcgen05.  So every fp32 row is written EXACTLY as
//     x[r][i] = sum_{k=1..S} slice_k[r][i] * 2^(e_r - 7k) + tail,   slice_k in [-127, 127] (int8),  |tail| < 2^(e_r - 7S)
// (e_r = exponent of the row maximum; S = 6 slices = 42 bits below the row maximum), and
//     G[t][s] = 2^(eA_t + eB_s) * sum_{d=2..S+1} 2^(-7d) * sum_{k+l=d} <slice_k(A_t), slice_l(B_s)>
// where every inner sum is an EXACT integer computed by int8 tensor-core MMAs into 32-bit accumulators (K chunks are
// sized so that no accumulator can overflow), shifted and added into two int64 planes with integer atomics
// (exact, order independent => deterministic), and finally combined in fp64.  Terms with k + l > S + 1 and the slice
// tails are dropped: relative to sum |a||b| the error is ~ (S+1) 2^(-7S) * (row max / row mean), i.e. ~1e-11 for
// activation-like rows -- two orders of magnitude inside the parity budget; tests/test_gpu_parity.py checks it against an
// fp64 Gram ("tf32-checked" in the north star's words, here int8-checked).
//
// Kernel anatomy (one CTA per work item = (Gram, 128 x 256 output tile, d, K chunk)):
//   warp 0    TMA producer: cp.async.bulk.tensor.3d (UTMALDG) of a 128 x 128 B A tile and two 128 x 128 B B tiles per
//             stage, 128B-swizzled, 4-stage mbarrier ring
//   warp 1    TMEM allocator + single-thread MMA issuer: 4 x tcgen05.mma.cta_group::1.kind::i8 (M128 N256 K32) per
//             stage, tcgen05.commit frees the stage / publishes the accumulator
//   warps 2-5 epilogue: tcgen05.ld 32x32b.x32 -> shift -> 64-bit integer atomics into the plane
#include <cuda.h>

#include "common.cuh"

namespace i8g {
constexpr int TM = 128, TN = 256, BK = 128;   // output tile, K bytes per stage
constexpr int STAGES = 4;
constexpr int A_BYTES = TM * BK, B_BYTES = TN * BK, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int THREADS = 192;
constexpr int MAX_S = 6;
constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;

struct Item {
    int32_t ti, tj;        // output tile (rows ti*TM.., cols tj*TN..)
    int32_t d;             // k + l
    int32_t gram;          // 0: A = Xq, B = Xq (G2);  1: A = Xq, B = X (G1)
    int32_t plane, shift;  // 0 = hi, 1 = lo; left shift applied before the atomic add
    int32_t k_begin, k_end;  // sample range, multiples of BK
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
            smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// K-major operand tile, 128-byte rows, SWIZZLE_128B: 8-row atoms of 1024 B (SBO), LBO unused (1), descriptor version 1
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
        "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
}  // namespace i8g

// ---- slicing ---------------------------------------------------------------------------------------------------
// e[r] = exponent with |x| < 2^e for the whole row (0 for an all-zero row)
__global__ void __launch_bounds__(256) i8_row_exponent_kernel(const float *__restrict__ X, int64_t ldx, int64_t m,
                                                              int32_t *__restrict__ e) {
    __shared__ float red[8];
    const float *row = X + (int64_t)blockIdx.x * ldx;
    float mx = 0.f;
    for (int64_t i = threadIdx.x; i < m; i += 256) mx = fmaxf(mx, fabsf(row[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
        int ex = 0;
        if (mx > 0.f && isfinite(mx)) frexpf(mx, &ex);  // mx = f * 2^ex, f in [0.5, 1)  =>  |x| <= mx < 2^ex
        e[blockIdx.x] = ex;
    }
}

// slices: (S, N0p, mp) int8, zero outside (N0, m).  One thread = 16 consecutive samples of one row.
__global__ void __launch_bounds__(256) i8_split_kernel(const float *__restrict__ X, int64_t ldx, int64_t N0, int64_t m,
                                                       const int32_t *__restrict__ e, int8_t *__restrict__ slices,
                                                       int64_t N0p, int64_t mp, int S) {
    const int64_t r = blockIdx.y;
    const int64_t i0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 16;
    if (i0 >= mp) return;
    uint32_t packed[i8g::MAX_S][4];
#pragma unroll
    for (int k = 0; k < i8g::MAX_S; ++k) packed[k][0] = packed[k][1] = packed[k][2] = packed[k][3] = 0u;
    if (r < N0) {
        const double scale = ldexp(1.0, 7 * S - e[r]);
        const float *row = X + r * ldx;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const int64_t i = i0 + c;
            const float x = i < m ? row[i] : 0.f;
            const long long mag = (long long)(fabs((double)x) * scale);  // exact scaling, truncation: mag < 2^(7S)
            const bool neg = x < 0.f;
#pragma unroll
            for (int k = 0; k < i8g::MAX_S; ++k) {
                if (k < S) {
                    int dgt = (int)((mag >> (7 * (S - 1 - k))) & 127);
                    if (neg) dgt = -dgt;
                    packed[k][c >> 2] |= ((uint32_t)(dgt & 0xff)) << (8 * (c & 3));
                }
            }
        }
    }
    for (int k = 0; k < S; ++k) {
        uint4 v = make_uint4(packed[k][0], packed[k][1], packed[k][2], packed[k][3]);
        *reinterpret_cast<uint4 *>(slices + ((int64_t)k * N0p + r) * mp + i0) = v;
    }
}

// ---- the tensor-core kernel ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(i8g::THREADS, 1)
gram_i8_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x,
               const i8g::Item *__restrict__ items, long long *__restrict__ planes_g2, long long *__restrict__ planes_g1,
               int64_t ldp, int64_t N0, int S) {
    using namespace i8g;
    extern __shared__ unsigned char i8_smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)i8_smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)STAGES * STAGE_BYTES);
    uint64_t *empty = full + STAGES;
    uint64_t *acc_ready = empty + STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_ready + 1);

    const Item it = items[blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k_lo = (it.d - S > 1) ? it.d - S : 1, k_hi = (it.d - 1 < S) ? it.d - 1 : S;  // slice pairs (k, d - k)
    const int n_kb = (it.k_end - it.k_begin) / BK;
    const int n_iter = (k_hi - k_lo + 1) * n_kb;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_ready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(TN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const CUtensorMap *map_b = it.gram ? &map_x : &map_q;
            int iter = 0;
            for (int k = k_lo; k <= k_hi; ++k) {
                const int l = it.d - k;
                for (int kb = 0; kb < n_kb; ++kb, ++iter) {
                    const int s = iter % STAGES;
                    if (iter >= STAGES) mbar_wait(&empty[s], ((iter / STAGES) - 1) & 1);
                    unsigned char *a = smem + (size_t)s * STAGE_BYTES, *b = a + A_BYTES;
                    mbar_expect_tx(&full[s], STAGE_BYTES);
                    const int kc = it.k_begin + kb * BK;
                    tma_load_3d(a, &map_q, &full[s], kc, it.ti * TM, k - 1);
                    tma_load_3d(b, map_b, &full[s], kc, it.tj * TN, l - 1);
                    tma_load_3d(b + 128 * BK, map_b, &full[s], kc, it.tj * TN + 128, l - 1);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D = s32, A = B = signed 8-bit, both K-major, N = 256, M = 128
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            for (int iter = 0; iter < n_iter; ++iter) {
                const int s = iter % STAGES;
                mbar_wait(&full[s], (iter / STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t a = smem_u32(smem + (size_t)s * STAGE_BYTES), b = a + A_BYTES;
#pragma unroll
                for (int ks = 0; ks < BK / 32; ++ks)
                    umma_i8(tmem_d, umma_desc_sw128(a + ks * 32), umma_desc_sw128(b + ks * 32), idesc, (iter | ks) ? 1u : 0u);
                umma_commit(&empty[s]);  // the stage may be refilled once these MMAs have read it
            }
            umma_commit(acc_ready);
        }
    } else {
        // ---- epilogue: TMEM lanes 32*(warp % 4) .. +31 are this warp's tile rows
        mbar_wait(acc_ready, 0);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const int quad = warp & 3;
        const int64_t row = (int64_t)it.ti * TM + quad * 32 + lane;
        long long *plane = (it.gram ? planes_g1 : planes_g2) + (int64_t)it.plane * ldp * ldp;
        for (int c = 0; c < TN / 32; ++c) {
            uint32_t v[32];
            tmem_ld32(tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)(c * 32), v);
            const int64_t col0 = (int64_t)it.tj * TN + c * 32;
            if (row < N0) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int64_t col = col0 + i;
                    const int val = (int)v[i];
                    if (col <= row && val != 0)
                        atomicAdd(reinterpret_cast<unsigned long long *>(plane + row * ldp + col),
                                  (unsigned long long)((long long)val << it.shift));
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"(TN));
    }
}

// G[t][s] (s <= t) = (hi * 2^(7 (D - Dhi)) + lo) * 2^(eA_t + eB_s - 7 D)
__global__ void i8_combine_kernel(const long long *__restrict__ planes, int64_t ldp, const int32_t *__restrict__ eA,
                                  const int32_t *__restrict__ eB, int64_t N0, int D, int Dhi, double *__restrict__ G) {
    const int64_t t = blockIdx.y;
    const long long *hi = planes + t * ldp, *lo = planes + ldp * ldp + t * ldp;
    const double up = ldexp(1.0, 7 * (D - Dhi));
    const int et = eA[t] - 7 * D;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s <= t; s += (int64_t)gridDim.x * blockDim.x)
        G[t * N0 + s] = ldexp(fma((double)hi[s], up, (double)lo[s]), et + eB[s]);
}

// ---- host side -------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static int make_slice_map(gpfq_ctx *ctx, CUtensorMap *map, int8_t *slices, int64_t N0p, int64_t mp, int S) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return gpfq_fail(ctx, GPFQ_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3] = {(cuuint64_t)mp, (cuuint64_t)N0p, (cuuint64_t)S};
    const cuuint64_t strides[2] = {(cuuint64_t)mp, (cuuint64_t)mp * (cuuint64_t)N0p};  // bytes, dims 1 and 2
    const cuuint32_t box[3] = {(cuuint32_t)i8g::BK, 128u, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult rc = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, slices, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return gpfq_fail(ctx, GPFQ_ERR_CUDA, "cuTensorMapEncodeTiled failed with code %d", (int)rc);
    return GPFQ_OK;
}

// Device workspace the int8 Gram needs beyond G1/G2 (bytes); the caller may prefer the DMMA path when this is too much.
size_t gram_i8_workspace_bytes(int64_t N0, int64_t m, bool same, int S) {
    const int64_t N0p = ceil_div64(N0, 128) * 128, mp = ceil_div64(m, i8g::BK) * i8g::BK;
    return (size_t)(same ? 1 : 2) * ((size_t)S * N0p * mp + (size_t)2 * N0p * N0p * 8);
}

// G2 = Xq Xq^T and (unless X == Xq) G1 = Xq X^T, lower triangles, fp64 (N0, N0).
int gram_i8_stage(gpfq_ctx *ctx, const float *X, const float *Xq, int64_t ldx, int64_t N0, int64_t m, double *G1, double *G2) {
    using namespace i8g;
    const bool same = (X == Xq);
    const int S = MAX_S, D = S + 1, Dhi = D < 4 ? D : 4;
    if (m >= ((int64_t)1 << 31) - BK || N0 >= ((int64_t)1 << 30))
        return gpfq_fail(ctx, GPFQ_ERR_ARG, "int8 Gram: shape too large");
    const int64_t N0p = ceil_div64(N0, 128) * 128, mp = ceil_div64(m, BK) * BK;
    cudaStream_t st = ctx->stream;
    int8_t *sl_q = nullptr, *sl_x = nullptr;
    int32_t *e_q = nullptr, *e_x = nullptr;
    long long *pl2 = nullptr, *pl1 = nullptr;
    GPFQ_TRY(gpfq_ws(ctx, WS_I8_SQ, (size_t)S * N0p * mp, (void **)&sl_q));
    GPFQ_TRY(gpfq_ws(ctx, WS_I8_E, (size_t)2 * N0p * sizeof(int32_t), (void **)&e_q));
    e_x = e_q + N0p;
    GPFQ_TRY(gpfq_ws(ctx, WS_I8_P2, (size_t)2 * N0p * N0p * sizeof(long long), (void **)&pl2));
    if (!same) {
        GPFQ_TRY(gpfq_ws(ctx, WS_I8_SX, (size_t)S * N0p * mp, (void **)&sl_x));
        GPFQ_TRY(gpfq_ws(ctx, WS_I8_P1, (size_t)2 * N0p * N0p * sizeof(long long), (void **)&pl1));
    } else {
        sl_x = sl_q;
        e_x = e_q;
    }
    // slicing
    CUDA_TRY(ctx, cudaMemsetAsync(e_q, 0, (size_t)2 * N0p * sizeof(int32_t), st));
    dim3 sgrid((unsigned)ceil_div64(mp, 256 * 16), (unsigned)N0p);
    i8_row_exponent_kernel<<<(unsigned)N0, 256, 0, st>>>(Xq, ldx, m, e_q);
    KERNEL_CHECK(ctx);
    i8_split_kernel<<<sgrid, 256, 0, st>>>(Xq, ldx, N0, m, e_q, sl_q, N0p, mp, S);
    KERNEL_CHECK(ctx);
    if (!same) {
        i8_row_exponent_kernel<<<(unsigned)N0, 256, 0, st>>>(X, ldx, m, e_x);
        KERNEL_CHECK(ctx);
        i8_split_kernel<<<sgrid, 256, 0, st>>>(X, ldx, N0, m, e_x, sl_x, N0p, mp, S);
        KERNEL_CHECK(ctx);
    }
    CUDA_TRY(ctx, cudaMemsetAsync(pl2, 0, (size_t)2 * N0p * N0p * sizeof(long long), st));
    if (!same) CUDA_TRY(ctx, cudaMemsetAsync(pl1, 0, (size_t)2 * N0p * N0p * sizeof(long long), st));

    // work items, heaviest first
    std::vector<Item> items;
    const int tiles_m = (int)ceil_div64(N0, TM), tiles_n = (int)ceil_div64(N0, TN);
    for (int d = D; d >= 2; --d) {
        const int k_lo = d - S > 1 ? d - S : 1, k_hi = d - 1 < S ? d - 1 : S, pairs = k_hi - k_lo + 1;
        int64_t kc = ((int64_t)2147483647 / ((int64_t)16129 * pairs)) / BK * BK;  // no s32 overflow: pairs * kc * 127^2 < 2^31
        // enough items to fill the machine: split K further while the tile count is small
        const int64_t tiles = (int64_t)tiles_m * tiles_n * (same ? 1 : 2);
        while (kc > 4 * BK && tiles * ceil_div64(mp, kc) * (D - 1) < 2 * ctx->sm_count) kc = ceil_div64(kc / 2, BK) * BK;
        for (int g = 0; g < (same ? 1 : 2); ++g)
            for (int ti = 0; ti < tiles_m; ++ti)
                for (int tj = 0; tj < tiles_n; ++tj) {
                    if ((int64_t)tj * TN > (int64_t)ti * TM + TM - 1) continue;  // wholly above the diagonal
                    for (int64_t kb = 0; kb < mp; kb += kc) {
                        Item it;
                        it.ti = ti; it.tj = tj; it.d = d; it.gram = g;
                        it.plane = d <= Dhi ? 0 : 1;
                        it.shift = 7 * ((d <= Dhi ? Dhi : D) - d);
                        it.k_begin = (int32_t)kb;
                        it.k_end = (int32_t)(kb + kc < mp ? kb + kc : mp);
                        items.push_back(it);
                    }
                }
    }
    Item *d_items = nullptr;
    GPFQ_TRY(gpfq_ws(ctx, WS_I8_ITEMS, items.size() * sizeof(Item), (void **)&d_items));
    // pageable source: cudaMemcpyAsync stages it before returning
    CUDA_TRY(ctx, cudaMemcpyAsync(d_items, items.data(), items.size() * sizeof(Item), cudaMemcpyHostToDevice, st));
    CUtensorMap map_q, map_x;
    GPFQ_TRY(make_slice_map(ctx, &map_q, sl_q, N0p, mp, S));
    GPFQ_TRY(make_slice_map(ctx, &map_x, sl_x, N0p, mp, S));
    CUDA_TRY(ctx, cudaFuncSetAttribute(gram_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    gram_i8_kernel<<<(unsigned)items.size(), THREADS, SMEM, st>>>(map_q, map_x, d_items, pl2, pl1, N0p, N0, S);
    KERNEL_CHECK(ctx);
    dim3 cgrid((unsigned)ceil_div64(N0, 256 * 4) > 0 ? (unsigned)ceil_div64(N0, 256 * 4) : 1u, (unsigned)N0);
    i8_combine_kernel<<<cgrid, 256, 0, st>>>(pl2, N0p, e_q, e_q, N0, D, Dhi, G2);
    KERNEL_CHECK(ctx);
    if (!same) {
        i8_combine_kernel<<<cgrid, 256, 0, st>>>(pl1, N0p, e_q, e_x, N0, D, Dhi, G1);
        KERNEL_CHECK(ctx);
    }
    return GPFQ_OK;
}
