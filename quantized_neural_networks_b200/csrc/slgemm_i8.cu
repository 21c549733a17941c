// slgemm_i8.cu -- rectangular "NT" contraction  C (+)= sum_seg alpha_seg * A_seg B_seg^T  with fp64-class accuracy on the 5th-generation
// tensor cores: int8 digit slices (Ozaki scheme, as gram_i8.cu) + tcgen05.mma.kind::i8 into TMEM, operands by TMA.
//
// Used by the carried-residual (low-rank) form of the Dense sweep (dense_gram.cu): per range of directions
//     U += W_r X_r - Q_r X~_r          (nj x m, K = the range's directions;  quantized_network.py:119 for a whole range)
//     D_r = U X~_r^T                   (nj x R, K = the m samples;           the residual dots of :86/:89 for a whole range)
// which north_star (b) wants "folded in as small GEMMs" -- here they run on tcgen05 instead of the fp64 DMMA pipe.
//
// Operands are stored as S signed base-256 digit slices per value with one exponent per ROW (row = output row or column,
// K contiguous):   x[r][k] * 2^-e_r = sum_{s=1..S} b_s[r][k] 2^(2 - 8 s)   (b in [-128, 127]; rounding at 2^-(8S-2) of 2^e_r).
// A "phase" is one (segment, d): every slice pair with s_a + s_b = d accumulates EXACTLY into one s32 TMEM accumulator
// (<= 5 pairs x K <= 26112), then the epilogue folds it into the tile's fp64 accumulators with the scale
// alpha 2^(4 - 8 d + eA_i + eB_j).  Unlike gram_i8.cu (K = all samples, few phases per byte of output) these products have a
// short K, so a read-modify-write of the fp64 tile in global memory per phase would dominate: the fp64 accumulators live in
// TENSOR MEMORY instead (tcgen05.ld / tcgen05.st by the epilogue warps; 128 x 128 tile = 256 columns of lo / hi words, next
// to two 128-column s32 MMA accumulators: all 512 columns), and C is touched once per tile.
//
// Kernel anatomy (one CTA per 128 x 128 output tile, one CTA per SM):
//   warp 0    TMA producer: 128 x 128 B boxes of one A slice and one B slice per stage (128B swizzle), 6-stage mbarrier ring
//   warp 1    TMEM allocator + single-thread MMA issuer: 4 x tcgen05.mma.cta_group::1.kind::i8 (M128 N128 K32) per stage
//   warps 2-5 epilogue: per phase tcgen05.ld (s32) -> fp64 scale -> fma into the TMEM-resident fp64 tile (tcgen05.ld/st);
//             after the last phase the tile is written (or added) to C, one row per thread
#include <algorithm>

#include "slgemm_i8.cuh"

namespace slg {
constexpr int TM = 128, TN = 128, BK = 128;
constexpr int STAGES = 6;
constexpr int A_BYTES = TM * BK, B_BYTES = TN * BK, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int THREADS = 192;
constexpr int S = 5;                 // digit slices per sliced value
constexpr int P_BITS = 8 * S - 2;
constexpr int KB_MAX = 204;          // K blocks per call: 5 pairs * 204 * 128 * 128^2 < 2^31
constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */ + 2 * TN * sizeof(double);
}  // namespace slg

struct SlSeg {
    int SA, SB, D;          // slices of A / B; slice pairs with s_a + s_b <= D
    int a_row0, b_row0;     // first row of this product inside the A / B slice tensors (the tile offset is added)
    int k0, kblocks;        // first K byte (both operands) and K blocks of 128 bytes
    double alpha;
    const int32_t *eA;      // row exponents, indexed like the slice rows (nullptr: eA_const)
    const int32_t *eB;
    int eA_const, eB_const;
};

struct SlArgs {
    SlSeg seg[2];
    int nseg, tiles_n;
    double *C;
    int64_t ldc;
    int M, N;               // valid rows / columns of C
    int accumulate, vec2;
};

__global__ void __launch_bounds__(slg::THREADS, 1)
slgemm_i8_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapB0,
                 const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapB1, const SlArgs args) {
    using namespace i8g;
    using namespace slg;
    extern __shared__ unsigned char sl_smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)sl_smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)STAGES * STAGE_BYTES);
    uint64_t *empty = full + STAGES;
    uint64_t *acc_full = empty + STAGES;    // [2]
    uint64_t *acc_empty = acc_full + 2;     // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
    double *colscale = reinterpret_cast<double *>(smem + (size_t)STAGES * STAGE_BYTES + 256);   // [2][TN]: 2^eB of the tile's columns

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ti = blockIdx.x / args.tiles_n, tj = blockIdx.x % args.tiles_n;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (warp >= 2) {
        const int c = threadIdx.x - 64;     // 0..127: one column of the tile per epilogue thread
        for (int s = 0; s < args.nseg; ++s) {
            const SlSeg &g = args.seg[s];
            const int e = g.eB ? g.eB[g.b_row0 + tj * TN + c] : g.eB_const;
            colscale[s * TN + c] = ldexp(1.0, e);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---- TMA producer
        if (lane == 0) {
            int iter = 0;
            for (int sg = 0; sg < args.nseg; ++sg) {
                const SlSeg &g = args.seg[sg];
                const CUtensorMap *ma = sg ? &mapA1 : &mapA0, *mb = sg ? &mapB1 : &mapB0;
                for (int d = 2; d <= g.D; ++d) {
                    const int k_lo = max(1, d - g.SB), k_hi = min(g.SA, d - 1);
                    for (int k = k_lo; k <= k_hi; ++k) {
                        const int l = d - k;
                        for (int kb = 0; kb < g.kblocks; ++kb, ++iter) {
                            const int s = iter % STAGES;
                            if (iter >= STAGES) mbar_wait(&empty[s], ((iter / STAGES) - 1) & 1);
                            unsigned char *a = smem + (size_t)s * STAGE_BYTES, *b = a + A_BYTES;
                            mbar_expect_tx(&full[s], STAGE_BYTES);
                            tma_load_3d(a, ma, &full[s], g.k0 + kb * BK, g.a_row0 + ti * TM, k - 1);
                            tma_load_3d(b, mb, &full[s], g.k0 + kb * BK, g.b_row0 + tj * TN, l - 1);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer
        if (lane == 0) {
            // instruction descriptor: D = s32, A = B = signed 8-bit, both K-major, N = 128, M = 128
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            int iter = 0, p = 0;
            for (int sg = 0; sg < args.nseg; ++sg) {
                const SlSeg &g = args.seg[sg];
                for (int d = 2; d <= g.D; ++d, ++p) {
                    const int buf = p & 1;
                    if (p >= 2) {  // the epilogue must have drained this accumulator
                        mbar_wait(&acc_empty[buf], ((p >> 1) - 1) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    }
                    const uint32_t tmem_d = tmem_base + (uint32_t)(buf * TN);
                    const int n_it = (min(g.SA, d - 1) - max(1, d - g.SB) + 1) * g.kblocks;
                    for (int it = 0; it < n_it; ++it, ++iter) {
                        const int s = iter % STAGES;
                        mbar_wait(&full[s], (iter / STAGES) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                        const uint32_t a = smem_u32(smem + (size_t)s * STAGE_BYTES), b = a + A_BYTES;
#pragma unroll
                        for (int ks = 0; ks < BK / 32; ++ks)
                            umma_i8(tmem_d, umma_desc_sw128(a + ks * 32), umma_desc_sw128(b + ks * 32), idesc, (it | ks) ? 1u : 0u);
                        umma_commit(&empty[s]);  // the stage may be refilled once these MMAs have read it
                    }
                    umma_commit(&acc_full[buf]);
                }
            }
        }
    } else {
        // ---- epilogue: TMEM lanes 32 * (warp % 4) .. + 31 are this warp's tile rows; one thread = one row.
        const int quad = warp & 3;
        const int row_t = quad * 32 + lane;                          // row inside the tile
        const uint32_t lane_field = (uint32_t)(quad * 32) << 16;
        const uint32_t acc64 = tmem_base + lane_field + 2 * TN;      // fp64 tile: column c at words 2 c (lo), 2 c + 1 (hi)
        int p = 0;
        for (int sg = 0; sg < args.nseg; ++sg) {
            const SlSeg &g = args.seg[sg];
            const int ea = g.eA ? g.eA[g.a_row0 + ti * TM + row_t] : g.eA_const;
            const double *cs = colscale + sg * TN;
            for (int d = 2; d <= g.D; ++d, ++p) {
                const int buf = p & 1;
                mbar_wait(&acc_full[buf], (p >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const double rs = ldexp(g.alpha, 4 - 8 * d + ea);
#pragma unroll 1
                for (int cc = 0; cc < TN / 32; ++cc) {
                    uint32_t v[32], a[64];
                    tmem_ld32(tmem_base + lane_field + (uint32_t)(buf * TN + cc * 32), v);
                    if (p > 0) tmem_ld64(acc64 + cc * 64, a);
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const double prev = p > 0 ? __hiloint2double((int)a[2 * i + 1], (int)a[2 * i]) : 0.0;
                        const double r = fma((double)(int)v[i] * cs[cc * 32 + i], rs, prev);
                        a[2 * i] = (uint32_t)__double2loint(r);
                        a[2 * i + 1] = (uint32_t)__double2hiint(r);
                    }
                    tmem_st64(acc64 + cc * 64, a);
                }
                asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[buf]);
            }
        }
        // ---- the finished tile: one row per thread, 32 consecutive doubles per chunk
        const int64_t row = (int64_t)ti * TM + row_t;
#pragma unroll 1
        for (int cc = 0; cc < TN / 32; ++cc) {
            uint32_t a[64];
            tmem_ld64(acc64 + cc * 64, a);
            const int64_t col0 = (int64_t)tj * TN + cc * 32;
            if (row < args.M && col0 < args.N) {
                double *o = args.C + row * args.ldc + col0;
                if (args.vec2 && col0 + 32 <= args.N) {
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        double2 r = make_double2(__hiloint2double((int)a[2 * i + 1], (int)a[2 * i]),
                                                 __hiloint2double((int)a[2 * i + 3], (int)a[2 * i + 2]));
                        if (args.accumulate) {
                            const double2 old = *reinterpret_cast<const double2 *>(o + i);
                            r.x += old.x;
                            r.y += old.y;
                        }
                        *reinterpret_cast<double2 *>(o + i) = r;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (col0 + i < args.N) {
                            const double r = __hiloint2double((int)a[2 * i + 1], (int)a[2 * i]);
                            o[i] = args.accumulate ? o[i] + r : r;
                        }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512));
    }
}

// ---- slicing kernels ---------------------------------------------------------------------------------------------
// Row-wise: one CTA per (padded) row of a row-major matrix (fp32 or fp64 values): exponent of the row maximum, then the S
// digit slices of every value, 16 consecutive K positions per thread and step.  Rows >= rows and columns >= cols are zeros.
template <typename T>
__global__ void __launch_bounds__(256) sl_rowsplit_kernel(const T *__restrict__ X, int64_t ldx, int64_t rows, int64_t cols,
                                                          int32_t *__restrict__ e, int8_t *__restrict__ slices, int64_t rowsP,
                                                          int64_t colsP) {
    using namespace slg;
    __shared__ double red[8];
    __shared__ int ex_s;
    const int64_t r = blockIdx.x;
    const T *row = X + r * ldx;
    double mx = 0.0;
    if (r < rows)
        for (int64_t i = threadIdx.x; i < cols; i += 256) mx = fmax(mx, fabs((double)row[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) mx = fmax(mx, red[w]);
        int ex = 0;
        if (mx > 0.0 && isfinite(mx)) frexp(mx, &ex);   // mx = f 2^ex, f in [0.5, 1)  =>  |x| < 2^ex
        e[r] = ex;
        ex_s = ex;
    }
    __syncthreads();
    const int ex = ex_s;
    for (int64_t i0 = (int64_t)threadIdx.x * 16; i0 < colsP; i0 += 256 * 16) {
        uint32_t packed[S][4];
#pragma unroll
        for (int k = 0; k < S; ++k) packed[k][0] = packed[k][1] = packed[k][2] = packed[k][3] = 0u;
        if (r < rows && i0 < cols) {
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const double x = (i0 + c < cols) ? (double)row[i0 + c] : 0.0;
                long long v = __double2ll_rn(ldexp(x, P_BITS - ex));   // |v| <= 2^38
#pragma unroll
                for (int k = S - 1; k >= 0; --k) {
                    const int dg = (int)((v + 128) & 255) - 128;       // balanced digit in [-128, 127]
                    v = (v - dg) >> 8;
                    packed[k][c >> 2] |= ((uint32_t)(dg & 0xff)) << (8 * (c & 3));
                }
            }
        }
#pragma unroll
        for (int k = 0; k < S; ++k)
            *reinterpret_cast<uint4 *>(slices + ((int64_t)k * rowsP + r) * colsP + i0) =
                make_uint4(packed[k][0], packed[k][1], packed[k][2], packed[k][3]);
    }
}

// Column maxima of |X| (X: (N0, m) fp32): non-negative floats order like their bit patterns, so an integer atomicMax does it.
__global__ void __launch_bounds__(256) sl_colmax_kernel(const float *__restrict__ X, int64_t ldx, int64_t N0, int64_t m,
                                                        int *__restrict__ colmax, int rows_per_cta) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= m) return;
    const int64_t t0 = (int64_t)blockIdx.y * rows_per_cta, t1 = min(N0, t0 + rows_per_cta);
    float mx = 0.f;
    for (int64_t t = t0; t < t1; ++t) mx = fmaxf(mx, fabsf(X[t * ldx + i]));
    atomicMax(colmax + i, __float_as_int(mx));
}

__global__ void sl_exp_from_max_kernel(const int *__restrict__ colmax, int64_t n, int64_t nP, int32_t *__restrict__ e) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nP) return;
    int ex = 0;
    if (i < n) {
        const float mx = __int_as_float(colmax[i]);
        if (mx > 0.f && isfinite(mx)) frexpf(mx, &ex);
    }
    e[i] = ex;
}

// Transposed slicing: X (N0, m) fp32 -> slices (S, mP, N0P) of X^T (row = sample, K = direction), exponent per sample.
// Tile of 64 directions x 32 samples through shared memory; a thread then owns one sample and 16 consecutive directions.
__global__ void __launch_bounds__(128) sl_transsplit_kernel(const float *__restrict__ X, int64_t ldx, int64_t N0, int64_t m,
                                                            const int32_t *__restrict__ e, int8_t *__restrict__ slices, int64_t mP,
                                                            int64_t N0P) {
    using namespace slg;
    __shared__ float tile[64][33];
    const int64_t t0 = (int64_t)blockIdx.x * 64, i0 = (int64_t)blockIdx.y * 32;
    for (int idx = threadIdx.x; idx < 64 * 32; idx += 128) {
        const int tr = idx >> 5, ic = idx & 31;
        const int64_t t = t0 + tr, i = i0 + ic;
        tile[tr][ic] = (t < N0 && i < m) ? X[t * ldx + i] : 0.f;
    }
    __syncthreads();
    const int ic = threadIdx.x & 31, tg = threadIdx.x >> 5;   // sample, group of 16 directions
    const int64_t i = i0 + ic;
    if (i >= mP) return;
    const int ex = e[i];
    uint32_t packed[S][4];
#pragma unroll
    for (int k = 0; k < S; ++k) packed[k][0] = packed[k][1] = packed[k][2] = packed[k][3] = 0u;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        long long v = __double2ll_rn(ldexp((double)tile[tg * 16 + c][ic], P_BITS - ex));
#pragma unroll
        for (int k = S - 1; k >= 0; --k) {
            const int dg = (int)((v + 128) & 255) - 128;
            v = (v - dg) >> 8;
            packed[k][c >> 2] |= ((uint32_t)(dg & 0xff)) << (8 * (c & 3));
        }
    }
#pragma unroll
    for (int k = 0; k < S; ++k)
        *reinterpret_cast<uint4 *>(slices + ((int64_t)k * mP + i) * N0P + t0 + tg * 16) =
            make_uint4(packed[k][0], packed[k][1], packed[k][2], packed[k][3]);
}

// Level indices of a range of decisions: q = h k', k' in [-(K-1), K-1] (symmetric equispaced alphabet; the literal 0 of a dead
// direction is k' = 0)  ->  one int8 "slice" (rows, N0P) of the whole decision matrix, byte columns [tb, tb + width) written
// (zeros for neurons >= nj and directions >= te).
__global__ void __launch_bounds__(256) sl_qindex_kernel(const double *__restrict__ Qt, int64_t N0, int64_t nj, int64_t tb, int64_t te,
                                                        double inv_h, int8_t *__restrict__ out, int64_t grid_rows, int64_t N0P,
                                                        int64_t width) {
    const int64_t per_row = width / 16;
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= grid_rows * per_row) return;
    const int64_t j = idx / per_row, c0 = (idx % per_row) * 16;
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    if (j < nj) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const int64_t t = tb + c0 + c;
            const int kq = t < te ? __double2int_rn(Qt[j * N0 + t] * inv_h) : 0;
            w[c >> 2] |= ((uint32_t)(kq & 0xff)) << (8 * (c & 3));
        }
    }
    *reinterpret_cast<uint4 *>(out + j * N0P + tb + c0) = make_uint4(w[0], w[1], w[2], w[3]);
}

// ---- host side -------------------------------------------------------------------------------------------------
int sl_make_operand(gpfq_ctx *ctx, SlOperand *op, const int8_t *slices, int64_t rowsP, int64_t kbytes, int n_slices, const int32_t *e,
                    int e_const) {
    op->slices = slices;
    op->rowsP = rowsP;
    op->kbytes = kbytes;
    op->n_slices = n_slices;
    op->e = e;
    op->e_const = e_const;
    return make_i8_slice_map(ctx, &op->map, slices, rowsP, kbytes, n_slices);
}

// C[M x N] (ldc) = or += sum of the products.  M, N: valid extents; the slice tensors are zero-padded to tile multiples.
int slgemm_i8(gpfq_ctx *ctx, const SlProduct *prod, int nprod, double *C, int64_t ldc, int64_t M, int64_t N, bool accumulate) {
    using namespace slg;
    if (nprod < 1 || nprod > 2) return gpfq_fail(ctx, GPFQ_ERR_ARG, "slgemm_i8: 1 or 2 products");
    SlArgs a = {};
    a.nseg = nprod;
    for (int s = 0; s < nprod; ++s) {
        const SlProduct &p = prod[s];
        if (p.K % BK || p.k0 % 16 || p.K / BK > KB_MAX || p.K < BK)
            return gpfq_fail(ctx, GPFQ_ERR_ARG, "slgemm_i8: K = %lld (offset %lld) must be a multiple of 128 (16), at most %d",
                             (long long)p.K, (long long)p.k0, KB_MAX * BK);
        if (p.a_row0 + ceil_div64(M, TM) * TM > p.A->rowsP || p.b_row0 + ceil_div64(N, TN) * TN > p.B->rowsP ||
            p.k0 + p.K > p.A->kbytes || p.k0 + p.K > p.B->kbytes)
            return gpfq_fail(ctx, GPFQ_ERR_ARG, "slgemm_i8: operand slices are not padded to the tile grid");
        SlSeg &g = a.seg[s];
        g.SA = p.A->n_slices;
        g.SB = p.B->n_slices;
        g.D = std::min(p.D, g.SA + g.SB);
        g.a_row0 = (int)p.a_row0;
        g.b_row0 = (int)p.b_row0;
        g.k0 = (int)p.k0;
        g.kblocks = (int)(p.K / BK);
        g.alpha = p.alpha;
        g.eA = p.A->e;
        g.eB = p.B->e;
        g.eA_const = p.A->e_const;
        g.eB_const = p.B->e_const;
    }
    a.tiles_n = (int)ceil_div64(N, TN);
    a.C = C;
    a.ldc = ldc;
    a.M = (int)M;
    a.N = (int)N;
    a.accumulate = accumulate ? 1 : 0;
    a.vec2 = (ldc % 2 == 0) && ((uintptr_t)C % 16 == 0);
    CUDA_TRY(ctx, cudaFuncSetAttribute(slgemm_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    const unsigned grid = (unsigned)(ceil_div64(M, TM) * a.tiles_n);
    const SlProduct &p1 = prod[nprod - 1];
    slgemm_i8_kernel<<<grid, THREADS, SMEM, ctx->stream>>>(prod[0].A->map, prod[0].B->map, p1.A->map, p1.B->map, a);
    KERNEL_CHECK(ctx);
    return GPFQ_OK;
}

// slices of a row-major fp64 / fp32 matrix (rows x cols, row = output row, K = cols); e: rowsP exponents
template <typename T>
int sl_rowsplit(gpfq_ctx *ctx, const T *X, int64_t ldx, int64_t rows, int64_t cols, int32_t *e, int8_t *slices, int64_t rowsP,
                int64_t colsP, int64_t grid_rows) {
    sl_rowsplit_kernel<T><<<(unsigned)grid_rows, 256, 0, ctx->stream>>>(X, ldx, rows, cols, e, slices, rowsP, colsP);
    KERNEL_CHECK(ctx);
    return GPFQ_OK;
}
template int sl_rowsplit<double>(gpfq_ctx *, const double *, int64_t, int64_t, int64_t, int32_t *, int8_t *, int64_t, int64_t, int64_t);
template int sl_rowsplit<float>(gpfq_ctx *, const float *, int64_t, int64_t, int64_t, int32_t *, int8_t *, int64_t, int64_t, int64_t);

// slices of X^T for X (N0, m) fp32: (S, mP, N0P), exponent per sample; scratch: mP ints
int sl_transsplit(gpfq_ctx *ctx, const float *X, int64_t ldx, int64_t N0, int64_t m, int32_t *e, int *scratch, int8_t *slices,
                  int64_t mP, int64_t N0P) {
    cudaStream_t st = ctx->stream;
    CUDA_TRY(ctx, cudaMemsetAsync(scratch, 0, (size_t)mP * sizeof(int), st));
    const int rows_per_cta = 256;
    dim3 g1((unsigned)ceil_div64(m, 256), (unsigned)ceil_div64(N0, rows_per_cta));
    sl_colmax_kernel<<<g1, 256, 0, st>>>(X, ldx, N0, m, scratch, rows_per_cta);
    KERNEL_CHECK(ctx);
    sl_exp_from_max_kernel<<<(unsigned)ceil_div64(mP, 256), 256, 0, st>>>(scratch, m, mP, e);
    KERNEL_CHECK(ctx);
    dim3 g2((unsigned)(N0P / 64), (unsigned)(mP / 32));
    sl_transsplit_kernel<<<g2, 128, 0, st>>>(X, ldx, N0, m, e, slices, mP, N0P);
    KERNEL_CHECK(ctx);
    return GPFQ_OK;
}

int sl_qindex(gpfq_ctx *ctx, const double *Qt, int64_t N0, int64_t nj, int64_t tb, int64_t te, double inv_h, int8_t *out,
              int64_t grid_rows, int64_t N0P, int64_t width) {
    const int64_t n = grid_rows * (width / 16);
    sl_qindex_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, ctx->stream>>>(Qt, N0, nj, tb, te, inv_h, out, grid_rows, N0P, width);
    KERNEL_CHECK(ctx);
    return GPFQ_OK;
}

// ---- diagnostics entry point (include/gpfq.h): C = A B^T through the int8-slice kernel, host arrays ------------------------
extern "C" int gpfq_debug_slgemm(gpfq_ctx *ctx, const double *A, const float *B, int64_t M, int64_t N, int64_t K, int32_t D,
                                 int32_t transposed_b, double *C_out) {
    // A: (M, K) fp64 row-major.  B: (N, K) fp32 row-major, or with transposed_b (K, N) fp32 row-major -- the X^T slicing path.
    using namespace slg;
    if (!ctx) return GPFQ_ERR_ARG;
    ctx->err.clear();
    if (!A || !B || !C_out || M < 1 || N < 1 || K < 1) return gpfq_fail(ctx, GPFQ_ERR_ARG, "bad arguments");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t MP = ceil_div64(M, TM) * TM, NP = ceil_div64(N, TN) * TN, KP = ceil_div64(K, BK) * BK;
    double *dA = nullptr, *dC = nullptr;
    float *dB = nullptr;
    int8_t *sA = nullptr, *sB = nullptr;
    int32_t *e = nullptr;
    GPFQ_TRY(gpfq_ws(ctx, WS_G1, (size_t)M * K * sizeof(double), (void **)&dA));
    GPFQ_TRY(gpfq_ws(ctx, WS_X, (size_t)N * K * sizeof(float), (void **)&dB));
    GPFQ_TRY(gpfq_ws(ctx, WS_G2, (size_t)M * N * sizeof(double), (void **)&dC));
    GPFQ_TRY(gpfq_ws(ctx, WS_I8_SQ, (size_t)S * MP * KP, (void **)&sA));
    GPFQ_TRY(gpfq_ws(ctx, WS_I8_SX, (size_t)S * NP * KP, (void **)&sB));
    GPFQ_TRY(gpfq_ws(ctx, WS_I8_E, (size_t)(MP + 2 * NP + 8) * sizeof(int32_t), (void **)&e));
    CUDA_TRY(ctx, cudaMemcpyAsync(dA, A, (size_t)M * K * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(dB, B, (size_t)N * K * sizeof(float), cudaMemcpyHostToDevice, st));
    GPFQ_TRY(sl_rowsplit<double>(ctx, dA, K, M, K, e, sA, MP, KP, MP));
    if (transposed_b) GPFQ_TRY(sl_transsplit(ctx, dB, N, K, N, e + MP, reinterpret_cast<int *>(e + MP + NP), sB, NP, KP));
    else GPFQ_TRY(sl_rowsplit<float>(ctx, dB, K, N, K, e + MP, sB, NP, KP, NP));
    SlOperand oa, ob;
    GPFQ_TRY(sl_make_operand(ctx, &oa, sA, MP, KP, S, e, 0));
    GPFQ_TRY(sl_make_operand(ctx, &ob, sB, NP, KP, S, e + MP, 0));
    int64_t done = 0;
    bool first = true;
    while (done < KP) {     // K chunks of at most KB_MAX blocks: the s32 accumulators cannot overflow
        const int64_t kc = std::min<int64_t>(KP - done, (int64_t)KB_MAX * BK);
        SlProduct p = {&oa, &ob, 0, 0, done, kc, D, 1.0};
        GPFQ_TRY(slgemm_i8(ctx, &p, 1, dC, N, M, N, !first));
        first = false;
        done += kc;
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(C_out, dC, (size_t)M * N * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return GPFQ_OK;
}
