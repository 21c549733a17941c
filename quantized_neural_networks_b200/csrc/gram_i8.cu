// gram_i8.cu -- Dense Gram stage on the 5th-generation tensor cores: error-free int8 slicing (Ozaki scheme) +
// tcgen05.mma.kind::i8 with s32 accumulators in TMEM, operands staged by TMA.
//
// The Gram stage G1 = Xq X^T, G2 = Xq Xq^T (the m-length dots and norms of quantized_network.py:83-89 in Gram form,
// SURVEY.md 7.2) is the only dense contraction of the path, but its accuracy requirement (<= 1e-9 relative, SURVEY.md
// App. C "Gram noise sweep") rules out single-pass TF32/BF16 tensor-core products, and tcgen05 has no fp64 kind.  So
// every fp32 row is written as S = 5 signed-digit slices in base 256,
//     x[r][i] * 2^-e_r = sum_{k=1..S} b_k[r][i] * 2^(2 - 8k) + rounding,   b_k in [-128, 127] (int8),
// with e_r the exponent of the row maximum (|x| < 2^e_r); the rounding is at 2^-38 relative to 2^e_r.  Then
//     G[t][s] = 2^(eA_t + eB_s) * sum_{d=2..D} 2^(4 - 8d) * sum_{k+l=d} <b_k(A_t), b_l(B_s)>
// where every inner sum is an EXACT integer from int8 tensor-core MMAs into 32-bit accumulators (K chunks are sized so
// that no accumulator can overflow) and only the last step -- scaling by a power of two and adding at most a few
// dozen terms -- is done in fp64, in a fixed order (bit-reproducible).  D = 6 keeps 15 of the 25 slice pairs; the
// dropped ones are below 2^-36 of 2^(eA+eB) per sample and of random sign (measured: 2e-11 of |Xq||X|^T on ReLU
// activations, 3e-13 with D = 7, 1e-13 with every pair; the parity budget is 1e-9).  Rows whose non-zero entries are
// tiny against the row maximum amplify the dropped terms by (max / mean)^2, so the slicing kernel measures max / mean
// per row and raises D for the whole call to 7 or to 2S (all pairs: only the 2^-38 slice rounding remains) -- on the
// device, without a host round trip.
//
// Kernel anatomy (one CTA per work item = (Gram, 128 x 256 output tile, K split); one CTA per SM):
//   warp 0    TMA producer: cp.async.bulk.tensor.3d (UTMALDG) of a 128 x 128 B A tile and two 128 x 128 B B tiles per
//             stage, 128B-swizzled, 4-stage mbarrier ring
//   warp 1    TMEM allocator + single-thread MMA issuer: 4 x tcgen05.mma.cta_group::1.kind::i8 (M128 N256 K32) per
//             stage into one of two 256-column accumulators; tcgen05.commit frees the stage / publishes the accumulator
//   warps 2-5 epilogue: tcgen05.ld 32x32b.x32 -> scale by 2^(4-8d) -> add into the CTA's own fp64 output tile (the CTA
//             is the only writer, so plain loads/stores in a fixed order), overlapped with the next phase's MMAs
// A "phase" is one (K chunk, d): all slice pairs with k + l = d accumulate into the same TMEM accumulator.
#include <algorithm>

#include "i8_common.cuh"

namespace i8g {
constexpr int TM = 128, TN = 256, BK = 128;   // output tile, K bytes (= samples) per stage
constexpr int STAGES = 4;
constexpr int A_BYTES = TM * BK, B_BYTES = TN * BK, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int THREADS = 192;
constexpr int S = 5;                          // slices per value
constexpr int P_BITS = 8 * S - 2;             // bits kept below 2^e
constexpr int D_NARROW = S + 1;               // slice pairs with k + l <= D: 15 of 25 (rows with max / mean <= 12)
constexpr int D_MEDIUM = S + 2;               // 19 of 25 (max / mean <= 128)
constexpr int D_ALL = 2 * S;                  // every pair
constexpr int KC_BLOCKS = 204;                // K blocks per chunk: 5 pairs * 204 * 128 * 128^2 < 2^31
constexpr int EPI_PITCH = 33;                 // words: per-warp 32 x 32 transposing tile of the epilogue
constexpr size_t EPI_BYTES = (size_t)4 * 32 * EPI_PITCH * 4;
constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */ + EPI_BYTES;
constexpr double RATIO_NARROW = 12.0, RATIO_MEDIUM = 128.0;  // 2^e / mean|non-zero x| of the worst row

}  // namespace i8g

// ---- slicing ---------------------------------------------------------------------------------------------------
// e[r] = exponent with |x| < 2^e for the whole row (0 for an all-zero row).  wide[0] = 0 / 1 / 2 is raised when the
// non-zero entries of some row are so small against 2^e that more slice pairs must be kept (see the header).
__global__ void __launch_bounds__(256) i8_row_exponent_kernel(const float *__restrict__ X, int64_t ldx, int64_t m,
                                                              int32_t *__restrict__ e, int *__restrict__ wide) {
    __shared__ float red_mx[8];
    __shared__ double red_s[8];
    __shared__ unsigned red_n[8];
    const float *row = X + (int64_t)blockIdx.x * ldx;
    float mx = 0.f;
    double sum = 0.0;
    unsigned nnz = 0;
    for (int64_t i = threadIdx.x; i < m; i += 256) {
        const float a = fabsf(row[i]);
        mx = fmaxf(mx, a);
        sum += (double)a;
        nnz += a != 0.f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        nnz += __shfl_xor_sync(0xffffffffu, nnz, o);
    }
    if ((threadIdx.x & 31) == 0) { red_mx[threadIdx.x >> 5] = mx; red_s[threadIdx.x >> 5] = sum; red_n[threadIdx.x >> 5] = nnz; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { mx = fmaxf(mx, red_mx[w]); sum += red_s[w]; nnz += red_n[w]; }
        int ex = 0;
        if (mx > 0.f && isfinite(mx)) frexpf(mx, &ex);  // mx = f * 2^ex, f in [0.5, 1)  =>  |x| <= mx < 2^ex
        if (!isfinite(mx) || !isfinite(sum)) atomicMax(wide, 3);  // Inf / NaN in the inputs: the digit slices would be garbage
        e[blockIdx.x] = ex;
        if (nnz > 0) {
            const double top = ldexp(1.0, ex) * (double)nnz;
            if (!(top <= i8g::RATIO_MEDIUM * sum)) atomicMax(wide, 2);
            else if (!(top <= i8g::RATIO_NARROW * sum)) atomicMax(wide, 1);
        }
    }
}

// slices: (S, N0p, mp) int8, zero outside (N0, m).  One thread = 16 consecutive samples of one row.
__global__ void __launch_bounds__(256) i8_split_kernel(const float *__restrict__ X, int64_t ldx, int64_t N0, int64_t m,
                                                       const int32_t *__restrict__ e, int8_t *__restrict__ slices,
                                                       int64_t N0p, int64_t mp, int vec) {
    using namespace i8g;
    const int64_t r = blockIdx.y;
    const int64_t i0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 16;
    if (i0 >= mp) return;
    uint32_t packed[S][4];
#pragma unroll
    for (int k = 0; k < S; ++k) packed[k][0] = packed[k][1] = packed[k][2] = packed[k][3] = 0u;
    if (r < N0 && i0 < m) {
        const double scale = pow2(P_BITS - e[r]);
        const float *row = X + r * ldx;
        float xs[16];
        if (vec && i0 + 16 <= m) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 v4 = __ldg(reinterpret_cast<const float4 *>(row + i0) + q);
                xs[4 * q] = v4.x; xs[4 * q + 1] = v4.y; xs[4 * q + 2] = v4.z; xs[4 * q + 3] = v4.w;
            }
        } else {
#pragma unroll
            for (int c = 0; c < 16; ++c) xs[c] = (i0 + c < m) ? __ldg(row + i0 + c) : 0.f;
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            long long v = __double2ll_rn((double)xs[c] * scale);  // exact scaling, |v| < 2^38
#pragma unroll
            for (int k = S - 1; k >= 0; --k) {
                const int dg = (int)((v + 128) & 255) - 128;     // balanced digit in [-128, 127]
                v = (v - dg) >> 8;
                packed[k][c >> 2] |= ((uint32_t)(dg & 0xff)) << (8 * (c & 3));
            }
        }
    }
#pragma unroll
    for (int k = 0; k < S; ++k)
        *reinterpret_cast<uint4 *>(slices + ((int64_t)k * N0p + r) * mp + i0) = make_uint4(packed[k][0], packed[k][1], packed[k][2], packed[k][3]);
}

// ---- the tensor-core kernel ------------------------------------------------------------------------------------
// Output tiles in bands of eight tile rows, walked column by column inside a band, so that the CTAs of a wave share A
// and B tiles in L2.  Tile (ti, tj) exists when tj <= ti / 2 (128-row x 256-column tiles touching the lower triangle).
__device__ __forceinline__ void i8_decode_tile(int idx, int tiles_m, int &ti, int &tj) {
    int b0 = 0;
    for (;;) {
        const int b1 = min(b0 + 8, tiles_m);
        int cnt = 0;
        for (int t = b0; t < b1; ++t) cnt += (t >> 1) + 1;
        if (idx < cnt || b1 >= tiles_m) {
            for (int j = 0;; ++j) {
                const int lo = max(b0, 2 * j), n = b1 - lo;
                if (idx < n || n <= 0) { ti = lo + (n > 0 ? idx : 0); tj = j; return; }
                idx -= n;
            }
        }
        idx -= cnt;
        b0 = b1;
    }
}

struct I8Args {
    int tiles_m;             // tile rows (128 directions each)
    int nsplit, ngram;       // K splits per tile, Grams (1: G2 only, 2: G2 then G1)
    int kb_total, kb_per;    // K blocks of BK samples in all / per split
    double *out[2];          // per Gram: nsplit partial matrices (N0 x N0, row stride N0), split_stride apart
    int64_t split_stride;
    int64_t N0;
    const int *wide;         // device level from the slicing kernels: 0 narrow rows, 1 medium, 2 keep every slice pair
    int d_force;             // > 0: use this D whatever the level (tests)
};

__global__ void __launch_bounds__(i8g::THREADS, 1)
gram_i8_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x, const I8Args args) {
    using namespace i8g;
    extern __shared__ unsigned char i8_smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)i8_smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)STAGES * STAGE_BYTES);
    uint64_t *empty = full + STAGES;
    uint64_t *acc_full = empty + STAGES;    // [2]
    uint64_t *acc_empty = acc_full + 2;     // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
    int *epi_tiles = reinterpret_cast<int *>(smem + (size_t)STAGES * STAGE_BYTES + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gram = (int)(blockIdx.x % args.ngram);
    const int rest = (int)(blockIdx.x / args.ngram);
    const int split = rest % args.nsplit;
    int ti, tj;
    i8_decode_tile(rest / args.nsplit, args.tiles_m, ti, tj);
    const int kb0 = split * args.kb_per;
    const int kb1 = min(kb0 + args.kb_per, args.kb_total);
    const int level = *args.wide;
    const int D = args.d_force > 0 ? args.d_force : (level >= 2 ? D_ALL : (level == 1 ? D_MEDIUM : D_NARROW));
    const int n_chunks = (kb1 - kb0 + KC_BLOCKS - 1) / KC_BLOCKS;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(2 * TN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---- TMA producer
        if (lane == 0) {
            const CUtensorMap *map_b = gram ? &map_x : &map_q;
            int iter = 0;
            for (int c = 0; c < n_chunks; ++c) {
                const int ck0 = kb0 + c * KC_BLOCKS, ck1 = min(ck0 + KC_BLOCKS, kb1);
                for (int d = 2; d <= D; ++d) {
                    const int k_lo = max(1, d - S), k_hi = min(S, d - 1);
                    for (int k = k_lo; k <= k_hi; ++k) {
                        const int l = d - k;
                        for (int kb = ck0; kb < ck1; ++kb, ++iter) {
                            const int s = iter % STAGES;
                            if (iter >= STAGES) mbar_wait(&empty[s], ((iter / STAGES) - 1) & 1);
                            unsigned char *a = smem + (size_t)s * STAGE_BYTES, *b = a + A_BYTES;
                            mbar_expect_tx(&full[s], STAGE_BYTES);
                            tma_load_3d(a, &map_q, &full[s], kb * BK, ti * TM, k - 1);
                            tma_load_3d(b, map_b, &full[s], kb * BK, tj * TN, l - 1);
                            tma_load_3d(b + 128 * BK, map_b, &full[s], kb * BK, tj * TN + 128, l - 1);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer: the whole warp runs the (uniform) loops and waits, one elected lane issues (i8_common.cuh: elect_one)
        {
            // instruction descriptor: D = s32, A = B = signed 8-bit, both K-major, N = 256, M = 128
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            int iter = 0, p = 0;
            for (int c = 0; c < n_chunks; ++c) {
                const int ck0 = kb0 + c * KC_BLOCKS, ck1 = min(ck0 + KC_BLOCKS, kb1);
                for (int d = 2; d <= D; ++d, ++p) {
                    const int buf = p & 1;
                    if (p >= 2) {  // the epilogue must have drained this accumulator
                        mbar_wait(&acc_empty[buf], ((p >> 1) - 1) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    }
                    const uint32_t tmem_d = tmem_base + (uint32_t)(buf * TN);
                    const int n_it = (min(S, d - 1) - max(1, d - S) + 1) * (ck1 - ck0);
                    for (int it = 0; it < n_it; ++it, ++iter) {
                        const int s = iter % STAGES;
                        mbar_wait(&full[s], (iter / STAGES) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                        const uint32_t a = smem_u32(smem + (size_t)s * STAGE_BYTES), b = a + A_BYTES;
                        if (elect_one()) {
#pragma unroll
                            for (int ks = 0; ks < BK / 32; ++ks)
                                umma_i8(tmem_d, umma_desc_sw128(a + ks * 32), umma_desc_sw128(b + ks * 32), idesc, (it | ks) ? 1u : 0u);
                            umma_commit(&empty[s]);  // the stage may be refilled once these MMAs have read it
                        }
                        __syncwarp();
                    }
                    if (elect_one()) umma_commit(&acc_full[buf]);
                    __syncwarp();
                }
            }
        }
    } else {
        // ---- epilogue: TMEM lanes 32*(warp % 4) .. +31 are this warp's tile rows.  tcgen05.ld hands a thread one ROW of
        // 32 accumulators; a 32 x 32 transposing tile in shared memory turns that into one COLUMN per lane, so that every
        // global access of the read-modify-write is a full 256-byte row segment.
        const int quad = warp & 3;
        const int64_t row_lo = (int64_t)ti * TM + quad * 32, row_hi = row_lo + 31;
        double *out = args.out[gram] + (int64_t)split * args.split_stride;
        int *tile = epi_tiles + quad * 32 * EPI_PITCH;
        int p = 0;
        for (int c = 0; c < n_chunks; ++c) {
            for (int d = 2; d <= D; ++d, ++p) {
                const int buf = p & 1;
                mbar_wait(&acc_full[buf], (p >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const double scale = pow2(4 - 8 * d);
                for (int cc = 0; cc < TN / 32; ++cc) {
                    const int64_t col0 = (int64_t)tj * TN + cc * 32;
                    if (col0 > row_hi || col0 >= args.N0) break;  // warp-uniform: the rest of the tile is above the diagonal
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * TN + cc * 32), v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) tile[lane * EPI_PITCH + i] = (int)v[i];
                    __syncwarp();
                    const int64_t col = col0 + lane;
                    double *o = out + row_lo * args.N0 + col;
                    const int n_rows = (int)min((int64_t)32, args.N0 - row_lo);
                    double acc[32];
#pragma unroll
                    for (int r = 0; r < 32; ++r)
                        acc[r] = (p > 0 && r < n_rows && col <= row_lo + r) ? o[(int64_t)r * args.N0] : 0.0;
#pragma unroll
                    for (int r = 0; r < 32; ++r)
                        if (r < n_rows && col <= row_lo + r) o[(int64_t)r * args.N0] = fma((double)tile[r * EPI_PITCH + lane], scale, acc[r]);
                    __syncwarp();
                }
                asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[buf]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(2 * TN));
    }
}

// G[t][s] (s <= t) = (sum of the K-split partials, in index order) * 2^(eA_t + eB_s)
// (part may alias G when there is a single split: every element is read and written by the same thread)
// Non-finite inputs (level 3 from the slicing kernels: a diverged activation collection) give an all-NaN Gram, which is what
// the fp64 contraction and the streaming walk make of them too -- never finite garbage from meaningless digits.
__global__ void i8_finish_kernel(const double *part, int nsplit, int64_t split_stride, const int32_t *__restrict__ eA,
                                 const int32_t *__restrict__ eB, int64_t N0, double *G, const int *__restrict__ wide) {
    const int64_t t = blockIdx.y;
    const int et = eA[t];
    const bool poisoned = *wide >= 3;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s <= t; s += (int64_t)gridDim.x * blockDim.x) {
        double tot = part[t * N0 + s];
        for (int k = 1; k < nsplit; ++k) tot += part[(int64_t)k * split_stride + t * N0 + s];
        G[t * N0 + s] = poisoned ? __longlong_as_double(0x7ff8000000000000LL) : ldexp(tot, et + eB[s]);
    }
}

// ---- host side -------------------------------------------------------------------------------------------------
static int make_slice_map(gpfq_ctx *ctx, CUtensorMap *map, int8_t *slices, int64_t N0p, int64_t mp) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return gpfq_fail(ctx, GPFQ_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3] = {(cuuint64_t)mp, (cuuint64_t)N0p, (cuuint64_t)i8g::S};
    const cuuint64_t strides[2] = {(cuuint64_t)mp, (cuuint64_t)mp * (cuuint64_t)N0p};  // bytes, dims 1 and 2
    const cuuint32_t box[3] = {(cuuint32_t)i8g::BK, 128u, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult rc = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, slices, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return gpfq_fail(ctx, GPFQ_ERR_CUDA, "cuTensorMapEncodeTiled failed with code %d", (int)rc);
    return GPFQ_OK;
}

struct I8Plan {
    int64_t N0p, mp;
    int tiles, nsplit, kb_total, kb_per;
    size_t slice_bytes, part_bytes;  // per matrix / per Gram
};

static I8Plan i8_plan(int sm_count, int64_t N0, int64_t m, bool same) {
    using namespace i8g;
    I8Plan p;
    p.N0p = ceil_div64(N0, TN) * TN;
    p.mp = ceil_div64(m, BK) * BK;
    p.kb_total = (int)(p.mp / BK);
    const int tiles_m = (int)ceil_div64(N0, TM);
    p.tiles = 0;
    for (int ti = 0; ti < tiles_m; ++ti) p.tiles += (ti >> 1) + 1;  // tj <= (ti*128 + 127) / 256
    const int64_t base = (int64_t)p.tiles * (same ? 1 : 2);
    int64_t ns = base >= sm_count ? 1 : sm_count / base;           // fill one wave when the tile count is small
    ns = std::min<int64_t>(ns, std::max<int64_t>(1, p.kb_total / 8));  // at least 8 K blocks per split
    ns = std::min<int64_t>(ns, 32);
    p.kb_per = (int)ceil_div64(p.kb_total, ns);
    p.nsplit = (int)ceil_div64(p.kb_total, p.kb_per);              // no empty split
    p.slice_bytes = (size_t)S * p.N0p * p.mp;
    p.part_bytes = p.nsplit > 1 ? (size_t)p.nsplit * N0 * N0 * sizeof(double) : 0;
    return p;
}

// Device workspace the int8 Gram needs beyond G1/G2 (bytes); the caller takes the DMMA path when this does not fit.
size_t gram_i8_workspace_bytes(gpfq_ctx *ctx, int64_t N0, int64_t m, bool same) {
    const I8Plan p = i8_plan(ctx->sm_count, N0, m, same);
    return (size_t)(same ? 1 : 2) * (p.slice_bytes + p.part_bytes) + ((size_t)1 << 20);
}

// G2 = Xq Xq^T and (unless X == Xq) G1 = Xq X^T, lower triangles, fp64 (N0, N0).
int gram_i8_stage(gpfq_ctx *ctx, const float *X, const float *Xq, int64_t ldx, int64_t N0, int64_t m, double *G1, double *G2) {
    using namespace i8g;
    const bool same = (X == Xq);
    if (m >= ((int64_t)1 << 31) - BK || N0 >= ((int64_t)1 << 22))
        return gpfq_fail(ctx, GPFQ_ERR_ARG, "int8 Gram: shape too large");
    const I8Plan pl = i8_plan(ctx->sm_count, N0, m, same);
    cudaStream_t st = ctx->stream;
    int8_t *sl_q = nullptr, *sl_x = nullptr;
    int32_t *e_q = nullptr, *e_x = nullptr;
    int *wide = nullptr;
    GPFQ_TRY(gpfq_ws(ctx, WS_I8_SQ, pl.slice_bytes, (void **)&sl_q));
    GPFQ_TRY(gpfq_ws(ctx, WS_I8_E, (size_t)(2 * pl.N0p + 4) * sizeof(int32_t), (void **)&e_q));
    e_x = e_q + pl.N0p;
    wide = e_q + 2 * pl.N0p;
    if (!same) {
        GPFQ_TRY(gpfq_ws(ctx, WS_I8_SX, pl.slice_bytes, (void **)&sl_x));
    } else {
        sl_x = sl_q;
        e_x = e_q;
    }
    double *part2 = G2, *part1 = G1;
    if (pl.nsplit > 1) {
        GPFQ_TRY(gpfq_ws(ctx, WS_PART, (size_t)(same ? 1 : 2) * pl.part_bytes, (void **)&part2));
        part1 = part2 + (size_t)pl.nsplit * N0 * N0;
    }
    // slicing
    CUDA_TRY(ctx, cudaMemsetAsync(wide, 0, 4 * sizeof(int32_t), st));
    const int vec = (ldx % 4 == 0) && ((uintptr_t)X % 16 == 0) && ((uintptr_t)Xq % 16 == 0);
    dim3 sgrid((unsigned)ceil_div64(pl.mp, 256 * 16), (unsigned)pl.N0p);
    i8_row_exponent_kernel<<<(unsigned)N0, 256, 0, st>>>(Xq, ldx, m, e_q, wide);
    KERNEL_CHECK(ctx);
    i8_split_kernel<<<sgrid, 256, 0, st>>>(Xq, ldx, N0, m, e_q, sl_q, pl.N0p, pl.mp, vec);
    KERNEL_CHECK(ctx);
    if (!same) {
        i8_row_exponent_kernel<<<(unsigned)N0, 256, 0, st>>>(X, ldx, m, e_x, wide);
        KERNEL_CHECK(ctx);
        i8_split_kernel<<<sgrid, 256, 0, st>>>(X, ldx, N0, m, e_x, sl_x, pl.N0p, pl.mp, vec);
        KERNEL_CHECK(ctx);
    }
    CUtensorMap map_q, map_x;
    GPFQ_TRY(make_slice_map(ctx, &map_q, sl_q, pl.N0p, pl.mp));
    GPFQ_TRY(make_slice_map(ctx, &map_x, sl_x, pl.N0p, pl.mp));
    I8Args a;
    a.tiles_m = (int)ceil_div64(N0, TM);
    a.nsplit = pl.nsplit;
    a.ngram = same ? 1 : 2;
    a.kb_total = pl.kb_total;
    a.kb_per = pl.kb_per;
    a.out[0] = part2;
    a.out[1] = part1;
    a.split_stride = N0 * N0;
    a.N0 = N0;
    a.wide = wide;
    a.d_force = ctx->i8_pairs_d >= 2 && ctx->i8_pairs_d <= 2 * S ? ctx->i8_pairs_d : 0;
    CUDA_TRY(ctx, cudaFuncSetAttribute(gram_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    const unsigned grid = (unsigned)((size_t)pl.tiles * pl.nsplit * a.ngram);
    gram_i8_kernel<<<grid, THREADS, SMEM, st>>>(map_q, map_x, a);
    KERNEL_CHECK(ctx);
    dim3 cgrid((unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(N0, 256), 64)), (unsigned)N0);
    i8_finish_kernel<<<cgrid, 256, 0, st>>>(part2, pl.nsplit, N0 * N0, e_q, e_q, N0, G2, wide);
    KERNEL_CHECK(ctx);
    if (!same) {
        i8_finish_kernel<<<cgrid, 256, 0, st>>>(part1, pl.nsplit, N0 * N0, e_q, e_x, N0, G1, wide);
        KERNEL_CHECK(ctx);
    }
    return GPFQ_OK;
}
