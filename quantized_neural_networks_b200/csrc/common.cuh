// common.cuh -- context, workspace pool, error plumbing and the scalar quantizer shared by all
// GPFQ kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/gpfq.h"

#define GPFQ_DEAD_NORM 1e-16   // quantized_network.py:83
#define GPFQ_PERP_DOT 1e-10    // quantized_network.py:86

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct ConvPtrs {
    const float *const *Xp;   // device array of n_channels device pointers
    const float *const *Xqp;
};

struct AlphEntry {
    std::vector<char> blob;  // levels (fp64), the prefix offsets (int32), one equispaced flag per alphabet (int32)
    size_t off = 0;          // byte offset inside the WS_ALPH device buffer
};

// One set of timing events + the static part of the report per API call; a ring of them lets callers
// enqueue many GPFQ_NO_SYNC calls and read every call's stage times afterwards (gpfq_query_stats).
struct CallRecord {
    cudaEvent_t e[8] = {};
    bool rec[8] = {};
    int kind = 0;  // 0 dense Gram+sweep, 1 dense stream, 2 conv
    gpfq_stats st = {};
};
static constexpr int GPFQ_RING = 128;

struct gpfq_ctx {
    int device = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t stream = nullptr;  // the one kernels launch on
    std::vector<CallRecord> ring;
    int64_t calls = 0;          // API calls begun so far
    CallRecord *cur = nullptr;  // record of the call in progress
    cudaEvent_t ev_copy[4] = {};
    cudaStream_t aux_stream[8] = {};   // residual-form sweep, up to 4 neuron groups: [g] the W part of group g's residual update,
                                       // [4 + g] the main chain of group g >= 1 (group 0 runs on `stream`)
    cudaEvent_t ev_chain[12] = {};     // ... [3 g] residuals sliced (main -> aux), [3 g + 1] W part landed (aux -> main), [3 g + 2] done
    int sweep_wq = 0;                  // ... 0 auto, 1 W part of the update on the aux stream, 2 W and Q parts as one two-product launch
    int sweep_range = 0;               // ... directions per range (0 auto; a multiple of 128)
    int sweep_groups = 0;              // ... neuron groups (0 auto, 1 / 2 / 4)
    std::string err = "";
    int launches = 0;
    int conv_variant = 0;         // 3x3 patch-Gram kernel: 0 TMA-staged / correlation form (default), 1 direct LDG, 2 generic,
                                  // 3 NHWC entry point: shared-memory planes kernel instead of the correlation form
    int corr_pack = 0;            // correlation form, images packed as virtual channels: 0 by shape, 1 always (tests), 2 never
    bool corr_direct_small = false;  // correlation form also on images below 128 pixels (tests)
    int corr_occ[2][9] = {};      // correlation-form kernels: resident CTAs per SM by [X~ == X][rows per band] (0: not queried yet)
    int corr_rb = 0;              // correlation-form conv Grams: rows per band (0: chosen per image height)
    int corr_strip_occ[2][2][3] = {};  // strip kernel: resident CTAs per SM by [X~ == X][rows per band 1 / 4][strip width 8 / 14 / 16]
    int corr_strip = 0;           // ... small images (W in {8, 14, 16, 28, 32, 56, 64}): 0 the strip kernel, 2 the band kernel
    int sweep_variant = 0;        // triangular sweep: 0 persistent neuron-tile kernel (default), 1 one launch pair per block
    bool stream_literal = false;  // streaming walk: reproduce the reference's fp32-rounded w*X products (set per call)
    int gram_variant = 0;         // Dense Gram stage: 0 auto, 1 fp64 DMMA (mma.sync), 2 int8 slices on tcgen05 (gram_i8.cu)
    int lowrank_variant = 0;      // sweep outer level: 0 auto, 1 Gram rows, 2 residual (low-rank) form
    int sweep_nt = 0;             // pipelined range walk: neurons per CTA (0 auto, 8 / 16 / 32)
    int sweep_walk = 0;           // range walk of the residual-form sweep: 0 auto, 1 tensor-core walk (sweep_tc.cu), 2 sweep_pipe / sweep_tile
    int last_sweep_tc = 0;        // the last residual-form sweep walked its ranges with sweep_tc_kernel
    int sweep_i8 = 0;             // sweep contractions of the residual form: 0 auto, 1 int8 slices on tcgen05, 2 fp64 DMMA
    int i8_pairs_d = 0;           // int8 Gram: keep slice pairs with k + l <= this (0: the default of gram_i8.cu)
    const double *h_alph = nullptr;  // host copy of the current call's alphabets (levels back to back) and their offsets
    const int *h_koff = nullptr, *h_flags = nullptr;
    int last_sweep_i8 = 0;        // the last residual-form sweep ran its contractions on tcgen05 (int8 slices)
    double *gram_only_out = nullptr;  // set for the duration of gpfq_conv_gram_nhwc: conv_finish hands the Grams out, no walk
    int last_gram_kernel = 0;     // what the last Dense Gram stage ran (1 DMMA, 2 int8 tcgen05)
    size_t i8_oom_bytes = 0;      // smallest int8-Gram workspace that failed to allocate (0: none yet)
    // grow-only named workspaces
    DevBuf ws[48];
    DevBuf pinned[4];
    // alphabets already resident on the device (steady-state calls re-use them: no copy, no sync)
    std::vector<AlphEntry> alph_cache;
    size_t alph_used = 0;
};

enum WsSlot {
    WS_X = 0, WS_XQ, WS_W, WS_Q, WS_WT, WS_QT, WS_G1, WS_G2, WS_PART, WS_DT, WS_NRM, WS_ALPH,
    WS_U, WS_PTRS, WS_CG, WS_CPART, WS_PATCH_A, WS_PATCH_B, WS_ACT_A, WS_ACT_B, WS_QIDX, WS_MISC,
    WS_I8_SQ, WS_I8_SX, WS_I8_E, WS_I8_TILES, WS_LR_XD, WS_LR_XT, WS_LR_U, WS_CORR_B,
    WS_SL_W, WS_SL_XT, WS_SL_XQT, WS_SL_XQ, WS_SL_U, WS_SL_KQ, WS_SL_E,
    WS_TC_TAB, WS_TC_G2S, WS_TC_G1S, WS_TC_E, WS_TC_P, WS_TC_W, WS_TC_G1M
};

int gpfq_fail(gpfq_ctx *ctx, int code, const char *fmt, ...);
cudaError_t gpfq_record(gpfq_ctx *ctx, int which, cudaStream_t s);  // record timing event `which` of the current call
int gpfq_ws(gpfq_ctx *ctx, int slot, size_t bytes, void **out);         // device workspace
int gpfq_pinned(gpfq_ctx *ctx, int slot, size_t bytes, void **out);     // pinned host staging

#define CUDA_TRY(ctx, expr)                                                                    \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return gpfq_fail((ctx), _e == cudaErrorMemoryAllocation ? GPFQ_ERR_OOM : GPFQ_ERR_CUDA, \
                             "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define GPFQ_TRY(expr)                \
    do {                              \
        int _rc = (expr);             \
        if (_rc != GPFQ_OK) return _rc; \
    } while (0)

#define KERNEL_CHECK(ctx)                                                                     \
    do {                                                                                      \
        (ctx)->launches++;                                                                    \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess)                                                                \
            return gpfq_fail((ctx), GPFQ_ERR_CUDA, "kernel launch failed: %s (%s:%d)",        \
                             cudaGetErrorString(_e), __FILE__, __LINE__);                     \
    } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// Scalar quantizer: alphabet[argmin_k |alphabet[k] - v|], first minimal index on ties
// (quantized_network.py:57).  Same fp64 subtraction/abs/compare as NumPy, so ties break alike.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double gpfq_bit_round(double v, const double *__restrict__ alph, int K) {
    double best = alph[0];
    double bd = fabs(__dsub_rn(best, v));
#pragma unroll 4
    for (int k = 1; k < K; ++k) {
        const double a = alph[k];
        const double d = fabs(__dsub_rn(a, v));
        if (d < bd) { bd = d; best = a; }
    }
    return best;
}

// out-of-line copy for the rarely taken branches of the fast paths (keeps their loop bodies small)
static __device__ __noinline__ double gpfq_bit_round_slow(double v, const double *__restrict__ alph, int K) {
    return gpfq_bit_round(v, alph, K);
}

// The same quantizer when the levels are ascending and equispaced (`rad * linspace(-1, 1, K)`, :396/:545 -- the host
// checks this per alphabet): |a_k - v| is unimodal in k, so the first minimal index of the full scan is the grid guess
// rint((v - a_0) / step) or one of its two neighbours (the guess is off by at most one, at a tie either side is in the
// window).  The three candidates are compared with the very same subtraction / abs / strict-less tests in ascending
// order, so ties still go to the lower index; indices clamped at the ends only repeat a level, which never wins a
// strict comparison against its own earlier copy.  Branch-free: this sits on the serial critical path of every walk.
// inv_step <= 0 (not equispaced) or a non-finite / huge argument: literal scan.
__device__ __forceinline__ double gpfq_bit_round_eq(double v, const double *__restrict__ alph, int K, double inv_step) {
    const double gpos = (v - alph[0]) * inv_step;
    if (!(inv_step > 0.0) || !(fabs(gpos) < 1e9)) return gpfq_bit_round_slow(v, alph, K);  // also NaN / inf arguments
    const int kr = __double2int_rn(gpos), hi = K - 1;
    const int i1 = min(max(kr, 0), hi), i0 = max(i1 - 1, 0), i2 = min(i1 + 1, hi);
    const double a0 = alph[i0], a1 = alph[i1], a2 = alph[i2];
    const double d0 = fabs(__dsub_rn(a0, v)), d1 = fabs(__dsub_rn(a1, v)), d2 = fabs(__dsub_rn(a2, v));
    double best = a0, bd = d0;
    if (d1 < bd) { bd = d1; best = a1; }
    if (d2 < bd) best = a2;
    return best;
}

// The same quantizer for the ternary alphabet {-a, 0, a} (bits = log2 3: the VGG16 and MNIST configurations): the literal
// three-level scan itself -- the same subtractions, absolute values and strict comparisons in ascending order, so every tie goes
// where _bit_round_parallel sends it -- with the levels in registers: no index arithmetic, no shared-memory loads on the
// serial chain of the walk (the windowed form above costs ~115 dependent cycles per step, this one ~25).
__device__ __forceinline__ double gpfq_bit_round_ternary(double v, double a) {
    const double dl = fabs(__dsub_rn(-a, v)), dm = fabs(__dsub_rn(0.0, v)), dh = fabs(__dsub_rn(a, v));
    double best = -a, bd = dl;
    if (dm < bd) { bd = dm; best = 0.0; }
    if (dh < bd) best = a;
    return best;
}
// a > 0 when the alphabet staged at `alph` is exactly {-a, 0, a}, else 0
__device__ __forceinline__ double gpfq_ternary_radius(const double *alph, int K) {
    return (K == 3 && alph[1] == 0.0 && alph[2] > 0.0 && alph[0] == -alph[2]) ? alph[2] : 0.0;
}

// inv_step of an alphabet staged in shared memory (flag from the host: 1 = ascending and equispaced)
__device__ __forceinline__ double gpfq_inv_step(const double *alph, int K, int equispaced) {
    return (equispaced && K >= 2) ? (double)(K - 1) / (alph[K - 1] - alph[0]) : 0.0;
}

// One greedy decision in Gram form:
//   nrm  = (double)(float)sqrt(G2[t,t])         (snrm2 result, quantized_network.py:83)
//   d    = <Xq_t, u_{t-1}>                       (:86)
//   num  = <Xq_t, u_{t-1} + w_t X_t>             (:89)
__device__ __forceinline__ double gpfq_decide(double nrm, double d, double num, double w,
                                              const double *__restrict__ alph, int K, double inv_step = 0.0) {
    if (nrm < GPFQ_DEAD_NORM) return 0.0;
    if (fabs(d) < GPFQ_PERP_DOT) return gpfq_bit_round_eq(w, alph, K, inv_step);
    return gpfq_bit_round_eq(num / (nrm * nrm), alph, K, inv_step);
}

// The same decision for the sweep's in-block walk, out of line (32 unrolled call sites share one copy) and with the
// division by nrm^2 done as a Markstein correction of num * RN(1/nrm^2): q0 = RN(num r), e = RN(num - q0 den) (exact,
// fma), v = RN(q0 + e r) is the correctly rounded quotient, i.e. bit-identical to num / den, while the reciprocal
// is computed once per direction instead of once per neuron and step.
static __device__ __noinline__ double gpfq_decide_rcp(double nrm, double rinv, double d, double num, double w,
                                               const double *__restrict__ alph, int K, double inv_step) {
    if (nrm < GPFQ_DEAD_NORM) return 0.0;
    double v = w;
    if (!(fabs(d) < GPFQ_PERP_DOT)) {
        const double den = nrm * nrm;
        const double q0 = num * rinv;
        const double e = fma(-q0, den, num);
        v = fma(e, rinv, q0);
    }
    return gpfq_bit_round_eq(v, alph, K, inv_step);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
