// conv.cu -- Conv2D / DepthwiseConv2D channels (K4): many channels per launch.
//
// Replaces the channel loop of _quantize_conv2D_layer_parallel_jit (quantized_network.py:844-860) and
// the per-filter pool of _quantize_channel_parallel_jit (:706-718), whose workers run the kk = kh*kw
// step walk of _quantize_filter2D_parallel_jit (:219-228).  kk is tiny (9 for 3x3) and n_patches is
// huge, so each channel is a pure HBM stream of its two (kk, n_patches) patch matrices:
//   stage 1  conv_gram_kernel     per channel G1 = Xq X^T (kk x kk), G2 = Xq Xq^T (lower), fp64
//                                 accumulation of exact fp32 products, per-CTA partials
//   stage 2  conv_finalize_kernel fixed-order sum of the partials (deterministic)
//   stage 3  conv_sweep_kernel    one thread per (channel, filter): the kk-step walk in registers
// Also here: the on-device single-channel im2col that serves gpfq_conv_layer_nhwc (the reference's
// _build_patch_array :729-809 / tf.image.extract_patches :158-172) and the MSQ baseline.
#include "common.cuh"

static constexpr int CONV_BLOCK = 256;
static constexpr int CONV_MAXKK = 9;

// ---- stage 1 ---------------------------------------------------------------------------------
// Thread group G (of NG) owns the Gram rows t = G, G+NG, ...; every group sweeps all columns of the
// CTA's chunk, so the 81+45 fp64 accumulators of a 3x3 channel are split over two thread groups.
template <int KK, bool SAME, int NG, int G>
__device__ __forceinline__ void conv_gram_body(const float *__restrict__ X, const float *__restrict__ Xq,
                                               int64_t n, int64_t c_beg, int64_t c_end, int tg, int gt,
                                               double *__restrict__ red) {
    constexpr int NR = (KK - G + NG - 1) / NG;
    double a1[NR][KK], a2[NR][KK];
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int s = 0; s < KK; ++s) a1[r][s] = a2[r][s] = 0.0;

    float cx[KK], cq[KK];
    int64_t p = c_beg + tg;
    if (p < c_end) {
#pragma unroll
        for (int s = 0; s < KK; ++s) {
            cq[s] = __ldg(Xq + (int64_t)s * n + p);
            if (!SAME) cx[s] = __ldg(X + (int64_t)s * n + p);
        }
    }
    for (; p < c_end; p += gt) {
        float nx[KK], nq[KK];
        const int64_t pn = p + gt;
        if (pn < c_end) {
#pragma unroll
            for (int s = 0; s < KK; ++s) {
                nq[s] = __ldg(Xq + (int64_t)s * n + pn);
                if (!SAME) nx[s] = __ldg(X + (int64_t)s * n + pn);
            }
        }
        double xd[KK], qd[KK];
#pragma unroll
        for (int s = 0; s < KK; ++s) {
            qd[s] = (double)cq[s];
            if (!SAME) xd[s] = (double)cx[s];
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int t = G + r * NG;
#pragma unroll
            for (int s = 0; s < KK; ++s) {
                if (!SAME) a1[r][s] = fma(qd[t], xd[s], a1[r][s]);
                if (s <= t) a2[r][s] = fma(qd[t], qd[s], a2[r][s]);
            }
        }
#pragma unroll
        for (int s = 0; s < KK; ++s) {
            cq[s] = nq[s];
            if (!SAME) cx[s] = nx[s];
        }
    }
    // fixed-order reduction: lanes (xor butterfly), then warps of this group in index order
    const int lane = threadIdx.x & 31, wig = tg >> 5, nwg = gt >> 5;
    double *myred = red + (size_t)(G * nwg + wig) * (2 * KK * KK);
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int t = G + r * NG;
#pragma unroll
        for (int s = 0; s < KK; ++s) {
            if (!SAME) {
                const double v = warp_sum(a1[r][s]);
                if (lane == 0) myred[t * KK + s] = v;
            }
            if (s <= t) {
                const double v = warp_sum(a2[r][s]);
                if (lane == 0) myred[KK * KK + t * KK + s] = v;
            }
        }
    }
}

// partial: (n_channels, n_chunks, 2*KK*KK)
template <int KK, bool SAME, int NG>
__global__ void __launch_bounds__(CONV_BLOCK, 1)
conv_gram_kernel(ConvPtrs ptrs, int64_t n, int64_t chunk_cols, double *__restrict__ partial) {
    __shared__ double red[(CONV_BLOCK / 32) * 2 * KK * KK];
    const int ch = blockIdx.y, chunk = blockIdx.x;
    const float *X = ptrs.Xp[ch];
    const float *Xq = SAME ? X : ptrs.Xqp[ch];
    const int64_t c_beg = (int64_t)chunk * chunk_cols;
    const int64_t c_end = (c_beg + chunk_cols < n) ? c_beg + chunk_cols : n;
    double *out = partial + ((size_t)ch * gridDim.x + chunk) * (2 * KK * KK);
    constexpr int GT = CONV_BLOCK / NG;
    const int g = threadIdx.x / GT, tg = threadIdx.x % GT;
    if (NG == 1) {
        conv_gram_body<KK, SAME, 1, 0>(X, Xq, n, c_beg, c_end, tg, GT, red);
    } else {
        if (g == 0) conv_gram_body<KK, SAME, NG, 0>(X, Xq, n, c_beg, c_end, tg, GT, red);
        else conv_gram_body<KK, SAME, NG, (NG > 1 ? 1 : 0)>(X, Xq, n, c_beg, c_end, tg, GT, red);
    }
    __syncthreads();
    // entry (t, s) was accumulated by group t % NG; sum that group's warps in index order
    constexpr int NWG = GT / 32;
    for (int e = threadIdx.x; e < 2 * KK * KK; e += CONV_BLOCK) {
        const int which = e / (KK * KK), t = (e % (KK * KK)) / KK, s = e % KK;
        if ((which == 0 && SAME) || (which == 1 && s > t)) continue;
        const int og = t % NG;
        double tot = 0.0;
        for (int w = 0; w < NWG; ++w) tot += red[(size_t)(og * NWG + w) * (2 * KK * KK) + e];
        out[e] = tot;
    }
}

// ---- stage 2 ---------------------------------------------------------------------------------
// gram: (n_channels, 2*kk*kk): [G1 | G2], lower triangle + diagonal of each valid.
__global__ void conv_finalize_kernel(const double *__restrict__ partial, int n_chunks, int kk, int same,
                                     double *__restrict__ gram) {
    const int ch = blockIdx.x, sz = 2 * kk * kk;
    for (int e = threadIdx.x; e < sz; e += blockDim.x) {
        const int which = e / (kk * kk), t = (e % (kk * kk)) / kk, s = e % kk;
        if (which == 1 && s > t) { gram[(size_t)ch * sz + e] = 0.0; continue; }
        const int src = (which == 0 && same) ? kk * kk + t * kk + s : e;
        if (which == 0 && same && s > t) { gram[(size_t)ch * sz + e] = 0.0; continue; }
        double tot = 0.0;
        for (int c = 0; c < n_chunks; ++c) tot += partial[((size_t)ch * n_chunks + c) * sz + src];
        gram[(size_t)ch * sz + e] = tot;
    }
}

// ---- stage 3 ---------------------------------------------------------------------------------
// One thread per (channel, filter, alphabet).  W/Q tap stride CF = C*F; channel c at offset c*F.
template <int KK>
__global__ void __launch_bounds__(128)
conv_sweep_kernel(const double *__restrict__ gram, const float *__restrict__ W, double *__restrict__ Q,
                  int64_t CF, int64_t F, int64_t c0, const double *__restrict__ alphabets,
                  const int *__restrict__ Koff, int64_t q_alph_stride) {
    __shared__ double g1[KK * KK], g2[KK * KK], nrm[KK], alph[GPFQ_MAX_K];
    const int ch = blockIdx.y, a = blockIdx.z;
    const int K = Koff[a + 1] - Koff[a];
    for (int e = threadIdx.x; e < KK * KK; e += blockDim.x) {
        g1[e] = gram[(size_t)ch * 2 * KK * KK + e];
        g2[e] = gram[(size_t)ch * 2 * KK * KK + KK * KK + e];
    }
    for (int e = threadIdx.x; e < K; e += blockDim.x) alph[e] = alphabets[Koff[a] + e];
    __syncthreads();
    if (threadIdx.x < KK) nrm[threadIdx.x] = (double)(float)sqrt(g2[threadIdx.x * KK + threadIdx.x]);
    __syncthreads();
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int64_t base = (c0 + ch) * F + f;
    double w[KK], q[KK];
#pragma unroll
    for (int t = 0; t < KK; ++t) w[t] = (double)W[(int64_t)t * CF + base];
#pragma unroll
    for (int t = 0; t < KK; ++t) {
        double d = 0.0;
#pragma unroll
        for (int s = 0; s < KK; ++s)
            if (s < t) d += w[s] * g1[t * KK + s] - q[s] * g2[t * KK + s];
        const double num = fma(w[t], g1[t * KK + t], d);
        q[t] = gpfq_decide(nrm[t], d, num, w[t], alph, K);
        Q[(int64_t)a * q_alph_stride + (int64_t)t * CF + base] = q[t];
    }
}

// ---- on-device single-channel im2col (extract_patches semantics) ------------------------------
// out[ch]: (kh*kw, n_img*Ho*Wo), row r*kw+c, column (img*Ho + i)*Wo + j.
__global__ void im2col_kernel(const float *__restrict__ act, int64_t n_img, int H, int Wd, int64_t C,
                              int64_t c_first, int kh, int kw, int sh, int sw, int rh, int rw, int pt, int pl,
                              int Ho, int Wo, float *__restrict__ out, int64_t ch_stride) {
    const int ch = blockIdx.y;
    const int tap = blockIdx.z;
    const int r = tap / kw, cc = tap % kw;
    const int64_t n = n_img * Ho * Wo;
    float *dst = out + (size_t)ch * ch_stride + (size_t)tap * n;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(p % Wo);
        const int i = (int)((p / Wo) % Ho);
        const int64_t img = p / ((int64_t)Wo * Ho);
        const int y = i * sh + r * rh - pt, x = j * sw + cc * rw - pl;
        float v = 0.f;
        if (y >= 0 && y < H && x >= 0 && x < Wd) v = act[((img * H + y) * Wd + x) * C + c_first + ch];
        dst[p] = v;
    }
}

// ---- MSQ -------------------------------------------------------------------------------------
template <typename T>
__global__ void msq_kernel(const T *__restrict__ W, int64_t n, const double *__restrict__ alphabet, int K,
                           double *__restrict__ Q) {
    __shared__ double alph[GPFQ_MAX_K];
    for (int e = threadIdx.x; e < K; e += blockDim.x) alph[e] = alphabet[e];
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        Q[i] = gpfq_bit_round((double)W[i], alph, K);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int KK>
static int launch_conv_gram(gpfq_ctx *ctx, ConvPtrs ptrs, bool same, int64_t n, int n_ch, int n_chunks,
                            int64_t chunk_cols, double *partial) {
    dim3 grid((unsigned)n_chunks, (unsigned)n_ch);
    constexpr int NG = (KK >= 9) ? 2 : 1;
    if (same) conv_gram_kernel<KK, true, 1><<<grid, CONV_BLOCK, 0, ctx->stream>>>(ptrs, n, chunk_cols, partial);
    else conv_gram_kernel<KK, false, NG><<<grid, CONV_BLOCK, 0, ctx->stream>>>(ptrs, n, chunk_cols, partial);
    KERNEL_CHECK(ctx);
    return GPFQ_OK;
}

int conv_supported_kk(int kk) { return kk == 1 || kk == 2 || kk == 3 || kk == 4 || kk == 6 || kk == 9; }

int conv_pick_chunks(gpfq_ctx *ctx, int64_t n, int n_ch, int64_t *chunk_cols) {
    int64_t want = ceil_div64(4LL * ctx->sm_count, n_ch);
    const int64_t max_chunks = ceil_div64(n, 4096) > 0 ? ceil_div64(n, 4096) : 1;
    if (want > max_chunks) want = max_chunks;
    if (want < 1) want = 1;
    int64_t cols = ceil_div64(n, want);
    cols = ceil_div64(cols, 128) * 128;
    *chunk_cols = cols;
    return (int)ceil_div64(n, cols);
}

// Gram partials of n_ch channels whose patch pointers (device) are in d_ptrs.
int conv_gram_stage(gpfq_ctx *ctx, int kk, ConvPtrs d_ptrs, bool same, int64_t n, int n_ch, int n_chunks,
                    int64_t chunk_cols, double *partial) {
    switch (kk) {
        case 1: return launch_conv_gram<1>(ctx, d_ptrs, same, n, n_ch, n_chunks, chunk_cols, partial);
        case 2: return launch_conv_gram<2>(ctx, d_ptrs, same, n, n_ch, n_chunks, chunk_cols, partial);
        case 3: return launch_conv_gram<3>(ctx, d_ptrs, same, n, n_ch, n_chunks, chunk_cols, partial);
        case 4: return launch_conv_gram<4>(ctx, d_ptrs, same, n, n_ch, n_chunks, chunk_cols, partial);
        case 6: return launch_conv_gram<6>(ctx, d_ptrs, same, n, n_ch, n_chunks, chunk_cols, partial);
        case 9: return launch_conv_gram<9>(ctx, d_ptrs, same, n, n_ch, n_chunks, chunk_cols, partial);
    }
    return gpfq_fail(ctx, GPFQ_ERR_UNSUPPORTED, "conv kernel size kk=%d has no specialised kernel", kk);
}

int conv_finalize_stage(gpfq_ctx *ctx, const double *partial, int n_ch, int n_chunks, int kk, bool same,
                        double *gram) {
    conv_finalize_kernel<<<(unsigned)n_ch, 192, 0, ctx->stream>>>(partial, n_chunks, kk, same ? 1 : 0, gram);
    KERNEL_CHECK(ctx);
    return GPFQ_OK;
}

int conv_sweep_stage(gpfq_ctx *ctx, int kk, const double *gram, const float *W, double *Q, int64_t C,
                     int64_t F, int64_t c0, int n_ch, const double *d_alph, const int *d_koff, int n_alph) {
    dim3 grid((unsigned)ceil_div64(F, 128), (unsigned)n_ch, (unsigned)n_alph);
    const int64_t CF = C * F, qs = (int64_t)kk * CF;
#define SWEEP_CASE(KKV) \
    case KKV: conv_sweep_kernel<KKV><<<grid, 128, 0, ctx->stream>>>(gram, W, Q, CF, F, c0, d_alph, d_koff, qs); break;
    switch (kk) {
        SWEEP_CASE(1) SWEEP_CASE(2) SWEEP_CASE(3) SWEEP_CASE(4) SWEEP_CASE(6) SWEEP_CASE(9)
        default: return gpfq_fail(ctx, GPFQ_ERR_UNSUPPORTED, "conv kernel size kk=%d has no specialised kernel", kk);
    }
#undef SWEEP_CASE
    KERNEL_CHECK(ctx);
    return GPFQ_OK;
}

int im2col_stage(gpfq_ctx *ctx, const float *act, int64_t n_img, int H, int Wd, int64_t C, int64_t c_first,
                 int n_ch, int kh, int kw, int sh, int sw, int rh, int rw, int pt, int pl, int Ho, int Wo,
                 float *out, int64_t ch_stride) {
    const int64_t n = n_img * Ho * Wo;
    int bx = (int)(ceil_div64(n, 256) < 2048 ? ceil_div64(n, 256) : 2048);
    dim3 grid((unsigned)bx, (unsigned)n_ch, (unsigned)(kh * kw));
    im2col_kernel<<<grid, 256, 0, ctx->stream>>>(act, n_img, H, Wd, C, c_first, kh, kw, sh, sw, rh, rw, pt, pl,
                                                 Ho, Wo, out, ch_stride);
    KERNEL_CHECK(ctx);
    return GPFQ_OK;
}

int msq_stage(gpfq_ctx *ctx, const void *W, int is_f64, int64_t n, const double *d_alph, int K, double *Q) {
    int bx = (int)(ceil_div64(n, 256) < 1184 ? ceil_div64(n, 256) : 1184);
    if (bx < 1) bx = 1;
    if (is_f64) msq_kernel<double><<<bx, 256, 0, ctx->stream>>>((const double *)W, n, d_alph, K, Q);
    else msq_kernel<float><<<bx, 256, 0, ctx->stream>>>((const float *)W, n, d_alph, K, Q);
    KERNEL_CHECK(ctx);
    return GPFQ_OK;
}
