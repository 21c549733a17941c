// dense_stream.cu -- Dense layer, streaming form (K3): the literal residual walk of
// _quantize_neuron_parallel (quantized_network.py:113-121) with the residual u kept on chip.
//
// One CTA walks J neurons through all N0 directions.  u (fp64, J x m) lives in shared memory (or an
// L2-resident global scratch when m is too long); every step streams the rows X_t, Xq_t once
// (coalesced) and
//   * applies the pending update of step t-1:  u += fl32(w X_{t-1}) - q Xq_{t-1}        (:119)
//   * accumulates  d = <Xq_t, u>  and  s = <Xq_t, u + fl32(w_t X_t)>                      (:86, :89)
// in one pass, then reduces across the CTA in a fixed order and takes the decision (:83-89).
// The arithmetic follows the reference's dtype ladder exactly (fp32 product w*X, fp64 everything
// else, no contraction in the update), so this kernel is the parity anchor of the library.
#include "common.cuh"

static constexpr int STREAM_T = 512;

// nrm[t] = (double)(float)sqrt(sum_i Xq[t][i]^2)  -- what scipy.linalg.norm(float32 row) returns (snrm2)
__global__ void __launch_bounds__(256) row_norms_kernel(const float *__restrict__ Xq, int64_t ldx,
                                                        int64_t m, double *__restrict__ nrm) {
    __shared__ double red[8];
    const float *row = Xq + (int64_t)blockIdx.x * ldx;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < m; i += 256) {
        const double v = (double)row[i];
        s = fma(v, v, s);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < 8; ++w) tot += red[w];
        nrm[blockIdx.x] = (double)(float)sqrt(tot);
    }
}

template <int J>
__global__ void __launch_bounds__(STREAM_T)
dense_stream_kernel(const float *__restrict__ X, const float *__restrict__ Xq, int64_t ldx, int64_t N0,
                    int64_t m, const float *__restrict__ W, int64_t ldw, int64_t j0, int64_t nj,
                    const double *__restrict__ nrm, const double *__restrict__ alphabet, int K,
                    double *__restrict__ Q, int64_t ldq, int64_t col0, double *__restrict__ u_scratch,
                    int u_in_smem) {
    extern __shared__ __align__(16) unsigned char stream_smem[];
    constexpr int NW = STREAM_T / 32;
    __shared__ double red[NW][2 * J];
    __shared__ double qsh[J];
    __shared__ double alph[GPFQ_MAX_K];

    const int64_t jb = (int64_t)blockIdx.x * J;
    double *u = u_in_smem ? reinterpret_cast<double *>(stream_smem)
                          : u_scratch + (int64_t)blockIdx.x * J * m;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int e = tid; e < K; e += STREAM_T) alph[e] = alphabet[e];
    for (int64_t e = tid; e < (int64_t)J * m; e += STREAM_T) u[e] = 0.0;
    __syncthreads();

    float wprev[J];
    double qprev[J];
#pragma unroll
    for (int j = 0; j < J; ++j) { wprev[j] = 0.f; qprev[j] = 0.0; }

    for (int64_t t = 0; t < N0; ++t) {
        float w[J];
#pragma unroll
        for (int j = 0; j < J; ++j) w[j] = (jb + j < nj) ? W[t * ldw + j0 + jb + j] : 0.f;
        const float *x = X + t * ldx, *xq = Xq + t * ldx;
        const float *xp = X + (t > 0 ? t - 1 : 0) * ldx, *xqp = Xq + (t > 0 ? t - 1 : 0) * ldx;
        double d[J], s[J];
#pragma unroll
        for (int j = 0; j < J; ++j) d[j] = s[j] = 0.0;

        for (int64_t i = tid; i < m; i += STREAM_T) {
            const float xv = x[i], xqv = xq[i];
            const double xqd = (double)xqv;
            float xpv = 0.f;
            double xqpd = 0.0;
            if (t > 0) { xpv = xp[i]; xqpd = (double)xqp[i]; }
#pragma unroll
            for (int j = 0; j < J; ++j) {
                double uj = u[(int64_t)j * m + i];
                if (t > 0) {
                    // u += w[t-1]*X_{t-1} - q[t-1]*Xq_{t-1}: fp32 product, fp64 product, sub, add (:119)
                    const double wx = (double)__fmul_rn(wprev[j], xpv);
                    const double qx = __dmul_rn(qprev[j], xqpd);
                    uj = __dadd_rn(uj, __dsub_rn(wx, qx));
                    u[(int64_t)j * m + i] = uj;
                }
                d[j] = fma(xqd, uj, d[j]);
                const double uw = __dadd_rn(uj, (double)__fmul_rn(w[j], xv));  // u + w*X  (:89)
                s[j] = fma(xqd, uw, s[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < J; ++j) {
            d[j] = warp_sum(d[j]);
            s[j] = warp_sum(s[j]);
        }
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < J; ++j) { red[warp][j] = d[j]; red[warp][J + j] = s[j]; }
        }
        __syncthreads();
        if (tid < J) {
            double dd = 0.0, ss = 0.0;
            for (int wv = 0; wv < NW; ++wv) { dd += red[wv][tid]; ss += red[wv][J + tid]; }
            const double q = gpfq_decide(nrm[t], dd, ss, (double)w[tid], alph, K);
            qsh[tid] = q;
            if (jb + tid < nj) Q[t * ldq + col0 + jb + tid] = q;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < J; ++j) { wprev[j] = w[j]; qprev[j] = qsh[j]; }
    }
}

// ---------------------------------------------------------------------------------------------
// Register-resident walk (m <= STREAM_T * 16).  One CTA walks J neurons; thread `tid` owns samples
// i = tid + e*STREAM_T (e < EPT) of every one of them, so the residual never leaves the register file and
// each step costs no shared-memory or global traffic beyond the two rows X_t, Xq_t, which are prefetched
// one step ahead (they do not depend on the decision).  Per step:
//   d_j = <Xq_t, u_j>                                  (:86)    one DFMA per sample and neuron
//   LITERAL: s_j = <Xq_t, u_j + fl32(w_j X_t)>         (:89)    the reference's dtype ladder, mul/sub/add unfused
//   else   : s_j = d_j + w_j <Xq_t, X_t>                         exact fp32 x fp32 products (what the Gram form computes)
//   one block reduction (warp butterflies, then every warp sums the per-warp partials in index order, so a
//   single barrier per step suffices), decision in lanes 0..J-1, broadcast by shuffle
//   u_j += fl32(w_j X_t) - q_j Xq_t   (:119)   [non-LITERAL: two DFMAs on the exact products]
// ---------------------------------------------------------------------------------------------
template <int EPT, int J, bool LITERAL>
__global__ void __launch_bounds__(STREAM_T)
dense_stream_reg_kernel(const float *__restrict__ X, const float *__restrict__ Xq, int64_t ldx, int64_t N0,
                        int64_t m, const float *__restrict__ W, int64_t ldw, int64_t j0, int64_t nj,
                        const double *__restrict__ nrm, const double *__restrict__ alphabet, int K,
                        double *__restrict__ Q, int64_t ldq, int64_t col0) {
    constexpr int NW = STREAM_T / 32;
    constexpr int V = LITERAL ? 2 * J : J + 1;  // values reduced per step
    __shared__ double red[2][NW][V];
    __shared__ double alph[GPFQ_MAX_K];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t jb = (int64_t)blockIdx.x * J;
    for (int e = tid; e < K; e += STREAM_T) alph[e] = alphabet[e];

    double u[J][EPT];
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
        for (int e = 0; e < EPT; ++e) u[j][e] = 0.0;

    auto load_rows = [&](int64_t t, float *rx, float *rq) {
        const float *x = X + t * ldx, *xq = Xq + t * ldx;
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const int64_t i = tid + (int64_t)e * STREAM_T;
            const bool ok = i < m;
            rx[e] = ok ? __ldg(x + i) : 0.f;
            rq[e] = ok ? __ldg(xq + i) : 0.f;
        }
    };
    auto load_w = [&](int64_t t, float *rw) {
#pragma unroll
        for (int j = 0; j < J; ++j) rw[j] = (jb + j < nj) ? __ldg(W + t * ldw + j0 + jb + j) : 0.f;
    };

    float cx[EPT], cq[EPT], w[J];
    double nrm_t = nrm[0];
    load_rows(0, cx, cq);
    load_w(0, w);
    __syncthreads();  // alphabet staged

    for (int64_t t = 0; t < N0; ++t) {
        float nx[EPT], nq[EPT], nw[J];
        double nrm_n = 0.0;
        if (t + 1 < N0) {
            load_rows(t + 1, nx, nq);
            load_w(t + 1, nw);
            nrm_n = nrm[t + 1];
        }
        double xqd[EPT], val[V];
#pragma unroll
        for (int v = 0; v < V; ++v) val[v] = 0.0;
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            xqd[e] = (double)cq[e];
            if (!LITERAL) val[J] = fma(xqd[e], (double)cx[e], val[J]);
#pragma unroll
            for (int j = 0; j < J; ++j) {
                val[j] = fma(xqd[e], u[j][e], val[j]);
                if (LITERAL) {
                    const double uw = __dadd_rn(u[j][e], (double)__fmul_rn(w[j], cx[e]));  // u + w*X  (:89)
                    val[J + j] = fma(xqd[e], uw, val[J + j]);
                }
            }
        }
#pragma unroll
        for (int v = 0; v < V; ++v) val[v] = warp_sum(val[v]);
        double(*rb)[V] = red[t & 1];
        if (lane == 0) {
#pragma unroll
            for (int v = 0; v < V; ++v) rb[warp][v] = val[v];
        }
        __syncthreads();
        // every warp finishes the reduction itself (same order everywhere => identical decisions)
        double q = 0.0;
        {
            double dd = 0.0, ss = 0.0;
            const int jj = lane < J ? lane : 0;
#pragma unroll
            for (int wv = 0; wv < NW; ++wv) {
                dd += rb[wv][jj];
                ss += rb[wv][LITERAL ? J + jj : J];
            }
            float wj = w[0];
#pragma unroll
            for (int j = 1; j < J; ++j) wj = (jj == j) ? w[j] : wj;
            const double num = LITERAL ? ss : fma((double)wj, ss, dd);
            q = gpfq_decide(nrm_t, dd, num, (double)wj, alph, K);
            if (warp == 0 && lane < J && jb + lane < nj) Q[t * ldq + col0 + jb + lane] = q;
        }
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const double qj = __shfl_sync(0xffffffffu, q, j);
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                if (LITERAL) {
                    // u += w[t]*X_t - q[t]*Xq_t: fp32 product, fp64 product, sub, add (:119)
                    const double wx = (double)__fmul_rn(w[j], cx[e]);
                    u[j][e] = __dadd_rn(u[j][e], __dsub_rn(wx, __dmul_rn(qj, xqd[e])));
                } else {
                    u[j][e] = fma(-qj, xqd[e], fma((double)w[j], (double)cx[e], u[j][e]));
                }
            }
        }
#pragma unroll
        for (int e = 0; e < EPT; ++e) { cx[e] = nx[e]; cq[e] = nq[e]; }
#pragma unroll
        for (int j = 0; j < J; ++j) w[j] = nw[j];
        nrm_t = nrm_n;
    }
}

template <int EPT, int J>
static int launch_stream_reg(gpfq_ctx *ctx, bool literal, const float *X, const float *Xq, int64_t ldx, int64_t N0,
                             int64_t m, const float *W, int64_t ldw, int64_t j0, int64_t nj, const double *nrm,
                             const double *d_alph, int K, double *Qd, int64_t ldq, int64_t col0) {
    const unsigned nblk = (unsigned)ceil_div64(nj, J);
    if (literal)
        dense_stream_reg_kernel<EPT, J, true><<<nblk, STREAM_T, 0, ctx->stream>>>(X, Xq, ldx, N0, m, W, ldw, j0, nj, nrm,
                                                                               d_alph, K, Qd, ldq, col0);
    else
        dense_stream_reg_kernel<EPT, J, false><<<nblk, STREAM_T, 0, ctx->stream>>>(X, Xq, ldx, N0, m, W, ldw, j0, nj, nrm,
                                                                                d_alph, K, Qd, ldq, col0);
    KERNEL_CHECK(ctx);
    return GPFQ_OK;
}

// (EPT, J) with EPT * J <= 24 residual doubles per thread
static int dispatch_stream_reg(gpfq_ctx *ctx, int ept, int J, bool literal, const float *X, const float *Xq, int64_t ldx,
                               int64_t N0, int64_t m, const float *W, int64_t ldw, int64_t j0, int64_t nj,
                               const double *nrm, const double *d_alph, int K, double *Qd, int64_t ldq, int64_t col0) {
#define SR(E, JJ) \
    if (ept == E && J == JJ) \
        return launch_stream_reg<E, JJ>(ctx, literal, X, Xq, ldx, N0, m, W, ldw, j0, nj, nrm, d_alph, K, Qd, ldq, col0);
    SR(2, 1) SR(2, 2) SR(2, 4) SR(2, 8)
    SR(4, 1) SR(4, 2) SR(4, 4)
    SR(6, 1) SR(6, 2) SR(6, 4)
    SR(8, 1) SR(8, 2)
    SR(12, 1) SR(12, 2)
    SR(16, 1)
#undef SR
    return gpfq_fail(ctx, GPFQ_ERR_ARG, "no streaming kernel for EPT=%d J=%d", ept, J);
}

template <int J>
static int launch_stream(gpfq_ctx *ctx, const float *X, const float *Xq, int64_t ldx, int64_t N0, int64_t m,
                         const float *W, int64_t ldw, int64_t j0, int64_t nj, const double *nrm,
                         const double *d_alph, int K, double *Qd, int64_t ldq, int64_t col0) {
    const int64_t nblk = ceil_div64(nj, J);
    const size_t need = (size_t)J * m * sizeof(double);
    const size_t static_smem = 8192;  // red/qsh/alph, generous
    const bool in_smem = need + static_smem <= ctx->smem_optin;
    double *scratch = nullptr;
    if (!in_smem) GPFQ_TRY(gpfq_ws(ctx, WS_U, (size_t)nblk * need, (void **)&scratch));
    auto k = dense_stream_kernel<J>;
    if (in_smem)
        CUDA_TRY(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    k<<<(unsigned)nblk, STREAM_T, in_smem ? need : 0, ctx->stream>>>(X, Xq, ldx, N0, m, W, ldw, j0, nj, nrm,
                                                                   d_alph, K, Qd, ldq, col0, scratch,
                                                                   in_smem ? 1 : 0);
    KERNEL_CHECK(ctx);
    return GPFQ_OK;
}

// Dense layer by the streaming walk.  Device pointers.  Alphabets are walked one after another
// (each needs its own residual); Qd: (n_alph, N0, ldq).
int dense_stream_path(gpfq_ctx *ctx, const float *X, const float *Xq, int64_t ldx, int64_t N0, int64_t m,
                      const float *W, int64_t ldw, int64_t j0, int64_t nj, const double *d_alph,
                      const int *h_koff, int n_alph, double *Qd, int64_t ldq, int64_t col0,
                      gpfq_stats *st) {
    double *nrm = nullptr;
    GPFQ_TRY(gpfq_ws(ctx, WS_NRM, (size_t)N0 * sizeof(double), (void **)&nrm));
    CUDA_TRY(ctx, gpfq_record(ctx, 2, ctx->stream));
    row_norms_kernel<<<(unsigned)N0, 256, 0, ctx->stream>>>(Xq, ldx, m, nrm);
    KERNEL_CHECK(ctx);
    const bool literal = ctx->stream_literal;
    if (m <= (int64_t)STREAM_T * 16) {
        // register-resident residual: EPT samples per thread, J neurons per CTA (EPT * J <= 24)
        const int need = (int)ceil_div64(m, STREAM_T);
        static const int epts[] = {2, 4, 6, 8, 12, 16};
        int ept = 16;
        for (int e : epts) if (e >= need) { ept = e; break; }
        int J = 1;
        while (J * 2 * ept <= 24 && J < 8 && ceil_div64(nj, J * 2) >= ctx->sm_count) J *= 2;
        for (int a = 0; a < n_alph; ++a) {
            const double *al = d_alph + h_koff[a];
            const int K = h_koff[a + 1] - h_koff[a];
            GPFQ_TRY(dispatch_stream_reg(ctx, ept, J, literal, X, Xq, ldx, N0, m, W, ldw, j0, nj, nrm, al, K,
                                         Qd + (int64_t)a * N0 * ldq, ldq, col0));
        }
    } else {
        // long sample axis: residual in shared memory (or an L2-resident scratch), literal arithmetic
        const size_t per_neuron = (size_t)m * sizeof(double);
        int J = 1;
        while (J < 4 && ceil_div64(nj, J * 2) >= ctx->sm_count && (size_t)(J * 2) * per_neuron + 8192 <= ctx->smem_optin)
            J *= 2;
        for (int a = 0; a < n_alph; ++a) {
            const double *al = d_alph + h_koff[a];
            const int K = h_koff[a + 1] - h_koff[a];
            double *Qa = Qd + (int64_t)a * N0 * ldq;
            if (J == 4) GPFQ_TRY(launch_stream<4>(ctx, X, Xq, ldx, N0, m, W, ldw, j0, nj, nrm, al, K, Qa, ldq, col0));
            else if (J == 2) GPFQ_TRY(launch_stream<2>(ctx, X, Xq, ldx, N0, m, W, ldw, j0, nj, nrm, al, K, Qa, ldq, col0));
            else GPFQ_TRY(launch_stream<1>(ctx, X, Xq, ldx, N0, m, W, ldw, j0, nj, nrm, al, K, Qa, ldq, col0));
        }
    }
    CUDA_TRY(ctx, gpfq_record(ctx, 3, ctx->stream));
    if (st) {
        st->method = GPFQ_METHOD_STREAM >> 4;
        st->flops_algorithmic = 6 * m * N0 * nj * n_alph;
        st->bytes_algorithmic = 8 * N0 * m + 12 * N0 * nj;
    }
    return GPFQ_OK;
}
