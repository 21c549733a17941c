// conv_corr.cu -- 3x3 / stride 1 / 'SAME' conv layers: per-channel Grams as shifted correlations of the activations.
//
// What the walk of _quantize_filter2D_parallel_jit (quantized_network.py:219-228) consumes per channel are the 9 x 9
// matrices G1[t,s] = <Xq_t, X_s>, G2[t,s] = <Xq_t, Xq_s> (s <= t) of the channel's patch matrix (:789-797, one row per
// tap t = r*3 + c, one column per output pixel p, entry = pixel p + tap - 1 of the zero-padded image).  Substituting
// the pixel s = p + a - 1 that tap a reads:
//
//     G[a, b] = sum over pixels s of  xq[s] * x[s + (b - a)]   restricted to  s - a + 1 inside the image,
//
// i.e. the correlation of the channel image with itself at displacement d = b - a, over all pixels s except ONE border
// row and/or ONE border column (tap row 0 never reads the bottom image row as its own pixel, tap row 2 never the top
// row; columns alike).  For s <= t the displacement is lexicographically <= 0: only 13 displacements exist, so a pixel
// costs 13 exact-product fp64 FMAs per Gram instead of the 81 + 45 of the patch form (126 useful MACs, 168 fp64-pipe
// slots in conv_gram9_nhwc_kernel).  The pixels s are split into nine disjoint regions -- interior, top / bottom row
// and left / right column without their ends, four corners -- whose 13 sums are kept apart and ADDED per (a, b) for the
// regions tap a may sit on.  No subtraction anywhere: a tap direction that only ever reads zeros (a dead channel, an
// image with black margins) gets an exactly zero Gram row, as in the patch form, so the dead-direction guard of
// quantized_network.py:83-84 fires identically.  Sums are exact fp32 x fp32 products accumulated in fp64 in a fixed
// order (tasks per warp in index order, warps in index order): bit-reproducible, and equal to the patch form up to fp64
// re-association (1e-16 relative; the parity budget is 1e-9).
//
//   conv_corr9_tma_kernel<RB, CROSS>  interior pixels.  lane = channel (NHWC: 32 channels = one 128-byte line per pixel);
//                                  a warp walks a band of RB image rows left to right with a (RB+2) x 5 fp64 register
//                                  window of the displaced operand and 13 accumulators per lane.  Operands arrive by TMA
//                                  (cp.async.bulk.tensor.4d over the (N, H, W, C) tensor, box = 32 channels x 5 columns x
//                                  RB+2 rows, zero fill outside the image = the 'SAME' padding) into a per-warp mbarrier
//                                  ring; the warp is its own producer (lane 0 issues the next box before multiplying the
//                                  current one).  CROSS: xq centre x x window (G1), else xq x xq (G2).  Bound by the fp64
//                                  pipe: 13 DFMA + ~1.3-2.3 F2F per pixel and pass.
//   conv_corr9_kernel<RB, CROSS>   the same walk with direct LDG loads (C % 4 != 0 or unaligned tensors: no tensor map)
//   conv_corr9_border_kernel       the same 13 sums for the eight border regions, bounds-checked loads
//   conv_corr9_assemble_kernel     fixed-order slot sums + the per-tap region sums -> [G1 | G2] per channel in the
//                                  layout conv_sweep_kernel reads
#include <cuda.h>

#include <algorithm>

#include "common.cuh"

namespace corr9 {
constexpr int ND = 13;        // displacements (dy, dx): dy in {-2, -1} x dx in [-2, 2], then dy = 0 x dx in {-2, -1, 0}
constexpr int WARPS = 4;      // warps (tasks in flight) per CTA
constexpr int NCLS = 8;       // border regions: top, bottom row and left, right column (ends excluded), corners TL, TR, BL, BR
constexpr int WC = 5;         // columns per TMA box = phases of the rotating register window
constexpr int REC = 2 * ND;   // doubles per (channel, slot): [cross sums | auto sums]
}  // namespace corr9

struct Corr9Geom {
    int H, W;
    int64_t C;         // channels of the activation tensor
    int64_t c_first;   // first channel of this launch
    int n_ch;          // channels of this launch
    int64_t img0, n_img;
    int slots;         // warps per channel group in this launch
};

// One band of RB image rows of one image, walked left to right.  pb: row y0 - 2, column 0 of the displaced operand
// (this lane's channel); pa: row y0, column 0 of xq (CROSS only).  EDGE: the band touches the top or the bottom of the
// image, rows are bounds-checked (interior bands skip the checks: one IMAD.WIDE + LDG per operand).
template <int RB, bool CROSS, bool EDGE>
__device__ __forceinline__ void corr9_band(double (&acc)[corr9::ND], const float *__restrict__ pb,
                                           const float *__restrict__ pa, int rowstride, int Cs, int W, int H, int y0,
                                           bool chok) {
    double win[RB + 2][5];
    double ac[RB];
    float rawb[RB + 2], rawa[RB];
#pragma unroll
    for (int r = 0; r < RB + 2; ++r)
#pragma unroll
        for (int c = 0; c < 5; ++c) win[r][c] = 0.0;
    (void)ac;
    (void)rawa;

    auto load = [&](int cx) {
        const bool colok = chok && cx < W;
        const float *pc = pb + cx * Cs;
#pragma unroll
        for (int r = 0; r < RB + 2; ++r) {
            const bool ok = colok && (!EDGE || (unsigned)(y0 - 2 + r) < (unsigned)H);
            rawb[r] = ok ? __ldg(pc + r * rowstride) : 0.f;
        }
        if (CROSS) {
            const bool aok = chok && (unsigned)(cx - 2) < (unsigned)W;
            const float *pac = pa + (cx - 2) * Cs;
#pragma unroll
            for (int i = 0; i < RB; ++i) {
                const bool ok = aok && (!EDGE || (y0 + i) < H);
                rawa[i] = ok ? __ldg(pac + i * rowstride) : 0.f;
            }
        }
    };

    load(0);
    for (int cx0 = 0; cx0 < W + 2; cx0 += 5) {
#pragma unroll
        for (int ph = 0; ph < 5; ++ph) {
            const int cx = cx0 + ph;  // newest window column (image column cx) lives in physical slot ph
            if (cx < W + 2) {
#pragma unroll
                for (int r = 0; r < RB + 2; ++r) win[r][ph] = (double)rawb[r];
                if (CROSS) {
#pragma unroll
                    for (int i = 0; i < RB; ++i) ac[i] = (double)rawa[i];
                }
                load(cx + 1);  // next column in flight while this one is multiplied
                if (cx >= 3 && cx <= W) {
                    // interior pixel column cx - 2 in [1, W - 2] = logical window column 2 = physical slot (ph + 3) % 5
#pragma unroll
                    for (int i = 0; i < RB; ++i) {
                        double a = CROSS ? ac[i] : win[i + 2][(ph + 3) % 5];
                        if (EDGE && !((unsigned)(y0 + i - 1) < (unsigned)(H - 2))) a = 0.0;  // rows 0 and H - 1: border kernel
#pragma unroll
                        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                            for (int dx = 0; dx < 5; ++dx)
                                acc[dy * 5 + dx] = fma(a, win[i + dy][(ph + 1 + dx) % 5], acc[dy * 5 + dx]);
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx)
                            acc[10 + dx] = fma(a, win[i + 2][(ph + 1 + dx) % 5], acc[10 + dx]);
                    }
                }
            }
        }
    }
}

template <int RB, bool CROSS>
__global__ void __launch_bounds__(corr9::WARPS * 32)
conv_corr9_kernel(const float *__restrict__ actq, const float *__restrict__ actx, Corr9Geom gm,
                  double *__restrict__ partial, int slot_stride, int slot0) {
    using namespace corr9;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = blockIdx.x * WARPS + warp;
    const int chl = blockIdx.y * 32 + lane;
    const bool chok = chl < gm.n_ch;
    const int64_t ch = gm.c_first + (chok ? chl : 0);
    const int H = gm.H, W = gm.W, Cs = (int)gm.C;
    const int nbands = (H + RB - 1) / RB;
    const int64_t ntasks = gm.n_img * nbands;
    const int rowstride = W * Cs;  // elements per image row (< 2^27, checked by corr9_plan)
    const float *bsrc = CROSS ? actx : actq;
    double acc[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) acc[d] = 0.0;

    for (int64_t task = slot; task < ntasks; task += gm.slots) {
        const int64_t img = gm.img0 + task / nbands;
        const int y0 = (int)(task % nbands) * RB;
        // row y0 - 2 of this lane's channel (may lie above the image: never dereferenced then)
        const float *pb = bsrc + (img * H + (y0 - 2)) * (int64_t)rowstride + ch;
        const float *pa = actq + (img * H + y0) * (int64_t)rowstride + ch;
        if (y0 >= 2 && y0 + RB <= H - 1) corr9_band<RB, CROSS, false>(acc, pb, pa, rowstride, Cs, W, H, y0, chok);
        else corr9_band<RB, CROSS, true>(acc, pb, pa, rowstride, Cs, W, H, y0, chok);
    }
    if (chok) {
        double *out = partial + ((size_t)chl * slot_stride + slot0 + slot) * REC + (CROSS ? 0 : ND);
#pragma unroll
        for (int d = 0; d < ND; ++d) out[d] = acc[d];
    }
}

// ---- TMA-fed variant ---------------------------------------------------------------------------------------------
namespace corr9 {
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
// box of the (C, W, H, N) tensor at (channel c, column x, row y, image n); coordinates outside the tensor read zeros
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c, int x, int y, int n) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(
            smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c), "r"(x), "r"(y), "r"(n)
        : "memory");
}
template <int RB, bool CROSS>
struct Ring {
    static constexpr int NS = CROSS ? 2 : 3;                 // boxes in flight per warp
    static constexpr int B_FLOATS = (RB + 2) * WC * 32;      // window operand: RB + 2 rows x 5 columns x 32 channels
    static constexpr int A_FLOATS = CROSS ? RB * WC * 32 : 0;  // xq centre pixels, two columns behind
    static constexpr int STAGE_FLOATS = B_FLOATS + A_FLOATS;
    static constexpr size_t WARP_BYTES = (size_t)NS * STAGE_FLOATS * sizeof(float);
    static constexpr size_t SMEM = WARPS * WARP_BYTES + WARPS * NS * sizeof(uint64_t);
};
}  // namespace corr9

template <int RB, bool CROSS>
__global__ void __launch_bounds__(corr9::WARPS * 32)
conv_corr9_tma_kernel(const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapA, Corr9Geom gm,
                      double *__restrict__ partial, int slot_stride, int slot0) {
    using namespace corr9;
    using R = Ring<RB, CROSS>;
    constexpr int NS = R::NS;
    extern __shared__ __align__(128) unsigned char corr9_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = blockIdx.x * WARPS + warp;
    const int chl = blockIdx.y * 32 + lane;
    const bool chok = chl < gm.n_ch;
    float *ring = reinterpret_cast<float *>(corr9_smem + warp * R::WARP_BYTES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(corr9_smem + WARPS * R::WARP_BYTES) + warp * NS;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();
    const int H = gm.H, W = gm.W;
    const int c0 = (int)(gm.c_first + (int64_t)blockIdx.y * 32);
    const int nbands = (H + RB - 1) / RB;
    const int64_t ntasks = gm.n_img * nbands;
    const int nstages = (W + 2 + WC - 1) / WC;
    uint32_t it = 0;  // boxes this warp has consumed so far: ring position and mbarrier parity
    double acc[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) acc[d] = 0.0;

    for (int64_t task = slot; task < ntasks; task += gm.slots) {
        const int img = (int)(gm.img0 + task / nbands);
        const int y0 = (int)(task % nbands) * RB;
        unsigned rowmask = 0;  // bit i: pixel row y0 + i is an interior row (1 .. H - 2)
#pragma unroll
        for (int i = 0; i < RB; ++i) rowmask |= ((unsigned)(y0 + i - 1) < (unsigned)(H - 2)) ? (1u << i) : 0u;
        double win[RB + 2][5];
        double ac[RB];
        (void)ac;
#pragma unroll
        for (int r = 0; r < RB + 2; ++r)
#pragma unroll
            for (int c = 0; c < 5; ++c) win[r][c] = 0.0;
        auto issue = [&](int k, uint32_t pos) {
            if (lane == 0) {
                const int s = (int)(pos % NS);
                float *dst = ring + s * R::STAGE_FLOATS;
                mbar_expect_tx(&bars[s], (uint32_t)(R::STAGE_FLOATS * sizeof(float)));
                tma_load_4d(dst, &mapB, &bars[s], c0, k * WC, y0 - 2, img);
                if (CROSS) tma_load_4d(dst + R::B_FLOATS, &mapA, &bars[s], c0, k * WC - 2, y0, img);
            }
        };
        __syncwarp();  // every lane has finished reading the previous task's boxes
#pragma unroll
        for (int p = 0; p < NS - 1; ++p)
            if (p < nstages) issue(p, it + p);
        for (int k = 0; k < nstages; ++k) {
            __syncwarp();  // the box consumed in iteration k - 1 is free again
            if (k + NS - 1 < nstages) issue(k + NS - 1, it + k + NS - 1);
            const uint32_t pos = it + k;
            const int s = (int)(pos % NS);
            mbar_wait(&bars[s], (pos / NS) & 1u);
            const float *tb = ring + s * R::STAGE_FLOATS + lane;
            const float *ta = tb + R::B_FLOATS;
#pragma unroll
            for (int ph = 0; ph < WC; ++ph) {
                const int cx = k * WC + ph;  // newest window column (image column cx) lives in physical slot ph
                if (cx < W + 2) {
#pragma unroll
                    for (int r = 0; r < RB + 2; ++r) win[r][ph] = (double)tb[(r * WC + ph) * 32];
                    if (CROSS) {
#pragma unroll
                        for (int i = 0; i < RB; ++i) ac[i] = (double)ta[(i * WC + ph) * 32];
                    }
                    if (cx >= 3 && cx <= W) {
                        // interior pixel column cx - 2 in [1, W - 2] = logical window column 2 = physical slot (ph + 3) % 5
#pragma unroll
                        for (int i = 0; i < RB; ++i) {
                            double a = CROSS ? ac[i] : win[i + 2][(ph + 3) % 5];
                            a = (rowmask >> i) & 1u ? a : 0.0;
#pragma unroll
                            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                                for (int dx = 0; dx < 5; ++dx)
                                    acc[dy * 5 + dx] = fma(a, win[i + dy][(ph + 1 + dx) % 5], acc[dy * 5 + dx]);
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx)
                                acc[10 + dx] = fma(a, win[i + 2][(ph + 1 + dx) % 5], acc[10 + dx]);
                        }
                    }
                }
            }
        }
        it += (uint32_t)nstages;
    }
    if (chok) {
        double *out = partial + ((size_t)chl * slot_stride + slot0 + slot) * REC + (CROSS ? 0 : ND);
#pragma unroll
        for (int d = 0; d < ND; ++d) out[d] = acc[d];
    }
}

// Border regions.  Slot s of a launch handles class s % 8 for images s / 8, s / 8 + slots / 8, ...
template <bool SAME>
__global__ void __launch_bounds__(corr9::WARPS * 32)
conv_corr9_border_kernel(const float *__restrict__ actq, const float *__restrict__ actx, Corr9Geom gm,
                         double *__restrict__ bpartial, int slot_stride, int slot0) {
    using namespace corr9;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = blockIdx.x * WARPS + warp;
    const int chl = blockIdx.y * 32 + lane;
    const bool chok = chl < gm.n_ch;
    const int64_t ch = gm.c_first + (chok ? chl : 0);
    const int H = gm.H, W = gm.W;
    const int64_t Cs = gm.C;
    const int cls = slot % NCLS;
    const int64_t first = slot / NCLS, step = gm.slots / NCLS;
    double a1[ND], a2[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) a1[d] = a2[d] = 0.0;
    // pixels of the class: (y, x) = (ys + k * yk, xs + k * xk), k < n
    int ys = 0, xs = 0, yk = 0, xk = 0, n = 1;
    switch (cls) {
        case 0: n = W - 2; xk = 1; xs = 1; break;                       // top row without its ends
        case 1: n = W - 2; xk = 1; xs = 1; ys = H - 1; break;           // bottom row
        case 2: n = H - 2; yk = 1; ys = 1; break;                       // left column without its ends
        case 3: n = H - 2; yk = 1; ys = 1; xs = W - 1; break;           // right column
        case 4: break;                                      // top-left
        case 5: xs = W - 1; break;                          // top-right
        case 6: ys = H - 1; break;                          // bottom-left
        default: ys = H - 1; xs = W - 1; break;             // bottom-right
    }
    if (chok) {
        for (int64_t im = first; im < gm.n_img; im += step) {
            const float *q = actq + (gm.img0 + im) * (int64_t)H * W * Cs + ch;
            const float *x = SAME ? q : actx + (gm.img0 + im) * (int64_t)H * W * Cs + ch;
            for (int k = 0; k < n; ++k) {
                const int y = ys + k * yk, xx = xs + k * xk;
                const double a = (double)__ldg(q + ((int64_t)y * W + xx) * Cs);
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    const int dy = d < 10 ? d / 5 - 2 : 0, dx = d < 10 ? d % 5 - 2 : d - 12;
                    const int yy = y + dy, xc = xx + dx;
                    if ((unsigned)yy < (unsigned)H && (unsigned)xc < (unsigned)W) {
                        const int64_t off = ((int64_t)yy * W + xc) * Cs;
                        a2[d] = fma(a, (double)__ldg(q + off), a2[d]);
                        if (!SAME) a1[d] = fma(a, (double)__ldg(x + off), a1[d]);
                    }
                }
            }
        }
        double *out = bpartial + ((size_t)chl * slot_stride + slot0 + slot) * REC;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            out[d] = a1[d];
            out[ND + d] = a2[d];
        }
    }
}

// gram: (n_channels, 2 * 81): [G1 | G2], lower triangle + diagonal valid, zeros above (conv_finalize_kernel's layout).
__global__ void __launch_bounds__(96)
conv_corr9_assemble_kernel(const double *__restrict__ partial, int slots, const double *__restrict__ bpartial, int bslots,
                           int same, double *__restrict__ gram) {
    using namespace corr9;
    __shared__ double T[REC], B[NCLS][REC];
    const int ch = blockIdx.x, tid = threadIdx.x;
    if (tid < REC) {
        double tot = 0.0;
        for (int s = 0; s < slots; ++s) tot += partial[((size_t)ch * slots + s) * REC + tid];  // slots in index order
        T[tid] = tot;
    }
    for (int e = tid; e < NCLS * REC; e += blockDim.x) {
        const int cls = e / REC, i = e % REC;
        double tot = 0.0;
        for (int s = cls; s < bslots; s += NCLS) tot += bpartial[((size_t)ch * bslots + s) * REC + i];
        B[cls][i] = tot;
    }
    __syncthreads();
    for (int e = tid; e < 162; e += blockDim.x) {
        const int which = e / 81, t = (e % 81) / 9, s = e % 9;
        double v = 0.0;
        if (s <= t) {
            const int ar = t / 3, ac = t % 3, dy = s / 3 - ar, dx = s % 3 - ac;
            const int id = (which == 0 && !same ? 0 : ND) + (dy < 0 ? (dy + 2) * 5 + dx + 2 : 12 + dx);
            // tap row 0 never sits on the bottom image row, tap row 2 never on the top row; columns alike
            const bool top = ar != 2, bot = ar != 0, lef = ac != 2, rig = ac != 0;
            v = T[id];                       // interior pixels
            if (top) v += B[0][id];
            if (bot) v += B[1][id];
            if (lef) v += B[2][id];
            if (rig) v += B[3][id];
            if (top && lef) v += B[4][id];
            if (top && rig) v += B[5][id];
            if (bot && lef) v += B[6][id];
            if (bot && rig) v += B[7][id];
        }
        gram[(size_t)ch * 162 + e] = v;
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
// Is the layer eligible, and with how many rows per band?
int corr9_plan(int kh, int kw, int sh, int sw, int rh, int rw, int padding_same, int H, int W, int64_t C, int n_ch) {
    if (kh != 3 || kw != 3 || sh != 1 || sw != 1 || rh != 1 || rw != 1 || !padding_same) return 0;
    if (H < 2 || W < 2) return 0;                                    // the nine pixel regions must be disjoint
    if (n_ch < 8 || (int64_t)W * C >= ((int64_t)1 << 27)) return 0;  // few channels: lanes idle (lane = channel)
    int best = 0, pad = 1 << 30;
    const int cand[3] = {8, 7, 4};
    for (int i = 0; i < 3; ++i) {
        const int p = (H + cand[i] - 1) / cand[i] * cand[i] - H;
        if (p < pad) { pad = p; best = cand[i]; }
    }
    return best;
}

// Slots (warps per channel group) that keep every SM busy for about `waves` rounds.
int corr9_pick_slots(gpfq_ctx *ctx, int n_ch, int64_t ntasks) {
    using namespace corr9;
    const int64_t groups = ceil_div64(n_ch, 32);
    int64_t per = ceil_div64((int64_t)ctx->sm_count * 2 * WARPS, groups);  // two CTAs of 4 warps resident per SM (180-200 registers)
    per = std::min<int64_t>(per, std::max<int64_t>(1, ntasks));
    return (int)(ceil_div64(per, WARPS) * WARPS);
}

typedef CUresult (*Corr9EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static Corr9EncodeFn corr9_encode_fn() {
    static Corr9EncodeFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<Corr9EncodeFn>(p);
    }
    return fn;
}

// (C, W, H, N) fp32 tensor map with a 32-channel x 5-column x `rows`-row box; false when the tensor cannot be mapped
static bool corr9_make_map(CUtensorMap *map, const float *act, int64_t n_img_total, int H, int W, int64_t C, int rows) {
    Corr9EncodeFn enc = corr9_encode_fn();
    if (!enc || C % 4 != 0 || C < 32 || ((uintptr_t)act & 15) != 0 || n_img_total < 1) return false;  // box = 32 channels
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_img_total};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};  // bytes, dims 1..3
    const cuuint32_t box[4] = {32u, (cuuint32_t)corr9::WC, (cuuint32_t)rows, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(act), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int RB>
static cudaError_t launch_corr9(const float *actq, const float *actx, bool same, const Corr9Geom &gm, int64_t n_img_total,
                                bool allow_tma, double *partial, int slot_stride, int slot0, cudaStream_t st) {
    using namespace corr9;
    dim3 grid((unsigned)(gm.slots / WARPS), (unsigned)ceil_div64(gm.n_ch, 32));
    CUtensorMap mq_win, mq_ctr, mx_win;
    bool tma = allow_tma && corr9_make_map(&mq_win, actq, n_img_total, gm.H, gm.W, gm.C, RB + 2);
    if (tma && !same)
        tma = corr9_make_map(&mq_ctr, actq, n_img_total, gm.H, gm.W, gm.C, RB) &&
              corr9_make_map(&mx_win, actx, n_img_total, gm.H, gm.W, gm.C, RB + 2);
    if (tma) {
        cudaError_t e = cudaFuncSetAttribute(conv_corr9_tma_kernel<RB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)Ring<RB, false>::SMEM);
        if (e != cudaSuccess) return e;
        conv_corr9_tma_kernel<RB, false><<<grid, WARPS * 32, Ring<RB, false>::SMEM, st>>>(mq_win, mq_win, gm, partial, slot_stride, slot0);
        if (!same) {
            e = cudaFuncSetAttribute(conv_corr9_tma_kernel<RB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)Ring<RB, true>::SMEM);
            if (e != cudaSuccess) return e;
            conv_corr9_tma_kernel<RB, true><<<grid, WARPS * 32, Ring<RB, true>::SMEM, st>>>(mx_win, mq_ctr, gm, partial, slot_stride, slot0);
        }
        return cudaSuccess;
    }
    conv_corr9_kernel<RB, false><<<grid, WARPS * 32, 0, st>>>(actq, actq, gm, partial, slot_stride, slot0);
    if (!same) conv_corr9_kernel<RB, true><<<grid, WARPS * 32, 0, st>>>(actq, actx, gm, partial, slot_stride, slot0);
    return cudaSuccess;
}

// Correlation sums of channels [c_first, c_first + n_ch) over images [img0, img0 + n_img): fills main slots
// [slot0, slot0 + slots) and border slots [bslot0, bslot0 + bslots) of every channel.
int conv_corr9_stage(gpfq_ctx *ctx, const float *act, const float *actq, bool same, int64_t img0, int64_t n_img,
                     int64_t n_img_total, int H, int Wd, int64_t C, int64_t c_first, int n_ch, int RB, double *partial,
                     int slot_stride, int slot0, int slots, double *bpartial, int bslot_stride, int bslot0, int bslots) {
    using namespace corr9;
    if (slots % WARPS || bslots % NCLS || bslots % WARPS)
        return gpfq_fail(ctx, GPFQ_ERR_ARG, "corr9: slot counts must be multiples of %d / %d", WARPS, NCLS);
    Corr9Geom gm;
    gm.H = H; gm.W = Wd; gm.C = C; gm.c_first = c_first; gm.n_ch = n_ch; gm.img0 = img0; gm.n_img = n_img;
    gm.slots = slots;
    // `act` / `actq` point at image 0 of a tensor of n_img_total images; the tensor maps cover all of it
    const bool allow_tma = ctx->corr_variant != 1 && n_img_total < ((int64_t)1 << 31);
    switch (RB) {
        case 8: CUDA_TRY(ctx, launch_corr9<8>(actq, act, same, gm, n_img_total, allow_tma, partial, slot_stride, slot0, ctx->stream)); break;
        case 7: CUDA_TRY(ctx, launch_corr9<7>(actq, act, same, gm, n_img_total, allow_tma, partial, slot_stride, slot0, ctx->stream)); break;
        case 4: CUDA_TRY(ctx, launch_corr9<4>(actq, act, same, gm, n_img_total, allow_tma, partial, slot_stride, slot0, ctx->stream)); break;
        default: return gpfq_fail(ctx, GPFQ_ERR_ARG, "corr9: no kernel for %d rows per band", RB);
    }
    KERNEL_CHECK(ctx);
    if (!same) ctx->launches++;
    gm.slots = bslots;
    dim3 bgrid((unsigned)(bslots / WARPS), (unsigned)ceil_div64(n_ch, 32));
    if (same) conv_corr9_border_kernel<true><<<bgrid, WARPS * 32, 0, ctx->stream>>>(actq, actq, gm, bpartial, bslot_stride, bslot0);
    else conv_corr9_border_kernel<false><<<bgrid, WARPS * 32, 0, ctx->stream>>>(actq, act, gm, bpartial, bslot_stride, bslot0);
    KERNEL_CHECK(ctx);
    return GPFQ_OK;
}

int conv_corr9_assemble_stage(gpfq_ctx *ctx, const double *partial, int slots, const double *bpartial, int bslots,
                              bool same, int n_ch, double *gram) {
    conv_corr9_assemble_kernel<<<(unsigned)n_ch, 96, 0, ctx->stream>>>(partial, slots, bpartial, bslots, same ? 1 : 0, gram);
    KERNEL_CHECK(ctx);
    return GPFQ_OK;
}
