// slgemm_i8.cuh -- host interface of the int8-slice tcgen05 contraction (slgemm_i8.cu).
#pragma once
#include "i8_common.cuh"

struct SlOperand {          // a sliced matrix on the device: (n_slices, rowsP, kbytes) int8 + row exponents
    const int8_t *slices = nullptr;
    int64_t rowsP = 0, kbytes = 0;
    int n_slices = 0;
    const int32_t *e = nullptr;   // nullptr: e_const for every row
    int e_const = 0;
    CUtensorMap map;
};

struct SlProduct {          // one segment: alpha * A[a_row0 : a_row0 + M, k0 : k0 + K] B[b_row0 : b_row0 + N, k0 : k0 + K]^T
    const SlOperand *A, *B;
    int64_t a_row0, b_row0, k0, K;
    int D;
    double alpha;
};

int sl_make_operand(gpfq_ctx *ctx, SlOperand *op, const int8_t *slices, int64_t rowsP, int64_t kbytes, int n_slices, const int32_t *e,
                    int e_const);
// C[M x N] (ldc) = or += sum of the products.  M, N: valid extents; the slice tensors are zero-padded to 128-row / 128-byte tiles.
int slgemm_i8(gpfq_ctx *ctx, const SlProduct *prod, int nprod, double *C, int64_t ldc, int64_t M, int64_t N, bool accumulate);
// 5 digit slices + row exponents of `grid_rows` rows (rows >= `rows` and columns >= cols are zeros) of a row-major matrix;
// `slices` / `e` point at the first of these rows inside a (5, rowsP, colsP) slice tensor
template <typename T>
int sl_rowsplit(gpfq_ctx *ctx, const T *X, int64_t ldx, int64_t rows, int64_t cols, int32_t *e, int8_t *slices, int64_t rowsP,
                int64_t colsP, int64_t grid_rows);
int sl_transsplit(gpfq_ctx *ctx, const float *X, int64_t ldx, int64_t N0, int64_t m, int32_t *e, int *scratch, int8_t *slices,
                  int64_t mP, int64_t N0P);
// int8 level indices k' = q / h of decisions [tb, te) of `grid_rows` neurons, written at byte columns tb.. of a (rowsP, N0P) tensor
int sl_qindex(gpfq_ctx *ctx, const double *Qt, int64_t N0, int64_t nj, int64_t tb, int64_t te, double inv_h, int8_t *out,
              int64_t grid_rows, int64_t N0P, int64_t width);
