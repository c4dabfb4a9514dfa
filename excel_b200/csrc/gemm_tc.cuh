// Interface of the tcgen05 split-fp16 GEMM engine (gemm_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace xl {

// C[z][m,n] = act(alpha * sum_k A[m,k] B[n,k] + bias[n]) (+ residual[z][m,n]);  z = (z1, z2) two-level batch.
// A and B are split-fp16 matrices addressed through their tensor maps: tile (k, row) of batch z starts at
// column a_col1*z1 + a_col2*z2 (+ a_lo_off for the lo half), row a_row1*z1 + a_row2*z2 (same for B).
struct TcParams {
    int M, N, kblocks;          // kblocks = Kp / 64
    int a_lo_off, b_lo_off;     // column distance hi -> lo (elements)
    int nb2;                    // size of the inner batch level
    int a_row0, a_row1, a_row2, a_col0, a_col1, a_col2;   // row / column of batch z: x0 + z1*x1 + z2*x2
    int b_row0, b_row1, b_row2, b_col0, b_col1, b_col2;
    float* C;                   // fp32 output (may be null)
    int64_t ldc, c1, c2;
    const float* bias;          // [N] or null; batch z1 uses bias + z1 * bias1
    int64_t bias1;
    const float* residual;      // same layout as C, or null (may alias C)
    int c_add;                  // C += ... : the fp32 tile leaves as a TMA reduce-add (the add is performed in L2) -- the in-place
                                // residual update without reading the residual in the epilogue; needs the TMA-store layout rules
    float alpha;
    int act;                    // 0 none, 1 QuickGELU
    __half* Cs;                 // split-fp16 output (may be null): hi at Cs, lo at Cs + cs_lo_off
    int64_t lds, cs1, cs2;
    int cs_lo_off;
    int b_mn;                   // B is stored [K rows, N columns] (MN-major, e.g. V inside the qkv matrix): b_row* address the K
                                // dimension, b_col* the N dimension; its tensor map has a 64-column x 64-row box
};

// tensor map of a split-fp16 operand [rows, cols_total] with row pitch ld_elems (box 64 k x box_rows, SWIZZLE_128B);
// box_rows = 128 for an A operand, = the tile N (64 or 128) for a B operand
int make_operand_map(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols_total, int64_t ld_elems, int box_rows);
int tc_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, int batch, int bn, cudaStream_t st);
// 256-wide tiles on CTA pairs (gemm_tc2.cu: tcgen05 cta_group::2, M = 256): chosen by tc_gemm when the shape fills the chip
int tc_gemm_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmS, const TcParams& p,
                 int batch, bool res, cudaStream_t st);
int tc_pick_bn(int64_t M, int N, int batch);   // tile width (64 / 128 / 256) by tile count
// x [rows, cols] fp32 (pitch ldx) -> out [rows, 2*Kp] fp16 (hi | lo), zero padded; Kp % 64 == 0
int split_f16(const float* x, int64_t ldx, int rows, int cols, int Kp, __half* out, cudaStream_t st, float scale = 1.f);

}  // namespace xl
