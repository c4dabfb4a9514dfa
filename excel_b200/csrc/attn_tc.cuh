// Interface of the tcgen05 attention-probability kernels (attn_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace xl {

// P[b,h] = softmax_j(alpha' * X_h Y_h^T), X/Y = column blocks (xo, yo; 64 columns per head) of the split-fp16
// qkv matrix [B*N, 2*lo_off] (hi | lo).  alpha = scale * log2(e) (the kernels work in the exp2 domain).
struct AttnParams {
    int B, H, N;
    int ntypes;             // score sets summed into `out` by one launch (1, or 3 for the surgery new path: qq, kk, vv)
    int xo[3], yo[3];       // column offsets of X / Y per score set
    int lo_off;
    float alpha;
    float* m;               // [ntypes,B,H,N] row statistic m + log2(l) - 10: written by the stats pass, read by the map pass
    float* out;             // row-padded map [B,N,Npad], Npad = round_up(N,4): coef * sum_types sum_h P[b,h]
    float coef;
    __half* out_split;      // instead of `out`: the map x 2^10 as a split-fp16 GEMM operand [B*N, 2*np] (hi | lo), np = round_up(N,64)
    int np;
};

// stats pass (always) + map pass (skipped with stats_only)
int attn_scores(const CUtensorMap& tmQ, const AttnParams& p, cudaStream_t st, bool stats_only = false);
int make_map_store(CUtensorMap* tm, float* base, int B, int N);

// Fused probabilities + head-reduced map + P V for one score set (attn_pv.cu); needs the stats pass's `ml`.
struct AttnPvParams {
    int B, H, N, D;         // D = 64 H
    int xo, yo, vo;         // column offsets of X (queries) / Y (keys) / V (values) in the split qkv matrix
    int lo_off;
    float alpha;            // scale * log2(e)
    const float* ml;        // [B,H,N]  m + log2(l) - 10
    float* out;             // [B,N,Npad], Npad = round_up(N,4): coef * sum_h P[b,h], written by TMA store / reduce-add
    float coef;
    int hpi;                // heads per group (1, 2 or 4) and
    int gsplit;             // 1: one work item per (image, query block, head group) -- small batches; the groups' partial maps go
    float* part;            //    to part [ngrp][B,N,Npad] and are summed into `out` in a fixed order (attn_pv_plan chooses)
    __half* o;              // split-fp16 [B*N, 2*D] (hi | lo): O[b, :, h*64..] = P[b,h] V[b,h]
};
// tmQ: split qkv [B*N, 6D], box 64 x 128 rows (V is consumed in place as an MN-major B operand: no transpose)
int attn_pv(const CUtensorMap& tmQ, const AttnPvParams& p, cudaStream_t st);
void attn_pv_plan(int B, int H, int N, int* hpi, int* gsplit);

}  // namespace xl
