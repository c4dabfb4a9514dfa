// Fused attention for the original (q k^T) path: probabilities, head-reduced attention map AND the P V product in one
// tcgen05 kernel -- the per-head probabilities never leave the SM (reference: nn.MultiheadAttention with need_weights
// in the standard blocks and `attn_ori` / `x_ori = attn_ori @ v` in the surgery blocks,
// clip/clip_surgery_model.py:101-102,151-154,297-307).
//
// Given the softmax row statistics of the stats pass (attn_tc.cu, MODE 0) a CTA owns one 128-row query block of one
// image and walks  head group (4 heads) -> key block (128 keys) -> head:
//   S  = X_h Y_h^T                     split-fp16 operands from shared memory (TMA), 3 MMA passes, fp32 in TMEM;
//   p  = exp2(alpha s - (m + log2 l))  exactly normalised, in the epilogue warps; summed over the heads in registers
//                                      (-> the attention map the API returns) and written BACK INTO THE S TILE's
//                                      tensor memory as split fp16 (hi | lo pairs, two keys per 32-bit column);
//   O_h += P V_h                       tcgen05.mma with the A operand read from TENSOR MEMORY and V_h^T from shared
//                                      memory (3 passes: P_hi V_lo + P_lo V_hi + P_hi V_hi), fp32 accumulators of the
//                                      4 heads of the group in the other half of TMEM (4 x 64 columns).
// TMEM budget: 2 x 128 columns (S / P double buffer) + 256 columns (O of 4 heads) = 512.  The map is written once
// per (head group, key block): plain stores for the first group, same-thread read-modify-write for the others (a
// fixed summation order, so the result is deterministic).  The last key block of an image (N = 128 q + r) runs with
// the MMA N / K extents rounded up to 16 instead of 128.
#include <cuda_fp16.h>

#include "attn_tc.cuh"
#include "common.cuh"
#include "excel_b200.h"
#include "tc.cuh"

namespace xl {

namespace {

constexpr int kHG = 4;                          // heads per group (O accumulators: kHG x 64 TMEM columns)
constexpr uint32_t kTile = 128 * 64 * 2;        // 16 KB: 128 rows x 64 halves (one SWIZZLE_128B operand tile)
constexpr uint32_t kXYStage = 4 * kTile;        // X_hi, X_lo, Y_hi, Y_lo
constexpr uint32_t kVBox = 64 * 64 * 2;         // 8 KB: 64 head-dim rows x 64 keys of V^T
constexpr uint32_t kVStage = 4 * kVBox;         // hi keys 0..63, hi keys 64..127, lo keys 0..63, lo keys 64..127
constexpr int kXYStages = 2, kVStages = 2;
constexpr uint32_t kStg = 2 * 16384;            // head-sum staging, one 128 x 32 fp32 block per epilogue team
constexpr int kPvThreads = 64 + 256;
constexpr size_t kPvSmem = kXYStages * kXYStage + kVStages * kVStage + kStg + 1024 /*align*/ + 256 /*barriers*/;

__device__ __forceinline__ float ex2a(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

}  // namespace

__global__ void __launch_bounds__(kPvThreads, 1)
attn_pv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmV, const AttnPvParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* xy = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* vs = xy + kXYStages * kXYStage;
    uint8_t* stg_base = vs + kVStages * kVStage;
    uint64_t* xy_full = reinterpret_cast<uint64_t*>(stg_base + kStg);
    uint64_t* xy_empty = xy_full + kXYStages;
    uint64_t* v_full = xy_empty + kXYStages;
    uint64_t* v_empty = v_full + kVStages;
    uint64_t* s_full = v_empty + kVStages;    // [2] S tile complete (MMA -> epilogue)
    uint64_t* p_ready = s_full + 2;           // [2] P written back into the S tile (epilogue -> MMA)
    uint64_t* o_full = p_ready + 2;           // O of the head group complete (MMA -> epilogue)
    uint64_t* o_empty = o_full + 1;           // O drained (epilogue -> MMA)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nblk = (p.N + 127) / 128;
    const int ngrp = (p.H + kHG - 1) / kHG;
    const int items = p.B * nblk;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kXYStages; ++s) { mbar_init(&xy_full[s], 1); mbar_init(&xy_empty[s], 1); }
        for (int s = 0; s < kVStages; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_ready[s], 8); }
        mbar_init(o_full, 1);
        mbar_init(o_empty, 8);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_o = tmem_base + 256;

    if (warp == 0) {
        // ---- TMA producer: per step (key block kb, head h) the X / Y tiles, then that step's V^T tiles
        if (lane == 0) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmV);
            uint32_t n = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                const int rb = item % nblk, b = item / nblk;
                for (int g = 0; g < ngrp; ++g) {
                    const int hc = min(kHG, p.H - g * kHG);
                    for (int kb = 0; kb < nblk; ++kb)
                        for (int hh = 0; hh < hc; ++hh, ++n) {
                            const int h = g * kHG + hh;
                            const int s = n % kXYStages, ph = (n / kXYStages) & 1;
                            mbar_wait(&xy_empty[s], ph ^ 1);
                            uint8_t* st = xy + s * kXYStage;
                            mbar_arrive_expect_tx(&xy_full[s], kXYStage);
                            const int xr = b * p.N + rb * 128, yr = b * p.N + kb * 128;
                            const int xc = p.xo + h * 64, yc = p.yo + h * 64;
                            tma_load_2d(st, &tmQ, &xy_full[s], xc, xr);
                            tma_load_2d(st + kTile, &tmQ, &xy_full[s], xc + p.lo_off, xr);
                            tma_load_2d(st + 2 * kTile, &tmQ, &xy_full[s], yc, yr);
                            tma_load_2d(st + 3 * kTile, &tmQ, &xy_full[s], yc + p.lo_off, yr);
                            const int sv = n % kVStages, pv = (n / kVStages) & 1;
                            mbar_wait(&v_empty[sv], pv ^ 1);
                            uint8_t* vt = vs + sv * kVStage;
                            mbar_arrive_expect_tx(&v_full[sv], kVStage);
                            const int vr = b * p.D + h * 64, k0 = kb * 128;   // V^T rows = head-dim channels, columns = keys
                            tma_load_2d(vt, &tmV, &v_full[sv], k0, vr);
                            tma_load_2d(vt + kVBox, &tmV, &v_full[sv], k0 + 64, vr);
                            tma_load_2d(vt + 2 * kVBox, &tmV, &v_full[sv], p.np + k0, vr);
                            tma_load_2d(vt + 3 * kVBox, &tmV, &v_full[sv], p.np + k0 + 64, vr);
                        }
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer.  Issue order S(0), S(1), PV(0), S(2), PV(1), ...: tcgen05.mma executes in issue order, so S(j+2)
        // may overwrite the buffer PV(j) reads its P from without a further barrier.
        if (lane == 0) {
            constexpr uint32_t kIdescPV = make_idesc(64);
            uint32_t ns = 0, npv = 0, ngd = 0;
            auto issue_s = [&](int kb) {
                const int s = ns % kXYStages, buf = ns & 1;
                mbar_wait(&xy_full[s], (ns / kXYStages) & 1);
                tc_fence_after();
                const int nvalid = min(128, p.N - kb * 128);
                const uint32_t idesc = make_idesc((nvalid + 15) & ~15);
                const uint32_t tacc = tmem_base + (uint32_t)(buf * 128);
                const uint32_t st = smem_u32(xy + s * kXYStage);
                const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + kTile);
                const uint64_t b_hi = umma_desc_sw128(st + 2 * kTile), b_lo = umma_desc_sw128(st + 3 * kTile);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t adv = (uint64_t)(k * 32 >> 4);
                    umma_f16(tacc, a_hi + adv, b_lo + adv, idesc, k != 0);
                    umma_f16(tacc, a_lo + adv, b_hi + adv, idesc, 1);
                    umma_f16(tacc, a_hi + adv, b_hi + adv, idesc, 1);
                }
                umma_commit(&xy_empty[s]);
                umma_commit(&s_full[buf]);
                ++ns;
            };
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                for (int g = 0; g < ngrp; ++g) {
                    const int hc = min(kHG, p.H - g * kHG);
                    const int steps = nblk * hc;
                    issue_s(0);
                    for (int j = 0; j < steps; ++j) {
                        if (j + 1 < steps) issue_s((j + 1) / hc);
                        const int kb = j / hc, hh = j - kb * hc;
                        const int buf = npv & 1, sv = npv % kVStages;
                        mbar_wait(&p_ready[buf], (npv >> 1) & 1);
                        mbar_wait(&v_full[sv], (npv / kVStages) & 1);
                        if (j == 0) mbar_wait(o_empty, (ngd & 1) ^ 1);   // the previous group's O has been read out
                        tc_fence_after();
                        const int nvalid = min(128, p.N - kb * 128);
                        const int ksteps = (nvalid + 15) >> 4;
                        const uint32_t pbase = tmem_base + (uint32_t)(buf * 128);
                        const uint32_t vst = smem_u32(vs + sv * kVStage);
                        const uint32_t d = tmem_o + (uint32_t)(hh * 64);
                        for (int k = 0; k < ksteps; ++k) {
                            // keys 16k..16k+15: chunk k/2 of the S tile holds hi at columns +0..15, lo at +16..31 (8 columns per k step)
                            const uint32_t a_hi = pbase + (uint32_t)((k >> 1) * 32 + (k & 1) * 8), a_lo = a_hi + 16;
                            const uint32_t box = vst + (uint32_t)((k >> 2) * kVBox) + (uint32_t)((k & 3) * 32);
                            const uint64_t b_hi = umma_desc_sw128(box), b_lo = umma_desc_sw128(box + 2 * kVBox);
                            umma_f16_ts(d, a_hi, b_lo, kIdescPV, (kb | k) != 0);
                            umma_f16_ts(d, a_lo, b_hi, kIdescPV, 1);
                            umma_f16_ts(d, a_hi, b_hi, kIdescPV, 1);
                        }
                        umma_commit(&v_empty[sv]);
                        ++npv;
                    }
                    umma_commit(o_full);
                    ++ngd;
                }
            }
        }
    } else {
        // ---- epilogue warps: warp (lg, half) owns TMEM lanes 32*lg..+31 (query rows) and columns 64*half..+63 of every S tile
        const int ew = warp - 2, lg = warp & 3, half = ew >> 2;
        const int trow = lg * 32 + lane;
        const int team_bar = 1 + half;
        const uint32_t lane_addr = (uint32_t)(lg * 32) << 16;
        float* stg = reinterpret_cast<float*>(stg_base + half * 16384);
        const int tid = (ew & 3) * 32 + lane, sub = tid >> 3, c4 = (tid & 7) * 4;
        uint32_t ns = 0, ngd = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int rb = item % nblk, b = item / nblk;
            const int row = rb * 128 + trow;
            const bool row_ok = row < p.N;
            for (int g = 0; g < ngrp; ++g) {
                const int hc = min(kHG, p.H - g * kHG);
                const float* mrow = p.ml + ((int64_t)b * p.H + g * kHG) * p.N + row;   // + hh * N
                for (int kb = 0; kb < nblk; ++kb) {
                    float acc[2][32];
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                        for (int e = 0; e < 32; ++e) acc[cc][e] = 0.f;
                    float m_next = row_ok ? __ldg(mrow) : INFINITY;   // rows past N: exp2(-inf) = 0
                    for (int hh = 0; hh < hc; ++hh, ++ns) {
                        const float m_row = m_next;
                        if (row_ok && hh + 1 < hc) m_next = __ldg(mrow + (int64_t)(hh + 1) * p.N);
                        const int buf = ns & 1;
                        mbar_wait(&s_full[buf], (ns >> 1) & 1);
                        tc_fence_after();
#pragma unroll
                        for (int cc = 0; cc < 2; ++cc) {
                            const int c = half * 2 + cc;
                            const int key0 = kb * 128 + c * 32;
                            if (key0 >= p.N) continue;   // (uniform) chunk of padding keys: neither S nor P columns are used
                            const uint32_t taddr = tmem_base + lane_addr + (uint32_t)(buf * 128 + c * 32);
                            uint32_t r[32];
                            tmem_ld32(taddr, r);
                            if (key0 + 32 > p.N) {
#pragma unroll
                                for (int e = 0; e < 32; ++e)
                                    if (key0 + e >= p.N) r[e] = 0xff800000u;  // -inf -> probability 0 for the padding keys
                            }
                            uint32_t ph[16], pl[16];
#pragma unroll
                            for (int e = 0; e < 16; ++e) {
                                // 2^10 p = exp2(alpha s - (m + log2 l - 10)): one FFMA + one MUFU per element
                                const float v0 = ex2a(fmaf(p.alpha, __uint_as_float(r[2 * e]), -m_row));
                                const float v1 = ex2a(fmaf(p.alpha, __uint_as_float(r[2 * e + 1]), -m_row));
                                acc[cc][2 * e] += v0;
                                acc[cc][2 * e + 1] += v1;
                                const __half2 hh2 = __floats2half2_rn(v0, v1);
                                const float2 hf = __half22float2(hh2);
                                const __half2 ll2 = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
                                ph[e] = *reinterpret_cast<const uint32_t*>(&hh2);
                                pl[e] = *reinterpret_cast<const uint32_t*>(&ll2);
                            }
                            tmem_st16(taddr, ph);        // P_hi: keys (2e, 2e+1) of the chunk in column e
                            tmem_st16(taddr + 16, pl);   // P_lo
                        }
                        tmem_st_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&p_ready[buf]);
                    }
                    // head-reduced map of this key block: out[b,row,key] (+)= coef * 2^-10 * sum over the group's heads.
                    // Staged per 32-column chunk in the team's buffer, written with 8 lanes per 128 B row segment.
                    const float cf = p.coef * (1.f / 1024.f);
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        const int key0 = kb * 128 + (half * 2 + cc) * 32;
                        if (key0 >= p.N) continue;  // (uniform)
                        bar_sync(team_bar, 128);    // the previous chunk has been read out of the staging buffer
#pragma unroll
                        for (int e = 0; e < 32; ++e)   // row-rotated columns: conflict-free without padding
                            stg[trow * 32 + ((e + trow) & 31)] = cf * acc[cc][e];
                        bar_sync(team_bar, 128);
                        const int nrow = min(128, p.N - rb * 128);
                        if (g == 0) {
                            for (int rr = sub; rr < nrow; rr += 16) {
                                float* o = p.out + ((int64_t)b * p.N + rb * 128 + rr) * p.N + key0 + c4;
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    if (key0 + c4 + e < p.N) o[e] = stg[rr * 32 + ((c4 + e + rr) & 31)];
                            }
                        } else {
                            for (int r0 = sub; r0 < nrow; r0 += 64) {   // 4 rows per thread in flight: batches the L2 round trips
                                float old[4][4];
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const int rr = r0 + 16 * q;
                                    const float* o = p.out + ((int64_t)b * p.N + rb * 128 + rr) * p.N + key0 + c4;
#pragma unroll
                                    for (int e = 0; e < 4; ++e) old[q][e] = (rr < nrow && key0 + c4 + e < p.N) ? o[e] : 0.f;
                                }
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const int rr = r0 + 16 * q;
                                    if (rr >= nrow) break;
                                    float* o = p.out + ((int64_t)b * p.N + rb * 128 + rr) * p.N + key0 + c4;
#pragma unroll
                                    for (int e = 0; e < 4; ++e)
                                        if (key0 + c4 + e < p.N) o[e] = old[q][e] + stg[rr * 32 + ((c4 + e + rr) & 31)];
                                }
                            }
                        }
                    }
                }
                // ---- O of the group's heads: TMEM -> split fp16 -> o[b*N + row, h*64 ..] (hi) / [.. + D] (lo)
                mbar_wait(o_full, ngd & 1);
                tc_fence_after();
#pragma unroll 1
                for (int q = 0; q < 4; ++q) {      // this warp's 128 columns = heads 2*half, 2*half+1 (two 32-column chunks each)
                    const int hh = half * 2 + (q >> 1);
                    if (hh >= hc) break;
                    uint32_t r[32];
                    tmem_ld32(tmem_o + lane_addr + (uint32_t)(half * 128 + q * 32), r);
                    if (row_ok) {
                        __half* oh = p.o + ((int64_t)b * p.N + row) * (2 * p.D) + (g * kHG + hh) * 64 + (q & 1) * 32;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            __align__(16) __half2 h2[4], l2[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float v0 = __uint_as_float(r[8 * j + 2 * e]) * (1.f / 1024.f);
                                const float v1 = __uint_as_float(r[8 * j + 2 * e + 1]) * (1.f / 1024.f);
                                h2[e] = __floats2half2_rn(v0, v1);
                                const float2 hf = __half22float2(h2[e]);
                                l2[e] = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
                            }
                            *reinterpret_cast<uint4*>(oh + 8 * j) = *reinterpret_cast<const uint4*>(h2);
                            *reinterpret_cast<uint4*>(oh + p.D + 8 * j) = *reinterpret_cast<const uint4*>(l2);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(o_empty);
                ++ngd;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

int attn_pv(const CUtensorMap& tmQ, const CUtensorMap& tmV, const AttnPvParams& p, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        XL_CUDA(cudaFuncSetAttribute(attn_pv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPvSmem));
        attr_set = true;
    }
    XL_REQUIRE(p.B > 0 && p.H > 0 && p.N > 0 && p.np % 64 == 0 && p.np >= p.N && p.D == p.H * 64, "attn_pv: bad shape");
    XL_REQUIRE(p.ml && p.out && p.o, "attn_pv: missing buffers");
    const int items = p.B * ((p.N + 127) / 128);
    attn_pv_kernel<<<items < kNumSMs ? items : kNumSMs, kPvThreads, kPvSmem, st>>>(tmQ, tmV, p);
    return check_launch("attn_pv_kernel");
}

}  // namespace xl
