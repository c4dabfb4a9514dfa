// PAR -- pixel-adaptive refinement (reference: utils/PAR.py) as HBM-roofline kernels for sm_100a.
//
//   par_affinity : img [B,3,H,W] -> aff [B,K,H,W]   (utils/PAR.py:67-86), K = 8 * n_dil
//   par_iterate  : planes <- sum_k aff_k * planes[clamped neighbour k]   (utils/PAR.py:88-90)
//   par_labels   : valid_key[argmax_c planes]                            (utils/affutils.py:86-87)
//
// Layout: every tensor is planar fp32, x fastest.  "Planes" are the mask channels of ALL images of
// a batch packed back to back ([P,H,W], P = sum_b C_b); plane_off[b]..plane_off[b+1] are image b's
// channels (C differs per image: background + present classes).
//
// Neighbour k = di*8 + t reads (clamp(y + DY[t]*d), clamp(x + DX[t]*d)), d = dilations[di]:
// replicate padding + one-hot dilated 3x3 conv of the reference (utils/PAR.py:10-24,39-49).
#include <type_traits>

#include "common.cuh"
#include "excel_b200.h"
#include "ptx.cuh"

namespace xl {

struct ParGeom {
    int n_dil;
    int dil[XL_PAR_MAX_DIL];
    float pos[8 * XL_PAR_MAX_DIL];  // w2 * softmax_k(-(pos_k/(std(pos)+1e-8)/w1)^2), utils/PAR.py:84-86
};


// ------------------------------------------------------------------------------------------------
// bilinear resize, align_corners=True (utils/PAR.py:67).  One thread per output pixel.
__global__ void par_resize_ac_kernel(const float* __restrict__ src, int64_t sb, int64_t sc, int64_t sy,
                                     float* __restrict__ dst, int hi, int wi, int H, int W, const int* __restrict__ img_index) {
    pdl_trigger();
    pdl_wait();
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int bc = blockIdx.z;  // b*3 + c
    if (x >= W || y >= H) return;
    const float ry = H > 1 ? (float)(hi - 1) / (float)(H - 1) : 0.f;
    const float rx = W > 1 ? (float)(wi - 1) / (float)(W - 1) : 0.f;
    const float fy = ry * y, fx = rx * x;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < hi - 1 ? 1 : 0), x1 = x0 + (x0 < wi - 1 ? 1 : 0);
    const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
    const float* p = src + (int64_t)(img_index ? img_index[bc / 3] : bc / 3) * sb + (int64_t)(bc % 3) * sc;
    const float v = hy * (hx * __ldg(p + y0 * sy + x0) + lx * __ldg(p + y0 * sy + x1)) +
                    ly * (hx * __ldg(p + y1 * sy + x0) + lx * __ldg(p + y1 * sy + x1));
    dst[((int64_t)bc * H + y) * W + x] = v;
}

// ------------------------------------------------------------------------------------------------
// Tiling shared by the two hot kernels: a block of 32x8 threads owns a TX x TY = 32x32 pixel tile
// (4 rows per thread) and stages the tile plus a halo of max(dilation) pixels in shared memory with
// the replicate padding already applied (coordinates clamped at load time), so the 8*n_dil gathers
// per pixel are conflict-free LDS (a warp reads 32 consecutive floats) with no bounds logic.
constexpr int kTX = 32, kTY = 32, kRows = 4;  // kTY == 8 * kRows

__host__ __device__ constexpr int tap_dy(int t) { return t < 3 ? -1 : (t < 5 ? 0 : 1); }
__host__ __device__ constexpr int tap_dx(int t) { return (t == 0 || t == 3 || t == 5) ? -1 : ((t == 1 || t == 6) ? 0 : 1); }

// stage `nplanes` planes (plane p at src + p*plane_stride, rows `sy` apart) into sm[p][TH][TW] with
// cp.async: every copy of the tile is in flight at once (one memory round trip instead of a chain of
// register-staged loads); rows whose whole span is inside the image and 16 B-aligned move as 16 B
// copies.  Called by the first 256 threads of the CTA (tid = 0..255); complete with
// cp_async_wait_all() + a barrier.
__device__ __forceinline__ void stage_tile_async(float* sm, const float* __restrict__ src, int64_t plane_stride, int64_t sy,
                                                 int nplanes, int x0, int y0, int halo, int TW, int TH, int H, int W,
                                                 int tid, int nwarps = 8) {
    const int lane = tid & 31, warp = tid >> 5;
    const int xl = x0 - halo;
    const bool inside = xl >= 0 && xl + TW <= W;
    // interior tiles (no clamping in x or y) whose rows are all 16 B-aligned: one pointer pair per plane, bumped per row --
    // the general loop below spends ~25 instructions per copy on clamps, 64-bit multiplies and the alignment test
    if (inside && y0 - halo >= 0 && y0 - halo + TH <= H && (sy & 3) == 0 && (plane_stride & 3) == 0 &&
        ((reinterpret_cast<uintptr_t>(src + (int64_t)(y0 - halo) * sy + xl) & 15) == 0) && (TW & 3) == 0) {
        if (lane * 4 < TW) {
            for (int p = 0; p < nplanes; ++p) {
                const float* sp = src + p * plane_stride + (int64_t)(y0 - halo + warp) * sy + xl + lane * 4;
                float* dp = sm + p * TH * TW + warp * TW + lane * 4;
                for (int r = warp; r < TH; r += nwarps, sp += (int64_t)nwarps * sy, dp += nwarps * TW) {
                    for (int c = 0; lane * 4 + c < TW; c += 128) cp_async16(dp + c, sp + c);
                }
            }
        }
        return;
    }
    for (int p = 0; p < nplanes; ++p) {
        const float* sp = src + p * plane_stride;
        float* dp = sm + p * TH * TW;
        for (int r = warp; r < TH; r += nwarps) {
            const float* row = sp + (int64_t)clampi(y0 - halo + r, 0, H - 1) * sy;
            if (inside && ((reinterpret_cast<uintptr_t>(row + xl) & 15) == 0)) {
                for (int c = lane * 4; c < TW; c += 128) cp_async16(dp + r * TW + c, row + xl + c);
            } else {
                for (int c = lane; c < TW; c += 32) cp_async4(dp + r * TW + c, row + clampi(xl + c, 0, W - 1));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Affinity (utils/PAR.py:67-86).  HBM traffic is the K output planes (192 B/pixel at K=48) against
// 12 B/pixel of image; the image tile + halo is read from L2.
//   pass 1: per channel, unbiased std of the K neighbours (centre excluded), accumulated on the
//           differences to the centre pixel -- small where the image is smooth, so the one-pass
//           sum / sum-of-squares form is well conditioned;
//   pass 2: a_k = -mean_c[(|I_k - I_0| / (std_c + 1e-8) / w1)^2] kept in registers, softmax over k,
//           + w2 * (constant positional softmax), streamed out.
// STD: the reference's dilation set (1, 2, 4, 8, 12, 24; scripts/train_voc.py:112, tools/infer_lam.py:168) is baked in, so
// that every neighbour address is an immediate offset of one base register -- the generic form spends a quarter of its
// issue slots on address arithmetic (ncu: the kernel is issue-bound, not LDS- or HBM-bound).
__device__ constexpr int kStdDil[6] = {1, 2, 4, 8, 12, 24};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int NDIL, bool STD>
__global__ void __launch_bounds__(256, 3)
par_affinity_kernel(const float* __restrict__ img, int64_t sb, int64_t sc, int64_t sy, float* __restrict__ aff,
                    int H, int W, int Wp, int halo_rt, ParGeom g, float w1, const int* __restrict__ img_index) {
    pdl_trigger();
    pdl_wait();
    constexpr int K = 8 * NDIL;
    extern __shared__ float sm[];
    const int halo = STD ? 24 : halo_rt;
    const int TW = kTX + 2 * halo, TH = kTY + 2 * halo;
    const int x0 = blockIdx.x * kTX, y0 = blockIdx.y * kTY, b = blockIdx.z;
    // image slot b of the launch reads image img_index[b] of the caller's batch (runs sorted by plane count: no gather copy)
    stage_tile_async(sm, img + (int64_t)(img_index ? img_index[b] : b) * sb, sc, sy, 3, x0, y0, halo, TW, TH, H, W,
                     threadIdx.y * 32 + threadIdx.x);
    cp_async_wait_all();
    __syncthreads();
    const int x = x0 + threadIdx.x;
    if (x >= W) return;
    const float* s0p = sm + (threadIdx.y + halo) * TW + threadIdx.x + halo;
    const int cs = TH * TW;
    const int64_t plane = (int64_t)H * Wp;  // the affinity workspace has a row pitch of Wp = round_up(W, 4)
    const float invk = 1.f / K, invk1 = 1.f / (K - 1), invw = 1.f / w1;
#pragma unroll 1
    for (int j = 0; j < kRows; ++j) {
        const int y = y0 + threadIdx.y + 8 * j;
        if (y >= H) break;
        const float* ctr = s0p + 8 * j * TW;
        const float c0 = ctr[0], c1 = ctr[cs], c2 = ctr[2 * cs];
        // The arithmetic runs on packed fp32x2 instructions (sm_100 FADD2 / FMUL2 / FFMA2 -- scalar FP32 issues at half rate and
        // this kernel is bound by the FMA pipe): taps t and t+1 of a dilation share one instruction per channel.
        float2 S0 = make_float2(0.f, 0.f), S1 = S0, S2 = S0, Q0 = S0, Q1 = S0, Q2 = S0;
#pragma unroll
        for (int di = 0; di < NDIL; ++di) {
            const int d = STD ? kStdDil[di] : g.dil[di], dW = d * TW;
#pragma unroll
            for (int t = 0; t < 8; t += 2) {
                const float* pa = ctr + tap_dy(t) * dW + tap_dx(t) * d;
                const float* pb = ctr + tap_dy(t + 1) * dW + tap_dx(t + 1) * d;
                const float2 D0 = __fadd2_rn(make_float2(pa[0], pb[0]), make_float2(-c0, -c0));
                const float2 D1 = __fadd2_rn(make_float2(pa[cs], pb[cs]), make_float2(-c1, -c1));
                const float2 D2 = __fadd2_rn(make_float2(pa[2 * cs], pb[2 * cs]), make_float2(-c2, -c2));
                S0 = __fadd2_rn(S0, D0); Q0 = __ffma2_rn(D0, D0, Q0);
                S1 = __fadd2_rn(S1, D1); Q1 = __ffma2_rn(D1, D1, Q1);
                S2 = __fadd2_rn(S2, D2); Q2 = __ffma2_rn(D2, D2, Q2);
            }
        }
        const float s0 = S0.x + S0.y, s1 = S1.x + S1.y, s2 = S2.x + S2.y;
        const float q0 = Q0.x + Q0.y, q1 = Q1.x + Q1.y, q2 = Q2.x + Q2.y;
        const float r0 = __fdividef(invw, sqrtf(fmaxf((q0 - s0 * s0 * invk) * invk1, 0.f)) + 1e-8f);
        const float r1 = __fdividef(invw, sqrtf(fmaxf((q1 - s1 * s1 * invk) * invk1, 0.f)) + 1e-8f);
        const float r2 = __fdividef(invw, sqrtf(fmaxf((q2 - s2 * s2 * invk) * invk1, 0.f)) + 1e-8f);
        // Pass 2 RE-READS the neighbours (the barrier stops the compiler from carrying the 144 differences of pass 1 across: it
        // did, at 128 registers and 2 CTAs per SM, with 212 B of spills).  Without the carry the kernel fits 85 registers: three
        // CTAs (24 warps) per SM -- the kernel is latency-bound (ncu: 6 stall cycles per issue at 4 warps per scheduler), not
        // pipe-bound, so the 50 % more warps are worth the 144 extra LDS per pixel.
        asm volatile("" ::: "memory");
        float a[K];
        float amax = -INFINITY;
        // a_k = -mean_c(t_c^2) * log2(e) (exp2 below) = sum_c d_c * (d_c * w_c), w_c = -(r_c^2) * log2(e) / 3: the channel mean, the
        // 1/w1 and the 1/(std + eps) factors fold into one scalar per channel (a true division costs a ~10-instruction slow-path
        // check 48 times per pixel for at most 1 ulp of the exponent)
        constexpr float kk = -1.4426950408889634f / 3.f;
        const float w0 = r0 * r0 * kk, w1c = r1 * r1 * kk, w2c = r2 * r2 * kk;
        const float2 W0 = make_float2(w0, w0), W1 = make_float2(w1c, w1c), W2 = make_float2(w2c, w2c);
#pragma unroll
        for (int di = 0; di < NDIL; ++di) {
            const int d = STD ? kStdDil[di] : g.dil[di], dW = d * TW;
#pragma unroll
            for (int t = 0; t < 8; t += 2) {
                const float* pa = ctr + tap_dy(t) * dW + tap_dx(t) * d;
                const float* pb = ctr + tap_dy(t + 1) * dW + tap_dx(t + 1) * d;
                const float2 D0 = __fadd2_rn(make_float2(pa[0], pb[0]), make_float2(-c0, -c0));
                const float2 D1 = __fadd2_rn(make_float2(pa[cs], pb[cs]), make_float2(-c1, -c1));
                const float2 D2 = __fadd2_rn(make_float2(pa[2 * cs], pb[2 * cs]), make_float2(-c2, -c2));
                const float2 V = __ffma2_rn(__fmul2_rn(D2, W2), D2, __ffma2_rn(__fmul2_rn(D1, W1), D1, __fmul2_rn(__fmul2_rn(D0, W0), D0)));
                a[di * 8 + t] = V.x;
                a[di * 8 + t + 1] = V.y;
                amax = fmaxf(amax, fmaxf(V.x, V.y));
            }
        }
        // softmax over k in the exp2 domain: a_k - amax <= 0, so the bare MUFU (ex2.approx.ftz: results below 2^-126 flush to 0
        // against a largest term of exactly 1) replaces exp2f's denormal-range scaling (an FSETP and two predicated FMULs
        // per neighbour); subtraction, sum and the final a * inv + pos run on packed fp32x2 instructions.
        const float2 NM = make_float2(-amax, -amax);
        float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < K; k += 2) {
            const float2 x = __fadd2_rn(make_float2(a[k], a[k + 1]), NM);
            a[k] = ex2_approx(x.x);
            a[k + 1] = ex2_approx(x.y);
            sum2 = __fadd2_rn(sum2, make_float2(a[k], a[k + 1]));
        }
        const float inv = __fdividef(1.f, sum2.x + sum2.y);   // sum in [1, K]: MUFU.RCP, no slow-path call
        const float2 INV = make_float2(inv, inv);
        // one 64-bit pointer bumped by the plane stride per store (k * plane as an index costs ~8 integer instructions per store:
        // a quarter of the kernel's instructions before this form)
        float* out = aff + ((int64_t)b * K * H + y) * Wp + x;
#pragma unroll
        for (int k = 0; k < K; k += 2) {
            const float2 v = __ffma2_rn(make_float2(a[k], a[k + 1]), INV, make_float2(g.pos[k], g.pos[k + 1]));
            // (asm: the compiler otherwise rewrites the bump as base + k * plane)
            asm volatile("st.global.f32 [%0], %1;" :: "l"(out), "f"(v.x) : "memory");
            out += plane;
            asm volatile("st.global.f32 [%0], %1;" :: "l"(out), "f"(v.y) : "memory");
            out += plane;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// One propagation step (utils/PAR.py:88-90): out[p,y,x] = sum_k aff[b,k,y,x] * in[p, nbr_k(y,x)].
// Algorithmic bytes per pixel and step: 4*(K + 2*C); the K affinity planes are the stream.
//
//  * The affinity tile [K][64][32] is streamed by TMA (cp.async.bulk.tensor.3d) through a shared-memory ring of
//    two-tap stages (16 KB), completion on mbarriers: the loads are asynchronous, cost no registers or LSU issue
//    slots, and keep 56-120 KB in flight per SM (HBM needs ~45 KB per microsecond of latency and SM).
//    The affinity workspace is internal, so its row pitch is padded to 4 floats (TMA stride rule) and
//    out-of-image elements are zero-filled by the TMA unit.
//  * A thread owns 4 consecutive pixels of one row: affinities are one LDS.128, mask neighbours one
//    LDS.128 per plane when the dilation is a multiple of 4 (4 of the 6 standard dilations), and a
//    compile-time-shifted LDS.128/LDS.64/LDS.32 combination for dilations 1 and 2.
//  * CCH mask planes (+ halo, replicate padding applied at load) are staged in shared memory; images
//    with more planes loop (the affinity re-read then comes from L2).
// Ring geometry (par_ty / par_nst / par_kg below): 16 KB TMA stages; NST-1 of them in flight per CTA.

// shared-memory loads on 32-bit shared addresses (keeps the address arithmetic 32-bit; `volatile` pins
// them behind the mbarrier waits)
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float2 lds64(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}

template <int O>
__device__ __forceinline__ void load4_shift(uint32_t p, float (&m)[4]) {
    // m[i] = smem_float[p/4 + i + O] for a 16 B-aligned byte address p
    if constexpr (O == 0) {
        const float4 v = lds128(p);
        m[0] = v.x; m[1] = v.y; m[2] = v.z; m[3] = v.w;
    } else if constexpr (O == -1) {
        const float l = lds32(p - 4);
        const float4 v = lds128(p);
        m[0] = l; m[1] = v.x; m[2] = v.y; m[3] = v.z;
    } else if constexpr (O == 1) {
        const float4 v = lds128(p);
        const float r = lds32(p + 16);
        m[0] = v.y; m[1] = v.z; m[2] = v.w; m[3] = r;
    } else if constexpr (O == -2) {
        const float2 l = lds64(p - 8);
        const float2 v = lds64(p);
        m[0] = l.x; m[1] = l.y; m[2] = v.x; m[3] = v.y;
    } else {  // O == 2
        const float2 v = lds64(p + 8);
        const float2 r = lds64(p + 16);
        m[0] = v.x; m[1] = v.y; m[2] = r.x; m[3] = r.y;
    }
}

// kG taps T0..T0+kG-1 of one dilation.  MODE: 1 / 2 = dilation 1 / 2 (compile-time shifts);
// 0 = dilation % 4 == 0 (aligned LDS.128); -1 = any dilation (scalar loads).
// as: byte address of this thread's affinities in the stage; ctr[c]: byte address of its centre pixel in
// plane c; d4 / dW4: byte offsets of one dilation step in x / y.
template <int CCH, int TY, int KG, int MODE>
__device__ __forceinline__ void par_taps(int t0, uint32_t as, const uint32_t (&ctr)[CCH], int d4, int dW4, float2 (&acc)[2][CCH]) {
#pragma unroll
    for (int tt = 0; tt < KG; ++tt) {
        const int t = t0 + tt;  // compile-time after unrolling (the caller's q loop is unrolled too)
        const float4 a4 = lds128(as + tt * TY * kTX * 4);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        const int dy = tap_dy(t), dx = tap_dx(t);
#pragma unroll
        for (int c = 0; c < CCH; ++c) {
            float m[4];
            const uint32_t row = ctr[c] + dy * dW4;
            if constexpr (MODE == 1 || MODE == 2) {
                if (dx < 0) load4_shift<-MODE>(row, m);
                else if (dx > 0) load4_shift<MODE>(row, m);
                else load4_shift<0>(row, m);
            } else if constexpr (MODE == 0) {
                load4_shift<0>(row + dx * d4, m);
            } else {
                const uint32_t p = row + dx * d4;
                m[0] = lds32(p); m[1] = lds32(p + 4); m[2] = lds32(p + 8); m[3] = lds32(p + 12);
            }
#pragma unroll
            // packed fp32x2 FMAs (sm_100 FFMA2: the scalar FFMA issues at half rate): same two IEEE fmas per instruction
            acc[0][c] = __ffma2_rn(make_float2(a[0], a[1]), make_float2(m[0], m[1]), acc[0][c]);
            acc[1][c] = __ffma2_rn(make_float2(a[2], a[3]), make_float2(m[2], m[3]), acc[1][c]);
        }
    }
}

template <int CCH, int TY, int KG>
__device__ __forceinline__ void par_taps_mode(int mode, int t0, uint32_t as, const uint32_t (&ctr)[CCH], int d4, int dW4,
                                              float2 (&acc)[2][CCH]) {
    if (mode == 0) par_taps<CCH, TY, KG, 0>(t0, as, ctr, d4, dW4, acc);
    else if (mode == 1) par_taps<CCH, TY, KG, 1>(t0, as, ctr, d4, dW4, acc);
    else if (mode == 2) par_taps<CCH, TY, KG, 2>(t0, as, ctr, d4, dW4, acc);
    else par_taps<CCH, TY, KG, -1>(t0, as, ctr, d4, dW4, acc);
}

// Replicate padding for a tile that TMA zero-filled outside the image: copy the nearest in-image
// column, then the nearest in-image row, inside shared memory.  Called by the 256 consumer threads.
__device__ __forceinline__ void fix_border(float* sm, int np, int xl, int yl, int TW, int TH, int H, int W, int tid,
                                           int nwarps) {
    const int lane = tid & 31, warp = tid >> 5;
    const int c_lo = max(0, -xl), c_hi = min(TW, W - xl) - 1;
    const int r_lo = max(0, -yl), r_hi = min(TH, H - yl) - 1;
    if (c_lo > 0 || c_hi < TW - 1) {
        for (int q = warp; q < np * TH; q += nwarps) {
            const int r = q % TH;
            if (r < r_lo || r > r_hi) continue;
            float* row = sm + q * TW;
            const float vl = row[c_lo], vr = row[c_hi];
            for (int c = lane; c < c_lo; c += 32) row[c] = vl;
            for (int c = c_hi + 1 + lane; c < TW; c += 32) row[c] = vr;
        }
    }
    bar_sync(1, nwarps * 32);
    if (r_lo > 0 || r_hi < TH - 1) {
        for (int q = warp; q < np * TH; q += nwarps) {
            const int r = q % TH;
            if (r >= r_lo && r <= r_hi) continue;
            const float* srow = sm + (q - r + (r < r_lo ? r_lo : r_hi)) * TW;
            float* row = sm + q * TW;
            for (int c = lane; c < TW; c += 32) row[c] = srow[c];
        }
    }
    bar_sync(1, nwarps * 32);
}

// TY = tile height; NST = ring depth; KG = taps per ring stage.
// STD: the reference's dilation set baked in (see par_affinity_kernel): neighbour addresses become immediate offsets and the
// six dilation rounds unroll.
template <int CCH, int TY, int NST, int KG, bool STD>
__global__ void __launch_bounds__(8 * TY + 32, 1)
par_iterate_kernel(const __grid_constant__ CUtensorMap tm_aff, const __grid_constant__ CUtensorMap tm_in, int tma_in,
                   const float* __restrict__ in, float* __restrict__ out, const int* __restrict__ plane_off, int img0,
                   int H, int W, int halo_rt, ParGeom g) {
    pdl_trigger();
    pdl_wait();
    const int halo = STD ? 24 : halo_rt;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[NST], empty_bar[NST], tile_full, tile_empty;
    constexpr int NC = 8 * TY, NW = TY / 4;  // consumer threads / warps
    // [NST][KG][TY][kTX] affinity ring (128 B-aligned TMA destinations), then the mask tile
    // (offset arithmetic, not an integer round trip, so the compiler keeps the shared address space)
    float* ring = reinterpret_cast<float*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    float* sm = ring + NST * KG * TY * kTX;
    const int TW = kTX + 2 * halo, TH = TY + 2 * halo, cs = TH * TW;
    const int x0 = blockIdx.x * kTX, y0 = blockIdx.y * TY, b = img0 + blockIdx.z;
    const int tid = threadIdx.x;
    const int K = 8 * g.n_dil, nchunk = K / KG;
    const int64_t plane = (int64_t)H * W;
    const int pbeg = plane_off[b], pend = plane_off[b + 1];
    const int npass = (pend - pbeg + CCH - 1) / CCH;
    constexpr uint32_t kStageBytes = KG * TY * kTX * sizeof(float);

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], NW);
        }
        mbar_init(&tile_full, 1);
        mbar_init(&tile_empty, NW);
        fence_barrier_init();
    }
    __syncthreads();

    if (tid >= NC) {  // ---- producer warp: one lane drives TMA
        if (tid == NC) {
            int i = 0;
            for (int pass = 0; pass < npass; ++pass) {
                if (tma_in) {  // mask tile (+halo) of this pass; zero-filled outside the image
                    if (pass > 0) mbar_wait(&tile_empty, (pass - 1) & 1);
                    mbar_arrive_expect_tx(&tile_full, (uint32_t)(CCH * cs * sizeof(float)));
                    tma_load_3d_hint(sm, &tm_in, &tile_full, x0 - halo, y0 - halo, pbeg + pass * CCH, kEvictLast);
                }
                for (int ch = 0; ch < nchunk; ++ch, ++i) {  // affinity chunks through the ring
                    const int s = i % NST;
                    if (i >= NST) mbar_wait(&empty_bar[s], ((i / NST) - 1) & 1);
                    mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
                    // the affinity planes are a pure stream: evict-first keeps the (re-read) mask planes in L2
                    tma_load_3d_hint(ring + s * KG * TY * kTX, &tm_aff, &full_bar[s], x0, y0, (int)blockIdx.z * K + ch * KG,
                                     kEvictFirst);
                }
            }
        }
        return;
    }

    // ---- consumers: thread (tx4, ty) owns pixels (y0+ty, x0+4*tx4 .. +3)
    const int tx4 = tid & 7, ty = tid >> 3;
    const bool border = x0 - halo < 0 || y0 - halo < 0 || x0 + kTX + halo > W || y0 + TY + halo > H;
    uint32_t ctr[CCH];
#pragma unroll
    for (int c = 0; c < CCH; ++c) ctr[c] = smem_u32(sm + c * cs + (ty + halo) * TW + 4 * tx4 + halo);
    const uint32_t as0 = smem_u32(ring + ty * kTX + 4 * tx4);
    int it = 0;
    for (int pass = 0; pass < npass; ++pass) {
        const int pc = pbeg + pass * CCH;
        const int np = min(CCH, pend - pc);
        if (tma_in) {
            mbar_wait(&tile_full, pass & 1);
            if (border) fix_border(sm, CCH, x0 - halo, y0 - halo, TW, TH, H, W, tid, NW);
        } else {
            if (pass > 0) bar_sync(1, NC);
            stage_tile_async(sm, in + (int64_t)pc * plane, plane, W, np, x0, y0, halo, TW, TH, H, W, tid, NW);
            cp_async_wait_all();
            bar_sync(1, NC);
        }
        float2 acc[2][CCH];   // pixels (0,1) and (2,3) of the thread's quad
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int c = 0; c < CCH; ++c) acc[i][c] = make_float2(0.f, 0.f);
        auto round = [&](int d) {   // the 8 taps of one dilation, KG per ring stage
            const int d4 = d * 4, dW4 = d * TW * 4;
            const int mode = d == 1 ? 1 : (d == 2 ? 2 : ((d & 3) == 0 ? 0 : -1));
#pragma unroll
            for (int q = 0; q < 8 / KG; ++q) {
                const int s = it % NST;
                mbar_wait(&full_bar[s], (it / NST) & 1);
                const uint32_t as = as0 + s * kStageBytes;
                par_taps_mode<CCH, TY, KG>(mode, q * KG, as, ctr, d4, dW4, acc);
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(&empty_bar[s]);  // this warp is done with stage s
                ++it;
            }
        };
        // Dilations 1 and 2 of the reference's set: the three taps of a mask row (dx = -d, 0, +d) overlap, and with four pixels per
        // thread every shared-memory load -- LDS.32, LDS.64 or LDS.128 -- costs the same four wavefronts (eight lanes of a row
        // hit eight banks).  One LDS.128 of the row is therefore kept in registers across the ring stages of the round and only
        // the d floats left / right of it are fetched per shifted tap: 9 loads per plane instead of 14.  Same taps, same
        // accumulation order per plane as the generic form: bit-identical results.
        auto round_shared = [&](auto dtag) {
            constexpr int D = decltype(dtag)::value;
            static_assert(KG == 2 && (D == 1 || D == 2), "row sharing: two taps per stage, dilation 1 or 2");
            constexpr int dW4 = D * (kTX + 2 * 24) * 4;          // STD: halo 24
            constexpr uint32_t tapB = TY * kTX * 4;                // second tap of a stage
            float4 keep[CCH];
            auto fma4 = [&](const float4& a, float m0, float m1, float m2, float m3, int c) {
                acc[0][c] = __ffma2_rn(make_float2(a.x, a.y), make_float2(m0, m1), acc[0][c]);
                acc[1][c] = __ffma2_rn(make_float2(a.z, a.w), make_float2(m2, m3), acc[1][c]);
            };
            auto left = [&](const float4& a, uint32_t row, const float4& v, int c) {    // pixels x-D .. x+3-D
                if constexpr (D == 1) { const float l = lds32(row - 4); fma4(a, l, v.x, v.y, v.z, c); }
                else { const float2 l = lds64(row - 8); fma4(a, l.x, l.y, v.x, v.y, c); }
            };
            auto right = [&](const float4& a, uint32_t row, const float4& v, int c) {   // pixels x+D .. x+3+D
                if constexpr (D == 1) { const float r = lds32(row + 16); fma4(a, v.y, v.z, v.w, r, c); }
                else { const float2 r = lds64(row + 16); fma4(a, v.z, v.w, r.x, r.y, c); }
            };
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int s = it % NST;
                mbar_wait(&full_bar[s], (it / NST) & 1);
                const uint32_t as = as0 + s * kStageBytes;
                const float4 a0 = lds128(as), a1 = lds128(as + tapB);
#pragma unroll
                for (int c = 0; c < CCH; ++c) {
                    if (q == 0) {          // taps (-1,-D), (-1,0)
                        keep[c] = lds128(ctr[c] - dW4);
                        left(a0, ctr[c] - dW4, keep[c], c);
                        fma4(a1, keep[c].x, keep[c].y, keep[c].z, keep[c].w, c);
                    } else if (q == 1) {   // taps (-1,+D), (0,-D)
                        right(a0, ctr[c] - dW4, keep[c], c);
                        keep[c] = lds128(ctr[c]);
                        left(a1, ctr[c], keep[c], c);
                    } else if (q == 2) {   // taps (0,+D), (+1,-D)
                        right(a0, ctr[c], keep[c], c);
                        keep[c] = lds128(ctr[c] + dW4);
                        left(a1, ctr[c] + dW4, keep[c], c);
                    } else {               // taps (+1,0), (+1,+D)
                        fma4(a0, keep[c].x, keep[c].y, keep[c].z, keep[c].w, c);
                        right(a1, ctr[c] + dW4, keep[c], c);
                    }
                }
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(&empty_bar[s]);
                ++it;
            }
        };
        if constexpr (STD) {
            round_shared(std::integral_constant<int, 1>{});
            round_shared(std::integral_constant<int, 2>{});
#pragma unroll
            for (int di = 2; di < 6; ++di) round(kStdDil[di]);
        } else {
#pragma unroll 1
            for (int di = 0; di < g.n_dil; ++di) round(g.dil[di]);
        }
        if (tma_in && pass + 1 < npass) {  // hand the mask tile back to the producer
            if (border) fence_proxy_async_smem();
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&tile_empty);
        }
        const int y = y0 + ty, x = x0 + 4 * tx4;
        if (y < H) {
#pragma unroll
            for (int c = 0; c < CCH; ++c) {
                if (c >= np) break;
                float* o = out + (int64_t)(pc + c) * plane + (int64_t)y * W + x;
                if (x + 3 < W && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
                    *reinterpret_cast<float4*>(o) = make_float4(acc[0][c].x, acc[0][c].y, acc[1][c].x, acc[1][c].y);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (x + i < W) o[i] = (i & 1) ? acc[i >> 1][c].y : acc[i >> 1][c].x;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// labels[b,y,x] = key[plane_off[b] + argmax_c planes[plane_off[b]+c, y, x]]; first maximum wins and
// NaN compares as the maximum (torch.argmax semantics).
__global__ void par_labels_kernel(const float* __restrict__ planes, const int* __restrict__ plane_off,
                                  const int64_t* __restrict__ key, int64_t* __restrict__ labels, int64_t hw,
                                  const int* __restrict__ out_index) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= hw) return;
    const int pbeg = plane_off[b], pend = plane_off[b + 1];
    float best = planes[(int64_t)pbeg * hw + i];
    int arg = pbeg;
    for (int p = pbeg + 1; p < pend; ++p) {
        const float v = planes[(int64_t)p * hw + i];
        if (!(best != best) && (v > best || v != v)) { best = v; arg = p; }
    }
    labels[(int64_t)(out_index ? out_index[b] : b) * hw + i] = pend > pbeg ? key[arg] : 0;
}

static int make_geom(const int* dilations, int n_dil, float w1, float w2, ParGeom* g) {
    XL_REQUIRE(n_dil >= 1 && n_dil <= XL_PAR_MAX_DIL, "PAR: n_dil=%d outside [1,%d]", n_dil, XL_PAR_MAX_DIL);
    g->n_dil = n_dil;
    const int K = 8 * n_dil;
    float pos[8 * XL_PAR_MAX_DIL];
    const float r2 = sqrtf(2.f);
    for (int di = 0; di < n_dil; ++di) {
        XL_REQUIRE(dilations[di] >= 1, "PAR: dilation %d < 1", dilations[di]);
        g->dil[di] = dilations[di];
        for (int t = 0; t < 8; ++t)
            pos[di * 8 + t] = (float)dilations[di] * ((t == 0 || t == 2 || t == 5 || t == 7) ? r2 : 1.f);
    }
    // fp32 arithmetic in the reference's order (utils/PAR.py:74,84,86)
    float mean = 0.f;
    for (int k = 0; k < K; ++k) mean += pos[k];
    mean /= K;
    float var = 0.f;
    for (int k = 0; k < K; ++k) var += (pos[k] - mean) * (pos[k] - mean);
    const float sd = K > 1 ? sqrtf(var / (K - 1)) : NAN;
    float mx = -INFINITY, e[8 * XL_PAR_MAX_DIL], sum = 0.f;
    for (int k = 0; k < K; ++k) {
        const float t = pos[k] / (sd + 1e-8f) / w1;
        e[k] = -(t * t);
        mx = fmaxf(mx, e[k]);
    }
    for (int k = 0; k < K; ++k) { e[k] = expf(e[k] - mx); sum += e[k]; }
    for (int k = 0; k < K; ++k) g->pos[k] = w2 * (e[k] / sum);
    return 0;
}

static int max_dilation(const ParGeom& g) {
    int m = 0;
    for (int i = 0; i < g.n_dil; ++i) m = g.dil[i] > m ? g.dil[i] : m;
    return (m + 3) & ~3;  // halo: multiple of 4 floats so that every tile row stays 16 B-aligned
}

static size_t tile_smem_bytes(int halo, int planes) {
    return (size_t)(kTX + 2 * halo) * (kTY + 2 * halo) * planes * sizeof(float);
}

template <typename Kern>
static int set_smem(Kern kern, size_t bytes, const char* what) {
    XL_REQUIRE(bytes <= 227 * 1024, "%s: %zu B of shared memory (dilation too large for the tile)", what, bytes);
    XL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    // whole 228 KB carve-out: three affinity CTAs of 75 KB (+1 KB reserved each) fill it exactly
    XL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    return 0;
}

static bool is_std_dilations(const ParGeom& g) {
    if (g.n_dil != 6) return false;
    for (int i = 0; i < 6; ++i)
        if (g.dil[i] != kStdDil[i]) return false;
    return true;
}

template <int NDIL>
static int launch_affinity(const float* img, int64_t sb, int64_t sc, int64_t sy, float* aff, int B, int H, int W,
                           int Wp, const ParGeom& g, float w1, const int* img_index, cudaStream_t st) {
    const int halo = max_dilation(g);
    const size_t smem = tile_smem_bytes(halo, 3);
    dim3 grid(ceil_div(W, kTX), ceil_div(H, kTY), B), block(32, 8);
    if constexpr (NDIL == 6) {
        if (is_std_dilations(g)) {
            if (int e = set_smem(par_affinity_kernel<6, true>, smem, "par_affinity")) return e;
            XL_CUDA(launch_pdl(par_affinity_kernel<6, true>, dim3(grid), dim3(block), smem, st, img, sb, sc, sy, aff, H, W, Wp, halo, g, w1, img_index));
            return check_launch("par_affinity_kernel<std>");
        }
    }
    if (int e = set_smem(par_affinity_kernel<NDIL, false>, smem, "par_affinity")) return e;
    XL_CUDA(launch_pdl(par_affinity_kernel<NDIL, false>, dim3(grid), dim3(block), smem, st, img, sb, sc, sy, aff, H, W, Wp, halo, g, w1, img_index));
    return check_launch("par_affinity_kernel");
}

// tile height / ring depth per channel count: CCH = 4 uses half-height tiles so that two CTAs still share an SM
// Tile geometry: 32 x 64 pixel tiles, one CTA (16 consumer warps + the producer warp) per SM.  Against 32 x 32 tiles at two
// CTAs per SM the halo shrinks from 5.25x to 3.4x the tile and a 4-plane pass keeps 16 warps per SM instead of 8 (measured at
// 512^2 x 16, 20 steps: 3.87 -> 3.71 ms at 2 planes, 4.47 -> 4.39 ms at 3, 6.79 -> 5.52 ms at 4).  The affinity ring holds
// 16 KB stages (two taps of the tile), as many as fit beside the mask planes.
__host__ __device__ constexpr int par_ty(int) { return 64; }
__host__ __device__ constexpr int par_nst(int cch) { return cch <= 2 ? 8 : (cch == 3 ? 6 : 4); }   // ring depth
__host__ __device__ constexpr int par_kg(int) { return 2; }                                     // taps per stage

template <int CCH>
static int launch_iterate_c(const CUtensorMap& tm, const CUtensorMap* tm_in, const float* in, float* out,
                            const int* plane_off, int img0, int nimg, int H, int W, const ParGeom& g, cudaStream_t st) {
    constexpr int TY = par_ty(CCH), NST = par_nst(CCH), KG = par_kg(CCH);
    const int halo = max_dilation(g);
    const size_t smem = (size_t)(kTX + 2 * halo) * (TY + 2 * halo) * CCH * sizeof(float) + NST * KG * TY * kTX * sizeof(float) + 128;
    dim3 grid(ceil_div(W, kTX), ceil_div(H, TY), nimg);
    if (is_std_dilations(g)) {
        if (int e = set_smem(par_iterate_kernel<CCH, TY, NST, KG, true>, smem, "par_iterate")) return e;
        XL_CUDA(launch_pdl(par_iterate_kernel<CCH, TY, NST, KG, true>, dim3(grid), dim3(8 * TY + 32), smem, st, tm, tm_in ? *tm_in : tm, tm_in != nullptr, in,
                                                                                  out, plane_off, img0, H, W, halo, g));
        return check_launch("par_iterate_kernel<std>");
    }
    if (int e = set_smem(par_iterate_kernel<CCH, TY, NST, KG, false>, smem, "par_iterate")) return e;
    XL_CUDA(launch_pdl(par_iterate_kernel<CCH, TY, NST, KG, false>, dim3(grid), dim3(8 * TY + 32), smem, st, tm, tm_in ? *tm_in : tm, tm_in != nullptr, in, out,
                                                                             plane_off, img0, H, W, halo, g));
    return check_launch("par_iterate_kernel");
}

static int launch_iterate(const CUtensorMap& tm, const CUtensorMap* tm_in, const float* in, float* out,
                          const int* plane_off, int img0, int nimg, int cch, int H, int W, const ParGeom& g,
                          cudaStream_t st) {
    switch (cch) {
        case 1: return launch_iterate_c<1>(tm, tm_in, in, out, plane_off, img0, nimg, H, W, g, st);
        case 2: return launch_iterate_c<2>(tm, tm_in, in, out, plane_off, img0, nimg, H, W, g, st);
        case 3: return launch_iterate_c<3>(tm, tm_in, in, out, plane_off, img0, nimg, H, W, g, st);
        default: return launch_iterate_c<4>(tm, tm_in, in, out, plane_off, img0, nimg, H, W, g, st);
    }
}

#define XL_NDIL_SWITCH(n, CALL)                                                    \
    switch (n) {                                                                   \
        case 1: { constexpr int ND = 1; CALL; } break;                             \
        case 2: { constexpr int ND = 2; CALL; } break;                             \
        case 3: { constexpr int ND = 3; CALL; } break;                             \
        case 4: { constexpr int ND = 4; CALL; } break;                             \
        case 5: { constexpr int ND = 5; CALL; } break;                             \
        case 6: { constexpr int ND = 6; CALL; } break;                             \
        case 7: { constexpr int ND = 7; CALL; } break;                             \
        default: { constexpr int ND = 8; CALL; } break;                            \
    }

}  // namespace xl

using namespace xl;

extern "C" int excel_par_forward(const float* img, int64_t stride_b, int64_t stride_c, int64_t stride_y, int B,
                                 int hi, int wi, int H, int W, const int* dilations, int n_dil, float w1, float w2,
                                 int num_iter, int group, float* resize_ws, float* aff_ws, const float* planes_in,
                                 float* planes_out, float* planes_tmp, const int* plane_off_dev, int total_planes,
                                 int max_c, const int* img_index_dev, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    ParGeom g;
    if (int e = make_geom(dilations, n_dil, w1, w2, &g)) return e;
    XL_REQUIRE(B >= 0 && H > 0 && W > 0 && hi > 0 && wi > 0, "PAR: bad shape B=%d img=%dx%d out=%dx%d", B, hi, wi, H, W);
    XL_REQUIRE(num_iter >= 0 && max_c >= 0, "PAR: num_iter=%d max_c=%d", num_iter, max_c);
    if (B == 0) return 0;
    if (group <= 0 || group > B) group = B;
    XL_REQUIRE(group <= 65535, "PAR: launch group %d > 65535", group);
    const bool iterate = planes_out != nullptr && num_iter > 0 && max_c > 0;
    if (iterate) {
        XL_REQUIRE(num_iter <= 1 || planes_tmp != nullptr, "PAR: num_iter=%d needs a ping-pong buffer", num_iter);
        XL_REQUIRE(planes_in != planes_out && planes_tmp != planes_out && planes_in != planes_tmp,
                   "PAR: in/out/tmp must not alias");
    }
    if (hi != H || wi != W) {
        XL_REQUIRE(resize_ws != nullptr, "PAR: image %dx%d != mask %dx%d needs a resize workspace", hi, wi, H, W);
        XL_REQUIRE(B * 3 <= 65535, "PAR: B too large for the resize launch");
        dim3 grid(ceil_div(W, 32), ceil_div(H, 8), B * 3), block(32, 8);
        XL_CUDA(launch_pdl(par_resize_ac_kernel, dim3(grid), dim3(block), 0, st, img, stride_b, stride_c, stride_y, resize_ws, hi, wi, H, W, img_index_dev));
        if (int e = check_launch("par_resize_ac_kernel")) return e;
        img = resize_ws;
        stride_y = W; stride_c = (int64_t)H * W; stride_b = 3 * stride_c;
        img_index_dev = nullptr;   // the resized copy is in slot order
    }
    // planes staged per pass: all of them up to 4; images with more planes run passes of 3 (measured at 512^2 x 16:
    // 5-6 planes take 8.1 ms / 20 steps in passes of 3 against 12.5 ms in passes of 4, whose half-height tiles need 2 passes too)
    const int cch = max_c <= 4 ? max_c : 3;
    const int Wp = (W + 3) & ~3;  // row pitch of the affinity workspace
    CUtensorMap tm;
    if (iterate) {
        const uint64_t dims[3] = {(uint64_t)W, (uint64_t)H, (uint64_t)group * 8 * n_dil};
        const uint64_t strides[2] = {(uint64_t)Wp * 4, (uint64_t)Wp * H * 4};
        const uint32_t box[3] = {kTX, (uint32_t)par_ty(cch), (uint32_t)par_kg(cch)};
        if (int e = encode_tensor_map(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, aff_ws, dims, strides, box,
                                      CU_TENSOR_MAP_SWIZZLE_NONE)) return e;
    }
    // mask tiles (+halo) also move by TMA when the planes satisfy its alignment rules (else cp.async):
    // one map per ping-pong buffer
    const float* bufs[3] = {planes_in, planes_out, planes_tmp};
    CUtensorMap tm_planes[3];
    bool tma_planes = iterate && (W % 4) == 0 && total_planes > 0;
    for (int i = 0; i < 3 && tma_planes; ++i)
        if (bufs[i] && (reinterpret_cast<uintptr_t>(bufs[i]) & 15)) tma_planes = false;
    if (tma_planes) {
        const int halo = max_dilation(g);
        const uint64_t dims[3] = {(uint64_t)W, (uint64_t)H, (uint64_t)total_planes};
        const uint64_t strides[2] = {(uint64_t)W * 4, (uint64_t)W * H * 4};
        const uint32_t box[3] = {(uint32_t)(kTX + 2 * halo), (uint32_t)(par_ty(cch) + 2 * halo), (uint32_t)cch};
        for (int i = 0; i < 3; ++i)
            if (bufs[i])
                if (int e = encode_tensor_map(&tm_planes[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, bufs[i], dims, strides, box,
                                              CU_TENSOR_MAP_SWIZZLE_NONE)) return e;
    }
    for (int b0 = 0; b0 < B; b0 += group) {
        const int nb = B - b0 < group ? B - b0 : group;
        int e = 0;
        XL_NDIL_SWITCH(n_dil, e = launch_affinity<ND>(img_index_dev ? img : img + (int64_t)b0 * stride_b, stride_b, stride_c, stride_y,
                                                      aff_ws, nb, H, W, Wp, g, w1, img_index_dev ? img_index_dev + b0 : nullptr, st));
        if (e) return e;
        if (!iterate) {  // affinity only: aff_ws holds all B images
            aff_ws += (int64_t)nb * 8 * n_dil * H * Wp;
            continue;
        }
        const float* src = planes_in;
        for (int it = 0; it < num_iter; ++it) {  // ping-pong so that the LAST step writes planes_out
            float* dst = ((num_iter - 1 - it) & 1) ? planes_tmp : planes_out;
            const CUtensorMap* tmi = !tma_planes ? nullptr
                                     : &tm_planes[src == planes_in ? 0 : (src == planes_out ? 1 : 2)];
            if ((e = launch_iterate(tm, tmi, src, dst, plane_off_dev, b0, nb, cch, H, W, g, st))) return e;
            src = dst;
        }
    }
    return 0;
}

extern "C" int excel_par_labels(const float* planes, const int* plane_off_dev, const int64_t* plane_key_dev,
                                int64_t* labels, int B, int H, int W, const int* out_index_dev, void* stream) {
    XL_REQUIRE(B >= 0 && H > 0 && W > 0, "PAR labels: bad shape");
    if (B == 0) return 0;
    XL_REQUIRE(B <= 65535, "PAR labels: B=%d > 65535", B);
    const int64_t hw = (int64_t)H * W;
    dim3 grid((unsigned)ceil_div64(hw, 256), B);
    XL_CUDA(launch_pdl(par_labels_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, planes, plane_off_dev, plane_key_dev, labels, hw, out_index_dev));
    return check_launch("par_labels_kernel");
}
