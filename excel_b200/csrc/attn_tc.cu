// Attention probabilities on tcgen05 without materialising the scores (reference: nn.MultiheadAttention with
// need_weights and the surgery Attention.forward, clip/clip_surgery_model.py:95-159,297-307).
//
// For X, Y in {q, k, v} the encoder needs  P[b,h] = softmax_j(scale * X_h Y_h^T)  reduced over heads (the attention map
// the API returns / the surgery new-path map) and, for the q k^T set, as the operand of P V (attn_pv.cu).  The N x N scores
// per head (857 MB per score set at 512^2 x 16) never leave the chip:
//   MODE 0 (stats): per (b, h, 128-row block) walk the key blocks, S tile = X Y^T on the tensor cores (split-fp16,
//           3 MMA passes) into TMEM; the epilogue warps keep a running row max / sum (exp2 domain) and write the single
//           row statistic  m + log2(l) - 10  [sets,B,H,N]  that turns a score into a normalised, 2^10-scaled probability.
//   MODE 1 (map): per (b, row block, key block) walk the (score set, head) pairs: S tile again, p = exp2(alpha s - stat),
//           summed in registers and written once as coef * sum p through a TMA store into the row-padded map [B,N,Npad].
// Skeleton = the persistent GEMM's (gemm_tc.cu): TMA producer warp, elected-lane tcgen05.mma issuer, FOUR TMEM
// accumulators (512 columns) so that the MMA -> epilogue round trip stays off the critical path; 16 epilogue warps, four
// per TMEM lane group (MODE 0: two groups of 8 ping-pong over the tiles, 64 columns per warp, the item's X tile resident
// in shared memory beside a Y-only ring; MODE 1: all 16 on every tile, 32 columns each).  The last key block of an image
// (N = 128 q + r) runs with the MMA N extent rounded up to 16.
#include <cuda_fp16.h>

#include <type_traits>

#include "attn_tc.cuh"
#include "common.cuh"
#include "excel_b200.h"
#include "tc.cuh"

// Timing experiments on the statistics pass (tools/experiments/stats_probe.py: private builds with -DXL_TUNING
// -DXL_TC_VARIANT=<mask>; the product build defines neither and every XL_TC(bit) is the constant false):
//   1 no MUFU.EX2   2 no tcgen05.ld   4 no S MMAs
#if defined(XL_TUNING) && defined(XL_TC_VARIANT)
#define XL_TC(bit) (((XL_TC_VARIANT) & (bit)) != 0)
#else
#define XL_TC(bit) false
#endif

namespace xl {

constexpr int kABK = 64;                                  // head dim == one 64-wide k block
constexpr int kAStages = 3;                               // MODE 1 operand ring (X and Y tiles per stage): 192 KB in flight per SM
constexpr int kAYStages = 4;                              // MODE 0: Y-only ring (32 KB stages) beside a double-buffered X tile
constexpr int kAAcc = 4;                                  // TMEM accumulators (4 x 128 columns)
constexpr uint32_t kATile = 128 * kABK * 2;               // 16 KB: one 128-row fp16 operand tile
constexpr uint32_t kAStage = 4 * kATile;                  // X_hi, X_lo, Y_hi, Y_lo
constexpr int kAEpiWarps = 16;
constexpr int kAThreads = 64 + 32 * kAEpiWarps;
constexpr uint32_t kAEpi = kAEpiWarps * 2048;             // per-warp block: stats exchange (MODE 0) / 32 x 16 fp32 TMA staging (MODE 1)
constexpr size_t kASmem = kAStages * kAStage + kAEpi + 1024 /*align*/ + 256 /*barriers*/;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int MODE>
__global__ void __launch_bounds__(kAThreads, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmO, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* ebuf = tiles + kAStages * kAStage;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ebuf + kAEpi);
    uint64_t* empty_bar = full_bar + kAYStages;
    uint64_t* acc_full = empty_bar + kAYStages;
    uint64_t* acc_empty = acc_full + kAAcc;
    uint64_t* x_full = acc_empty + kAAcc;                 // [2] MODE 0: the item's X tile (shared by all its key blocks)
    uint64_t* x_empty = x_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_empty + 2);
    // MODE 0 shared-memory plan: X double buffer (2 x 32 KB: hi | lo), then kAYStages Y stages (32 KB: hi | lo).  An SM ingests
    // ~64 B/clk from L2; not re-fetching X for each of the item's key blocks halves the bytes per S tile.
    uint8_t* xbuf = tiles;
    uint8_t* yring = tiles + 2 * 2 * kATile;
    constexpr int NST = MODE == 0 ? kAYStages : kAStages;

    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nblk = (p.N + 127) / 128;                   // row blocks == key blocks
    const int TH = p.ntypes * p.H;                        // (score set, head) pairs
    const int inner = MODE == 0 ? nblk : TH;              // consecutive tiles of one work item
    const int items = MODE == 0 ? p.B * TH * nblk : p.B * nblk * nblk;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&x_full[s], 1);
            mbar_init(&x_empty[s], 1);
        }
        for (int s = 0; s < kAAcc; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], MODE == 0 ? kAEpiWarps / 2 : kAEpiWarps);  // one arrival per epilogue warp that reads the tile
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 128 * kAAcc);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    // work item -> (b, th, rb, kb); th = type * H + head; j = position inside the item
    auto decode = [&](int item, int j, int& b, int& h, int& rb, int& kb) {
        if (MODE == 0) {   // item = (b, th, rb), j = kb
            rb = item % nblk; const int r = item / nblk; h = r % TH; b = r / TH; kb = j;
        } else {           // item = (b, rb, kb), j = th
            kb = item % nblk; const int r = item / nblk; rb = r % nblk; b = r / nblk; h = j;
        }
    };

    if (warp == 0) {
        // whole warp runs the control flow, one elected lane issues (keeps the TMA / tcgen05 operands on the uniform datapath)
        const bool leader = elect_one_sync();
        if (leader) tma_prefetch_desc(&tmQ);
        int it = 0, ni = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++ni)
            for (int j = 0; j < inner; ++j, ++it) {
                int b, h, rb, kb;
                decode(item, j, b, h, rb, kb);
                const int xr = b * p.N + rb * 128, yr = b * p.N + kb * 128;
                const int ty = h / p.H, hd = h - ty * p.H;
                const int xc = p.xo[ty] + hd * kABK, yc = p.yo[ty] + hd * kABK;
                if (MODE == 0) {
                    if (j == 0) {   // the item's X tile, once
                        const int xs = ni & 1;
                        mbar_wait(&x_empty[xs], ((ni >> 1) & 1) ^ 1);
                        if (leader) {
                            mbar_arrive_expect_tx(&x_full[xs], 2 * kATile);
                            tma_load_2d(xbuf + xs * 2 * kATile, &tmQ, &x_full[xs], xc, xr);
                            tma_load_2d(xbuf + xs * 2 * kATile + kATile, &tmQ, &x_full[xs], xc + p.lo_off, xr);
                        }
                    }
                    const int s = it % kAYStages;
                    mbar_wait(&empty_bar[s], ((it / kAYStages) & 1) ^ 1);
                    uint8_t* st = yring + s * 2 * kATile;
                    if (leader) {
                        mbar_arrive_expect_tx(&full_bar[s], 2 * kATile);
                        tma_load_2d(st, &tmQ, &full_bar[s], yc, yr);
                        tma_load_2d(st + kATile, &tmQ, &full_bar[s], yc + p.lo_off, yr);
                    }
                } else {
                    const int s = it % kAStages;
                    mbar_wait(&empty_bar[s], ((it / kAStages) & 1) ^ 1);
                    uint8_t* st = tiles + s * kAStage;
                    if (leader) {
                        mbar_arrive_expect_tx(&full_bar[s], kAStage);
                        tma_load_2d(st, &tmQ, &full_bar[s], xc, xr);
                        tma_load_2d(st + kATile, &tmQ, &full_bar[s], xc + p.lo_off, xr);
                        tma_load_2d(st + 2 * kATile, &tmQ, &full_bar[s], yc, yr);
                        tma_load_2d(st + 3 * kATile, &tmQ, &full_bar[s], yc + p.lo_off, yr);
                    }
                }
            }
    } else if (warp == 1) {
        const bool leader = elect_one_sync();
        const uint32_t tiles0 = smem_u32(tiles);
        int it = 0, ni = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++ni)
            for (int j = 0; j < inner; ++j, ++it) {
                int b, h, rb, kb;
                decode(item, j, b, h, rb, kb);
                const int buf = it % kAAcc, s = it % NST;
                mbar_wait(&acc_empty[buf], ((it / kAAcc) & 1) ^ 1);
                if (MODE == 0 && j == 0) mbar_wait(&x_full[ni & 1], (ni >> 1) & 1);
                mbar_wait(&full_bar[s], (it / NST) & 1);
                tc_fence_after();
                const uint32_t idesc = make_idesc((min(128, p.N - kb * 128) + 15) & ~15);   // padding keys are not computed
                const uint32_t tacc = tmem_base + (uint32_t)(buf * 128);
                const uint32_t xa = MODE == 0 ? tiles0 + (ni & 1) * 2 * kATile : tiles0 + s * kAStage;
                const uint32_t yb = MODE == 0 ? tiles0 + 4 * kATile + s * 2 * kATile : tiles0 + s * kAStage + 2 * kATile;
                const uint64_t a_hi = umma_desc_sw128(xa), a_lo = umma_desc_sw128(xa + kATile);
                const uint64_t b_hi = umma_desc_sw128(yb), b_lo = umma_desc_sw128(yb + kATile);
                if (leader) {
                    if (!(XL_TC(4) && MODE == 0))
#pragma unroll
                    for (int k = 0; k < kABK / 16; ++k) {
                        const uint64_t adv = (uint64_t)(k * 32 >> 4);
                        umma_f16(tacc, a_hi + adv, b_lo + adv, idesc, k != 0);
                        umma_f16(tacc, a_lo + adv, b_hi + adv, idesc, 1);
                        umma_f16(tacc, a_hi + adv, b_hi + adv, idesc, 1);
                    }
                    umma_commit(&empty_bar[s]);
                    if (MODE == 0 && j == inner - 1) umma_commit(&x_empty[ni & 1]);   // the item's X tile may be replaced
                    umma_commit(&acc_full[buf]);
                }
            }
    } else {
        // ---- epilogue: warp (lg, qt): TMEM lanes 32*lg..+31 (rows), columns 32*qt..+31 of every S tile
        const int ew = warp - 2, lg = warp & 3, qt = ew >> 2;   // lane group = warp % 4 (hardware rule), column quarter = ew / 4
        const int trow = lg * 32 + lane;
        float* wbuf = reinterpret_cast<float*>(ebuf) + ew * 512;   // this warp's 2 KB block
        int it = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            int b, h, rb, kb;
            decode(item, 0, b, h, rb, kb);
            const int row = rb * 128 + trow;
            const bool row_ok = row < p.N;
            const bool rows_empty = rb * 128 + lg * 32 >= p.N;    // this warp's 32 rows lie past the image: no TMEM reads, no exp stream
            float m_run = -INFINITY, l_run = 0.f;                 // MODE 0
            float acc[32];                                        // MODE 1: head-summed probabilities
            if (MODE == 1) {
#pragma unroll
                for (int e = 0; e < 32; ++e) acc[e] = 0.f;
            }
            // MODE 1: the statistic of the NEXT tile is fetched before this tile's wait so that the global-load latency
            // stays off the critical path
            float m_next = INFINITY;   // rows past N: exp2(-inf) = 0
            if (MODE == 1 && row_ok) m_next = __ldg(p.m + ((int64_t)b * p.H) * p.N + row);   // (type 0, head 0)
            for (int j = 0; j < inner; ++j, ++it) {
                decode(item, j, b, h, rb, kb);
                const int buf = it % kAAcc;
                const float m_row = m_next;
                if (MODE == 1 && row_ok && j + 1 < inner) {
                    const int hn = j + 1, ty = hn / p.H, hd = hn - ty * p.H;
                    m_next = __ldg(p.m + (((int64_t)ty * p.B + b) * p.H + hd) * p.N + row);
                }
                if (MODE == 0) {
                    // Two groups of 8 warps ping-pong over the tiles (group = tile parity): while one group is in the TMEM
                    // round trip / barrier hand-off of a tile the other is in the exp2 stream of the previous one.  A warp
                    // covers 64 columns (cq) of its group's tiles as two 32-column chunks.
                    const int grp = qt >> 1, cq = qt & 1;
                    if ((it & 1) != grp) continue;
                    mbar_wait(&acc_full[buf], (it / kAAcc) & 1);
                    tc_fence_after();
                    if (rows_empty) {   // (warp-uniform) none of this warp's 32 rows exists (N = 128 q + 1: the last row block holds one row)
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_empty[buf]);
                        continue;
                    }
                    // Only the LAST key block can hold padding keys: the chunk code is instantiated twice so that the per-element
                    // `key >= N ? -inf : s` selection (if-converted by the compiler) does not run for the other key blocks.
                    auto chunks = [&](auto tail_tag) {
                    constexpr bool TAIL = decltype(tail_tag)::value;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int key0 = kb * 128 + cq * 64 + c * 32;
                        uint32_t r[32];
                        if (XL_TC(2)) {
#pragma unroll
                            for (int e = 0; e < 32; ++e) r[e] = 0x3c000000u + (uint32_t)(lane + e + c);
                        } else
                        if (!TAIL || key0 < p.N)   // (uniform) else: padding keys only -- not even computed by the MMA
                            tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * 128 + cq * 64 + c * 32), r);
                        if (c == 1) {     // this warp's TMEM reads of the tile are done
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&acc_empty[buf]);
                        }
                        if (TAIL && key0 >= p.N) continue;
                        if constexpr (TAIL) {
#pragma unroll
                            for (int e = 0; e < 32; ++e)
                                if (key0 + e >= p.N) r[e] = 0xff800000u;  // -inf: padding keys drop out of max and sum
                        }
                        // running row max / sum in the exp2 domain; alpha > 0, so max(alpha*a) = alpha*max(a).
                        // Four independent max / sum chains keep the FMNMX / FADD latency off the critical path.
                        float c0 = -INFINITY, c1 = -INFINITY, c2 = -INFINITY, c3 = -INFINITY;
#pragma unroll
                        for (int e = 0; e < 32; e += 4) {
                            c0 = fmaxf(c0, __uint_as_float(r[e]));
                            c1 = fmaxf(c1, __uint_as_float(r[e + 1]));
                            c2 = fmaxf(c2, __uint_as_float(r[e + 2]));
                            c3 = fmaxf(c3, __uint_as_float(r[e + 3]));
                        }
                        const float m_new = fmaxf(m_run, p.alpha * fmaxf(fmaxf(c0, c1), fmaxf(c2, c3)));
                        if (m_new > -INFINITY) {
                            // packed fp32x2 FMA / ADD (sm_100: scalar FP32 issues at half rate); four independent sum chains
                            const float2 al = make_float2(p.alpha, p.alpha), mm = make_float2(-m_new, -m_new);
                            float2 sa = make_float2(0.f, 0.f), sb = sa;
#pragma unroll
                            for (int e = 0; e < 32; e += 4) {
                                const float2 xa = __ffma2_rn(al, make_float2(__uint_as_float(r[e]), __uint_as_float(r[e + 1])), mm);
                                const float2 xb = __ffma2_rn(al, make_float2(__uint_as_float(r[e + 2]), __uint_as_float(r[e + 3])), mm);
                                sa = __fadd2_rn(sa, XL_TC(1) ? xa : make_float2(ex2_approx(xa.x), ex2_approx(xa.y)));
                                sb = __fadd2_rn(sb, XL_TC(1) ? xb : make_float2(ex2_approx(xb.x), ex2_approx(xb.y)));
                            }
                            l_run = l_run * ex2_approx(m_run - m_new) + ((sa.x + sb.x) + (sa.y + sb.y));
                            m_run = m_new;
                        }
                    }
                    };
                    if (kb == nblk - 1) chunks(std::true_type{});
                    else chunks(std::false_type{});
                    continue;
                }
                mbar_wait(&acc_full[buf], (it / kAAcc) & 1);
                tc_fence_after();
                if (rows_empty) {       // (warp-uniform) see MODE 0
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[buf]);
                    continue;
                }
                const int key0 = kb * 128 + qt * 32;
                auto tile = [&](auto tail_tag) {   // (two instantiations: see MODE 0)
                constexpr bool TAIL = decltype(tail_tag)::value;
                if (!TAIL || key0 < p.N) {   // (uniform) else: padding keys only -- not even computed by the MMA
                    uint32_t r[32];
                    tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * 128 + qt * 32), r);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[buf]);   // this warp's TMEM reads of the tile are done
                    if constexpr (TAIL) {
#pragma unroll
                        for (int e = 0; e < 32; ++e)
                            if (key0 + e >= p.N) r[e] = 0xff800000u;  // -inf -> probability 0 for the padding keys
                    }
                    // 2^10 p = exp2(alpha*s - (m + log2 l - 10)): one FFMA + one MUFU + one FADD per element
#pragma unroll
                    for (int e = 0; e < 32; ++e) acc[e] += ex2_approx(fmaf(p.alpha, __uint_as_float(r[e]), -m_row));
                } else {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[buf]);
                }
                };
                if (kb == nblk - 1) tile(std::true_type{});
                else tile(std::false_type{});
            }
            if (MODE == 0) {
                // merge the four column quarters of each row; write m + log2(l) - 10 (what the map / P V passes subtract)
                float* xch = reinterpret_cast<float*>(ebuf);   // [3][2][128]
                bar_sync(3, 32 * kAEpiWarps);                  // the previous item's merge has been read
                if (qt > 0) { xch[(qt - 1) * 256 + trow] = m_run; xch[(qt - 1) * 256 + 128 + trow] = l_run; }
                bar_sync(3, 32 * kAEpiWarps);
                if (qt == 0 && row_ok) {
                    float mf = m_run;
#pragma unroll
                    for (int q = 0; q < 3; ++q) mf = fmaxf(mf, xch[q * 256 + trow]);
                    float lf = l_run * ex2_approx(m_run - mf);
#pragma unroll
                    for (int q = 0; q < 3; ++q) lf += xch[q * 256 + 128 + trow] * ex2_approx(xch[q * 256 + trow] - mf);
                    const int ty = h / p.H, hd = h - ty * p.H;
                    p.m[(((int64_t)ty * p.B + b) * p.H + hd) * p.N + row] = mf + __log2f(lf) - 10.f;
                }
            } else {
                // head-reduced map block [32 rows x 32 keys] of this warp: two 16-column halves through the warp's staging
                // block (SWIZZLE_64B rows) and out by TMA store (clipped at the row / padded-column extents of the map)
                const int key0 = kb * 128 + qt * 32;
                if (p.out_split) {
                    // ... or straight into the split-fp16 A operand of the new-path P V GEMM (no fp32 map, no split pass):
                    // 2^10 p = hi + lo, [32 rows x 16 keys] halves each (32 B rows), columns [0, np) hi | [np, 2 np) lo; the
                    // padding keys N .. np-1 are written as zeros (K padding of the GEMM).
                    if (key0 < p.np && rb * 128 + lg * 32 < p.N) {
                        uint8_t* wb = reinterpret_cast<uint8_t*>(wbuf);
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf) {
                            if (lane == 0) tma_store_wait_read<0>();
                            __syncwarp();
#pragma unroll
                            for (int q = 0; q < 2; ++q) {
                                __align__(16) __half2 h2[4], l2[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float v0 = p.coef * acc[hf * 16 + 8 * q + 2 * e], v1 = p.coef * acc[hf * 16 + 8 * q + 2 * e + 1];
                                    h2[e] = __floats2half2_rn(v0, v1);
                                    const float2 hf2 = __half22float2(h2[e]);
                                    l2[e] = __floats2half2_rn(v0 - hf2.x, v1 - hf2.y);
                                }
                                *reinterpret_cast<uint4*>(wb + lane * 32 + q * 16) = *reinterpret_cast<const uint4*>(h2);
                                *reinterpret_cast<uint4*>(wb + 1024 + lane * 32 + q * 16) = *reinterpret_cast<const uint4*>(l2);
                            }
                            fence_proxy_async_smem();
                            __syncwarp();
                            if (lane == 0) {
                                tma_store_3d(&tmO, wb, key0 + hf * 16, rb * 128 + lg * 32, b);
                                tma_store_3d(&tmO, wb + 1024, p.np + key0 + hf * 16, rb * 128 + lg * 32, b);
                                tma_store_commit();
                            }
                        }
                    }
                } else if (key0 < p.N && rb * 128 + lg * 32 < p.N) {
                    const float cf = p.coef * (1.f / 1024.f);
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        if (key0 + hf * 16 >= p.N) break;   // (uniform)
                        if (lane == 0) tma_store_wait_read<0>();   // the previous block has left the staging buffer
                        __syncwarp();
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            *reinterpret_cast<float4*>(wbuf + lane * 16 + ((q ^ ((lane >> 1) & 3)) << 2)) =
                                make_float4(cf * acc[hf * 16 + 4 * q], cf * acc[hf * 16 + 4 * q + 1], cf * acc[hf * 16 + 4 * q + 2],
                                            cf * acc[hf * 16 + 4 * q + 3]);
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_3d(&tmO, wbuf, key0 + hf * 16, rb * 128 + lg * 32, b);
                            tma_store_commit();
                        }
                    }
                }
            }
        }
        if (MODE == 1 && lane == 0) tma_store_wait_all<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 128 * kAAcc);
    }
}

int attn_scores(const CUtensorMap& tmQ, const AttnParams& p, cudaStream_t st, bool stats_only) {
    static unsigned long long attr_once = 0;
    if (first_use_on_device(attr_once)) {
        XL_CUDA(cudaFuncSetAttribute(attn_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kASmem));
        XL_CUDA(cudaFuncSetAttribute(attn_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kASmem));
    }
    XL_REQUIRE(p.B > 0 && p.H > 0 && p.N > 0 && p.ntypes >= 1 && p.ntypes <= 3, "attn_scores: bad shape");
    XL_REQUIRE(p.m && (p.out || p.out_split || stats_only), "attn_scores: missing buffers");
    const int nblk = (p.N + 127) / 128;
    const int items0 = p.B * p.ntypes * p.H * nblk, items1 = p.B * nblk * nblk;
    XL_CUDA(launch_pdl(attn_tc_kernel<0>, dim3(items0 < kNumSMs ? items0 : kNumSMs), dim3(kAThreads), kASmem, st, tmQ, tmQ, p));
    if (int e = check_launch("attn_tc_kernel<stats>")) return e;
    if (stats_only) return 0;
    CUtensorMap tmO;
    if (p.out_split) {
        // split-fp16 map [B, N, 2 np] (hi | lo): 16-column x 32-row fp16 boxes (32 B rows, no swizzle)
        XL_REQUIRE(p.np % 64 == 0 && p.np >= p.N, "attn_scores: bad split pitch");
        const uint64_t dims[3] = {(uint64_t)2 * p.np, (uint64_t)p.N, (uint64_t)p.B};
        const uint64_t strides[2] = {(uint64_t)2 * p.np * 2, (uint64_t)2 * p.np * 2 * p.N};
        const uint32_t box[3] = {16, 32, 1};
        if (int e = encode_tensor_map(&tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, p.out_split, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return e;
    } else if (int e = make_map_store(&tmO, p.out, p.B, p.N)) return e;
    XL_CUDA(launch_pdl(attn_tc_kernel<1>, dim3(items1 < kNumSMs ? items1 : kNumSMs), dim3(kAThreads), kASmem, st, tmQ, tmO, p));
    return check_launch("attn_tc_kernel<map>");
}

// tensor map of a row-padded map [B, N, Npad] fp32 for 16-column x 32-row warp blocks (SWIZZLE_64B staging)
int make_map_store(CUtensorMap* tm, float* base, int B, int N) {
    const int Npad = (N + 3) & ~3;
    const uint64_t dims[3] = {(uint64_t)Npad, (uint64_t)N, (uint64_t)B};
    const uint64_t strides[2] = {(uint64_t)Npad * 4, (uint64_t)Npad * 4 * N};
    const uint32_t box[3] = {16, 32, 1};
    return encode_tensor_map(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
}

}  // namespace xl
