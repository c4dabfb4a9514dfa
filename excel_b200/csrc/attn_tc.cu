// Attention probabilities on tcgen05 without materialising the scores (reference: nn.MultiheadAttention with
// need_weights and the surgery Attention.forward, clip/clip_surgery_model.py:95-159,297-307).
//
// For X, Y in {q, k, v} the encoder needs  P[b,h] = softmax_j(scale * X_h Y_h^T)  twice over: as the operand of the
// P V product (per head) and reduced over heads (the attention map the API returns / the new-path map).  The
// N x N scores per head (857 MB per score set at 512^2 x 16) never leave the chip:
//   MODE 0 (stats): per (b, h, 128-row block) walk the key blocks, S tile = X Y^T on the tensor cores (split-fp16,
//           3 MMA passes) into TMEM; the epilogue warps keep a running row max / sum (exp2 domain) -> m, l [B,H,N].
//   MODE 1 (probs): per (b, row block, key block) walk the HEADS: S tile again, p = exp2(s - m) / l exactly
//           normalised; (a) written as the split-fp16 P operand through a TMA store, (b) summed over heads in
//           registers and written once as coef * sum_h p (+ the previous content) -> out [B,N,N].
// Skeleton = the persistent GEMM's (gemm_tc.cu): TMA producer warp, single-thread tcgen05.mma issuer, double-
// buffered TMEM accumulators; 8 epilogue warps (two per TMEM lane group, each taking half of the 128 columns).
#include <cuda_fp16.h>

#include "attn_tc.cuh"
#include "common.cuh"
#include "excel_b200.h"
#include "tc.cuh"

namespace xl {

constexpr int kABK = 64;                                  // head dim == one 64-wide k block
constexpr int kAStages = 3;                               // the kernels are TMA-latency bound: 128 KB in flight per SM
constexpr uint32_t kATile = 128 * kABK * 2;               // 16 KB: one 128-row fp16 operand tile
constexpr uint32_t kAStage = 4 * kATile;                  // X_hi, X_lo, Y_hi, Y_lo
constexpr uint32_t kAEpi = 2 * 16384;                     // one TMA-store staging buffer per epilogue team (+1 KB spill-over)
// stats pass: 16 epilogue warps (four per TMEM lane group, 32 columns each) hide the MUFU / FMNMX latency chains;
// probs pass: 8 (two per lane group, 64 columns each -- its epilogue carries 64 head-sum accumulators per thread)
__host__ __device__ constexpr int attn_epi_warps(int mode) { return mode == 0 ? 16 : 8; }
__host__ __device__ constexpr int attn_threads(int mode) { return 64 + 32 * attn_epi_warps(mode); }
constexpr size_t kASmem = kAStages * kAStage + kAEpi + 1024 /*row stats exchange / staging tail*/ + 1024 /*align*/ + 256 /*barriers*/;
constexpr float kProbScaleA = 1024.f;                     // 2^10; must match vit.cu: kProbScale

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int MODE>
__global__ void __launch_bounds__(attn_threads(MODE), 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmP, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* ebuf = tiles + kAStages * kAStage;
    float* xch = reinterpret_cast<float*>(ebuf);          // stats mode: [NPART-1][2][128] (m, l) of the other column parts
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ebuf + kAEpi + 1024);
    uint64_t* empty_bar = full_bar + kAStages;
    uint64_t* acc_full = empty_bar + kAStages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nblk = (p.N + 127) / 128;                   // row blocks == key blocks
    const int TH = p.ntypes * p.H;                        // (score set, head) pairs
    const int inner = MODE == 0 ? nblk : TH;              // consecutive tiles of one work item
    const int items = MODE == 0 ? p.B * TH * nblk : p.B * nblk * nblk;
    constexpr uint32_t kIdesc = make_idesc(128);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kAStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], attn_epi_warps(MODE));  // one arrival per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

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
        int it = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x)
            for (int j = 0; j < inner; ++j, ++it) {
                int b, h, rb, kb;
                decode(item, j, b, h, rb, kb);
                const int s = it % kAStages;
                mbar_wait(&empty_bar[s], ((it / kAStages) & 1) ^ 1);
                uint8_t* st = tiles + s * kAStage;
                const int xr = b * p.N + rb * 128, yr = b * p.N + kb * 128;
                const int ty = h / p.H, hd = h - ty * p.H;
                const int xc = p.xo[ty] + hd * kABK, yc = p.yo[ty] + hd * kABK;
                if (leader) {
                    mbar_arrive_expect_tx(&full_bar[s], kAStage);
                    tma_load_2d(st, &tmQ, &full_bar[s], xc, xr);
                    tma_load_2d(st + kATile, &tmQ, &full_bar[s], xc + p.lo_off, xr);
                    tma_load_2d(st + 2 * kATile, &tmQ, &full_bar[s], yc, yr);
                    tma_load_2d(st + 3 * kATile, &tmQ, &full_bar[s], yc + p.lo_off, yr);
                }
            }
    } else if (warp == 1) {
        const bool leader = elect_one_sync();
        const uint32_t tiles0 = smem_u32(tiles);
        int it = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x)
            for (int j = 0; j < inner; ++j, ++it) {
                const int buf = it & 1, s = it % kAStages;
                mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
                mbar_wait(&full_bar[s], (it / kAStages) & 1);
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(buf * 128);
                const uint32_t st = tiles0 + s * kAStage;
                const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + kATile);
                const uint64_t b_hi = umma_desc_sw128(st + 2 * kATile), b_lo = umma_desc_sw128(st + 3 * kATile);
                if (leader) {
#pragma unroll
                    for (int k = 0; k < kABK / 16; ++k) {
                        const uint64_t adv = (uint64_t)(k * 32 >> 4);
                        umma_f16(tacc, a_hi + adv, b_lo + adv, kIdesc, k != 0);
                        umma_f16(tacc, a_lo + adv, b_hi + adv, kIdesc, 1);
                        umma_f16(tacc, a_hi + adv, b_hi + adv, kIdesc, 1);
                    }
                    umma_commit(&empty_bar[s]);
                    umma_commit(&acc_full[buf]);
                }
            }
    } else {
        // ---- epilogue: warp (lg, half): TMEM lanes 32*lg..+31 (rows), columns 32*CPW*half..+32*CPW-1 of every S tile
        constexpr int NPART = attn_epi_warps(MODE) / 4, CPW = 4 / NPART;   // column parts; 32-column chunks per warp
        const int ew = warp - 2, lg = warp & 3, half = ew >> 2;   // part = ew / 4 (lane group = warp % 4)
        const int trow = lg * 32 + lane;
        const int team_bar = 1 + half;
        const bool leader = (ew & 3) == 0 && lane == 0;           // first warp of the team
        uint8_t* tbuf = ebuf + half * 16384;
        int it = 0, ck = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            int b, h, rb, kb;
            decode(item, 0, b, h, rb, kb);
            const int row = rb * 128 + trow;
            const bool row_ok = row < p.N;
            float m_run = -INFINITY, l_run = 0.f;                 // MODE 0
            float acc[2][32];                                     // MODE 1: head-summed probabilities
            if (MODE == 1) {
#pragma unroll
                for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                    for (int e = 0; e < 32; ++e) acc[cc][e] = 0.f;
            }
            float m_next = INFINITY;   // rows past N: exp2(-inf) = 0
            if (MODE == 1 && row_ok) m_next = __ldg(p.m + ((int64_t)b * p.H) * p.N + row);   // (type 0, head 0)
            for (int j = 0; j < inner; ++j, ++it) {
                decode(item, j, b, h, rb, kb);
                const int buf = it & 1;
                // m_row = m + log2(l) - 10 (stats pass): folds 1/l and the 2^10 operand scale.  The value for the NEXT
                // tile is fetched before this tile's wait so that the global-load latency stays off the critical path.
                const float m_row = m_next;
                if (MODE == 1 && row_ok && j + 1 < inner) {
                    const int hn = j + 1, ty = hn / p.H, hd = hn - ty * p.H;
                    m_next = __ldg(p.m + (((int64_t)ty * p.B + b) * p.H + hd) * p.N + row);
                }
                mbar_wait(&acc_full[buf], (it >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int cc = 0; cc < CPW; ++cc) {
                    const int c = half * CPW + cc;
                    uint32_t r[32];
                    tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * 128 + c * 32), r);
                    if (cc == CPW - 1) {  // this warp's TMEM reads of the tile are done
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_empty[buf]);
                    }
                    const int key0 = kb * 128 + c * 32;
                    const bool full = key0 + 32 <= p.N;   // (uniform) no key of this chunk is padding
                    if (MODE == 0) {
                        // running row max / sum in the exp2 domain; alpha > 0, so max(alpha*a) = alpha*max(a).
                        // Four independent max / sum chains keep the FMNMX / FADD latency off the critical path.
                        if (!full) {
#pragma unroll
                            for (int e = 0; e < 32; ++e)
                                if (key0 + e >= p.N) r[e] = 0xff800000u;  // -inf: padding keys drop out of max and sum
                        }
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
                            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                            for (int e = 0; e < 32; e += 4) {
                                s0 += ex2_approx(fmaf(p.alpha, __uint_as_float(r[e]), -m_new));
                                s1 += ex2_approx(fmaf(p.alpha, __uint_as_float(r[e + 1]), -m_new));
                                s2 += ex2_approx(fmaf(p.alpha, __uint_as_float(r[e + 2]), -m_new));
                                s3 += ex2_approx(fmaf(p.alpha, __uint_as_float(r[e + 3]), -m_new));
                            }
                            l_run = l_run * ex2_approx(m_run - m_new) + ((s0 + s1) + (s2 + s3));
                            m_run = m_new;
                        }
                    } else {
                        // pr = 2^10 * p = exp2(alpha*s - (m + log2 l - 10)): one FFMA + one MUFU per element
                        float pr[32];
                        if (!full) {
#pragma unroll
                            for (int e = 0; e < 32; ++e)
                                if (key0 + e >= p.N) r[e] = 0xff800000u;  // -inf -> probability 0 for the padding keys
                        }
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            pr[e] = ex2_approx(fmaf(p.alpha, __uint_as_float(r[e]), -m_row));
                            acc[cc][e] += pr[e];
                        }
                        if (p.write_p && key0 < p.np) {  // (uniform) P operand tile: split fp16, scaled, via TMA store
                            uint8_t* sb = tbuf;
                            if (leader) tma_store_wait_read<0>();   // the previous store has finished reading the buffer
                            bar_sync(team_bar, 128);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                __align__(16) __half2 hh[4], ll[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float v0 = pr[8 * q + 2 * e], v1 = pr[8 * q + 2 * e + 1];
                                    hh[e] = __floats2half2_rn(v0, v1);
                                    const float2 hf = __half22float2(hh[e]);
                                    ll[e] = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
                                }
                                const int off = trow * 64 + ((q ^ ((trow >> 1) & 3)) << 4);   // SWIZZLE_64B
                                *reinterpret_cast<uint4*>(sb + off) = *reinterpret_cast<const uint4*>(hh);
                                *reinterpret_cast<uint4*>(sb + 8192 + off) = *reinterpret_cast<const uint4*>(ll);
                            }
                            fence_proxy_async_smem();
                            bar_sync(team_bar, 128);
                            if (leader) {
                                tma_store_3d(&tmP, sb, key0, rb * 128, b * p.H + h);
                                tma_store_3d(&tmP, sb + 8192, p.np + key0, rb * 128, b * p.H + h);
                                tma_store_commit();
                            }
                            ++ck;
                        }
                    }
                }
            }
            if (MODE == 0) {
                // merge the column parts of each row; write m + log2(l) - 10 (what the probs pass subtracts)
                if (half > 0) { xch[(half - 1) * 256 + trow] = m_run; xch[(half - 1) * 256 + 128 + trow] = l_run; }
                bar_sync(3, 32 * attn_epi_warps(MODE));
                if (half == 0 && row_ok) {
                    float mf = m_run;
#pragma unroll
                    for (int q = 0; q < NPART - 1; ++q) mf = fmaxf(mf, xch[q * 256 + trow]);
                    float lf = l_run * ex2_approx(m_run - mf);
#pragma unroll
                    for (int q = 0; q < NPART - 1; ++q) lf += xch[q * 256 + 128 + trow] * ex2_approx(xch[q * 256 + trow] - mf);
                    const int ty = h / p.H, hd = h - ty * p.H;
                    p.m[(((int64_t)ty * p.B + b) * p.H + hd) * p.N + row] = mf + __log2f(lf) - 10.f;
                }
                bar_sync(3, 32 * attn_epi_warps(MODE));
            } else {
                // head-reduced map out[b,row,key] = coef * sum p: staged per 32-column chunk in the team's buffer
                // and written with 8 lanes per row segment -> 128 B coalesced stores
                float* stg = reinterpret_cast<float*>(tbuf);
                const int tid = (ew & 3) * 32 + lane, sub = tid >> 3, c4 = (tid & 7) * 4;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int key0 = kb * 128 + (half * 2 + cc) * 32;
                    if (key0 >= p.N) continue;  // (uniform)
                    if (leader) tma_store_wait_read<0>();  // P stores may still be reading the staging buffers
                    bar_sync(team_bar, 128);
#pragma unroll
                    for (int e = 0; e < 32; ++e)   // row-rotated columns: conflict-free without padding (16 KB exactly)
                        stg[trow * 32 + ((e + trow) & 31)] = (p.coef * (1.f / kProbScaleA)) * acc[cc][e];
                    bar_sync(team_bar, 128);
                    for (int rr = sub; rr < 128; rr += 16) {
                        const int orow = rb * 128 + rr;
                        if (orow >= p.N) break;
                        float* o = p.out + ((int64_t)b * p.N + orow) * p.N + key0 + c4;
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (key0 + c4 + e < p.N) o[e] = stg[rr * 32 + ((c4 + e + rr) & 31)];
                    }
                }
                bar_sync(team_bar, 128);  // staging buffers are reused by the next item's P tiles
            }
        }
        if (MODE == 1 && leader) tma_store_wait_read<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

int attn_scores(const CUtensorMap& tmQ, const AttnParams& p, __half* Ps, cudaStream_t st, bool stats_only) {
    static bool attr_set = false;
    if (!attr_set) {
        XL_CUDA(cudaFuncSetAttribute(attn_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kASmem));
        XL_CUDA(cudaFuncSetAttribute(attn_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kASmem));
        attr_set = true;
    }
    XL_REQUIRE(p.B > 0 && p.H > 0 && p.N > 0 && p.np % 64 == 0 && p.np >= p.N && p.ntypes >= 1 && p.ntypes <= 3 &&
                   (!p.write_p || p.ntypes == 1), "attn_scores: bad shape");
    XL_REQUIRE(p.m && (p.out || stats_only), "attn_scores: missing buffers");
    const int nblk = (p.N + 127) / 128;
    CUtensorMap tmP = tmQ;
    if (p.write_p) {
        XL_REQUIRE(Ps != nullptr, "attn_scores: write_p without a P buffer");
        const uint64_t dims[3] = {(uint64_t)2 * p.np, (uint64_t)p.N, (uint64_t)p.B * p.H};
        const uint64_t strides[2] = {(uint64_t)2 * p.np * 2, (uint64_t)2 * p.np * 2 * p.N};
        const uint32_t box[3] = {32, 128, 1};
        if (int e = encode_tensor_map(&tmP, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, Ps, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B))
            return e;
    }
    const int items0 = p.B * p.ntypes * p.H * nblk, items1 = p.B * nblk * nblk;
    attn_tc_kernel<0><<<items0 < kNumSMs ? items0 : kNumSMs, attn_threads(0), kASmem, st>>>(tmQ, tmP, p);
    if (int e = check_launch("attn_tc_kernel<stats>")) return e;
    if (stats_only) return 0;
    attn_tc_kernel<1><<<items1 < kNumSMs ? items1 : kNumSMs, attn_threads(1), kASmem, st>>>(tmQ, tmP, p);
    return check_launch("attn_tc_kernel<probs>");
}

}  // namespace xl
