// Fused attention for the original (q k^T) path: probabilities, head-reduced attention map AND the P V product in one
// tcgen05 kernel -- the per-head probabilities never leave the SM (reference: nn.MultiheadAttention with need_weights
// in the standard blocks and `attn_ori` / `x_ori = attn_ori @ v` in the surgery blocks,
// clip/clip_surgery_model.py:101-102,151-154,297-307).
//
// Given the softmax row statistics of the stats pass (attn_tc.cu, MODE 0) a CTA owns one 128-row query block of one
// image and walks  head group (4 heads) -> key block (128 keys) -> head -> 64-key half:
//   S  = X_h Y_h^T                     split-fp16 operands from shared memory (TMA), 3 MMA passes, fp32 in TMEM
//                                      (64-key sub-tiles: four 64-column S / P buffers in flight);
//   p  = exp2(alpha s - (m + log2 l))  exactly normalised, in the epilogue warps; summed over the heads in registers
//                                      (-> the attention map the API returns) and written BACK INTO THE S TILE's
//                                      tensor memory as split fp16 (hi | lo pairs, two keys per 32-bit column);
//   O_h += P V_h                       tcgen05.mma with the A operand read from TENSOR MEMORY and V_h [keys, head dim]
//                                      as an MN-major shared-memory operand straight from the qkv matrix (3 passes:
//                                      P_hi V_lo + P_lo V_hi + P_hi V_hi), fp32 accumulators of the 4 heads of the
//                                      group in the other half of TMEM (4 x 64 columns).
// TMEM budget: 4 x 64 columns (S / P) + 256 columns (O of 4 heads) = 512.  Warps: TMA producer (X / Y ring and V ring
// polled independently), MMA issuer (S runs up to four sub-steps ahead of P V), 16 epilogue warps in two groups that
// ping-pong over the key halves.  The head-reduced map leaves once per (head group, key block) through the warp's staging
// block and TMA: a plain store for the first head group, a reduce-add in L2 for the others (fixed order per address, so
// the result is bit-reproducible), straight into the API's attention tensor, whose rows are padded to Npad = round_up(N,4)
// floats (TMA needs 16 B-aligned row pitches; the Python side returns the [.., :N] view).
// Key blocks are trimmed to the valid keys rounded up to 16 (N = 1025: the ninth block is one 16-wide MMA).
#include <cuda_fp16.h>

#include <type_traits>

#include "attn_tc.cuh"
#include "common.cuh"
#include "excel_b200.h"
#include "tc.cuh"

// Timing experiments (tools/attn_probe.py builds private copies of this file with -DXL_TUNING -DXL_PV_VARIANT=<mask>; the
// product build defines neither, every XL_PV(bit) is the constant 0 and the guarded code is the plain path):
//   1 no P V MMAs   2 no S MMAs   4 no exp / split math   8 no tcgen05.ld   16 no tcgen05.st   32 no map store
//   64 X tiles loaded only for the first two steps   128 V tiles loaded only for the first two steps
//   256 no MUFU.EX2   512 no lo half (P_lo = P_hi)   1024 no fp16 conversions at all   2048 Veltkamp split on the FMA pipe
//   4096 clock trace of CTA 0 (MMA warp + one epilogue warp per group) into g_pv_trace, read back by excel_dev_pv_trace
// (variants whose P is garbage / NaN also change the chip's POWER draw and with it the clocks of every other kernel: under
//  the board power cap only variants that keep the data finite are valid timing comparisons)
#if defined(XL_TUNING) && defined(XL_PV_VARIANT)
#define XL_PV(bit) (((XL_PV_VARIANT) & (bit)) != 0)
#else
#define XL_PV(bit) false
#endif

#if defined(XL_TUNING)
// timing trace (dev builds only): CTA 0 records SM clock stamps of its MMA warp and of two epilogue warps (one per group)
__device__ unsigned long long g_pv_trace[3 * 4096];
#define XL_TRACE(tag) do { if (XL_PV(4096) && blockIdx.x == 0 && lane == 0 && tpos < tend) g_pv_trace[tpos++] = ((unsigned long long)(tag) << 56) | ((unsigned long long)clock64() & 0xffffffffffffffull); } while (0)
extern "C" int excel_dev_pv_trace(unsigned long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, g_pv_trace, sizeof(g_pv_trace));
}
#else
#define XL_TRACE(tag) do { } while (0)
#endif

namespace xl {

namespace {

constexpr int kHG = 4;                          // max heads per group (O accumulators: kHG x 64 TMEM columns); p.hpi <= kHG
constexpr uint32_t kTile = 128 * 64 * 2;        // 16 KB: 128 rows x 64 halves (one SWIZZLE_128B operand tile)
constexpr uint32_t kXYStage = 4 * kTile;        // X_hi, X_lo, Y_hi, Y_lo
constexpr uint32_t kVStage = 2 * kTile;         // V_hi, V_lo: 128 keys x 64 head-dim channels, straight from the qkv matrix
constexpr int kXYStages = 2, kVStages = 2;
constexpr uint32_t kStg = 16 * 2048;             // head-sum staging: one 32 x 16 fp32 block per epilogue warp
constexpr int kPvEpiWarps = 16;              // four per TMEM lane group, 32 columns of every S tile each
constexpr int kPvThreads = 64 + 32 * kPvEpiWarps;
constexpr size_t kPvSmem = kXYStages * kXYStage + kVStages * kVStage + kStg + 1024 /*align*/ + 256 /*barriers*/;

__device__ __forceinline__ float ex2a(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

}  // namespace

__global__ void __launch_bounds__(kPvThreads, 1)
attn_pv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmO, const AttnPvParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* xy = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* vs = xy + kXYStages * kXYStage;
    uint8_t* stg_base = vs + kVStages * kVStage;
    uint64_t* xy_full = reinterpret_cast<uint64_t*>(stg_base + kStg);
    uint64_t* xy_empty = xy_full + kXYStages;
    uint64_t* v_full = xy_empty + kXYStages;
    uint64_t* v_empty = v_full + kVStages;
    uint64_t* s_full = v_empty + kVStages;    // [4] S sub-tile complete (MMA -> epilogue)
    uint64_t* p_ready = s_full + 4;           // [4] P written back into the S sub-tile (epilogue -> MMA)
    uint64_t* o_full = p_ready + 4;           // O of the head group complete (MMA -> epilogue)
    uint64_t* o_empty = o_full + 1;           // O drained (epilogue -> MMA)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 1);

    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nblk = (p.N + 127) / 128;
    const int hpi = p.hpi;                                  // heads per group (1, 2 or 4)
    const int ngrp = (p.H + hpi - 1) / hpi;
    // Work item = (image, 128-row query block) walking all head groups, or -- p.gsplit, small batches -- one (image, query
    // block, head group) each, so that B * nblk * ngrp CTAs share the chip; a split item writes its partial head-sum into
    // slice g of the scratch map [ngrp][B,N,Npad] with plain stores (attn_combine_kernel adds the slices in a fixed order).
    const int gsp = p.gsplit ? ngrp : 1;
    const int items = p.B * nblk * gsp;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kXYStages; ++s) { mbar_init(&xy_full[s], 1); mbar_init(&xy_empty[s], 1); }
        for (int s = 0; s < kVStages; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
        for (int s = 0; s < 4; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_ready[s], kPvEpiWarps / 2); }
        mbar_init(o_full, 1);
        mbar_init(o_empty, kPvEpiWarps);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_o = tmem_base + 256;
    pdl_wait();
    constexpr uint32_t kIdescPV = make_idesc_bmn(64);   // B = V [keys, head dim]: MN-major

    // A "load step" is one (key block kb, head hh) pair: one X / Y stage and one V^T stage.  It is consumed as one or two
    // 64-key SUB-STEPS (the last key block of an image may hold <= 64 valid keys), each with its own 64-column S / P buffer:
    // four buffers in flight hide the MMA -> epilogue -> MMA round trip that a 128-key double buffer exposes.
    const int last_valid = p.N - (nblk - 1) * 128;          // valid keys of the last key block (1..128)
    const int nsub_last = 2;   // every load step has both 64-key sub-steps (the second may hold no valid key): sub-step parity ==
                               // key half, which is what lets the two epilogue groups below each own one half

    if (warp == 0) {
        // ---- TMA producer: the X / Y ring and the V^T ring advance independently (each polled without blocking), so a
        // V^T stage waiting for its P V MMAs never holds back the X / Y tiles of later steps.
        {
            const bool leader = elect_one_sync();
            if (leader) {
                tma_prefetch_desc(&tmQ);
            }
            struct It { int item, g, kb, hh; };
            auto advance = [&](It& it) {
                const int hc = min(hpi, p.H - it.g * hpi);
                if (++it.hh < hc) return;
                it.hh = 0;
                if (++it.kb < nblk) return;
                it.kb = 0;
                if (!p.gsplit && ++it.g < ngrp) return;
                it.item += gridDim.x;
                it.g = p.gsplit ? it.item % gsp : 0;
            };
            It ix = {(int)blockIdx.x, p.gsplit ? (int)blockIdx.x % gsp : 0, 0, 0}, iv = ix;
            uint32_t nx = 0, nv = 0;
            while (ix.item < items || iv.item < items) {
                if (ix.item < items) {
                    const int s = nx % kXYStages;
                    if (mbar_try(&xy_empty[s], ((nx / kXYStages) & 1) ^ 1)) {
                        const int rb = (ix.item / gsp) % nblk, b = ix.item / (gsp * nblk), h = ix.g * hpi + ix.hh;
                        uint8_t* st = xy + s * kXYStage;
                        const int xr = b * p.N + rb * 128, yr = b * p.N + ix.kb * 128;
                        const int xc = p.xo + h * 64, yc = p.yo + h * 64;
                        if (leader) {
                            const bool skip_x = XL_PV(64) && nx >= (uint32_t)kXYStages;
                            mbar_arrive_expect_tx(&xy_full[s], skip_x ? kXYStage / 2 : kXYStage);
                            if (!skip_x) {
                                tma_load_2d(st, &tmQ, &xy_full[s], xc, xr);
                                tma_load_2d(st + kTile, &tmQ, &xy_full[s], xc + p.lo_off, xr);
                            }
                            tma_load_2d(st + 2 * kTile, &tmQ, &xy_full[s], yc, yr);
                            tma_load_2d(st + 3 * kTile, &tmQ, &xy_full[s], yc + p.lo_off, yr);
                        }
                        ++nx;
                        advance(ix);
                    }
                }
                if (iv.item < items) {
                    const int sv = nv % kVStages;
                    if (mbar_try(&v_empty[sv], ((nv / kVStages) & 1) ^ 1)) {
                        const int b = iv.item / (gsp * nblk), h = iv.g * hpi + iv.hh;
                        uint8_t* vt = vs + sv * kVStage;
                        const int vr = b * p.N + iv.kb * 128, vc = p.vo + h * 64;   // V rows = keys, columns = head-dim channels
                        if (leader) {
                            if (XL_PV(128) && nv >= (uint32_t)kVStages) {
                                mbar_arrive(&v_full[sv]);
                            } else {
                                mbar_arrive_expect_tx(&v_full[sv], kVStage);
                                tma_load_2d(vt, &tmQ, &v_full[sv], vc, vr);
                                tma_load_2d(vt + kTile, &tmQ, &v_full[sv], vc + p.lo_off, vr);
                            }
                        }
                        ++nv;
                        advance(iv);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer (whole warp runs the control flow, one elected lane issues).  Within a head group the S tiles run
        // up to four sub-steps ahead of the P V products: S(0..3), PV(0), S(4), PV(1), S(5), ...  tcgen05.mma executes in
        // issue order, so S(u+4) may overwrite the buffer PV(u) reads its P from without a further barrier.
        const bool leader = elect_one_sync();
        int tpos = 0; const int tend = 4096; (void)tpos; (void)tend;
        struct Sub { int kb, hh, half; };             // position inside a head group
        uint32_t gs = 0, gp = 0;          // sub-steps issued (S / PV), across items: buffer = count & 3
        uint32_t ls = 0, lp = 0;          // load steps consumed (S / PV): ring stages and phases
        uint32_t ngd = 0;
        const uint32_t xy0 = smem_u32(xy), vs0 = smem_u32(vs);
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int g_lo = p.gsplit ? item % gsp : 0, g_hi = p.gsplit ? g_lo + 1 : ngrp;
            for (int g = g_lo; g < g_hi; ++g) {
                const int hc = min(hpi, p.H - g * hpi);
                const int U = ((nblk - 1) * 2 + nsub_last) * hc;
                auto next = [&](Sub& q) {     // sub-step order: key block, head, 64-key half
                    const int nsub = q.kb < nblk - 1 ? 2 : nsub_last;
                    if (++q.half < nsub) return;
                    q.half = 0;
                    if (++q.hh < hc) return;
                    q.hh = 0;
                    ++q.kb;
                };
                Sub qs = {0, 0, 0}, qp = {0, 0, 0};
                int us = 0;
                for (int up = 0; up < U; ++up) {
                    // ---- S of one LOAD STEP (both 64-key halves: sub-steps us, us + 1) as ONE N = 128 instruction per k-step into the
                    // adjacent buffer pair (gs, gs + 1): an M128 x N64 x K16 MMA re-reads the 4 KB X slice for 2 KB of Y and is bound by
                    // the shared-memory operand bandwidth (65-80 clk against 32 of math); at N = 128 the X slice is read once per 4 KB of Y.
                    for (; us < U && us + 1 < up + 4; us += 2, gs += 2) {
                        const uint32_t s = ls % kXYStages;
                        XL_TRACE(1);
                        mbar_wait(&xy_full[s], (ls / kXYStages) & 1);
                        tc_fence_after();
                        XL_TRACE(2);
                        const int nvalid = min(128, p.N - qs.kb * 128);              // >= 1; padding keys are not computed
                        const uint32_t idesc = make_idesc((nvalid + 15) & ~15);
                        const uint32_t tacc = tmem_base + (gs & 3) * 64;             // gs even: buffers (0, 1) or (2, 3)
                        const uint32_t st = xy0 + s * kXYStage, yb = st + 2 * kTile;
                        const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + kTile);
                        const uint64_t b_hi = umma_desc_sw128(yb), b_lo = umma_desc_sw128(yb + kTile);
                        if (leader) {
                            if (!XL_PV(2)) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                umma_f16(tacc, a_hi + 2 * k, b_lo + 2 * k, idesc, k != 0);
                                umma_f16(tacc, a_lo + 2 * k, b_hi + 2 * k, idesc, 1);
                                umma_f16(tacc, a_hi + 2 * k, b_hi + 2 * k, idesc, 1);
                            }
                            }
                            umma_commit(&xy_empty[s]);
                            umma_commit(&s_full[gs & 3]);
                            umma_commit(&s_full[(gs + 1) & 3]);
                        }
                        XL_TRACE(3);
                        ++ls;
                        next(qs);
                        next(qs);
                    }
                    // ---- PV(up): O_hh += P[64 keys] V_hh
                    const int nsub = qp.kb < nblk - 1 ? 2 : nsub_last;
                    const uint32_t buf = gp & 3, sv = lp % kVStages;
                    XL_TRACE(4);
                    mbar_wait(&p_ready[buf], (gp >> 2) & 1);
                    XL_TRACE(5);
                    if (qp.half == 0) mbar_wait(&v_full[sv], (lp / kVStages) & 1);
                    XL_TRACE(6);
                    if (up == 0) mbar_wait(o_empty, (ngd & 1) ^ 1);   // the previous group's O has been read out
                    tc_fence_after();
                    const int nvalid = min(64, p.N - qp.kb * 128 - qp.half * 64);
                    const uint32_t pbase = tmem_base + buf * 64;
                    const uint32_t vbox = vs0 + sv * kVStage + (uint32_t)qp.half * 8192u;   // keys 64..127: +64 rows x 128 B
                    const uint64_t b_hi = umma_desc_sw128(vbox), b_lo = umma_desc_sw128(vbox + kTile);
                    const uint32_t d = tmem_o + (uint32_t)(qp.hh * 64);
                    const uint32_t first = (qp.kb | qp.half) != 0;
                    if (leader && !XL_PV(1)) {
                        // keys 16k..16k+15 of the sub-tile live in its columns 16k..16k+15: hi pairs in +0..7, lo pairs in +8..15
                        if (nvalid > 48) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                umma_f16_ts(d, pbase + 16 * k, b_lo + 128 * k, kIdescPV, k ? 1u : first);
                                umma_f16_ts(d, pbase + 16 * k + 8, b_hi + 128 * k, kIdescPV, 1);
                                umma_f16_ts(d, pbase + 16 * k, b_hi + 128 * k, kIdescPV, 1);
                            }
                        } else {
                            for (int k = 0; k < ((nvalid + 15) >> 4); ++k) {
                                umma_f16_ts(d, pbase + 16 * k, b_lo + 128 * k, kIdescPV, k ? 1u : first);
                                umma_f16_ts(d, pbase + 16 * k + 8, b_hi + 128 * k, kIdescPV, 1);
                                umma_f16_ts(d, pbase + 16 * k, b_hi + 128 * k, kIdescPV, 1);
                            }
                        }
                    }
                    if (leader && qp.half == nsub - 1) umma_commit(&v_empty[sv]);
                    XL_TRACE(7);
                    if (qp.half == nsub - 1) ++lp;
                    ++gp;
                    next(qp);
                }
                if (leader) umma_commit(o_full);
                ++ngd;
            }
        }
    } else {
        // ---- epilogue warps in TWO GROUPS that ping-pong: group grp = qt / 2 owns the sub-tiles of key half grp (buffers grp and
        // grp + 2), so while one group is in the latency-bound tail of a sub-tile (TMEM store, fence, barrier hand-off) the other
        // is in the middle of the next one.  Warp (lg, grp, cq): TMEM lanes 32*lg..+31 (query rows), columns 32*cq..+31 of its
        // group's 64-key sub-tiles, as two 16-column chunks: S -> registers -> p -> split fp16 written back over the same 16
        // columns (hi pairs in +0..7, lo pairs in +8..15), so a warp only ever overwrites scores it has already read.
        const int ew = warp - 2, lg = warp & 3, qt = ew >> 2, grp = qt >> 1, cq = qt & 1;
        const int trow = lg * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(lg * 32) << 16;
        float* stg = reinterpret_cast<float*>(stg_base) + ew * 512;   // warp-private 32 rows x 16 floats
        uint32_t gl = 0, ngd = 0;                                      // load steps seen: this group's sub-tile is 2 * gl + grp
        int tpos = ew == 0 ? 4096 : 8192; const int tend = (ew == 0 || ew == 8) ? tpos + 4096 : 0; (void)tpos; (void)tend;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int rb = (item / gsp) % nblk, b = item / (gsp * nblk);
            const int row = rb * 128 + trow;
            const bool row_ok = row < p.N;
            const int nrow = min(32, p.N - rb * 128 - lg * 32);       // valid rows of this warp's 32 (may be <= 0)
            const int g_lo = p.gsplit ? item % gsp : 0, g_hi = p.gsplit ? g_lo + 1 : ngrp;
            for (int g = g_lo; g < g_hi; ++g) {
                const int hc = min(hpi, p.H - g * hpi);
                const int zo = p.gsplit ? g * p.B + b : b;            // map slice of this item
                const float* mrow = p.ml + ((int64_t)b * p.H + g * hpi) * p.N + row;   // + hh * N
                // The head-reduced map block of a key block (this warp: 32 rows x 32 keys) leaves as two 16-column chunks through the
                // warp's single staging block.  Chunk 0 goes out right after the key block's last head; chunk 1 is DEFERRED until
                // after the first head step of the next key block: by then the bulk engine (whose queue is full of operand loads) has
                // long read chunk 0 out of the staging block, where a back-to-back second chunk waited ~3000 clocks for it.
                float pend[16];
                int pend_kc = -1;                                     // key column of the deferred chunk (-1: none)
                const float cf = p.coef * (1.f / 1024.f);
                auto flush_chunk = [&](const float (&v)[16], int kc) {
                    if (lane == 0) tma_store_wait_read<0>();          // the previous chunk has left the staging buffer
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<float4*>(stg + lane * 16 + ((j ^ ((lane >> 1) & 3)) << 2)) =
                            make_float4(cf * v[4 * j], cf * v[4 * j + 1], cf * v[4 * j + 2], cf * v[4 * j + 3]);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        // a plain store for the first head group, a reduce-add (performed in L2) for the others
                        if (g == 0 || p.gsplit) tma_store_3d(&tmO, stg, kc, rb * 128 + lg * 32, zo);
                        else tma_reduce_add_3d(&tmO, stg, kc, rb * 128 + lg * 32, zo);
                        tma_store_commit();
                    }
                };
                for (int kb = 0; kb < nblk; ++kb) {
                    const int key0 = kb * 128 + grp * 64 + cq * 32;   // this warp's 32 keys of the key block
                    float acc[2][16];
#pragma unroll
                    for (int c = 0; c < 2; ++c)
#pragma unroll
                        for (int e = 0; e < 16; ++e) acc[c][e] = 0.f;
                    float m_next = row_ok ? __ldg(mrow) : INFINITY;   // rows past N: exp2(-inf) = 0
                    for (int hh = 0; hh < hc; ++hh, ++gl) {
                        const float m_row = m_next;
                        if (row_ok && hh + 1 < hc) m_next = __ldg(mrow + (int64_t)(hh + 1) * p.N);
                        const uint32_t u = 2 * gl + grp, buf = u & 3;
                        XL_TRACE(8);
                        mbar_wait(&s_full[buf], (u >> 2) & 1);
                        tc_fence_after();
                        XL_TRACE(9);
                        if (nrow <= 0) {   // (warp-uniform) none of this warp's rows exists: their P (and O) rows are never read back
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&p_ready[buf]);
                            continue;
                        }
                        // Only the LAST key block of an image can hold padding keys.  The two cases are separate instantiations of the
                        // chunk code: written as one body, the per-element `key >= N ? -inf : s` test was if-converted and ran
                        // (ISETP + SEL per element, a quarter of the epilogue's instructions) for every key block.
                        auto chunks = [&](auto tail_tag) {
                        constexpr bool TAIL = decltype(tail_tag)::value;
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            if (TAIL && key0 + c * 16 >= p.N) break;   // (uniform) padding keys only: columns unused
                            const uint32_t taddr = tmem_base + lane_addr + (uint32_t)(buf * 64 + cq * 32 + c * 16);
                            uint32_t r[16];
                            if (!XL_PV(8)) tmem_ld16(taddr, r);
                            else {
#pragma unroll
                                for (int e = 0; e < 16; ++e) r[e] = 0x3c000000u + (uint32_t)(lane + e + c);
                            }
                            if constexpr (TAIL) {
#pragma unroll
                                for (int e = 0; e < 16; ++e)
                                    if (key0 + c * 16 + e >= p.N) r[e] = 0xff800000u;  // -inf -> probability 0 for the padding keys
                            }
                            uint32_t ph[8], pl[8];
                            if (XL_PV(4)) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) { ph[e] = r[2 * e]; pl[e] = r[2 * e + 1]; acc[c][2 * e] += __uint_as_float(r[e]); }
                            } else
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                // 2^10 p = exp2(alpha s - (m + log2 l - 10)): one FFMA + one MUFU per element
                                // (packed fp32x2 FMA / ADD: scalar FP32 issues at half rate on sm_100)
                                const float2 x = __ffma2_rn(make_float2(p.alpha, p.alpha),
                                                            make_float2(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1])),
                                                            make_float2(-m_row, -m_row));
                                const float2 v = XL_PV(256) ? x : make_float2(ex2a(x.x), ex2a(x.y));
                                const float2 a2 = __fadd2_rn(make_float2(acc[c][2 * e], acc[c][2 * e + 1]), v);
                                acc[c][2 * e] = a2.x;
                                acc[c][2 * e + 1] = a2.y;
                                if (XL_PV(1024)) {
                                    ph[e] = __float_as_uint(v.x); pl[e] = __float_as_uint(v.y);
                                } else if (XL_PV(2048)) {
                                    // Veltkamp split: hi = v rounded to 11 significant bits, computed on the FMA pipe (no fp16 -> fp32 unpack)
                                    const float2 cc = __fmul2_rn(v, make_float2(8193.f, 8193.f));
                                    const float2 dd = __ffma2_rn(v, make_float2(-1.f, -1.f), cc);
                                    const float2 hf2 = __ffma2_rn(dd, make_float2(-1.f, -1.f), cc);
                                    const float2 lo2 = __ffma2_rn(hf2, make_float2(-1.f, -1.f), v);
                                    const __half2 hh2 = __floats2half2_rn(hf2.x, hf2.y);
                                    const __half2 ll2 = __floats2half2_rn(lo2.x, lo2.y);
                                    ph[e] = *reinterpret_cast<const uint32_t*>(&hh2);
                                    pl[e] = *reinterpret_cast<const uint32_t*>(&ll2);
                                } else {
                                const __half2 hh2 = __floats2half2_rn(v.x, v.y);
                                ph[e] = *reinterpret_cast<const uint32_t*>(&hh2);
                                if (XL_PV(512)) { pl[e] = ph[e]; } else {
                                const float2 hf2 = __half22float2(hh2);
                                const float2 lo2 = __fadd2_rn(v, make_float2(-hf2.x, -hf2.y));
                                const __half2 ll2 = __floats2half2_rn(lo2.x, lo2.y);
                                pl[e] = *reinterpret_cast<const uint32_t*>(&ll2);
                                }
                                }
                            }
                            if (!XL_PV(16)) {
                            tmem_st8(taddr, ph);       // P_hi: keys (2e, 2e+1) of the chunk in column e
                            tmem_st8(taddr + 8, pl);   // P_lo
                            } else if (ph[0] == 0x12345678u && pl[3] == 0x9abcdef0u) acc[c][0] += 1.f;   // keep the values alive
                        }
                        };
                        if (kb == nblk - 1) chunks(std::true_type{});
                        else chunks(std::false_type{});
                        XL_TRACE(10);
                        tmem_st_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&p_ready[buf]);
                        XL_TRACE(11);
                        if (hh == 0 && pend_kc >= 0) {               // the previous key block's second chunk
                            flush_chunk(pend, pend_kc);
                            pend_kc = -1;
                        }
                    }
                    // head-reduced map of this key block: coef * 2^-10 * sum over the group's heads -- no thread waits on global memory
                    XL_TRACE(12);
                    if (nrow > 0 && !XL_PV(32) && key0 < p.N) {
                        flush_chunk(acc[0], key0);
                        XL_TRACE(13);
                        if (key0 + 16 < p.N) {
#pragma unroll
                            for (int e = 0; e < 16; ++e) pend[e] = acc[1][e];
                            pend_kc = key0 + 16;
                        }
                    }
                    XL_TRACE(14);
                }
                if (pend_kc >= 0) {                                   // last key block of the group
                    flush_chunk(pend, pend_kc);
                    pend_kc = -1;
                }
                if (lane == 0) tma_store_wait_all<0>();   // this group's map blocks are performed before the next group adds to them
                // ---- O of head hh = qt of the group: TMEM -> split fp16 -> o[b*N + row, h*64 ..] (hi) / [.. + D] (lo)
                mbar_wait(o_full, ngd & 1);
                tc_fence_after();
                if (qt < hc) {
#pragma unroll 1
                    for (int q = 0; q < 4; ++q) {
                        uint32_t r[16];
                        tmem_ld16(tmem_o + lane_addr + (uint32_t)(qt * 64 + q * 16), r);
                        if (row_ok) {
                            __half* oh = p.o + ((int64_t)b * p.N + row) * (2 * p.D) + (g * hpi + qt) * 64 + q * 16;
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                __align__(16) __half2 h2[4], l2[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float v0 = __uint_as_float(r[8 * j + 2 * e]) * (1.f / 1024.f);
                                    const float v1 = __uint_as_float(r[8 * j + 2 * e + 1]) * (1.f / 1024.f);
                                    h2[e] = __floats2half2_rn(v0, v1);
                                    const float2 hf2 = __half22float2(h2[e]);
                                    l2[e] = __floats2half2_rn(v0 - hf2.x, v1 - hf2.y);
                                }
                                *reinterpret_cast<uint4*>(oh + 8 * j) = *reinterpret_cast<const uint4*>(h2);
                                *reinterpret_cast<uint4*>(oh + p.D + 8 * j) = *reinterpret_cast<const uint4*>(l2);
                            }
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

// out[i] = sum_g part[g][i] in the fixed order g = 0, 1, ... (split launches: one partial head-sum map per head group)
__global__ void __launch_bounds__(256)
attn_combine_kernel(const float4* __restrict__ part, int ngrp, int64_t slice4, float4* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < slice4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 a = __ldcs(part + i);
        for (int g = 1; g < ngrp; ++g) {
            const float4 v = __ldcs(part + g * slice4 + i);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        out[i] = a;
    }
}

int attn_pv(const CUtensorMap& tmQ, const AttnPvParams& p, cudaStream_t st) {
    static unsigned long long attr_once = 0;
    if (first_use_on_device(attr_once)) {
        XL_CUDA(cudaFuncSetAttribute(attn_pv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPvSmem));
    }
    XL_REQUIRE(p.B > 0 && p.H > 0 && p.N > 0 && p.D == p.H * 64, "attn_pv: bad shape");
    XL_REQUIRE(p.ml && p.out && p.o, "attn_pv: missing buffers");
    XL_REQUIRE(p.hpi == 1 || p.hpi == 2 || p.hpi == 4, "attn_pv: heads per group must be 1, 2 or 4");
    const int ngrp = (p.H + p.hpi - 1) / p.hpi, nblk = (p.N + 127) / 128, Npad = (p.N + 3) & ~3;
    XL_REQUIRE(!p.gsplit || p.part, "attn_pv: split launches need the partial-map scratch");
    CUtensorMap tmO;
    if (int e = make_map_store(&tmO, p.gsplit ? p.part : p.out, p.gsplit ? p.B * ngrp : p.B, p.N)) return e;
    const int items = p.B * nblk * (p.gsplit ? ngrp : 1);
    XL_CUDA(launch_pdl(attn_pv_kernel, dim3(items < kNumSMs ? items : kNumSMs), dim3(kPvThreads), kPvSmem, st, tmQ, tmO, p));
    if (int e = check_launch("attn_pv_kernel")) return e;
    if (!p.gsplit) return 0;
    const int64_t slice4 = (int64_t)p.B * p.N * Npad / 4;
    const int64_t blocks = ceil_div64(slice4, 256 * 2);
    XL_CUDA(launch_pdl(attn_combine_kernel, dim3((unsigned)(blocks < 148 * 8 ? blocks : 148 * 8)), dim3(256), 0, st,
                       reinterpret_cast<const float4*>(p.part), ngrp, slice4, reinterpret_cast<float4*>(p.out)));
    return check_launch("attn_combine_kernel");
}

// heads per group / split decision for a batch: keep the whole-item form (3 head groups of 4 walked by one CTA, the map
// reduced in L2 in a fixed order) while (image, query block) items fill the chip; below that one item per head group, with
// as few heads per group as it takes to reach one CTA per SM.
void attn_pv_plan(int B, int H, int N, int* hpi, int* gsplit) {
    const int nblk = (N + 127) / 128;
    *hpi = 4; *gsplit = 0;
    if (B * nblk >= (kNumSMs * 3) / 4) return;
    *gsplit = 1;
    for (int h = 4; h >= 1; h /= 2) {
        *hpi = h;
        if (B * nblk * ((H + h - 1) / h) >= kNumSMs) break;
    }
}

}  // namespace xl
