// tcgen05 GEMM engine for the encoder's dense contractions (linear layers, q k^T, P V) on sm_100a.
//
// Precision: the reference is fp32 end to end (clip/build_model.py:72) and the parity bar is 1e-3 on
// min-max-normalised CAMs through 12 layers, which single-pass TF32/BF16 misses (measured on the oracle:
// 4e-3 / 2e-2).  Operands are therefore SPLIT fp16 pairs x = hi + lo (hi = fp16(x), lo = fp16(x - hi),
// 22 significant bits) and every tile computes  A_hi B_hi + A_hi B_lo + A_lo B_hi  with fp32
// accumulation in TMEM: fp32-quality products (rel. error ~7e-7) at three tensor-core passes, and only
// four tile loads per k-block (each operand half is reused by two of the three MMAs).
//
// Split matrices are row-major fp16 [rows, 2*Kp]: columns [0,K) hold hi, [Kp, Kp+K) hold lo, Kp a multiple
// of 64, padding zero.  Both operands are K-major ("TN": C[m,n] = sum_k A[m,k] B[n,k]), i.e. activations
// [tokens, features] against nn.Linear weights [out, in], q against k, P against V^T.
//
// Kernel anatomy (persistent: one CTA per SM walks 128 x BN output tiles, BN = 64 / 128 / 256; 320 threads):
//   warp 0   : TMA producer -- A_hi, A_lo (64 k x 128 rows) and B_hi, B_lo (64 k x BN rows, or 64-column x 64-row boxes of
//              an MN-major B) per stage, 128B swizzle, 3-stage mbarrier ring (2 stages of 96 KB at BN = 256);
//   warp 1   : allocates 2 x BN TMEM columns (two accumulators); its elected lane issues 12 tcgen05.mma
//              (M128 x BN x K16, kind::f16) per stage and tcgen05.commit's the stage back to the producer / the finished
//              accumulator to the epilogue, which then overlaps the next tile's main loop;
//   warps 2-9: epilogue, two teams of four warps on alternate 32-column chunks -- tcgen05.ld 32 lanes x 32 columns,
//              alpha / bias / QuickGELU or ReLU / residual in registers, swizzled staging, TMA store (fp32 or split fp16).
#include <cuda_fp16.h>

#include "common.cuh"
#include "excel_b200.h"
#include "gemm_tc.cuh"
#include "ptx.cuh"
#include "tc.cuh"

namespace xl {

constexpr int kBK = 64;
constexpr uint32_t kTileBytes = kBM * kBK * 2;        // one 128-row x 64-k fp16 tile: 16 KB
constexpr int kTcThreads = 64 + 256;                 // TMA producer warp, MMA warp, 8 epilogue warps
constexpr uint32_t kEpiBytes = 2 * 16384;             // epilogue staging: two 128-row x 128 B blocks (TMA store sources)
// Per tile width BN: a stage holds A_hi, A_lo (16 KB each) and B_hi, B_lo (BN x 128 B each).  An SM ingests ~64 B/clk from
// L2, about what 128 x 128 tiles need to keep the tensor pipe busy (64 KB per 36 MMAs); 256-wide tiles move 96 KB per 72
// MMAs and leave the pipe as the limit.  They take 2 x 256 TMEM columns and a 2-stage ring (2 x 96 KB).
__host__ __device__ constexpr uint32_t tc_b_bytes(int bn) { return (uint32_t)bn * kBK * 2; }
__host__ __device__ constexpr uint32_t tc_stage_bytes(int bn) { return 2 * kTileBytes + 2 * tc_b_bytes(bn); }
__host__ __device__ constexpr int tc_stages(int bn) { return bn > 128 ? 2 : 3; }
__host__ __device__ constexpr size_t tc_smem(int bn) { return tc_stages(bn) * tc_stage_bytes(bn) + kEpiBytes + 1024 /*align*/ + 256 /*barriers*/; }

// Persistent kernel: one CTA per SM walks the output tiles t = blockIdx.x + i*gridDim.x (n fastest, so
// neighbouring CTAs share A rows in L2).  Two TMEM accumulators (2 x kBN columns) let the epilogue of tile i
// overlap the main loop of tile i+1; the operand ring keeps running across tile boundaries.
// TMA_EPI: the output leaves through TMA stores (fp32 tile: 4D map tmC; split-fp16: 3D map tmS) -- the epilogue
// threads only move TMEM -> registers -> swizzled shared memory; otherwise (oddly pitched outputs) they store
// to global themselves.
// RES: fp32 output with a residual term (its prefetch registers and control flow stay out of the other instantiations --
// folding it into a run-time branch cost the GELU / split-output GEMM 20 % through code scheduling alone).
template <int kBN, bool TMA_EPI, bool RES>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmS, const TcParams p,
               int tiles_n, int tiles_m, int num_tiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // SWIZZLE_128B wants 1024 B alignment
    constexpr int kTcStages = tc_stages(kBN);
    constexpr uint32_t kStageBytes = tc_stage_bytes(kBN), kBBytes = tc_b_bytes(kBN);
    float* stage = reinterpret_cast<float*>(tiles + kTcStages * kStageBytes);      // epilogue staging, 128 x 36 floats
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + kTcStages * kStageBytes + kEpiBytes);
    uint64_t* empty_bar = full_bar + kTcStages;
    uint64_t* acc_full = empty_bar + kTcStages;   // [2]
    uint64_t* acc_empty = acc_full + 2;           // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t kIdesc = make_idesc(kBN);
    constexpr uint32_t kTxBytes = 2 * kTileBytes + 2 * kBN * kBK * 2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kTcStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], TMA_EPI ? 8 : 4);  // one arrival per (active) epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * kBN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();   // the producers of A / B / residual have completed (set-up above overlapped their tail)

    auto decode = [&](int t, int& m0, int& n0, int& z1, int& z2) {
        const int nt = t % tiles_n, r = t / tiles_n;
        const int mt = r % tiles_m, z = r / tiles_m;
        m0 = mt * kBM; n0 = nt * kBN; z1 = z / p.nb2; z2 = z % p.nb2;
    };

    if (warp == 0) {
        // whole warp runs the control flow, one elected lane issues (keeps the TMA / tcgen05 operands on the uniform datapath)
        const bool leader = elect_one_sync();
        if (leader) {
            tma_prefetch_desc(&tmA);
            tma_prefetch_desc(&tmB);
        }
        int it = 0;  // running k-block counter across tiles
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            int m0, n0, z1, z2;
            decode(t, m0, n0, z1, z2);
            const int a_row = p.a_row0 + z1 * p.a_row1 + z2 * p.a_row2 + m0, a_col = p.a_col0 + z1 * p.a_col1 + z2 * p.a_col2;
            const int b_row = p.b_row0 + z1 * p.b_row1 + z2 * p.b_row2 + n0, b_col = p.b_col0 + z1 * p.b_col1 + z2 * p.b_col2;
            for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
                const int s = it % kTcStages;
                mbar_wait(&empty_bar[s], ((it / kTcStages) & 1) ^ 1);
                uint8_t* st = tiles + s * kStageBytes;
                if (leader) {
                    mbar_arrive_expect_tx(&full_bar[s], kTxBytes);
                    tma_load_2d(st, &tmA, &full_bar[s], a_col + kb * kBK, a_row);
                    tma_load_2d(st + kTileBytes, &tmA, &full_bar[s], a_col + p.a_lo_off + kb * kBK, a_row);
                    if (p.b_mn) {
                        // MN-major B: 64-column x 64-row boxes, one per 64 columns of the tile (b_row / b_col then hold the
                        // K row / N column of the tile: n0 moves along the columns)
#pragma unroll
                        for (int c = 0; c < kBN; c += 64) {
                            tma_load_2d(st + 2 * kTileBytes + c * 128, &tmB, &full_bar[s], b_col + n0 + c, b_row - n0 + kb * kBK);
                            tma_load_2d(st + 2 * kTileBytes + kBBytes + c * 128, &tmB, &full_bar[s], b_col + n0 + c + p.b_lo_off,
                                        b_row - n0 + kb * kBK);
                        }
                    } else {
                        // B tiles wider than 128 rows arrive as stacked 128-row boxes (the operand maps keep a 128-row box)
#pragma unroll
                        for (int r = 0; r < kBN; r += 128) {
                            tma_load_2d(st + 2 * kTileBytes + r * 128, &tmB, &full_bar[s], b_col + kb * kBK, b_row + r);
                            tma_load_2d(st + 2 * kTileBytes + kBBytes + r * 128, &tmB, &full_bar[s], b_col + p.b_lo_off + kb * kBK, b_row + r);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        const bool leader = elect_one_sync();
        const uint32_t tiles0 = smem_u32(tiles);
        int it = 0, i = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++i) {
            const int buf = i & 1;
            mbar_wait(&acc_empty[buf], ((i >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t tacc = tmem_base + (uint32_t)(buf * kBN);
            for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
                const int s = it % kTcStages;
                mbar_wait(&full_bar[s], (it / kTcStages) & 1);
                tc_fence_after();
                const uint32_t st = tiles0 + s * kStageBytes;
                const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + kTileBytes);
                // K-major B: a 16-element K step is 32 B inside the swizzle atom; MN-major B: 16 rows = 2048 B, atoms along N 8 KB apart
                const uint64_t b_hi = p.b_mn ? umma_desc_sw128_mn(st + 2 * kTileBytes, 8192) : umma_desc_sw128(st + 2 * kTileBytes);
                const uint64_t b_lo = p.b_mn ? umma_desc_sw128_mn(st + 2 * kTileBytes + kBBytes, 8192)
                                             : umma_desc_sw128(st + 2 * kTileBytes + kBBytes);
                const uint64_t b_step = p.b_mn ? 128 : 2;
                const uint32_t idesc = p.b_mn ? make_idesc_bmn(kBN) : kIdesc;
                if (leader) {
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
                        const uint64_t adv = (uint64_t)(k * 32 >> 4);  // A: 16 fp16 = 32 B further along K inside the swizzle atom
                        umma_f16(tacc, a_hi + adv, b_lo + k * b_step, idesc, (kb | k) != 0);
                        umma_f16(tacc, a_lo + adv, b_hi + k * b_step, idesc, 1);
                        umma_f16(tacc, a_hi + adv, b_hi + k * b_step, idesc, 1);
                    }
                    umma_commit(&empty_bar[s]);  // stage s may be refilled once these MMAs have read it
                }
            }
            if (leader) umma_commit(&acc_full[buf]);     // accumulator complete
        }
    } else {
        // ---- epilogue (warps 2..9).  Warp w owns TMEM lanes 32*(w%4)..+31 == tile rows 32*(w%4)+lane; the two teams
        // of four warps (2..5, 6..9) take alternate 32-column chunks of the tile.
        const int lg = warp & 3, team = (warp - 2) >> 2;
        const int trow = lg * 32 + lane;
        const float alpha = p.alpha;
        const int act = p.act;
        int i = 0;
        if constexpr (TMA_EPI) {
            // Per 32-column chunk each thread pulls its row slice out of TMEM, applies alpha / bias / QuickGELU /
            // residual and writes it into its team's swizzled staging buffer; the team's elected thread then hands
            // the 128 x 32 block to the TMA store engine, which also clips rows >= M and columns >= N.
            uint8_t* sb = reinterpret_cast<uint8_t*>(stage) + team * 16384;
            const bool leader = lane == 0 && ((warp - 2) & 3) == 0;
            const int team_bar = 1 + team;
            // The residual slice of a chunk (this thread's row, 32 columns) is fetched one chunk AHEAD into rq[]: the global
            // round trip overlaps the previous chunk's staging / store and the TMEM wait instead of sitting between
            // tcgen05.ld and the staging store (measured: out_proj 102 -> 90 us).
            float rq[32];
            auto fetch_residual = [&](int tt, int c) {
                int m0, n0, z1, z2;
                decode(tt, m0, n0, z1, z2);
                const int nb = n0 + c * 32;
                const bool ok = m0 + trow < p.M;
                const float* r = p.residual + (int64_t)z1 * p.c1 + (int64_t)z2 * p.c2 + (int64_t)(m0 + trow) * p.ldc + nb;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (ok && nb + j + 3 < p.N) {
                        const float4 q = *reinterpret_cast<const float4*>(r + j);
                        rq[j] = q.x; rq[j + 1] = q.y; rq[j + 2] = q.z; rq[j + 3] = q.w;
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) rq[j + e] = (ok && nb + j + e < p.N) ? r[j + e] : 0.f;
                    }
                }
            };
            auto fetch_next_residual = [&](int tt, int c) {   // next chunk of this team: same tile, or the first chunk of this CTA's next tile
                if (c + 2 < kBN / 32) fetch_residual(tt, c + 2);
                else if (tt + (int)gridDim.x < num_tiles) fetch_residual(tt + gridDim.x, team);
            };
            constexpr bool has_res = RES;
            if constexpr (has_res) {
                if ((int)blockIdx.x < num_tiles) fetch_residual(blockIdx.x, team);
            }
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++i) {
                int m0, n0, z1, z2;
                decode(t, m0, n0, z1, z2);
                const int buf = i & 1;
                const float* bias = p.bias ? p.bias + (int64_t)z1 * p.bias1 : nullptr;
                mbar_wait(&acc_full[buf], (i >> 1) & 1);
                tc_fence_after();
#pragma unroll 1
                for (int c = team; c < kBN / 32; c += 2) {
                    uint32_t r[32];
                    tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * kBN + c * 32), r);
                    if (c + 2 >= kBN / 32) {  // all of this warp's TMEM reads for the tile are done
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_empty[buf]);
                    }
                    const int nb = n0 + c * 32;
                    if (nb >= p.N) {           // (team-uniform) nothing to store for this chunk
                        if constexpr (has_res) fetch_next_residual(t, c);
                        continue;
                    }
                    if (leader) tma_store_wait_read<0>();  // the store that last read the team's buffer has drained
                    bar_sync(team_bar, 128);
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = alpha * __uint_as_float(r[j]);
                    if (bias) {
                        if (nb + 32 <= p.N && (reinterpret_cast<uintptr_t>(bias + nb) & 15) == 0) {   // (uniform) 8 x 16 B broadcast loads
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 q = __ldg(reinterpret_cast<const float4*>(bias + nb + j));
                                v[j] += q.x; v[j + 1] += q.y; v[j + 2] += q.z; v[j + 3] += q.w;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (nb + j < p.N) v[j] += __ldg(bias + nb + j);
                        }
                    }
                    // the activation as its own (warp-uniform) branch around a whole pass over the chunk: inside the element loop
                    // a three-way choice was compiled to predicated code that ran the GELU's two MUFUs for every GEMM
                    if (act == 1) {          // QuickGELU x * sigmoid(1.702 x) (clip_surgery_model.py:280-282), exp2 domain
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __fdividef(v[j], 1.f + exp2f(-2.4554669595930156f * v[j]));
                    } else if (act == 2) {   // ReLU (model/segformer_head.py:24)
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                    }
                    if (p.C) {
                        if constexpr (has_res) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] += rq[j];
                            fetch_next_residual(t, c);
                        }
                        // fp32 tile [128][32]: 128 B rows, SWIZZLE_128B (16 B chunk index ^= row % 8): conflict-free
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            *reinterpret_cast<float4*>(sb + trow * 128 + ((j ^ (trow & 7)) << 4)) =
                                make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    } else {
                        // split fp16 tiles hi | lo, [128][32] halves each: 64 B rows, SWIZZLE_64B (chunk ^= (row/2) % 4)
                        uint8_t* sh = sb;
                        uint8_t* sl = sb + 8192;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            __align__(16) __half2 h[4], l[4];   // packed conversions: one F2FP per pair
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float v0 = v[8 * j + 2 * e], v1 = v[8 * j + 2 * e + 1];
                                h[e] = __floats2half2_rn(v0, v1);
                                const float2 hf = __half22float2(h[e]);
                                l[e] = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
                            }
                            const int off = trow * 64 + ((j ^ ((trow >> 1) & 3)) << 4);
                            *reinterpret_cast<uint4*>(sh + off) = *reinterpret_cast<const uint4*>(h);
                            *reinterpret_cast<uint4*>(sl + off) = *reinterpret_cast<const uint4*>(l);
                        }
                    }
                    fence_proxy_async_smem();
                    bar_sync(team_bar, 128);
                    if (leader) {
                        if (p.C) {
                            if (p.c_add) tma_reduce_add_4d(&tmC, sb, nb, m0, z2, z1);
                            else tma_store_4d(&tmC, sb, nb, m0, z2, z1);
                        } else {
                            const int col = z2 * (int)p.cs2 + nb;
                            tma_store_3d(&tmS, sb, col, m0, z1);
                            tma_store_3d(&tmS, sb + 8192, col + p.cs_lo_off, m0, z1);
                        }
                        tma_store_commit();
                    }
                }
            }
            if (leader) tma_store_wait_read<0>();
        } else if (team == 0) {
        // Fallback for outputs TMA cannot address (row pitch not a multiple of 16 B), warps 2..5 only: stage 128 x 32
        // blocks in shared memory (pitch 36 floats) and stream them out with 8 lanes per row (128 B segments).
        constexpr int kPitch = 36;
        const int ew = warp - 2, sub = lane >> 3, c4 = (lane & 7) * 4;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++i) {
            int m0, n0, z1, z2;
            decode(t, m0, n0, z1, z2);
            const int buf = i & 1;
            mbar_wait(&acc_full[buf], (i >> 1) & 1);
            tc_fence_after();
            const int64_t zoffc = (int64_t)z1 * p.c1 + (int64_t)z2 * p.c2;
            const int64_t zoffs = (int64_t)z1 * p.cs1 + (int64_t)z2 * p.cs2;
            const float* bias = p.bias ? p.bias + (int64_t)z1 * p.bias1 : nullptr;
#pragma unroll 1
            for (int c = 0; c < kBN / 32; ++c) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * kBN + c * 32), r);
                if (c == kBN / 32 - 1) {  // all of this thread's TMEM reads for the tile are done
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[buf]);
                }
                const int nb = n0 + c * 32;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float x = alpha * __uint_as_float(r[j + e]);
                        if (bias && nb + j + e < p.N) x += __ldg(bias + nb + j + e);
                        if (act == 1) x = __fdividef(x, 1.f + exp2f(-2.4554669595930156f * x));  // QuickGELU
                        else if (act == 2) x = fmaxf(x, 0.f);
                        v[e] = x;
                    }
                    *reinterpret_cast<float4*>(stage + trow * kPitch + j) = make_float4(v[0], v[1], v[2], v[3]);
                }
                bar_sync(1, 128);
                const int ncols = min(32, p.N - nb);
                if (ncols > 0) {
#pragma unroll 1
                    for (int rr = ew * 4 + sub; rr < kBM; rr += 16) {
                        const int m = m0 + rr;
                        if (m >= p.M) break;
                        const float* srow = stage + rr * kPitch + c4;
                        if (p.C) {
                            float* dst = p.C + zoffc + (int64_t)m * p.ldc + nb + c4;
                            const float* res = p.residual ? p.residual + zoffc + (int64_t)m * p.ldc + nb + c4 : nullptr;
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (c4 + e < ncols) dst[e] = srow[e] + (res ? res[e] : 0.f);
                        }
                        if (p.Cs) {  // split-fp16 copy (operand of the next GEMM); no residual on this path
                            __half* hi = p.Cs + zoffs + (int64_t)m * p.lds + nb + c4;
                            __half* lo = hi + p.cs_lo_off;
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (c4 + e < ncols) {
                                    const __half h = __float2half_rn(srow[e]);
                                    hi[e] = h;
                                    lo[e] = __float2half_rn(srow[e] - __half2float(h));
                                }
                        }
                    }
                }
                bar_sync(1, 128);  // staging buffer is reused by the next chunk
            }
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * kBN);
    }
}

// ---- fp32 -> split fp16 ---------------------------------------------------------------------------------
// out [rows, 2*Kp] : hi | lo, zero padded to Kp columns each
__global__ void split_f16_kernel(const float* __restrict__ x, int64_t ldx, int rows, int cols, int Kp, __half* __restrict__ out,
                                 float scale) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= Kp) return;
    float v = c < cols ? x[(int64_t)r * ldx + c] * scale : 0.f;
    const __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));   // hi saturates: |v| <= 2 x 65504 stays finite
    out[(int64_t)r * 2 * Kp + c] = h;
    out[(int64_t)r * 2 * Kp + Kp + c] = __float2half_rn(v - __half2float(h));
}

int make_operand_map(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols_total, int64_t ld_elems, int box_rows) {
    const uint64_t dims[2] = {(uint64_t)cols_total, (uint64_t)rows};
    const uint64_t strides[1] = {(uint64_t)ld_elems * 2};
    const uint32_t box[2] = {kBK, (uint32_t)box_rows};
    return encode_tensor_map(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

int tc_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, int batch, int bn, cudaStream_t st) {
    static unsigned long long attr_once = 0;
    if (first_use_on_device(attr_once)) {
#define XL_TC_ATTR(BN_) \
        XL_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN_, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem(BN_))); \
        XL_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN_, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem(BN_))); \
        XL_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN_, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem(BN_)))
        XL_TC_ATTR(256);
        XL_TC_ATTR(128);
        XL_TC_ATTR(64);
#undef XL_TC_ATTR
    }
    XL_REQUIRE(p.M > 0 && p.N > 0 && p.kblocks > 0 && batch > 0 && p.nb2 >= 1 && batch % p.nb2 == 0,
               "tc_gemm: bad shape M=%d N=%d kblocks=%d batch=%d", p.M, p.N, p.kblocks, batch);
    XL_REQUIRE(bn == 64 || bn == 128 || bn == 256, "tc_gemm: tile N must be 64, 128 or 256 (B tensor map box: 64 rows for 64, else 128)");
    XL_REQUIRE((p.C != nullptr) != (p.Cs != nullptr), "tc_gemm: exactly one of the fp32 / split-fp16 outputs");
    const int tiles_n = ceil_div(p.N, bn), tiles_m = ceil_div(p.M, kBM);
    const int64_t total = (int64_t)tiles_n * tiles_m * batch;
    XL_REQUIRE(total < (1ll << 31), "tc_gemm: too many tiles");
    const int grid = (int)(total < kNumSMs ? total : kNumSMs);  // persistent: one CTA per SM
    const int nb1 = batch / p.nb2;
    // output through TMA stores when the layout satisfies the tensor-map rules (16 B-aligned base and strides)
    CUtensorMap tmC = tmA, tmS = tmA;  // placeholders when unused
    bool tma_epi = false;
    if (p.C) {
        const bool ok = (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 && p.ldc % 4 == 0 && p.c1 % 4 == 0 && p.c2 % 4 == 0 &&
                        (!p.residual || (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0) &&
                        (nb1 == 1 || p.c1 > 0) && (p.nb2 == 1 || p.c2 > 0);
        if (ok) {
            const uint64_t dims[4] = {(uint64_t)p.N, (uint64_t)p.M, (uint64_t)p.nb2, (uint64_t)nb1};
            const uint64_t strides[3] = {(uint64_t)p.ldc * 4, (uint64_t)(p.nb2 > 1 ? p.c2 : p.ldc * (int64_t)p.M) * 4,
                                         (uint64_t)(nb1 > 1 ? p.c1 : p.ldc * (int64_t)p.M * p.nb2) * 4};
            const uint32_t box[4] = {32, kBM, 1, 1};
            if (int e = encode_tensor_map(&tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p.C, dims, strides, box,
                                          CU_TENSOR_MAP_SWIZZLE_128B)) return e;
            tma_epi = true;
        }
    } else {
        const bool ok = (reinterpret_cast<uintptr_t>(p.Cs) & 15) == 0 && p.lds % 8 == 0 && p.cs1 % 8 == 0 && p.N % 32 == 0 &&
                        p.cs_lo_off % 8 == 0 && p.cs2 % 8 == 0 && (nb1 == 1 || p.cs1 > 0);
        if (ok) {
            const uint64_t dims[3] = {(uint64_t)p.lds, (uint64_t)p.M, (uint64_t)nb1};
            const uint64_t strides[2] = {(uint64_t)p.lds * 2, (uint64_t)(nb1 > 1 ? p.cs1 : p.lds * (int64_t)p.M) * 2};
            const uint32_t box[3] = {32, kBM, 1};
            if (int e = encode_tensor_map(&tmS, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, p.Cs, dims, strides, box,
                                          CU_TENSOR_MAP_SWIZZLE_64B)) return e;
            tma_epi = true;
        }
    }
    XL_REQUIRE(!p.c_add || (tma_epi && p.C && !p.residual), "tc_gemm: C += needs an fp32 output that TMA can address and no residual");
    const bool res = tma_epi && p.C != nullptr && p.residual != nullptr;   // (the non-TMA fallback epilogue reads the residual itself)
    // 256-wide tiles with a K-major B: CTA pairs (M = 256) fetch a third less from L2 per MMA, which is what bounds these shapes
    // -- when their rounds cost less than the single-CTA waves: a pair tile takes ~0.9 of a single tile's time (measured), but
    // pair tiles quantise on 74 pairs (odd row-block counts add a ghost block per batch item)
    if (bn == 256 && tma_epi && !p.b_mn && p.N % 256 == 0) {
        const int64_t ptiles = (int64_t)(p.N / 256) * ((tiles_m + 1) / 2) * batch;
        const int64_t rounds_pair = ceil_div64(ptiles, kNumSMs / 2), waves_single = ceil_div64(total, kNumSMs);
        if (ptiles >= kNumSMs / 2 && 9 * rounds_pair < 10 * waves_single) return tc_gemm_pair(tmA, tmB, tmC, tmS, p, batch, res, st);
    }
#define XL_TC_LAUNCH(BN_, EPI_, RES_) \
    XL_CUDA(launch_pdl(gemm_tc_kernel<BN_, EPI_, RES_>, dim3(grid), dim3(kTcThreads), tc_smem(BN_), st, tmA, tmB, tmC, tmS, p, tiles_n, tiles_m, (int)total))
#define XL_TC_PICK(BN_) \
    do { if (!tma_epi) XL_TC_LAUNCH(BN_, false, false); else if (res) XL_TC_LAUNCH(BN_, true, true); else XL_TC_LAUNCH(BN_, true, false); } while (0)
    if (bn == 256) XL_TC_PICK(256);
    else if (bn == 128) XL_TC_PICK(128);
    else XL_TC_PICK(64);
#undef XL_TC_PICK
#undef XL_TC_LAUNCH
    return check_launch("gemm_tc_kernel");
}

// Tile width of a [M, N] x batch GEMM on 148 persistent CTAs: minimise  waves(bn) x (bn + 64)  -- the time of a tile grows
// with its width plus the fixed cost of streaming the A rows -- over the widths the shape allows; ties go to the wider tile
// (least L2 -> SM operand traffic per MMA).  Large batches end up on 256-wide tiles; the per-image loops of the drop-in
// surface (M = one image's tokens: 9 row tiles) get one full wave of narrow tiles for the N = 768 layers and a single wave
// of 256-wide tiles for in_proj / c_fc instead of two waves of 128-wide ones.
int tc_pick_bn(int64_t M, int N, int batch) {
    if (N <= 64) return 64;
    const int64_t mt = ceil_div64(M, kBM) * batch;
    int best = 0;
    int64_t best_cost = 0;
    for (int bn = 256; bn >= 64; bn /= 2) {
        if (bn == 256 && N % 256 != 0) continue;
        const int64_t tiles = mt * ceil_div(N, bn);
        const int64_t cost = ceil_div64(tiles, kNumSMs) * (bn + 64);
        if (best == 0 || cost < best_cost) { best = bn; best_cost = cost; }
    }
    return best;
}

int split_f16(const float* x, int64_t ldx, int rows, int cols, int Kp, __half* out, cudaStream_t st, float scale) {
    XL_REQUIRE(rows >= 0 && cols >= 0 && Kp >= cols && Kp % 64 == 0, "split_f16: bad shape");
    if (rows == 0) return 0;
    // rows ride on grid.y (<= 65535): fold larger row counts into several launches
    for (int r0 = 0; r0 < rows; r0 += 65535) {
        const int nr = rows - r0 < 65535 ? rows - r0 : 65535;
        dim3 grid(ceil_div(Kp, 256), nr);
        split_f16_kernel<<<grid, 256, 0, st>>>(x + (int64_t)r0 * ldx, ldx, nr, cols, Kp, out + (int64_t)r0 * 2 * Kp, scale);
        if (int e = check_launch("split_f16_kernel")) return e;
    }
    return 0;
}

}  // namespace xl

using namespace xl;

// Stand-alone entry point (tests, small callers): C = alpha * A B^T (+bias, act, +residual) for fp32 row-major
// A [M,K], B [N,K]; operands are split to fp16 pairs into ws (>= 2*(M+N)*Kp halves, Kp = round_up(K,64)).
extern "C" int excel_gemm_tc(const float* A, const float* B, float* C, const float* bias, const float* residual, int M, int N,
                             int K, int64_t lda, int64_t ldb, int64_t ldc, float alpha, int act, void* ws, int64_t ws_bytes,
                             void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    XL_REQUIRE(M >= 1 && N >= 1 && K >= 1, "gemm_tc: bad shape");
    const int Kp = (K + 63) & ~63;
    XL_REQUIRE(ws_bytes >= (int64_t)2 * (M + N) * Kp * 2, "gemm_tc: workspace too small");
    __half* As = reinterpret_cast<__half*>(ws);
    __half* Bs = As + (int64_t)M * 2 * Kp;
    if (int e = split_f16(A, lda, M, K, Kp, As, st)) return e;
    if (int e = split_f16(B, ldb, N, K, Kp, Bs, st)) return e;
    CUtensorMap tmA, tmB;
    const int bn = N <= 64 ? 64 : (N % 256 == 0 ? 256 : 128);
    if (int e = make_operand_map(&tmA, As, M, 2 * Kp, 2 * Kp, 128)) return e;
    if (int e = make_operand_map(&tmB, Bs, N, 2 * Kp, 2 * Kp, bn == 64 ? 64 : 128)) return e;
    TcParams p = {};
    p.M = M; p.N = N; p.kblocks = Kp / 64; p.a_lo_off = Kp; p.b_lo_off = Kp; p.nb2 = 1;
    p.C = C; p.ldc = ldc; p.bias = bias; p.residual = residual; p.alpha = alpha; p.act = act;
    return tc_gemm(tmA, tmB, p, 1, bn, st);
}

// Batched GEMM on operands that are ALREADY in the engine's split-fp16 format (include/excel_b200.h).
extern "C" int excel_gemm_tc_split(const void* As, int64_t lda, int a_lo_off, int64_t a_rows_z, const void* Bs, int64_t ldb,
                                   int b_lo_off, int64_t b_rows_z, float* C, int64_t ldc, int64_t c_z, void* Cs, int64_t lds,
                                   int cs_lo_off, int64_t cs_z, const float* bias, int64_t bias_z, int M, int N, int K, int batch,
                                   float alpha, int act, void* stream) {
    XL_REQUIRE(M >= 1 && N >= 1 && K >= 64 && K % 64 == 0 && batch >= 1, "gemm_tc_split: bad shape M=%d N=%d K=%d batch=%d (K %% 64 == 0)", M, N, K, batch);
    XL_REQUIRE(As && Bs && lda % 8 == 0 && ldb % 8 == 0 && (reinterpret_cast<uintptr_t>(As) & 15) == 0 && (reinterpret_cast<uintptr_t>(Bs) & 15) == 0,
               "gemm_tc_split: operands must be 16 B-aligned with row pitches that are multiples of 8 halves");
    XL_REQUIRE(a_rows_z * (batch - 1) + M < (1ll << 31) && b_rows_z * (batch - 1) + N < (1ll << 31), "gemm_tc_split: too many rows");
    const int bn = tc_pick_bn(M, N, batch);
    CUtensorMap tmA, tmB;
    if (int e = make_operand_map(&tmA, As, a_rows_z * (batch - 1) + M, a_lo_off + K, lda, 128)) return e;
    if (int e = make_operand_map(&tmB, Bs, b_rows_z * (batch - 1) + N, b_lo_off + K, ldb, bn == 64 ? 64 : 128)) return e;
    TcParams p = {};
    p.M = M; p.N = N; p.kblocks = K / 64; p.a_lo_off = a_lo_off; p.b_lo_off = b_lo_off; p.nb2 = 1;
    p.a_row1 = (int)a_rows_z; p.b_row1 = (int)b_rows_z;
    p.C = C; p.ldc = ldc; p.c1 = c_z; p.bias = bias; p.bias1 = bias_z; p.alpha = alpha; p.act = act;
    p.Cs = reinterpret_cast<__half*>(Cs); p.lds = lds; p.cs1 = cs_z; p.cs_lo_off = cs_lo_off;
    return tc_gemm(tmA, tmB, p, batch, bn, (cudaStream_t)stream);
}
