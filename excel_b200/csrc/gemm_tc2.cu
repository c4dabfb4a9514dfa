// CTA-PAIR (tcgen05 cta_group::2) form of the split-fp16 GEMM engine of gemm_tc.cu for its 256-wide tiles.
//
// Why: a 128 x 256 tile moves 96 KB of operands per 64-k block (A hi | lo 32 KB + B hi | lo 64 KB) for 1536 clocks of tensor
// work.  On 148 SMs that is 6.6 KB per clock from L2 -- the chip's L2 -> SM throughput ceiling (measured: in_proj 1.34 GB in
// 202.9 k cycles, c_fc 1.78 GB in 270 k cycles, both 6.6 KB/clk), so the 256-wide GEMMs of the encoder were L2-bound at a
// tensor pipe 81-84 % active.  Two CTAs of a cluster (the two SMs of a TPC) that own ADJACENT 128-row blocks of the same
// 256 output columns run one tcgen05.mma.cta_group::2 (M = 256, N = 256): each CTA fetches its own A rows and only HALF of
// the B tile (64 KB per k-block and CTA instead of 96 KB), and the ring holds three stages instead of two.
//
// Protocol (rank 0 = leader), as in the attention pair experiment (tools/experiments/attn_tc2.cu):
//   * both CTAs' producer warps issue their own TMA loads; all of them complete on the LEADER's full barrier of the stage
//     (cp.async.bulk.tensor ... .cta_group::2), whose expect-tx covers both CTAs' bytes;
//   * only the leader's MMA warp issues tcgen05.mma.cta_group::2; tcgen05.commit ... .multicast::cluster arrives on the
//     stage-empty / accumulator-full barriers of BOTH CTAs;
//   * each CTA's epilogue warps read their own tensor memory (their own 128 output rows) and arrive REMOTELY on the leader's
//     accumulator-empty barrier; the epilogue itself (alpha / bias / activation / residual / TMA store or reduce-add, fp32 or
//     split-fp16 output) is the single-CTA kernel's.
// A pair tile is (batch z, row-block pair, 256-column tile); with an odd number of row blocks the last pair's second CTA owns
// rows past M: its loads are clipped / harmless, its stores are clipped by the tensor map.
#include <cuda_fp16.h>

#include "common.cuh"
#include "excel_b200.h"
#include "gemm_tc.cuh"
#include "ptx.cuh"
#include "tc.cuh"

namespace xl {

namespace {

constexpr int k2BN = 256;
constexpr int k2BK = 64;
constexpr uint32_t k2Tile = kBM * k2BK * 2;              // 16 KB: 128 rows x 64 k fp16
constexpr uint32_t k2Stage = 4 * k2Tile;                 // A hi | A lo | B-half hi | B-half lo
constexpr int k2Stages = 3;
constexpr int k2Threads = 64 + 256;
constexpr uint32_t k2Epi = 2 * 16384;
constexpr size_t k2Smem = k2Stages * k2Stage + k2Epi + 1024 /*align*/ + 256 /*barriers*/;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(const void* p, uint32_t rank) {
    uint32_t a;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(smem_u32(p)), "r"(rank));
    return a;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* tm, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// D fp32, A/B fp16, K-major, M = 256 (two CTAs x 128 rows), N = 256
constexpr uint32_t k2Idesc = (1u << 4) | ((uint32_t)(k2BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

}  // namespace

// num_tiles pair tiles; tiles_mp = row-block PAIRS per batch item
template <bool RES>
__global__ void __launch_bounds__(k2Threads, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmS, const TcParams p,
                int tiles_n, int tiles_mp, int num_tiles) {
    constexpr int kBN = k2BN;
    constexpr bool TMA_EPI = true;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* stage = reinterpret_cast<float*>(tiles + k2Stages * k2Stage);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + k2Stages * k2Stage + k2Epi);   // the leader's are used
    uint64_t* empty_bar = full_bar + k2Stages;                                              // per CTA (multicast commit)
    uint64_t* acc_full = empty_bar + k2Stages;                                              // [2] per CTA (multicast commit)
    uint64_t* acc_empty = acc_full + 2;                                                     // [2] the leader's are used
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool is_leader = rank == 0;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < k2Stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], 16);   // 8 epilogue warps of each CTA
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair(tmem_slot, 2 * kBN);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // both CTAs' barriers exist before any remote arrive / multicast commit / remote complete_tx
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    // pair tile t -> this CTA's 128-row block (m0), the 256-column tile (n0), the batch item (z1, z2)
    auto decode = [&](int t, int& m0, int& n0, int& z1, int& z2) {
        const int nt = t % tiles_n, r = t / tiles_n;
        const int mp = r % tiles_mp, z = r / tiles_mp;
        m0 = (2 * mp + (int)rank) * kBM; n0 = nt * kBN; z1 = z / p.nb2; z2 = z % p.nb2;
    };

    if (warp == 0) {
        // ---- TMA producer of this CTA: its own A rows and its half (128 of the 256 rows) of the B tile
        const bool leader = elect_one_sync();
        if (leader) {
            tma_prefetch_desc(&tmA);
            tma_prefetch_desc(&tmB);
        }
        int it = 0;
        for (int t = pair; t < num_tiles; t += npairs) {
            int m0, n0, z1, z2;
            decode(t, m0, n0, z1, z2);
            const int a_row = p.a_row0 + z1 * p.a_row1 + z2 * p.a_row2 + m0, a_col = p.a_col0 + z1 * p.a_col1 + z2 * p.a_col2;
            const int b_row = p.b_row0 + z1 * p.b_row1 + z2 * p.b_row2 + n0 + (int)rank * 128;
            const int b_col = p.b_col0 + z1 * p.b_col1 + z2 * p.b_col2;
            for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
                const int s = it % k2Stages;
                mbar_wait(&empty_bar[s], ((it / k2Stages) & 1) ^ 1);
                uint8_t* st = tiles + s * k2Stage;
                if (leader) {
                    const uint32_t fb = map_to_rank(&full_bar[s], 0);
                    if (is_leader) mbar_arrive_expect_tx(&full_bar[s], 2 * k2Stage);   // both CTAs' bytes
                    tma_load_2d_pair(st, &tmA, fb, a_col + kb * k2BK, a_row);
                    tma_load_2d_pair(st + k2Tile, &tmA, fb, a_col + p.a_lo_off + kb * k2BK, a_row);
                    tma_load_2d_pair(st + 2 * k2Tile, &tmB, fb, b_col + kb * k2BK, b_row);
                    tma_load_2d_pair(st + 3 * k2Tile, &tmB, fb, b_col + p.b_lo_off + kb * k2BK, b_row);
                }
            }
        }
    } else if (warp == 1) {
        if (is_leader) {
            // ---- MMA issuer of the pair
            const bool leader = elect_one_sync();
            const uint32_t tiles0 = smem_u32(tiles);
            int it = 0, i = 0;
            for (int t = pair; t < num_tiles; t += npairs, ++i) {
                const int buf = i & 1;
                mbar_wait(&acc_empty[buf], ((i >> 1) & 1) ^ 1);  // both CTAs' epilogues have drained this accumulator
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(buf * kBN);
                for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
                    const int s = it % k2Stages;
                    mbar_wait(&full_bar[s], (it / k2Stages) & 1);
                    tc_fence_after();
                    const uint32_t st = tiles0 + s * k2Stage;
                    const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + k2Tile);
                    const uint64_t b_hi = umma_desc_sw128(st + 2 * k2Tile), b_lo = umma_desc_sw128(st + 3 * k2Tile);
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < k2BK / 16; ++k) {
                            umma_f16_pair(tacc, a_hi + 2 * k, b_lo + 2 * k, k2Idesc, (kb | k) != 0);
                            umma_f16_pair(tacc, a_lo + 2 * k, b_hi + 2 * k, k2Idesc, 1);
                            umma_f16_pair(tacc, a_hi + 2 * k, b_hi + 2 * k, k2Idesc, 1);
                        }
                        umma_commit_pair(&empty_bar[s]);
                    }
                }
                if (leader) umma_commit_pair(&acc_full[buf]);
            }
        }
    } else {
        // ---- epilogue of this CTA's 128 rows (warps 2..9): the single-CTA kernel's, with a remote accumulator-empty arrival
        const int lg = warp & 3, team = (warp - 2) >> 2;
        const int trow = lg * 32 + lane;
        const float alpha = p.alpha;
        const int act = p.act;
        int i = 0;
        {
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
                else if (tt + npairs < num_tiles) fetch_residual(tt + npairs, team);
            };
            constexpr bool has_res = RES;
            if constexpr (has_res) {
                if (pair < num_tiles) fetch_residual(pair, team);
            }
            for (int t = pair; t < num_tiles; t += npairs, ++i) {
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
                        if (lane == 0) mbar_arrive_cluster(map_to_rank(&acc_empty[buf], 0));   // the leader's barrier
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
        }
    }
    // neither CTA may leave (or free its tensor memory) while the pair's MMAs can still read its shared memory
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 2 * kBN);
    }
}

template <bool RES>
static int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmS, const TcParams& p,
                       int tiles_n, int tiles_mp, int total, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kNumSMs);
    cfg.blockDim = dim3(k2Threads);
    cfg.dynamicSmemBytes = k2Smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    static int max_pairs[2] = {0, 0};
    if (max_pairs[RES] == 0) {   // persistent pairs: as many clusters as the device holds at once
        int n = 0;
        cfg.numAttrs = 1;
        XL_CUDA(cudaOccupancyMaxActiveClusters(&n, gemm_tc2_kernel<RES>, &cfg));
        cfg.numAttrs = 2;
        max_pairs[RES] = n > 0 ? n : 1;
    }
    const int np = total < max_pairs[RES] ? total : max_pairs[RES];
    cfg.gridDim = dim3(2 * np);
    XL_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc2_kernel<RES>, tmA, tmB, tmC, tmS, p, tiles_n, tiles_mp, total));
    return 0;
}

// Called by tc_gemm for 256-wide, K-major-B, TMA-addressable outputs with enough pair tiles to fill the chip.
int tc_gemm_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmS, const TcParams& p,
                 int batch, bool res, cudaStream_t st) {
    static unsigned long long attr_once = 0;
    if (first_use_on_device(attr_once)) {
        XL_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2Smem));
        XL_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2Smem));
    }
    const int tiles_n = p.N / k2BN, tiles_mp = (ceil_div(p.M, kBM) + 1) / 2;
    const int total = tiles_n * tiles_mp * batch;
    if (res) { if (int e = launch_pair<true>(tmA, tmB, tmC, tmS, p, tiles_n, tiles_mp, total, st)) return e; }
    else { if (int e = launch_pair<false>(tmA, tmB, tmC, tmS, p, tiles_n, tiles_mp, total, st)) return e; }
    return check_launch("gemm_tc2_kernel");
}

}  // namespace xl
