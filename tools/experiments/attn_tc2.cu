// EXPERIMENT (round 2), not part of the product build: measured SLOWER than the single-CTA kernels of csrc/attn_tc.cu.
//   1 x B200, cfg2 (16 x 512^2, N = 1025), ncu launch list, same box:  stats pass 1 set 104 us (single CTA: 97), 3 sets 294 (265),
//   map pass 360 (320); all 65 GPU parity tests pass with it routed in (vit.cu: scores() -> attn_scores_pair when N > 128).
// What it showed: relieving the shared-memory operand bandwidth (the pair reads 6 KB instead of 8 KB per k-step and fetches
// half of every key tile) buys nothing, so the S MMAs are NOT what bounds these passes.  A 128 x 128 fp32 score tile costs 1024
// clocks of tensor-memory read (tcgen05.ld: 64 B/clk per SM) and 1024 clocks of MUFU.EX2 (16 per clock per SM); the kernels
// run at 1750-1930 clocks per tile -- between the overlapped and the serialised sum of those two -- and the odd ninth row block
// costs a pair a whole ghost CTA (five pairs for nine blocks).  Also measured here: polling the pair's barriers with
// mbarrier.try_wait.acquire.cluster costs +90 % (map pass 612 us); plain CTA-scope waits are what CUTLASS uses too.  Software-
// pipelining the tcgen05.ld's under the exp2 stream (below, MODE 0) changed nothing (107 us).
// To rebuild: copy into excel_b200/csrc/, declare attn_scores_pair in attn_tc.cuh, route it from vit.cu scores().
//
// CTA-PAIR (tcgen05 cta_group::2) form of the attention statistics / map passes of attn_tc.cu (reference: nn.MultiheadAttention
// with need_weights and the surgery Attention.forward, clip/clip_surgery_model.py:95-159,297-307).
//
// Why: at head dim 64 a single-CTA S tile (M128 x N128 x K64, three split-fp16 passes) reads 96 KB of operands from shared
// memory for 768 clocks of tensor-core math -- exactly the SM's 128 B/clk, so with the TMA fills and the epilogue staging on
// the same banks the MMAs run shared-memory-bound (tensor pipe 50 % in the single-CTA kernels).  Two CTAs of a cluster (the
// two SMs of a TPC) that own ADJACENT 128-row query blocks of the same (image, head) share every key tile: one
// tcgen05.mma.cta_group::2 (M = 256) reads each CTA's own 128 rows of X and only HALF of the key tile from each CTA's shared
// memory (6 KB per 64-clock k-step instead of 8 KB), and each CTA fetches only its half of Y through TMA.
//
// Protocol (per pair; rank 0 = leader):
//   * both CTAs' producer warps issue their own TMA loads; all loads complete on the LEADER's full barriers
//     (cp.async.bulk.tensor ... .cta_group::2 with the leader's mbarrier address), whose expect-tx covers both CTAs' bytes;
//   * only the leader's MMA warp issues tcgen05.mma.cta_group::2; tcgen05.commit ... .multicast::cluster arrives on the
//     stage-empty / accumulator-full barriers of BOTH CTAs (same shared-memory offsets);
//   * each CTA's epilogue warps read their own tensor memory (rows of their own query block) and arrive REMOTELY on the
//     leader's accumulator-empty barrier.
// Work item of a pair: MODE 0 (stats) = (image, score set x head, row-block pair) walking the key blocks;
//                      MODE 1 (map)   = (image, row-block pair, key block) walking the (score set, head) pairs.
// With an odd number of row blocks (N = 1025: nine) the last pair's second CTA owns no rows: it still supplies its half of
// the key tiles, its epilogue warps skip their arithmetic.  Key blocks are trimmed to the valid keys rounded up to 16; CTA r
// supplies keys [r n/2, (r+1) n/2) of an n-wide block, so the columns of the S tile stay in key order.
#include <cuda_fp16.h>

#include <type_traits>

#include "attn_tc.cuh"
#include "common.cuh"
#include "excel_b200.h"
#include "tc.cuh"

namespace xl {

// (declare in attn_tc.cuh when building this into the library)
int attn_scores_pair(const CUtensorMap& tmQ, const CUtensorMap& tmY, const AttnParams& p, cudaStream_t st, bool stats_only);

namespace {

constexpr int k2Stages1 = 4;                              // MODE 1 ring: X (hi | lo, 32 KB) + Y half (hi | lo, 16 KB) per stage
constexpr int k2YStages0 = 6;                             // MODE 0: Y-half ring (16 KB stages) beside a double-buffered X tile
constexpr int k2Acc = 4;                                  // TMEM accumulators (4 x 128 columns per CTA)
constexpr uint32_t k2Tile = 128 * 64 * 2;                 // 16 KB: one 128-row x 64-k fp16 tile
constexpr uint32_t k2Half = 64 * 64 * 2;                  // 8 KB: 64 keys x 64 k
constexpr uint32_t k2Stage1 = 2 * k2Tile + 2 * k2Half;    // 48 KB
constexpr int k2EpiWarps = 16;
constexpr int k2Threads = 64 + 32 * k2EpiWarps;
constexpr uint32_t k2Epi = k2EpiWarps * 2048;
constexpr size_t k2Smem = k2Stages1 * k2Stage1 + k2Epi + 1024 /*align*/ + 256 /*barriers*/;
static_assert(2 * 2 * k2Tile + k2YStages0 * 2 * k2Half <= k2Stages1 * k2Stage1, "MODE 0 plan fits the MODE 1 ring");
constexpr int k2MaxSt = k2YStages0 > k2Stages1 ? k2YStages0 : k2Stages1;

__device__ __forceinline__ float ex2_approx2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// asynchronous tcgen05.ld (issue only; complete with tmem_ld_wait) -- lets a warp keep a load in flight under its arithmetic
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- cluster / cta_group::2 PTX ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared-memory object of this CTA's layout) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_rank(const void* p, uint32_t rank) {
    uint32_t a;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(smem_u32(p)), "r"(rank));
    return a;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a barrier of this CTA that the pair arrives on (plain CTA-scope acquire: the data these barriers guard moves through
// the async proxy / tensor memory and is ordered by complete_tx and the tcgen05 fences; a cluster-scope acquire per poll is costly)
__device__ __forceinline__ void mbar_wait_cl(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// TMA tile load of this CTA into its own shared memory, completing on an mbarrier of EITHER CTA of the pair
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
// arrive on the barrier at this shared-memory offset in BOTH CTAs when all previously issued MMAs of this thread are complete
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// instruction descriptor: D fp32, A/B fp16, K-major, M = 256 (two CTAs x 128 rows), N = n
__device__ __forceinline__ uint32_t make_idesc_pair(int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

}  // namespace

template <int MODE>
__global__ void __launch_bounds__(k2Threads, 1)
attn_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmO,
                const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* ebuf = tiles + k2Stages1 * k2Stage1;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ebuf + k2Epi);   // leader's are used (both CTAs' loads complete there)
    uint64_t* empty_bar = full_bar + k2MaxSt;                         // per CTA (multicast commit)
    uint64_t* acc_full = empty_bar + k2MaxSt;                         // per CTA (multicast commit)
    uint64_t* acc_empty = acc_full + k2Acc;                           // leader's are used (remote arrivals of both CTAs' epilogues)
    uint64_t* x_full = acc_empty + k2Acc;                             // [2] MODE 0, leader's
    uint64_t* x_empty = x_full + 2;                                   // [2] MODE 0, per CTA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_empty + 2);
    uint8_t* xbuf = tiles;                                            // MODE 0: X double buffer (2 x 32 KB), then the Y-half ring
    uint8_t* yring = tiles + 2 * 2 * k2Tile;
    constexpr int NST = MODE == 0 ? k2YStages0 : k2Stages1;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool is_leader = rank == 0;
    const int pair = blockIdx.x >> 1, npairs_grid = gridDim.x >> 1;
    const int nblk = (p.N + 127) / 128, npr = (nblk + 1) / 2;         // row blocks == key blocks; row-block pairs
    const int TH = p.ntypes * p.H;
    const int inner = MODE == 0 ? nblk : TH;
    const int items = MODE == 0 ? p.B * TH * npr : p.B * npr * nblk;
    constexpr int kReaders = MODE == 0 ? k2EpiWarps / 2 : k2EpiWarps;   // epilogue warps per CTA that read one accumulator

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&x_full[s], 1);
            mbar_init(&x_empty[s], 1);
        }
        for (int s = 0; s < k2Acc; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], 2 * kReaders);   // both CTAs' reader warps
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair(tmem_slot, 128 * k2Acc);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();           // both CTAs' barriers are initialised before any remote arrive / multicast commit / remote complete_tx
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto decode = [&](int item, int j, int& b, int& h, int& rb, int& kb) {
        if (MODE == 0) {   // item = (b, th, row-block pair), j = kb
            const int rp = item % npr, r = item / npr;
            h = r % TH; b = r / TH; kb = j; rb = 2 * rp + (int)rank;
        } else {           // item = (b, row-block pair, kb), j = th
            kb = item % nblk; const int r = item / nblk;
            rb = 2 * (r % npr) + (int)rank; b = r / npr; h = j;
        }
    };
    // width of key block kb as the MMA computes it (valid keys rounded up to 16)
    auto kb_width = [&](int kb) { return (min(128, p.N - kb * 128) + 15) & ~15; };

    if (warp == 0) {
        // ---- TMA producer of this CTA: its own X rows and its HALF of every key tile; completion on the leader's barriers
        const bool leader_lane = elect_one_sync();
        if (leader_lane) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmY); }
        int it = 0, ni = 0;
        for (int item = pair; item < items; item += npairs_grid, ++ni)
            for (int j = 0; j < inner; ++j, ++it) {
                int b, h, rb, kb;
                decode(item, j, b, h, rb, kb);
                const int nw = kb_width(kb);
                const int xr = b * p.N + rb * 128, yr = b * p.N + kb * 128 + (int)rank * (nw >> 1);
                const int ty = h / p.H, hd = h - ty * p.H;
                const int xc = p.xo[ty] + hd * 64, yc = p.yo[ty] + hd * 64;
                if (MODE == 0) {
                    if (j == 0) {
                        const int xs = ni & 1;
                        mbar_wait_cl(&x_empty[xs], ((ni >> 1) & 1) ^ 1);
                        if (leader_lane) {
                            const uint32_t xf = map_to_rank(&x_full[xs], 0);
                            if (is_leader) mbar_arrive_expect_tx(&x_full[xs], 2 * 2 * k2Tile);   // both CTAs' X tiles
                            tma_load_2d_pair(xbuf + xs * 2 * k2Tile, &tmQ, xf, xc, xr);
                            tma_load_2d_pair(xbuf + xs * 2 * k2Tile + k2Tile, &tmQ, xf, xc + p.lo_off, xr);
                        }
                    }
                    const int s = it % NST;
                    mbar_wait_cl(&empty_bar[s], ((it / NST) & 1) ^ 1);
                    uint8_t* st = yring + s * 2 * k2Half;
                    if (leader_lane) {
                        const uint32_t fb = map_to_rank(&full_bar[s], 0);
                        if (is_leader) mbar_arrive_expect_tx(&full_bar[s], 2 * 2 * k2Half);
                        tma_load_2d_pair(st, &tmY, fb, yc, yr);
                        tma_load_2d_pair(st + k2Half, &tmY, fb, yc + p.lo_off, yr);
                    }
                } else {
                    const int s = it % NST;
                    mbar_wait_cl(&empty_bar[s], ((it / NST) & 1) ^ 1);
                    uint8_t* st = tiles + s * k2Stage1;
                    if (leader_lane) {
                        const uint32_t fb = map_to_rank(&full_bar[s], 0);
                        if (is_leader) mbar_arrive_expect_tx(&full_bar[s], 2 * k2Stage1);
                        tma_load_2d_pair(st, &tmQ, fb, xc, xr);
                        tma_load_2d_pair(st + k2Tile, &tmQ, fb, xc + p.lo_off, xr);
                        tma_load_2d_pair(st + 2 * k2Tile, &tmY, fb, yc, yr);
                        tma_load_2d_pair(st + 2 * k2Tile + k2Half, &tmY, fb, yc + p.lo_off, yr);
                    }
                }
            }
    } else if (warp == 1) {
        if (is_leader) {
            // ---- MMA issuer of the pair
            const bool leader_lane = elect_one_sync();
            const uint32_t tiles0 = smem_u32(tiles);
            int it = 0, ni = 0;
            for (int item = pair; item < items; item += npairs_grid, ++ni)
                for (int j = 0; j < inner; ++j, ++it) {
                    int b, h, rb, kb;
                    decode(item, j, b, h, rb, kb);
                    const int buf = it % k2Acc, s = it % NST;
                    mbar_wait_cl(&acc_empty[buf], ((it / k2Acc) & 1) ^ 1);
                    if (MODE == 0 && j == 0) mbar_wait_cl(&x_full[ni & 1], (ni >> 1) & 1);
                    mbar_wait_cl(&full_bar[s], (it / NST) & 1);
                    tc_fence_after();
                    const uint32_t idesc = make_idesc_pair(kb_width(kb));
                    const uint32_t tacc = tmem_base + (uint32_t)(buf * 128);
                    const uint32_t xa = MODE == 0 ? tiles0 + (ni & 1) * 2 * k2Tile : tiles0 + s * k2Stage1;
                    const uint32_t yb = MODE == 0 ? tiles0 + 4 * k2Tile + s * 2 * k2Half : tiles0 + s * k2Stage1 + 2 * k2Tile;
                    const uint64_t a_hi = umma_desc_sw128(xa), a_lo = umma_desc_sw128(xa + k2Tile);
                    const uint64_t b_hi = umma_desc_sw128(yb), b_lo = umma_desc_sw128(yb + k2Half);
                    if (leader_lane) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            umma_f16_pair(tacc, a_hi + 2 * k, b_lo + 2 * k, idesc, k != 0);
                            umma_f16_pair(tacc, a_lo + 2 * k, b_hi + 2 * k, idesc, 1);
                            umma_f16_pair(tacc, a_hi + 2 * k, b_hi + 2 * k, idesc, 1);
                        }
                        umma_commit_pair(&empty_bar[s]);
                        if (MODE == 0 && j == inner - 1) umma_commit_pair(&x_empty[ni & 1]);
                        umma_commit_pair(&acc_full[buf]);
                    }
                }
        }
    } else {
        // ---- epilogue of this CTA's row block: warp (lg, qt): TMEM lanes 32*lg..+31 (rows), columns 32*qt..+31 of every S tile
        const int ew = warp - 2, lg = warp & 3, qt = ew >> 2;
        const int trow = lg * 32 + lane;
        float* wbuf = reinterpret_cast<float*>(ebuf) + ew * 512;
        int it = 0;
        for (int item = pair; item < items; item += npairs_grid) {
            int b, h, rb, kb;
            decode(item, 0, b, h, rb, kb);
            const int row = rb * 128 + trow;
            const bool row_ok = row < p.N;
            const bool rows_empty = rb * 128 + lg * 32 >= p.N;    // none of this warp's rows exists (also: the ghost CTA of an odd pair)
            float m_run = -INFINITY, l_run = 0.f;
            float acc[32];
            if (MODE == 1) {
#pragma unroll
                for (int e = 0; e < 32; ++e) acc[e] = 0.f;
            }
            float m_next = INFINITY;
            if (MODE == 1 && row_ok) m_next = __ldg(p.m + ((int64_t)b * p.H) * p.N + row);
            for (int j = 0; j < inner; ++j, ++it) {
                decode(item, j, b, h, rb, kb);
                const int buf = it % k2Acc;
                const uint32_t ae = map_to_rank(&acc_empty[buf], 0);   // the leader's accumulator-empty barrier
                const float m_row = m_next;
                if (MODE == 1 && row_ok && j + 1 < inner) {
                    const int hn = j + 1, ty = hn / p.H, hd = hn - ty * p.H;
                    m_next = __ldg(p.m + (((int64_t)ty * p.B + b) * p.H + hd) * p.N + row);
                }
                if (MODE == 0) {
                    const int grp = qt >> 1, cq = qt & 1;
                    if ((it & 1) != grp) continue;
                    mbar_wait_cl(&acc_full[buf], (it / k2Acc) & 1);
                    tc_fence_after();
                    if (rows_empty) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(ae);
                        continue;
                    }
                    // both 32-column chunks of the tile are requested before the first is consumed: the second tcgen05.ld is in flight
                    // under the first chunk's exp2 stream, and the accumulator goes back to the MMA warp before any arithmetic
                    auto chunks = [&](auto tail_tag) {
                    constexpr bool TAIL = decltype(tail_tag)::value;
                    const int keyb = kb * 128 + cq * 64;
                    const uint32_t ta = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * 128 + cq * 64);
                    // running row max / sum over W consecutive columns (exp2 domain; alpha > 0, so max(alpha a) = alpha max(a))
                    auto consume = [&](auto& r, int key0) {
                        constexpr int W = sizeof(r) / 4;
                        if constexpr (TAIL) {
#pragma unroll
                            for (int e = 0; e < W; ++e)
                                if (key0 + e >= p.N) r[e] = 0xff800000u;
                        }
                        float c0 = -INFINITY, c1 = -INFINITY, c2 = -INFINITY, c3 = -INFINITY;
#pragma unroll
                        for (int e = 0; e < W; e += 4) {
                            c0 = fmaxf(c0, __uint_as_float(r[e]));
                            c1 = fmaxf(c1, __uint_as_float(r[e + 1]));
                            c2 = fmaxf(c2, __uint_as_float(r[e + 2]));
                            c3 = fmaxf(c3, __uint_as_float(r[e + 3]));
                        }
                        const float m_new = fmaxf(m_run, p.alpha * fmaxf(fmaxf(c0, c1), fmaxf(c2, c3)));
                        if (m_new > -INFINITY) {
                            const float2 al = make_float2(p.alpha, p.alpha), mm = make_float2(-m_new, -m_new);
                            float2 sa = make_float2(0.f, 0.f), sb = sa;
#pragma unroll
                            for (int e = 0; e < W; e += 4) {
                                const float2 xa = __ffma2_rn(al, make_float2(__uint_as_float(r[e]), __uint_as_float(r[e + 1])), mm);
                                const float2 xb = __ffma2_rn(al, make_float2(__uint_as_float(r[e + 2]), __uint_as_float(r[e + 3])), mm);
                                sa = __fadd2_rn(sa, make_float2(ex2_approx2(xa.x), ex2_approx2(xa.y)));
                                sb = __fadd2_rn(sb, make_float2(ex2_approx2(xb.x), ex2_approx2(xb.y)));
                            }
                            l_run = l_run * ex2_approx2(m_run - m_new) + ((sa.x + sb.x) + (sa.y + sb.y));
                            m_run = m_new;
                        }
                    };
                    if constexpr (!TAIL) {
                        // four 16-column chunks, two register buffers: the load of chunk c + 2 is in flight under the exp2 stream of
                        // chunk c + 1 (tcgen05.ld completes asynchronously until tcgen05.wait::ld)
                        uint32_t r0[16], r1[16];
                        tmem_ld16_issue(ta, r0);
                        tmem_ld16_issue(ta + 16, r1);
                        tmem_ld_wait();
                        consume(r0, keyb);
                        tmem_ld16_issue(ta + 32, r0);
                        consume(r1, keyb + 16);
                        tmem_ld_wait();
                        tmem_ld16_issue(ta + 48, r1);
                        consume(r0, keyb + 32);
                        tmem_ld_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(ae);   // the accumulator goes back to the MMA warp
                        consume(r1, keyb + 48);
                    } else {   // last key block (padding keys): one chunk at a time
                        uint32_t r[32];
                        const bool c0ok = keyb < p.N, c1ok = keyb + 32 < p.N;   // (uniform)
                        if (c0ok) tmem_ld32(ta, r);
                        if (!c1ok) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_cluster(ae);
                        }
                        if (c0ok) consume(r, keyb);
                        if (c1ok) {
                            tmem_ld32(ta + 32, r);
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_cluster(ae);
                            consume(r, keyb + 32);
                        }
                    }
                    };
                    if (kb == nblk - 1) chunks(std::true_type{});
                    else chunks(std::false_type{});
                    continue;
                }
                mbar_wait_cl(&acc_full[buf], (it / k2Acc) & 1);
                tc_fence_after();
                if (rows_empty) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(ae);
                    continue;
                }
                const int key0 = kb * 128 + qt * 32;
                auto tile = [&](auto tail_tag) {
                constexpr bool TAIL = decltype(tail_tag)::value;
                if (!TAIL || key0 < p.N) {
                    uint32_t r[32];
                    tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * 128 + qt * 32), r);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(ae);
                    if constexpr (TAIL) {
#pragma unroll
                        for (int e = 0; e < 32; ++e)
                            if (key0 + e >= p.N) r[e] = 0xff800000u;
                    }
#pragma unroll
                    for (int e = 0; e < 32; ++e) acc[e] += ex2_approx2(fmaf(p.alpha, __uint_as_float(r[e]), -m_row));
                } else {
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(ae);
                }
                };
                if (kb == nblk - 1) tile(std::true_type{});
                else tile(std::false_type{});
            }
            if (MODE == 0) {
                float* xch = reinterpret_cast<float*>(ebuf);   // [3][2][128]
                bar_sync(3, 32 * k2EpiWarps);
                if (qt > 0) { xch[(qt - 1) * 256 + trow] = m_run; xch[(qt - 1) * 256 + 128 + trow] = l_run; }
                bar_sync(3, 32 * k2EpiWarps);
                if (qt == 0 && row_ok) {
                    float mf = m_run;
#pragma unroll
                    for (int q = 0; q < 3; ++q) mf = fmaxf(mf, xch[q * 256 + trow]);
                    float lf = l_run * ex2_approx2(m_run - mf);
#pragma unroll
                    for (int q = 0; q < 3; ++q) lf += xch[q * 256 + 128 + trow] * ex2_approx2(xch[q * 256 + trow] - mf);
                    const int ty = h / p.H, hd = h - ty * p.H;
                    p.m[(((int64_t)ty * p.B + b) * p.H + hd) * p.N + row] = mf + __log2f(lf) - 10.f;
                }
            } else {
                const int key0 = kb * 128 + qt * 32;
                if (p.out_split) {
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
                        if (key0 + hf * 16 >= p.N) break;
                        if (lane == 0) tma_store_wait_read<0>();
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
    // Neither CTA may leave (or free its tensor memory) while the other can still read its shared memory through the pair's
    // MMAs or arrive on its barriers.
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 128 * k2Acc);
    }
}

template <int MODE>
static int launch_pair(const CUtensorMap& tmQ, const CUtensorMap& tmY, const CUtensorMap& tmO, const AttnParams& p, int pair_items,
                       cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kNumSMs);
    cfg.blockDim = dim3(k2Threads);
    cfg.dynamicSmemBytes = k2Smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // persistent pairs: as many clusters as the device can hold at once (a TPC with one usable SM cannot host a pair)
    static int max_pairs[2] = {0, 0};
    if (max_pairs[MODE] == 0) {
        int n = 0;
        XL_CUDA(cudaOccupancyMaxActiveClusters(&n, attn_tc2_kernel<MODE>, &cfg));
        max_pairs[MODE] = n > 0 ? n : 1;
    }
    const int npairs = pair_items < max_pairs[MODE] ? pair_items : max_pairs[MODE];
    cfg.gridDim = dim3(2 * npairs);
    XL_CUDA(cudaLaunchKernelEx(&cfg, attn_tc2_kernel<MODE>, tmQ, tmY, tmO, p));
    return 0;
}

// Same contract as attn_scores (attn_tc.cu); tmY: the same split-fp16 qkv matrix with 64-row boxes (key half-tiles).
int attn_scores_pair(const CUtensorMap& tmQ, const CUtensorMap& tmY, const AttnParams& p, cudaStream_t st, bool stats_only) {
    static unsigned long long attr_once = 0;
    if (first_use_on_device(attr_once)) {
        XL_CUDA(cudaFuncSetAttribute(attn_tc2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2Smem));
        XL_CUDA(cudaFuncSetAttribute(attn_tc2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2Smem));
    }
    XL_REQUIRE(p.B > 0 && p.H > 0 && p.N > 0 && p.ntypes >= 1 && p.ntypes <= 3, "attn_scores: bad shape");
    XL_REQUIRE(p.m && (p.out || p.out_split || stats_only), "attn_scores: missing buffers");
    const int nblk = (p.N + 127) / 128, npr = (nblk + 1) / 2;
    if (int e = launch_pair<0>(tmQ, tmY, tmQ, p, p.B * p.ntypes * p.H * npr, st)) return e;
    if (int e = check_launch("attn_tc2_kernel<stats>")) return e;
    if (stats_only) return 0;
    CUtensorMap tmO;
    if (p.out_split) {
        XL_REQUIRE(p.np % 64 == 0 && p.np >= p.N, "attn_scores: bad split pitch");
        const uint64_t dims[3] = {(uint64_t)2 * p.np, (uint64_t)p.N, (uint64_t)p.B};
        const uint64_t strides[2] = {(uint64_t)2 * p.np * 2, (uint64_t)2 * p.np * 2 * p.N};
        const uint32_t box[3] = {16, 32, 1};
        if (int e = encode_tensor_map(&tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, p.out_split, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return e;
    } else if (int e = make_map_store(&tmO, p.out, p.B, p.N)) return e;
    if (int e = launch_pair<1>(tmQ, tmY, tmO, p, p.B * npr * nblk, st)) return e;
    return check_launch("attn_tc2_kernel<map>");
}

}  // namespace xl
