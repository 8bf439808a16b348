// CTA-pair (cta_group::2) variant of the tcgen05 jet kernel for H = 256: forward jet, value / grid query, reverse jet.
//
// Same mathematics, operand formats and saved layouts as jet_tc.cu (see there for the reference citations and the
// transposed formulation).  What changes is how a tile is mapped to the machine:
//
//   * two CTAs of a cluster (one TPC) work on the SAME tile: CTA r owns units [128 r, 128 r + 128).  One
//     tcgen05.mma.cta_group::2 (M = 256, N = 128) per K step covers all 256 units; each CTA feeds its 128 weight rows (A)
//     and HALF of the activation columns (B): 64 columns x 256 k x (hi, lo) = 64 KB instead of 128 KB per tile, and 6 KB
//     instead of 8 KB of shared-memory operand reads per 128x128x16 of math;
//   * the halved operand buffer makes room for TWO tile contexts per pair, so the MMA of one tile runs while the
//     epilogue warps of both CTAs work on the other tile: the MMA / epilogue serialisation of the single-CTA kernel
//     (profiles/r01_full_tc_kernels.md) goes away;
//   * a thread produces the next layer's operand row k = its unit for all 128 columns: the 64 columns that live in
//     the other CTA are written through distributed shared memory (st.shared::cluster), published with
//     fence.proxy.async + a cluster-scope release arrive on the leader's mbarrier;
//   * weights: each CTA's TMA loads its own 128 rows, both loads complete on the leader's `w_full` barrier; the leader's
//     MMA thread commits with a multicast arrive to `w_empty` / `acc_ready` of both CTAs;
//   * the last layer's per-unit partial sums of both CTAs meet in the leader's shared memory (DSMEM stores + mbarrier),
//     the leader applies the output head.
//   * the rows a CTA produces for the PEER's column half are staged locally and moved by one DSMEM bulk copy per part
//     (cp.async.bulk.shared::cluster) issued by a dedicated exchange warp; per-thread st.shared::cluster stores of the
//     8-byte pieces ran at ~1.5 clk per piece and dominated the kernel.
// Every mechanism was validated in isolation by tests/cuda/umma2_probe.cu before this kernel was written.
//
// STATUS: correct (same parity suite as the single-CTA kernel) but SLOWER on the B200, so it is opt-in
// (DIGS_B200_PAIR=1 in the environment).  C2 off-surface cloud, bf16x2, forward without saving: 0.214 ms against 0.170 ms.
// Measured with the MMAs / the weight ring compiled out: epilogue + exchange alone 0.142 ms, with MMAs but no ring
// waits 0.169 ms.  The pair's cross-SM traffic is the limit: every tcgen05.mma.cta_group::2 pulls the peer's 2 KB half
// of B across the SM-to-SM fabric (96 KB per step and direction) on top of the 32 KB operand exchange, and the smaller
// shared-memory budget leaves 3 ring stages and one staging buffer.  See DESIGN.md section 5a.
//
// Warp roles per CTA (576 threads): warps 0-15 epilogue, warp 16 TMA producer, warp 17 MMA issuer (leader CTA) + TMEM owner.
#include "jet_tc_shared.cuh"
#include <cuda.h>
#include <type_traits>

namespace digs {
namespace tcpath {

constexpr int PSTAGES = 3;                             // weight ring stages of the pair kernel (16 KB each)
constexpr int PTHREADS = EPI_WARPS * 32 + 96;          // + TMA producer, MMA issuer, operand-exchange warp
constexpr int XWARP = EPI_WARPS + 2;                   // the exchange warp

template <int D, int LAP, int PASSES>
struct PCfg {
    static constexpr int C = 1 + D + LAP;
    static constexpr int H = 256;
    static constexpr int N = 128;                     // operand columns per tile (UMMA N); each CTA stores N/2
    static constexpr int GP = (C == 1) ? 16 : 4;      // points per epilogue group
    static constexpr int T = (N / C) / GP * GP;       // points per tile
    static constexpr int NGRP = T / GP;
    static constexpr int WQ = EPI_WARPS / 4;          // epilogue warps per TMEM lane quarter
    static constexpr int SK = 1024;                   // byte stride between 8-deep k groups (64 columns x 8 k)
    static constexpr int B_PART = (H / 8) * SK;       // 32 KB: one operand part (hi or lo) of one context
    static constexpr int NPARTS = (PASSES == 3) ? 2 : 1;
    static constexpr int B_CTX = B_PART * NPARTS;
    static constexpr int KSLABS = H / 64;
    static constexpr int ETHREADS = EPI_WARPS * 32;
    static constexpr int X_PART = 128 * 128;          // 16 KB: this CTA's 128 operand rows x the PEER's 64 columns, one part
    static constexpr int OFF_STAGE = 2 * B_CTX;       // staging for the operand exchange (X_PART * NPARTS), 1 KB aligned
    static constexpr int OFF_RING = OFF_STAGE + X_PART * NPARTS;
    static constexpr int OFF_XYZ = OFF_RING + PSTAGES * STAGE_BYTES;     // [2][N*3] floats
    static constexpr int OFF_HEAD = OFF_XYZ + 2 * N * 3 * 4;              // [2][8 unit blocks][N] floats (used in the leader)
    static constexpr int OFF_UBAR = OFF_HEAD + 2 * 8 * N * 4;             // [2][N] floats
    static constexpr int OFF_BAR = OFF_UBAR + 2 * N * 4;
    static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
    static constexpr int TMEM_COLS = 256;             // two contexts x 128 accumulator columns
    static_assert(NGRP >= WQ, "every epilogue warp needs at least one point group per step");
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

// ---- cluster / pair primitives ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void stc_f32(uint32_t caddr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(caddr), "f"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t caddr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(caddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_2d_pair(void* dst_smem, const void* tmap, uint64_t* leader_bar, int c0, int c1) {
    // issued by both CTAs into their own shared memory; the barrier address has its peer bit cleared = CTA 0's barrier
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(smem_u32(leader_bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {   // arrives on `bar` of BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

template <int D, int LAP, int PASSES, bool BWD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PTHREADS, 1)
k_tc_pair(const __grid_constant__ CUtensorMap tmapW, Args a) {
    using G = PCfg<D, LAP, PASSES>;
    constexpr int C = G::C, T = G::T, N = G::N, H = G::H, GP = G::GP;
    constexpr int P2 = (C * GP <= 4) ? 4 : ((C * GP <= 16) ? 16 : 32);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sRing = smem + G::OFF_RING;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G::OFF_BAR);
    uint64_t* w_full = bars;                          // [PSTAGES]  leader: both CTAs' weight tiles landed
    uint64_t* w_empty = bars + PSTAGES;                // [PSTAGES]  each CTA: MMAs reading the stage completed
    uint64_t* b_ready = bars + 2 * PSTAGES;            // [2]       leader: operand rows of context c written in both CTAs (3 arrivals)
    uint64_t* acc_ready = bars + 2 * PSTAGES + 2;      // [2]       each CTA: accumulators of context c complete
    uint64_t* head_ready = bars + 2 * PSTAGES + 4;     // [2]       leader: last-layer partial sums of both CTAs stored
    uint64_t* in_ready = bars + 2 * PSTAGES + 6;       // [2]       each CTA: the peer's half of the operand rows landed here (bulk copy bytes)
    uint64_t* stage_free = bars + 2 * PSTAGES + 8;     //           each CTA: the exchange copy has finished reading the staging buffer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * PSTAGES + 9);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_rank();
    const bool leader = (rank == 0);
    if (tid == 0) {
        for (int i = 0; i < PSTAGES; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&b_ready[i], 3);                  // both exchange warps (own halves written) + the peer's relay (its copy landed)
            mbar_init(&acc_ready[i], 1);
            mbar_init(&head_ready[i], 2 * EPI_WARPS);
            mbar_init(&in_ready[i], 1);
        }
        mbar_init(stage_free, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmapW);
    }
    if (warp == EPI_WARPS + 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(G::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < G::OFF_RING / 16; i += PTHREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // both CTAs: barriers initialised, operand buffers zeroed, before any remote store / arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int64_t num_tiles = (a.n_pad + T - 1) / T;
    const int64_t num_rounds = (num_tiles + 1) / 2;          // a round = two consecutive tiles = the pair's two contexts
    const int64_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int Lh = a.Lh;

    if (warp == EPI_WARPS) {
        // ================= TMA producer (both CTAs): this CTA's 128 weight rows, in the order the MMA consumes them =========
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t r = pair; r < num_rounds; r += npairs)
                for (int j = 0; j < Lh; ++j) {
                    const int l = BWD ? (Lh - j) : (j + 1);
                    for (int ctx = 0; ctx < 2; ++ctx) {
                        if (2 * r + ctx >= num_tiles) continue;
                        for (int ks = 0; ks < G::KSLABS; ++ks)
                            for (int part = 0; part < G::NPARTS; ++part) {
                                mbar_wait(&w_empty[stage], phase ^ 1);
                                if (leader) mbar_arrive_expect_tx(&w_full[stage], 2 * STAGE_BYTES);
                                // rows: [dir = fwd|transposed][part = hi|lo][layer][H]
                                tma_load_2d_pair(sRing + stage * STAGE_BYTES, &tmapW, &w_full[stage], ks * 64,
                                                 (((BWD ? 2 : 0) + part) * Lh + (l - 1)) * H + (int)rank * 128);
                                if (++stage == PSTAGES) { stage = 0; phase ^= 1; }
                            }
                    }
                }
        }
    } else if (warp == EPI_WARPS + 1) {
        // ================= MMA issuer (leader CTA only): alternates between the two tile contexts ============================
        if (leader && lane == 0) {
            const uint32_t idesc = make_idesc_bf16(256, N, /*A K-major*/ 0, /*B MN-major*/ 1);
            int stage = 0;
            uint32_t phase = 0, pb[2] = {0, 0};
            for (int64_t r = pair; r < num_rounds; r += npairs)
                for (int j = 0; j < Lh; ++j)
                    for (int ctx = 0; ctx < 2; ++ctx) {
                        if (2 * r + ctx >= num_tiles) continue;
                        const uint32_t sB_hi = smem_u32(smem) + ctx * G::B_CTX, sB_lo = sB_hi + G::B_PART;
                        const uint32_t d_tmem = tmem_base + ctx * N;
                        mbar_wait_cluster(&b_ready[ctx], pb[ctx]);
                        mbar_wait(&in_ready[ctx], pb[ctx]);        // the peer's rows for this CTA's 64 columns
                        pb[ctx] ^= 1;
                        tc_fence_after();
                        for (int ks = 0; ks < G::KSLABS; ++ks) {
                            mbar_wait(&w_full[stage], phase);   // A_hi tiles of both CTAs: hi*hi (+ hi*lo)
                            tc_fence_after();
                            {
                                const uint32_t sA = smem_u32(sRing + stage * STAGE_BYTES);
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk) {
                                    uint64_t da = make_smem_desc(sA + kk * 32, 16, 1024, LAYOUT_SW128);
                                    uint32_t koff = (ks * 8 + kk * 2) * G::SK;
                                    uint64_t db = make_smem_desc(sB_hi + koff, G::SK, G::SK, LAYOUT_SW128);
                                    umma_bf16_pair(d_tmem, da, db, idesc, (ks | kk) != 0);
                                    if constexpr (PASSES == 3) {
                                        uint64_t dbl = make_smem_desc(sB_lo + koff, G::SK, G::SK, LAYOUT_SW128);
                                        umma_bf16_pair(d_tmem, da, dbl, idesc, 1);
                                    }
                                }
                            }
                            umma_commit_pair(&w_empty[stage]);
                            if (++stage == PSTAGES) { stage = 0; phase ^= 1; }
                            if constexpr (PASSES == 3) {
                                mbar_wait(&w_full[stage], phase);   // A_lo tiles: lo*hi
                                tc_fence_after();
                                const uint32_t sA = smem_u32(sRing + stage * STAGE_BYTES);
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk) {
                                    uint64_t da = make_smem_desc(sA + kk * 32, 16, 1024, LAYOUT_SW128);
                                    uint32_t koff = (ks * 8 + kk * 2) * G::SK;
                                    uint64_t db = make_smem_desc(sB_hi + koff, G::SK, G::SK, LAYOUT_SW128);
                                    umma_bf16_pair(d_tmem, da, db, idesc, 1);
                                }
                                umma_commit_pair(&w_empty[stage]);
                                if (++stage == PSTAGES) { stage = 0; phase ^= 1; }
                            }
                        }
                        umma_commit_pair(&acc_ready[ctx]);
                    }
        }
    } else if (warp == XWARP) {
        // ================= operand exchange (both CTAs) ===========================================================================
        // Per producing (step, context): once every epilogue warp has arrived on named barrier 2 (its rows of the own column
        // half are in this CTA's operand buffer, its rows of the peer's half in the staging buffer), ONE bulk copy per part
        // moves the staged rows into the peer's operand buffer -- contiguous there, because a CTA's 128 k rows are
        // contiguous in the MN-major layout.  (Writing them with st.shared::cluster from the epilogue threads was measured
        // at ~1.5 clk per 8-byte piece: 60 us of a 245 us forward.)
        const uint32_t peer = rank ^ 1u;
        const uint32_t sStage = smem_u32(smem + G::OFF_STAGE);
        const uint32_t peer_in_ready = map_to_rank(smem_u32(in_ready), peer);
        const uint32_t b_ready_leader = map_to_rank(smem_u32(b_ready), 0);
        uint32_t pin[2] = {0, 0};
        for (int64_t r = pair; r < num_rounds; r += npairs)
            for (int j = 0; j < Lh; ++j)
                for (int ctx = 0; ctx < 2; ++ctx) {
                    if (2 * r + ctx >= num_tiles) continue;
                    asm volatile("bar.sync 2, %0;" ::"r"(G::ETHREADS + 32) : "memory");
                    if (lane == 0) {
                        const uint32_t dst = map_to_rank(smem_u32(smem) + ctx * G::B_CTX + rank * G::X_PART, peer);
                        const uint32_t mb = peer_in_ready + ctx * 8;
                        asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(mb),
                                     "r"(G::X_PART * G::NPARTS) : "memory");
#pragma unroll
                        for (int part = 0; part < G::NPARTS; ++part)
                            asm volatile(
                                "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                ::"r"(dst + part * G::B_PART), "r"(sStage + part * G::X_PART), "r"(G::X_PART), "r"(mb) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        mbar_arrive_cluster(b_ready_leader + ctx * 8);          // this CTA's own column half is complete
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        mbar_arrive(stage_free);                                // epilogue may overwrite the staging buffer
                        if (!leader) {                                          // relay: the leader's rows have landed HERE
                            mbar_wait(&in_ready[ctx], pin[ctx]);
                            pin[ctx] ^= 1;
                            mbar_arrive_cluster(b_ready_leader + ctx * 8);
                        }
                    }
                    __syncwarp();
                }
    } else {
        // ================= epilogue warps (both CTAs): this CTA's 128 units x all 128 columns of both contexts ================
        const int q = warp & 3;                     // TMEM lane quarter (hardware restriction: warp % 4)
        const int wr = warp >> 2;                   // rank among the WQ warps of the quarter
        const int u = (int)rank * 128 + q * 32 + lane;          // unit index = k row of the next operand
        const int ub = (int)rank * 4 + q;                       // 32-unit block
        const uint32_t sB = smem_u32(smem);
        float* sXyz = reinterpret_cast<float*>(smem + G::OFF_XYZ);
        float* sHead = reinterpret_cast<float*>(smem + G::OFF_HEAD);
        float* sUbar = reinterpret_cast<float*>(smem + G::OFF_UBAR);
        const uint32_t sHead_leader = map_to_rank(smem_u32(sHead), 0);
        const uint32_t head_ready_leader = map_to_rank(smem_u32(head_ready), 0);
        uint32_t pa[2] = {0, 0}, ph_head[2] = {0, 0}, pfree = 0;
        bool staged_once = false;                     // the staging buffer is free until the first exchange has been issued
        const uint32_t rowoff = (uint32_t)(u >> 3) * G::SK + (uint32_t)(u & 7) * 128;
        const int ul = q * 32 + lane;                 // row within this CTA's 128 rows = row of the staging buffer
        const uint32_t srow = smem_u32(smem + G::OFF_STAGE) + (uint32_t)(ul >> 3) * G::SK + (uint32_t)(ul & 7) * 128;
        const uint32_t usw = (uint32_t)(u & 7) << 4;
        const int64_t plane_stride = a.n_pad * (int64_t)H;
        uint32_t* words = reinterpret_cast<uint32_t*>(BWD ? a.GP_words : a.actP);
        const float4 w0v = __ldg(reinterpret_cast<const float4*>(a.w0_30) + u);
        const float wl_u = __ldg(a.w_last + u);

        for (int64_t r = pair; r < num_rounds; r += npairs) {
            const int nctx = (2 * r + 1 < num_tiles) ? 2 : 1;
            // ---- stage the coordinates (and head adjoints) of both tiles -------------------------------------------------
            for (int e = tid; e < nctx * T; e += G::ETHREADS) {
                const int ctx = e / T, ee = e % T;
                int64_t gp = (2 * r + ctx) * T + ee;
                if (gp >= a.n) gp = a.n - 1;
                float x, y, z;
                point_coords(a.src, gp, x, y, z);
                float* d = sXyz + ctx * N * 3 + ee * 3;
                d[0] = x; d[1] = y; d[2] = z;
            }
            if constexpr (BWD) {
                for (int e = tid; e < nctx * T * C; e += G::ETHREADS) {
                    const int ctx = e / (T * C), ee = e % (T * C);
                    int64_t gp = (2 * r + ctx) * T + ee / C;
                    sUbar[ctx * N + ee] = (gp < a.n) ? __ldg(a.ubar + gp * C + (ee % C)) : 0.f;
                }
            }
            named_bar_sync(1, G::ETHREADS);

            for (int j = 0; j <= Lh; ++j) {
                const int l = BWD ? (Lh - j) : j;          // layer whose (pre-)activations this step touches
                const bool last = (j == Lh);
                for (int ctx = 0; ctx < nctx; ++ctx) {
                    const int64_t p0 = (2 * r + ctx) * T;
                    const bool full = (p0 + T <= a.n);
                    if (j > 0) {
                        mbar_wait(&acc_ready[ctx], pa[ctx]);
                        pa[ctx] ^= 1;
                        tc_fence_after();
                    }
                    const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + ctx * N;
                    const float* cXyz = sXyz + ctx * N * 3;
                    const float* cUbar = sUbar + ctx * N;
                    const uint32_t cB = sB + ctx * G::B_CTX;
                    float* pl_l = a.planes ? a.planes + (int64_t)(l - 1) * C * plane_stride + p0 * H : nullptr;
                    uint32_t* wd_l = words ? words + (int64_t)(BWD ? (l - 1) : l) * C * plane_stride + p0 * H : nullptr;
                    const float blv = (!BWD && j > 0) ? __ldg(a.bias30 + (l - 1) * H + u) : 0.f;
                    float bsum = 0.f, wlsum = 0.f, w0s0 = 0.f, w0s1 = 0.f, w0s2 = 0.f;
                    const bool produces = !((!BWD && last) || (BWD && l == 0));
                    if (produces) {                    // the previous exchange copy must have drained the staging buffer
                        if (staged_once) { mbar_wait(stage_free, pfree); pfree ^= 1; }
                        staged_once = true;
                    }

                    auto run_items = [&](auto full_tag) {
                        constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll 1
                        for (int g = (wr + 2 * ctx) % G::WQ; g < G::NGRP; g += G::WQ) {   // rotated per context: 6 groups balance over 4 warps
                            const int64_t pg = p0 + g * GP;              // first point of the group
                            const int goff = g * GP * H + u;             // word offset ([col][unit] words)
                            const int poff = g * GP * H + u * 4;         // plane offset ([point / 4][unit][point % 4] floats)
                            float acc[C][GP];                             // TMEM values: fwd pre-activations, bwd incoming adjoint
                            if (j > 0) {
#pragma unroll
                                for (int c = 0; c < C; ++c) {
                                    if constexpr (GP == 16) tmem_ld16(t_base + c * T + g * GP, acc[c]);
                                    else tmem_ld4(t_base + c * T + g * GP, acc[c]);
                                }
                                tmem_ld_wait();
                            }
                            float pre[C][GP];                             // scaled pre-activations (30 z | 30 Z_k | 30 Q) of layer l
                            if ((!BWD && j == 0) || (BWD && l == 0)) {
#pragma unroll
                                for (int i = 0; i < GP; ++i) {
                                    const float* xp = cXyz + (g * GP + i) * 3;
                                    float bz = w0v.w;
                                    if (a.src.shape_bias) {
                                        int64_t gp = pg + i < a.n ? pg + i : a.n - 1;
                                        bz = 30.f * __ldg(a.src.shape_bias + (gp / a.src.pps) * H + u);
                                    }
                                    pre[0][i] = fmaf(w0v.z, xp[2], fmaf(w0v.y, xp[1], fmaf(w0v.x, xp[0], bz)));
                                    if constexpr (D > 0) {
                                        pre[1][i] = w0v.x;
                                        if constexpr (D > 1) pre[2][i] = w0v.y;
                                        if constexpr (D > 2) pre[3][i] = w0v.z;
                                        if constexpr (LAP) pre[1 + D][i] = 0.f;
                                    }
                                }
                            } else if (!BWD) {
#pragma unroll
                                for (int c = 0; c < C; ++c)
#pragma unroll
                                    for (int i = 0; i < GP; ++i) pre[c][i] = (c == 0) ? acc[c][i] + blv : acc[c][i];
                            } else if constexpr (GP % 4 == 0) {
                                const float* srcp = pl_l + poff;
#pragma unroll
                                for (int c = 0; c < C; ++c)
#pragma unroll
                                    for (int h = 0; h < GP / 4; ++h) {
                                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                                        if (FULL || pg + 4 * h < a.n_pad)
                                            v = __ldg(reinterpret_cast<const float4*>(srcp + c * plane_stride + h * 4 * H));
                                        pre[c][4 * h] = v.x; pre[c][4 * h + 1] = v.y; pre[c][4 * h + 2] = v.z; pre[c][4 * h + 3] = v.w;
                                    }
                            }
                            if (BWD && j == 0) {
#pragma unroll
                                for (int c = 0; c < C; ++c)
#pragma unroll
                                    for (int i = 0; i < GP; ++i) acc[c][i] = cUbar[(g * GP + i) * C + c] * wl_u;
                            }
                            // ---------------- per-point jet math --------------------------------------------------------
                            float out[C][GP];  // fwd: layer outputs (h | A_k | P); bwd: scaled adjoints (zhat | Zhat_k | Qhat)
#pragma unroll
                            for (int i = 0; i < GP; ++i) {
                                float s, c;
                                sincos_reduced(pre[0][i], s, c);
                                float S2 = 0.f;
#pragma unroll
                                for (int k = 0; k < D; ++k) S2 = fmaf(pre[1 + k][i], pre[1 + k][i], S2);
                                const float Qp = LAP ? pre[C - 1][i] : 0.f;
                                if (!BWD) {
                                    out[0][i] = s;
#pragma unroll
                                    for (int k = 0; k < D; ++k) out[1 + k][i] = c * pre[1 + k][i];
                                    if constexpr (LAP) out[C - 1][i] = fmaf(c, Qp, -s * S2);
                                } else {
                                    const float Pb = LAP ? acc[C - 1][i] : 0.f;
                                    float ZA = 0.f;
#pragma unroll
                                    for (int k = 0; k < D; ++k) ZA = fmaf(pre[1 + k][i], acc[1 + k][i], ZA);
                                    if (j == 0) {
                                        // dL/dw_last: ubar . (h | A_k | P) of the last sine layer
                                        const float* ubp = cUbar + (g * GP + i) * C;
                                        float t = ubp[0] * s;
#pragma unroll
                                        for (int k = 0; k < D; ++k) t = fmaf(ubp[1 + k], c * pre[1 + k][i], t);
                                        if constexpr (LAP) t = fmaf(ubp[C - 1], fmaf(c, Qp, -s * S2), t);
                                        wlsum += t;
                                    }
                                    float zh = c * acc[0][i] - s * ZA;
                                    if constexpr (LAP) zh -= Pb * fmaf(s, Qp, c * S2);
                                    out[0][i] = zh;
#pragma unroll
                                    for (int k = 0; k < D; ++k) {
                                        float v = c * acc[1 + k][i];
                                        if constexpr (LAP) v -= 2.f * s * pre[1 + k][i] * Pb;
                                        out[1 + k][i] = v;
                                    }
                                    if constexpr (LAP) out[C - 1][i] = c * Pb;
                                }
                            }
                            // ---------------- scatter -------------------------------------------------------------------
                            if (BWD && l == 0) {
                                // layer-0 parameter gradients: dW0[u][k] += 30 (zhat x_k + Zhat_k), db0[u] += 30 zhat
#pragma unroll
                                for (int i = 0; i < GP; ++i) {
                                    const float* xp = cXyz + (g * GP + i) * 3;
                                    bsum += out[0][i];
                                    if constexpr (D > 0) w0s0 += fmaf(out[0][i], xp[0], out[1][i]);
                                    if constexpr (D > 1) w0s1 += fmaf(out[0][i], xp[1], out[2][i]);
                                    if constexpr (D > 2) w0s2 += fmaf(out[0][i], xp[2], out[3][i]);
                                    if (a.sb_grad && pg + i < a.n)
                                        atomicAdd(a.sb_grad + ((pg + i) / a.src.pps) * H + u, 30.f * out[0][i]);
                                }
                            } else if (!BWD && last) {
                                // last layer: this warp's 32-unit partial sums go to the LEADER's shared memory
                                float part[P2];
#pragma unroll
                                for (int v = 0; v < P2; ++v) part[v] = (v < C * GP) ? out[v / GP][v % GP] * wl_u : 0.f;
                                warp_reduce_scatter<P2>(part, lane);
                                if (lane < C * GP)
                                    stc_f32(sHead_leader + (uint32_t)(((ctx * 8 + ub) * N + (lane / GP) * T + g * GP + (lane % GP)) * 4), part[0]);
                                if (pl_l) {        // the last sine layer's pre-activations are needed by the reverse pass too
#pragma unroll
                                    for (int c = 0; c < C; ++c) {
                                        float* dst = pl_l + c * plane_stride + poff;
#pragma unroll
                                        for (int h = 0; h < GP / 4; ++h)
                                            if (FULL || pg + 4 * h < a.n_pad)
                                                *reinterpret_cast<float4*>(dst + h * 4 * H) =
                                                    make_float4(pre[c][4 * h], pre[c][4 * h + 1], pre[c][4 * h + 2], pre[c][4 * h + 3]);
                                    }
                                }
                            } else {
                                if (BWD) {
#pragma unroll
                                    for (int i = 0; i < GP; ++i) bsum += out[0][i];
                                }
                                // (1) next operand: row k = u; columns of the other half live in the peer CTA (DSMEM)
                                uint32_t hiw[C][GP / 2], low[C][GP / 2];
#pragma unroll
                                for (int c = 0; c < C; ++c) {
                                    const int ncol = c * T + g * GP;       // a group never straddles the 64-column halves
#pragma unroll
                                    for (int i = 0; i < GP / 2; ++i) split2(out[c][2 * i], out[c][2 * i + 1], hiw[c][i], low[c][i]);
                                    const uint32_t half = (uint32_t)ncol >> 6;
                                    // own column half -> this CTA's operand buffer; the peer's half -> staging (same row layout)
                                    const uint32_t rbase = (half == rank) ? (cB + rowoff) : srow;
                                    const uint32_t pstep = (half == rank) ? (uint32_t)G::B_PART : (uint32_t)G::X_PART;
                                    if constexpr (GP >= 8) {
#pragma unroll
                                        for (int h = 0; h < GP / 8; ++h) {       // one 16-byte chunk per 8 points
                                            const int nc = (ncol & 63) + 8 * h;
                                            const uint32_t off = rbase + ((((uint32_t)nc >> 3) << 4) ^ usw);
                                            sts_v4(off, hiw[c][4 * h], hiw[c][4 * h + 1], hiw[c][4 * h + 2], hiw[c][4 * h + 3]);
                                            if constexpr (PASSES == 3)
                                                sts_v4(off + pstep, low[c][4 * h], low[c][4 * h + 1], low[c][4 * h + 2], low[c][4 * h + 3]);
                                        }
                                    } else {
                                        const uint32_t nc = (uint32_t)ncol & 63;
                                        const uint32_t off = rbase + (((nc >> 3) << 4) ^ usw) + (nc & 4) * 2;
                                        sts_v2(off, hiw[c][0], hiw[c][1]);
                                        if constexpr (PASSES == 3) sts_v2(off + pstep, low[c][0], low[c][1]);
                                    }
                                }
                                // (2) this warp's last group of the step: publish to the leader's MMA warp before the global stores
                                // (non-blocking arrive on named barrier 2; the exchange warp is its only waiter)
                                if (g + G::WQ >= G::NGRP) {
                                    fence_proxy_async();
                                    tc_fence_before();
                                    asm volatile("bar.arrive 2, %0;" ::"r"(G::ETHREADS + 32) : "memory");
                                }
                                // (3) what the reverse pass / wgrad need, to HBM
                                if (!BWD && j > 0 && pl_l) {   // layer 0 is recomputed from the coordinates, never stored
#pragma unroll
                                    for (int c = 0; c < C; ++c) {
                                        float* dst = pl_l + c * plane_stride + poff;
#pragma unroll
                                        for (int h = 0; h < GP / 4; ++h)
                                            if (FULL || pg + 4 * h < a.n_pad)
                                                *reinterpret_cast<float4*>(dst + h * 4 * H) =
                                                    make_float4(pre[c][4 * h], pre[c][4 * h + 1], pre[c][4 * h + 2], pre[c][4 * h + 3]);
                                    }
                                }
                                if (wd_l) {
                                    // wgrad operand words (hi | lo << 16); forward tail points (>= n) must be exact zeros
#pragma unroll
                                    for (int c = 0; c < C; ++c) {
                                        uint32_t* dst = wd_l + c * plane_stride + goff;
#pragma unroll
                                        for (int i = 0; i < GP; ++i) {
                                            uint32_t w = __byte_perm(hiw[c][i >> 1], low[c][i >> 1], (i & 1) ? 0x7632 : 0x5410);
                                            if constexpr (FULL) {
                                                dst[i * H] = w;
                                            } else {
                                                if (!BWD && pg + i >= a.n) w = 0;
                                                if (pg + i < a.n_pad) dst[i * H] = w;
                                            }
                                        }
                                    }
                                }
                            }
                        }
                    };
                    if (full) run_items(std::true_type{});
                    else run_items(std::false_type{});

                    // ---------------- end of the (context, step): per-unit parameter gradients --------------------------------
                    if (BWD) {
                        if (j == 0) atomicAdd(a.dw_last + u, wlsum);
                        if (l >= 1) {
                            atomicAdd(a.db[l] + u, 30.f * bsum);
                        } else {
                            if (!a.sb_grad) atomicAdd(a.db[0] + u, 30.f * bsum);
                            if constexpr (D > 0) atomicAdd(a.dW0 + (int64_t)u * a.w0_stride + 0, 30.f * w0s0);
                            if constexpr (D > 1) atomicAdd(a.dW0 + (int64_t)u * a.w0_stride + 1, 30.f * w0s1);
                            if constexpr (D > 2) atomicAdd(a.dW0 + (int64_t)u * a.w0_stride + 2, 30.f * w0s2);
                        }
                    }
                    if (!BWD && last) {
                        // ---- output head: both CTAs' partial sums are in the leader's shared memory ------------------------
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(head_ready_leader + ctx * 8);
                        if (leader) {
                            mbar_wait_cluster(&head_ready[ctx], ph_head[ctx]);
                            ph_head[ctx] ^= 1;
                            const float* cHead = sHead + ctx * 8 * N;
                            for (int e = tid; e < T; e += G::ETHREADS) {
                                int64_t gp = p0 + e;
                                if (gp < a.n) {
                                    float v[C];
#pragma unroll
                                    for (int c = 0; c < C; ++c) {
                                        float s = 0.f;
#pragma unroll
                                        for (int w = 0; w < 8; ++w) s += cHead[w * N + c * T + e];
                                        v[c] = s;
                                    }
                                    float uu = v[0] + __ldg(a.b_last);
                                    if (a.head_saved) {
                                        a.head_saved[gp * C] = uu;
#pragma unroll
                                        for (int c = 1; c < C; ++c) a.head_saved[gp * C + c] = v[c];
                                    }
                                    if (a.head == DIGS_HEAD_IDENTITY) {
                                        a.f[gp] = uu;
                                        if constexpr (D > 0) {
#pragma unroll
                                            for (int k = 0; k < D; ++k) a.grad[gp * D + k] = v[1 + k];
                                            if constexpr (LAP) a.lap[gp] = v[1 + D];
                                        }
                                    } else {
                                        float aa = fabsf(uu) + 1e-8f;
                                        float sg = (uu > 0.f) ? 1.f : ((uu < 0.f) ? -1.f : 0.f);
                                        float ra = sqrtf(aa);
                                        a.f[gp] = a.scale * (sg * ra - a.radius);
                                        if constexpr (D > 0) {
                                            float g1 = 0.5f / ra, gg = 0.f;
#pragma unroll
                                            for (int k = 0; k < D; ++k) {
                                                a.grad[gp * D + k] = a.scale * g1 * v[1 + k];
                                                gg = fmaf(v[1 + k], v[1 + k], gg);
                                            }
                                            if constexpr (LAP) {
                                                float g2 = -0.25f * sg / (aa * ra);
                                                a.lap[gp] = a.scale * (g1 * v[1 + D] + g2 * gg);
                                            }
                                        }
                                    }
                                }
                            }
                        }
                    }
                }
            }
            named_bar_sync(1, G::ETHREADS);   // sXyz / sUbar are rewritten by the next round
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // no CTA of the pair exits (or frees TMEM) while the other may still address it
    if (warp == EPI_WARPS + 1)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(G::TMEM_COLS) : "memory");
}

// ---- host -----------------------------------------------------------------------------------------------------------------
bool pair_supported(const DigsNet* net) { return net->H == 256; }

template <int D, int LAP, int PASSES, bool BWD>
static int launch_one(const CUtensorMap& tmap, const Args& a, cudaStream_t st) {
    using G = PCfg<D, LAP, PASSES>;
    auto kern = k_tc_pair<D, LAP, PASSES, BWD>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES);
        if (e != cudaSuccess) { set_error("tc pair: smem attribute (%d B): %s", G::SMEM_BYTES, cudaGetErrorString(e)); return DIGS_ECUDA; }
        attr_set = true;
    }
    static int max_pairs = 0;     // CTA pairs that can be resident at once (persistent kernel: never launch more)
    if (!max_pairs) {
        cudaLaunchConfig_t qc = {};
        qc.gridDim = dim3(2 * (unsigned)num_sms());
        qc.blockDim = dim3(PTHREADS);
        qc.dynamicSmemBytes = G::SMEM_BYTES;
        int nc = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, kern, &qc);
        if (e != cudaSuccess || nc < 1) { cudaGetLastError(); nc = num_sms() / 2; }
        max_pairs = nc;
    }
    const int64_t tiles = (a.n_pad + G::T - 1) / G::T;
    const int64_t rounds = (tiles + 1) / 2;
    const int grid = 2 * (int)(rounds < max_pairs ? rounds : max_pairs);
    kern<<<grid, PTHREADS, G::SMEM_BYTES, st>>>(tmap, a);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("tc pair launch: %s", cudaGetErrorString(e)); return DIGS_ECUDA; }
    return DIGS_OK;
}

template <int D, int LAP, int PASSES>
static int launch_dir(const CUtensorMap& tmap, const Args& a, bool bwd, cudaStream_t st) {
    if constexpr (D == 0) {
        if (bwd) { set_error("tc reverse needs jet channels"); return DIGS_EINVAL; }
        return launch_one<D, LAP, PASSES, false>(tmap, a, st);
    } else {
        return bwd ? launch_one<D, LAP, PASSES, true>(tmap, a, st) : launch_one<D, LAP, PASSES, false>(tmap, a, st);
    }
}

int launch_pair(const DigsNet* net, const CUtensorMap& tmap, const Args& a, int want_jet, int want_lap, int passes, bool bwd,
                cudaStream_t st) {
    if (!pair_supported(net)) { set_error("tc pair: H=%d not supported (256)", net->H); return DIGS_ENOTIMPL; }
    const bool p3 = (passes == 3);
    if (!want_jet) return p3 ? launch_dir<0, 0, 3>(tmap, a, bwd, st) : launch_dir<0, 0, 1>(tmap, a, bwd, st);
    if (net->d == 3 && want_lap) return p3 ? launch_dir<3, 1, 3>(tmap, a, bwd, st) : launch_dir<3, 1, 1>(tmap, a, bwd, st);
    if (net->d == 3) return p3 ? launch_dir<3, 0, 3>(tmap, a, bwd, st) : launch_dir<3, 0, 1>(tmap, a, bwd, st);
    if (want_lap) return p3 ? launch_dir<2, 1, 3>(tmap, a, bwd, st) : launch_dir<2, 1, 1>(tmap, a, bwd, st);
    return p3 ? launch_dir<2, 0, 3>(tmap, a, bwd, st) : launch_dir<2, 0, 1>(tmap, a, bwd, st);
}

}  // namespace tcpath
}  // namespace digs
