// tcgen05 / TMEM / TMA implementation of the jet path (DIGS_MODE_BF16X2, DIGS_MODE_BF16): forward jet, grid / value
// query, and the reverse pass through the network (dgrad); the weight-gradient GEMM lives in wgrad_tc.cu.
//
// Reference behaviour reproduced: FCBlock.forward (models/DiGS.py:122-134) for a whole network per point tile,
// the gradient / Laplacian channels the reference obtains from utils.gradient (utils/utils.py:108-117,
// models/losses.py:72-76,100-106), the chunk loop of utils.implicit2mesh (utils/utils.py:343-348), and the
// network part of loss.backward() (train_surface_reconstruction.py:128).
//
// One persistent CTA per SM walks point tiles.  A tile is T points x C jet channels <= 128 operand columns
// (column = channel*T + point).  Every hidden layer is evaluated TRANSPOSED on the tensor cores:
//
//   forward   D[unit][col]   = sum_k (30 W_l)[unit][k]  * X[col][k]        A = 30 W_l      (K-major,  TMA ring)
//   reverse   D[k][col]      = sum_u (30 W_l)[u][k]     * Zhat[col][u]     A = (30 W_l)^T  (K-major,  TMA ring)
//
// so units sit on TMEM lanes and the C jet channels of a point sit in TMEM columns of the same lane: the sin/cos
// + jet recurrences (forward) and their adjoints (reverse) of one (unit, point) need only one thread's registers.
//   A   bf16 hi [+ lo] tiles of 128 rows x 64 k, SWIZZLE_128B, streamed from L2 by TMA through a 5-stage mbarrier ring;
//   B   bf16 hi [+ lo], MN-major SWIZZLE_128B, written IN PLACE by the epilogue warps as the next layer's operand:
//       a thread owns one unit = one k row of B and stores 4 (value-only: 16) consecutive points per st.shared;
//   D   fp32 in TMEM (H/128 tiles of N columns, two buffers), read back with tcgen05.ld.32x32b.x4 / x16.
// DIGS_MODE_BF16X2 runs three MMA passes per product (hi*hi + hi*lo + lo*hi); DIGS_MODE_BF16 runs hi*hi only.
// Layer 0 (K = d) and the last layer (N = 1) are fp32 FMAs / warp reductions from registers.
//
// A layer step runs in two operand phases (Cfg::NPH): the epilogue first finishes the units of the lower half of the next
// layer's K range and releases them (b_ready[0]), so the MMA warp starts the next layer on those K slabs -- into the
// other TMEM accumulator buffer -- while the epilogue is still producing the upper half (b_ready[1]).
//
// What the forward leaves in `saved` for the reverse pass (per cloud, n_pad = n rounded up to 8):
//   planes   [Lh][C][n_pad/4][H][4] fp32  pre-activations scaled by 30: (30 z | 30 Z_k | 30 Q); a thread's 4 points = one float4
//   actP     [Lh][C*n_pad][H]  uint32  layer inputs (h | A_k | P) of layers 1..Lh as (bf16 hi | bf16 lo << 16) = wgrad operand
//   head     [n][C] fp32               (u | gu_k | lu)
// and the reverse kernel writes GP [Lh][C*n_pad][H] uint32 = adjoints (zbar | Zbar_k | Qbar)/30, the other wgrad operand.
// Every global access of an epilogue warp covers 32 consecutive units of one (channel, point): one 128-byte line.
//
// Warp roles (576 threads): warps 0-15 epilogue, warp 16 TMA producer, warp 17 MMA issuer + TMEM owner.
#include "jet_tc_shared.cuh"
#include <cuda.h>
#include <type_traits>
#include <cstdlib>

namespace digs {
namespace tcpath {

template <int D, int LAP, int H, int PASSES, bool BWD>
__global__ void __launch_bounds__(NTHREADS, 1) k_tc_net(const __grid_constant__ CUtensorMap tmapW, Args a) {
    using G = Cfg<D, LAP, H, PASSES>;
    constexpr int C = G::C, T = G::T, N = G::N, MT = G::MT, NPH = G::NPH;
    constexpr int GP = G::GP;
    constexpr int P2 = (C * GP <= 4) ? 4 : ((C * GP <= 16) ? 16 : 32);   // padded size of the head partial-sum vector
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1 KB aligned, still a shared-space pointer
    uint8_t* sRing = smem + G::OFF_RING;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G::OFF_BAR);
    uint64_t* w_full = bars;                          // [STAGES]
    uint64_t* w_empty = bars + STAGES;                // [STAGES]
    uint64_t* b_ready = bars + 2 * STAGES;            // [2] epilogue -> MMA: operand rows of phase p complete (ETHREADS arrivals)
    uint64_t* acc_ready = bars + 2 * STAGES + 2;      // [2] MMA -> epilogue: accumulators of the M tiles of phase p complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        mbar_init(&b_ready[0], G::ETHREADS);
        mbar_init(&b_ready[1], G::ETHREADS);
        mbar_init(&acc_ready[0], 1);
        mbar_init(&acc_ready[1], 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmapW);
    }
    if (warp == EPI_WARPS + 1) tmem_alloc<G::TMEM_COLS>(tmem_slot);
    for (int i = tid; i < G::OFF_RING / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    // Reverse pass: per-unit gradients (biases, layer 0, last layer) are summed per CTA in shared memory over all its
    // tiles and flushed once at the end -- one global atomic per (CTA, unit, tensor) instead of one per (item, unit):
    // the per-item global atomics all hit the same few hundred addresses and cost 67 us of a 300 us kernel.
    float* sAcc = reinterpret_cast<float*>(smem + G::OFF_ACC);
    const int n_acc = (BWD && a.acc_smem) ? (a.Lh + 5) * H : 0;
    for (int i = tid; i < n_acc; i += NTHREADS) sAcc[i] = 0.f;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int64_t num_tiles = (a.n_pad + T - 1) / T;   // tiles cover [0, n_pad): the wgrad operands have no unwritten column
    const int Lh = a.Lh;

    if (warp == EPI_WARPS) {
        // ================= TMA producer: stream weight tiles (hi, lo) in the order the MMA warp consumes them =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t r = blockIdx.x; r < num_tiles; r += gridDim.x)
                for (int j = 0; j < Lh; ++j) {
                    const int l = BWD ? (Lh - j) : (j + 1);
                    auto load = [&](int ks, int m) {
                        for (int part = 0; part < G::NPARTS; ++part) {
                            mbar_wait(&w_empty[stage], phase ^ 1);
                            mbar_arrive_expect_tx(&w_full[stage], STAGE_BYTES);
                            // rows: [dir = fwd|transposed][part = hi|lo][layer][H]
                            tma_load_2d(sRing + stage * STAGE_BYTES, &tmapW, &w_full[stage], ks * 64,
                                        (((BWD ? 2 : 0) + part) * Lh + (l - 1)) * H + m * 128);
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        }
                    };
                    weight_tile_order<G>(load, [](int) {}, [](int) {});
                }
        }
    } else if (warp == EPI_WARPS + 1) {
        // ================= MMA issuer ===================================================================================
        // Step j+1's accumulators go to TMEM buffer (j+1)&1; its K slabs of phase p start once b_ready[p] completes.
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, N, /*A K-major*/ 0, /*B MN-major*/ 1);
            const uint32_t sB_hi = smem_u32(smem), sB_lo = sB_hi + G::B_PART;
            int stage = 0;
            uint32_t phase = 0, pb = 0;
            for (int64_t r = blockIdx.x; r < num_tiles; r += gridDim.x)
                for (int j = 0; j < Lh; ++j) {
                    const uint32_t d_buf = tmem_base + ((j + 1) & 1) * G::ACC_COLS;
                    auto issue = [&](int ks, int m) {
                        const uint32_t d_tmem = d_buf + m * N;
                        mbar_wait(&w_full[stage], phase);   // A_hi tile: hi*hi (+ hi*lo)
                        tc_fence_after();
                        {
                            const uint32_t sA = smem_u32(sRing + stage * STAGE_BYTES);
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                uint64_t da = make_smem_desc(sA + kk * 32, 16, 1024, LAYOUT_SW128);
                                uint32_t koff = (ks * 8 + kk * 2) * G::SK;
                                uint64_t db = make_smem_desc(sB_hi + koff, G::SN, G::SK, LAYOUT_SW128);
                                umma_bf16(d_tmem, da, db, idesc, (ks | kk) != 0);
                                if constexpr (PASSES == 3) {
                                    uint64_t dbl = make_smem_desc(sB_lo + koff, G::SN, G::SK, LAYOUT_SW128);
                                    umma_bf16(d_tmem, da, dbl, idesc, 1);
                                }
                            }
                        }
                        umma_commit(&w_empty[stage]);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        if constexpr (PASSES == 3) {
                            mbar_wait(&w_full[stage], phase);   // A_lo tile: lo*hi
                            tc_fence_after();
                            const uint32_t sA = smem_u32(sRing + stage * STAGE_BYTES);
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                uint64_t da = make_smem_desc(sA + kk * 32, 16, 1024, LAYOUT_SW128);
                                uint32_t koff = (ks * 8 + kk * 2) * G::SK;
                                uint64_t db = make_smem_desc(sB_hi + koff, G::SN, G::SK, LAYOUT_SW128);
                                umma_bf16(d_tmem, da, db, idesc, 1);
                            }
                            umma_commit(&w_empty[stage]);
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        }
                    };
                    weight_tile_order<G>(issue,
                                         [&](int ph) { mbar_wait(&b_ready[ph], pb); tc_fence_after(); },
                                         [&](int ph) { umma_commit(&acc_ready[ph]); });
                    pb ^= 1;               // both phase barriers complete exactly once per operand-producing step
                }
        }
    } else {
        // ================= epilogue warps ===============================================================================
        // Warp w serves TMEM lane quarter q = w % 4 (hardware restriction) and shares that quarter's work items
        // (M tile, point group) of each phase with the other WQ - 1 warps of the quarter.
        const int q = warp & 3;
        const int wr = warp >> 2;
        const uint32_t sB = smem_u32(smem);        // operand stores use shared-space addresses (st.shared)
        float* sXyz = reinterpret_cast<float*>(smem + G::OFF_XYZ);
        float* sHead = reinterpret_cast<float*>(smem + G::OFF_HEAD);
        float* sUbar = reinterpret_cast<float*>(smem + G::OFF_UBAR);
        uint32_t pa = 0;
        const uint32_t usw = (uint32_t)(lane & 7) << 4;
        const int64_t plane_stride = a.n_pad * (int64_t)H;     // one (layer, channel) plane of `planes` / actP / GP
        uint32_t* words = reinterpret_cast<uint32_t*>(BWD ? a.GP_words : a.actP);

        // ---- stage a tile's coordinates (and head adjoints) into coordinate buffer xb --------------------------------------
        auto stage = [&](const int64_t tile, const int xb) {
            const int64_t q0 = tile * T;
            float* xs = sXyz + xb * (N * 3);
            for (int e = tid; e < T; e += G::ETHREADS) {
                int64_t gp = q0 + e;
                if (gp >= a.n) gp = a.n - 1;
                float x, y, z;
                point_coords(a.src, gp, x, y, z);
                xs[e * 3] = x; xs[e * 3 + 1] = y; xs[e * 3 + 2] = z;
            }
            if constexpr (BWD) {
                for (int e = tid; e < T * C; e += G::ETHREADS) {
                    int64_t gp = q0 + e / C;
                    sUbar[e] = (gp < a.n) ? __ldg(a.ubar + gp * C + (e % C)) : 0.f;
                }
            }
            named_bar_sync(1, G::ETHREADS);
        };

        // ---- layer 0 of a tile on its own (forward direction): pre-activations from the coordinates, activation, operand for
        // the first hidden layer.  Used to start the NEXT tile before the current tile's last-layer step (see the schedule below);
        // the same arithmetic as step j = 0 of run_tile.
        auto layer0 = [&](const int64_t tile, const int xb) {
            const int64_t q0 = tile * T;
            const float* xs = sXyz + xb * (N * 3);
            uint32_t* wd0 = words ? words + q0 * H : nullptr;       // actP slot 0
#pragma unroll 1
            for (int ph = 0; ph < NPH; ++ph) {
#pragma unroll 1
                for (int it = (wr + G::WQ - (ph * G::ROT) % G::WQ) % G::WQ; it < G::NITEMS; it += G::WQ) {
                    const int mt = ph * G::MTP + it / G::NGRP;
                    const int g = it % G::NGRP;
                    const int u = (mt * 4 + q) * 32 + lane;
                    const int64_t pg = q0 + g * GP;
                    const float4 w0v = __ldg(reinterpret_cast<const float4*>(a.w0_30) + u);
                    float out[C][GP];
#pragma unroll
                    for (int i = 0; i < GP; ++i) {
                        const float* xp = xs + (g * GP + i) * 3;
                        float bz = w0v.w;
                        if (a.src.shape_bias) {
                            int64_t gp = pg + i < a.n ? pg + i : a.n - 1;
                            bz = 30.f * __ldg(a.src.shape_bias + (gp / a.src.pps) * H + u);
                        }
                        float s, c;
                        sincos_reduced(fmaf(w0v.z, xp[2], fmaf(w0v.y, xp[1], fmaf(w0v.x, xp[0], bz))), s, c);
                        out[0][i] = s;
                        if constexpr (D > 0) {
                            out[1][i] = c * w0v.x;
                            if constexpr (D > 1) out[2][i] = c * w0v.y;
                            if constexpr (D > 2) out[3][i] = c * w0v.z;
                            float S2 = w0v.x * w0v.x;
                            if constexpr (D > 1) S2 = fmaf(w0v.y, w0v.y, S2);
                            if constexpr (D > 2) S2 = fmaf(w0v.z, w0v.z, S2);
                            if constexpr (LAP) out[C - 1][i] = -s * S2;
                        }
                    }
                    const uint32_t rowoff = (uint32_t)(u >> 3) * G::SK + (uint32_t)(u & 7) * 128;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const int ncol = c * T + g * GP;
                        uint32_t hiw[GP / 2], low[GP / 2];
#pragma unroll
                        for (int i = 0; i < GP / 2; ++i) split2(out[c][2 * i], out[c][2 * i + 1], hiw[i], low[i]);
                        if constexpr (GP >= 8) {
#pragma unroll
                            for (int h = 0; h < GP / 8; ++h) {
                                const int nc = ncol + 8 * h;
                                const uint32_t off = rowoff + (uint32_t)(nc >> 6) * G::SN + (((((uint32_t)nc & 63) >> 3) << 4) ^ usw);
                                sts_v4(sB + off, hiw[4 * h], hiw[4 * h + 1], hiw[4 * h + 2], hiw[4 * h + 3]);
                                if constexpr (PASSES == 3) sts_v4(sB + G::B_PART + off, low[4 * h], low[4 * h + 1], low[4 * h + 2], low[4 * h + 3]);
                            }
                        } else {
                            const uint32_t off = rowoff + (uint32_t)(ncol >> 6) * G::SN +
                                                 (((((uint32_t)ncol & 63) >> 3) << 4) ^ usw) + (uint32_t)(ncol & 4) * 2;
                            sts_v2(sB + off, hiw[0], hiw[1]);
                            if constexpr (PASSES == 3) sts_v2(sB + G::B_PART + off, low[0], low[1]);
                        }
                        if (wd0) {      // wgrad operand words of layer 1's input; padding points are exact zeros
                            uint32_t* dst = wd0 + c * plane_stride + g * GP * H + u;
#pragma unroll
                            for (int i = 0; i < GP; ++i) {
                                uint32_t w = __byte_perm(hiw[i >> 1], low[i >> 1], (i & 1) ? 0x7632 : 0x5410);
                                if (pg + i >= a.n) w = 0;
                                if (pg + i < a.n_pad) dst[i * H] = w;
                            }
                        }
                    }
                    if (it + G::WQ >= G::NITEMS) {
                        fence_proxy_async();
                        tc_fence_before();
                        mbar_arrive(&b_ready[ph]);
                    }
                }
            }
        };

        // Forward kernels with an even number of MMA steps per tile run layer 0 of the NEXT tile before the last-layer step of
        // the current one: that step and the head use no tensor core, so the MMA warp works on the next tile's first hidden
        // layer meanwhile (into the other TMEM buffer; the operand buffer is free once all accumulators of the tile are complete).
        const bool pipe = !BWD && ((Lh & 1) == 0);
        bool layer0_done = false;      // layer 0 of the tile about to start has been run already
        int xb = 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int64_t p0 = tile * T;
            const bool full = (p0 + T <= a.n);     // no tail handling needed anywhere in this tile
            const float* xs = sXyz + xb * (N * 3);
            const int j_start = layer0_done ? 1 : 0;
            if (!layer0_done) stage(tile, xb);
            const int64_t nxt = tile + gridDim.x;
            const bool pipe_next = pipe && nxt < num_tiles;

            auto run_tile = [&](auto full_tag) {
                constexpr bool FULL = decltype(full_tag)::value;
                for (int j = j_start; j <= Lh; ++j) {
                    const int l = BWD ? (Lh - j) : j;          // layer whose (pre-)activations this step touches
                    const bool last = (j == Lh);
                    bool awaited = false;
                    if (!BWD && last && pipe_next) {
                        // every MMA of this tile has to be complete before the next tile's layer 0 rewrites the operand
#pragma unroll 1
                        for (int ph = 0; ph < NPH; ++ph) mbar_wait(&acc_ready[ph], pa);
                        tc_fence_after();
                        stage(nxt, xb ^ 1);
                        layer0(nxt, xb ^ 1);
                        awaited = true;
                    }
                    const uint32_t t_buf = tmem_base + ((uint32_t)(q * 32) << 16) + (j & 1) * G::ACC_COLS;
                    // layer slot bases: planes slot = l-1 (both directions); GP slot = l-1 (reverse), actP slot = l (forward)
                    float* pl_l = a.planes ? a.planes + (int64_t)(l - 1) * C * plane_stride + p0 * H : nullptr;
                    uint32_t* wd_l = words ? words + (int64_t)(BWD ? (l - 1) : l) * C * plane_stride + p0 * H : nullptr;

                    // stored pre-activations of this thread's unit for the 4-point pieces of group g (float4 each).
                    // (Issuing the first item's loads before the accumulator wait was tried: the longer live range spilled
                    //  ~500 B per thread and the reverse kernel lost 40 %.)
                    auto load_planes = [&](float (&pre)[C][GP], int mt, int g) {
                        const float* srcp = pl_l + (g * GP * H + (mt * 128 + q * 32 + lane) * 4);
#pragma unroll
                        for (int c = 0; c < C; ++c)
#pragma unroll
                            for (int h = 0; h < GP / 4; ++h) {
                                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (FULL || p0 + g * GP + 4 * h < a.n_pad)
                                    v = __ldg(reinterpret_cast<const float4*>(srcp + c * plane_stride + h * 4 * H));
                                pre[c][4 * h] = v.x; pre[c][4 * h + 1] = v.y; pre[c][4 * h + 2] = v.z; pre[c][4 * h + 3] = v.w;
                            }
                    };
#pragma unroll 1
                    for (int ph = 0; ph < NPH; ++ph) {
                        if (j > 0 && !awaited) {   // the accumulators of this phase's M tiles (the MMA warp may still be filling the others)
                            mbar_wait(&acc_ready[ph], pa);
                            tc_fence_after();
                        }
#pragma unroll 1
                        for (int it = (wr + G::WQ - (ph * G::ROT) % G::WQ) % G::WQ; it < G::NITEMS; it += G::WQ) {
                            const int mt = ph * G::MTP + it / G::NGRP;
                            const int g = it % G::NGRP;
                            const int ub = mt * 4 + q;                   // 32-unit block
                            const int u = ub * 32 + lane;                // unit index
                            const int64_t pg = p0 + g * GP;              // first point of the group
                            const int goff = g * GP * H + u;             // word offset from the tile / layer base ([col][unit] words)
                            const int poff = g * GP * H + u * 4;         // plane offset ([point / 4][unit][point % 4] floats)
                            DIGS_ASSERT(g < G::NGRP && mt < MT && u < H && (FULL || p0 < a.n_pad));
                            const uint32_t t_base = t_buf + mt * N;
                            float bsum = 0.f, wlsum = 0.f, w0s0 = 0.f, w0s1 = 0.f, w0s2 = 0.f;
                            float acc[C][GP];                             // TMEM values: fwd pre-activations, bwd incoming adjoint
                            if (j > 0) {
#pragma unroll
                                for (int c = 0; c < C; ++c) {
                                    if constexpr (GP == 16) tmem_ld16(t_base + c * T + g * GP, acc[c]);
                                    else if constexpr (GP == 8) tmem_ld8(t_base + c * T + g * GP, acc[c]);
                                    else tmem_ld4(t_base + c * T + g * GP, acc[c]);
                                }
                                tmem_ld_wait();
                            }
                            float pre[C][GP];                             // scaled pre-activations (30 z | 30 Z_k | 30 Q) of layer l
                            if ((!BWD && j == 0) || (BWD && l == 0)) {
                                const float4 w0v = __ldg(reinterpret_cast<const float4*>(a.w0_30) + u);
#pragma unroll
                                for (int i = 0; i < GP; ++i) {
                                    const float* xp = xs + (g * GP + i) * 3;
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
                                const float blv = __ldg(a.bias30 + (l - 1) * H + u);
#pragma unroll
                                for (int c = 0; c < C; ++c)
#pragma unroll
                                    for (int i = 0; i < GP; ++i) pre[c][i] = (c == 0) ? acc[c][i] + blv : acc[c][i];
                            } else if constexpr (GP % 4 == 0) {
                                load_planes(pre, mt, g);
                            }
                            float wl_u = 0.f;
                            if ((BWD && j == 0) || (!BWD && last)) wl_u = __ldg(a.w_last + u);
                            if (BWD && j == 0) {
#pragma unroll
                                for (int c = 0; c < C; ++c)
#pragma unroll
                                    for (int i = 0; i < GP; ++i) acc[c][i] = sUbar[(g * GP + i) * C + c] * wl_u;
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
                                        const float* ubp = sUbar + (g * GP + i) * C;
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
                                    const float* xp = xs + (g * GP + i) * 3;
                                    bsum += out[0][i];
                                    if constexpr (D > 0) w0s0 += fmaf(out[0][i], xp[0], out[1][i]);
                                    if constexpr (D > 1) w0s1 += fmaf(out[0][i], xp[1], out[2][i]);
                                    if constexpr (D > 2) w0s2 += fmaf(out[0][i], xp[2], out[3][i]);
                                    if (a.sb_grad && pg + i < a.n)
                                        atomicAdd(a.sb_grad + ((pg + i) / a.src.pps) * H + u, 30.f * out[0][i]);
                                }
                            } else if (!BWD && last) {
                                float part[P2];
#pragma unroll
                                for (int v = 0; v < P2; ++v) part[v] = (v < C * GP) ? out[v / GP][v % GP] * wl_u : 0.f;
                                warp_reduce_scatter<P2>(part, lane);
                                if (lane < C * GP) sHead[ub * N + (lane / GP) * T + g * GP + (lane % GP)] = part[0];
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
                                // (1) next operand into shared memory first: it is what the MMA warp waits for
                                const uint32_t rowoff = (uint32_t)(u >> 3) * G::SK + (uint32_t)(u & 7) * 128;
                                uint32_t hiw[C][GP / 2], low[C][GP / 2];
#pragma unroll
                                for (int c = 0; c < C; ++c) {
                                    const int ncol = c * T + g * GP;
#pragma unroll
                                    for (int i = 0; i < GP / 2; ++i) split2(out[c][2 * i], out[c][2 * i + 1], hiw[c][i], low[c][i]);
                                    if constexpr (GP >= 8) {
#pragma unroll
                                        for (int h = 0; h < GP / 8; ++h) {       // one 16-byte chunk per 8 points
                                            const int nc = ncol + 8 * h;
                                            const uint32_t off = rowoff + (uint32_t)(nc >> 6) * G::SN + (((((uint32_t)nc & 63) >> 3) << 4) ^ usw);
                                            sts_v4(sB + off, hiw[c][4 * h], hiw[c][4 * h + 1], hiw[c][4 * h + 2], hiw[c][4 * h + 3]);
                                            if constexpr (PASSES == 3)
                                                sts_v4(sB + G::B_PART + off, low[c][4 * h], low[c][4 * h + 1], low[c][4 * h + 2], low[c][4 * h + 3]);
                                        }
                                    } else {
                                        const uint32_t off = rowoff + (uint32_t)(ncol >> 6) * G::SN +
                                                             (((((uint32_t)ncol & 63) >> 3) << 4) ^ usw) + (uint32_t)(ncol & 4) * 2;
                                        sts_v2(sB + off, hiw[c][0], hiw[c][1]);
                                        if constexpr (PASSES == 3) sts_v2(sB + G::B_PART + off, low[c][0], low[c][1]);
                                    }
                                }
                                // (2) this warp's last item of the phase: release the MMA warp before the global stores
                                if (it + G::WQ >= G::NITEMS) {
                                    fence_proxy_async();
                                    tc_fence_before();
                                    mbar_arrive(&b_ready[ph]);
                                }
                                // (3) what the reverse pass / wgrad need, to HBM (128-byte lines per warp)
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
                            // ---------------- per-unit parameter gradients of this item ---------------------------------
                            if (BWD) {
                                if (n_acc) {
                                    if (j == 0) atomicAdd(sAcc + (Lh + 4) * H + u, wlsum);
                                    if (l >= 1 || !a.sb_grad) atomicAdd(sAcc + l * H + u, 30.f * bsum);
                                    if (l == 0) {
                                        if constexpr (D > 0) atomicAdd(sAcc + (Lh + 1) * H + u, 30.f * w0s0);
                                        if constexpr (D > 1) atomicAdd(sAcc + (Lh + 2) * H + u, 30.f * w0s1);
                                        if constexpr (D > 2) atomicAdd(sAcc + (Lh + 3) * H + u, 30.f * w0s2);
                                    }
                                } else {
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
                            }
                        }
                    }
                    if (j > 0) pa ^= 1;     // acc_ready[0..NPH) complete once per MMA step: one parity for both
                }
            };
            if (full) run_tile(std::true_type{});
            else run_tile(std::false_type{});

            if constexpr (!BWD) {
                // ---- last layer: sum the per-unit-block partials, apply the output head ------------------------------
                named_bar_sync(1, G::ETHREADS);
                for (int e = tid; e < T; e += G::ETHREADS) {
                    int64_t gp = p0 + e;
                    if (gp < a.n) {
                        float v[C];
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            float s = 0.f;
#pragma unroll
                            for (int w = 0; w < G::NUB; ++w) s += sHead[w * N + c * T + e];
                            v[c] = s;
                        }
                        DIGS_ASSERT(gp >= 0 && gp < a.n);
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
            named_bar_sync(1, G::ETHREADS);   // sHead / sUbar are rewritten by the next tile
            layer0_done = pipe_next;
            if (pipe_next) xb ^= 1;
        }
        if (n_acc) {
            // every epilogue warp has passed the barrier above for its last tile: the CTA's sums are complete
            for (int i = tid; i < n_acc; i += G::ETHREADS) {
                const float v = sAcc[i];
                const int row = i / H, uu = i % H;
                if (row <= Lh) {
                    if (row >= 1 || !a.sb_grad) atomicAdd(a.db[row] + uu, v);
                } else if (row <= Lh + 3) {
                    if (row - (Lh + 1) < D) atomicAdd(a.dW0 + (int64_t)uu * a.w0_stride + (row - (Lh + 1)), v);
                } else {
                    atomicAdd(a.dw_last + uu, v);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == EPI_WARPS + 1) tmem_dealloc<G::TMEM_COLS>(tmem_base);
}

// ---- weight preparation: 30*W (forward operand) and (30*W)^T (reverse operand) as bf16 hi/lo; 30*b; 30*W0 ---------
struct PrepPtrs {
    const float* W[DIGS_MAX_LAYERS];
    const float* b[DIGS_MAX_LAYERS];
};
// Wq layout: [4][Lh][H][H] bf16 with slot 0 = hi, 1 = lo, 2 = hi transposed, 3 = lo transposed
__global__ void k_tc_prep(PrepPtrs p, int Lh, int H, int d, int w0_stride, __nv_bfloat16* Wq, float* bias30, float* w0_30) {
    const int64_t per = (int64_t)H * H;
    const int64_t total = (int64_t)Lh * per;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int l = (int)(e / per);
        int64_t r = e - (int64_t)l * per;
        int u = (int)(r / H), k = (int)(r % H);
        float w = 30.f * p.W[1 + l][r];
        __nv_bfloat16 hi = __float2bfloat16_rn(w);
        __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
        Wq[e] = hi;
        Wq[total + e] = lo;
        int64_t et = (int64_t)l * per + (int64_t)k * H + u;
        Wq[2 * total + et] = hi;
        Wq[3 * total + et] = lo;
    }
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Lh * H) bias30[i] = 30.f * p.b[1 + i / H][i % H];
    if (i < H) {
        const float* w = p.W[0] + (int64_t)i * w0_stride;
        w0_30[i * 4 + 0] = 30.f * w[0];
        w0_30[i * 4 + 1] = 30.f * w[1];
        w0_30[i * 4 + 2] = (d > 2) ? 30.f * w[2] : 0.f;
        w0_30[i * 4 + 3] = 30.f * p.b[0][i];
    }
}

// ---- host ---------------------------------------------------------------------------------------------------------------
EncodeFn get_encode() {
    static EncodeFn fn = nullptr;
    if (!fn) {
        cudaDriverEntryPointQueryResult q;
        void* p = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeFn>(p);
    }
    return fn;
}

bool supported(const DigsNet* net) { return net->H == 256 || net->H == 128 || net->H == 512; }

static size_t wq_bytes(const DigsNet* net) { return align_up((size_t)4 * net->Lh * net->H * net->H * 2, 256); }
static size_t bias_bytes(const DigsNet* net) { return align_up((size_t)net->Lh * net->H * 4, 256); }
static size_t w0_bytes(const DigsNet* net) { return align_up((size_t)net->H * 16, 256); }
size_t prep_ws_bytes(const DigsNet* net) { return wq_bytes(net) + bias_bytes(net) + w0_bytes(net); }

SavedLayout saved_layout(const DigsNet* net, int64_t n, int want_lap) {
    SavedLayout s;
    s.C = 1 + net->d + (want_lap ? 1 : 0);
    s.n_pad = (n + 7) / 8 * 8;
    s.planes_bytes = align_up((size_t)net->Lh * s.C * net->H * s.n_pad * 4, 256);
    s.act_bytes = align_up((size_t)net->Lh * 2 * net->H * s.C * s.n_pad * 2, 256);
    s.head_bytes = align_up((size_t)n * s.C * 4, 256);
    return s;
}
size_t saved_bytes(const DigsNet* net, int64_t n, int want_lap) {
    SavedLayout s = saved_layout(net, n, want_lap);
    return s.planes_bytes + s.act_bytes + s.head_bytes;
}
size_t bwd_ws_bytes(const DigsNet* net, int64_t n, int want_lap) {
    SavedLayout s = saved_layout(net, n, want_lap);
    return prep_ws_bytes(net) + s.act_bytes /* GP_words */ + s.head_bytes /* ubar */;
}

int num_sms() {      // of the CURRENT device (cached per device ordinal)
    static int n[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int& slot = n[dev & 63];
    if (!slot) cudaDeviceGetAttribute(&slot, cudaDevAttrMultiProcessorCount, dev);
    return slot;
}

template <int D, int LAP, int H, int PASSES, bool BWD>
static int launch(const CUtensorMap& tmap, const Args& a, cudaStream_t st) {
    using G = Cfg<D, LAP, H, PASSES>;
    auto kern = k_tc_net<D, LAP, H, PASSES, BWD>;
    static bool attr_set[64] = {false};      // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) { set_error("tc net: smem attribute: %s", cudaGetErrorString(e)); return DIGS_ECUDA; }
        attr_set[dev & 63] = true;
    }
    Args aa = a;
    int smem_bytes = G::SMEM_BYTES;
    if (BWD) {      // shared-memory gradient accumulators when they fit next to the operand buffer and the weight ring
        const int acc_bytes = (a.Lh + 5) * H * 4;
        const int with_acc = G::OFF_ACC + acc_bytes + 1024;    // the accumulators take the place of the forward's head scratch
        aa.acc_smem = (with_acc <= 232448) ? 1 : 0;
        if (aa.acc_smem && with_acc > smem_bytes) smem_bytes = with_acc;
    }
    int64_t tiles = (a.n_pad + G::T - 1) / G::T;
    int grid = (int)(tiles < num_sms() ? tiles : num_sms());
    kern<<<grid, NTHREADS, smem_bytes, st>>>(tmap, aa);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("tc net launch: %s", cudaGetErrorString(e)); return DIGS_ECUDA; }
    return DIGS_OK;
}

template <int D, int LAP, int PASSES, bool BWD>
static int launch_h(const DigsNet* net, const CUtensorMap& tmap, const Args& a, cudaStream_t st) {
    switch (net->H) {
        case 128: return launch<D, LAP, 128, PASSES, BWD>(tmap, a, st);
        case 256: return launch<D, LAP, 256, PASSES, BWD>(tmap, a, st);
        case 512: return launch<D, LAP, 512, PASSES, BWD>(tmap, a, st);
    }
    set_error("tc net: H=%d not supported (128, 256, 512)", net->H);
    return DIGS_ENOTIMPL;
}

// The CTA-pair kernel (jet_tc_pair.cu, H = 256) is correct but measured slower than this file's kernel on the B200.  It is
// NOT part of the default build: `make PAIR=1` compiles it in, and it then runs only when DIGS_B200_PAIR=1 is set.
#ifdef DIGS_WITH_PAIR
static bool use_pair(const DigsNet* net) {
    static int pair = -1;
    if (pair < 0) {
        const char* e = getenv("DIGS_B200_PAIR");
        pair = (e && e[0] == '1') ? 1 : 0;
    }
    return pair && pair_supported(net);
}
#endif

template <bool BWD>
static int dispatch(const DigsNet* net, const CUtensorMap& tmap, const Args& a, int want_jet, int want_lap, int passes,
                    cudaStream_t st) {
#ifdef DIGS_WITH_PAIR
    if (use_pair(net)) return launch_pair(net, tmap, a, want_jet, want_lap, passes, BWD, st);
#endif
    if (!want_jet) {
        if constexpr (BWD) { set_error("tc reverse needs jet channels"); return DIGS_EINVAL; }
        else return (passes == 3) ? launch_h<0, 0, 3, false>(net, tmap, a, st) : launch_h<0, 0, 1, false>(net, tmap, a, st);
    }
    if (net->d == 3 && want_lap) return (passes == 3) ? launch_h<3, 1, 3, BWD>(net, tmap, a, st) : launch_h<3, 1, 1, BWD>(net, tmap, a, st);
    if (net->d == 3) return (passes == 3) ? launch_h<3, 0, 3, BWD>(net, tmap, a, st) : launch_h<3, 0, 1, BWD>(net, tmap, a, st);
    if (want_lap) return (passes == 3) ? launch_h<2, 1, 3, BWD>(net, tmap, a, st) : launch_h<2, 1, 1, BWD>(net, tmap, a, st);
    return (passes == 3) ? launch_h<2, 0, 3, BWD>(net, tmap, a, st) : launch_h<2, 0, 1, BWD>(net, tmap, a, st);
}

// Weight operands of the tensor-core kernels: taken from net->prepared (digs_prepare_weights, hoisted out of the per-call
// path) or, when that is NULL, built in `ws` by one k_tc_prep launch; then the TMA map over them is encoded.
int run_prep(const DigsNet* net, void* dst, cudaStream_t st) {
    char* w = reinterpret_cast<char*>(dst);
    __nv_bfloat16* Wq = reinterpret_cast<__nv_bfloat16*>(w);
    float* bias30 = reinterpret_cast<float*>(w + wq_bytes(net));
    float* w0_30 = reinterpret_cast<float*>(w + wq_bytes(net) + bias_bytes(net));
    PrepPtrs pp;
    for (int l = 0; l <= net->Lh + 1; ++l) { pp.W[l] = net->W[l]; pp.b[l] = net->b[l]; }
    k_tc_prep<<<num_sms() * 4, 256, 0, st>>>(pp, net->Lh, net->H, net->d, net->w0_stride, Wq, bias30, w0_30);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("tc prep launch: %s", cudaGetErrorString(e)); return DIGS_ECUDA; }
    return DIGS_OK;
}

int prepare_weights(const DigsNet* net, void* ws, CUtensorMap* tmap, Prep* out, cudaStream_t st) {
    EncodeFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return DIGS_ECUDA; }
    char* w = net->prepared ? reinterpret_cast<char*>(const_cast<void*>(net->prepared)) : reinterpret_cast<char*>(ws);
    if (!net->prepared) {
        int rc = run_prep(net, ws, st);
        if (rc) return rc;
    }
    __nv_bfloat16* Wq = reinterpret_cast<__nv_bfloat16*>(w);
    cuuint64_t gdim[2] = {(cuuint64_t)net->H, (cuuint64_t)4 * net->Lh * net->H};
    cuuint64_t gstr[1] = {(cuuint64_t)net->H * 2};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Wq, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights) failed (%d)", (int)r); return DIGS_ECUDA; }
    out->bias30 = reinterpret_cast<float*>(w + wq_bytes(net));
    out->w0_30 = reinterpret_cast<float*>(w + wq_bytes(net) + bias_bytes(net));
    return DIGS_OK;
}

static int prepare(const DigsNet* net, void* ws, CUtensorMap* tmap, Args* a, cudaStream_t st) {
    Prep p;
    int rc = prepare_weights(net, ws, tmap, &p, st);
    if (rc) return rc;
    a->Lh = net->Lh;
    a->w0_30 = p.w0_30;
    a->bias30 = p.bias30;
    a->w_last = net->W[net->Lh + 1];
    a->b_last = net->b[net->Lh + 1];
    a->head = net->head;
    a->radius = net->radius;
    a->scale = net->scale;
    return DIGS_OK;
}

// Forward jet / value query.  want_jet = 0 -> value channel only (C = 1).  saved: NULL or saved_bytes() bytes.
int forward(const DigsNet* net, const PreSrc& src, int64_t n, int want_jet, int want_lap, float* f, float* grad,
            float* lap, void* saved, void* ws, int passes, cudaStream_t st) {
    if (n == 0) return DIGS_OK;
    CUtensorMap tmap;
    Args a{};
    int rc = prepare(net, ws, &tmap, &a, st);
    if (rc) return rc;
    a.src = src;
    a.n = n;
    a.f = f; a.grad = grad; a.lap = lap;
    if (saved) {
        SavedLayout s = saved_layout(net, n, want_lap);
        char* sp = reinterpret_cast<char*>(saved);
        a.n_pad = s.n_pad;
        a.planes = reinterpret_cast<float*>(sp);
        a.actP = reinterpret_cast<__nv_bfloat16*>(sp + s.planes_bytes);
        a.head_saved = reinterpret_cast<float*>(sp + s.planes_bytes + s.act_bytes);
    } else {
        a.n_pad = (n + 7) / 8 * 8;
    }
    return dispatch<false>(net, tmap, a, want_jet, want_lap, passes, st);
}

// Reverse pass: head adjoint (fp32 kernel) -> fused dgrad kernel (bias / layer-0 / last-layer gradients, GP_words) -> wgrad GEMMs.
int backward(const DigsNet* net, const PreSrc& src, int64_t n, int want_lap, const float* f_bar, const float* g_bar,
             const float* l_bar, const void* saved, void* ws, const DigsGrads* grads, float* sb_grad, int passes,
             cudaStream_t st) {
    if (n == 0) return DIGS_OK;
    SavedLayout s = saved_layout(net, n, want_lap);
    const char* sp = reinterpret_cast<const char*>(saved);
    float* planes = reinterpret_cast<float*>(const_cast<char*>(sp));
    __nv_bfloat16* actP = reinterpret_cast<__nv_bfloat16*>(const_cast<char*>(sp) + s.planes_bytes);
    const float* head_saved = reinterpret_cast<const float*>(sp + s.planes_bytes + s.act_bytes);
    char* w = reinterpret_cast<char*>(ws);
    __nv_bfloat16* GP_words = reinterpret_cast<__nv_bfloat16*>(w + prep_ws_bytes(net));
    float* ubar = reinterpret_cast<float*>(w + prep_ws_bytes(net) + s.act_bytes);
    int rc = simt::head_seed(net, n, want_lap, head_saved, f_bar, g_bar, l_bar, ubar, grads->db[net->Lh + 1], st);
    if (rc) return rc;
    CUtensorMap tmap;
    Args a{};
    rc = prepare(net, ws, &tmap, &a, st);
    if (rc) return rc;
    a.src = src;
    a.n = n;
    a.n_pad = s.n_pad;
    a.planes = planes;
    a.ubar = ubar;
    a.GP_words = GP_words;
    for (int l = 0; l <= net->Lh; ++l) a.db[l] = grads->db[l];
    a.dW0 = grads->dW[0];
    a.w0_stride = net->w0_stride;
    a.dw_last = grads->dW[net->Lh + 1];
    a.sb_grad = sb_grad;
    rc = dispatch<true>(net, tmap, a, 1, want_lap, passes, st);
    if (rc) return rc;
    return wgrad(net, GP_words, actP, s.C, s.n_pad, grads, passes, st);
}

}  // namespace tcpath
}  // namespace digs
