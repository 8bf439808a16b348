// The whole hot-path step of one batch, fused per point tile (DIGS_MODE_BF16X2 / DIGS_MODE_BF16): digs_step_fused.
//
// Reference behaviour reproduced, per iteration (train_surface_reconstruction.py:121-128):
//   DiGSNetwork.forward on both clouds        models/DiGS.py:43-66, 122-134
//   utils.gradient x (1 + d) inside DiGSLoss  utils/utils.py:108-117, models/losses.py:72-76, 100-106
//   DiGSLoss.forward terms                    models/losses.py:107-184
//   loss.backward()                           train_surface_reconstruction.py:128
//
// Every DiGS loss term is a mean of per-point functions, so the adjoint seeds of a point depend only on its own
// (f, grad f, lap f) and on global constants (weights, global point counts).  One persistent CTA per SM therefore walks
// point tiles of BOTH clouds and runs, per tile,
//
//     forward jet (Lh tcgen05 layer steps) -> last layer + output head -> loss terms + seeds + head adjoint
//     -> reverse jet (Lh tcgen05 layer steps with (30 W)^T) -> layer-0 / bias / last-layer gradients
//
// with the same warp roles, operand layouts and two-phase layer step as jet_tc.cu (see there).  What changes is where
// the state between the two directions lives:
//   * the tile's pre-activation planes (30 z | 30 Z_k | 30 Q per hidden layer) go to a per-CTA scratch block
//     [Lh][C][T/4][H][4] fp32 (about 0.5 MB, rewritten by every tile) that stays resident in L2 -- they never have to
//     reach HBM; the two-call API (digs_jet_forward / digs_jet_backward) stores 2 x 553 MB of them per step;
//   * the head values (u | gu_k | lu) and their adjoints stay in shared memory;
//   * only the two operand streams of the weight-gradient GEMM are written to global memory, tile-major
//     ([layer][column][unit] words of (bf16 hi | bf16 lo << 16), column = tile_col0 + channel * T + point), and ONE
//     k_tc_wgrad launch over the columns of both clouds finishes the step.
// The loss scalars (models/losses.py:186-188) are summed per CTA in shared memory and added to terms_out once.
#include "jet_tc_shared.cuh"
#include <cuda.h>
#include <type_traits>
#include <cstdlib>

namespace digs {
namespace tcpath {

struct StepSeg {                // one cloud of the step
    const float* xyz;
    int64_t xyz_stride;
    int64_t n;
    const float* normals;       // surface cloud: (n, d) or NULL
    const float* shape_bias;    // NULL or (n_shapes, H)
    int64_t pps;
    float* sb_grad;             // NULL or (n_shapes, H), accumulated
    float* f;                   // outputs (n), (n, d), (n) | NULL
    float* grad;
    float* lap;
    int lap_chan;               // 1: this cloud carries the Laplacian channel (C = 2 + d)
    int kind;                   // 0 = off-surface cloud (inter, eikonal, div); 1 = surface cloud (sdf, eikonal, normal)
    int64_t tile0;              // first tile of this cloud in the step's tile list
    int64_t tiles;
    int64_t col0;               // first operand column of this cloud in the weight-gradient operands
};

struct StepArgs {
    StepSeg seg[2];
    int64_t total_tiles;
    int64_t total_cols;
    int Lh;
    const float* w0_30;
    const float* bias30;
    const float* w_last;
    const float* b_last;
    int head;
    float radius, scale;
    int d_in;
    // loss (models/losses.py:107-184)
    float w_sdf, w_inter, w_normal, w_eik, w_div;
    const float* w_dev;
    float inv_nm, inv_nn, inv_nall, div_clamp;
    int add_normal, use_div, div_l2;
    float* terms;               // [8], accumulated
    // state between the directions / towards the weight-gradient GEMM
    float* scratch;             // per-CTA plane blocks
    int64_t scratch_stride;     // floats per CTA
    uint32_t* actP;             // [Lh][total_cols][H]
    uint32_t* GPw;              // [Lh][total_cols][H]
    // per-unit gradients
    float* db[DIGS_MAX_LAYERS];
    float* dW0;
    int w0_stride;
    float* dw_last;
    float* db_last;
    int acc_smem;
    unsigned int* round_ctr;    // NULL or WgradSync counters: [round] += 1 per finished tile, [WG_MAX_ROUNDS] += 1 per finished CTA
};

// Two experiments on this kernel that did NOT pay (profiles/r02_experiments.md):
//  * register redistribution with setmaxnreg (20 warps, helpers 32 / epilogue 112 registers -- the pool is the launch-time
//    allocation, 5 x 96 per scheduler, so 4 x 112 + 32 is the most that fits; 120 / 32 and 112 / 56 never get their
//    allocation and hang): spills gone, tile kernel 0.690 ms against 0.675 ms without it;
//  * a compact run-time loop over the weight-tile order in the MMA / TMA warps instead of the unrolled callbacks:
//    the issue thread's index arithmetic sits on the critical path (0.732 ms, grid query 228 ms instead of 204 ms).
template <int D, int H_, int PASSES>
struct SCfg {
    using G1 = Cfg<D, 1, H_, PASSES>;
    using G0 = Cfg<D, 0, H_, PASSES>;
    static constexpr int H = H_;
    static constexpr int N = G1::N;
    static constexpr int OFF_RING = G1::B_BYTES;
    static constexpr int OFF_XYZ = OFF_RING + STAGES * STAGE_BYTES;   // [N*3] floats
    static constexpr int OFF_UBAR = OFF_XYZ + N * 3 * 4;              // [N] floats: head adjoints (u_bar | gu_bar | lu_bar) per point
    static constexpr int OFF_BAR = OFF_UBAR + N * 4;
    static constexpr int OFF_HEAD = OFF_BAR + 256;                    // [NUB][N] floats: last-layer partial sums
    static constexpr int OFF_LOSS = OFF_HEAD + G1::NUB * N * 4;       // 16 floats: loss partial sums, db_last
    static constexpr int OFF_ACC = OFF_LOSS + 64;                     // [(Lh+5)][H] floats (optional)
    static constexpr int SMEM_BASE = OFF_ACC + 1024;                  // + alignment slack
    static constexpr int PLANE_FLOATS_1 = G1::C * G1::T * H;          // floats per layer slot, C = 2 + D
    static constexpr int PLANE_FLOATS_0 = G0::C * G0::T * H;
    static constexpr int PLANE_FLOATS = PLANE_FLOATS_1 > PLANE_FLOATS_0 ? PLANE_FLOATS_1 : PLANE_FLOATS_0;
    static_assert(G0::N == G1::N && G0::B_BYTES == G1::B_BYTES && G0::TMEM_COLS == G1::TMEM_COLS, "clouds must share the operand tile");
    static_assert(G0::GP == 4 && G1::GP == 4, "jet tiles use 4-point groups");
    static_assert(G0::T * 3 <= N * 3 / 2 && G1::T * 3 <= N * 3 / 2, "coordinates and normals share the staging block");
};

// L2 load that the compiler must leave where it is written (the plane prefetch of the NEXT item belongs at the end of the
// current one: hoisted, its 4 x C registers would stay live across the whole item and spill)
__device__ __forceinline__ float4 ldcg_pinned(const float* p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ float sgnf(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

// ---- epilogue warps: one tile, both directions -------------------------------------------------------------------------
// SEG selects a.seg[SEG] statically: its fields are then constant-bank operands (a run-time index would turn every
// access into an indexed constant load on the long scoreboard).
template <int D, int LAP, int H, int PASSES, int SEG>
__device__ __forceinline__ void epi_tile(const StepArgs& a, const int64_t tl, uint8_t* smem,
                                         const uint32_t tmem_base, uint64_t* b_ready, uint64_t* acc_ready, uint32_t& pa,
                                         float* planes_cta, const int n_acc, float* lsum) {
    using G = Cfg<D, LAP, H, PASSES>;
    using S = SCfg<D, H, PASSES>;
    constexpr int C = G::C, T = G::T, N = G::N, NPH = G::NPH, GP = 4;
    constexpr int P2 = (C * GP <= 16) ? 16 : 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, wr = warp >> 2;
    const uint32_t sB = smem_u32(smem);
    float* sXyz = reinterpret_cast<float*>(smem + S::OFF_XYZ);
    float* sUbar = reinterpret_cast<float*>(smem + S::OFF_UBAR);
    float* sHead = reinterpret_cast<float*>(smem + S::OFF_HEAD);
    float* sAcc = reinterpret_cast<float*>(smem + S::OFF_ACC);
    const uint32_t usw = (uint32_t)(lane & 7) << 4;
    const StepSeg& sg = a.seg[SEG];
    const int Lh = a.Lh;
    const int64_t p0 = tl * T;
    const int64_t col_tile = sg.col0 + tl * (C * T);                  // first operand column of this tile
    const int64_t slot_stride = a.total_cols * (int64_t)H;            // words per layer slot of actP / GPw

    // ---- stage the tile's coordinates -------------------------------------------------------------------------------
    for (int e = tid; e < T; e += G::ETHREADS) {
        int64_t gp = p0 + e;
        if (gp >= sg.n) gp = sg.n - 1;
        const float* xp = sg.xyz + gp * sg.xyz_stride;
        sXyz[e * 3] = __ldg(xp);
        sXyz[e * 3 + 1] = __ldg(xp + 1);
        sXyz[e * 3 + 2] = (D > 2) ? __ldg(xp + 2) : 0.f;
        if (sg.kind == 1 && sg.normals) {      // the loss section between the two directions is serial: its inputs are staged here
            const float* np_ = sg.normals + gp * D;
#pragma unroll
            for (int k = 0; k < D; ++k) sXyz[N * 3 / 2 + e * D + k] = __ldg(np_ + k);
        }
    }
    named_bar_sync(1, G::ETHREADS);

    auto run_pass = [&](auto bwd_tag) {
        constexpr bool BWD = decltype(bwd_tag)::value;
#ifndef DIGS_DIAG_NO_PREFETCH
        // Reverse direction: the stored planes of an item do not depend on the accumulators.  They are fetched from L2 one item
        // ahead -- at the end of the item before it in this warp's program order, across phase and layer boundaries -- so the
        // L2 latency runs under the wait for the tensor pipe instead of after it.  Every load is unconditional (clamped to a
        // valid address when there is no next item): a conditionally defined array would be demoted to local memory.
        float4 pv[C];
        auto fetch_planes = [&](int ln, int phn, int itn) {
            const int mtn = phn * G::MTP + itn / G::NGRP, gn = itn % G::NGRP;
            const float* src = planes_cta + (int64_t)(ln >= 1 ? ln - 1 : 0) * (C * T * H) + gn * GP * H + ((mtn * 4 + q) * 32 + lane) * 4;
#pragma unroll
            for (int c = 0; c < C; ++c) pv[c] = ldcg_pinned(src + c * (T * H));
        };
        if (BWD) fetch_planes(Lh, 0, wr % G::WQ);
#endif
        for (int j = 0; j <= Lh; ++j) {
            const int l = BWD ? (Lh - j) : j;
            const bool last = (j == Lh);
            const uint32_t t_buf = tmem_base + ((uint32_t)(q * 32) << 16) + (j & 1) * G::ACC_COLS;
            float* pl_l = planes_cta + (int64_t)(l - 1) * (C * T * H);         // scratch slot of layer l (l >= 1)
            uint32_t* wd_l = (BWD ? a.GPw : a.actP) + (int64_t)(BWD ? (l - 1) : l) * slot_stride + col_tile * H;
#pragma unroll 1
            for (int ph = 0; ph < NPH; ++ph) {
                const int it_first = (wr + G::WQ - (ph * G::ROT) % G::WQ) % G::WQ;
                if (j > 0) {      // the accumulators of this phase's M tiles (the MMA warp may still be filling the others)
                    mbar_wait(&acc_ready[ph], pa);
                    tc_fence_after();
                }
#pragma unroll 1
                for (int it = it_first; it < G::NITEMS; it += G::WQ) {
                    const int mt = ph * G::MTP + it / G::NGRP;
                    const int g = it % G::NGRP;
                    const int ub = mt * 4 + q;
                    const int u = ub * 32 + lane;
                    const int64_t pg = p0 + g * GP;
                    const int poff = g * GP * H + u * 4;                  // plane offset ([point / 4][unit][point % 4])
                    DIGS_ASSERT((int64_t)(Lh - 1) * (C * T * H) + (C - 1) * (T * H) + poff + 4 <= a.scratch_stride && g < G::NGRP && mt < G::MT);
                    const uint32_t t_base = t_buf + mt * N;
                    float bsum = 0.f, wlsum = 0.f, w0s0 = 0.f, w0s1 = 0.f, w0s2 = 0.f;
                    float acc[C][GP];
                    // reverse hidden steps: sin / cos / |Z|^2 of the stored planes do not need the accumulators -- they run under
                    // the TMEM load (the wait follows them); the empty asm pins every use of acc[] behind the wait
                    constexpr bool kLateWaitDir = BWD;
                    const bool late_wait = kLateWaitDir && l >= 1;
                    auto acc_wait = [&]() {
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < C; ++c)
#pragma unroll
                            for (int i = 0; i < GP; ++i) asm volatile("" : "+f"(acc[c][i]));
                    };
                    if (j > 0) {
#pragma unroll
                        for (int c = 0; c < C; ++c) tmem_ld4(t_base + c * T + g * GP, acc[c]);
                        if (!late_wait) acc_wait();
                    }
                    float pre[C][GP];
                    if (l == 0) {
                        const float4 w0v = __ldg(reinterpret_cast<const float4*>(a.w0_30) + u);
#pragma unroll
                        for (int i = 0; i < GP; ++i) {
                            const float* xp = sXyz + (g * GP + i) * 3;
                            float bz = w0v.w;
                            if (sg.shape_bias) {
                                int64_t gp = pg + i < sg.n ? pg + i : sg.n - 1;
                                bz = 30.f * __ldg(sg.shape_bias + (gp / sg.pps) * H + u);
                            }
                            pre[0][i] = fmaf(w0v.z, xp[2], fmaf(w0v.y, xp[1], fmaf(w0v.x, xp[0], bz)));
                            pre[1][i] = w0v.x;
                            pre[2][i] = w0v.y;
                            if constexpr (D > 2) pre[3][i] = w0v.z;
                            if constexpr (LAP) pre[1 + D][i] = 0.f;
                        }
                    } else if (!BWD) {
                        const float blv = __ldg(a.bias30 + (l - 1) * H + u);
#pragma unroll
                        for (int c = 0; c < C; ++c)
#pragma unroll
                            for (int i = 0; i < GP; ++i) pre[c][i] = (c == 0) ? acc[c][i] + blv : acc[c][i];
                    } else {
#pragma unroll
                        for (int c = 0; c < C; ++c) {
#ifdef DIGS_DIAG_NO_PLANES
                            const float4 v = make_float4(0.1f, 0.2f, 0.3f, 0.4f);
#elif !defined(DIGS_DIAG_NO_PREFETCH)
                            const float4 v = pv[c];
#else
                            const float4 v = __ldcg(reinterpret_cast<const float4*>(pl_l + c * (T * H) + poff));   // written by this CTA: L2, not L1
#endif
                            pre[c][0] = v.x; pre[c][1] = v.y; pre[c][2] = v.z; pre[c][3] = v.w;
                        }
                    }
                    float wl_u = 0.f;
                    if ((BWD && j == 0) || (!BWD && last)) wl_u = __ldg(a.w_last + u);
                    if (BWD && j == 0) {
#pragma unroll
                        for (int c = 0; c < C; ++c)
#pragma unroll
                            for (int i = 0; i < GP; ++i) acc[c][i] = sUbar[(g * GP + i) * C + c] * wl_u;
                    }
                    // ---------------- per-point jet math ------------------------------------------------
                    float out[C][GP];
                    float sn[GP], cs[GP], S2v[GP];
#pragma unroll
                    for (int i = 0; i < GP; ++i) {
                        sincos_reduced(pre[0][i], sn[i], cs[i]);
                        float S2 = 0.f;
#pragma unroll
                        for (int k = 0; k < D; ++k) S2 = fmaf(pre[1 + k][i], pre[1 + k][i], S2);
                        S2v[i] = S2;
                    }
                    if (late_wait && j > 0) acc_wait();
#pragma unroll
                    for (int i = 0; i < GP; ++i) {
                        const float s = sn[i], c = cs[i], S2 = S2v[i];
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
                    // ---------------- scatter -----------------------------------------------------------------------
                    if (BWD && l == 0) {
#pragma unroll
                        for (int i = 0; i < GP; ++i) {
                            const float* xp = sXyz + (g * GP + i) * 3;
                            bsum += out[0][i];
                            w0s0 += fmaf(out[0][i], xp[0], out[1][i]);
                            w0s1 += fmaf(out[0][i], xp[1], out[2][i]);
                            if constexpr (D > 2) w0s2 += fmaf(out[0][i], xp[2], out[3][i]);
                            if (sg.sb_grad && pg + i < sg.n)
                                atomicAdd(sg.sb_grad + ((pg + i) / sg.pps) * H + u, 30.f * out[0][i]);
                        }
                    } else if (!BWD && last) {
                        float part[P2];
#pragma unroll
                        for (int v = 0; v < P2; ++v) part[v] = (v < C * GP) ? out[v / GP][v % GP] * wl_u : 0.f;
                        warp_reduce_scatter<P2>(part, lane);
                        if (lane < C * GP) sHead[ub * N + (lane / GP) * T + g * GP + (lane % GP)] = part[0];
#ifndef DIGS_DIAG_NO_PLANES
#pragma unroll
                        for (int c = 0; c < C; ++c)
                            *reinterpret_cast<float4*>(pl_l + c * (T * H) + poff) = make_float4(pre[c][0], pre[c][1], pre[c][2], pre[c][3]);
#endif
                    } else {
                        if (BWD) {
#pragma unroll
                            for (int i = 0; i < GP; ++i) bsum += out[0][i];
                        }
                        // (1) next operand into shared memory first: it is what the MMA warp waits for
                        const uint32_t rowoff = (uint32_t)(u >> 3) * G::SK + (uint32_t)(u & 7) * 128;
                        uint32_t hiw[C][2], low[C][2];
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            const int ncol = c * T + g * GP;
                            split2(out[c][0], out[c][1], hiw[c][0], low[c][0]);
                            split2(out[c][2], out[c][3], hiw[c][1], low[c][1]);
                            const uint32_t off = rowoff + (uint32_t)(ncol >> 6) * G::SN +
                                                 (((((uint32_t)ncol & 63) >> 3) << 4) ^ usw) + (uint32_t)(ncol & 4) * 2;
                            DIGS_ASSERT(ncol + GP <= N && off + 8 <= (uint32_t)G::B_PART && u < H);
                            sts_v2(sB + off, hiw[c][0], hiw[c][1]);
                            if constexpr (PASSES == 3) sts_v2(sB + G::B_PART + off, low[c][0], low[c][1]);
                        }
                        // (2) this warp's last item of the phase: release the MMA warp before the global stores
                        if (it + G::WQ >= G::NITEMS) {
                            fence_proxy_async();
                            tc_fence_before();
                            mbar_arrive(&b_ready[ph]);
                        }
                        // (3) forward: this layer's pre-activations to the L2-resident scratch (layer 0 is recomputed)
#ifndef DIGS_DIAG_NO_PLANES
                        if (!BWD && j > 0) {
#pragma unroll
                            for (int c = 0; c < C; ++c)
                                *reinterpret_cast<float4*>(pl_l + c * (T * H) + poff) = make_float4(pre[c][0], pre[c][1], pre[c][2], pre[c][3]);
                        }
#endif
                        // (4) weight-gradient operand words (hi | lo << 16), tile-major columns; streamed (evict-first)
#ifndef DIGS_DIAG_NO_WORDS      // diagnostic builds (make DIAG=NO_WORDS): timing only, results are wrong
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            uint32_t* dst = wd_l + (c * T + g * GP) * H + u;
                            DIGS_ASSERT(col_tile + c * T + g * GP + GP <= a.total_cols && (BWD ? l - 1 : l) >= 0 && (BWD ? l - 1 : l) < Lh);
#pragma unroll
                            for (int i = 0; i < GP; ++i) {
                                uint32_t w = __byte_perm(hiw[c][i >> 1], low[c][i >> 1], (i & 1) ? 0x7632 : 0x5410);
                                // padding points (>= n) need no masking: their head adjoints are exact zeros (sUbar), so
                                // every adjoint word of such a column is 0 and the finite layer-input words drop out
                                __stcs(dst + i * H, w);
                            }
                        }
#endif
                    }
                    // ---------------- per-unit parameter gradients of this item ---------------------------------------
                    if (BWD) {
                        if (n_acc) {
                            if (j == 0) atomicAdd(sAcc + (Lh + 4) * H + u, wlsum);
                            if (l >= 1 || !sg.sb_grad) atomicAdd(sAcc + l * H + u, 30.f * bsum);
                            if (l == 0) {
                                atomicAdd(sAcc + (Lh + 1) * H + u, 30.f * w0s0);
                                atomicAdd(sAcc + (Lh + 2) * H + u, 30.f * w0s1);
                                if constexpr (D > 2) atomicAdd(sAcc + (Lh + 3) * H + u, 30.f * w0s2);
                            }
                        } else {
                            if (j == 0) atomicAdd(a.dw_last + u, wlsum);
                            if (l >= 1) {
                                atomicAdd(a.db[l] + u, 30.f * bsum);
                            } else {
                                if (!sg.sb_grad) atomicAdd(a.db[0] + u, 30.f * bsum);
                                atomicAdd(a.dW0 + (int64_t)u * a.w0_stride + 0, 30.f * w0s0);
                                atomicAdd(a.dW0 + (int64_t)u * a.w0_stride + 1, 30.f * w0s1);
                                if constexpr (D > 2) atomicAdd(a.dW0 + (int64_t)u * a.w0_stride + 2, 30.f * w0s2);
                            }
                        }
                    }
#ifndef DIGS_DIAG_NO_PREFETCH
                    if (BWD) {      // this warp's next item: same phase, else the first item of the next phase / layer step
                        int itn = it + G::WQ, phn = ph, ln = l;
                        if (itn >= G::NITEMS) {
                            if (++phn == NPH) { phn = 0; --ln; }
                            itn = (wr + G::WQ - (phn * G::ROT) % G::WQ) % G::WQ;
                        }
                        fetch_planes(ln, phn, itn);
                    }
#endif
                }
            }
            if (j > 0) pa ^= 1;     // acc_ready[0..NPH) complete once per MMA step: one parity for both
        }
    };

    // ================= forward ==========================================================================================
    run_pass(std::false_type{});

    // ================= last layer sums, output head, loss terms, seeds, head adjoint (one warp: T <= 32 points) ========
    named_bar_sync(1, G::ETHREADS);
    if (warp < (T + 31) / 32) {
        const int e = tid;
        const int64_t gp = p0 + e;
        const bool valid = (e < T) && (gp < sg.n);
        float t_sdf = 0.f, t_inter = 0.f, t_eik = 0.f, t_nrm = 0.f, t_div = 0.f, ub_sum = 0.f;
        float ubv[C];
#pragma unroll
        for (int c = 0; c < C; ++c) ubv[c] = 0.f;
        if (valid) {
            float v[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < G::NUB; ++w) s += sHead[w * N + c * T + e];
                v[c] = s;
            }
            const float uu = v[0] + __ldg(a.b_last);
            // ---- output head (models/DiGS.py:128-132) and its derivatives
            float f, gr[D], lp = 0.f;
            float aa = 1.f, sg_u = 0.f, ra = 1.f, g1 = 1.f, g2 = 0.f, g3 = 0.f, gg = 0.f;
            if (a.head == DIGS_HEAD_IDENTITY) {
                f = uu;
#pragma unroll
                for (int k = 0; k < D; ++k) gr[k] = v[1 + k];
                if constexpr (LAP) lp = v[1 + D];
            } else {
                aa = fabsf(uu) + 1e-8f;
                sg_u = sgnf(uu);
                ra = sqrtf(aa);
                g1 = 0.5f / ra;
                g2 = -0.25f * sg_u / (aa * ra);
                g3 = 0.375f / (aa * aa * ra);
                f = a.scale * (sg_u * ra - a.radius);
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    gr[k] = a.scale * g1 * v[1 + k];
                    gg = fmaf(v[1 + k], v[1 + k], gg);
                }
                if constexpr (LAP) lp = a.scale * (g1 * v[1 + D] + g2 * gg);
            }
            DIGS_ASSERT(gp >= 0 && gp < sg.n);
            sg.f[gp] = f;
#pragma unroll
            for (int k = 0; k < D; ++k) sg.grad[gp * D + k] = gr[k];
            if constexpr (LAP) if (sg.lap) sg.lap[gp] = lp;
            // ---- loss terms and adjoint seeds (models/losses.py:107-168; same arithmetic as k_loss)
            float w_sdf = a.w_sdf, w_inter = a.w_inter, w_normal = a.w_normal, w_eik = a.w_eik, w_div = a.w_div;
            if (a.w_dev) {
                w_sdf = __ldg(a.w_dev); w_inter = __ldg(a.w_dev + 1); w_normal = __ldg(a.w_dev + 2);
                w_eik = __ldg(a.w_dev + 3); w_div = __ldg(a.w_dev + 4);
            }
            float nr2 = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) nr2 = fmaf(gr[k], gr[k], nr2);
            const float nr = sqrtf(nr2);
            t_eik = fabsf(nr - 1.f) * a.inv_nall;
            const float ke = (nr > 0.f) ? w_eik * a.inv_nall * sgnf(nr - 1.f) / nr : 0.f;
            float fb, gb[D], lb = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) gb[k] = ke * gr[k];
            if (sg.kind == 1) {
                t_sdf = fabsf(f) * a.inv_nm;
                fb = w_sdf * a.inv_nm * sgnf(f);
                if (sg.normals) {
                    float nv[D], nn2 = 0.f, dot = 0.f;
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        nv[k] = sXyz[N * 3 / 2 + e * D + k];
                        nn2 = fmaf(nv[k], nv[k], nn2);
                        dot = fmaf(gr[k], nv[k], dot);
                    }
                    const float gc = fmaxf(nr, 1e-8f), nc = fmaxf(sqrtf(nn2), 1e-8f);
                    const float cs = dot / (gc * nc);
                    t_nrm = (1.f - fabsf(cs)) * a.inv_nm;
                    if (a.add_normal) {
                        const float k2 = -w_normal * a.inv_nm * sgnf(cs);
#pragma unroll
                        for (int k = 0; k < D; ++k) gb[k] += k2 * (nv[k] / (gc * nc) - cs / (gc * gc) * gr[k]);
                    }
                }
            } else {
                const float ex = expf(-100.f * fabsf(f));
                t_inter = ex * a.inv_nn;
                fb = -100.f * w_inter * a.inv_nn * sgnf(f) * ex;
                if constexpr (LAP) {
                    if (a.use_div) {
                        float lv = lp;
                        const bool isn = (lv != lv);
                        if (isn) lv = 0.f;                                   // models/losses.py:107
                        const float t = a.div_l2 ? lv * lv : fabsf(lv);
                        const float dt = a.div_l2 ? 2.f * lv : sgnf(lv);
                        t_div = fminf(fmaxf(t, 0.1f), a.div_clamp) * a.inv_nn;
                        const bool inside = (t >= 0.1f) && (t <= a.div_clamp) && !isn;
                        lb = inside ? w_div * a.inv_nn * dt : 0.f;
                    }
                }
            }
            // ---- head adjoint (SURVEY 8a "head adjoint")
            if (a.head == DIGS_HEAD_IDENTITY) {
                ubv[0] = fb;
#pragma unroll
                for (int k = 0; k < D; ++k) ubv[1 + k] = gb[k];
                if constexpr (LAP) ubv[1 + D] = lb;
            } else {
                float gbg = 0.f;
#pragma unroll
                for (int k = 0; k < D; ++k) gbg = fmaf(gb[k], v[1 + k], gbg);
                const float lu = LAP ? v[C - 1] : 0.f;
                ubv[0] = a.scale * (g1 * fb + g2 * gbg + lb * (g2 * lu + g3 * gg));
#pragma unroll
                for (int k = 0; k < D; ++k) ubv[1 + k] = a.scale * (g1 * gb[k] + 2.f * g2 * lb * v[1 + k]);
                if constexpr (LAP) ubv[1 + D] = a.scale * g1 * lb;
            }
            ub_sum = ubv[0];
        }
        if (e < T) {
#pragma unroll
            for (int c = 0; c < C; ++c) sUbar[e * C + c] = ubv[c];
        }
        float red[6] = {t_sdf, t_inter, t_eik, t_nrm, t_div, ub_sum};
#pragma unroll
        for (int k = 0; k < 6; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) red[k] += __shfl_xor_sync(0xffffffffu, red[k], o);
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 6; ++k) atomicAdd(lsum + k, red[k]);
        }
    }
    named_bar_sync(1, G::ETHREADS);

    // ================= reverse ==========================================================================================
    run_pass(std::true_type{});
    named_bar_sync(1, G::ETHREADS);   // sXyz / sHead / sUbar are rewritten by the next tile
}

template <int D, int H, int PASSES>
__global__ void __launch_bounds__(NTHREADS, 1) k_tc_step(const __grid_constant__ CUtensorMap tmapW, const StepArgs a) {
    using S = SCfg<D, H, PASSES>;
    using G = typename S::G1;
    constexpr int N = S::N, MT = G::MT;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sRing = smem + S::OFF_RING;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
    uint64_t* w_full = bars;
    uint64_t* w_empty = bars + STAGES;
    uint64_t* b_ready = bars + 2 * STAGES;
    uint64_t* acc_ready = bars + 2 * STAGES + 2;      // [2]: accumulators of the M tiles of phase p complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    float* sLoss = reinterpret_cast<float*>(smem + S::OFF_LOSS);
    float* sAcc = reinterpret_cast<float*>(smem + S::OFF_ACC);

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
    for (int i = tid; i < S::OFF_RING / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    const int n_acc = a.acc_smem ? (a.Lh + 5) * H : 0;
    for (int i = tid; i < n_acc; i += NTHREADS) sAcc[i] = 0.f;
    if (tid < 16) sLoss[tid] = 0.f;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int Lh = a.Lh;
    // the weight-gradient GEMM is a programmatic dependent launch: its CTAs may be dispatched as soon as SMs free up (they
    // synchronise on round_ctr, not on the completion of this grid)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#ifdef DIGS_DIAG_TRACE
    if (a.round_ctr && tid == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        reinterpret_cast<unsigned long long*>(a.round_ctr + WG_MAX_ROUNDS + 1)[256 + blockIdx.x] = t;
    }
#endif

    if (warp == EPI_WARPS) {
        // ================= TMA producer: per tile the forward weights (layers 1..Lh), then the transposed ones (Lh..1) ==
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t r = blockIdx.x; r < a.total_tiles; r += gridDim.x)
                for (int dir = 0; dir < 2; ++dir)
                    for (int j = 0; j < Lh; ++j) {
                        const int l = dir ? (Lh - j) : (j + 1);
                        auto load = [&](int ks, int m) {
                            for (int part = 0; part < G::NPARTS; ++part) {
                                mbar_wait(&w_empty[stage], phase ^ 1);
                                mbar_arrive_expect_tx(&w_full[stage], STAGE_BYTES);
                                tma_load_2d(sRing + stage * STAGE_BYTES, &tmapW, &w_full[stage], ks * 64,
                                            (((dir ? 2 : 0) + part) * Lh + (l - 1)) * H + m * 128);
                                if (++stage == STAGES) { stage = 0; phase ^= 1; }
                            }
                        };
                        weight_tile_order<G>(load, [](int) {}, [](int) {});
                    }
        }
    } else if (warp == EPI_WARPS + 1) {
        // ================= MMA issuer: 2 * Lh layer steps per tile =========================================================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, N, /*A K-major*/ 0, /*B MN-major*/ 1);
            const uint32_t sB_hi = smem_u32(smem), sB_lo = sB_hi + G::B_PART;
            int stage = 0;
            uint32_t phase = 0, pb = 0;
            for (int64_t r = blockIdx.x; r < a.total_tiles; r += gridDim.x)
                for (int dir = 0; dir < 2; ++dir)
                    for (int j = 0; j < Lh; ++j) {
                        const uint32_t d_buf = tmem_base + ((j + 1) & 1) * G::ACC_COLS;
                        auto issue = [&](int ks, int m) {
                            const uint32_t d_tmem = d_buf + m * N;
                            mbar_wait(&w_full[stage], phase);
                            tc_fence_after();
                            {
                                const uint32_t sA = smem_u32(sRing + stage * STAGE_BYTES);
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk) {
                                    uint64_t da = make_smem_desc(sA + kk * 32, 16, 1024, LAYOUT_SW128);
                                    uint32_t koff = (ks * 8 + kk * 2) * G::SK;
                                    uint64_t db = make_smem_desc(sB_hi + koff, G::SN, G::SK, LAYOUT_SW128);
#ifndef DIGS_DIAG_NO_MMA
                                    umma_bf16(d_tmem, da, db, idesc, (ks | kk) != 0);
                                    if constexpr (PASSES == 3) {
                                        uint64_t dbl = make_smem_desc(sB_lo + koff, G::SN, G::SK, LAYOUT_SW128);
                                        umma_bf16(d_tmem, da, dbl, idesc, 1);
                                    }
#else
                                    (void)da; (void)db; (void)d_tmem;
#endif
                                }
                            }
                            umma_commit(&w_empty[stage]);
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                            if constexpr (PASSES == 3) {
                                mbar_wait(&w_full[stage], phase);
                                tc_fence_after();
                                const uint32_t sA = smem_u32(sRing + stage * STAGE_BYTES);
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk) {
                                    uint64_t da = make_smem_desc(sA + kk * 32, 16, 1024, LAYOUT_SW128);
                                    uint32_t koff = (ks * 8 + kk * 2) * G::SK;
                                    uint64_t db = make_smem_desc(sB_hi + koff, G::SN, G::SK, LAYOUT_SW128);
#ifndef DIGS_DIAG_NO_MMA
                                    umma_bf16(d_tmem, da, db, idesc, 1);
#else
                                    (void)da; (void)db;
#endif
                                }
                                umma_commit(&w_empty[stage]);
                                if (++stage == STAGES) { stage = 0; phase ^= 1; }
                            }
                        };
                        weight_tile_order<G>(issue,
                                             [&](int ph) { mbar_wait(&b_ready[ph], pb); tc_fence_after(); },
                                             [&](int ph) { umma_commit(&acc_ready[ph]); });
                        pb ^= 1;
                    }
        }
    } else {
        // ================= epilogue warps ===================================================================================
        uint32_t pa = 0;
        float* planes_cta = a.scratch + (int64_t)blockIdx.x * a.scratch_stride;
        float* lsum = sLoss;                                 // per-CTA sums: sdf, inter, eikonal, normal, div, sum of u_bar
        for (int64_t tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
            // cloud 0 carries the Laplacian channel iff the loss has a divergence term; cloud 1 never does
            if (tile >= a.seg[1].tile0)
                epi_tile<D, 0, H, PASSES, 1>(a, tile - a.seg[1].tile0, smem, tmem_base, b_ready, acc_ready, pa, planes_cta, n_acc, lsum);
            else if (a.seg[0].lap_chan)
                epi_tile<D, 1, H, PASSES, 0>(a, tile, smem, tmem_base, b_ready, acc_ready, pa, planes_cta, n_acc, lsum);
            else
                epi_tile<D, 0, H, PASSES, 0>(a, tile, smem, tmem_base, b_ready, acc_ready, pa, planes_cta, n_acc, lsum);
            // every epilogue thread has issued the tile's operand words (barrier at the end of epi_tile): publish the tile
            if (a.round_ctr && tid == 0) {
                __threadfence();
                atomicAdd(a.round_ctr + tile / gridDim.x, 1u);
            }
        }
        if (tid == 0) {
            // terms: [0] loss [1] sdf [2] inter [3] eikonal [4] normal [5] div (same slots as k_loss)
            float w_sdf = a.w_sdf, w_inter = a.w_inter, w_normal = a.w_normal, w_eik = a.w_eik, w_div = a.w_div;
            if (a.w_dev) {
                w_sdf = a.w_dev[0]; w_inter = a.w_dev[1]; w_normal = a.w_dev[2]; w_eik = a.w_dev[3]; w_div = a.w_dev[4];
            }
            atomicAdd(a.terms + 1, lsum[0]);
            atomicAdd(a.terms + 2, lsum[1]);
            atomicAdd(a.terms + 3, lsum[2]);
            atomicAdd(a.terms + 4, lsum[3]);
            atomicAdd(a.terms + 5, lsum[4]);
            atomicAdd(a.terms, w_sdf * lsum[0] + w_inter * lsum[1] + w_eik * lsum[2] +
                                   (a.add_normal ? w_normal * lsum[3] : 0.f) + (a.use_div ? w_div * lsum[4] : 0.f));
            atomicAdd(a.db_last, lsum[5]);
        }
        if (n_acc) {
            const bool any_sb = a.seg[0].sb_grad || a.seg[1].sb_grad;
            for (int i = tid; i < n_acc; i += G::ETHREADS) {
                const float v = sAcc[i];
                const int row = i / H, uu = i % H;
                if (row <= Lh) {
                    if (row >= 1 || !any_sb) atomicAdd(a.db[row] + uu, v);
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
    if (a.round_ctr && tid == 0) {      // the per-unit gradients and loss sums of this CTA are out as well
        __threadfence();
        atomicAdd(a.round_ctr + WG_MAX_ROUNDS, 1u);
    }
    if (warp == EPI_WARPS + 1) tmem_dealloc<G::TMEM_COLS>(tmem_base);
#ifdef DIGS_DIAG_TRACE
    if (a.round_ctr && tid == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        reinterpret_cast<unsigned long long*>(a.round_ctr + WG_MAX_ROUNDS + 1)[blockIdx.x] = t;
    }
#endif
}

// ---- host ---------------------------------------------------------------------------------------------------------------
struct StepLayout {
    int C[2], T[2];
    int64_t tiles[2], cols[2];
    int64_t total_tiles, total_cols;
    size_t words_bytes;       // one operand stream (actP or GPw)
    size_t scratch_bytes;     // all CTAs
    size_t scratch_stride;    // floats per CTA
};

static StepLayout step_layout(const DigsNet* net, int64_t nn, int64_t nm, int want_lap) {
    StepLayout L{};
    const int ncols = (net->H > 256) ? 64 : 128;
    const int64_t n[2] = {nn, nm};
    const int lap[2] = {want_lap ? 1 : 0, 0};
    int64_t plane_floats = 0;
    for (int s = 0; s < 2; ++s) {
        L.C[s] = 1 + net->d + lap[s];
        L.T[s] = (ncols / L.C[s]) / 4 * 4;
        L.tiles[s] = (n[s] + L.T[s] - 1) / L.T[s];
        L.cols[s] = L.tiles[s] * L.C[s] * L.T[s];
        int64_t pf = (int64_t)L.C[s] * L.T[s] * net->H;
        if (pf > plane_floats) plane_floats = pf;
    }
    L.total_tiles = L.tiles[0] + L.tiles[1];
    L.total_cols = L.cols[0] + L.cols[1];
    L.words_bytes = align_up((size_t)net->Lh * L.total_cols * net->H * 4, 256);
    L.scratch_stride = (size_t)net->Lh * plane_floats;
    L.scratch_bytes = align_up((size_t)num_sms() * L.scratch_stride * 4, 256);
    return L;
}

size_t step_ws_bytes(const DigsNet* net, int64_t nn, int64_t nm, int want_lap) {
    StepLayout L = step_layout(net, nn, nm, want_lap);
    return prep_ws_bytes(net) + 2 * L.words_bytes + L.scratch_bytes + WG_CTR_BYTES;
}

template <int D, int H, int PASSES>
static int launch_step(const CUtensorMap& tmap, StepArgs& a, cudaStream_t st) {
    using S = SCfg<D, H, PASSES>;
    auto kern = k_tc_step<D, H, PASSES>;
    int smem_bytes = S::SMEM_BASE;
    const int with_acc = S::OFF_ACC + (a.Lh + 5) * H * 4 + 1024;
    a.acc_smem = (with_acc <= 232448) ? 1 : 0;
    if (a.acc_smem) smem_bytes = with_acc;
    static bool attr_set[64] = {false};      // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    cudaError_t e;
    if (!attr_set[dev & 63]) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (e != cudaSuccess) { set_error("tc step: smem attribute: %s", cudaGetErrorString(e)); return DIGS_ECUDA; }
        attr_set[dev & 63] = true;
    }
    int grid = (int)(a.total_tiles < num_sms() ? a.total_tiles : num_sms());
    kern<<<grid, NTHREADS, smem_bytes, st>>>(tmap, a);
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("tc step launch: %s", cudaGetErrorString(e)); return DIGS_ECUDA; }
    return DIGS_OK;
}

int step_fused(const DigsNet* net, const float* nonmnfld, int64_t nn, const float* mnfld, int64_t nm, const float* normals,
               int64_t xyz_stride, const float* sb_n, const float* sb_m, int64_t pps_n, int64_t pps_m, int want_lap,
               const DigsLossCfg* cfg, float* f_n, float* g_n, float* lap_n, float* f_m, float* g_m, float* terms,
               const DigsGrads* grads, float* sbg_n, float* sbg_m, void* ws, int passes, cudaStream_t st) {
    if (nn + nm == 0) return DIGS_OK;
    StepLayout L = step_layout(net, nn, nm, want_lap);
    char* w = reinterpret_cast<char*>(ws);
    StepArgs a{};
    CUtensorMap tmap;
    Prep pr;
    int rc = prepare_weights(net, ws, &tmap, &pr, st);
    if (rc) return rc;
    a.Lh = net->Lh;
    a.w0_30 = pr.w0_30;
    a.bias30 = pr.bias30;
    a.w_last = net->W[net->Lh + 1];
    a.b_last = net->b[net->Lh + 1];
    a.head = net->head;
    a.radius = net->radius;
    a.scale = net->scale;
    a.d_in = net->d;
    a.actP = reinterpret_cast<uint32_t*>(w + prep_ws_bytes(net));
    a.GPw = reinterpret_cast<uint32_t*>(w + prep_ws_bytes(net) + L.words_bytes);
    a.scratch = reinterpret_cast<float*>(w + prep_ws_bytes(net) + 2 * L.words_bytes);
    a.scratch_stride = (int64_t)L.scratch_stride;
    // tile-round counters for the overlapped weight-gradient GEMM (DIGS_B200_WGRAD_OVERLAP=0: plain stream order)
    const int grid_ctas = (int)(L.total_tiles < num_sms() ? L.total_tiles : num_sms());
    // Worth it only when the last round leaves a visible share of the SMs idle (C2: 1094 tiles = 7.4 rounds on 148 CTAs, 7.6 %
    // of the kernel's SM time; C4 / C5: 14.8 rounds, 1.5 % -- there the finer work units of the overlapped GEMM cost more
    // than the tail returns: measured +2 % / +0.5 % step time).  DIGS_B200_WGRAD_OVERLAP=0 / 1 forces either way.
    const int64_t rounds = (L.total_tiles + grid_ctas - 1) / grid_ctas;
    const double idle_share = (double)(rounds * grid_ctas - L.total_tiles) / (double)(rounds * grid_ctas);
    bool overlap = idle_share >= 0.03;
    if (const char* e = getenv("DIGS_B200_WGRAD_OVERLAP")) overlap = (e[0] != '0');
    overlap = overlap && rounds <= WG_QUEUE_IDX;
    if (const char* e = getenv("DIGS_B200_STEP_NO_WGRAD")) overlap = overlap && e[0] != '1';
    if (overlap) {
        a.round_ctr = reinterpret_cast<unsigned int*>(w + prep_ws_bytes(net) + 2 * L.words_bytes + L.scratch_bytes);
        cudaError_t me = cudaMemsetAsync(a.round_ctr, 0, WG_CTR_BYTES, st);
        if (me != cudaSuccess) { set_error("tc step: counter reset: %s", cudaGetErrorString(me)); return DIGS_ECUDA; }
    }
    a.total_tiles = L.total_tiles;
    a.total_cols = L.total_cols;
    const float* xyz[2] = {nonmnfld, mnfld};
    const int64_t n[2] = {nn, nm};
    for (int s = 0; s < 2; ++s) {
        StepSeg& g = a.seg[s];
        g.xyz = xyz[s];
        g.xyz_stride = xyz_stride;
        g.n = n[s];
        g.normals = s ? normals : nullptr;
        g.shape_bias = s ? sb_m : sb_n;
        g.pps = (s ? pps_m : pps_n) > 0 ? (s ? pps_m : pps_n) : 1;
        g.sb_grad = s ? sbg_m : sbg_n;
        g.f = s ? f_m : f_n;
        g.grad = s ? g_m : g_n;
        g.lap = s ? nullptr : lap_n;
        g.lap_chan = s ? 0 : (want_lap ? 1 : 0);
        g.kind = s;
        g.tile0 = s ? L.tiles[0] : 0;
        g.tiles = L.tiles[s];
        g.col0 = s ? L.cols[0] : 0;
    }
    a.w_sdf = cfg->weights[0]; a.w_inter = cfg->weights[1]; a.w_normal = cfg->weights[2];
    a.w_eik = cfg->weights[3]; a.w_div = cfg->weights[4];
    a.w_dev = cfg->weights_device;
    const int64_t Nm = cfg->Nm_total > 0 ? cfg->Nm_total : nm, Nn = cfg->Nn_total > 0 ? cfg->Nn_total : nn;
    a.inv_nm = Nm > 0 ? 1.f / (float)Nm : 0.f;
    a.inv_nn = Nn > 0 ? 1.f / (float)Nn : 0.f;
    a.inv_nall = (Nm + Nn) > 0 ? 1.f / (float)(Nm + Nn) : 0.f;
    a.div_clamp = cfg->div_clamp;
    a.add_normal = cfg->add_normal; a.use_div = cfg->use_div; a.div_l2 = cfg->div_l2;
    a.terms = terms;
    for (int l = 0; l <= net->Lh; ++l) a.db[l] = grads->db[l];
    a.db_last = grads->db[net->Lh + 1];
    a.dW0 = grads->dW[0];
    a.w0_stride = net->w0_stride;
    a.dw_last = grads->dW[net->Lh + 1];

    const int key = net->d * 10 + passes;
    switch (net->H) {
        case 128:
            if (key == 33) rc = launch_step<3, 128, 3>(tmap, a, st);
            else if (key == 31) rc = launch_step<3, 128, 1>(tmap, a, st);
            else if (key == 23) rc = launch_step<2, 128, 3>(tmap, a, st);
            else rc = launch_step<2, 128, 1>(tmap, a, st);
            break;
        case 256:
            if (key == 33) rc = launch_step<3, 256, 3>(tmap, a, st);
            else if (key == 31) rc = launch_step<3, 256, 1>(tmap, a, st);
            else if (key == 23) rc = launch_step<2, 256, 3>(tmap, a, st);
            else rc = launch_step<2, 256, 1>(tmap, a, st);
            break;
        case 512:
            if (key == 33) rc = launch_step<3, 512, 3>(tmap, a, st);
            else if (key == 31) rc = launch_step<3, 512, 1>(tmap, a, st);
            else if (key == 23) rc = launch_step<2, 512, 3>(tmap, a, st);
            else rc = launch_step<2, 512, 1>(tmap, a, st);
            break;
        default:
            set_error("tc step: H=%d not supported (128, 256, 512)", net->H);
            return DIGS_ENOTIMPL;
    }
    if (rc) return rc;
    // measurement-only switch (bench.py times the tile kernel alone for its roofline entry): skip the GEMM launch
    if (const char* e = getenv("DIGS_B200_STEP_NO_WGRAD")) {
        if (e[0] == '1') return DIGS_OK;
    }
    // one weight-gradient GEMM over the columns of both clouds (tile-major columns: C = 1, n_pad = total_cols)
    WgradSync sync{};
    sync.ctr = a.round_ctr;
    sync.grid = grid_ctas;
    sync.total_tiles = L.total_tiles;
    sync.tiles0 = L.tiles[0];
    sync.cols0 = L.cols[0];
    sync.ct0 = L.C[0] * L.T[0];
    sync.ct1 = L.C[1] * L.T[1];
    return wgrad(net, reinterpret_cast<const __nv_bfloat16*>(a.GPw), reinterpret_cast<const __nv_bfloat16*>(a.actP), 1,
                 L.total_cols, grads, passes, st, &sync);
}

}  // namespace tcpath
}  // namespace digs
