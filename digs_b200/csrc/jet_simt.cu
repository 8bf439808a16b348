// fp32 CUDA-core implementation of the DiGS jet path (DIGS_MODE_FP32).
//
// Reference behaviour reproduced (SURVEY.md section 8a closed form):
//   forward   models/DiGS.py:122-157 (+ head :128-132) and the derivative channels autograd would give
//             utils/utils.py:108-117, models/losses.py:72-76,100-106
//   backward  loss.backward() from (dL/df, dL/dgrad, dL/dlap) to the parameter gradients
//
// Layout in HBM: a "jet plane set" is [C][n][H] fp32, C = 1 + d + want_lap channels in the order
// (z | Z_0..Z_{d-1} | Q).  Only PRE-activations are ever stored; each consumer re-applies
// sin/cos on load, so one stored tensor serves the next layer's GEMM, the backward adjoint
// formulas and the weight-gradient GEMM.
//
// Every GEMM is a shared-memory tiled SGEMM whose loader performs the per-element jet math
// (activation in the forward pass, adjoint recurrence in the reverse pass) while staging the
// operand, so no activated / adjoint tensor is materialised beyond what the next kernel needs.
#include "digs_internal.h"
#include <cstdio>

namespace digs {
namespace simt {

constexpr float OMEGA = 30.0f;
constexpr int TN = 64;   // output columns per CTA
constexpr int KS = 16;   // reduction slab
constexpr int NTHREADS = 256;

template <int D, int LAP>
struct Cfg {
    static constexpr int C = 1 + D + LAP;
    static constexpr int TP = (C == 1) ? 128 : 32;  // points per CTA
    static constexpr int R = C * TP;                // operand rows per CTA
    static constexpr int RT = R / 16;               // rows per thread
    static constexpr int RPAD = R + 2;              // smem row pitch: conflict-free transposed stores
    static_assert(RT % 2 == 0, "rows per thread must be even (float2 smem loads)");
};

// ------------------------------------------------------------------------------------------
// element fetch: pre-activations (z, Z_k, Q) of one (point, unit)
// ------------------------------------------------------------------------------------------
template <int D, int LAP, bool FROM_XYZ>
__device__ __forceinline__ void fetch_pre(const PreSrc& s, int64_t gp, int k, int H, float& z, float* Z, float& Q) {
    constexpr int DD = (D > 0) ? D : 1;
    if constexpr (!FROM_XYZ) {
        const float* base = s.planes + gp * H + k;
        z = __ldg(base);
#pragma unroll
        for (int j = 0; j < D; ++j) Z[j] = __ldg(base + (1 + j) * s.plane_stride);
        if constexpr (LAP) Q = __ldg(base + (1 + D) * s.plane_stride);
        else Q = 0.f;
    } else {
        float x[3];
        if (s.mode == 2) {
            int64_t g = s.grid_base + gp;
            int iz = (int)(g % s.nz);
            int64_t t = g / s.nz;
            int ix = (int)(t % s.nx);
            int iy = (int)(t / s.nx);
            x[0] = __half2float(s.ax[ix]);
            x[1] = __half2float(s.ay[iy]);
            x[2] = __half2float(s.az[iz]);
        } else {
            const float* xp = s.xyz + gp * s.xyz_stride;
            x[0] = __ldg(xp);
            x[1] = __ldg(xp + 1);
            x[2] = (s.din > 2) ? __ldg(xp + 2) : 0.f;
        }
        const float* w = s.W0 + (int64_t)k * s.w0_stride;
        float bz = s.shape_bias ? __ldg(s.shape_bias + (gp / s.pps) * H + k) : __ldg(s.b0 + k);
        float w0 = __ldg(w), w1 = __ldg(w + 1), w2 = (s.din > 2) ? __ldg(w + 2) : 0.f;
        z = fmaf(w2, x[2], fmaf(w1, x[1], fmaf(w0, x[0], bz)));
        if constexpr (D > 0) {
            Z[0] = w0;
            if constexpr (DD > 1) Z[1] = w1;
            if constexpr (DD > 2) Z[2] = w2;
        }
        Q = 0.f;
    }
}

// activation jet: (z,Z,Q) -> out[C] = (h | A_k | P)
template <int D, int LAP>
__device__ __forceinline__ void activate(float z, const float* Z, float Q, float* out) {
    float s, c;
    sincosf(OMEGA * z, &s, &c);
    out[0] = s;
    if constexpr (D > 0) {
        float oc = OMEGA * c;
        float S2 = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) {
            out[1 + j] = oc * Z[j];
            S2 = fmaf(Z[j], Z[j], S2);
        }
        if constexpr (LAP) out[1 + D] = oc * Q - (OMEGA * OMEGA) * s * S2;
    }
}

// adjoint of the activation jet: incoming (hb, Ab_k, Pb) -> (zb | Zb_k | Qb)
template <int D, int LAP>
__device__ __forceinline__ void activate_adjoint(float z, const float* Z, float Q, float hb, const float* Ab, float Pb,
                                                 float* out) {
    float s, c;
    sincosf(OMEGA * z, &s, &c);
    const float w1 = OMEGA, w2 = OMEGA * OMEGA, w3 = OMEGA * OMEGA * OMEGA;
    float oc = w1 * c;
    float S2 = 0.f, ZA = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        S2 = fmaf(Z[j], Z[j], S2);
        ZA = fmaf(Z[j], Ab[j], ZA);
    }
    float zb = oc * hb - w2 * s * ZA;
    if constexpr (LAP) zb -= Pb * (w2 * s * Q + w3 * c * S2);
    out[0] = zb;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        float v = oc * Ab[j];
        if constexpr (LAP) v -= 2.f * w2 * s * Z[j] * Pb;
        out[1 + j] = v;
    }
    if constexpr (LAP) out[1 + D] = oc * Pb;
}

// one KS-deep slab of the register-tiled product: acc[r][j] += Xs[kk][row0+r] * Ws[kk][col0+j]
template <int RT, int RPAD>
__device__ __forceinline__ void slab_fma(const float (*Xs)[RPAD], const float (*Ws)[TN], float (*acc)[4], int row0,
                                         int col0) {
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
        float4 w = *reinterpret_cast<const float4*>(&Ws[kk][col0]);
        float xr[RT];
#pragma unroll
        for (int r = 0; r < RT; r += 2) {
            float2 v = *reinterpret_cast<const float2*>(&Xs[kk][row0 + r]);
            xr[r] = v.x;
            xr[r + 1] = v.y;
        }
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            acc[r][0] = fmaf(xr[r], w.x, acc[r][0]);
            acc[r][1] = fmaf(xr[r], w.y, acc[r][1]);
            acc[r][2] = fmaf(xr[r], w.z, acc[r][2]);
            acc[r][3] = fmaf(xr[r], w.w, acc[r][3]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// K1: hidden layer forward.  pre_out[c][p][u] = sum_k act_c(pre_in)[p][k] * W[u][k]  (+ b[u] on c = 0)
//   Wt is the transposed weight (k-major) so slab loads are coalesced.
// ------------------------------------------------------------------------------------------
template <int D, int LAP, bool FROM_XYZ>
__global__ void __launch_bounds__(NTHREADS) k_hidden_fwd(PreSrc src, const float* __restrict__ Wt,
                                                         const float* __restrict__ bias, float* __restrict__ pre_out,
                                                         int64_t n, int H) {
    using G = Cfg<D, LAP>;
    constexpr int C = G::C, TP = G::TP, RT = G::RT, RPAD = G::RPAD;
    __shared__ __align__(16) float Xs[KS][RPAD];
    __shared__ __align__(16) float Ws[KS][TN];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t p0 = (int64_t)blockIdx.x * TP;
    const int n0 = blockIdx.y * TN;
    float acc[RT][4];
#pragma unroll
    for (int r = 0; r < RT; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;

    for (int k0 = 0; k0 < H; k0 += KS) {
        for (int e = tid; e < TP * KS; e += NTHREADS) {
            int p = e / KS, kk = e % KS;
            int64_t gp = p0 + p;
            float out[C];
#pragma unroll
            for (int c = 0; c < C; ++c) out[c] = 0.f;
            if (gp < n) {
                float z, Z[D > 0 ? D : 1], Q;
                fetch_pre<D, LAP, FROM_XYZ>(src, gp, k0 + kk, H, z, Z, Q);
                activate<D, LAP>(z, Z, Q, out);
            }
#pragma unroll
            for (int c = 0; c < C; ++c) Xs[kk][c * TP + p] = out[c];
        }
        for (int e = tid; e < KS * TN / 4; e += NTHREADS) {
            int kk = e / (TN / 4), j4 = e % (TN / 4);
            *reinterpret_cast<float4*>(&Ws[kk][j4 * 4]) =
                __ldg(reinterpret_cast<const float4*>(Wt + (int64_t)(k0 + kk) * H + n0 + j4 * 4));
        }
        __syncthreads();
        slab_fma<RT, RPAD>(Xs, Ws, acc, ty * RT, tx * 4);
        __syncthreads();
    }
    const int64_t ps = n * (int64_t)H;
    float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n0 + tx * 4));
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        int row = ty * RT + r;
        int c = row / TP, p = row % TP;
        int64_t gp = p0 + p;
        if (gp < n) {
            float4 v = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
            if (c == 0) { v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w; }
            *reinterpret_cast<float4*>(pre_out + c * ps + gp * H + n0 + tx * 4) = v;
        }
    }
}

// ------------------------------------------------------------------------------------------
// K2: last layer + head.  One warp per point.
//   head_saved[p][0..C) = (u | gu_k | lu) for the reverse pass.
// ------------------------------------------------------------------------------------------
template <int D, int LAP, bool FROM_XYZ>
__global__ void __launch_bounds__(256) k_head_fwd(PreSrc src, const float* __restrict__ w_last,
                                                  const float* __restrict__ b_last, int head, float radius, float scale,
                                                  float* __restrict__ f, float* __restrict__ grad,
                                                  float* __restrict__ lap, float* __restrict__ head_saved, int64_t n,
                                                  int H) {
    constexpr int C = 1 + D + LAP;
    const int lane = threadIdx.x & 31;
    const int64_t gp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gp >= n) return;
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    for (int k = lane; k < H; k += 32) {
        float z, Z[D > 0 ? D : 1], Q, out[C];
        fetch_pre<D, LAP, FROM_XYZ>(src, gp, k, H, z, Z, Q);
        activate<D, LAP>(z, Z, Q, out);
        float w = __ldg(w_last + k);
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = fmaf(out[c], w, acc[c]);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    }
    if (lane == 0) {
        float u = acc[0] + __ldg(b_last);
        if (head_saved) {
            head_saved[gp * C] = u;
#pragma unroll
            for (int c = 1; c < C; ++c) head_saved[gp * C + c] = acc[c];
        }
        if (head == DIGS_HEAD_IDENTITY) {
            f[gp] = u;
            if constexpr (D > 0) {
#pragma unroll
                for (int j = 0; j < D; ++j) grad[gp * D + j] = acc[1 + j];
                if constexpr (LAP) lap[gp] = acc[1 + D];
            }
        } else {
            float a = fabsf(u) + 1e-8f;
            float sg = (u > 0.f) ? 1.f : ((u < 0.f) ? -1.f : 0.f);
            float ra = sqrtf(a);
            f[gp] = scale * (sg * ra - radius);
            if constexpr (D > 0) {
                float g1 = 0.5f / ra;
                float gg = 0.f;
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    grad[gp * D + j] = scale * g1 * acc[1 + j];
                    gg = fmaf(acc[1 + j], acc[1 + j], gg);
                }
                if constexpr (LAP) {
                    float g2 = -0.25f * sg / (a * ra);
                    lap[gp] = scale * (g1 * acc[1 + D] + g2 * gg);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// K3a: head adjoint per point: (f_bar, g_bar, l_bar) -> ubar[p][C] = (u_bar | gu_bar_k | lu_bar); db_last += sum u_bar
// ------------------------------------------------------------------------------------------
template <int D, int LAP>
__global__ void __launch_bounds__(256) k_head_seed(const float* __restrict__ head_saved, const float* __restrict__ f_bar,
                                                   const float* __restrict__ g_bar, const float* __restrict__ l_bar,
                                                   int head, float scale, float* __restrict__ ubar,
                                                   float* __restrict__ db_last, int64_t n) {
    constexpr int C = 1 + D + LAP;
    int64_t gp = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float ub = 0.f;
    if (gp < n) {
        float u = head_saved[gp * C];
        float gu[D], gb[D];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            gu[j] = head_saved[gp * C + 1 + j];
            gb[j] = g_bar ? g_bar[gp * D + j] : 0.f;
        }
        float lu = LAP ? head_saved[gp * C + 1 + D] : 0.f;
        float fb = f_bar ? f_bar[gp] : 0.f;
        float lb = (LAP && l_bar) ? l_bar[gp] : 0.f;
        float gub[D], lub;
        if (head == DIGS_HEAD_IDENTITY) {
            ub = fb;
#pragma unroll
            for (int j = 0; j < D; ++j) gub[j] = gb[j];
            lub = lb;
        } else {
            float a = fabsf(u) + 1e-8f;
            float sg = (u > 0.f) ? 1.f : ((u < 0.f) ? -1.f : 0.f);
            float ra = sqrtf(a);
            float g1 = 0.5f / ra;
            float g2 = -0.25f * sg / (a * ra);
            float g3 = 0.375f / (a * a * ra);
            float gg = 0.f, gbg = 0.f;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                gg = fmaf(gu[j], gu[j], gg);
                gbg = fmaf(gb[j], gu[j], gbg);
            }
            ub = scale * (g1 * fb + g2 * gbg + lb * (g2 * lu + g3 * gg));
#pragma unroll
            for (int j = 0; j < D; ++j) gub[j] = scale * (g1 * gb[j] + 2.f * g2 * lb * gu[j]);
            lub = scale * g1 * lb;
        }
        ubar[gp * C] = ub;
#pragma unroll
        for (int j = 0; j < D; ++j) ubar[gp * C + 1 + j] = gub[j];
        if constexpr (LAP) ubar[gp * C + 1 + D] = lub;
    }
    // block reduce u_bar for db_last
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ub += __shfl_xor_sync(0xffffffffu, ub, o);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ub;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        atomicAdd(db_last, t);
    }
}

// K3b: dw_last[k] += sum_p ubar[p] . act(pre_L)[p][k]
template <int D, int LAP, bool FROM_XYZ>
__global__ void k_head_wgrad(PreSrc src, const float* __restrict__ ubar, float* __restrict__ dw_last, int64_t n, int H,
                             int pts_per_block) {
    constexpr int C = 1 + D + LAP;
    extern __shared__ float ub_s[];  // [pts_per_block][C]
    int64_t p0 = (int64_t)blockIdx.y * pts_per_block;
    int np = (int)min((int64_t)pts_per_block, n - p0);
    for (int e = threadIdx.x; e < np * C; e += blockDim.x) ub_s[e] = ubar[p0 * C + e];
    __syncthreads();
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= H) return;
    float acc = 0.f;
    for (int p = 0; p < np; ++p) {
        float z, Z[D > 0 ? D : 1], Q, out[C];
        fetch_pre<D, LAP, FROM_XYZ>(src, p0 + p, k, H, z, Z, Q);
        activate<D, LAP>(z, Z, Q, out);
#pragma unroll
        for (int c = 0; c < C; ++c) acc = fmaf(ub_s[p * C + c], out[c], acc);
    }
    atomicAdd(dw_last + k, acc);
}

// ------------------------------------------------------------------------------------------
// K4: hidden layer reverse (dgrad).  Loader forms the adjoint (zb|Zb|Qb) of layer i from the incoming
// adjoint and pre(i); stores it to G (for wgrad) from the blockIdx.y == 0 column of CTAs; multiplies by W_i.
//   out[c][p][k] = sum_u adj_c[p][u] * W[u][k]
// TOP: incoming adjoint is the rank-1 product ubar[p][c] * w_last[u].
// ------------------------------------------------------------------------------------------
template <int D, int LAP, bool TOP>
__global__ void __launch_bounds__(NTHREADS) k_hidden_bwd(PreSrc pre, const float* __restrict__ inc, /* [C][n][H] */
                                                         const float* __restrict__ ubar, const float* __restrict__ w_last,
                                                         const float* __restrict__ W, float* __restrict__ G,
                                                         float* __restrict__ out, int64_t n, int H, int do_gemm) {
    using Gc = Cfg<D, LAP>;
    constexpr int C = Gc::C, TP = Gc::TP, RT = Gc::RT, RPAD = Gc::RPAD;
    __shared__ __align__(16) float Xs[KS][RPAD];
    __shared__ __align__(16) float Ws[KS][TN];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t p0 = (int64_t)blockIdx.x * TP;
    const int n0 = blockIdx.y * TN;
    const int64_t ps = n * (int64_t)H;
    float acc[RT][4];
#pragma unroll
    for (int r = 0; r < RT; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;

    for (int u0 = 0; u0 < H; u0 += KS) {
        for (int e = tid; e < TP * KS; e += NTHREADS) {
            int p = e / KS, uu = e % KS;
            int64_t gp = p0 + p;
            float o[C];
#pragma unroll
            for (int c = 0; c < C; ++c) o[c] = 0.f;
            if (gp < n) {
                int u = u0 + uu;
                float z, Z[D], Q;
                fetch_pre<D, LAP, false>(pre, gp, u, H, z, Z, Q);
                float hb, Ab[D], Pb = 0.f;
                if constexpr (TOP) {
                    float w = __ldg(w_last + u);
                    hb = ubar[gp * C] * w;
#pragma unroll
                    for (int j = 0; j < D; ++j) Ab[j] = ubar[gp * C + 1 + j] * w;
                    if constexpr (LAP) Pb = ubar[gp * C + 1 + D] * w;
                } else {
                    const float* b = inc + gp * H + u;
                    hb = __ldg(b);
#pragma unroll
                    for (int j = 0; j < D; ++j) Ab[j] = __ldg(b + (1 + j) * ps);
                    if constexpr (LAP) Pb = __ldg(b + (1 + D) * ps);
                }
                activate_adjoint<D, LAP>(z, Z, Q, hb, Ab, Pb, o);
                if (blockIdx.y == 0) {
#pragma unroll
                    for (int c = 0; c < C; ++c) G[c * ps + gp * H + u] = o[c];
                }
            }
#pragma unroll
            for (int c = 0; c < C; ++c) Xs[uu][c * TP + p] = o[c];
        }
        if (do_gemm) {
            for (int e = tid; e < KS * TN / 4; e += NTHREADS) {
                int uu = e / (TN / 4), j4 = e % (TN / 4);
                *reinterpret_cast<float4*>(&Ws[uu][j4 * 4]) =
                    __ldg(reinterpret_cast<const float4*>(W + (int64_t)(u0 + uu) * H + n0 + j4 * 4));
            }
            __syncthreads();
            slab_fma<RT, RPAD>(Xs, Ws, acc, ty * RT, tx * 4);
        }
        __syncthreads();
    }
    if (!do_gemm) return;
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        int row = ty * RT + r;
        int c = row / TP, p = row % TP;
        int64_t gp = p0 + p;
        if (gp < n)
            *reinterpret_cast<float4*>(out + c * ps + gp * H + n0 + tx * 4) =
                make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    }
}

// ------------------------------------------------------------------------------------------
// K5: weight gradient.  dW[u][k] += sum_{c,p} G[c][p][u] * act_c(pre_in)[p][k]   (split-K over points)
// ------------------------------------------------------------------------------------------
constexpr int WG_KP = 8;  // points per slab
template <int D, int LAP, bool FROM_XYZ>
__global__ void __launch_bounds__(NTHREADS) k_wgrad(PreSrc src, const float* __restrict__ G, float* __restrict__ dW,
                                                    int64_t n, int H, int64_t pts_per_split) {
    constexpr int C = 1 + D + LAP;
    constexpr int ROWS = C * WG_KP;
    __shared__ __align__(16) float As[ROWS][TN];
    __shared__ __align__(16) float Bs[ROWS][TN];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int u0 = blockIdx.x * TN, k0 = blockIdx.y * TN;
    const int64_t pb = (int64_t)blockIdx.z * pts_per_split;
    const int64_t pe = min(n, pb + pts_per_split);
    const int64_t ps = n * (int64_t)H;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

    for (int64_t p0 = pb; p0 < pe; p0 += WG_KP) {
        // A: adjoint rows
        for (int e = tid; e < ROWS * (TN / 4); e += NTHREADS) {
            int rr = e / (TN / 4), j4 = e % (TN / 4);
            int c = rr / WG_KP, pp = rr % WG_KP;
            int64_t gp = p0 + pp;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gp < pe) v = __ldg(reinterpret_cast<const float4*>(G + c * ps + gp * H + u0 + j4 * 4));
            *reinterpret_cast<float4*>(&As[rr][j4 * 4]) = v;
        }
        // B: activated inputs
        for (int e = tid; e < WG_KP * TN; e += NTHREADS) {
            int pp = e / TN, k = e % TN;
            int64_t gp = p0 + pp;
            float o[C];
#pragma unroll
            for (int c = 0; c < C; ++c) o[c] = 0.f;
            if (gp < pe) {
                float z, Z[D > 0 ? D : 1], Q;
                fetch_pre<D, LAP, FROM_XYZ>(src, gp, k0 + k, H, z, Z, Q);
                activate<D, LAP>(z, Z, Q, o);
            }
#pragma unroll
            for (int c = 0; c < C; ++c) Bs[c * WG_KP + pp][k] = o[c];
        }
        __syncthreads();
#pragma unroll 8
        for (int rr = 0; rr < ROWS; ++rr) {
            float4 a = *reinterpret_cast<const float4*>(&As[rr][ty * 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[rr][tx * 4]);
            float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i][0] = fmaf(av[i], b.x, acc[i][0]);
                acc[i][1] = fmaf(av[i], b.y, acc[i][1]);
                acc[i][2] = fmaf(av[i], b.z, acc[i][2]);
                acc[i][3] = fmaf(av[i], b.w, acc[i][3]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float* dst = dW + (int64_t)(u0 + ty * 4 + i) * H + k0 + tx * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(dst + j, acc[i][j]);
    }
}

// K6: db[u] += sum_p G[0][p][u]
__global__ void k_colsum(const float* __restrict__ G0, float* __restrict__ db, int64_t n, int H, int pts_per_block) {
    int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= H) return;
    int64_t p0 = (int64_t)blockIdx.y * pts_per_block;
    int64_t p1 = min(n, p0 + pts_per_block);
    float a = 0.f;
    for (int64_t p = p0; p < p1; ++p) a += __ldg(G0 + p * H + u);
    atomicAdd(db + u, a);
}

// ------------------------------------------------------------------------------------------
// K7: layer-0 parameter gradients from the adjoint arriving at layer-0 outputs.
//   dW0[u][j] += sum_p zb x_j + Zb_j ; db0[u] += sum_p zb ; shape_bias_grad[s][u] += sum_{p in s} zb
// ------------------------------------------------------------------------------------------
template <int D, int LAP>
__global__ void k_layer0_bwd(PreSrc src, const float* __restrict__ inc, float* __restrict__ dW0, int w0_stride,
                             float* __restrict__ db0, float* __restrict__ sb_grad, int64_t n, int H,
                             int pts_per_block) {
    int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= H) return;
    const int64_t ps = n * (int64_t)H;
    int64_t p0 = (int64_t)blockIdx.y * pts_per_block;
    int64_t p1 = min(n, p0 + pts_per_block);
    float aw[D], ab = 0.f, asb = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j) aw[j] = 0.f;
    int64_t cur_shape = sb_grad ? (p0 / src.pps) : 0;
    for (int64_t p = p0; p < p1; ++p) {
        float z, Z[D], Q;
        fetch_pre<D, LAP, true>(src, p, u, H, z, Z, Q);
        const float* b = inc + p * H + u;
        float hb = __ldg(b), Ab[D], Pb = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) Ab[j] = __ldg(b + (1 + j) * ps);
        if constexpr (LAP) Pb = __ldg(b + (1 + D) * ps);
        float o[1 + D + LAP];
        activate_adjoint<D, LAP>(z, Z, Q, hb, Ab, Pb, o);
        const float* xp = src.xyz + p * src.xyz_stride;
#pragma unroll
        for (int j = 0; j < D; ++j) aw[j] += fmaf(o[0], __ldg(xp + j), o[1 + j]);
        ab += o[0];
        if (sb_grad) {
            int64_t sh = p / src.pps;
            if (sh != cur_shape) {
                atomicAdd(sb_grad + cur_shape * H + u, asb);
                asb = 0.f;
                cur_shape = sh;
            }
            asb += o[0];
        }
    }
#pragma unroll
    for (int j = 0; j < D; ++j) atomicAdd(dW0 + (int64_t)u * w0_stride + j, aw[j]);
    if (sb_grad) atomicAdd(sb_grad + cur_shape * H + u, asb);
    else atomicAdd(db0 + u, ab);
}

// K8: batched transpose of the hidden weights: Wt[l][k][u] = W_l[u][k]
struct WPtrs { const float* w[DIGS_MAX_LAYERS]; };
__global__ void k_transpose(WPtrs wp, float* __restrict__ Wt, int H) {
    __shared__ float t[32][33];
    const float* W = wp.w[blockIdx.z];
    float* o = Wt + (int64_t)blockIdx.z * H * H;
    int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8) t[j][threadIdx.x] = W[(int64_t)(y0 + j) * H + x];
    __syncthreads();
    x = blockIdx.y * 32 + threadIdx.x;
    y0 = blockIdx.x * 32;
    for (int j = threadIdx.y; j < 32; j += 8) o[(int64_t)(y0 + j) * H + x] = t[threadIdx.x][j];
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static inline int chans(const DigsNet* net, int want_lap) { return 1 + net->d + (want_lap ? 1 : 0); }

size_t saved_bytes(const DigsNet* net, int64_t n, int want_lap) {
    size_t C = chans(net, want_lap);
    size_t planes = (size_t)net->Lh * C * n * net->H * sizeof(float);
    size_t headsv = align_up((size_t)n * C * sizeof(float), 256);
    return align_up(planes, 256) + headsv;
}
static size_t wt_bytes(const DigsNet* net) { return align_up((size_t)net->Lh * net->H * net->H * sizeof(float), 256); }
size_t fwd_ws_bytes(const DigsNet* net, int64_t n, int want_lap) {
    // transposed weights + (inference only) two ping-pong plane sets
    size_t C = chans(net, want_lap);
    return wt_bytes(net) + 2 * align_up((size_t)C * n * net->H * sizeof(float), 256);
}
size_t bwd_ws_bytes(const DigsNet* net, int64_t n, int want_lap) {
    size_t C = chans(net, want_lap);
    size_t plane = align_up((size_t)C * n * net->H * sizeof(float), 256);
    return 3 * plane + align_up((size_t)n * C * sizeof(float), 256);
}
size_t value_ws_bytes(const DigsNet* net, int64_t n) {
    return wt_bytes(net) + 2 * align_up((size_t)n * net->H * sizeof(float), 256);
}
constexpr int64_t GRID_CHUNK = 1 << 16;
size_t grid_ws_bytes(const DigsNet* net) { return value_ws_bytes(net, GRID_CHUNK); }

static int launch_transpose(const DigsNet* net, float* Wt, cudaStream_t st) {
    WPtrs wp;
    for (int l = 0; l < net->Lh; ++l) wp.w[l] = net->W[1 + l];
    dim3 g(net->H / 32, net->H / 32, net->Lh), b(32, 8);
    k_transpose<<<g, b, 0, st>>>(wp, Wt, net->H); count_launch();
    return 0;
}

static PreSrc make_xyz_src(const DigsNet* net, const float* xyz, int64_t xyz_stride, const float* shape_bias,
                           int64_t pps) {
    PreSrc s{};
    s.mode = 1;
    s.din = net->d;
    s.xyz = xyz;
    s.xyz_stride = xyz_stride;
    s.W0 = net->W[0];
    s.w0_stride = net->w0_stride;
    s.b0 = net->b[0];
    s.shape_bias = shape_bias;
    s.pps = pps > 0 ? pps : 1;
    return s;
}
static PreSrc make_plane_src(const float* planes, int64_t n, int H) {
    PreSrc s{};
    s.mode = 0;
    s.planes = planes;
    s.plane_stride = n * (int64_t)H;
    return s;
}

template <int D, int LAP>
static int forward_t(const DigsNet* net, const PreSrc& src0, int64_t n, float* f, float* grad, float* lap,
                     float* planes /* [Lh][C][n][H] or ping-pong */, bool pingpong, float* head_saved, float* Wt,
                     cudaStream_t st) {
    using G = Cfg<D, LAP>;
    constexpr int C = G::C;
    const int H = net->H;
    const size_t plane_set = (size_t)C * n * H;
    dim3 grid((unsigned)((n + G::TP - 1) / G::TP), H / TN);
    for (int l = 1; l <= net->Lh; ++l) {
        float* outp = planes + (pingpong ? ((l & 1) ? 0 : align_up(plane_set * 4, 256) / 4) : (size_t)(l - 1) * plane_set);
        const float* Wl = Wt + (size_t)(l - 1) * H * H;
        if (l == 1) {
            k_hidden_fwd<D, LAP, true><<<grid, NTHREADS, 0, st>>>(src0, Wl, net->b[l], outp, n, H); count_launch();
        } else {
            float* inp = planes + (pingpong ? (((l - 1) & 1) ? 0 : align_up(plane_set * 4, 256) / 4)
                                            : (size_t)(l - 2) * plane_set);
            k_hidden_fwd<D, LAP, false><<<grid, NTHREADS, 0, st>>>(make_plane_src(inp, n, H), Wl, net->b[l], outp, n, H); count_launch();
        }
    }
    int L = net->Lh;
    float* lastp = planes + (pingpong ? ((L & 1) ? 0 : align_up(plane_set * 4, 256) / 4) : (size_t)(L - 1) * plane_set);
    unsigned hb = (unsigned)((n + 7) / 8);
    k_head_fwd<D, LAP, false><<<hb, 256, 0, st>>>(make_plane_src(lastp, n, H), net->W[L + 1], net->b[L + 1], net->head,
                                                  net->radius, net->scale, f, grad, lap, head_saved, n, H); count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("simt forward launch: %s", cudaGetErrorString(e)); return DIGS_ECUDA; }
    return DIGS_OK;
}

int jet_forward(const DigsNet* net, const float* xyz, int64_t n, int64_t xyz_stride, const float* shape_bias,
                int64_t pps, int want_lap, float* f, float* grad, float* lap, void* saved, void* ws, cudaStream_t st) {
    if (n == 0) return DIGS_OK;
    float* Wt = reinterpret_cast<float*>(ws);
    launch_transpose(net, Wt, st);
    PreSrc src0 = make_xyz_src(net, xyz, xyz_stride, shape_bias, pps);
    const int C = chans(net, want_lap);
    float* planes;
    float* headsv = nullptr;
    bool pingpong;
    if (saved) {
        planes = reinterpret_cast<float*>(saved);
        headsv = reinterpret_cast<float*>(reinterpret_cast<char*>(saved) +
                                          align_up((size_t)net->Lh * C * n * net->H * sizeof(float), 256));
        pingpong = false;
    } else {
        planes = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + wt_bytes(net));
        pingpong = true;
    }
    if (net->d == 3 && want_lap) return forward_t<3, 1>(net, src0, n, f, grad, lap, planes, pingpong, headsv, Wt, st);
    if (net->d == 3) return forward_t<3, 0>(net, src0, n, f, grad, lap, planes, pingpong, headsv, Wt, st);
    if (net->d == 2 && want_lap) return forward_t<2, 1>(net, src0, n, f, grad, lap, planes, pingpong, headsv, Wt, st);
    return forward_t<2, 0>(net, src0, n, f, grad, lap, planes, pingpong, headsv, Wt, st);
}

int value_forward(const DigsNet* net, const PreSrc& src0, int64_t n, float* f, void* ws, cudaStream_t st) {
    if (n == 0) return DIGS_OK;
    float* Wt = reinterpret_cast<float*>(ws);
    launch_transpose(net, Wt, st);
    float* planes = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + wt_bytes(net));
    return forward_t<0, 0>(net, src0, n, f, nullptr, nullptr, planes, true, nullptr, Wt, st);
}

int grid_query(const DigsNet* net, const uint16_t* xh, int nx, const uint16_t* yh, int ny, const uint16_t* zh, int nz,
               int iy_begin, int iy_end, const float* shape_bias, float* out, void* ws, cudaStream_t st) {
    (void)ny;
    float* Wt = reinterpret_cast<float*>(ws);
    launch_transpose(net, Wt, st);
    float* planes = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + wt_bytes(net));
    int64_t total = (int64_t)(iy_end - iy_begin) * nx * nz;
    int64_t base = (int64_t)iy_begin * nx * nz;
    for (int64_t off = 0; off < total; off += GRID_CHUNK) {
        int64_t m = total - off < GRID_CHUNK ? total - off : GRID_CHUNK;
        PreSrc s{};
        s.mode = 2;
        s.din = 3;
        s.W0 = net->W[0];
        s.w0_stride = net->w0_stride;
        s.b0 = net->b[0];
        s.shape_bias = shape_bias;
        s.pps = (int64_t)1 << 62;  // one shape
        s.ax = reinterpret_cast<const __half*>(xh);
        s.ay = reinterpret_cast<const __half*>(yh);
        s.az = reinterpret_cast<const __half*>(zh);
        s.nx = nx;
        s.nz = nz;
        s.grid_base = base + off;
        int rc = forward_t<0, 0>(net, s, m, out + off, nullptr, nullptr, planes, true, nullptr, Wt, st);
        if (rc) return rc;
    }
    return DIGS_OK;
}

template <int D, int LAP>
static int backward_t(const DigsNet* net, const PreSrc& src0, int64_t n, const float* f_bar, const float* g_bar,
                      const float* l_bar, const float* saved_planes, const float* head_saved, float* ws,
                      const DigsGrads* gr, float* sb_grad, cudaStream_t st) {
    using G = Cfg<D, LAP>;
    constexpr int C = G::C;
    const int H = net->H, L = net->Lh;
    const size_t plane_set = (size_t)C * n * H;
    const size_t plane_al = align_up(plane_set * 4, 256) / 4;
    float* Gbuf = ws;
    float* inc[2] = {ws + plane_al, ws + 2 * plane_al};
    float* ubar = ws + 3 * plane_al;

    k_head_seed<D, LAP><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(head_saved, f_bar, g_bar, l_bar, net->head,
                                                                      net->scale, ubar, gr->db[L + 1], n); count_launch();
    {
        const int ppb = 64;
        int bt = H < 256 ? H : 256;
        dim3 g(H / bt, (unsigned)((n + ppb - 1) / ppb));
        k_head_wgrad<D, LAP, false><<<g, bt, ppb * C * sizeof(float), st>>>(
            make_plane_src(saved_planes + (size_t)(L - 1) * plane_set, n, H), ubar, gr->dW[L + 1], n, H, ppb); count_launch();
    }
    dim3 grid((unsigned)((n + G::TP - 1) / G::TP), H / TN);
    int nsplit = (int)((n + 1023) / 1024);
    if (nsplit > 64) nsplit = 64;
    if (nsplit < 1) nsplit = 1;
    int64_t pps_split = ((n + nsplit - 1) / nsplit + WG_KP - 1) / WG_KP * WG_KP;
    nsplit = (int)((n + pps_split - 1) / pps_split);
    dim3 wgrid(H / TN, H / TN, nsplit);
    for (int l = L; l >= 1; --l) {
        PreSrc pre = make_plane_src(saved_planes + (size_t)(l - 1) * plane_set, n, H);
        float* outp = inc[l & 1];
        if (l == L) {
            k_hidden_bwd<D, LAP, true><<<grid, NTHREADS, 0, st>>>(pre, nullptr, ubar, net->W[L + 1], net->W[l], Gbuf,
                                                                  outp, n, H, 1);
        } else {
            k_hidden_bwd<D, LAP, false><<<grid, NTHREADS, 0, st>>>(pre, inc[(l + 1) & 1], nullptr, nullptr, net->W[l],
                                                                   Gbuf, outp, n, H, 1);
        }
        if (l == 1) {
            k_wgrad<D, LAP, true><<<wgrid, NTHREADS, 0, st>>>(src0, Gbuf, gr->dW[l], n, H, pps_split);
        } else {
            k_wgrad<D, LAP, false><<<wgrid, NTHREADS, 0, st>>>(
                make_plane_src(saved_planes + (size_t)(l - 2) * plane_set, n, H), Gbuf, gr->dW[l], n, H, pps_split);
        }
        count_launch(2);
        {
            int bt = H < 256 ? H : 256;
            dim3 g(H / bt, (unsigned)((n + 255) / 256));
            k_colsum<<<g, bt, 0, st>>>(Gbuf, gr->db[l], n, H, 256); count_launch();
        }
    }
    {
        int bt = H < 128 ? H : 128;
        const int ppb = 128;
        dim3 g(H / bt, (unsigned)((n + ppb - 1) / ppb));
        k_layer0_bwd<D, LAP><<<g, bt, 0, st>>>(src0, inc[1], gr->dW[0], net->w0_stride, gr->db[0], sb_grad, n, H, ppb); count_launch();
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("simt backward launch: %s", cudaGetErrorString(e)); return DIGS_ECUDA; }
    return DIGS_OK;
}

int jet_backward(const DigsNet* net, const float* xyz, int64_t n, int64_t xyz_stride, const float* shape_bias,
                 int64_t pps, int want_lap, const float* f_bar, const float* g_bar, const float* l_bar,
                 const void* saved, void* ws, const DigsGrads* grads, float* shape_bias_grad, cudaStream_t st) {
    if (n == 0) return DIGS_OK;
    PreSrc src0 = make_xyz_src(net, xyz, xyz_stride, shape_bias, pps);
    const int C = chans(net, want_lap);
    const float* planes = reinterpret_cast<const float*>(saved);
    const float* headsv = reinterpret_cast<const float*>(
        reinterpret_cast<const char*>(saved) + align_up((size_t)net->Lh * C * n * net->H * sizeof(float), 256));
    float* w = reinterpret_cast<float*>(ws);
    if (net->d == 3 && want_lap)
        return backward_t<3, 1>(net, src0, n, f_bar, g_bar, l_bar, planes, headsv, w, grads, shape_bias_grad, st);
    if (net->d == 3)
        return backward_t<3, 0>(net, src0, n, f_bar, g_bar, l_bar, planes, headsv, w, grads, shape_bias_grad, st);
    if (net->d == 2 && want_lap)
        return backward_t<2, 1>(net, src0, n, f_bar, g_bar, l_bar, planes, headsv, w, grads, shape_bias_grad, st);
    return backward_t<2, 0>(net, src0, n, f_bar, g_bar, l_bar, planes, headsv, w, grads, shape_bias_grad, st);
}

int head_seed(const DigsNet* net, int64_t n, int want_lap, const float* head_saved, const float* f_bar,
              const float* g_bar, const float* l_bar, float* ubar, float* db_last, cudaStream_t st) {
    unsigned blocks = (unsigned)((n + 255) / 256);
    if (net->d == 3 && want_lap)
        k_head_seed<3, 1><<<blocks, 256, 0, st>>>(head_saved, f_bar, g_bar, l_bar, net->head, net->scale, ubar, db_last, n);
    else if (net->d == 3)
        k_head_seed<3, 0><<<blocks, 256, 0, st>>>(head_saved, f_bar, g_bar, l_bar, net->head, net->scale, ubar, db_last, n);
    else if (want_lap)
        k_head_seed<2, 1><<<blocks, 256, 0, st>>>(head_saved, f_bar, g_bar, l_bar, net->head, net->scale, ubar, db_last, n);
    else
        k_head_seed<2, 0><<<blocks, 256, 0, st>>>(head_saved, f_bar, g_bar, l_bar, net->head, net->scale, ubar, db_last, n);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("head_seed launch: %s", cudaGetErrorString(e)); return DIGS_ECUDA; }
    return DIGS_OK;
}

}  // namespace simt
}  // namespace digs
