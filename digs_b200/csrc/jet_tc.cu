// tcgen05 / TMEM / TMA implementation of the forward jet and the grid query (DIGS_MODE_BF16X2, DIGS_MODE_BF16).
//
// Reference behaviour reproduced: FCBlock.forward (models/DiGS.py:122-134) for a whole network per point tile,
// plus the gradient / Laplacian channels the reference obtains from utils.gradient (utils/utils.py:108-117,
// models/losses.py:72-76,100-106), and the chunk loop of utils.implicit2mesh (utils/utils.py:343-348).
//
// One persistent CTA per SM walks point tiles.  A tile is T points x C jet channels = N <= 128 operand columns
// (column = channel*T + point).  For every hidden layer the CTA computes, on the tensor cores,
//
//        D[unit][column] = sum_k (30 W)[unit][k] * X[column][k]          M = H units (H/128 UMMA tiles), N columns, K = H
//
// i.e. the layer is evaluated TRANSPOSED: units sit on TMEM lanes, jet channels of a point sit in TMEM columns of
// the same lane, so the sin/cos + jet recurrences of one (unit, point) need only the registers of one thread.
//   A = 30*W   bf16 (hi [+ lo]) K-major SWIZZLE_128B tiles of 128 units x 64 k, streamed from L2 by TMA through a
//              5-stage mbarrier ring (weights never live in shared memory as a whole);
//   B = X      bf16 (hi [+ lo]) MN-major SWIZZLE_128B, written in place by the epilogue warps as the NEXT layer's
//              operand: a thread owns one unit = one k row and stores 8 consecutive points as one 16-byte chunk;
//   D          fp32 in TMEM, H/128 tiles of N columns; read back with tcgen05.ld.32x32b.
// DIGS_MODE_BF16X2 runs three MMA passes per product (hi*hi + hi*lo + lo*hi, fp32 accumulate) which keeps ~16
// mantissa bits per operand; DIGS_MODE_BF16 runs the hi*hi pass only.  Layer 0 (K = d) and the last layer (N = 1)
// are fp32 FMAs / warp reductions from registers and never touch the tensor cores.
//
// Warp roles (320 threads): warps 0-7 epilogue (TMEM -> registers -> jet math -> next operand / outputs),
// warp 8 TMA producer, warp 9 MMA issuer + TMEM owner.
#include "digs_internal.h"
#include "tc_common.cuh"
#include <cuda.h>

namespace digs {
namespace tcpath {

using namespace digs::tc;

constexpr int NTHREADS = 320;
constexpr int EPI_THREADS = 256;
constexpr int STAGES = 5;
constexpr int STAGE_BYTES = 128 * 64 * 2;  // one A tile: 128 units x 64 k, bf16

template <int D, int LAP, int H_, int PASSES>
struct Cfg {
    static constexpr int C = 1 + D + LAP;
    static constexpr int H = H_;
    static constexpr int N = (H_ > 256) ? 64 : 128;  // UMMA N (operand columns); 64 keeps X within shared memory at H=512
    static constexpr int T = (N / C) / 8 * 8;         // points per tile (multiple of 8: one 16 B chunk of a k row)
    static constexpr int MT = H / 128;                // UMMA M tiles
    static constexpr int NG = N / 64;                 // 64-column groups of the MN-major operand
    static constexpr int SN = 1024;                   // byte stride between column groups   (descriptor LBO)
    static constexpr int SK = NG * 1024;              // byte stride between 8-deep k groups (descriptor SBO)
    static constexpr int B_PART = (H / 8) * SK;       // bytes of one operand part (hi or lo)
    static constexpr int NPARTS = (PASSES == 3) ? 2 : 1;
    static constexpr int KSLABS = H / 64;
    static constexpr int TMEM_COLS = (MT * N <= 128) ? 128 : ((MT * N <= 256) ? 256 : 512);
    static constexpr int UPW = (MT >= 2) ? MT / 2 : 1;   // 32-unit blocks per epilogue warp
    static constexpr int GSPLIT = (MT == 1) ? 2 : 1;     // point groups split between the two warp halves
    static constexpr int NSLOT = (MT == 1) ? 4 : 8;      // head partial-sum slots
    static constexpr int OFF_RING = B_PART * NPARTS;
    static constexpr int OFF_XYZ = OFF_RING + STAGES * STAGE_BYTES;
    static constexpr int OFF_HEAD = OFF_XYZ + 128 * 3 * 4;
    static constexpr int OFF_BAR = OFF_HEAD + NSLOT * N * 4;
    static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;   // + alignment slack
    static_assert(H % 128 == 0 && H <= 512, "H must be 128, 256, 384 or 512");
    static_assert(MT * N <= 512, "accumulators exceed TMEM");
};

struct Args {
    PreSrc src;               // layer-0 input (xyz or grid), W0/b0 fields unused here
    int64_t n;                // points
    int Lh;
    const float* w0_30;       // [H][4]: 30*W0[u][0..2], 30*b0[u]
    const float* bias30;      // [Lh][H]: 30*b_l[u], l = 1..Lh
    const float* w_last;      // [H]
    const float* b_last;      // [1]
    int head;
    float radius, scale;
    float* f;                 // (n)
    float* grad;              // (n, D)
    float* lap;               // (n)
    float* planes;            // NULL or [Lh][C][n][H] pre-activations (z, Z_k, Q) for the reverse pass
    float* head_saved;        // NULL or [n][C]
};

// sin/cos of a phase of any magnitude: two-constant Cody-Waite reduction to [-pi, pi], then MUFU
__device__ __forceinline__ void sincos_reduced(float t, float& s, float& c) {
    const float magic = 12582912.f;  // 1.5 * 2^23: round-to-nearest-integer trick
    float k = fmaf(t, 0.15915494309189535f, magic) - magic;
    float r = fmaf(k, -6.2831854820251465f, t);
    r = fmaf(k, 1.7484556000744883e-7f, r);
    s = __sinf(r);
    c = __cosf(r);
}

// value t = 30 z, Zp = 30 Z_k, Qp = 30 Q  ->  next-layer inputs (h, A_k, P)
template <int D, int LAP>
__device__ __forceinline__ void act_scaled(float t, const float* Zp, float Qp, float* o) {
    float s, c;
    sincos_reduced(t, s, c);
    o[0] = s;
    if constexpr (D > 0) {
        float S2 = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) {
            o[1 + j] = c * Zp[j];
            S2 = fmaf(Zp[j], Zp[j], S2);
        }
        if constexpr (LAP) o[1 + D] = fmaf(c, Qp, -s * S2);
    }
}

// sum over the 32 lanes of NV (multiple of 8) per-lane values; afterwards lane j holds total of value j for the
// first 32-blocks, and lane (j % 8) of every 8-lane group holds value j of a trailing 8-block.
template <int NV>
__device__ __forceinline__ void warp_reduce_scatter(float* v, int lane) {
    constexpr int FULL = NV / 32;
#pragma unroll
    for (int b = 0; b < FULL; ++b) {
        float* x = v + 32 * b;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int j = 0; j < off; ++j) {
                float send = up ? x[j] : x[j + off];
                float keep = up ? x[j + off] : x[j];
                x[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
    }
    constexpr int REM = (NV % 32) / 8;
#pragma unroll
    for (int b = 0; b < REM; ++b) {
        float* x = v + 32 * FULL + 8 * b;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            x[j] += __shfl_xor_sync(0xffffffffu, x[j], 16);
            x[j] += __shfl_xor_sync(0xffffffffu, x[j], 8);
        }
#pragma unroll
        for (int off = 4; off >= 1; off >>= 1) {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int j = 0; j < off; ++j) {
                float send = up ? x[j] : x[j + off];
                float keep = up ? x[j + off] : x[j];
                x[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
    }
}

template <int D, int LAP, int H, int PASSES>
__global__ void __launch_bounds__(NTHREADS, 1) k_tc_forward(const __grid_constant__ CUtensorMap tmapW, Args a) {
    using G = Cfg<D, LAP, H, PASSES>;
    constexpr int C = G::C, T = G::T, N = G::N, MT = G::MT;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sB = smem;
    uint8_t* sRing = smem + G::OFF_RING;
    float* sXyz = reinterpret_cast<float*>(smem + G::OFF_XYZ);
    float* sHead = reinterpret_cast<float*>(smem + G::OFF_HEAD);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G::OFF_BAR);
    uint64_t* w_full = bars;                 // [STAGES]
    uint64_t* w_empty = bars + STAGES;       // [STAGES]
    uint64_t* b_ready = bars + 2 * STAGES;   // epilogue -> MMA: next operand complete (256 arrivals)
    uint64_t* acc_ready = bars + 2 * STAGES + 1;  // MMA -> epilogue: accumulators complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        mbar_init(b_ready, EPI_THREADS);
        mbar_init(acc_ready, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmapW);
    }
    if (warp == 9) tmem_alloc<G::TMEM_COLS>(tmem_slot);
    for (int i = tid; i < G::OFF_RING / 16; i += NTHREADS) reinterpret_cast<uint4*>(sB)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int64_t num_tiles = (a.n + T - 1) / T;
    const int Lh = a.Lh;

    if (warp == 8) {
        // ================= TMA producer: stream 30*W tiles (hi, lo) in the order the MMA warp consumes them ========
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int l = 1; l <= Lh; ++l)
                    for (int ks = 0; ks < G::KSLABS; ++ks)
                        for (int m = 0; m < MT; ++m)
                            for (int part = 0; part < G::NPARTS; ++part) {
                                mbar_wait(&w_empty[stage], phase ^ 1);
                                mbar_arrive_expect_tx(&w_full[stage], STAGE_BYTES);
                                tma_load_2d(sRing + stage * STAGE_BYTES, &tmapW, &w_full[stage], ks * 64,
                                            (part * Lh + (l - 1)) * H + m * 128);
                                if (++stage == STAGES) { stage = 0; phase ^= 1; }
                            }
            }
        }
    } else if (warp == 9) {
        // ================= MMA issuer ============================================================================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, N, /*A K-major*/ 0, /*B MN-major*/ 1);
            const uint32_t sB_hi = smem_u32(sB), sB_lo = smem_u32(sB) + G::B_PART;
            int stage = 0;
            uint32_t phase = 0, pb = 0;
            for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int l = 1; l <= Lh; ++l) {
                    mbar_wait(b_ready, pb);
                    pb ^= 1;
                    tc_fence_after();
                    for (int ks = 0; ks < G::KSLABS; ++ks)
                        for (int m = 0; m < MT; ++m) {
                            const uint32_t d_tmem = tmem_base + m * N;
                            // --- A_hi tile: hi*hi (+ hi*lo)
                            mbar_wait(&w_full[stage], phase);
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
                                // --- A_lo tile: lo*hi
                                mbar_wait(&w_full[stage], phase);
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
                        }
                    umma_commit(acc_ready);
                }
            }
        }
    } else {
        // ================= epilogue warps 0..7 ===================================================================
        const int q = warp & 3;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int slot = (MT == 1) ? q : warp;
        const int g0 = (MT == 1) ? (warp >> 2) : 0;
        uint32_t pa = 0;
        // per-unit constants
        int units[G::UPW];
        float w0x[G::UPW], w0y[G::UPW], w0z[G::UPW], b0[G::UPW], wl[G::UPW];
#pragma unroll
        for (int ui = 0; ui < G::UPW; ++ui) {
            int mt = (MT >= 2) ? ((warp >> 2) + 2 * ui) : 0;
            int u = mt * 128 + q * 32 + lane;
            units[ui] = u;
            float4 w = __ldg(reinterpret_cast<const float4*>(a.w0_30) + u);
            w0x[ui] = w.x; w0y[ui] = w.y; w0z[ui] = w.z; b0[ui] = w.w;
            wl[ui] = __ldg(a.w_last + u);
        }
        const int64_t ps = a.n * (int64_t)H;

        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int64_t p0 = tile * T;
            // ---- stage the tile's coordinates -------------------------------------------------------------------
            for (int e = tid; e < T; e += EPI_THREADS) {
                int64_t gp = p0 + e;
                if (gp >= a.n) gp = a.n - 1;
                float x, y, z;
                if (a.src.mode == 2) {
                    int64_t g = a.src.grid_base + gp;
                    int iz = (int)(g % a.src.nz);
                    int64_t t2 = g / a.src.nz;
                    int ix = (int)(t2 % a.src.nx);
                    int iy = (int)(t2 / a.src.nx);
                    x = __half2float(a.src.ax[ix]);
                    y = __half2float(a.src.ay[iy]);
                    z = __half2float(a.src.az[iz]);
                } else {
                    const float* xp = a.src.xyz + gp * a.src.xyz_stride;
                    x = __ldg(xp);
                    y = __ldg(xp + 1);
                    z = (a.src.din > 2) ? __ldg(xp + 2) : 0.f;
                }
                sXyz[e * 3] = x; sXyz[e * 3 + 1] = y; sXyz[e * 3 + 2] = z;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");

            for (int l = 0; l <= Lh; ++l) {
                if (l > 0) {
                    mbar_wait(acc_ready, pa);
                    pa ^= 1;
                    tc_fence_after();
                }
                const bool last = (l == Lh);
                for (int g = g0; g < T / 8; g += G::GSPLIT) {
                    float part[C * 8];
                    if (last) {
#pragma unroll
                        for (int i = 0; i < C * 8; ++i) part[i] = 0.f;
                    }
#pragma unroll
                    for (int ui = 0; ui < G::UPW; ++ui) {
                        const int u = units[ui];
                        const int mt = u >> 7;
                        float acc[C][8];
                        if (l == 0) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float* xp = sXyz + (g * 8 + i) * 3;
                                float bz = b0[ui];
                                if (a.src.shape_bias) {
                                    int64_t gp = p0 + g * 8 + i;
                                    if (gp >= a.n) gp = a.n - 1;
                                    bz = 30.f * __ldg(a.src.shape_bias + (gp / a.src.pps) * H + u);
                                }
                                acc[0][i] = fmaf(w0z[ui], xp[2], fmaf(w0y[ui], xp[1], fmaf(w0x[ui], xp[0], bz)));
                                if constexpr (D > 0) {
                                    acc[1][i] = w0x[ui];
                                    if constexpr (D > 1) acc[2][i] = w0y[ui];
                                    if constexpr (D > 2) acc[3][i] = w0z[ui];
                                    if constexpr (LAP) acc[1 + D][i] = 0.f;
                                }
                            }
                        } else {
#pragma unroll
                            for (int c = 0; c < C; ++c)
                                tmem_ld8(tmem_base + lane_base + mt * N + c * T + g * 8, acc[c]);
                            tmem_ld_wait();
                            float bl = __ldg(a.bias30 + (l - 1) * H + u);
#pragma unroll
                            for (int i = 0; i < 8; ++i) acc[0][i] += bl;
                            if (a.planes) {
                                // pre-activations for the reverse pass, unscaled: [l-1][c][point][unit]
                                const float inv = 1.f / 30.f;
#pragma unroll
                                for (int c = 0; c < C; ++c)
#pragma unroll
                                    for (int i = 0; i < 8; ++i) {
                                        int64_t gp = p0 + g * 8 + i;
                                        if (gp < a.n)
                                            a.planes[((int64_t)(l - 1) * C + c) * ps + gp * H + u] = acc[c][i] * inv;
                                    }
                            }
                        }
                        // ---- jet activation, in place: acc[c][i] -> next-layer inputs --------------------------------
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float Zp[D > 0 ? D : 1], o[C];
#pragma unroll
                            for (int j = 0; j < D; ++j) Zp[j] = acc[1 + j][i];
                            act_scaled<D, LAP>(acc[0][i], Zp, LAP ? acc[C - 1][i] : 0.f, o);
#pragma unroll
                            for (int c = 0; c < C; ++c) acc[c][i] = o[c];
                        }
                        if (!last) {
                            // ---- next operand: row k = u, 8 consecutive columns = one 16 B chunk, SWIZZLE_128B --------
                            const uint32_t rowoff = (uint32_t)(u >> 3) * G::SK + (uint32_t)(u & 7) * 128;
#pragma unroll
                            for (int c = 0; c < C; ++c) {
                                const int ncol = c * T + g * 8;
                                const uint32_t off = rowoff + (uint32_t)(ncol >> 6) * G::SN +
                                                     ((((uint32_t)(ncol & 63) >> 3) ^ (uint32_t)(u & 7)) << 4);
                                uint4 hi, lo;
                                split2(acc[c][0], acc[c][1], hi.x, lo.x);
                                split2(acc[c][2], acc[c][3], hi.y, lo.y);
                                split2(acc[c][4], acc[c][5], hi.z, lo.z);
                                split2(acc[c][6], acc[c][7], hi.w, lo.w);
                                *reinterpret_cast<uint4*>(sB + off) = hi;
                                if constexpr (PASSES == 3) *reinterpret_cast<uint4*>(sB + G::B_PART + off) = lo;
                            }
                        } else {
#pragma unroll
                            for (int c = 0; c < C; ++c)
#pragma unroll
                                for (int i = 0; i < 8; ++i) part[c * 8 + i] = fmaf(acc[c][i], wl[ui], part[c * 8 + i]);
                        }
                    }
                    if (last) {
                        warp_reduce_scatter<C * 8>(part, lane);
                        // lane j holds value j of the first 32; lanes (j%8) hold value j of each trailing 8-block
                        constexpr int FULL = (C * 8) / 32, REM = ((C * 8) % 32) / 8;
#pragma unroll
                        for (int b = 0; b < FULL; ++b) {
                            int v = 32 * b + lane;
                            sHead[slot * N + (v >> 3) * T + g * 8 + (v & 7)] = part[32 * b];
                        }
#pragma unroll
                        for (int b = 0; b < REM; ++b) {
                            if (lane < 8) {
                                int v = 32 * FULL + 8 * b + lane;
                                sHead[slot * N + (v >> 3) * T + g * 8 + (v & 7)] = part[32 * FULL + 8 * b];
                            }
                        }
                    }
                }
                if (!last) {
                    fence_proxy_async();
                    tc_fence_before();
                    mbar_arrive(b_ready);
                }
            }
            // ---- last layer sum over units + output head -------------------------------------------------------------
            asm volatile("bar.sync 1, 256;" ::: "memory");
            for (int e = tid; e < T; e += EPI_THREADS) {
                int64_t gp = p0 + e;
                if (gp < a.n) {
                    float v[C];
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        float s = 0.f;
#pragma unroll
                        for (int w = 0; w < G::NSLOT; ++w) s += sHead[w * N + c * T + e];
                        v[c] = s;
                    }
                    float u = v[0] + __ldg(a.b_last);
                    if (a.head_saved) {
                        a.head_saved[gp * C] = u;
#pragma unroll
                        for (int c = 1; c < C; ++c) a.head_saved[gp * C + c] = v[c];
                    }
                    if (a.head == DIGS_HEAD_IDENTITY) {
                        a.f[gp] = u;
                        if constexpr (D > 0) {
#pragma unroll
                            for (int j = 0; j < D; ++j) a.grad[gp * D + j] = v[1 + j];
                            if constexpr (LAP) a.lap[gp] = v[1 + D];
                        }
                    } else {
                        float aa = fabsf(u) + 1e-8f;
                        float sg = (u > 0.f) ? 1.f : ((u < 0.f) ? -1.f : 0.f);
                        float ra = sqrtf(aa);
                        a.f[gp] = a.scale * (sg * ra - a.radius);
                        if constexpr (D > 0) {
                            float g1 = 0.5f / ra, gg = 0.f;
#pragma unroll
                            for (int j = 0; j < D; ++j) {
                                a.grad[gp * D + j] = a.scale * g1 * v[1 + j];
                                gg = fmaf(v[1 + j], v[1 + j], gg);
                            }
                            if constexpr (LAP) {
                                float g2 = -0.25f * sg / (aa * ra);
                                a.lap[gp] = a.scale * (g1 * v[1 + D] + g2 * gg);
                            }
                        }
                    }
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");   // sXyz / sHead are rewritten by the next tile
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc<G::TMEM_COLS>(tmem_base);
}

// ---- weight preparation: 30*W -> bf16 hi/lo, 30*b, 30*W0 -----------------------------------------------------------
struct PrepPtrs {
    const float* W[DIGS_MAX_LAYERS];
    const float* b[DIGS_MAX_LAYERS];
};
__global__ void k_tc_prep(PrepPtrs p, int Lh, int H, int d, int w0_stride, __nv_bfloat16* Whl /*[2][Lh*H][H]*/,
                          float* bias30 /*[Lh][H]*/, float* w0_30 /*[H][4]*/) {
    const int64_t per = (int64_t)H * H;
    const int64_t total = (int64_t)Lh * per;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int l = (int)(e / per);
        float w = 30.f * p.W[1 + l][e - (int64_t)l * per];
        __nv_bfloat16 hi = __float2bfloat16_rn(w);
        Whl[e] = hi;
        Whl[total + e] = __float2bfloat16_rn(w - __bfloat162float(hi));
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
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn get_encode() {
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

static size_t whl_bytes(const DigsNet* net) { return align_up((size_t)2 * net->Lh * net->H * net->H * 2, 256); }
static size_t bias_bytes(const DigsNet* net) { return align_up((size_t)net->Lh * net->H * 4, 256); }
static size_t w0_bytes(const DigsNet* net) { return align_up((size_t)net->H * 16, 256); }
size_t prep_ws_bytes(const DigsNet* net) { return whl_bytes(net) + bias_bytes(net) + w0_bytes(net); }

static int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    }
    return n;
}

template <int D, int LAP, int H, int PASSES>
static int launch(const CUtensorMap& tmap, const Args& a, cudaStream_t st) {
    using G = Cfg<D, LAP, H, PASSES>;
    auto kern = k_tc_forward<D, LAP, H, PASSES>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES);
        if (e != cudaSuccess) { set_error("tc forward: smem attribute (%d B): %s", G::SMEM_BYTES, cudaGetErrorString(e)); return DIGS_ECUDA; }
        attr_set = true;
    }
    int64_t tiles = (a.n + G::T - 1) / G::T;
    int grid = (int)(tiles < num_sms() ? tiles : num_sms());
    kern<<<grid, NTHREADS, G::SMEM_BYTES, st>>>(tmap, a);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("tc forward launch: %s", cudaGetErrorString(e)); return DIGS_ECUDA; }
    return DIGS_OK;
}

template <int D, int LAP, int PASSES>
static int launch_h(const DigsNet* net, const CUtensorMap& tmap, const Args& a, cudaStream_t st) {
    switch (net->H) {
        case 128: return launch<D, LAP, 128, PASSES>(tmap, a, st);
        case 256: return launch<D, LAP, 256, PASSES>(tmap, a, st);
        case 512: return launch<D, LAP, 512, PASSES>(tmap, a, st);
    }
    set_error("tc forward: H=%d not supported (128, 256, 512)", net->H);
    return DIGS_ENOTIMPL;
}

// Forward jet / value query on the tensor cores.  want_jet = 0 -> value channel only (C = 1).
int forward(const DigsNet* net, const PreSrc& src, int64_t n, int want_jet, int want_lap, float* f, float* grad,
            float* lap, float* planes, float* head_saved, void* ws, int passes, cudaStream_t st) {
    if (n == 0) return DIGS_OK;
    EncodeFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return DIGS_ECUDA; }
    char* w = reinterpret_cast<char*>(ws);
    __nv_bfloat16* Whl = reinterpret_cast<__nv_bfloat16*>(w);
    float* bias30 = reinterpret_cast<float*>(w + whl_bytes(net));
    float* w0_30 = reinterpret_cast<float*>(w + whl_bytes(net) + bias_bytes(net));
    PrepPtrs pp;
    for (int l = 0; l <= net->Lh + 1; ++l) { pp.W[l] = net->W[l]; pp.b[l] = net->b[l]; }
    k_tc_prep<<<148 * 4, 256, 0, st>>>(pp, net->Lh, net->H, net->d, net->w0_stride, Whl, bias30, w0_30);
    count_launch();
    CUtensorMap tmap;
    cuuint64_t gdim[2] = {(cuuint64_t)net->H, (cuuint64_t)2 * net->Lh * net->H};
    cuuint64_t gstr[1] = {(cuuint64_t)net->H * 2};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Whl, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return DIGS_ECUDA; }
    Args a{};
    a.src = src;
    a.n = n;
    a.Lh = net->Lh;
    a.w0_30 = w0_30;
    a.bias30 = bias30;
    a.w_last = net->W[net->Lh + 1];
    a.b_last = net->b[net->Lh + 1];
    a.head = net->head;
    a.radius = net->radius;
    a.scale = net->scale;
    a.f = f; a.grad = grad; a.lap = lap; a.planes = planes; a.head_saved = head_saved;
    if (!want_jet) return (passes == 3) ? launch_h<0, 0, 3>(net, tmap, a, st) : launch_h<0, 0, 1>(net, tmap, a, st);
    if (net->d == 3 && want_lap) return (passes == 3) ? launch_h<3, 1, 3>(net, tmap, a, st) : launch_h<3, 1, 1>(net, tmap, a, st);
    if (net->d == 3) return (passes == 3) ? launch_h<3, 0, 3>(net, tmap, a, st) : launch_h<3, 0, 1>(net, tmap, a, st);
    if (want_lap) return (passes == 3) ? launch_h<2, 1, 3>(net, tmap, a, st) : launch_h<2, 1, 1>(net, tmap, a, st);
    return (passes == 3) ? launch_h<2, 0, 3>(net, tmap, a, st) : launch_h<2, 0, 1>(net, tmap, a, st);
}

}  // namespace tcpath
}  // namespace digs
