// Pieces shared by the tcgen05 jet kernels (jet_tc.cu: one CTA per tile; jet_tc_pair.cu: CTA pairs, cta_group::2).
#pragma once
#include "digs_internal.h"
#include "tc_common.cuh"

#ifdef DIGS_DEBUG
#include <cassert>
#define DIGS_ASSERT(cond) assert(cond)      // make DEBUG=1: device-side bounds checks (compute-sanitizer is closed on the GPU pool)
#else
#define DIGS_ASSERT(cond) ((void)0)
#endif

namespace digs {
namespace tcpath {

using namespace digs::tc;

constexpr int EPI_WARPS = 16;
constexpr int NTHREADS = EPI_WARPS * 32 + 64;
constexpr int STAGES = 5;
constexpr int STAGE_BYTES = 128 * 64 * 2;  // one A tile: 128 rows x 64 k, bf16

struct Args {
    PreSrc src;               // layer-0 input (points or grid indices)
    int64_t n, n_pad;
    int Lh;
    const float* w0_30;       // [H][4]: 30*W0[u][0..2], 30*b0[u]
    const float* bias30;      // [Lh][H]: 30*b_l[u], l = 1..Lh
    const float* w_last;      // [H]
    const float* b_last;      // [1]
    int head;
    float radius, scale;
    float* f;                 // forward outputs
    float* grad;
    float* lap;
    float* planes;            // forward: written (or NULL); reverse: read
    __nv_bfloat16* actP;      // forward: written (or NULL)
    float* head_saved;        // forward: written (or NULL)
    const float* ubar;        // reverse: [n][C] head adjoints (u_bar | gu_bar_k | lu_bar)
    __nv_bfloat16* GP_words;  // reverse: written
    float* db[DIGS_MAX_LAYERS];  // reverse: bias gradients, layers 0..Lh (accumulated)
    float* dW0;               // reverse: (H, w0_stride), first d columns accumulated
    int w0_stride;
    float* dw_last;           // reverse: (H)
    float* sb_grad;           // reverse: (n_shapes, H) or NULL
    int acc_smem;             // reverse: per-CTA shared-memory accumulators for the per-unit gradients (they fit)
};

template <int D, int LAP, int H_, int PASSES>
struct Cfg {
    static constexpr int C = 1 + D + LAP;
    static constexpr int H = H_;
    // One tile context per CTA.  (Two 64-column contexts ping-ponging on the MMA warp were measured SLOWER: a
    // tcgen05.mma with N = 64 re-reads the 4 KB A tile from shared memory for half the math, and the operand fetch,
    // not the tensor pipe, becomes the limit.)
    static constexpr int N = (H_ > 256) ? 64 : 128;   // UMMA N = operand columns per tile (64 keeps X in smem at H=512)
    static constexpr int GP = (C == 1) ? 16 : 4;      // points per epilogue group (one tcgen05.ld per channel)
    static constexpr int T = (N / C) / GP * GP;       // points per tile (whole groups)
    static constexpr int NGRP = T / GP;
    static constexpr int MT = H / 128;                // UMMA M tiles
    // Operand phases: the epilogue finishes the units of the first MT/2 M tiles (= the first half of the next layer's K
    // range) before it touches the rest, so the MMA warp starts the next layer on those K slabs while the epilogue
    // is still working on the second half.  Accumulators are double-buffered in TMEM for that.
    static constexpr int NPH = (MT >= 2) ? 2 : 1;
    static constexpr int MTP = MT / NPH;              // M tiles per phase
    static constexpr int WQ = EPI_WARPS / 4;          // epilogue warps per TMEM lane quarter
    static constexpr int NITEMS = MTP * NGRP;         // (M tile, point group) work items per quarter and phase
    static constexpr int ROT = WQ / 2;                // item rotation between phases (balances 6 items over 4 warps)
    static constexpr int NG = N / 64;                 // 64-column groups of the MN-major operand
    static constexpr int SN = 1024;                   // byte stride between column groups   (descriptor LBO)
    static constexpr int SK = NG * 1024;              // byte stride between 8-deep k groups (descriptor SBO)
    static constexpr int B_PART = (H / 8) * SK;       // bytes of one operand part (hi or lo)
    static constexpr int NPARTS = (PASSES == 3) ? 2 : 1;
    static constexpr int B_BYTES = B_PART * NPARTS;
    static constexpr int KSLABS = H / 64;
    static constexpr int KS_PER_PH = KSLABS / NPH;
    static constexpr int ACC_COLS = MT * N;           // TMEM columns of one accumulator buffer
    static constexpr int TMEM_COLS = (2 * ACC_COLS <= 256) ? 256 : 512;
    static constexpr int ETHREADS = EPI_WARPS * 32;
    static constexpr int NUB = H / 32;                // 32-unit blocks
    static constexpr int OFF_RING = B_BYTES;
    static constexpr int OFF_XYZ = OFF_RING + STAGES * STAGE_BYTES;      // [2][N*3] floats (the next tile's coordinates are staged early)
    static constexpr int OFF_UBAR = OFF_XYZ + 2 * N * 3 * 4;              // [N] floats
    static constexpr int OFF_BAR = OFF_UBAR + N * 4;
    static constexpr int OFF_HEAD = OFF_BAR + 256;                        // forward: [NUB][N] floats of last-layer partial sums
    static constexpr int OFF_ACC = OFF_HEAD;                              // reverse: [(Lh+1) db | 3 dW0 | dw_last][H] floats (optional)
    static constexpr int SMEM_BYTES = OFF_HEAD + NUB * N * 4 + 1024;      // + alignment slack; the reverse kernel adds the accumulators
    static_assert(H == 128 || H == 256 || H == 512, "H must be 128, 256 or 512");
    static_assert(2 * ACC_COLS <= 512, "accumulators exceed TMEM");
    static_assert(NITEMS >= WQ, "every epilogue warp needs at least one item per phase (it arrives on b_ready after its last)");
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

// Order in which the weight tiles (K slab ks, M tile m) of one layer step are produced (TMA warp) and consumed (MMA warp):
//   part 0: the K slabs of operand phase 0 for all M tiles                                    [after b_ready[0]]
//   then, per M-tile half h: its M tiles over the K slabs of operand phase 1                   [after b_ready[1]]
//   and the accumulators of half h are complete right after it -> ready(h): the epilogue of the next step starts on
//   those M tiles while the tensor pipe is still working on the other half.
template <class G, class Tile, class Wait, class Ready>
__device__ __forceinline__ void weight_tile_order(Tile&& tile, Wait&& wait, Ready&& ready) {
    if constexpr (G::NPH == 1) {
        wait(0);
        for (int ks = 0; ks < G::KSLABS; ++ks)
            for (int m = 0; m < G::MT; ++m) tile(ks, m);
        ready(0);
    } else {
        wait(0);
        for (int ks = 0; ks < G::KS_PER_PH; ++ks)
            for (int m = 0; m < G::MT; ++m) tile(ks, m);
        wait(1);
        for (int h = 0; h < 2; ++h) {
            for (int m = h * G::MTP; m < (h + 1) * G::MTP; ++m)
                for (int ks = G::KS_PER_PH; ks < G::KSLABS; ++ks) tile(ks, m);
            ready(h);
        }
    }
}

// sin/cos of a phase of any magnitude: two-constant Cody-Waite reduction to [-pi, pi], then MUFU
__device__ __forceinline__ void sincos_reduced(float t, float& s, float& c) {
    const float magic = 12582912.f;  // 1.5 * 2^23: round-to-nearest-integer trick
    float k = fmaf(t, 0.15915494309189535f, magic) - magic;
    float r = fmaf(k, -6.2831854820251465f, t);
    r = fmaf(k, 1.7484556000744883e-7f, r);
    s = __sinf(r);
    c = __cosf(r);
}

// Sum NV (<= 32) per-lane values over the 32 lanes.  x[] is padded to P = 4, 16 or 32 entries; afterwards x[0] of
// lane j (j < P) holds the total of value j.
template <int P>
__device__ __forceinline__ void warp_reduce_scatter(float* x, int lane) {
#pragma unroll
    for (int off = 16; off >= P; off >>= 1)
#pragma unroll
        for (int j = 0; j < P; ++j) x[j] += __shfl_xor_sync(0xffffffffu, x[j], off);
#pragma unroll
    for (int off = P / 2; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int j = 0; j < off; ++j) {
            float send = up ? x[j] : x[j + off];
            float keep = up ? x[j + off] : x[j];
            x[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
}

__device__ __forceinline__ void point_coords(const PreSrc& s, int64_t gp, float& x, float& y, float& z) {
    if (s.mode == 2) {
        int64_t g = s.grid_base + gp;
        int iz = (int)(g % s.nz);
        int64_t t2 = g / s.nz;
        int ix = (int)(t2 % s.nx);
        int iy = (int)(t2 / s.nx);
        x = __half2float(s.ax[ix]);
        y = __half2float(s.ay[iy]);
        z = __half2float(s.az[iz]);
    } else {
        const float* xp = s.xyz + gp * s.xyz_stride;
        x = __ldg(xp);
        y = __ldg(xp + 1);
        z = (s.din > 2) ? __ldg(xp + 2) : 0.f;
    }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// weight operands (jet_tc.cu)
struct Prep {
    const float* bias30;
    const float* w0_30;
};
int run_prep(const DigsNet* net, void* dst, cudaStream_t st);
int prepare_weights(const DigsNet* net, void* ws, CUtensorMap* tmap, Prep* out, cudaStream_t st);

// paired-CTA variant of the network kernel (jet_tc_pair.cu); H = 256 only
bool pair_supported(const DigsNet* net);
int launch_pair(const DigsNet* net, const CUtensorMap& tmap, const Args& a, int want_jet, int want_lap, int passes, bool bwd,
                cudaStream_t st);

}  // namespace tcpath
}  // namespace digs
