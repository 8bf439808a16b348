// tcgen05 weight-gradient GEMM of the reverse pass (DIGS_MODE_BF16X2 / DIGS_MODE_BF16).
//
//   dW_l[u][k] += 30 * sum_col  GT_l[u][col] * actT_{l-1}[k][col]        col = (channel, point), K = C * n_pad
//
// which is the sum the reference's autograd accumulates into decoder.fc_block.net.{l}.0.weight.grad through the
// value, Jacobian and Laplacian channels (SURVEY.md 8a "W_bar += z_bar^T h + sum_k Z_bar_k^T A_k + Q_bar^T P").
// Both operands were written point-contiguous (K-major) as bf16 hi/lo rows by the forward / reverse jet kernels, so
// this kernel is a plain TMA-fed split-K GEMM: CTA (split, tile, layer) owns a 128 x NT output tile and a K range,
// accumulates it in TMEM over a 2..4-stage TMA ring, and adds it to the gradient with vector fp32 reductions.
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-5 epilogue.
#include "digs_internal.h"
#include "tc_common.cuh"
#include <cuda.h>

namespace digs {
namespace tcpath {

using namespace digs::tc;

constexpr int WG_THREADS = 192;

template <int NT, int PASSES>
struct WCfg {
    static constexpr int NPARTS = (PASSES == 3) ? 2 : 1;
    static constexpr int A_BYTES = 128 * 64 * 2;
    static constexpr int B_BYTES = NT * 64 * 2;
    static constexpr int STAGE_BYTES = (A_BYTES + B_BYTES) * NPARTS;
    static constexpr int STAGES = (196608 / STAGE_BYTES) > 4 ? 4 : (196608 / STAGE_BYTES);
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;
    static constexpr int TMEM_COLS = NT;
};

struct WArgs {
    int H, C;
    int64_t kcols;        // C * n_pad
    int nsplit, n_ntiles;
    float* dW[DIGS_MAX_LAYERS];
};

template <int NT, int PASSES>
__global__ void __launch_bounds__(WG_THREADS, 1) k_tc_wgrad(const __grid_constant__ CUtensorMap tmapG,
                                                            const __grid_constant__ CUtensorMap tmapA, WArgs a) {
    using G = WCfg<NT, PASSES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G::STAGES * G::STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + G::STAGES;
    uint64_t* acc_bar = bars + 2 * G::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * G::STAGES + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int split = blockIdx.x;
    const int m = blockIdx.y / a.n_ntiles, nt = blockIdx.y % a.n_ntiles;
    const int li = blockIdx.z;                      // layer slot: layer l = li + 1
    const int64_t kblocks = (a.kcols + 63) / 64;
    const int64_t kb0 = kblocks * split / a.nsplit, kb1 = kblocks * (split + 1) / a.nsplit;
    if (kb0 >= kb1) return;                         // uniform over the CTA

    if (tid == 0) {
        for (int i = 0; i < G::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(acc_bar, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmapG);
        tma_prefetch_desc(&tmapA);
    }
    if (warp == 1) tmem_alloc<G::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&full[stage], G::STAGE_BYTES);
                uint8_t* s = smem + stage * G::STAGE_BYTES;
                for (int part = 0; part < G::NPARTS; ++part) {
                    tma_load_2d(s + part * G::A_BYTES, &tmapG, &full[stage], (int)(kb * 64), (li * 2 + part) * a.H + m * 128);
                    tma_load_2d(s + G::NPARTS * G::A_BYTES + part * G::B_BYTES, &tmapA, &full[stage], (int)(kb * 64),
                                (li * 2 + part) * a.H + nt * NT);
                }
                if (++stage == G::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, NT, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t sA = smem_u32(smem + stage * G::STAGE_BYTES);
                const uint32_t sBm = sA + G::NPARTS * G::A_BYTES;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    uint64_t da_hi = make_smem_desc(sA + kk * 32, 16, 1024, LAYOUT_SW128);
                    uint64_t db_hi = make_smem_desc(sBm + kk * 32, 16, 1024, LAYOUT_SW128);
                    umma_bf16(tmem_base, da_hi, db_hi, idesc, (kb > kb0) || (kk > 0));
                    if constexpr (PASSES == 3) {
                        uint64_t da_lo = make_smem_desc(sA + G::A_BYTES + kk * 32, 16, 1024, LAYOUT_SW128);
                        uint64_t db_lo = make_smem_desc(sBm + G::B_BYTES + kk * 32, 16, 1024, LAYOUT_SW128);
                        umma_bf16(tmem_base, da_hi, db_lo, idesc, 1);
                        umma_bf16(tmem_base, da_lo, db_hi, idesc, 1);
                    }
                }
                umma_commit(&empty[stage]);
                if (++stage == G::STAGES) { stage = 0; phase ^= 1; }
            }
            umma_commit(acc_bar);
        }
    } else {
        mbar_wait(acc_bar, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int u = m * 128 + q * 32 + lane;
        float* dst = a.dW[li + 1] + (int64_t)u * a.H + nt * NT;
#pragma unroll 4
        for (int c0 = 0; c0 < NT; c0 += 8) {
            float v[8];
            tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + c0, v);
            tmem_ld_wait();
            red_add_v4(dst + c0, 30.f * v[0], 30.f * v[1], 30.f * v[2], 30.f * v[3]);
            red_add_v4(dst + c0 + 4, 30.f * v[4], 30.f * v[5], 30.f * v[6], 30.f * v[7]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<G::TMEM_COLS>(tmem_base);
}

template <int NT, int PASSES>
static int launch_wgrad(const CUtensorMap& tG, const CUtensorMap& tA, const WArgs& a, int MT, int Lh, cudaStream_t st) {
    using G = WCfg<NT, PASSES>;
    auto kern = k_tc_wgrad<NT, PASSES>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES);
        if (e != cudaSuccess) { set_error("tc wgrad: smem attribute: %s", cudaGetErrorString(e)); return DIGS_ECUDA; }
        attr_set = true;
    }
    dim3 grid(a.nsplit, MT * a.n_ntiles, Lh);
    kern<<<grid, WG_THREADS, G::SMEM_BYTES, st>>>(tG, tA, a);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("tc wgrad launch: %s", cudaGetErrorString(e)); return DIGS_ECUDA; }
    return DIGS_OK;
}

int wgrad(const DigsNet* net, const __nv_bfloat16* GT, const __nv_bfloat16* actT, int C, int64_t n_pad,
          const DigsGrads* grads, int passes, cudaStream_t st) {
    EncodeFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return DIGS_ECUDA; }
    const int H = net->H, Lh = net->Lh;
    const int NT = H >= 256 ? 256 : 128;
    const int64_t kcols = (int64_t)C * n_pad;
    CUtensorMap tG, tA;
    cuuint64_t gdim[2] = {(cuuint64_t)kcols, (cuuint64_t)Lh * 2 * H};
    cuuint64_t gstr[1] = {(cuuint64_t)kcols * 2};
    cuuint32_t estr[2] = {1, 1};
    cuuint32_t boxA[2] = {64, 128}, boxB[2] = {64, (cuuint32_t)NT};
    CUresult r1 = enc(&tG, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(GT), gdim, gstr, boxA, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&tA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(actT), gdim, gstr, boxB, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(wgrad) failed (%d, %d)", (int)r1, (int)r2); return DIGS_ECUDA; }
    WArgs a{};
    a.H = H;
    a.C = C;
    a.kcols = kcols;
    a.n_ntiles = H / NT;
    const int MT = H / 128;
    int tiles = MT * a.n_ntiles * Lh;
    int64_t kblocks = (kcols + 63) / 64;
    int ns = num_sms() / tiles;
    if (ns < 1) ns = 1;
    if (ns > kblocks) ns = (int)kblocks;
    a.nsplit = ns;
    for (int l = 0; l <= Lh + 1; ++l) a.dW[l] = grads->dW[l];
    if (NT == 256) return passes == 3 ? launch_wgrad<256, 3>(tG, tA, a, MT, Lh, st) : launch_wgrad<256, 1>(tG, tA, a, MT, Lh, st);
    return passes == 3 ? launch_wgrad<128, 3>(tG, tA, a, MT, Lh, st) : launch_wgrad<128, 1>(tG, tA, a, MT, Lh, st);
}

}  // namespace tcpath
}  // namespace digs
