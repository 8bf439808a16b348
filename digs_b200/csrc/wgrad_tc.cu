// tcgen05 weight-gradient GEMM of the reverse pass (DIGS_MODE_BF16X2 / DIGS_MODE_BF16).
//
//   dW_l[u][k] += 30 * sum_col  Zhat_l[col][u] * X_{l-1}[col][k]         col = (channel, point), K = C * n_pad
//
// which is the sum the reference's autograd accumulates into decoder.fc_block.net.{l}.0.weight.grad through the
// value, Jacobian and Laplacian channels (SURVEY.md 8a "W_bar += z_bar^T h + sum_k Z_bar_k^T A_k + Q_bar^T P").
//
// Operands: GP (adjoints / 30, written by the reverse jet kernel) and actP (layer inputs, written by the forward jet
// kernel), both [slot][col][unit] words of (bf16 hi | bf16 lo << 16).  Read as bf16 they are MN-major matrices with
// 2H "split rows" m' = 2*unit + part per column, so ONE tcgen05.mma over (m', n') forms all four hi/lo cross products
// in separate accumulators; the epilogue adds each 2x2 block into one fp32 gradient entry.  TMA boxes of 64 split rows
// x 64 columns with SWIZZLE_128B land directly in the canonical MN-major UMMA layout (LBO = 8192, SBO = 1024).
//
// A work unit (K split, layer, output tile) is a 256 x 256 (m', n') = 128 x 128 (u, k) output tile over a K range: it is
// accumulated in all 512 TMEM columns over a 3-stage TMA ring and added to the gradient with fp32 vector reductions.
// Plain launch (two-call API): one unit per CTA.  Behind the fused step: a persistent grid draws units from a queue.
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-9 epilogue (two warps per TMEM
// lane quarter, each draining half of the accumulator columns).
//
// Behind the fused step (step_tc.cu) the launch is a programmatic dependent launch: the persistent tile kernel walks its
// tiles in rounds of gridDim.x, the last round is partial (C2: 1094 tiles on 148 CTAs = 7 full rounds + 58 tiles), and
// the CTAs without a tile in it exit early.  The GEMM's CTAs are dispatched onto those SMs and draw units split index
// slowest, so the first units cover the K ranges (operand columns) of the early rounds; each unit waits on the per-round
// completion counters the tile kernel publishes (release / acquire at gpu scope) instead of on the whole grid.  The units
// of the last split also wait for every tile-kernel CTA to have finished, so that completion of this grid implies
// completion of that one.  (Timeline and A/B numbers: profiles/r02b_experiments.md.)
#include "digs_internal.h"
#include "tc_common.cuh"
#include <cuda.h>
#include <cstdlib>

namespace digs {
namespace tcpath {

using namespace digs::tc;

constexpr int WG_THREADS = 320;             // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter)
constexpr int WG_BOX = 64 * 64 * 2;            // one TMA box: 64 split rows x 64 columns, bf16
constexpr int WG_OPER = 4 * WG_BOX;            // one operand tile: 256 split rows x 64 columns
constexpr int WG_STAGE = 2 * WG_OPER;          // A + B
constexpr int WG_STAGES = 3;
constexpr int WG_SMEM = WG_STAGES * WG_STAGE + 256 + 1024;

struct WArgs {
    int H;
    int64_t kcols;        // C * n_pad
    int nsplit, n_ntiles, n_mtiles, Lh;
    float* dW[DIGS_MAX_LAYERS];
    WgradSync sync;       // ctr == NULL: every operand column is final at launch
};

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Spin until *p >= want.  The producer (a co-resident or already finished tile-kernel CTA) never waits on this grid, so
// the wait always ends; the time bound turns a protocol bug into a trap instead of a hung GPU.
__device__ __forceinline__ void wait_counter(const unsigned int* p, unsigned int want) {
    if (ld_acquire_gpu(p) >= want) return;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_gpu(p) < want) {
        __nanosleep(256);
        if (global_ns() - t0 > 30000000000ull) __trap();
    }
}

__device__ __forceinline__ int64_t tile_of_col(const WgradSync& s, int64_t col) {
    return (col < s.cols0) ? col / s.ct0 : s.tiles0 + (col - s.cols0) / s.ct1;
}

__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// Work unit u = (K split, layer, output tile), split slowest.  Plain launch: one unit per CTA (u = blockIdx.x).  Behind the
// fused step: a persistent grid draws units from a global queue, so the CTAs that land on early-freed SMs simply take more
// of them; the next unit's operand loads start while the epilogue warps still drain the accumulators of the current one.
__global__ void __launch_bounds__(WG_THREADS, 1) k_tc_wgrad(const __grid_constant__ CUtensorMap tmapG,
                                                            const __grid_constant__ CUtensorMap tmapA, WArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE);
    uint64_t* full = bars;
    uint64_t* empty = bars + WG_STAGES;
    uint64_t* acc_bar = bars + 2 * WG_STAGES;           // accumulators of the current unit complete (tcgen05.commit)
    uint64_t* acc_free = bars + 2 * WG_STAGES + 1;      // ... and read out by the 256 epilogue threads
    uint64_t* unit_ready = bars + 2 * WG_STAGES + 2;    // [2]: s_unit[i & 1] published by the producer
    uint64_t* unit_free = bars + 2 * WG_STAGES + 4;     // [2]: ... and read by the MMA thread and the 256 epilogue threads
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG_STAGES + 6);
    volatile int* s_unit = reinterpret_cast<volatile int*>(tmem_slot + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_out = a.n_mtiles * a.n_ntiles;
    const int per_split = n_out * a.Lh;
    const int total_units = per_split * a.nsplit;
    const int64_t kblocks = (a.kcols + 63) / 64;
    const bool dyn = a.sync.ctr != nullptr;
    unsigned int* queue = dyn ? a.sync.ctr + WG_QUEUE_IDX : nullptr;

    if (tid == 0) {
        for (int i = 0; i < WG_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(acc_bar, 1);
        mbar_init(acc_free, 256);
        mbar_init(&unit_ready[0], 1);
        mbar_init(&unit_ready[1], 1);
        mbar_init(&unit_free[0], 257);
        mbar_init(&unit_free[1], 257);
        fence_barrier_init();
        tma_prefetch_desc(&tmapG);
        tma_prefetch_desc(&tmapA);
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0;; ++i) {
                int u = dyn ? (int)atomicAdd(queue, 1u) : (i == 0 ? (int)blockIdx.x : total_units);
                if (u > total_units) u = total_units;
                // a slot is rewritten only after both consumers have read its previous unit (with K ranges shorter than the
                // ring the producer would otherwise run several units ahead and a barrier phase would be skipped)
                if (i >= 2) mbar_wait(&unit_free[i & 1], (uint32_t)((i >> 1) - 1) & 1u);
                s_unit[i & 1] = u;
                mbar_arrive(&unit_ready[i & 1]);         // release: the two consumers read s_unit after their wait
                if (u >= total_units) break;
                const int split = u / per_split, li = (u % per_split) / n_out;
                const int mt = (u % n_out) / a.n_ntiles, nt = u % a.n_ntiles;
                const int64_t kb0 = kblocks * split / a.nsplit, kb1 = kblocks * (split + 1) / a.nsplit;
#ifdef DIGS_DIAG_TRACE
                unsigned long long* trace = dyn ? reinterpret_cast<unsigned long long*>(a.sync.ctr + WG_MAX_ROUNDS + 1) + 512 + 4 * u : nullptr;
                if (trace) {
                    unsigned int smid;
                    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
                    trace[0] = global_ns();
                    trace[3] = smid;
                }
#endif
                if (dyn && kb0 < kb1) {
                    // the rounds of tiles whose operand columns this K range covers
                    int64_t c_hi = kb1 * 64 < a.kcols ? kb1 * 64 : a.kcols;
                    const int64_t r0 = tile_of_col(a.sync, kb0 * 64) / a.sync.grid;
                    const int64_t r1 = tile_of_col(a.sync, c_hi - 1) / a.sync.grid;
                    for (int64_t r = r0; r <= r1; ++r) {
                        int64_t left = a.sync.total_tiles - r * a.sync.grid;
                        wait_counter(a.sync.ctr + r, (unsigned int)(left < a.sync.grid ? left : a.sync.grid));
                    }
                    fence_proxy_async_global();         // the words were written by generic-proxy stores; TMA reads them
                }
#ifdef DIGS_DIAG_TRACE
                if (trace) trace[1] = global_ns();
#endif
                for (int64_t kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full[stage], WG_STAGE);
                    uint8_t* s = smem + stage * WG_STAGE;
                    const int row = (int)(kb * 64);     // columns past kcols are zero-filled by TMA (3-D map: no bleed into slot li+1)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        tma_load_3d(s + b * WG_BOX, &tmapG, &full[stage], mt * 256 + b * 64, row, li);
                        tma_load_3d(s + WG_OPER + b * WG_BOX, &tmapA, &full[stage], nt * 256 + b * 64, row, li);
                    }
                    if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
                }
#ifdef DIGS_DIAG_TRACE
                if (trace) trace[2] = global_ns();
#endif
                // last split: this grid must not complete before the tile kernel has published everything it writes
                if (dyn && split == a.nsplit - 1) wait_counter(a.sync.ctr + WG_MAX_ROUNDS, (unsigned int)a.sync.grid);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, 256, /*A MN-major*/ 1, /*B MN-major*/ 1);
            int stage = 0;
            uint32_t phase = 0;
            int done = 0;                                // units whose accumulators this CTA has committed
            for (int i = 0;; ++i) {
                mbar_wait(&unit_ready[i & 1], (uint32_t)(i >> 1) & 1u);
                const int u = s_unit[i & 1];
                mbar_arrive(&unit_free[i & 1]);
                if (u >= total_units) break;
                const int split = u / per_split;
                const int64_t kb0 = kblocks * split / a.nsplit, kb1 = kblocks * (split + 1) / a.nsplit;
                if (kb0 >= kb1) continue;                // empty K range (uniform decision in all roles)
                if (done > 0) {                          // the previous unit's accumulators have been read out
                    mbar_wait(acc_free, (uint32_t)(done - 1) & 1u);
                    tc_fence_after();
                }
                for (int64_t kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sA = smem_u32(smem + stage * WG_STAGE);
                    const uint32_t sBm = sA + WG_OPER;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        uint64_t db = make_smem_desc(sBm + kk * 2048, WG_BOX, 1024, LAYOUT_SW128);
#pragma unroll
                        for (int m = 0; m < 2; ++m) {
                            uint64_t da = make_smem_desc(sA + m * 2 * WG_BOX + kk * 2048, WG_BOX, 1024, LAYOUT_SW128);
                            umma_bf16(tmem_base + m * 256, da, db, idesc, (kb > kb0) || (kk > 0));
                        }
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(acc_bar);
                ++done;
            }
        }
    } else {
        const int q = warp & 3;                         // TMEM lane quarter this warp may read
        const int chalf = (warp - 2) >> 2;              // which half of the 256 accumulator columns of each M block
        const int part = lane & 1;
        int done = 0;
        for (int i = 0;; ++i) {
            mbar_wait(&unit_ready[i & 1], (uint32_t)(i >> 1) & 1u);
            const int u = s_unit[i & 1];
            mbar_arrive(&unit_free[i & 1]);
            if (u >= total_units) break;
            const int split = u / per_split, li = (u % per_split) / n_out;
            const int mt = (u % n_out) / a.n_ntiles, nt = u % a.n_ntiles;
            const int64_t kb0 = kblocks * split / a.nsplit, kb1 = kblocks * (split + 1) / a.nsplit;
            if (kb0 >= kb1) continue;
            mbar_wait(acc_bar, (uint32_t)done & 1u);
            tc_fence_after();
#pragma unroll 1
            for (int m = 0; m < 2; ++m) {
                // TMEM lane = split row m' = 2*unit + part within this 128-lane tile
                const int uu = mt * 128 + m * 64 + q * 16 + (lane >> 1);
                float* dst = a.dW[li + 1] + (int64_t)uu * a.H + nt * 128;
#pragma unroll 4
                for (int c0 = chalf * 128; c0 < chalf * 128 + 128; c0 += 8) {
                    float v[8];
                    tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + m * 256 + c0, v);
                    tmem_ld_wait();
                    float s0 = v[0] + v[1], s1 = v[2] + v[3], s2 = v[4] + v[5], s3 = v[6] + v[7];   // B hi + lo
                    s0 += __shfl_xor_sync(0xffffffffu, s0, 1);                                      // A hi + lo
                    s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
                    s3 += __shfl_xor_sync(0xffffffffu, s3, 1);
                    const int k = c0 >> 1;
                    if (part == 0) red_add_v2(dst + k, 30.f * s0, 30.f * s1);
                    else red_add_v2(dst + k + 2, 30.f * s2, 30.f * s3);
                }
            }
            tc_fence_before();
            mbar_arrive(acc_free);
            ++done;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

int wgrad(const DigsNet* net, const __nv_bfloat16* GP_words, const __nv_bfloat16* actP, int C, int64_t n_pad,
          const DigsGrads* grads, int passes, cudaStream_t st, const WgradSync* sync) {
    (void)passes;   // both operands always carry hi and lo
    EncodeFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return DIGS_ECUDA; }
    const int H = net->H, Lh = net->Lh;
    const int64_t kcols = (int64_t)C * n_pad;
    CUtensorMap tG, tA;
    cuuint64_t gdim[3] = {(cuuint64_t)2 * H, (cuuint64_t)kcols, (cuuint64_t)Lh};
    cuuint64_t gstr[2] = {(cuuint64_t)H * 4, (cuuint64_t)kcols * H * 4};
    cuuint32_t estr[3] = {1, 1, 1};
    cuuint32_t box[3] = {64, 64, 1};
    CUresult r1 = enc(&tG, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(GP_words), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&tA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(actP), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(wgrad) failed (%d, %d)", (int)r1, (int)r2); return DIGS_ECUDA; }
    WArgs a{};
    a.H = H;
    a.Lh = Lh;
    a.kcols = kcols;
    a.n_ntiles = 2 * H / 256;
    a.n_mtiles = 2 * H / 256;
    const int n_mtiles = a.n_mtiles;
    int tiles = n_mtiles * a.n_ntiles * Lh;
    int64_t kblocks = (kcols + 63) / 64;
    int ns = num_sms() / tiles;
    if (sync && sync->ctr) {
        // overlapped launch (persistent grid + unit queue): units of roughly a quarter of a CTA's share, so that the SMs that
        // free up early take more of them and the grid ends evenly (DIGS_B200_WGRAD_SPLITS overrides, for measurements)
        ns = (4 * num_sms() + tiles - 1) / tiles;
        if (const char* e = getenv("DIGS_B200_WGRAD_SPLITS")) {
            int v = atoi(e);
            if (v > 0) ns = v;
        }
        a.sync = *sync;
    }
    if (ns < 1) ns = 1;
    if (ns > kblocks) ns = (int)kblocks;
    if (ns > 4096) ns = 4096;
    a.nsplit = ns;
    for (int l = 0; l <= Lh + 1; ++l) a.dW[l] = grads->dW[l];
    static bool attr_set[64] = {false};      // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(k_tc_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
        if (e != cudaSuccess) { set_error("tc wgrad: smem attribute: %s", cudaGetErrorString(e)); return DIGS_ECUDA; }
        attr_set[dev & 63] = true;
    }
    const int units = tiles * ns;
    dim3 grid(a.sync.ctr ? (units < num_sms() ? units : num_sms()) : units, 1, 1);
    if (a.sync.ctr) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(WG_THREADS, 1, 1);
        cfg.dynamicSmemBytes = WG_SMEM;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        cudaError_t le = cudaLaunchKernelEx(&cfg, k_tc_wgrad, tG, tA, a);
        if (le != cudaSuccess) { set_error("tc wgrad dependent launch: %s", cudaGetErrorString(le)); return DIGS_ECUDA; }
    } else {
        k_tc_wgrad<<<grid, WG_THREADS, WG_SMEM, st>>>(tG, tA, a);
    }
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("tc wgrad launch: %s", cudaGetErrorString(e)); return DIGS_ECUDA; }
    return DIGS_OK;
}

}  // namespace tcpath
}  // namespace digs
