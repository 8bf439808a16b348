// Stand-alone probe for the CTA-pair (cta_group::2) mechanisms a paired jet kernel needs, checked against a CPU GEMM:
//   * tcgen05.alloc / dealloc with cta_group::2, cluster of two CTAs;
//   * A (K-major, SWIZZLE_128B): each CTA loads ITS 128 rows by TMA, both loads complete on the LEADER's mbarrier;
//   * B (MN-major, SWIZZLE_128B): each CTA stores 64 of the 128 columns; the rows (k) are produced by the threads of
//     BOTH CTAs, half of them through distributed shared memory (st.shared::cluster), and published to the leader's
//     mbarrier with a cluster-scope release arrive after fence.proxy.async;
//   * one elected leader thread issues tcgen05.mma.cta_group::2 (M = 256, N = 128) and commits with a multicast
//     arrive to both CTAs; each CTA reads its 128 accumulator rows from its own TMEM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma2_probe umma2_probe.cu -I../../digs_b200/csrc
#include "tc_common.cuh"
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
using namespace digs::tc;

constexpr int M = 256, N = 128, K = 64;

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
__device__ __forceinline__ void st_cluster_b16(uint32_t caddr, uint16_t v) {
    asm volatile("st.shared::cluster.b16 [%0], %1;" ::"r"(caddr), "h"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t caddr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(caddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_2d_pair(void* dst_smem, const void* tmap, uint64_t* leader_bar, int c0, int c1) {
    // executed by both CTAs; the peer bit of the barrier address is cleared so the bytes are counted on CTA 0's barrier
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
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
probe2(const __grid_constant__ CUtensorMap tmapA, const __nv_bfloat16* B /*[N][K]*/, float* D) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;              // 16 KB: this CTA's 128 rows x 64 k
    uint8_t* sB = smem + 16384;      // 8 KB: this CTA's 64 columns x 64 k, MN-major
    __shared__ uint64_t bar_tma, bar_b, bar_mma;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = cluster_rank();
    if (tid == 0) {
        mbar_init(&bar_tma, 1);
        mbar_init(&bar_b, 2 * 128);      // every thread of both CTAs arrives once
        mbar_init(&bar_mma, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();                  // barriers of both CTAs initialised before anyone signals them
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    // ---- A: each CTA loads its own 128 rows; both loads are counted on the leader's barrier ----------------------------
    if (tid == 0) {
        if (rank == 0) mbar_arrive_expect_tx(&bar_tma, 2 * 128 * K * 2);
        tma_load_2d_pair(sA, &tmapA, &bar_tma, 0, (int)rank * 128);
    }
    // ---- B: CTA r produces the k rows [32 r, 32 r + 32) for ALL 128 columns; column half h lives in CTA h -----------------
    for (int e = tid; e < 32 * N; e += 128) {
        const int k = (int)rank * 32 + e / N, n = e % N;
        const int h = n >> 6, n_lo = n & 63, k_lo = k & 7, k_hi = k >> 3;
        const uint32_t off = k_hi * 1024 + k_lo * 128 + ((((n_lo >> 3) ^ k_lo)) << 4) + (n_lo & 7) * 2;
        const __nv_bfloat16 v = B[n * K + k];
        st_cluster_b16(map_to_rank(smem_u32(sB) + off, (uint32_t)h), *reinterpret_cast<const uint16_t*>(&v));
    }
    fence_proxy_async();                 // generic-proxy stores (local and remote) -> visible to the tensor cores' async reads
    mbar_arrive_cluster(map_to_rank(smem_u32(&bar_b), 0));

    if (rank == 0 && tid == 0) {
        mbar_wait_cluster(&bar_b, 0);
        mbar_wait(&bar_tma, 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(M, N, 0, 1);
        for (int kk = 0; kk < K / 16; ++kk) {
            uint64_t da = make_smem_desc(smem_u32(sA) + kk * 32, 16, 1024, LAYOUT_SW128);
            uint64_t db = make_smem_desc(smem_u32(sB) + kk * 2 * 1024, 1024, 1024, LAYOUT_SW128);
            umma_bf16_pair(tmem_base, da, db, idesc, kk > 0);
        }
        umma_commit_pair(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    for (int c = 0; c < N; c += 8) {
        float v[8];
        tmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
        tmem_ld_wait();
        for (int i = 0; i < 8; ++i) D[((int)rank * 128 + warp * 32 + (tid & 31)) * N + c + i] = v[i];
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(128) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    std::vector<__nv_bfloat16> hA(M * K), hB(N * K);
    std::vector<float> fA(M * K), fB(N * K), ref(M * N), out(M * N);
    srand(2);
    for (int i = 0; i < M * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2bfloat16(v); fA[i] = __bfloat162float(hA[i]); }
    for (int i = 0; i < N * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]); }
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)fA[m * K + k] * fB[n * K + k];
            ref[m * N + n] = (float)s;
        }
    __nv_bfloat16 *dA, *dB;
    float* dD;
    cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres);
    CUtensorMap tmap;
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)M};
    cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    const int smem_bytes = 16384 + 8192 + 1024;
    cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemset(dD, 0xff, M * N * 4);
        probe2<<<2, 128, smem_bytes>>>(tmap, dB, dD);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA ERROR %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(out.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
        double md = 0;
        int bad = 0, bad_lo = 0, bad_hi = 0;
        for (int i = 0; i < M * N; ++i) {
            double d = fabs(out[i] - ref[i]);
            if (!(d < 1e-2)) { ++bad; if (i / N < 128) ++bad_lo; else ++bad_hi; }
            if (d > md || d != d) md = d;
        }
        printf("pair M=256 N=128 K=64 rep %d: maxdiff %.3e  bad %d/%d (rows<128: %d, rows>=128: %d)  %s\n", rep, md, bad, M * N,
               bad_lo, bad_hi, bad == 0 ? "OK" : "MISMATCH");
    }
    return 0;
}
