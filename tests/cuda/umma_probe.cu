// Stand-alone probe: validates the tcgen05 descriptor encodings used by digs_b200/csrc/jet_tc.cu against a CPU GEMM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu -I../../digs_b200/csrc
#include "tc_common.cuh"
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
using namespace digs::tc;

constexpr int M = 128, N = 128, K = 64;

struct Cfg {
    int a_tma;        // A via TMA (1) or manual swizzled stores (0)
    int b_mn;         // B MN-major (1) or K-major (0)
    uint32_t b_sn;    // MN-major layout: byte stride between 64-wide n groups
    uint32_t b_sk;    // MN-major layout: byte stride between 8-deep k groups
    uint32_t b_lbo;   // descriptor fields for B
    uint32_t b_sbo;
    int a_mn;         // A MN-major (layout sn=8192, sk=1024, descriptor LBO=8192 SBO=1024)
};

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tmapA, const __nv_bfloat16* A,
                                             const __nv_bfloat16* B /*[N][K]*/, float* D, Cfg cfg) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;              // 16 KB
    uint8_t* sB = smem + 16384;      // 16 KB (+ slack for strided variants)
    __shared__ uint64_t bar_tma, bar_mma;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&bar_tma, 1);
        mbar_init(&bar_mma, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<128>(&tmem_base);
    // zero B region (32 KB) so strided variants read zeros in gaps
    for (int i = tid; i < 32768 / 4; i += 128) ((uint32_t*)sB)[i] = 0;
    __syncthreads();
    // ---- A: K-major SW128: row m at m*128 B, 16B chunk j stored at chunk (j ^ (m & 7))
    if (cfg.a_mn) {
        for (int e = tid; e < M * K; e += 128) {
            int m = e / K, k = e % K;
            int m_lo = m & 63, m_hi = m >> 6, k_lo = k & 7, k_hi = k >> 3;
            uint32_t off = m_hi * 8192 + k_hi * 1024 + k_lo * 128 + ((((m_lo >> 3) ^ k_lo)) << 4) + (m_lo & 7) * 2;
            *reinterpret_cast<__nv_bfloat16*>(sA + off) = A[m * K + k];
        }
    } else if (!cfg.a_tma) {
        for (int e = tid; e < M * 8; e += 128) {
            int m = e >> 3, j = e & 7;
            uint4 v = *reinterpret_cast<const uint4*>(A + m * K + j * 8);
            *reinterpret_cast<uint4*>(sA + m * 128 + ((j ^ (m & 7)) << 4)) = v;
        }
    }
    // ---- B
    if (!cfg.b_mn) {
        for (int e = tid; e < N * 8; e += 128) {
            int n = e >> 3, j = e & 7;
            uint4 v = *reinterpret_cast<const uint4*>(B + n * K + j * 8);
            *reinterpret_cast<uint4*>(sB + n * 128 + ((j ^ (n & 7)) << 4)) = v;
        }
    } else {
        // MN-major SW128: element (n,k): n = n_lo + 64 n_hi, k = k_lo + 8 k_hi
        //   byte = n_hi*sn + k_hi*sk + k_lo*128 + (((n_lo/8) ^ k_lo) * 16) + (n_lo%8)*2
        for (int e = tid; e < N * K; e += 128) {
            int n = e / K, k = e % K;
            int n_lo = n & 63, n_hi = n >> 6, k_lo = k & 7, k_hi = k >> 3;
            uint32_t off = n_hi * cfg.b_sn + k_hi * cfg.b_sk + k_lo * 128 + ((((n_lo >> 3) ^ k_lo)) << 4) + (n_lo & 7) * 2;
            *reinterpret_cast<__nv_bfloat16*>(sB + off) = B[n * K + k];
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        if (cfg.a_tma) {
            mbar_arrive_expect_tx(&bar_tma, M * K * 2);
            tma_load_2d(sA, &tmapA, &bar_tma, 0, 0);
            mbar_wait(&bar_tma, 0);
        }
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(M, N, cfg.a_mn, cfg.b_mn);
        for (int kk = 0; kk < K / 16; ++kk) {
            uint64_t da = cfg.a_mn ? make_smem_desc(smem_u32(sA) + kk * 2048, 8192, 1024, LAYOUT_SW128)
                                   : make_smem_desc(smem_u32(sA) + kk * 32, 16, 1024, LAYOUT_SW128);
            uint64_t db;
            if (!cfg.b_mn) db = make_smem_desc(smem_u32(sB) + kk * 32, 16, 1024, LAYOUT_SW128);
            else db = make_smem_desc(smem_u32(sB) + kk * 2 * cfg.b_sk, cfg.b_lbo, cfg.b_sbo, LAYOUT_SW128);
            umma_bf16(tmem_base, da, db, idesc, kk > 0);
        }
        umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    for (int c = 0; c < N; c += 8) {
        float v[8];
        tmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
        tmem_ld_wait();
        for (int i = 0; i < 8; ++i) D[(warp * 32 + (tid & 31)) * N + c + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<128>(tmem_base);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    std::vector<__nv_bfloat16> hA(M * K), hB(N * K);
    std::vector<float> fA(M * K), fB(N * K), ref(M * N), out(M * N);
    srand(1);
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
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 1024);
    struct { const char* name; Cfg c; } tests[] = {
        {"A manual K-major, B manual K-major", {0, 0, 0, 0, 0, 0}},
        {"A TMA    K-major, B manual K-major", {1, 0, 0, 0, 0, 0}},
        {"B MN-major sn=8192 sk=1024, desc LBO=sn SBO=sk", {0, 1, 8192, 1024, 8192, 1024}},
        {"B MN-major sn=8192 sk=1024, desc LBO=sk SBO=sn", {0, 1, 8192, 1024, 1024, 8192}},
        {"B MN-major sn=1024 sk=2048, desc LBO=sn SBO=sk", {0, 1, 1024, 2048, 1024, 2048}},
        {"B MN-major sn=1024 sk=2048, desc LBO=sk SBO=sn", {0, 1, 1024, 2048, 2048, 1024}},
        {"A TMA, B MN-major sn=8192 sk=1024, LBO=sn SBO=sk", {1, 1, 8192, 1024, 8192, 1024, 0}},
        {"A MN-major, B K-major", {0, 0, 0, 0, 0, 0, 1}},
        {"A MN-major, B MN-major sn=8192 sk=1024", {0, 1, 8192, 1024, 8192, 1024, 1}},
    };
    for (auto& t : tests) {
        cudaMemset(dD, 0xff, M * N * 4);
        probe<<<1, 128, 16384 + 32768 + 1024>>>(tmap, dA, dB, dD, t.c);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-55s CUDA ERROR %s\n", t.name, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(out.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
        double md = 0;
        int bad = 0;
        for (int i = 0; i < M * N; ++i) { double d = fabs(out[i] - ref[i]); if (!(d < 1e-2)) ++bad; if (d > md || d != d) md = d; }
        printf("%-55s maxdiff %.3e  bad %d/%d  %s\n", t.name, md, bad, M * N, bad == 0 ? "OK" : "MISMATCH");
    }
    return 0;
}
