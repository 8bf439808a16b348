// On-device batch sampling.
//
// Reference behaviour reproduced IN DISTRIBUTION (the RNG stream itself is numpy's and is not reproduced):
// ReconDataset.__getitem__ (surface_reconstruction/recon_dataset.py:143-174) draws, every iteration,
//   * n_points surface samples = the first n_points entries of a fresh random permutation of the cloud (so: distinct
//     indices, sampled without replacement), with their normals;
//   * n_points off-surface samples uniform in [-grid_range, grid_range]^d;
// on the host, in 4 DataLoader workers, followed by a pinned H2D copy.  Here the cloud stays resident in HBM and one
// launch fills the three batch buffers:
//   * the permutation is a keyed Feistel bijection on [0, 2^k) >= n_cloud with cycle walking, evaluated point-wise
//     (thread i computes perm(i)); a new key per iteration gives a new permutation, and perm(0..n_points-1) are distinct
//     by construction;
//   * the uniform samples come from a counter-based hash (two rounds of a 64-bit mix per coordinate).
// The iteration counter is read from DEVICE memory and incremented by the kernel, so the launch can sit inside a CUDA
// graph (digs_b200.GraphedStep) and still produce a new batch at every replay.
#include "digs_internal.h"

namespace digs {

__device__ __forceinline__ uint64_t mix64(uint64_t x) {   // splitmix64 finaliser
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// keyed bijection on [0, 2^(2*half_bits)): 4-round balanced Feistel network
__device__ __forceinline__ uint64_t feistel(uint64_t v, int half_bits, uint64_t key) {
    const uint64_t mask = (1ull << half_bits) - 1;
    uint64_t l = v >> half_bits, r = v & mask;
#pragma unroll
    for (int round = 0; round < 4; ++round) {
        uint64_t f = mix64(r ^ (key + 0x632BE59BD9B4E019ull * (round + 1))) & mask;
        uint64_t nl = r;
        r = l ^ f;
        l = nl;
    }
    return (l << half_bits) | r;
}

__global__ void __launch_bounds__(256) k_sample_batch(const float* __restrict__ cloud, const float* __restrict__ normals,
                                                      int64_t n_cloud, int d, int64_t n_pts, float box, uint64_t seed,
                                                      unsigned long long* __restrict__ iter_dev, int half_bits,
                                                      float* __restrict__ mnfld, float* __restrict__ nrm,
                                                      float* __restrict__ nonmnfld) {
    const uint64_t it = *iter_dev;
    const uint64_t key = mix64(seed ^ mix64(it));
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pts) {
        uint64_t j = (uint64_t)i;
        do { j = feistel(j, half_bits, key); } while (j >= (uint64_t)n_cloud);   // cycle walking keeps it a bijection on [0, n_cloud)
        for (int k = 0; k < d; ++k) {
            mnfld[i * d + k] = cloud[j * d + k];
            if (normals) nrm[i * d + k] = normals[j * d + k];
            uint64_t h = mix64(key ^ mix64((uint64_t)(i * 4 + k) + 0xD1B54A32D192ED03ull));
            float u01 = (float)(h >> 40) * (1.0f / 16777216.0f);                 // 24 random bits -> [0, 1)
            nonmnfld[i * d + k] = (2.f * u01 - 1.f) * box;
        }
    }
    // the last block to finish bumps the counter; every block has already read it (grid-wide ordering via a ticket)
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long ticket = atomicAdd(iter_dev + 1, 1ull);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        iter_dev[1] = 0;
        iter_dev[0] = it + 1;
    }
}

int sample_batch(const float* cloud, const float* normals, int64_t n_cloud, int d, int64_t n_pts, float box, uint64_t seed,
                 unsigned long long* iter_dev, float* mnfld, float* nrm, float* nonmnfld, cudaStream_t st) {
    if (n_pts == 0) return DIGS_OK;
    int bits = 2;
    while ((1ll << bits) < n_cloud) ++bits;
    int half_bits = (bits + 1) / 2;
    unsigned blocks = (unsigned)((n_pts + 255) / 256);
    k_sample_batch<<<blocks, 256, 0, st>>>(cloud, normals, n_cloud, d, n_pts, box, seed, iter_dev, half_bits, mnfld, nrm,
                                           nonmnfld);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("sample_batch launch: %s", cudaGetErrorString(e)); return DIGS_ECUDA; }
    return DIGS_OK;
}

}  // namespace digs
