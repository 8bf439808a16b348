// Optimizer tail of the training step: global-norm gradient clipping followed by Adam, over the
// flat parameter / gradient buffers, as two launches.
//
// Reference behaviour reproduced: torch.nn.utils.clip_grad_norm_(params, 10.) (train_surface_reconstruction.py:130,
// train_shapespace.py:147; hard-coded, the --grad_clip_norm flag is ignored) followed by torch.optim.Adam.step()
// (train_surface_reconstruction.py:63,132: lr 5e-5, betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad).
//   clip:  g *= min(1, max_norm / (||g||_2 + 1e-6))
//   adam:  m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include "digs_internal.h"

namespace digs {

__global__ void k_step_inc(int32_t* step_dev, float* sumsq) {
    *step_dev += 1;
    *sumsq = 0.f;
}

__global__ void __launch_bounds__(256) k_sumsq(const float* __restrict__ g, int64_t n, float* __restrict__ out) {
    float a = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float v = g[i];
        a = fmaf(v, v, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        atomicAdd(out, t);
    }
}

__global__ void __launch_bounds__(256) k_clip_adam(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, float lr, float b1, float b2,
                                                   float eps, float max_norm, float bc1, float bc2_sqrt,
                                                   const float* __restrict__ sumsq, const int32_t* __restrict__ step_dev) {
    float coef = 1.f;
    if (max_norm > 0.f) {
        float c = max_norm / (sqrtf(*sumsq) + 1e-6f);
        coef = c > 1.f ? 1.f : c;                    // torch.clamp(c, max=1.0): a NaN norm propagates into the gradients
    }
    if (step_dev) {                                  // replayable form: the step count is read from device memory
        const float t = (float)(*step_dev);          // already incremented for this step by k_sumsq_step
        bc1 = 1.f - powf(b1, t);
        bc2_sqrt = sqrtf(1.f - powf(b2, t));
    }
    const float step_size = lr / bc1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float gi = g[i] * coef;
        g[i] = gi;                                   // clip_grad_norm_ rescales the gradients in place
        float mi = m[i] + (1.f - b1) * (gi - m[i]);  // torch: exp_avg.lerp_(grad, 1 - beta1)
        float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] -= step_size * (mi / denom);
    }
}

int clip_adam(float* p, float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps, float max_norm,
              int step, int32_t* step_dev, float* scratch, cudaStream_t st) {
    if (n == 0) return DIGS_OK;
    int blocks = (int)((n + 255) / 256);
    const int cap = tcpath::num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (step_dev) {
        k_step_inc<<<1, 1, 0, st>>>(step_dev, scratch);
        count_launch();
    } else if (max_norm > 0.f) {
        cudaMemsetAsync(scratch, 0, sizeof(float), st);
    }
    if (max_norm > 0.f) {
        k_sumsq<<<blocks, 256, 0, st>>>(g, n, scratch);
        count_launch();
    }
    double bc1 = 1.0 - pow((double)b1, (double)(step > 0 ? step : 1));
    double bc2 = 1.0 - pow((double)b2, (double)(step > 0 ? step : 1));
    k_clip_adam<<<blocks, 256, 0, st>>>(p, g, m, v, n, lr, b1, b2, eps, max_norm, (float)bc1, (float)sqrt(bc2), scratch,
                                        step_dev);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("clip_adam launch: %s", cudaGetErrorString(e)); return DIGS_ECUDA; }
    return DIGS_OK;
}

}  // namespace digs
