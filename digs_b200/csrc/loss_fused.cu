// DiGS loss terms + adjoint seeds in one launch.
//
// Restates models/losses.py:107-184 for the in-scope loss types (siren, siren_wo_n, siren_w_div,
// siren_wo_n_w_div; div_type l1|l2) given the jet outputs (f, grad f, lap f), and emits the seeds
// (dL/df, dL/dgrad f, dL/dlap f) that autograd would have propagated back into the network
// (SURVEY.md section 8a "adjoint seeds").  Means are over GLOBAL point counts so that a SUM
// all-reduce over point shards reproduces the single-GPU loss and gradient exactly.
#include "digs_internal.h"

namespace digs {

struct LossArgs {
    const float *f_m, *g_m, *f_n, *g_n, *lap_n, *normals;
    float *fbar_m, *gbar_m, *fbar_n, *gbar_n, *lbar_n, *terms;
    int64_t nm, nn;
    float w_sdf, w_inter, w_normal, w_eik, w_div;
    const float* w_dev;   // optional DEVICE copy of the weights (CUDA-graph replays re-read it every launch)
    float inv_nm, inv_nn, inv_nall;
    float div_clamp;
    int add_normal, use_div, div_l2;
};

__device__ __forceinline__ float sgnf(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

template <int D>
__global__ void __launch_bounds__(256) k_loss(LossArgs a) {
    if (a.w_dev) {
        a.w_sdf = a.w_dev[0]; a.w_inter = a.w_dev[1]; a.w_normal = a.w_dev[2]; a.w_eik = a.w_dev[3]; a.w_div = a.w_dev[4];
    }
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float t_sdf = 0.f, t_inter = 0.f, t_eik = 0.f, t_nrm = 0.f, t_div = 0.f;
    if (i < a.nm) {
        float f = a.f_m[i];
        float g[D], nr2 = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) { g[j] = a.g_m[i * D + j]; nr2 = fmaf(g[j], g[j], nr2); }
        float nr = sqrtf(nr2);
        t_sdf = fabsf(f) * a.inv_nm;
        t_eik += fabsf(nr - 1.f) * a.inv_nall;
        a.fbar_m[i] = a.w_sdf * a.inv_nm * sgnf(f);
        float ke = (nr > 0.f) ? a.w_eik * a.inv_nall * sgnf(nr - 1.f) / nr : 0.f;
        float gb[D];
#pragma unroll
        for (int j = 0; j < D; ++j) gb[j] = ke * g[j];
        if (a.normals) {
            float nv[D], nn2 = 0.f, dot = 0.f;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                nv[j] = a.normals[i * D + j];
                nn2 = fmaf(nv[j], nv[j], nn2);
                dot = fmaf(g[j], nv[j], dot);
            }
            float gc = fmaxf(nr, 1e-8f), nc = fmaxf(sqrtf(nn2), 1e-8f);
            float cs = dot / (gc * nc);
            t_nrm = (1.f - fabsf(cs)) * a.inv_nm;
            if (a.add_normal) {
                float k = -a.w_normal * a.inv_nm * sgnf(cs);
#pragma unroll
                for (int j = 0; j < D; ++j) gb[j] += k * (nv[j] / (gc * nc) - cs / (gc * gc) * g[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < D; ++j) a.gbar_m[i * D + j] = gb[j];
    }
    if (i < a.nn) {
        float f = a.f_n[i];
        float g[D], nr2 = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) { g[j] = a.g_n[i * D + j]; nr2 = fmaf(g[j], g[j], nr2); }
        float nr = sqrtf(nr2);
        float e = expf(-100.f * fabsf(f));
        t_inter = e * a.inv_nn;
        t_eik += fabsf(nr - 1.f) * a.inv_nall;
        a.fbar_n[i] = -100.f * a.w_inter * a.inv_nn * sgnf(f) * e;
        float ke = (nr > 0.f) ? a.w_eik * a.inv_nall * sgnf(nr - 1.f) / nr : 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) a.gbar_n[i * D + j] = ke * g[j];
        if (a.use_div) {
            float l = a.lap_n[i];
            bool isn = (l != l);
            if (isn) l = 0.f;  // losses.py:107
            float t = a.div_l2 ? l * l : fabsf(l);
            float dt = a.div_l2 ? 2.f * l : sgnf(l);
            t_div = fminf(fmaxf(t, 0.1f), a.div_clamp) * a.inv_nn;
            bool inside = (t >= 0.1f) && (t <= a.div_clamp) && !isn;
            a.lbar_n[i] = inside ? a.w_div * a.inv_nn * dt : 0.f;
        } else if (a.lbar_n) {
            a.lbar_n[i] = 0.f;
        }
    }
    float v[5] = {t_sdf, t_inter, t_eik, t_nrm, t_div};
    __shared__ float red[5][8];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
        // terms: [0] loss [1] sdf [2] inter [3] eikonal [4] normal [5] div
        const int slot[5] = {1, 2, 3, 4, 5};
        atomicAdd(a.terms + slot[threadIdx.x], t);
        float wt;
        switch (threadIdx.x) {
            case 0: wt = a.w_sdf; break;
            case 1: wt = a.w_inter; break;
            case 2: wt = a.w_eik; break;
            case 3: wt = a.add_normal ? a.w_normal : 0.f; break;
            default: wt = a.use_div ? a.w_div : 0.f; break;
        }
        atomicAdd(a.terms, wt * t);
    }
}

int loss_fused(const float* f_m, const float* g_m, int64_t nm, const float* f_n, const float* g_n,
               const float* lap_n, int64_t nn, const float* normals, int d, const float* weights, int add_normal,
               int use_div, int div_l2, float div_clamp, int64_t Nm_total, int64_t Nn_total, float* fbar_m,
               float* gbar_m, float* fbar_n, float* gbar_n, float* lbar_n, float* terms_out,
               const float* weights_device, cudaStream_t st) {
    LossArgs a;
    a.w_dev = weights_device;
    static const float zero_w[6] = {0, 0, 0, 0, 0, 0};
    if (!weights) weights = zero_w;
    a.f_m = f_m; a.g_m = g_m; a.f_n = f_n; a.g_n = g_n; a.lap_n = lap_n; a.normals = normals;
    a.fbar_m = fbar_m; a.gbar_m = gbar_m; a.fbar_n = fbar_n; a.gbar_n = gbar_n; a.lbar_n = lbar_n;
    a.terms = terms_out;
    a.nm = nm; a.nn = nn;
    a.w_sdf = weights[0]; a.w_inter = weights[1]; a.w_normal = weights[2]; a.w_eik = weights[3]; a.w_div = weights[4];
    a.inv_nm = Nm_total > 0 ? 1.f / (float)Nm_total : 0.f;
    a.inv_nn = Nn_total > 0 ? 1.f / (float)Nn_total : 0.f;
    a.inv_nall = (Nm_total + Nn_total) > 0 ? 1.f / (float)(Nm_total + Nn_total) : 0.f;
    a.div_clamp = div_clamp;
    a.add_normal = add_normal; a.use_div = use_div; a.div_l2 = div_l2;
    int64_t m = nm > nn ? nm : nn;
    if (m == 0) return DIGS_OK;
    unsigned blocks = (unsigned)((m + 255) / 256);
    if (d == 3) { k_loss<3><<<blocks, 256, 0, st>>>(a); }
    else if (d == 2) { k_loss<2><<<blocks, 256, 0, st>>>(a); }
    else { set_error("loss_fused: d must be 2 or 3"); return DIGS_EINVAL; }
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("loss_fused launch: %s", cudaGetErrorString(e)); return DIGS_ECUDA; }
    return DIGS_OK;
}

}  // namespace digs
