// C-ABI of libdigs_b200.so (see include/digs_b200.h): argument checking and dispatch on precision mode.
//
// DIGS_MODE_FP32   every kernel is the fp32 CUDA-core implementation (jet_simt.cu).
// DIGS_MODE_BF16X2 / DIGS_MODE_BF16
//                  forward jet, value / grid query, reverse pass and weight gradients run on the tcgen05 kernels
//                  (jet_tc.cu, wgrad_tc.cu); only the per-point head adjoint reuses a small fp32 kernel.
#include "digs_internal.h"
#include <cstdarg>
#include <cstdio>
#include <atomic>

namespace digs {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static int check_net(const DigsNet* net) {
    if (!net) { set_error("net is NULL"); return DIGS_EINVAL; }
    if (net->d != 2 && net->d != 3) { set_error("net.d must be 2 or 3 (got %d)", net->d); return DIGS_EINVAL; }
    if (net->H < 64 || net->H > 1024 || (net->H % 64)) {
        set_error("net.H must be a multiple of 64 in [64,1024] (got %d)", net->H);
        return DIGS_EINVAL;
    }
    if (net->Lh < 1 || net->Lh + 2 > DIGS_MAX_LAYERS) {
        set_error("net.Lh must be in [1,%d] (got %d)", DIGS_MAX_LAYERS - 2, net->Lh);
        return DIGS_EINVAL;
    }
    if (net->head != DIGS_HEAD_IDENTITY && net->head != DIGS_HEAD_SQRT) { set_error("bad head"); return DIGS_EINVAL; }
    if (net->w0_stride < net->d) { set_error("w0_stride < d"); return DIGS_EINVAL; }
    for (int l = 0; l <= net->Lh + 1; ++l)
        if (!net->W[l] || !net->b[l]) { set_error("net.W[%d]/b[%d] is NULL", l, l); return DIGS_EINVAL; }
    return DIGS_OK;
}
// mode must be known; tensor-core modes additionally need a supported width (never a silent fallback)
static int check_mode(const DigsNet* net, int mode) {
    if (mode == DIGS_MODE_FP32) return DIGS_OK;
    if (mode == DIGS_MODE_BF16X2 || mode == DIGS_MODE_BF16) {
        if (!tcpath::supported(net)) {
            set_error("precision mode %d (tcgen05) needs H in {128,256,512}; got H=%d", mode, net->H);
            return DIGS_ENOTIMPL;
        }
        return DIGS_OK;
    }
    set_error("unknown precision mode %d", mode);
    return DIGS_EINVAL;
}
static inline bool is_tc(int mode) { return mode != DIGS_MODE_FP32; }
static inline int passes_of(int mode) { return mode == DIGS_MODE_BF16X2 ? 3 : 1; }

static PreSrc xyz_src(const DigsNet* net, const float* xyz, int64_t xyz_stride, const float* shape_bias, int64_t pps) {
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
}  // namespace digs

using namespace digs;

extern "C" {

int digs_abi_version(void) { return DIGS_ABI_VERSION; }
const char* digs_last_error(void) { return g_err; }
long long digs_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

size_t digs_saved_bytes(const DigsNet* net, int64_t n, int want_lap, int mode) {
    if (check_net(net) || check_mode(net, mode) || n < 0) return 0;
    if (is_tc(mode)) return tcpath::saved_bytes(net, n, want_lap);
    return simt::saved_bytes(net, n, want_lap);
}
size_t digs_forward_workspace_bytes(const DigsNet* net, int64_t n, int want_lap, int mode) {
    if (check_net(net) || check_mode(net, mode) || n < 0) return 0;
    if (is_tc(mode)) return tcpath::prep_ws_bytes(net);
    return simt::fwd_ws_bytes(net, n, want_lap);
}
size_t digs_backward_workspace_bytes(const DigsNet* net, int64_t n, int want_lap, int mode) {
    if (check_net(net) || check_mode(net, mode) || n < 0) return 0;
    if (is_tc(mode)) return tcpath::bwd_ws_bytes(net, n, want_lap);
    return simt::bwd_ws_bytes(net, n, want_lap);
}
size_t digs_value_workspace_bytes(const DigsNet* net, int64_t n, int mode) {
    if (check_net(net) || check_mode(net, mode) || n < 0) return 0;
    if (is_tc(mode)) return tcpath::prep_ws_bytes(net);
    return simt::value_ws_bytes(net, n);
}
size_t digs_grid_workspace_bytes(const DigsNet* net, int mode) {
    if (check_net(net) || check_mode(net, mode)) return 0;
    if (is_tc(mode)) return tcpath::prep_ws_bytes(net);
    return simt::grid_ws_bytes(net);
}

size_t digs_prepared_bytes(const DigsNet* net, int mode) {
    if (check_net(net) || check_mode(net, mode)) return 0;
    return is_tc(mode) ? tcpath::prep_ws_bytes(net) : 0;
}
int digs_prepare_weights(const DigsNet* net, void* prepared, size_t prepared_bytes, int mode, void* stream) {
    int rc;
    if ((rc = check_net(net)) || (rc = check_mode(net, mode))) return rc;
    if (!is_tc(mode)) return DIGS_OK;
    if (!prepared || prepared_bytes < tcpath::prep_ws_bytes(net)) { set_error("prepare_weights: buffer too small"); return DIGS_ENOSPACE; }
    return tcpath::run_prep(net, prepared, (cudaStream_t)stream);
}

size_t digs_step_workspace_bytes(const DigsNet* net, int64_t nn, int64_t nm, int want_lap, int mode) {
    if (check_net(net) || check_mode(net, mode) || nn < 0 || nm < 0 || !is_tc(mode)) return 0;
    return tcpath::step_ws_bytes(net, nn, nm, want_lap);
}
int digs_step_fused(const DigsNet* net, const float* nonmnfld, int64_t nn, const float* mnfld, int64_t nm,
                    const float* normals, int64_t xyz_stride, const float* shape_bias_n, const float* shape_bias_m,
                    int64_t pts_per_shape_n, int64_t pts_per_shape_m, int want_lap, const DigsLossCfg* cfg, float* f_n,
                    float* g_n, float* lap_n, float* f_m, float* g_m, float* terms_out, const DigsGrads* grads,
                    float* sb_grad_n, float* sb_grad_m, void* workspace, size_t workspace_bytes, int mode, void* stream) {
    int rc;
    if ((rc = check_net(net)) || (rc = check_mode(net, mode))) return rc;
    if (!is_tc(mode)) { set_error("step_fused: tensor-core modes only (use the two-call API in fp32 mode)"); return DIGS_ENOTIMPL; }
    if (nn < 0 || nm < 0 || xyz_stride < net->d) { set_error("step_fused: bad counts / xyz_stride"); return DIGS_EINVAL; }
    if (nn + nm == 0) return DIGS_OK;
    if (!cfg || !terms_out || !grads) { set_error("step_fused: NULL cfg / terms / grads"); return DIGS_EINVAL; }
    if (nn > 0 && (!nonmnfld || !f_n || !g_n)) { set_error("step_fused: NULL off-surface buffer"); return DIGS_EINVAL; }
    if (nm > 0 && (!mnfld || !f_m || !g_m)) { set_error("step_fused: NULL surface buffer"); return DIGS_EINVAL; }
    if (cfg->use_div && !want_lap) { set_error("step_fused: divergence loss needs want_lap"); return DIGS_EINVAL; }
    if (cfg->add_normal && !normals) { set_error("step_fused: normal term without normals"); return DIGS_EINVAL; }
    for (int l = 0; l <= net->Lh + 1; ++l)
        if (!grads->dW[l] || !grads->db[l]) { set_error("step_fused: grads.dW[%d]/db[%d] NULL", l, l); return DIGS_EINVAL; }
    if ((shape_bias_n && (!sb_grad_n || pts_per_shape_n <= 0)) || (shape_bias_m && (!sb_grad_m || pts_per_shape_m <= 0)) ||
        ((shape_bias_n != nullptr) != (shape_bias_m != nullptr) && nn > 0 && nm > 0)) {
        set_error("step_fused: shape_bias needs shape_bias_grad and pts_per_shape on both clouds");
        return DIGS_EINVAL;
    }
    size_t need = digs_step_workspace_bytes(net, nn, nm, want_lap, mode);
    if (!workspace || workspace_bytes < need) {
        set_error("step_fused: workspace too small (%zu < %zu)", workspace_bytes, need);
        return DIGS_ENOSPACE;
    }
    return tcpath::step_fused(net, nonmnfld, nn, mnfld, nm, normals, xyz_stride, shape_bias_n, shape_bias_m, pts_per_shape_n,
                              pts_per_shape_m, want_lap, cfg, f_n, g_n, lap_n, f_m, g_m, terms_out, grads, sb_grad_n,
                              sb_grad_m, workspace, passes_of(mode), (cudaStream_t)stream);
}

int digs_jet_forward(const DigsNet* net, const float* xyz, int64_t n, int64_t xyz_stride, const float* shape_bias,
                     int64_t pts_per_shape, int want_lap, float* f, float* grad, float* lap, void* saved,
                     size_t saved_bytes, void* workspace, size_t workspace_bytes, int mode, void* stream) {
    int rc;
    if ((rc = check_net(net)) || (rc = check_mode(net, mode))) return rc;
    if (n < 0 || xyz_stride < net->d) { set_error("jet_forward: bad n/xyz_stride"); return DIGS_EINVAL; }
    if (n == 0) return DIGS_OK;
    if (!xyz || !f || !grad || (want_lap && !lap)) { set_error("jet_forward: NULL buffer"); return DIGS_EINVAL; }
    if (shape_bias && pts_per_shape <= 0) { set_error("jet_forward: pts_per_shape <= 0"); return DIGS_EINVAL; }
    if (saved && saved_bytes < digs_saved_bytes(net, n, want_lap, mode)) {
        set_error("jet_forward: saved buffer too small (%zu < %zu)", saved_bytes, digs_saved_bytes(net, n, want_lap, mode));
        return DIGS_ENOSPACE;
    }
    size_t need = digs_forward_workspace_bytes(net, n, want_lap, mode);
    if (!workspace || workspace_bytes < need) {
        set_error("jet_forward: workspace too small (%zu < %zu)", workspace_bytes, need);
        return DIGS_ENOSPACE;
    }
    if (is_tc(mode))
        return tcpath::forward(net, xyz_src(net, xyz, xyz_stride, shape_bias, pts_per_shape), n, 1, want_lap, f, grad,
                               lap, saved, workspace, passes_of(mode), (cudaStream_t)stream);
    return simt::jet_forward(net, xyz, n, xyz_stride, shape_bias, pts_per_shape, want_lap, f, grad, lap, saved,
                             workspace, (cudaStream_t)stream);
}

int digs_jet_backward(const DigsNet* net, const float* xyz, int64_t n, int64_t xyz_stride, const float* shape_bias,
                      int64_t pts_per_shape, int want_lap, const float* f_bar, const float* grad_bar,
                      const float* lap_bar, const void* saved, size_t saved_bytes, void* workspace,
                      size_t workspace_bytes, const DigsGrads* grads, float* shape_bias_grad, int mode,
                      void* stream) {
    int rc;
    if ((rc = check_net(net)) || (rc = check_mode(net, mode))) return rc;
    if (n < 0 || xyz_stride < net->d) { set_error("jet_backward: bad n/xyz_stride"); return DIGS_EINVAL; }
    if (n == 0) return DIGS_OK;
    if (!xyz || !saved || !grads) { set_error("jet_backward: NULL buffer"); return DIGS_EINVAL; }
    for (int l = 0; l <= net->Lh + 1; ++l)
        if (!grads->dW[l] || !grads->db[l]) { set_error("jet_backward: grads.dW[%d]/db[%d] NULL", l, l); return DIGS_EINVAL; }
    if (shape_bias && (!shape_bias_grad || pts_per_shape <= 0)) {
        set_error("jet_backward: shape_bias given without shape_bias_grad/pts_per_shape");
        return DIGS_EINVAL;
    }
    if (saved_bytes < digs_saved_bytes(net, n, want_lap, mode)) { set_error("jet_backward: saved buffer too small"); return DIGS_ENOSPACE; }
    size_t need = digs_backward_workspace_bytes(net, n, want_lap, mode);
    if (!workspace || workspace_bytes < need) {
        set_error("jet_backward: workspace too small (%zu < %zu)", workspace_bytes, need);
        return DIGS_ENOSPACE;
    }
    if (is_tc(mode))
        return tcpath::backward(net, xyz_src(net, xyz, xyz_stride, shape_bias, pts_per_shape), n, want_lap, f_bar,
                                grad_bar, lap_bar, saved, workspace, grads, shape_bias ? shape_bias_grad : nullptr,
                                passes_of(mode), (cudaStream_t)stream);
    return simt::jet_backward(net, xyz, n, xyz_stride, shape_bias, pts_per_shape, want_lap, f_bar, grad_bar, lap_bar,
                              saved, workspace, grads, shape_bias ? shape_bias_grad : nullptr, (cudaStream_t)stream);
}

static int loss_fused_checked(const float* f_m, const float* g_m, int64_t nm, const float* f_n, const float* g_n,
                             const float* lap_n, int64_t nn, const float* normals, int d, const float* weights,
                             const float* weights_device, int add_normal, int use_div, int div_l2, float div_clamp,
                             int64_t Nm_total, int64_t Nn_total, float* fbar_m, float* gbar_m, float* fbar_n,
                             float* gbar_n, float* lbar_n, float* terms_out, void* stream) {
    if ((!weights && !weights_device) || !terms_out) { set_error("loss_fused: NULL weights/terms"); return DIGS_EINVAL; }
    if (nm < 0 || nn < 0) { set_error("loss_fused: negative count"); return DIGS_EINVAL; }
    if (nm > 0 && (!f_m || !g_m || !fbar_m || !gbar_m)) { set_error("loss_fused: NULL manifold buffer"); return DIGS_EINVAL; }
    if (nn > 0 && (!f_n || !g_n || !fbar_n || !gbar_n)) { set_error("loss_fused: NULL non-manifold buffer"); return DIGS_EINVAL; }
    if (use_div && nn > 0 && (!lap_n || !lbar_n)) { set_error("loss_fused: div loss without lap buffers"); return DIGS_EINVAL; }
    if (add_normal && !normals) { set_error("loss_fused: normal term without normals"); return DIGS_EINVAL; }
    return loss_fused(f_m, g_m, nm, f_n, g_n, lap_n, nn, normals, d, weights, add_normal, use_div, div_l2, div_clamp,
                      Nm_total, Nn_total, fbar_m, gbar_m, fbar_n, gbar_n, lbar_n, terms_out, weights_device,
                      (cudaStream_t)stream);
}

int digs_loss_fused(const float* f_m, const float* g_m, int64_t nm, const float* f_n, const float* g_n,
                    const float* lap_n, int64_t nn, const float* normals, int d, const float* weights, int add_normal,
                    int use_div, int div_l2, float div_clamp, int64_t Nm_total, int64_t Nn_total, float* fbar_m,
                    float* gbar_m, float* fbar_n, float* gbar_n, float* lbar_n, float* terms_out, void* stream) {
    return loss_fused_checked(f_m, g_m, nm, f_n, g_n, lap_n, nn, normals, d, weights, nullptr, add_normal, use_div, div_l2,
                              div_clamp, Nm_total, Nn_total, fbar_m, gbar_m, fbar_n, gbar_n, lbar_n, terms_out, stream);
}

int digs_loss_fused_dw(const float* f_m, const float* g_m, int64_t nm, const float* f_n, const float* g_n,
                       const float* lap_n, int64_t nn, const float* normals, int d, const float* weights_device,
                       int add_normal, int use_div, int div_l2, float div_clamp, int64_t Nm_total, int64_t Nn_total,
                       float* fbar_m, float* gbar_m, float* fbar_n, float* gbar_n, float* lbar_n, float* terms_out,
                       void* stream) {
    return loss_fused_checked(f_m, g_m, nm, f_n, g_n, lap_n, nn, normals, d, nullptr, weights_device, add_normal, use_div,
                              div_l2, div_clamp, Nm_total, Nn_total, fbar_m, gbar_m, fbar_n, gbar_n, lbar_n, terms_out,
                              stream);
}

int digs_value_forward(const DigsNet* net, const float* xyz, int64_t n, int64_t xyz_stride, const float* shape_bias,
                       int64_t pts_per_shape, float* f, void* workspace, size_t workspace_bytes, int mode,
                       void* stream) {
    int rc;
    if ((rc = check_net(net)) || (rc = check_mode(net, mode))) return rc;
    if (n < 0 || xyz_stride < net->d) { set_error("value_forward: bad n/xyz_stride"); return DIGS_EINVAL; }
    if (n == 0) return DIGS_OK;
    if (!xyz || !f) { set_error("value_forward: NULL buffer"); return DIGS_EINVAL; }
    size_t need = digs_value_workspace_bytes(net, n, mode);
    if (!workspace || workspace_bytes < need) { set_error("value_forward: workspace too small"); return DIGS_ENOSPACE; }
    PreSrc s = xyz_src(net, xyz, xyz_stride, shape_bias, pts_per_shape);
    if (is_tc(mode))
        return tcpath::forward(net, s, n, 0, 0, f, nullptr, nullptr, nullptr, workspace, passes_of(mode),
                               (cudaStream_t)stream);
    return simt::value_forward(net, s, n, f, workspace, (cudaStream_t)stream);
}

int digs_grid_query(const DigsNet* net, const uint16_t* x_fp16, int nx, const uint16_t* y_fp16, int ny,
                    const uint16_t* z_fp16, int nz, int iy_begin, int iy_end, const float* shape_bias, float* sdf_out,
                    void* workspace, size_t workspace_bytes, int mode, void* stream) {
    int rc;
    if ((rc = check_net(net)) || (rc = check_mode(net, mode))) return rc;
    if (net->d != 3) { set_error("grid_query: d must be 3"); return DIGS_EINVAL; }
    if (nx <= 0 || ny <= 0 || nz <= 0 || iy_begin < 0 || iy_end > ny || iy_begin > iy_end) {
        set_error("grid_query: bad grid extents");
        return DIGS_EINVAL;
    }
    if (iy_begin == iy_end) return DIGS_OK;
    if (!x_fp16 || !y_fp16 || !z_fp16 || !sdf_out) { set_error("grid_query: NULL buffer"); return DIGS_EINVAL; }
    size_t need = digs_grid_workspace_bytes(net, mode);
    if (!workspace || workspace_bytes < need) { set_error("grid_query: workspace too small"); return DIGS_ENOSPACE; }
    if (is_tc(mode)) {
        PreSrc s{};
        s.mode = 2;
        s.din = 3;
        s.shape_bias = shape_bias;
        s.pps = (int64_t)1 << 62;
        s.ax = reinterpret_cast<const __half*>(x_fp16);
        s.ay = reinterpret_cast<const __half*>(y_fp16);
        s.az = reinterpret_cast<const __half*>(z_fp16);
        s.nx = nx;
        s.nz = nz;
        s.grid_base = (int64_t)iy_begin * nx * nz;
        int64_t total = (int64_t)(iy_end - iy_begin) * nx * nz;
        return tcpath::forward(net, s, total, 0, 0, sdf_out, nullptr, nullptr, nullptr, workspace, passes_of(mode),
                               (cudaStream_t)stream);
    }
    return simt::grid_query(net, x_fp16, nx, y_fp16, ny, z_fp16, nz, iy_begin, iy_end, shape_bias, sdf_out, workspace,
                            (cudaStream_t)stream);
}

int digs_sample_batch(const float* cloud, const float* normals, int64_t n_cloud, int d, int64_t n_pts, float box,
                      uint64_t seed, uint64_t* iter_counter, float* mnfld, float* mnfld_normals, float* nonmnfld,
                      void* stream) {
    if (d != 2 && d != 3) { set_error("sample_batch: d must be 2 or 3"); return DIGS_EINVAL; }
    if (n_cloud <= 0 || n_pts < 0 || n_pts > n_cloud) { set_error("sample_batch: need 0 <= n_pts <= n_cloud"); return DIGS_EINVAL; }
    if (n_pts == 0) return DIGS_OK;
    if (!cloud || !iter_counter || !mnfld || !nonmnfld || (normals && !mnfld_normals)) {
        set_error("sample_batch: NULL buffer");
        return DIGS_EINVAL;
    }
    return sample_batch(cloud, normals, n_cloud, d, n_pts, box, seed, reinterpret_cast<unsigned long long*>(iter_counter),
                        mnfld, mnfld_normals, nonmnfld, (cudaStream_t)stream);
}

int digs_mt_classify(const float* vol, int n0, int n1, int n2, float level, uint8_t* edge_flags, uint8_t* tri_counts,
                     void* stream) {
    if (!vol || !edge_flags || !tri_counts) { set_error("mt_classify: NULL buffer"); return DIGS_EINVAL; }
    return mesh::classify(vol, n0, n1, n2, level, edge_flags, tri_counts, (cudaStream_t)stream);
}

int digs_mt_emit(const float* vol, int n0, int n1, int n2, float level, const int* axis_of, const float* origin,
                 const float* spacing, const uint8_t* edge_flags, const uint8_t* tri_counts, const int32_t* vert_offsets,
                 const int32_t* tri_offsets, float* verts, float* normals, int32_t* faces, void* stream) {
    if (!vol || !axis_of || !origin || !spacing || !edge_flags || !vert_offsets) { set_error("mt_emit: NULL buffer"); return DIGS_EINVAL; }
    if (faces && (!tri_counts || !tri_offsets)) { set_error("mt_emit: faces need tri_counts and tri_offsets"); return DIGS_EINVAL; }
    if (normals && !verts) { set_error("mt_emit: normals are written together with verts"); return DIGS_EINVAL; }
    return mesh::emit(vol, n0, n1, n2, level, axis_of, origin, spacing, edge_flags, tri_counts, vert_offsets, tri_offsets,
                      verts, normals, faces, (cudaStream_t)stream);
}

int digs_nn_query(const float* ref, int64_t n_ref, const float* query, int64_t n_query, int d, float* dist, int32_t* idx,
                  void* stream) {
    if (n_query < 0) { set_error("nn_query: negative n_query"); return DIGS_EINVAL; }
    if (n_query > 0 && (!ref || !query || !dist)) { set_error("nn_query: NULL buffer"); return DIGS_EINVAL; }
    return mesh::nn_query(ref, n_ref, query, n_query, d, dist, idx, (cudaStream_t)stream);
}

int digs_clip_adam(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                   float beta2, float eps, float max_norm, int step, float* scratch, void* stream) {
    if (n < 0 || step < 1) { set_error("clip_adam: bad n/step"); return DIGS_EINVAL; }
    if (n == 0) return DIGS_OK;
    if (!params || !grads || !exp_avg || !exp_avg_sq || (max_norm > 0.f && !scratch)) {
        set_error("clip_adam: NULL buffer");
        return DIGS_EINVAL;
    }
    return clip_adam(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, max_norm, step, nullptr, scratch,
                     (cudaStream_t)stream);
}

int digs_clip_adam_dev(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                       float beta2, float eps, float max_norm, int32_t* step_counter, float* scratch, void* stream) {
    if (n < 0 || !step_counter) { set_error("clip_adam_dev: bad n / NULL step counter"); return DIGS_EINVAL; }
    if (n == 0) return DIGS_OK;
    if (!params || !grads || !exp_avg || !exp_avg_sq || !scratch) { set_error("clip_adam_dev: NULL buffer"); return DIGS_EINVAL; }
    return clip_adam(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, max_norm, 0, step_counter, scratch,
                     (cudaStream_t)stream);
}

}  // extern "C"
