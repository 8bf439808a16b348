/*
 * digs_b200.h -- C-ABI of libdigs_b200.so: the B200-native DiGS hot path.
 *
 * Every entry point takes plain device pointers, sizes and a CUDA stream (as void*); there are no
 * torch types in any signature.  All buffers are BORROWED for the duration of the call and are
 * stream-ordered: the library never allocates or frees caller memory, never synchronises the
 * host, and is CUDA-graph capturable.  Return value: 0 on success, a negative DIGS_E* code on
 * failure, with a thread-local message available from digs_last_error().
 *
 * Reference interfaces replaced (paths relative to the DiGS repository):
 *   digs_jet_forward      FCBlock.forward (models/DiGS.py:122-134) applied to a point cloud, PLUS the
 *                         utils.gradient calls DiGSLoss makes on its output: first order
 *                         (models/losses.py:72,76; utils/utils.py:108-117) and the per-axis second
 *                         pass whose trace is the divergence (models/losses.py:100-106).
 *   digs_jet_backward     the part of loss.backward() (train_surface_reconstruction.py:128,
 *                         train_basic_shape.py:132, train_shapespace.py:145) that runs from
 *                         (dL/df, dL/dgrad f, dL/dlap f) down to the 2(Lh+2) parameter gradients.
 *   digs_loss_fused       DiGSLoss.forward's term evaluation (models/losses.py:107-184) and the
 *                         adjoint seeds autograd would derive from it.
 *   digs_value_forward    decoder(points) under no_grad (utils/utils.py:348,
 *                         compute_metrics_shapenet.py:134).
 *   digs_grid_query       the chunk loop of utils.implicit2mesh (utils/utils.py:343-350) over the
 *                         grid of utils.get_3d_grid (utils/utils.py:186-215), coordinates
 *                         generated on the device from the three fp16-rounded axes.
 *
 * Channel convention ("jet"): per point the kernels carry the value h, d Jacobian columns A_k and
 * (optionally) the Laplacian P through every layer: C = 1 + d + want_lap channels.
 */
#ifndef DIGS_B200_H_
#define DIGS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIGS_ABI_VERSION 2
#define DIGS_MAX_LAYERS 18 /* Lh + 2 <= 18 */

/* error codes */
#define DIGS_OK 0
#define DIGS_EINVAL (-1)  /* bad argument / unsupported configuration */
#define DIGS_ENOSPACE (-2) /* caller-provided workspace too small */
#define DIGS_ECUDA (-3)    /* CUDA runtime error at launch */
#define DIGS_ENOTIMPL (-4) /* precision mode not available for this configuration */

/* precision modes */
#define DIGS_MODE_FP32 0      /* CUDA-core fp32 FMAs, fp32 accumulate (reference-grade) */
#define DIGS_MODE_BF16X2 1    /* tcgen05 tensor cores, bf16 hi+lo split operands, 3 MMA passes */
#define DIGS_MODE_BF16 2      /* tcgen05 tensor cores, single-pass bf16 operands (stated looser bound) */

/* output head (models/DiGS.py:128-132) */
#define DIGS_HEAD_IDENTITY 0 /* init_type siren: f = u */
#define DIGS_HEAD_SQRT 1     /* init_type mfgi / geometric_sine: f = scale*(sign(u)sqrt(|u|+1e-8) - radius) */

/* Network description: nn.Linear layout, weight (out,in) row-major fp32, exactly the tensors
 * named decoder.fc_block.net.{i}.0.{weight,bias} (models/DiGS.py:88-99). */
typedef struct DigsNet {
    int32_t d;         /* spatial input dims: 2 or 3 */
    int32_t H;         /* hidden width, multiple of 64, <= 1024 */
    int32_t Lh;        /* number of HxH layers (decoder_n_hidden_layers), >= 1 */
    int32_t head;      /* DIGS_HEAD_* */
    float radius;      /* sphere_init_params[0] */
    float scale;       /* sphere_init_params[1] */
    int32_t w0_stride; /* row stride (floats) of W[0]: d + latent_size */
    int32_t reserved;
    const float* W[DIGS_MAX_LAYERS]; /* W[0] (H,w0_stride); W[1..Lh] (H,H); W[Lh+1] (1,H) */
    const float* b[DIGS_MAX_LAYERS]; /* b[0..Lh] (H); b[Lh+1] (1) */
    /* NULL, or digs_prepared_bytes() bytes filled by digs_prepare_weights() for THESE parameter values: the
     * tensor-core modes then skip their per-call operand preparation (bf16 hi/lo copies of 30 W and (30 W)^T).
     * The caller re-runs digs_prepare_weights after every parameter update (ABI 2). */
    const void* prepared;
} DigsNet;

/* Gradient destinations, same indexing as DigsNet.W/b.  Kernels ACCUMULATE (+=) into them.
 * Only the first d columns of dW[0] rows are written (latent columns are handled by the caller
 * through shape_bias_grad). */
typedef struct DigsGrads {
    float* dW[DIGS_MAX_LAYERS];
    float* db[DIGS_MAX_LAYERS];
} DigsGrads;

int digs_abi_version(void);
const char* digs_last_error(void);
/* Number of CUDA kernels this library has launched in this process (for bench.py's gpu_launches). */
long long digs_launch_count(void);

/* Bytes of the `saved` buffer digs_jet_forward fills for digs_jet_backward (0 if no backward wanted). */
size_t digs_saved_bytes(const DigsNet* net, int64_t n, int want_lap, int precision_mode);
/* Bytes of scratch digs_jet_forward / digs_jet_backward / digs_value_forward / digs_grid_query need. */
size_t digs_forward_workspace_bytes(const DigsNet* net, int64_t n, int want_lap, int precision_mode);
size_t digs_backward_workspace_bytes(const DigsNet* net, int64_t n, int want_lap, int precision_mode);
size_t digs_value_workspace_bytes(const DigsNet* net, int64_t n, int precision_mode);
size_t digs_grid_workspace_bytes(const DigsNet* net, int precision_mode);

/* Tensor-core operand preparation, hoisted out of the per-call path (DIGS_MODE_BF16X2 / DIGS_MODE_BF16; a no-op
 * returning 0 bytes in DIGS_MODE_FP32).  One launch; see DigsNet.prepared. */
size_t digs_prepared_bytes(const DigsNet* net, int precision_mode);
int digs_prepare_weights(const DigsNet* net, void* prepared, size_t prepared_bytes, int precision_mode, void* stream);

/* Forward jet of one point cloud.
 *  xyz          (n, xyz_stride) fp32, the first d columns are the coordinates.
 *  shape_bias   NULL, or (n_shapes, H): per-shape replacement for b[0] (= W0[:,d:] latent + b0,
 *               SURVEY 8a); point i belongs to shape i / pts_per_shape.
 *  f (n), grad (n,d), lap (n or NULL when !want_lap): outputs.
 *  saved        NULL (inference) or digs_saved_bytes() bytes. */
int digs_jet_forward(const DigsNet* net, const float* xyz, int64_t n, int64_t xyz_stride,
                     const float* shape_bias, int64_t pts_per_shape, int want_lap,
                     float* f, float* grad, float* lap,
                     void* saved, size_t saved_bytes, void* workspace, size_t workspace_bytes,
                     int precision_mode, void* stream);

/* Reverse pass.  f_bar (n), grad_bar (n,d), lap_bar (n or NULL): adjoint seeds.
 * Accumulates into grads->dW/db; shape_bias_grad (n_shapes,H) or NULL, accumulated. */
int digs_jet_backward(const DigsNet* net, const float* xyz, int64_t n, int64_t xyz_stride,
                      const float* shape_bias, int64_t pts_per_shape, int want_lap,
                      const float* f_bar, const float* grad_bar, const float* lap_bar,
                      const void* saved, size_t saved_bytes, void* workspace, size_t workspace_bytes,
                      const DigsGrads* grads, float* shape_bias_grad,
                      int precision_mode, void* stream);

/* Loss terms + adjoint seeds in one launch (in-scope loss types; SURVEY 8a "adjoint seeds").
 *  weights[6] = {sdf, inter, normal, eikonal, div, latent}; add_normal: loss type adds the normal
 *  term; use_div: loss type has 'div'; div_l2: 0 = l1, 1 = l2; normals may be NULL.
 *  Nm_total / Nn_total: GLOBAL point counts the means are taken over (>= local n when sharded).
 *  terms_out[8] (device): loss(partial, without latent term), sdf, inter, eikonal, normal, div, 0, 0
 *  -- accumulated with atomics, caller zeroes them first. */
int digs_loss_fused(const float* f_m, const float* g_m, int64_t nm, const float* f_n, const float* g_n,
                    const float* lap_n, int64_t nn, const float* normals, int d,
                    const float* weights, int add_normal, int use_div, int div_l2, float div_clamp,
                    int64_t Nm_total, int64_t Nn_total,
                    float* fbar_m, float* gbar_m, float* fbar_n, float* gbar_n, float* lbar_n,
                    float* terms_out, void* stream);

/* Same, with the six weights read from DEVICE memory at kernel time: for CUDA-graph replays, where the div weight
 * that update_div_weight (models/losses.py:190-228) rewrites every iteration must not be baked into the graph. */
int digs_loss_fused_dw(const float* f_m, const float* g_m, int64_t nm, const float* f_n, const float* g_n,
                       const float* lap_n, int64_t nn, const float* normals, int d,
                       const float* weights_device, int add_normal, int use_div, int div_l2, float div_clamp,
                       int64_t Nm_total, int64_t Nn_total,
                       float* fbar_m, float* gbar_m, float* fbar_n, float* gbar_n, float* lbar_n,
                       float* terms_out, void* stream);

/* ---- The whole hot-path step of one batch as TWO launches (tensor-core modes) ---------------------------------------
 * Replaces, for one iteration, train_surface_reconstruction.py:121-128 (train_basic_shape.py:105-132,
 * train_shapespace.py:130-145): DiGSNetwork.forward on both clouds, DiGSLoss.forward and loss.backward().
 * Every DiGS loss term is a mean of per-point functions (models/losses.py:107-168), so a point's adjoint seeds need only
 * its own (f, grad f, lap f) and global constants: a persistent kernel walks point tiles and runs, per tile,
 *   forward jet -> output head -> loss terms + adjoint seeds -> reverse jet
 * with the tile's pre-activations parked in a per-CTA scratch ring that stays in L2; only the two operand streams of the
 * weight-gradient GEMM (layer inputs, adjoints) reach HBM, and one GEMM launch over both clouds finishes the step.
 *   nonmnfld (nn, xyz_stride) / mnfld (nm, xyz_stride): the first d columns are coordinates; normals (nm, d) or NULL.
 *   shape_bias_n / shape_bias_m: NULL, or (n_shapes, H) per-shape layer-0 bias (auto-decoder), points_per_shape each;
 *   sb_grad_n / sb_grad_m (n_shapes, H) accumulate its gradient.
 *   cfg: loss configuration (weights on the host or, for CUDA-graph replays, in device memory; global point counts).
 *   outputs f_n (nn), g_n (nn,d), lap_n (nn or NULL), f_m (nm), g_m (nm,d); terms_out[8] as digs_loss_fused (accumulated).
 *   grads: accumulated (+=) like digs_jet_backward.
 * The GEMM is a programmatic dependent launch where that pays: the persistent tile kernel leaves part of the SMs idle in its
 * last round of tiles, the GEMM's CTAs start there and synchronise on per-round completion counters in the workspace (reset by
 * a memset on `stream` in front of the tile kernel) instead of on the whole grid; work queued on `stream` after this call sees
 * both launches complete.  Results do not depend on it (DIGS_B200_WGRAD_OVERLAP=0 forces plain stream order).
 * DIGS_MODE_FP32 returns DIGS_ENOTIMPL (the two-call API covers it). */
typedef struct DigsLossCfg {
    float weights[6];            /* sdf, inter, normal, eikonal, div, latent (host copy) */
    const float* weights_device; /* NULL, or 6 floats in device memory read at kernel time (overrides weights) */
    int32_t add_normal;          /* loss type adds the normal term (siren, siren_w_div) */
    int32_t use_div;             /* loss type has 'div' */
    int32_t div_l2;              /* 0 = l1, 1 = l2 */
    float div_clamp;
    int64_t Nm_total, Nn_total;  /* GLOBAL point counts of the means (>= local counts when sharded over ranks) */
} DigsLossCfg;
size_t digs_step_workspace_bytes(const DigsNet* net, int64_t nn, int64_t nm, int want_lap, int precision_mode);
int digs_step_fused(const DigsNet* net, const float* nonmnfld, int64_t nn, const float* mnfld, int64_t nm,
                    const float* normals, int64_t xyz_stride,
                    const float* shape_bias_n, const float* shape_bias_m, int64_t pts_per_shape_n, int64_t pts_per_shape_m,
                    int want_lap, const DigsLossCfg* cfg,
                    float* f_n, float* g_n, float* lap_n, float* f_m, float* g_m, float* terms_out,
                    const DigsGrads* grads, float* sb_grad_n, float* sb_grad_m,
                    void* workspace, size_t workspace_bytes, int precision_mode, void* stream);

/* decoder(points) -> (n) values, no derivative channels, nothing saved. */
int digs_value_forward(const DigsNet* net, const float* xyz, int64_t n, int64_t xyz_stride,
                       const float* shape_bias, int64_t pts_per_shape, float* f,
                       void* workspace, size_t workspace_bytes, int precision_mode, void* stream);

/* Dense grid query.  Axes are IEEE fp16 bit patterns (utils/utils.py:208-211 rounds the grid to
 * fp16).  Flat output order is [iy][ix][iz] (np.meshgrid indexing='xy', utils/utils.py:208,349);
 * this call writes rows iy in [iy_begin, iy_end) to sdf_out + (iy - iy_begin)*nx*nz. */
int digs_grid_query(const DigsNet* net, const uint16_t* x_fp16, int nx, const uint16_t* y_fp16, int ny,
                    const uint16_t* z_fp16, int nz, int iy_begin, int iy_end,
                    const float* shape_bias, float* sdf_out,
                    void* workspace, size_t workspace_bytes, int precision_mode, void* stream);

/* One iteration's batch, sampled on the device like ReconDataset.__getitem__ (recon_dataset.py:143-174) does on the host:
 * n_pts distinct cloud points (+ normals, may be NULL) as the first n_pts entries of a fresh pseudo-random permutation,
 * and n_pts points uniform in [-box, box]^d.  iter_counter: TWO device uint64 (iteration, ticket), zero-initialised by
 * the caller; the kernel reads the iteration, derives that iteration's key from (seed, iteration) and increments it, so
 * the call is CUDA-graph replayable.  The distribution, not numpy's RNG stream, is reproduced. */
int digs_sample_batch(const float* cloud, const float* normals, int64_t n_cloud, int d, int64_t n_pts, float box,
                      uint64_t seed, uint64_t* iter_counter, float* mnfld, float* mnfld_normals, float* nonmnfld,
                      void* stream);

/* Optimizer tail over flat buffers: clip_grad_norm_(params, max_norm) (train_surface_reconstruction.py:130; the loops
 * hard-code 10.) followed by Adam.step() (train_surface_reconstruction.py:63,132).  grads are rescaled in place like
 * clip_grad_norm_ does (a NaN / Inf norm propagates into them, as in torch); max_norm <= 0 skips clipping
 * (train_basic_shape.py:132-133 does not clip).  step is the 1-based Adam step count; scratch is one device float. */
int digs_clip_adam(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                   float beta2, float eps, float max_norm, int step, float* scratch, void* stream);
/* Same, CUDA-graph replayable: the Adam step count lives in device memory (step_counter: one int32, zero-initialised
 * by the caller); the kernel increments it and derives the bias corrections itself. */
int digs_clip_adam_dev(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                       float beta2, float eps, float max_norm, int32_t* step_counter, float* scratch, void* stream);

/* Iso-surface extraction on a dense fp32 volume in memory order (n0, n1, n2), n2 fastest: replaces
 * skimage.measure.marching_cubes(volume, level, spacing) + the origin shift of utils/utils.py:354-358.  Vertices sit on
 * the grid edges whose endpoint values straddle `level`, at the linear interpolation point (the marching-cubes vertex
 * rule); cells are triangulated by marching tetrahedra on the Kuhn split (watertight, no case tables), so edges also
 * include face and body diagonals.
 *   digs_mt_classify: edge_flags[p] = 7-bit mask of straddling edges that start at grid point p (bit m-1 <-> offset
 *     ((m>>2)&1, (m>>1)&1, m&1) along (axis0, axis1, axis2)); tri_counts[p] = triangles of the cell starting at p.
 *   The caller forms EXCLUSIVE prefix sums vert_offsets = scan(popcount(edge_flags)), tri_offsets = scan(tri_counts).
 *   digs_mt_emit: verts (V,3), unit normals (V,3; NULL to skip; normalised volume gradient, pointing to larger values),
 *     faces (F,3) int32 vertex ids, oriented so the face normal points to larger values.  Memory axis k is logical axis
 *     axis_of[k] (a permutation of 0,1,2; the reference's volume is [iy][ix][iz] -> {1,0,2}); position along it is
 *     origin[k] + spacing[k] * index.  axis_of / origin / spacing are HOST arrays of 3. */
int digs_mt_classify(const float* vol, int n0, int n1, int n2, float level, uint8_t* edge_flags, uint8_t* tri_counts,
                     void* stream);
int digs_mt_emit(const float* vol, int n0, int n1, int n2, float level, const int* axis_of, const float* origin,
                 const float* spacing, const uint8_t* edge_flags, const uint8_t* tri_counts, const int32_t* vert_offsets,
                 const int32_t* tri_offsets, float* verts, float* normals, int32_t* faces, void* stream);

/* Exact nearest neighbour of every query point in a reference cloud (d = 2 or 3, row-major fp32): dist[i] = Euclidean
 * distance, idx[i] = index of the neighbour (idx may be NULL).  Replaces KDTree(ref).query(query) of
 * surface_reconstruction/compute_metrics_srb.py:38-42 and utils/utils.py:250-251,274-275. */
int digs_nn_query(const float* ref, int64_t n_ref, const float* query, int64_t n_query, int d, float* dist, int32_t* idx,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIGS_B200_H_ */
