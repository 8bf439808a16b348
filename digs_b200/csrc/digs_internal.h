// Internal declarations shared by the translation units of libdigs_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/digs_b200.h"

namespace digs {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);  // process-wide count of kernels launched by this library

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Where the layer-input jets come from.
//   mode 0: stored pre-activation planes [C][n][H] (z, Z_0..Z_{d-1}, Q)
//   mode 1: layer 0 evaluated on the fly from point coordinates
//   mode 2: layer 0 evaluated on the fly from grid indices + three fp16 axes
struct PreSrc {
    int mode;
    int din;                  // spatial dims used by layer 0
    const float* planes;      // mode 0
    int64_t plane_stride;     // n*H
    const float* xyz;         // mode 1
    int64_t xyz_stride;
    const float* W0;          // modes 1,2: (H, w0_stride)
    int w0_stride;
    const float* b0;          // (H)
    const float* shape_bias;  // NULL or (n_shapes,H), replaces b0
    int64_t pps;              // points per shape
    const __half* ax;         // mode 2
    const __half* ay;
    const __half* az;
    int nx, nz;
    int64_t grid_base;        // flat index of local point 0 in [iy][ix][iz] order
};

// ---- fp32 CUDA-core path (jet_simt.cu) ------------------------------------------------------
namespace simt {
size_t saved_bytes(const DigsNet* net, int64_t n, int want_lap);
size_t fwd_ws_bytes(const DigsNet* net, int64_t n, int want_lap);
size_t bwd_ws_bytes(const DigsNet* net, int64_t n, int want_lap);
size_t value_ws_bytes(const DigsNet* net, int64_t n);
size_t grid_ws_bytes(const DigsNet* net);
int jet_forward(const DigsNet* net, const float* xyz, int64_t n, int64_t xyz_stride, const float* shape_bias,
                int64_t pps, int want_lap, float* f, float* grad, float* lap, void* saved, void* ws,
                cudaStream_t st);
int jet_backward(const DigsNet* net, const float* xyz, int64_t n, int64_t xyz_stride, const float* shape_bias,
                 int64_t pps, int want_lap, const float* f_bar, const float* g_bar, const float* l_bar,
                 const void* saved, void* ws, const DigsGrads* grads, float* shape_bias_grad, cudaStream_t st);
int value_forward(const DigsNet* net, const PreSrc& src0, int64_t n, float* f, void* ws, cudaStream_t st);
int grid_query(const DigsNet* net, const uint16_t* xh, int nx, const uint16_t* yh, int ny, const uint16_t* zh,
               int nz, int iy_begin, int iy_end, const float* shape_bias, float* out, void* ws, cudaStream_t st);
// head adjoint per point: (f_bar, g_bar, l_bar) + head_saved -> ubar [n][C]; db_last += sum u_bar
int head_seed(const DigsNet* net, int64_t n, int want_lap, const float* head_saved, const float* f_bar,
              const float* g_bar, const float* l_bar, float* ubar, float* db_last, cudaStream_t st);
}  // namespace simt

// ---- tcgen05 path (jet_tc.cu, wgrad_tc.cu) ----------------------------------------------------
namespace tcpath {
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn get_encode();   // cuTensorMapEncodeTiled through cudaGetDriverEntryPoint (no link-time libcuda dependency)
int num_sms();
struct SavedLayout {     // what the tensor-core forward leaves for the reverse pass
    int C;
    int64_t n_pad;
    size_t planes_bytes, act_bytes, head_bytes;
};
SavedLayout saved_layout(const DigsNet* net, int64_t n, int want_lap);
bool supported(const DigsNet* net);
size_t prep_ws_bytes(const DigsNet* net);
size_t saved_bytes(const DigsNet* net, int64_t n, int want_lap);
size_t bwd_ws_bytes(const DigsNet* net, int64_t n, int want_lap);
// want_jet = 0: value channel only.  saved: NULL or saved_bytes() bytes.  passes: 3 (bf16x2) or 1 (bf16).
int forward(const DigsNet* net, const PreSrc& src, int64_t n, int want_jet, int want_lap, float* f, float* grad,
            float* lap, void* saved, void* ws, int passes, cudaStream_t st);
int backward(const DigsNet* net, const PreSrc& src, int64_t n, int want_lap, const float* f_bar, const float* g_bar,
             const float* l_bar, const void* saved, void* ws, const DigsGrads* grads, float* sb_grad, int passes,
             cudaStream_t st);
size_t step_ws_bytes(const DigsNet* net, int64_t nn, int64_t nm, int want_lap);
int run_prep(const DigsNet* net, void* dst, cudaStream_t st);
int step_fused(const DigsNet* net, const float* nonmnfld, int64_t nn, const float* mnfld, int64_t nm, const float* normals,
               int64_t xyz_stride, const float* sb_n, const float* sb_m, int64_t pps_n, int64_t pps_m, int want_lap,
               const DigsLossCfg* cfg, float* f_n, float* g_n, float* lap_n, float* f_m, float* g_m, float* terms,
               const DigsGrads* grads, float* sbg_n, float* sbg_m, void* ws, int passes, cudaStream_t st);
// Tile-round completion counters of the fused step: lets the weight-gradient GEMM start (programmatic dependent launch)
// on the SMs whose tile-kernel CTA has already exited, while the other CTAs still work on the last round of tiles.
constexpr int WG_MAX_ROUNDS = 1023;             // ctr[0 .. rounds): tiles finished per round; ctr[WG_MAX_ROUNDS]: CTAs done
constexpr int WG_QUEUE_IDX = WG_MAX_ROUNDS - 1; // ctr[WG_QUEUE_IDX]: next work unit of the weight-gradient GEMM (rounds use [0, WG_QUEUE_IDX))
#ifdef DIGS_DIAG_TRACE                          // make DIAG=TRACE: globaltimer stamps of both grids behind the counters (scripts/overlap_trace.py)
constexpr size_t WG_CTR_BYTES = (WG_MAX_ROUNDS + 1) * sizeof(unsigned int) + 8192 * 8;
#else
constexpr size_t WG_CTR_BYTES = (WG_MAX_ROUNDS + 1) * sizeof(unsigned int);
#endif
struct WgradSync {
    unsigned int* ctr;        // NULL: plain launch, no flags
    int grid;                 // CTAs of the tile kernel (tile t belongs to round t / grid)
    int64_t total_tiles;
    int64_t tiles0, cols0;    // cloud 0: tiles, operand columns
    int ct0, ct1;             // operand columns per tile of cloud 0 / cloud 1
};
int wgrad(const DigsNet* net, const __nv_bfloat16* GP_words, const __nv_bfloat16* actP, int C, int64_t n_pad,
          const DigsGrads* grads, int passes, cudaStream_t st, const WgradSync* sync = nullptr);
}  // namespace tcpath

// ---- fused loss (loss_fused.cu) -------------------------------------------------------------
int loss_fused(const float* f_m, const float* g_m, int64_t nm, const float* f_n, const float* g_n,
               const float* lap_n, int64_t nn, const float* normals, int d, const float* weights,
               int add_normal, int use_div, int div_l2, float div_clamp, int64_t Nm_total, int64_t Nn_total,
               float* fbar_m, float* gbar_m, float* fbar_n, float* gbar_n, float* lbar_n, float* terms_out,
               const float* weights_device, cudaStream_t st);

// ---- on-device batch sampling (sampler.cu) -----------------------------------------------------
int sample_batch(const float* cloud, const float* normals, int64_t n_cloud, int d, int64_t n_pts, float box, uint64_t seed,
                 unsigned long long* iter_dev, float* mnfld, float* nrm, float* nonmnfld, cudaStream_t st);

// ---- optimizer tail (optim.cu) ------------------------------------------------------------------
int clip_adam(float* p, float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps, float max_norm,
              int step, int32_t* step_dev, float* scratch, cudaStream_t st);

// ---- iso-surface extraction and nearest neighbours (mesh.cu) ------------------------------------
namespace mesh {
int classify(const float* vol, int n0, int n1, int n2, float level, uint8_t* flags, uint8_t* ntri, cudaStream_t st);
int emit(const float* vol, int n0, int n1, int n2, float level, const int* axis_of, const float* origin, const float* spacing,
         const uint8_t* flags, const uint8_t* ntri, const int32_t* voff, const int32_t* toff, float* verts, float* normals,
         int32_t* faces, cudaStream_t st);
int nn_query(const float* ref, int64_t nr, const float* qry, int64_t nq, int d, float* dist, int32_t* idx, cudaStream_t st);
}  // namespace mesh

}  // namespace digs
