// Iso-surface extraction and nearest-neighbour distances on the device.
//
// Reference call sites replaced:
//   * utils/utils.py:354-358  skimage.measure.marching_cubes on the (nx, ny, nz) SDF volume, then the
//     spacing / origin / scale / translate of the vertices;
//   * surface_reconstruction/compute_metrics_srb.py:38-57  scipy KDTree.query both ways (Chamfer / Hausdorff).
// skimage's Lewiner marching cubes is third-party code that is not part of the reference tree ("parity unpinned",
// SURVEY.md section 8c).  What is reproduced exactly is its vertex rule: every grid edge whose endpoint values straddle
// the level gets one vertex at the linear interpolation point.  The surface inside a cell is triangulated by
// marching TETRAHEDRA on the Kuhn split of the cell (6 tetrahedra around the main diagonal), which needs no case
// tables, is translation invariant (faces of neighbouring cells match, the mesh is watertight) and adds vertices
// on face / body diagonals.  Vertices are unique: each one belongs to the grid point at the low end of its edge.
//
// Three passes over a dense volume in MEMORY order (n0, n1, n2), n2 fastest:
//   classify   per grid point: 7-bit mask of straddling edges (offsets 1..7 = bit pattern over the three axes) and the
//              number of triangles of the cell that starts there;
//   (caller)   exclusive prefix sums of popcount(mask) and of the triangle counts (any scan; torch.cumsum in the host layer);
//   vertices   per flagged edge: position (origin + spacing * index, mapped to logical x/y/z) and a unit normal from the
//              central-difference gradient of the volume;
//   faces      per cell: triangles as vertex-id triples, oriented so that the normal points to larger values.
#include "digs_internal.h"

namespace digs {
namespace mesh {

struct Vol {
    const float* v;
    int n0, n1, n2;
    float level;
};
struct Frame {        // memory axis k -> logical axis ax[k]; logical position = org[k] + sp[k] * index_k
    int ax[3];
    float org[3], sp[3];
};

__device__ __forceinline__ int64_t lin(const Vol& V, int a0, int a1, int a2) { return ((int64_t)a0 * V.n1 + a1) * V.n2 + a2; }
__device__ __forceinline__ int o0(int m) { return (m >> 2) & 1; }
__device__ __forceinline__ int o1(int m) { return (m >> 1) & 1; }
__device__ __forceinline__ int o2(int m) { return m & 1; }

// the six Kuhn tetrahedra of a cell: corners 0, a, a|b, 7 for the ordered axis pairs (a, b)
__constant__ int c_tet_a[6] = {1, 1, 2, 2, 4, 4};
__constant__ int c_tet_b[6] = {2, 4, 1, 4, 1, 2};

__global__ void __launch_bounds__(256) k_mt_classify(Vol V, uint8_t* __restrict__ flags, uint8_t* __restrict__ ntri) {
    const int64_t total = (int64_t)V.n0 * V.n1 * V.n2;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int a2 = (int)(p % V.n2);
        const int64_t t = p / V.n2;
        const int a1 = (int)(t % V.n1), a0 = (int)(t / V.n1);
        const bool in0 = a0 + 1 < V.n0, in1 = a1 + 1 < V.n1, in2 = a2 + 1 < V.n2;
        const bool s0 = V.v[p] < V.level;
        unsigned inside = s0 ? 1u : 0u, fl = 0;
#pragma unroll
        for (int m = 1; m < 8; ++m) {
            const bool ok = (!o0(m) || in0) && (!o1(m) || in1) && (!o2(m) || in2);
            if (ok) {
                const bool sm = V.v[lin(V, a0 + o0(m), a1 + o1(m), a2 + o2(m))] < V.level;
                inside |= (sm ? 1u : 0u) << m;
                if (sm != s0) fl |= 1u << (m - 1);
            }
        }
        unsigned nt = 0;
        if (in0 && in1 && in2 && inside != 0u && inside != 0xFFu) {
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                const int a = c_tet_a[q], b = c_tet_b[q];
                const int k = (int)(inside & 1u) + (int)((inside >> a) & 1u) + (int)((inside >> (a | b)) & 1u) + (int)((inside >> 7) & 1u);
                nt += (k == 2) ? 2u : ((k == 1 || k == 3) ? 1u : 0u);
            }
        }
        flags[p] = (uint8_t)fl;
        ntri[p] = (uint8_t)nt;
    }
}

// central (one-sided at the border) difference of the volume along memory axis `ax` at (a0, a1, a2), per index step
__device__ __forceinline__ float ddx(const Vol& V, int a0, int a1, int a2, int ax) {
    int n = (ax == 0) ? V.n0 : ((ax == 1) ? V.n1 : V.n2);
    int a = (ax == 0) ? a0 : ((ax == 1) ? a1 : a2);
    int lo = a > 0 ? a - 1 : a, hi = a + 1 < n ? a + 1 : a;
    int l0 = a0, l1 = a1, l2 = a2, h0 = a0, h1 = a1, h2 = a2;
    if (ax == 0) { l0 = lo; h0 = hi; } else if (ax == 1) { l1 = lo; h1 = hi; } else { l2 = lo; h2 = hi; }
    float d = V.v[lin(V, h0, h1, h2)] - V.v[lin(V, l0, l1, l2)];
    return (hi > lo) ? d / (float)(hi - lo) : 0.f;
}

__global__ void __launch_bounds__(256) k_mt_vertices(Vol V, Frame F, const uint8_t* __restrict__ flags,
                                                      const int32_t* __restrict__ voff, float* __restrict__ verts,
                                                      float* __restrict__ normals) {
    const int64_t total = (int64_t)V.n0 * V.n1 * V.n2;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const unsigned fl = flags[p];
        if (!fl) continue;
        const int a2 = (int)(p % V.n2);
        const int64_t t = p / V.n2;
        const int a1 = (int)(t % V.n1), a0 = (int)(t / V.n1);
        const float v0 = V.v[p];
        float g0[3];
        if (normals) {
#pragma unroll
            for (int k = 0; k < 3; ++k) g0[k] = ddx(V, a0, a1, a2, k);
        }
        int64_t vid = voff[p];
        for (int m = 1; m < 8; ++m) {
            if (!((fl >> (m - 1)) & 1u)) continue;
            const int b0 = a0 + o0(m), b1 = a1 + o1(m), b2 = a2 + o2(m);
            const float v1 = V.v[lin(V, b0, b1, b2)];
            const float tt = (V.level - v0) / (v1 - v0);            // endpoints straddle the level: v1 != v0
            const float idx[3] = {a0 + tt * o0(m), a1 + tt * o1(m), a2 + tt * o2(m)};
            float pos[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) pos[F.ax[k]] = fmaf(F.sp[k], idx[k], F.org[k]);
            verts[vid * 3 + 0] = pos[0]; verts[vid * 3 + 1] = pos[1]; verts[vid * 3 + 2] = pos[2];
            if (normals) {
                float g[3], nn = 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    float g1 = ddx(V, b0, b1, b2, k);
                    float gk = fmaf(tt, g1 - g0[k], g0[k]) / F.sp[k];
                    g[F.ax[k]] = gk;
                    nn = fmaf(gk, gk, nn);
                }
                float inv = nn > 0.f ? rsqrtf(nn) : 0.f;
                normals[vid * 3 + 0] = g[0] * inv; normals[vid * 3 + 1] = g[1] * inv; normals[vid * 3 + 2] = g[2] * inv;
            }
            ++vid;
        }
    }
}

struct CellCtx {
    const Vol* V;
    const Frame* F;
    const uint8_t* flags;
    const int32_t* voff;
    int a0, a1, a2;
    float val[8];
};
// vertex id and logical position of the crossing on the cell edge between corners ca < cb (cb ^ ca = the edge's offset mask)
__device__ __forceinline__ int32_t edge_vertex(const CellCtx& c, int ca, int cb, float* pos) {
    const int m = cb ^ ca;
    const int w0 = c.a0 + o0(ca), w1 = c.a1 + o1(ca), w2 = c.a2 + o2(ca);
    const int64_t w = lin(*c.V, w0, w1, w2);
    const unsigned fl = c.flags[w];
    const int32_t id = c.voff[w] + __popc(fl & ((1u << (m - 1)) - 1u));
    const float tt = (c.V->level - c.val[ca]) / (c.val[cb] - c.val[ca]);
    const float idx[3] = {w0 + tt * o0(m), w1 + tt * o1(m), w2 + tt * o2(m)};
#pragma unroll
    for (int k = 0; k < 3; ++k) pos[c.F->ax[k]] = fmaf(c.F->sp[k], idx[k], c.F->org[k]);
    return id;
}
__device__ __forceinline__ void corner_pos(const CellCtx& c, int m, float* pos) {
    const float idx[3] = {(float)(c.a0 + o0(m)), (float)(c.a1 + o1(m)), (float)(c.a2 + o2(m))};
#pragma unroll
    for (int k = 0; k < 3; ++k) pos[c.F->ax[k]] = fmaf(c.F->sp[k], idx[k], c.F->org[k]);
}
// writes one triangle, flipped if its normal does not point along `dir` (from the inside corners to the outside ones)
__device__ __forceinline__ void put_tri(int32_t* faces, int64_t t, int32_t i0, int32_t i1, int32_t i2, const float* p0,
                                        const float* p1, const float* p2, const float* dir) {
    const float e1[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
    const float e2[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
    const float nx = e1[1] * e2[2] - e1[2] * e2[1], ny = e1[2] * e2[0] - e1[0] * e2[2], nz = e1[0] * e2[1] - e1[1] * e2[0];
    const bool flip = (nx * dir[0] + ny * dir[1] + nz * dir[2]) < 0.f;
    faces[t * 3 + 0] = i0;
    faces[t * 3 + 1] = flip ? i2 : i1;
    faces[t * 3 + 2] = flip ? i1 : i2;
}

__global__ void __launch_bounds__(256) k_mt_faces(Vol V, Frame F, const uint8_t* __restrict__ flags,
                                                   const uint8_t* __restrict__ ntri, const int32_t* __restrict__ voff,
                                                   const int32_t* __restrict__ toff, int32_t* __restrict__ faces) {
    const int64_t total = (int64_t)V.n0 * V.n1 * V.n2;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        if (!ntri[p]) continue;
        CellCtx c;
        c.V = &V; c.F = &F; c.flags = flags; c.voff = voff;
        c.a2 = (int)(p % V.n2);
        const int64_t t = p / V.n2;
        c.a1 = (int)(t % V.n1);
        c.a0 = (int)(t / V.n1);
        unsigned inside = 0;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            c.val[m] = V.v[lin(V, c.a0 + o0(m), c.a1 + o1(m), c.a2 + o2(m))];
            inside |= (c.val[m] < V.level ? 1u : 0u) << m;
        }
        int64_t tri = toff[p];
        for (int q = 0; q < 6; ++q) {
            const int cr[4] = {0, c_tet_a[q], c_tet_a[q] | c_tet_b[q], 7};   // nested masks: cr[i] is a subset of cr[j] for i < j
            int in_c[4], out_c[4], ni = 0, no = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if ((inside >> cr[i]) & 1u) in_c[ni++] = cr[i];
                else out_c[no++] = cr[i];
            }
            if (ni == 0 || ni == 4) continue;
            // direction from the inside corners to the outside corners (towards larger values)
            float dir[3] = {0.f, 0.f, 0.f}, cp[3];
            for (int i = 0; i < ni; ++i) { corner_pos(c, in_c[i], cp); for (int k = 0; k < 3; ++k) dir[k] -= cp[k] / ni; }
            for (int i = 0; i < no; ++i) { corner_pos(c, out_c[i], cp); for (int k = 0; k < 3; ++k) dir[k] += cp[k] / no; }
            auto ev = [&](int x, int y, float* pos) { return (x < y) ? edge_vertex(c, x, y, pos) : edge_vertex(c, y, x, pos); };
            float p0[3], p1[3], p2[3], p3[3];
            if (ni == 1 || no == 1) {
                const int apex = (ni == 1) ? in_c[0] : out_c[0];
                const int* base = (ni == 1) ? out_c : in_c;
                int32_t i0 = ev(apex, base[0], p0), i1 = ev(apex, base[1], p1), i2 = ev(apex, base[2], p2);
                put_tri(faces, tri++, i0, i1, i2, p0, p1, p2, dir);
            } else {
                int32_t i0 = ev(in_c[0], out_c[0], p0), i1 = ev(in_c[0], out_c[1], p1);
                int32_t i2 = ev(in_c[1], out_c[1], p2), i3 = ev(in_c[1], out_c[0], p3);
                put_tri(faces, tri++, i0, i1, i2, p0, p1, p2, dir);
                put_tri(faces, tri++, i0, i2, i3, p0, p2, p3, dir);
            }
        }
    }
}

static int check_vol(const float* vol, int n0, int n1, int n2) {
    if (!vol || n0 < 2 || n1 < 2 || n2 < 2) { set_error("marching tetrahedra: volume must be at least 2x2x2"); return DIGS_EINVAL; }
    if ((int64_t)n0 * n1 * n2 > ((int64_t)1 << 40)) { set_error("marching tetrahedra: volume too large"); return DIGS_EINVAL; }
    return DIGS_OK;
}
static unsigned blocks_for(int64_t total) {
    int64_t b = (total + 255) / 256;
    int64_t cap = (int64_t)tcpath::num_sms() * 16;
    return (unsigned)(b < cap ? b : cap);
}

int classify(const float* vol, int n0, int n1, int n2, float level, uint8_t* flags, uint8_t* ntri, cudaStream_t st) {
    int rc = check_vol(vol, n0, n1, n2);
    if (rc) return rc;
    Vol V{vol, n0, n1, n2, level};
    k_mt_classify<<<blocks_for((int64_t)n0 * n1 * n2), 256, 0, st>>>(V, flags, ntri);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("mt_classify launch: %s", cudaGetErrorString(e)); return DIGS_ECUDA; }
    return DIGS_OK;
}

int emit(const float* vol, int n0, int n1, int n2, float level, const int* axis_of, const float* origin, const float* spacing,
         const uint8_t* flags, const uint8_t* ntri, const int32_t* voff, const int32_t* toff, float* verts, float* normals,
         int32_t* faces, cudaStream_t st) {
    int rc = check_vol(vol, n0, n1, n2);
    if (rc) return rc;
    int seen = 0;
    for (int k = 0; k < 3; ++k) {
        if (axis_of[k] < 0 || axis_of[k] > 2) { set_error("marching tetrahedra: axis_of must be a permutation of 0,1,2"); return DIGS_EINVAL; }
        seen |= 1 << axis_of[k];
        if (!(spacing[k] > 0.f)) { set_error("marching tetrahedra: spacing must be positive"); return DIGS_EINVAL; }
    }
    if (seen != 7) { set_error("marching tetrahedra: axis_of must be a permutation of 0,1,2"); return DIGS_EINVAL; }
    Vol V{vol, n0, n1, n2, level};
    Frame F;
    for (int k = 0; k < 3; ++k) { F.ax[k] = axis_of[k]; F.org[k] = origin[k]; F.sp[k] = spacing[k]; }
    const unsigned nb = blocks_for((int64_t)n0 * n1 * n2);
    if (verts) {
        k_mt_vertices<<<nb, 256, 0, st>>>(V, F, flags, voff, verts, normals);
        count_launch();
    }
    if (faces) {
        k_mt_faces<<<nb, 256, 0, st>>>(V, F, flags, ntri, voff, toff, faces);
        count_launch();
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("mt_emit launch: %s", cudaGetErrorString(e)); return DIGS_ECUDA; }
    return DIGS_OK;
}

// ---- exact nearest neighbour by tiled brute force -------------------------------------------------------------------
// One thread owns QPT query points; the reference cloud streams through shared memory in tiles that every thread of the
// block reads as broadcasts.  Distances are (q - r)^2 sums, not |q|^2 + |r|^2 - 2 q.r: Chamfer distances are ~1e-3 of the
// coordinate scale and the expanded form loses them to cancellation.
constexpr int NN_TILE = 2048, NN_QPT = 2, NN_THREADS = 256;

template <int D>
__global__ void __launch_bounds__(NN_THREADS) k_nn(const float* __restrict__ ref, int64_t nr, const float* __restrict__ qry,
                                                   int64_t nq, float* __restrict__ dist, int32_t* __restrict__ idx) {
    __shared__ float4 tile[NN_TILE];
    float q[NN_QPT][3], best[NN_QPT];
    int32_t bi[NN_QPT];
    const int64_t q0 = ((int64_t)blockIdx.x * NN_THREADS + threadIdx.x) * NN_QPT;
#pragma unroll
    for (int j = 0; j < NN_QPT; ++j) {
        const int64_t qi = (q0 + j < nq) ? q0 + j : nq - 1;
        q[j][0] = qry[qi * D]; q[j][1] = qry[qi * D + 1]; q[j][2] = (D > 2) ? qry[qi * D + 2] : 0.f;
        best[j] = 3.4e38f; bi[j] = -1;
    }
    for (int64_t base = 0; base < nr; base += NN_TILE) {
        const int cnt = (int)((nr - base < NN_TILE) ? nr - base : NN_TILE);
        __syncthreads();
        for (int e = threadIdx.x; e < NN_TILE; e += NN_THREADS) {
            float4 r = make_float4(3e18f, 3e18f, 3e18f, 0.f);     // padding: farther than any real point
            if (e < cnt) {
                const float* rp = ref + (base + e) * D;
                r = make_float4(rp[0], rp[1], (D > 2) ? rp[2] : 0.f, 0.f);
            }
            tile[e] = r;
        }
        __syncthreads();
#pragma unroll 8
        for (int e = 0; e < NN_TILE; ++e) {
            const float4 r = tile[e];
#pragma unroll
            for (int j = 0; j < NN_QPT; ++j) {
                const float dx = q[j][0] - r.x, dy = q[j][1] - r.y, dz = q[j][2] - r.z;
                const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                if (d2 < best[j]) { best[j] = d2; bi[j] = (int32_t)(base + e); }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NN_QPT; ++j)
        if (q0 + j < nq) {
            dist[q0 + j] = sqrtf(best[j]);
            if (idx) idx[q0 + j] = bi[j];
        }
}

int nn_query(const float* ref, int64_t nr, const float* qry, int64_t nq, int d, float* dist, int32_t* idx, cudaStream_t st) {
    if (d != 2 && d != 3) { set_error("nn_query: d must be 2 or 3"); return DIGS_EINVAL; }
    if (nr <= 0) { set_error("nn_query: empty reference cloud"); return DIGS_EINVAL; }
    if (nr > 0x7fffffffLL) { set_error("nn_query: reference cloud exceeds int32 indices"); return DIGS_EINVAL; }
    if (nq == 0) return DIGS_OK;
    const int64_t per_block = (int64_t)NN_THREADS * NN_QPT;
    const unsigned nb = (unsigned)((nq + per_block - 1) / per_block);
    if (d == 3) k_nn<3><<<nb, NN_THREADS, 0, st>>>(ref, nr, qry, nq, dist, idx);
    else k_nn<2><<<nb, NN_THREADS, 0, st>>>(ref, nr, qry, nq, dist, idx);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("nn_query launch: %s", cudaGetErrorString(e)); return DIGS_ECUDA; }
    return DIGS_OK;
}

}  // namespace mesh
}  // namespace digs
