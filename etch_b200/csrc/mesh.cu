// Mesh -> point cloud: bbox centring + area-weighted surface sampling, straight from vertex / face buffers on the device.
//
// SURVEY.md section 8(f) row 2 -- the step in front of the network in both callers:
//   src/inference_demo.py:19-34   preprocess_scan: centre = (min + max) / 2 over the vertices, vertices -= centre
//   src/inference_demo.py:36-39   trimesh.sample.sample_surface(mesh, num_points)
//   src/data_utils/GT_dataloader.py:100-102   the same with seed = self.seed + 15
// trimesh is a third-party dependency of the reference (environment.yml:21, unpinned, not vendored); its published algorithm
// (trimesh/sample.py::sample_surface, trimesh/triangles.py::area) is restated here in float64, operation for operation, so that
// with the SAME random draws the sampled points are bit-identical to the CPU restatement in oracle/mesh_sample.py:
//   crosses = cross(v1 - v0, v2 - v1); area = sqrt((crosses**2).sum()) / 2        (np.diff of the triangle, then np.cross)
//   cum = cumsum(area)  (sequential, as numpy);  pick = u_face * cum[-1];  face = searchsorted(cum, pick)   (side = 'left')
//   (r1, r2) = u_len; if r1 + r2 > 1: r1 -= 1, r2 -= 1; r = |r|;  p = ((v1 - v0) * r1 + (v2 - v0) * r2) + v0
// The host draws the uniforms with numpy's generator (what trimesh does), so seeded calls reproduce the reference's clouds.
// No product of two doubles is fused with an add here (numpy never contracts): explicit __dmul_rn / __dadd_rn / __dsub_rn.
#include "common.cuh"

namespace {

constexpr int MT = 1024;

// ---- bbox centre: block-wide min / max of the vertices (any number of vertices, one block), then v -= centre ----
__global__ void __launch_bounds__(MT) mesh_center_kernel(const double* __restrict__ verts, int V, double* __restrict__ centre,
                                                         double* __restrict__ centred) {
    __shared__ double s_min[3][MT / 32], s_max[3][MT / 32];
    __shared__ double s_c[3];
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = threadIdx.x; i < V; i += MT)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double v = verts[(size_t)i * 3 + c];
            mn[c] = fmin(mn[c], v); mx[c] = fmax(mx[c], v);
        }
#pragma unroll
    for (int c = 0; c < 3; ++c)
        for (int o = 16; o > 0; o >>= 1) {
            mn[c] = fmin(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
            mx[c] = fmax(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
        }
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int c = 0; c < 3; ++c) { s_min[c][threadIdx.x >> 5] = mn[c]; s_max[c][threadIdx.x >> 5] = mx[c]; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double a = s_min[threadIdx.x][0], b = s_max[threadIdx.x][0];
        for (int w = 1; w < MT / 32; ++w) { a = fmin(a, s_min[threadIdx.x][w]); b = fmax(b, s_max[threadIdx.x][w]); }
        const double c = __ddiv_rn(__dadd_rn(a, b), 2.0);     // (min + max) / 2.0
        s_c[threadIdx.x] = c;
        centre[threadIdx.x] = c;
    }
    __syncthreads();
    if (centred)
        for (int i = threadIdx.x; i < V * 3; i += MT) centred[i] = __dsub_rn(verts[i], s_c[i % 3]);
}

// ---- face areas (trimesh.triangles.area over np.diff'ed triangles) ----
__global__ void mesh_area_kernel(const double* __restrict__ verts, const int* __restrict__ faces, int F, double* __restrict__ area) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int i0 = faces[(size_t)f * 3], i1 = faces[(size_t)f * 3 + 1], i2 = faces[(size_t)f * 3 + 2];
    double a[3], b[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double v0 = verts[(size_t)i0 * 3 + c], v1 = verts[(size_t)i1 * 3 + c], v2 = verts[(size_t)i2 * 3 + c];
        a[c] = __dsub_rn(v1, v0);     // np.diff(triangle)[0]
        b[c] = __dsub_rn(v2, v1);     // np.diff(triangle)[1]
    }
    const double cx = __dsub_rn(__dmul_rn(a[1], b[2]), __dmul_rn(a[2], b[1]));
    const double cy = __dsub_rn(__dmul_rn(a[2], b[0]), __dmul_rn(a[0], b[2]));
    const double cz = __dsub_rn(__dmul_rn(a[0], b[1]), __dmul_rn(a[1], b[0]));
    const double ss = __dadd_rn(__dadd_rn(__dmul_rn(cx, cx), __dmul_rn(cy, cy)), __dmul_rn(cz, cz));
    area[f] = __ddiv_rn(__dsqrt_rn(ss), 2.0);
}

// ---- cumulative sum in numpy's (sequential) order: the block stages 1024 areas at a time in shared memory, thread 0 runs the
// dependent DADD chain (the only part that cannot be parallelised without changing the roundings), the block writes back ----
__global__ void __launch_bounds__(MT) mesh_cumsum_kernel(const double* __restrict__ area, int F, double* __restrict__ cum) {
    __shared__ double s[MT];
    __shared__ double carry;
    if (threadIdx.x == 0) carry = 0.0;
    for (int base = 0; base < F; base += MT) {
        const int i = base + threadIdx.x;
        __syncthreads();
        s[threadIdx.x] = i < F ? area[i] : 0.0;
        __syncthreads();
        if (threadIdx.x == 0) {
            double acc = carry;
            const int n = min(MT, F - base);
            for (int k = 0; k < n; ++k) { acc = __dadd_rn(acc, s[k]); s[k] = acc; }
            carry = acc;
        }
        __syncthreads();
        if (i < F) cum[i] = s[threadIdx.x];
    }
}

// ---- the samples ----
__global__ void mesh_sample_kernel(const double* __restrict__ verts, const int* __restrict__ faces, int F,
                                   const double* __restrict__ cum, const double* __restrict__ u_face,
                                   const double* __restrict__ u_len, int count, double* __restrict__ out64,
                                   float* __restrict__ out32, int* __restrict__ face_index) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double pick = __dmul_rn(u_face[i], cum[F - 1]);
    int lo = 0, hi = F;                                   // np.searchsorted(cum, pick, side='left'): first index with cum >= pick
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cum[mid] < pick) lo = mid + 1; else hi = mid;
    }
    const int f = lo < F ? lo : F - 1;                    // pick == cum[-1] * (1 - eps) at most: lo < F always; clamp defensively
    const int i0 = faces[(size_t)f * 3], i1 = faces[(size_t)f * 3 + 1], i2 = faces[(size_t)f * 3 + 2];
    double r1 = u_len[(size_t)i * 2], r2 = u_len[(size_t)i * 2 + 1];
    if (__dadd_rn(r1, r2) > 1.0) { r1 = __dsub_rn(r1, 1.0); r2 = __dsub_rn(r2, 1.0); }
    r1 = fabs(r1); r2 = fabs(r2);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double v0 = verts[(size_t)i0 * 3 + c];
        const double e1 = __dsub_rn(verts[(size_t)i1 * 3 + c], v0), e2 = __dsub_rn(verts[(size_t)i2 * 3 + c], v0);
        const double p = __dadd_rn(__dadd_rn(__dmul_rn(e1, r1), __dmul_rn(e2, r2)), v0);
        if (out64) out64[(size_t)i * 3 + c] = p;
        if (out32) out32[(size_t)i * 3 + c] = (float)p;   // torch.from_numpy(points).float(): round to nearest
    }
    if (face_index) face_index[i] = f;
}

}  // namespace

// preprocess_scan (src/inference_demo.py:19-34): centre[3] = (min + max) / 2 over verts [V,3] (float64); centred = verts - centre
// (may be NULL, may alias verts).
ETCH_API int etch_mesh_center(const double* verts, int V, double* centre, double* centred, cudaStream_t stream) {
    if (!verts || !centre || V <= 0) return ETCH_EINVAL;
    mesh_center_kernel<<<1, MT, 0, stream>>>(verts, V, centre, centred);
    ETCH_RETURN_LAST();
}

// trimesh.sample.sample_surface (src/inference_demo.py:36-39, src/data_utils/GT_dataloader.py:102) on device buffers.
// verts [V,3] f64, faces [F,3] i32, u_face [count] and u_len [count,2] = the uniform draws in trimesh's order (first `count`, then
// `2*count`), scratch [2*F] f64 (areas | cumulative areas).  Outputs (each may be NULL): out64 [count,3], out32 [count,3], face_index.
ETCH_API int etch_mesh_sample(const double* verts, const int* faces, int V, int F, const double* u_face, const double* u_len,
                              int count, double* scratch, double* out64, float* out32, int* face_index, cudaStream_t stream) {
    if (!verts || !faces || !u_face || !u_len || !scratch || V <= 0 || F <= 0 || count <= 0) return ETCH_EINVAL;
    double* area = scratch;
    double* cum = scratch + F;
    mesh_area_kernel<<<(unsigned)etch_cdiv(F, 256), 256, 0, stream>>>(verts, faces, F, area);
    mesh_cumsum_kernel<<<1, MT, 0, stream>>>(area, F, cum);
    mesh_sample_kernel<<<(unsigned)etch_cdiv(count, 256), 256, 0, stream>>>(verts, faces, F, cum, u_face, u_len, count, out64, out32, face_index);
    ETCH_RETURN_LAST();
}

// ================================================================================================================
// Closest point on a triangle mesh + nearest info point: the "VECTORS" block of the evaluation dataset
// (src/data_utils/GT_dataloader.py:104-124):
//     dists, indices = cKDTree(info_points).query(sample_points, k=1)
//     closest_points, _, _ = trimesh.proximity.closest_point(smpl_mesh, sample_points)
//     vectors = info_vectors[indices] if dists < 0.01 else sample_points - closest_points
// Brute force in float64 over all triangles / info points (a B200 does the 5000 x 13776 point-triangle tests of one scan in
// well under a millisecond; the reference walks an r-tree and a KD-tree on the CPU).  Point-triangle projection: the Voronoi-
// region algorithm of Ericson, "Real-Time Collision Detection" 5.1.5 (what trimesh.triangles.closest_point vectorises), with
// every dot product evaluated as ((x + y) + z) without fusion so that oracle/mesh_sample.py reproduces it bit for bit.
namespace {

struct D3 { double x, y, z; };
__device__ __forceinline__ D3 sub3(D3 a, D3 b) { return {__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y), __dsub_rn(a.z, b.z)}; }
__device__ __forceinline__ double dot3(D3 a, D3 b) { return __dadd_rn(__dadd_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)), __dmul_rn(a.z, b.z)); }
__device__ __forceinline__ D3 madd3(D3 a, D3 d, double t) { return {__dadd_rn(a.x, __dmul_rn(d.x, t)), __dadd_rn(a.y, __dmul_rn(d.y, t)), __dadd_rn(a.z, __dmul_rn(d.z, t))}; }

__device__ D3 closest_on_triangle(D3 p, D3 a, D3 b, D3 c) {
    const D3 ab = sub3(b, a), ac = sub3(c, a), ap = sub3(p, a);
    const double d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    if (d1 <= 0.0 && d2 <= 0.0) return a;
    const D3 bp = sub3(p, b);
    const double d3 = dot3(ab, bp), d4 = dot3(ac, bp);
    if (d3 >= 0.0 && d4 <= d3) return b;
    const double vc = __dsub_rn(__dmul_rn(d1, d4), __dmul_rn(d3, d2));
    if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) return madd3(a, ab, __ddiv_rn(d1, __dsub_rn(d1, d3)));
    const D3 cp = sub3(p, c);
    const double d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= 0.0 && d5 <= d6) return c;
    const double vb = __dsub_rn(__dmul_rn(d5, d2), __dmul_rn(d1, d6));
    if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) return madd3(a, ac, __ddiv_rn(d2, __dsub_rn(d2, d6)));
    const double va = __dsub_rn(__dmul_rn(d3, d6), __dmul_rn(d5, d4));
    const double e43 = __dsub_rn(d4, d3), e56 = __dsub_rn(d5, d6);
    if (va <= 0.0 && e43 >= 0.0 && e56 >= 0.0) return madd3(b, sub3(c, b), __ddiv_rn(e43, __dadd_rn(e43, e56)));
    const double denom = __ddiv_rn(1.0, __dadd_rn(__dadd_rn(va, vb), vc));
    const double v = __dmul_rn(vb, denom), w = __dmul_rn(vc, denom);
    const D3 q = madd3(a, ab, v);
    return madd3(q, ac, w);
}

constexpr int CP_T = 128;      // query points per block (one per thread)
constexpr int CP_TILE = 128;   // triangles staged per shared-memory tile

// grid (ceil(n / CP_T), chunks): every block scans triangles [chunk*len, (chunk+1)*len) for its 128 points
__global__ void __launch_bounds__(CP_T) closest_partial_kernel(const double* __restrict__ verts, const int* __restrict__ faces, int F,
                                                               const double* __restrict__ pts, int n, int chunk_len,
                                                               double* __restrict__ part_d, int* __restrict__ part_f) {
    __shared__ double tri[CP_TILE][9];
    const int i = blockIdx.x * CP_T + threadIdx.x;
    D3 p = {0, 0, 0};
    if (i < n) p = {pts[(size_t)i * 3], pts[(size_t)i * 3 + 1], pts[(size_t)i * 3 + 2]};
    const int f0 = blockIdx.y * chunk_len, f1 = min(F, f0 + chunk_len);
    double best = INFINITY;
    int bf = -1;
    for (int base = f0; base < f1; base += CP_TILE) {
        const int cnt = min(CP_TILE, f1 - base);
        __syncthreads();
        for (int e = threadIdx.x; e < cnt * 9; e += CP_T) {
            const int t = e / 9, r = e % 9;
            tri[t][r] = verts[(size_t)faces[(size_t)(base + t) * 3 + r / 3] * 3 + r % 3];
        }
        __syncthreads();
        if (i < n)
            for (int t = 0; t < cnt; ++t) {
                const D3 a = {tri[t][0], tri[t][1], tri[t][2]}, b = {tri[t][3], tri[t][4], tri[t][5]}, c = {tri[t][6], tri[t][7], tri[t][8]};
                const D3 d = sub3(p, closest_on_triangle(p, a, b, c));
                const double dd = dot3(d, d);
                if (dd < best) { best = dd; bf = base + t; }      // strict: the lowest face index wins a tie
            }
    }
    if (i < n) { part_d[(size_t)blockIdx.y * n + i] = best; part_f[(size_t)blockIdx.y * n + i] = bf; }
}

__global__ void closest_finish_kernel(const double* __restrict__ verts, const int* __restrict__ faces, const double* __restrict__ pts,
                                      int n, int chunks, const double* __restrict__ part_d, const int* __restrict__ part_f,
                                      double* __restrict__ closest, double* __restrict__ dist, int* __restrict__ face) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double best = INFINITY;
    int bf = -1;
    for (int c = 0; c < chunks; ++c) {
        const double d = part_d[(size_t)c * n + i];
        if (d < best) { best = d; bf = part_f[(size_t)c * n + i]; }
    }
    const D3 p = {pts[(size_t)i * 3], pts[(size_t)i * 3 + 1], pts[(size_t)i * 3 + 2]};
    const int i0 = faces[(size_t)bf * 3], i1 = faces[(size_t)bf * 3 + 1], i2 = faces[(size_t)bf * 3 + 2];
    const D3 a = {verts[(size_t)i0 * 3], verts[(size_t)i0 * 3 + 1], verts[(size_t)i0 * 3 + 2]};
    const D3 b = {verts[(size_t)i1 * 3], verts[(size_t)i1 * 3 + 1], verts[(size_t)i1 * 3 + 2]};
    const D3 c = {verts[(size_t)i2 * 3], verts[(size_t)i2 * 3 + 1], verts[(size_t)i2 * 3 + 2]};
    const D3 q = closest_on_triangle(p, a, b, c);
    closest[(size_t)i * 3] = q.x; closest[(size_t)i * 3 + 1] = q.y; closest[(size_t)i * 3 + 2] = q.z;
    if (dist) dist[i] = __dsqrt_rn(best);
    if (face) face[i] = bf;
}

// nearest of m reference points (float64), brute force through shared memory; ties -> lowest index
__global__ void __launch_bounds__(CP_T) nearest_point_kernel(const double* __restrict__ ref, int m, const double* __restrict__ pts, int n,
                                                             double* __restrict__ dist, int* __restrict__ index) {
    __shared__ double tile[256][3];
    const int i = blockIdx.x * CP_T + threadIdx.x;
    D3 p = {0, 0, 0};
    if (i < n) p = {pts[(size_t)i * 3], pts[(size_t)i * 3 + 1], pts[(size_t)i * 3 + 2]};
    double best = INFINITY;
    int bi = -1;
    for (int base = 0; base < m; base += 256) {
        const int cnt = min(256, m - base);
        __syncthreads();
        for (int e = threadIdx.x; e < cnt * 3; e += CP_T) tile[e / 3][e % 3] = ref[(size_t)base * 3 + e];
        __syncthreads();
        if (i < n)
            for (int t = 0; t < cnt; ++t) {
                const D3 d = sub3(p, D3{tile[t][0], tile[t][1], tile[t][2]});
                const double dd = dot3(d, d);
                if (dd < best) { best = dd; bi = base + t; }
            }
    }
    if (i < n) { dist[i] = __dsqrt_rn(best); index[i] = bi; }
}

// vectors = info_vectors[idx] where the nearest info point is closer than `threshold`, else sample - closest (GT_dataloader.py:112-124)
__global__ void gt_vectors_kernel(const double* __restrict__ pts, const double* __restrict__ closest, const double* __restrict__ info_vectors,
                                  const double* __restrict__ nn_dist, const int* __restrict__ nn_idx, int n, double threshold,
                                  double* __restrict__ vectors) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool near = nn_dist[i] < threshold;
#pragma unroll
    for (int c = 0; c < 3; ++c)
        vectors[(size_t)i * 3 + c] = near ? info_vectors[(size_t)nn_idx[i] * 3 + c] : __dsub_rn(pts[(size_t)i * 3 + c], closest[(size_t)i * 3 + c]);
}

}  // namespace

// trimesh.proximity.closest_point(mesh, points) (GT_dataloader.py:110): closest [n,3], dist [n] (optional), face [n] (optional).
// scratch: chunks * n doubles followed by chunks * n ints, chunks = etch_mesh_closest_chunks(F).
ETCH_API int etch_mesh_closest_chunks(int F) { int c = (F + 2047) / 2048; return c < 1 ? 1 : (c > 64 ? 64 : c); }

ETCH_API int etch_mesh_closest_point(const double* verts, const int* faces, int V, int F, const double* pts, int n, void* scratch,
                                     double* closest, double* dist, int* face, cudaStream_t stream) {
    if (!verts || !faces || !pts || !scratch || !closest || V <= 0 || F <= 0 || n <= 0) return ETCH_EINVAL;
    const int chunks = etch_mesh_closest_chunks(F);
    const int len = etch_cdiv(F, chunks);
    double* pd = reinterpret_cast<double*>(scratch);
    int* pf = reinterpret_cast<int*>(pd + (size_t)chunks * n);
    dim3 grid((unsigned)etch_cdiv(n, CP_T), (unsigned)chunks);
    closest_partial_kernel<<<grid, CP_T, 0, stream>>>(verts, faces, F, pts, n, len, pd, pf);
    closest_finish_kernel<<<(unsigned)etch_cdiv(n, 128), 128, 0, stream>>>(verts, faces, pts, n, chunks, pd, pf, closest, dist, face);
    ETCH_RETURN_LAST();
}

// cKDTree(ref).query(pts, k=1) (GT_dataloader.py:106-107): dist [n] float64, index [n] int32.
ETCH_API int etch_nearest_point(const double* ref, int m, const double* pts, int n, double* dist, int* index, cudaStream_t stream) {
    if (!ref || !pts || !dist || !index || m <= 0 || n <= 0) return ETCH_EINVAL;
    nearest_point_kernel<<<(unsigned)etch_cdiv(n, CP_T), CP_T, 0, stream>>>(ref, m, pts, n, dist, index);
    ETCH_RETURN_LAST();
}

// the vector assembly of GT_dataloader.py:112-124
ETCH_API int etch_gt_vectors(const double* pts, const double* closest, const double* info_vectors, const double* nn_dist, const int* nn_idx,
                             int n, double threshold, double* vectors, cudaStream_t stream) {
    if (!pts || !closest || !info_vectors || !nn_dist || !nn_idx || !vectors || n <= 0) return ETCH_EINVAL;
    gt_vectors_kernel<<<(unsigned)etch_cdiv(n, 256), 256, 0, stream>>>(pts, closest, info_vectors, nn_dist, nn_idx, n, threshold, vectors);
    ETCH_RETURN_LAST();
}
