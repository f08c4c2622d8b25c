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
