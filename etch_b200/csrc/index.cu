// Index kernels of the ETCH hot path: furthest point sampling (vgtk + pointops variants), ordered ball query,
// channel-major gather, brute-force kNN.  All index-exact against the reference kernels, including tie rules:
//   external/vgtk/vgtk/cuda/grouping_cuda_kernel.cu:67-113   ball_query_cuda_kernel
//   external/vgtk/vgtk/cuda/grouping_cuda_kernel.cu:339-466  furthest_point_sampling_cuda_kernel (+ __update)
//   external/vgtk/vgtk/cuda/gathering_cuda_kernel.cu:42-68   gather_points_forward_kernel
//   external/pointops/src/knnquery/knnquery_cuda_kernel.cu:21-108
//   external/pointops/src/sampling/sampling_cuda_kernel.cu:5-129
//
// B200 design notes
//  * FPS is a chain of m-1 dependent arg-max steps, i.e. latency bound.  One CTA per scan keeps every point and its
//    running min-distance in REGISTERS (x,y staged in shared memory for clouds > 10k points), reduces with
//    redux.sync (2 REDUX per level) and needs ONE __syncthreads per step (double-buffered warp slots).  The
//    reference needs 11 barriers + 2 global round trips of temp[] per step.
//  * The reference arg-max winner among equal distances is decided by its strided scan (lowest k inside a thread)
//    and its shared-memory tree (slot 0 vs slot s: lower slot wins on ties => the winner is the thread whose
//    BIT-REVERSED id is smallest).  We reproduce that with a packed tie key instead of replaying the tree.
//  * Ball query: one warp per query, ballot + popc ordered compaction, early exit.
//  * kNN: one thread per query running the reference's max-heap verbatim (so equal-distance order is identical),
//    supports staged through shared memory tiles.
#include "common.cuh"
#include <cooperative_groups.h>

namespace {

// ------------------------------------------------------------------------------------------------ FPS
// tie key: larger wins.  bits[30:21] = 1023 - bitrev_log2T(k mod T), bits[20:0] = 0x1FFFFF - k  (k < 2^21 - 1)
__device__ __forceinline__ unsigned fps_tie_key(int k, int Tm1, int shift) {
    const unsigned br = shift >= 32 ? 0u : (__brev((unsigned)(k & Tm1)) >> shift);
    return ((1023u - br) << 21) | (0x1FFFFFu - (unsigned)k);
}

struct FpsSeg {
    const float* xyz;  // BCN: channel-major base of this scan; PACKED: row-major base of the whole packed array
    int n;             // points in this segment
    int m;             // samples to draw
    int base;          // PACKED: first row of the segment (indices are emitted as base + k)
    int* out;          // where idx[0..m) of this segment go
};

template <bool PACKED>
__device__ __forceinline__ void fps_load(const FpsSeg& s, int k, float& x, float& y, float& z) {
    if (PACKED) {
        const float* p = s.xyz + (size_t)(s.base + k) * 3;
        x = __ldg(p); y = __ldg(p + 1); z = __ldg(p + 2);
    } else {
        x = __ldg(s.xyz + k); y = __ldg(s.xyz + s.n + k); z = __ldg(s.xyz + 2 * (size_t)s.n + k);
    }
}

// PPT points per thread.  SMEM_XY: x,y of every point live in shared memory, z and the running distance in
// registers (for clouds that do not fit the 64-register budget of a 1024-thread CTA).
template <int PPT, bool PACKED, bool SMEM_XY>
__device__ __forceinline__ void fps_segment(const FpsSeg& s, int T_ref, float* sx, float* sy, uint2 (*slots)[32]) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NT = blockDim.x, NW = NT >> 5;
    const int Tm1 = T_ref - 1;
    int lg = 0;
    while ((1 << lg) < T_ref) ++lg;
    const int shift = 32 - lg;

    float px[SMEM_XY ? 1 : PPT], py[SMEM_XY ? 1 : PPT], pz[PPT], tmp[PPT];
    unsigned valid = 0;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        const int k = tid + i * NT;
        float x = 0.f, y = 0.f, z = 0.f;
        bool ok = k < s.n;
        if (ok) {
            fps_load<PACKED>(s, k, x, y, z);
            if (!PACKED) {  // vgtk variant: points with |p|^2 <= 1e-3 never update temp and never win (:385-387)
                const float mag = etch_sqdist3(x, y, z);   // grouping_cuda_kernel.cu:384, same contraction
                if ((double)mag <= 1e-3) ok = false;
            }
        }
        if constexpr (SMEM_XY) {
            if (k < s.n) { sx[k] = x; sy[k] = y; }
        } else {
            px[i] = x; py[i] = y;
            if (k < s.n) { sx[k] = x; sy[k] = y; sy[s.n + (k)] = z; }   // winner look-up table (x | y | z), 12*n bytes
        }
        pz[i] = z;
        tmp[i] = 1e10f;
        if (ok) valid |= 1u << i;
    }
    if (tid == 0 && s.m > 0) s.out[0] = s.base;
    for (int i = tid; i < 64; i += NT) slots[i >> 5][i & 31] = make_uint2(0u, 0u);
    __syncthreads();

    int old = 0;
    for (int j = 1; j < s.m; ++j) {
        float x1, y1, z1;
        if constexpr (SMEM_XY) fps_load<PACKED>(s, old, x1, y1, z1);        // large clouds: z is not staged, read it back from L1/L2
        else { x1 = sx[old]; y1 = sy[old]; z1 = sy[s.n + old]; }             // broadcast shared-memory reads (~30 cycles)
        unsigned bv = 0u, bt = 0u;
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            if (valid & (1u << i)) {
                const int k = tid + i * NT;
                float x2, y2;
                if constexpr (SMEM_XY) { x2 = sx[k]; y2 = sy[k]; } else { x2 = px[i]; y2 = py[i]; }
                const float d = etch_sqdist3(x2 - x1, y2 - y1, pz[i] - z1);
                const float d2 = fminf(d, tmp[i]);
                tmp[i] = d2;
                const unsigned vb = __float_as_uint(d2);  // d2 >= 0: bit pattern is monotone
                const unsigned tk = fps_tie_key(k, Tm1, shift);
                if (vb > bv || (vb == bv && tk > bt)) { bv = vb; bt = tk; }
            }
        }
        const unsigned wv = __reduce_max_sync(0xffffffffu, bv);
        const unsigned wt = __reduce_max_sync(0xffffffffu, bv == wv ? bt : 0u);
        const int buf = j & 1;
        if (lane == 0) slots[buf][warp] = make_uint2(wv, wt);
        __syncthreads();
        const uint2 sl = lane < NW ? slots[buf][lane] : make_uint2(0u, 0u);
        const unsigned gv = __reduce_max_sync(0xffffffffu, sl.x);
        const unsigned gt = __reduce_max_sync(0xffffffffu, sl.x == gv ? sl.y : 0u);
        old = gt ? (int)(0x1FFFFFu - (gt & 0x1FFFFFu)) : 0;  // no candidate at all: reference keeps besti init
        if (tid == 0) s.out[j] = s.base + old;
    }
}

template <int PPT, bool SMEM_XY>
__global__ void __launch_bounds__(1024, 1) fps_bcn_kernel(const float* __restrict__ xyz, int n, int m, int T_ref,
                                                          int* __restrict__ idx) {
    extern __shared__ float fps_smem[];
    __shared__ uint2 slots[2][32];
    FpsSeg s;
    s.xyz = xyz + (size_t)blockIdx.x * 3 * n;
    s.n = n; s.m = m; s.base = 0;
    s.out = idx + (size_t)blockIdx.x * m;
    fps_segment<PPT, false, SMEM_XY>(s, T_ref, fps_smem, fps_smem + n, slots);
}

template <int PPT, bool SMEM_XY>
__global__ void __launch_bounds__(1024, 1) fps_packed_kernel(const float* __restrict__ xyz, const int* __restrict__ offset,
                                                             const int* __restrict__ new_offset, int T_ref,
                                                             int* __restrict__ idx) {
    extern __shared__ float fps_smem[];
    __shared__ uint2 slots[2][32];
    const int b = blockIdx.x;
    const int start_n = b == 0 ? 0 : offset[b - 1], end_n = offset[b];
    const int start_m = b == 0 ? 0 : new_offset[b - 1], end_m = new_offset[b];
    FpsSeg s;
    s.xyz = xyz; s.n = end_n - start_n; s.m = end_m - start_m; s.base = start_n;
    s.out = idx + start_m;
    fps_segment<PPT, true, SMEM_XY>(s, T_ref, fps_smem, fps_smem + s.n, slots);
}

// ------------------------------------------------------------------------------------------------ cluster FPS (large clouds)
// FPS is a chain of m-1 dependent steps, each a distance update of every point + an argmax.  One CTA keeps up to ~8k points
// in registers; beyond that (BASELINE configs[2..4]: 10k / 20k-point scans) the register file of one SM is too small and the
// per-step work (n x ~20 instructions) too long for one SM.  Here a thread-block CLUSTER of CL = 4 or 8 CTAs owns one scan: each
// CTA holds n / CL points in registers, finds its local winner exactly as fps_segment does, posts (value, tie key, xyz)
// into every peer's shared memory (distributed shared memory) and one cluster barrier later every CTA picks the same global
// winner from the CL candidates -- no global-memory round trip on the critical path.  Same (value, tie key) order as the
// single-CTA kernel, hence the same indices as the reference.
constexpr int FPS_CL_MAX = 8;               // portable cluster size limit
constexpr int FPS_CLUSTER_MIN = 8192;      // clouds above this size use the cluster kernel
struct __align__(16) FpsCand { unsigned v, t; float x, y, z; unsigned pad[3]; };

template <int PPT, bool PACKED, int FPS_CL>
__global__ void __launch_bounds__(1024, 1)
fps_cluster_kernel(const float* __restrict__ xyz, const int* __restrict__ offset, const int* __restrict__ new_offset, int n_bcn,
                   int m_bcn, int T_ref, int* __restrict__ idx) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ float fps_smem[];            // [3][per] coordinates of this CTA's points (winner look-up)
    __shared__ uint2 slots[2][32];
    __shared__ FpsCand cand[2][FPS_CL];
    const int rank = (int)cluster.block_rank();
    const int scan = blockIdx.x / FPS_CL;
    FpsSeg s;
    if (PACKED) {
        const int start_n = scan == 0 ? 0 : offset[scan - 1], end_n = offset[scan];
        const int start_m = scan == 0 ? 0 : new_offset[scan - 1], end_m = new_offset[scan];
        s.xyz = xyz; s.n = end_n - start_n; s.m = end_m - start_m; s.base = start_n; s.out = idx + start_m;
    } else {
        s.xyz = xyz + (size_t)scan * 3 * n_bcn; s.n = n_bcn; s.m = m_bcn; s.base = 0; s.out = idx + (size_t)scan * m_bcn;
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (s.n + FPS_CL - 1) / FPS_CL, kbase = rank * per;
    float* sx = fps_smem; float* sy = sx + per; float* sz = sy + per;
    const int Tm1 = T_ref - 1;
    int lg = 0;
    while ((1 << lg) < T_ref) ++lg;
    const int shift = 32 - lg;

    float px[PPT], py[PPT], pz[PPT], tmp[PPT];
    unsigned valid = 0;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        const int jl = tid + i * 1024, k = kbase + jl;
        float x = 0.f, y = 0.f, z = 0.f;
        bool ok = jl < per && k < s.n;
        if (ok) {
            fps_load<PACKED>(s, k, x, y, z);
            sx[jl] = x; sy[jl] = y; sz[jl] = z;
            if (!PACKED) {
                const float mag = etch_sqdist3(x, y, z);   // grouping_cuda_kernel.cu:384, same contraction
                if ((double)mag <= 1e-3) ok = false;
            }
        }
        px[i] = x; py[i] = y; pz[i] = z; tmp[i] = 1e10f;
        if (ok) valid |= 1u << i;
    }
    if (rank == 0 && tid == 0 && s.m > 0) s.out[0] = s.base;
    for (int i = tid; i < 64; i += 1024) slots[i >> 5][i & 31] = make_uint2(0u, 0u);
    float x1, y1, z1;
    fps_load<PACKED>(s, 0, x1, y1, z1);
    __syncthreads();
    cluster.sync();                                // every CTA's shared memory is live before the first remote store

    for (int j = 1; j < s.m; ++j) {
        unsigned bv = 0u, bt = 0u;
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            if (valid & (1u << i)) {
                const int k = kbase + tid + i * 1024;
                const float d = etch_sqdist3(px[i] - x1, py[i] - y1, pz[i] - z1);
                const float d2 = fminf(d, tmp[i]);
                tmp[i] = d2;
                const unsigned vb = __float_as_uint(d2);
                const unsigned tk = fps_tie_key(k, Tm1, shift);
                if (vb > bv || (vb == bv && tk > bt)) { bv = vb; bt = tk; }
            }
        }
        const unsigned wv = __reduce_max_sync(0xffffffffu, bv);
        const unsigned wt = __reduce_max_sync(0xffffffffu, bv == wv ? bt : 0u);
        const int buf = j & 1;
        if (lane == 0) slots[buf][warp] = make_uint2(wv, wt);
        __syncthreads();
        if (warp == 0) {
            const uint2 sl = slots[buf][lane];
            const unsigned gv = __reduce_max_sync(0xffffffffu, sl.x);
            const unsigned gt = __reduce_max_sync(0xffffffffu, sl.x == gv ? sl.y : 0u);
            if (lane < FPS_CL) {                   // post this CTA's candidate into peer `lane`
                FpsCand c;
                c.v = gv; c.t = gt; c.x = 0.f; c.y = 0.f; c.z = 0.f; c.pad[0] = c.pad[1] = c.pad[2] = 0u;
                if (gt) {
                    const int kl = (int)(0x1FFFFFu - (gt & 0x1FFFFFu)) - kbase;
                    c.x = sx[kl]; c.y = sy[kl]; c.z = sz[kl];
                }
                FpsCand* dst = cluster.map_shared_rank(&cand[buf][rank], lane);
                *reinterpret_cast<uint4*>(dst) = make_uint4(c.v, c.t, __float_as_uint(c.x), __float_as_uint(c.y));
                reinterpret_cast<float*>(dst)[4] = c.z;
            }
        }
        cluster.sync();
        unsigned gv = 0u, gt = 0u;
        int win = 0;
#pragma unroll
        for (int r = 0; r < FPS_CL; ++r) {
            const unsigned v = cand[buf][r].v, t = cand[buf][r].t;
            if (v > gv || (v == gv && t > gt)) { gv = v; gt = t; win = r; }
        }
        int old = 0;
        if (gt) {
            old = (int)(0x1FFFFFu - (gt & 0x1FFFFFu));
            x1 = cand[buf][win].x; y1 = cand[buf][win].y; z1 = cand[buf][win].z;
        } else {
            fps_load<PACKED>(s, 0, x1, y1, z1);    // no candidate at all: the reference keeps besti = 0
        }
        if (rank == 0 && tid == 0) s.out[j] = s.base + old;
    }
    cluster.sync();                                // nobody exits while a peer may still write into its shared memory
}


// launch a cluster kernel (cluster size is a launch attribute so that one template serves 4- and 8-CTA clusters)
template <typename Kern, typename... Args>
static int fps_cluster_launch(Kern kern, int clusters, int cl, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * cl));
    cfg.blockDim = dim3(1024);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (smem > 48 * 1024) ETCH_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ETCH_TRY(cudaLaunchKernelEx(&cfg, kern, args...));
    return ETCH_OK;
}

// ------------------------------------------------------------------------------------------------ ball query
// one warp per query; idx [B,m,nsample] fully written (zero-fill semantics of grouping_cuda.cpp:80-82 included)
__global__ void __launch_bounds__(256) ball_query_kernel(const float* __restrict__ new_xyz, const float* __restrict__ xyz,
                                                         int n, int m, float radius2, int nsample, int* __restrict__ idx) {
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= m) return;
    const float* X = xyz + (size_t)b * 3 * n;
    const float* Q = new_xyz + (size_t)b * 3 * m;
    int* out = idx + ((size_t)b * m + j) * nsample;
    const float qx = __ldg(Q + j), qy = __ldg(Q + m + j), qz = __ldg(Q + 2 * (size_t)m + j);
    int cnt = 0;
    for (int k0 = 0; k0 < n && cnt < nsample; k0 += 32) {
        const int k = k0 + lane;
        bool hit = false;
        if (k < n) {
            const float d2 = etch_sqdist3(qx - __ldg(X + k), qy - __ldg(X + n + k), qz - __ldg(X + 2 * (size_t)n + k));
            hit = d2 < radius2;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, hit);
        const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
        if (hit && pos < nsample) out[pos] = k;
        cnt += __popc(mask);
    }
    if (cnt > nsample) cnt = nsample;
    __syncwarp();
    if (cnt < nsample) {
        if (cnt > 0 && cnt < nsample - 1) {
            // reference: idx[cnt+k] = idx[k] sequentially == periodic extension of the found prefix (:96-102)
            for (int i = cnt + lane; i < nsample; i += 32) out[i] = out[i % cnt];
        } else {
            for (int i = cnt + lane; i < nsample; i += 32) out[i] = 0;  // cnt == nsample-1 (or 0): slot keeps zero init
        }
    }
}

// ------------------------------------------------------------------------------------------------ gather
__global__ void gather_bcn_kernel(const float* __restrict__ points, const int* __restrict__ idx, int C, int n, int m,
                                  float* __restrict__ out) {
    const int b = blockIdx.y;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)C * m) return;
    const int c = (int)(t / m), j = (int)(t % m);
    out[((size_t)b * C + c) * m + j] = __ldg(points + ((size_t)b * C + c) * n + __ldg(idx + (size_t)b * m + j));
}

// ------------------------------------------------------------------------------------------------ kNN
template <int NS>
__device__ __forceinline__ void knn_reheap(float* dist, int* idx, int k) {
    int root = 0, child = 1;
    while (child < k) {
        if (child + 1 < k && dist[child + 1] > dist[child]) child++;
        if (dist[root] > dist[child]) return;
        const float tf = dist[root]; dist[root] = dist[child]; dist[child] = tf;
        const int ti = idx[root]; idx[root] = idx[child]; idx[child] = ti;
        root = child;
        child = root * 2 + 1;
    }
}

constexpr int KNN_TILE = 512;

// NS > 0: compile-time heap size; NS == 0: runtime nsample (<= 100, as the reference's best_dist[100])
template <int NS>
__global__ void __launch_bounds__(64) knn_packed_kernel(int m, int nsample_rt, const float* __restrict__ xyz,
                                                         const float* __restrict__ new_xyz, const int* __restrict__ offset,
                                                         const int* __restrict__ new_offset, int nbatch,
                                                         int* __restrict__ idx, float* __restrict__ dist2) {
    constexpr int CAP = NS > 0 ? NS : 100;
    const int nsample = NS > 0 ? NS : nsample_rt;
    __shared__ float tile[KNN_TILE * 3];
    __shared__ int s_lo, s_hi;
    const int pt = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = pt < m;
    int start = 0, end = 0;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) {
        int bt = 0;
        while (bt < nbatch - 1 && !(pt < __ldg(new_offset + bt))) bt++;
        start = bt == 0 ? 0 : __ldg(offset + bt - 1);
        end = __ldg(offset + bt);
        qx = __ldg(new_xyz + (size_t)pt * 3); qy = __ldg(new_xyz + (size_t)pt * 3 + 1); qz = __ldg(new_xyz + (size_t)pt * 3 + 2);
    }
    if (threadIdx.x == 0) { s_lo = 0x7fffffff; s_hi = 0; }
    __syncthreads();
    if (active) { atomicMin(&s_lo, start); atomicMax(&s_hi, end); }
    __syncthreads();
    const int lo = s_lo, hi = s_hi;

    float best_dist[CAP];
    int best_idx[CAP];
    for (int i = 0; i < nsample; ++i) { best_dist[i] = 1e10f; best_idx[i] = start; }
    float root = 1e10f;

    for (int t0 = lo; t0 < hi; t0 += KNN_TILE) {
        const int cntp = min(KNN_TILE, hi - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < cntp * 3; i += blockDim.x) tile[i] = __ldg(xyz + (size_t)t0 * 3 + i);
        __syncthreads();
        if (active) {
            const int a = max(start, t0) - t0, e = min(end, t0 + cntp) - t0;
            int i = a;
            // 4 candidates per trip: the common case (none beats the heap root) costs one compare; insertions replay the
            // candidates in index order against the *current* root, exactly like the reference's serial scan
            for (; i + 3 < e; i += 4) {
                float d[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) d[u] = etch_sqdist3(qx - tile[(i + u) * 3], qy - tile[(i + u) * 3 + 1], qz - tile[(i + u) * 3 + 2]);
                if (fminf(fminf(d[0], d[1]), fminf(d[2], d[3])) < root) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (d[u] < root) {
                            best_dist[0] = d[u];
                            best_idx[0] = t0 + i + u;
                            knn_reheap<NS>(best_dist, best_idx, nsample);
                            root = best_dist[0];
                        }
                    }
                }
            }
            for (; i < e; ++i) {
                const float d2 = etch_sqdist3(qx - tile[i * 3], qy - tile[i * 3 + 1], qz - tile[i * 3 + 2]);
                if (d2 < root) {
                    best_dist[0] = d2;
                    best_idx[0] = t0 + i;
                    knn_reheap<NS>(best_dist, best_idx, nsample);
                    root = best_dist[0];
                }
            }
        }
    }
    if (!active) return;
    for (int i = nsample - 1; i > 0; i--) {  // heap_sort (:39-48)
        const float tf = best_dist[0]; best_dist[0] = best_dist[i]; best_dist[i] = tf;
        const int ti = best_idx[0]; best_idx[0] = best_idx[i]; best_idx[i] = ti;
        knn_reheap<NS>(best_dist, best_idx, i);
    }
    for (int i = 0; i < nsample; ++i) {
        idx[(size_t)pt * nsample + i] = best_idx[i];
        dist2[(size_t)pt * nsample + i] = best_dist[i];
    }
}

// ---- warp-per-query kNN (nsample <= 32) ----
// The reference keeps a max-heap of the nsample best and replaces its root whenever a candidate is strictly closer
// (knnquery_cuda_kernel.cu:50-75): the kept set is "the nsample smallest by (distance, scan order)" and, when no two kept
// distances are equal, the heap-sorted output is simply ascending distance.  A warp therefore keeps the list SORTED, one
// entry per lane: 32 candidates are evaluated per step, the few that beat the current nsample-th distance are inserted in
// index order with a ballot + shuffle shift.  If an insertion ever meets an equal distance already in the list (the only
// situation in which heap layout decides the result), the query is recomputed by knn_serial_exact, the literal emulation.
template <int NS>
__device__ __noinline__ void knn_serial_exact(int start, int end, float qx, float qy, float qz, const float* __restrict__ xyz,
                                              int* __restrict__ out_idx, float* __restrict__ out_d) {
    float best_dist[NS];
    int best_idx[NS];
    for (int i = 0; i < NS; ++i) { best_dist[i] = 1e10f; best_idx[i] = start; }
    for (int i = start; i < end; ++i) {
        const float d2 = etch_sqdist3(qx - __ldg(xyz + (size_t)i * 3), qy - __ldg(xyz + (size_t)i * 3 + 1), qz - __ldg(xyz + (size_t)i * 3 + 2));
        if (d2 < best_dist[0]) {
            best_dist[0] = d2;
            best_idx[0] = i;
            knn_reheap<NS>(best_dist, best_idx, NS);
        }
    }
    for (int i = NS - 1; i > 0; i--) {
        const float tf = best_dist[0]; best_dist[0] = best_dist[i]; best_dist[i] = tf;
        const int ti = best_idx[0]; best_idx[0] = best_idx[i]; best_idx[i] = ti;
        knn_reheap<NS>(best_dist, best_idx, i);
    }
    for (int i = 0; i < NS; ++i) { out_idx[i] = best_idx[i]; out_d[i] = best_dist[i]; }
}

// The same literal emulation, executed by a whole (converged) warp for ONE query with warp-uniform arguments: the 32 lanes compute
// the distances of 256 consecutive candidates into `buf` (shared memory, 256 floats owned by this warp), and lane 0 replays them in
// index order against its heap -- but only when some candidate of the chunk beats the heap root as it stood at the start of the
// chunk (the root only shrinks, so a chunk without such a candidate cannot change the heap).  After the first few chunks almost
// every chunk is skipped: a 5000-point fallback costs ~10 us instead of the ~550 us of the one-thread scan, which used to set the
// duration of the whole kernel as soon as a single query of the batch met a tie.
constexpr int KNNX_CHUNK = 256;
template <int NS>
__device__ __noinline__ void knn_exact_warp(int start, int end, float qx, float qy, float qz, const float* __restrict__ xyz,
                                            int* __restrict__ out_idx, float* __restrict__ out_d, float* buf) {
    const int lane = threadIdx.x & 31;
    float best_dist[NS];
    int best_idx[NS];
    for (int i = 0; i < NS; ++i) { best_dist[i] = 1e10f; best_idx[i] = start; }
    float root = 1e10f;                     // warp-uniform copy of lane 0's heap root
    for (int c0 = start; c0 < end; c0 += KNNX_CHUNK) {
        const int cnt = min(KNNX_CHUNK, end - c0);
        bool hit = false;
#pragma unroll
        for (int u = 0; u < KNNX_CHUNK / 32; ++u) {
            const int j = u * 32 + lane;
            if (j < cnt) {
                const size_t i = (size_t)(c0 + j) * 3;
                const float d2 = etch_sqdist3(qx - __ldg(xyz + i), qy - __ldg(xyz + i + 1), qz - __ldg(xyz + i + 2));
                buf[j] = d2;
                hit |= d2 < root;
            }
        }
        if (__any_sync(0xffffffffu, hit)) {
            __syncwarp();
            if (lane == 0) {
                for (int j = 0; j < cnt; ++j) {
                    const float d2 = buf[j];
                    if (d2 < best_dist[0]) {
                        best_dist[0] = d2;
                        best_idx[0] = c0 + j;
                        knn_reheap<NS>(best_dist, best_idx, NS);
                    }
                }
            }
            root = __shfl_sync(0xffffffffu, best_dist[0], 0);
        }
        __syncwarp();
    }
    if (lane == 0) {
        for (int i = NS - 1; i > 0; i--) {
            const float tf = best_dist[0]; best_dist[0] = best_dist[i]; best_dist[i] = tf;
            const int ti = best_idx[0]; best_idx[0] = best_idx[i]; best_idx[i] = ti;
            knn_reheap<NS>(best_dist, best_idx, i);
        }
        for (int i = 0; i < NS; ++i) { out_idx[i] = best_idx[i]; out_d[i] = best_dist[i]; }
    }
}

constexpr int KNNW_TILE = 1024;   // candidates staged per step (SoA, 12 KB)
constexpr int KNNW_WARPS = 8;

template <int NS>
__global__ void __launch_bounds__(KNNW_WARPS * 32) knn_warp_kernel(int m, const float* __restrict__ xyz, const float* __restrict__ new_xyz,
                                                                   const int* __restrict__ offset, const int* __restrict__ new_offset,
                                                                   int nbatch, int* __restrict__ idx, float* __restrict__ dist2) {
    static_assert(NS >= 1 && NS <= 32, "one list entry per lane");
    __shared__ float tx[KNNW_TILE], ty[KNNW_TILE], tz[KNNW_TILE];
    __shared__ float s_exact[KNNW_WARPS][KNNX_CHUNK];
    __shared__ int s_lo, s_hi;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pt = blockIdx.x * KNNW_WARPS + warp;
    const bool active = pt < m;
    int start = 0, end = 0;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) {
        int bt = 0;
        while (bt < nbatch - 1 && !(pt < __ldg(new_offset + bt))) bt++;
        start = bt == 0 ? 0 : __ldg(offset + bt - 1);
        end = __ldg(offset + bt);
        qx = __ldg(new_xyz + (size_t)pt * 3); qy = __ldg(new_xyz + (size_t)pt * 3 + 1); qz = __ldg(new_xyz + (size_t)pt * 3 + 2);
    }
    if (threadIdx.x == 0) { s_lo = 0x7fffffff; s_hi = 0; }
    __syncthreads();
    if (active && lane == 0) { atomicMin(&s_lo, start); atomicMax(&s_hi, end); }
    __syncthreads();
    const int lo = s_lo, hi = s_hi;

    float ld = 1e10f;       // lane l < NS: l-th smallest distance so far
    int li = start;
    float tau = 1e10f;      // NS-th smallest (warp-uniform)
    bool tie = false;

    for (int t0 = lo; t0 < hi; t0 += KNNW_TILE) {
        const int cntp = min(KNNW_TILE, hi - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < cntp * 3; i += KNNW_WARPS * 32) {
            const float v = __ldg(xyz + (size_t)t0 * 3 + i);
            const int c = i / 3, k = i - c * 3;
            (k == 0 ? tx : (k == 1 ? ty : tz))[c] = v;
        }
        __syncthreads();
        if (!active) continue;
        const int a = max(start, t0) - t0, e = min(end, t0 + cntp) - t0;
        for (int c0 = a; c0 < e; c0 += 32) {
            const int c = c0 + lane;
            float d = 3e38f;
            if (c < e) d = etch_sqdist3(qx - tx[c], qy - ty[c], qz - tz[c]);
            unsigned mask = __ballot_sync(0xffffffffu, d < tau);
            while (mask) {
                const int l = __ffs(mask) - 1;
                mask &= mask - 1;
                const float dc = __shfl_sync(0xffffffffu, d, l);
                if (dc < tau) {   // against the CURRENT nsample-th distance, as the serial scan would see it
                    const unsigned le = __ballot_sync(0xffffffffu, lane < NS && ld <= dc);
                    const int pos = __popc(le);
                    const unsigned eq = __ballot_sync(0xffffffffu, lane < NS && ld == dc);
                    tie |= eq != 0u;
                    const float pd = __shfl_up_sync(0xffffffffu, ld, 1);
                    const int pi = __shfl_up_sync(0xffffffffu, li, 1);
                    if (lane > pos) { ld = pd; li = pi; }
                    if (lane == pos) { ld = dc; li = t0 + c0 + l; }
                    tau = __shfl_sync(0xffffffffu, ld, NS - 1);
                }
            }
        }
    }
    if (!active) return;
    if (tie) {   // warp-uniform
        knn_exact_warp<NS>(start, end, qx, qy, qz, xyz, idx + (size_t)pt * NS, dist2 + (size_t)pt * NS, s_exact[warp]);
        return;
    }
    if (lane < NS) {
        idx[(size_t)pt * NS + lane] = li;
        dist2[(size_t)pt * NS + lane] = ld;
    }
}

// ---- grid-accelerated kNN (same results as the brute-force scan) ----
// The reference tests every query against every point of its segment (m x n distance tests, quadratic in the scan size).
// Here the candidates of each segment are binned into a uniform grid (cell size ~ the expected nsample-neighbour radius of a
// surface sampling), a query walks the cells around it shell by shell and stops when its nsample-th distance is strictly
// smaller than the distance to everything unexplored.  Without equal distances that set, sorted ascending, IS the reference's
// heap-sorted output; any tie inside the result or at its boundary (or a segment shorter than nsample) sends the query to
// knn_serial_exact, the literal emulation, so indices and distances stay bit-exact.
constexpr int KG_MAXCELLS = 65536;   // per segment
constexpr int KG_MAXDIM = 64;

struct __align__(16) KnnGrid { float ox, oy, oz, inv_h; int nx, ny, nz, ncell; float h; int pad0, pad1, pad2; };

__global__ void __launch_bounds__(256) knn_grid_setup_kernel(const float* __restrict__ xyz, const int* __restrict__ offset, int nsample,
                                                             KnnGrid* __restrict__ grids) {
    const int b = blockIdx.x;
    const int start = b == 0 ? 0 : __ldg(offset + b - 1), end = __ldg(offset + b);
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    for (int i = start + threadIdx.x; i < end; i += 256)
#pragma unroll
        for (int a = 0; a < 3; ++a) { const float v = __ldg(xyz + (size_t)i * 3 + a); lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v); }
    __shared__ float s_lo[3][8], s_hi[3][8];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o)); hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o)); }
        if ((threadIdx.x & 31) == 0) { s_lo[a][threadIdx.x >> 5] = lo[a]; s_hi[a][threadIdx.x >> 5] = hi[a]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        KnnGrid g;
        float e[3];
        float o3[3];
        for (int a = 0; a < 3; ++a) {
            float l = s_lo[a][0], h = s_hi[a][0];
            for (int w = 1; w < 8; ++w) { l = fminf(l, s_lo[a][w]); h = fmaxf(h, s_hi[a][w]); }
            if (end <= start) { l = 0.f; h = 0.f; }
            o3[a] = l;
            e[a] = fmaxf(h - l, 1e-6f);
        }
        const int n = max(end - start, 1);
        // expected nsample-neighbour radius of a surface sampling: half of the bounding-box surface as the area estimate
        const float area = e[0] * e[1] + e[1] * e[2] + e[2] * e[0];
        float h = 1.1f * sqrtf((float)nsample * area / (3.14159265f * (float)n));
        h = fmaxf(h, fmaxf(e[0], fmaxf(e[1], e[2])) / (float)KG_MAXDIM);
        int nx, ny, nz;
        for (;;) {
            nx = min(KG_MAXDIM, (int)(e[0] / h) + 1); ny = min(KG_MAXDIM, (int)(e[1] / h) + 1); nz = min(KG_MAXDIM, (int)(e[2] / h) + 1);
            if ((long long)nx * ny * nz <= KG_MAXCELLS) break;
            h *= 1.26f;
        }
        g.ox = o3[0]; g.oy = o3[1]; g.oz = o3[2]; g.inv_h = 1.0f / h; g.h = h;
        g.nx = nx; g.ny = ny; g.nz = nz; g.ncell = nx * ny * nz; g.pad0 = g.pad1 = g.pad2 = 0;
        grids[b] = g;
    }
}

__device__ __forceinline__ int kg_cell_coord(float v, float o, float inv_h, int n) {
    const int c = (int)floorf((v - o) * inv_h);
    return c < 0 ? 0 : (c >= n ? n - 1 : c);
}
__device__ __forceinline__ int kg_segment(int i, const int* __restrict__ offset, int nbatch) {
    int b = 0;
    while (b < nbatch - 1 && !(i < __ldg(offset + b))) ++b;
    return b;
}

// MODE 0: histogram; MODE 1: scatter (x, y, z, index) into cell order
template <int MODE>
__global__ void __launch_bounds__(256) knn_grid_bin_kernel(const float* __restrict__ xyz, const int* __restrict__ offset, int nbatch, int n,
                                                           const KnnGrid* __restrict__ grids, int* __restrict__ cells, float4* __restrict__ sorted) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int b = kg_segment(i, offset, nbatch);
    const KnnGrid g = grids[b];
    const float x = __ldg(xyz + (size_t)i * 3), y = __ldg(xyz + (size_t)i * 3 + 1), z = __ldg(xyz + (size_t)i * 3 + 2);
    const int c = (kg_cell_coord(z, g.oz, g.inv_h, g.nz) * g.ny + kg_cell_coord(y, g.oy, g.inv_h, g.ny)) * g.nx + kg_cell_coord(x, g.ox, g.inv_h, g.nx);
    int* cell = cells + (size_t)b * (KG_MAXCELLS + 1) + c;
    if (MODE == 0) atomicAdd(cell, 1);
    else sorted[atomicAdd(cell, 1)] = make_float4(x, y, z, __int_as_float(i));
}

// exclusive scan of the per-cell counts of one segment -> first slot of every cell (global slot = segment start + prefix);
// `starts` keeps the result, `cursor` is a working copy for the scatter pass
__global__ void __launch_bounds__(1024) knn_grid_scan_kernel(const int* __restrict__ offset, const KnnGrid* __restrict__ grids,
                                                             int* __restrict__ starts, int* __restrict__ cursor) {
    const int b = blockIdx.x, tid = threadIdx.x;
    const int base = b == 0 ? 0 : __ldg(offset + b - 1);
    const int ncell = grids[b].ncell;
    int* st = starts + (size_t)b * (KG_MAXCELLS + 1);
    int* cu = cursor + (size_t)b * (KG_MAXCELLS + 1);
    const int per = (ncell + 1023) / 1024;
    const int c0 = tid * per, c1 = min(ncell, c0 + per);
    int sum = 0;
    for (int c = c0; c < c1; ++c) sum += st[c];
    __shared__ int s_part[1024];
    s_part[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = tid >= o ? s_part[tid - o] : 0;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    int run = base + s_part[tid] - sum;
    for (int c = c0; c < c1; ++c) { const int cnt = st[c]; st[c] = run; cu[c] = run; run += cnt; }
    if (tid == 1023) st[ncell] = base + s_part[1023];
}

// diagnostic counters (tools/knn_probe.py): queries answered by the exact brute-force fallback, and candidates visited by the grid walk
__device__ unsigned long long g_knn_fallbacks = 0ull, g_knn_candidates = 0ull;

template <int NS>
__global__ void __launch_bounds__(64) knn_grid_query_kernel(int m, const float* __restrict__ xyz, const float* __restrict__ new_xyz,
                                                             const int* __restrict__ offset, const int* __restrict__ new_offset, int nbatch,
                                                             const KnnGrid* __restrict__ grids, const int* __restrict__ starts,
                                                             const float4* __restrict__ sorted, int* __restrict__ idx, float* __restrict__ dist2) {
    __shared__ float s_exact[2][KNNX_CHUNK];
    const int pt = blockIdx.x * 64 + threadIdx.x;
    const bool active = pt < m;
    const int ptc = active ? pt : m - 1;                 // inactive lanes of the last block shadow the last query and write nothing
    const int b = kg_segment(ptc, new_offset, nbatch);
    const int start = b == 0 ? 0 : __ldg(offset + b - 1), end = __ldg(offset + b);
    const float qx = __ldg(new_xyz + (size_t)ptc * 3), qy = __ldg(new_xyz + (size_t)ptc * 3 + 1), qz = __ldg(new_xyz + (size_t)ptc * 3 + 2);
    int* oi = idx + (size_t)ptc * NS;
    float* od = dist2 + (size_t)ptc * NS;
    bool need_exact = active && (end - start < NS);      // short segment: the reference's (1e10, start) fillers come out of the heap emulation
    const bool walk = active && !need_exact;
    const KnnGrid g = grids[b];
    const int* st = starts + (size_t)b * (KG_MAXCELLS + 1);
    const int cx = kg_cell_coord(qx, g.ox, g.inv_h, g.nx), cy = kg_cell_coord(qy, g.oy, g.inv_h, g.ny), cz = kg_cell_coord(qz, g.oz, g.inv_h, g.nz);
    float ld[NS];
    int li[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) { ld[i] = 3e38f; li[i] = start; }
    float rej = 3e38f;      // smallest distance that is NOT in the list
    int cnt = 0;
    const int rmax = walk ? max(g.nx, max(g.ny, g.nz)) : -1;
    for (int r = 0; r <= rmax; ++r) {
        const int z0 = max(cz - r, 0), z1 = min(cz + r, g.nz - 1), y0 = max(cy - r, 0), y1 = min(cy + r, g.ny - 1);
        const int x0 = max(cx - r, 0), x1 = min(cx + r, g.nx - 1);
        for (int z = z0; z <= z1; ++z)
            for (int y = y0; y <= y1; ++y) {
                const bool face = (z == cz - r) || (z == cz + r) || (y == cy - r) || (y == cy + r);
                // cells of this (z, y) row that belong to shell r: the whole row on a face, else only the two x ends
                for (int xr = 0; xr < 2; ++xr) {
                    int xa, xb;
                    if (face) { if (xr == 1) break; xa = x0; xb = x1; }
                    else {
                        const int xe = xr == 0 ? cx - r : cx + r;
                        if (xe < 0 || xe >= g.nx || (xr == 1 && r == 0)) continue;
                        xa = xb = xe;
                    }
                    const int crow = (z * g.ny + y) * g.nx;
                    const int s0 = __ldg(st + crow + xa), s1 = __ldg(st + crow + xb + 1);   // consecutive cells are consecutive slots
                    for (int sb = s0; sb < s1; sb += 4) {       // 4 candidates in flight: the walk is bound by load latency
                        float4 cv[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) cv[u] = __ldg(sorted + min(sb + u, s1 - 1));
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (sb + u >= s1) break;
                            const float4 c = cv[u];
                            const float d = etch_sqdist3(qx - c.x, qy - c.y, qz - c.z);
                            ++cnt;
                            if (d < ld[NS - 1]) {
                                rej = fminf(rej, ld[NS - 1]);
                                float cd = d;
                                int ci = __float_as_int(c.w);
#pragma unroll
                                for (int i = 0; i < NS; ++i) {      // sorted insertion by swapping through the list
                                    if (cd < ld[i]) { const float td = ld[i]; const int ti = li[i]; ld[i] = cd; li[i] = ci; cd = td; ci = ti; }
                                }
                            } else {
                                rej = fminf(rej, d);
                            }
                        }
                    }
                }
            }
        if (cnt >= NS) {
            // distance from the query to everything outside the explored cube (faces beyond the grid have nothing behind them)
            float dout = 3e38f;
            if (cx - r > 0) dout = fminf(dout, qx - (g.ox + (float)(cx - r) * g.h));
            if (cx + r + 1 < g.nx) dout = fminf(dout, g.ox + (float)(cx + r + 1) * g.h - qx);
            if (cy - r > 0) dout = fminf(dout, qy - (g.oy + (float)(cy - r) * g.h));
            if (cy + r + 1 < g.ny) dout = fminf(dout, g.oy + (float)(cy + r + 1) * g.h - qy);
            if (cz - r > 0) dout = fminf(dout, qz - (g.oz + (float)(cz - r) * g.h));
            if (cz + r + 1 < g.nz) dout = fminf(dout, g.oz + (float)(cz + r + 1) * g.h - qz);
            if (dout >= 3e38f) break;                                  // the cube covers the whole grid
            dout -= 1e-3f * g.h;                                       // margin for the rounding of the cell assignment
            if (dout > 0.f && ld[NS - 1] < dout * dout) break;
        }
    }
    bool tie = rej <= ld[NS - 1];
#pragma unroll
    for (int i = 1; i < NS; ++i) tie |= ld[i] == ld[i - 1];
    if (walk) {
#ifdef ETCH_KNN_STATS
        atomicAdd(&g_knn_candidates, (unsigned long long)cnt);
        if (tie || cnt < NS) atomicAdd(&g_knn_fallbacks, 1ull);
#endif
        if (tie || cnt < NS) need_exact = true;
        else {
#pragma unroll
            for (int i = 0; i < NS; ++i) { oi[i] = li[i]; od[i] = ld[i]; }
        }
    }
    // the rare queries that need the literal emulation are served one after the other by the whole warp
    unsigned todo = __ballot_sync(0xffffffffu, need_exact);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int xs = __shfl_sync(0xffffffffu, start, src), xe = __shfl_sync(0xffffffffu, end, src), xp = __shfl_sync(0xffffffffu, ptc, src);
        const float x = __shfl_sync(0xffffffffu, qx, src), y = __shfl_sync(0xffffffffu, qy, src), z = __shfl_sync(0xffffffffu, qz, src);
        knn_exact_warp<NS>(xs, xe, x, y, z, xyz, idx + (size_t)xp * NS, dist2 + (size_t)xp * NS, s_exact[threadIdx.x >> 5]);
    }
}

}  // namespace

static int g_sm_budget = 148;
int etch_sm_budget() { return g_sm_budget; }

// ================================================================================================ C ABI
// number of SMs the persistent kernels size their grids for (default 148 = all of a B200)
ETCH_API int etch_set_sm_budget(int sms) {
    if (sms < 1 || sms > 148) return ETCH_EINVAL;
    g_sm_budget = sms;
    return ETCH_OK;
}

// replaces epn_grouping.furthest_point_sampling (external/vgtk/vgtk/cuda/grouping_cuda.cpp:160-174)
ETCH_API int etch_fps_bcn(const float* xyz, int B, int n, int m, int* idx, cudaStream_t stream) {
    if (!xyz || !idx || B <= 0 || n <= 0 || m < 0 || n > 28672) return ETCH_EINVAL;
    if (m == 0) return ETCH_OK;
    const int T = etch_opt_n_threads(n);
    if (n > FPS_CLUSTER_MIN) {   // large clouds: a cluster of 4 (n <= 16k) or 8 CTAs per scan, <= 4 points per thread
        const int cl = n <= 16384 ? 4 : 8;
        const int per = (n + cl - 1) / cl, cppt = (per + 1023) / 1024;
        const size_t sm = (size_t)per * 12;
        const int* np_ = nullptr;
#define LC(P)                                                                                                              \
    return cl == 4 ? fps_cluster_launch(fps_cluster_kernel<P, false, 4>, B, 4, sm, stream, xyz, np_, np_, n, m, T, idx)    \
                   : fps_cluster_launch(fps_cluster_kernel<P, false, 8>, B, 8, sm, stream, xyz, np_, np_, n, m, T, idx);
        if (cppt <= 1) { LC(1) } else if (cppt <= 2) { LC(2) } else if (cppt <= 3) { LC(3) } else { LC(4) }
#undef LC
    }
    const int nt = n >= 1024 ? 1024 : ((n + 31) / 32) * 32;
    const int ppt = (n + nt - 1) / nt;
#define L(P, SX)                                                                                               \
    {                                                                                                          \
        auto kern = fps_bcn_kernel<P, SX>;                                                                     \
        const size_t sm = (SX) ? (size_t)n * 8 : (size_t)n * 12;                                               \
        if (sm > 48 * 1024) ETCH_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
        kern<<<B, nt, sm, stream>>>(xyz, n, m, T, idx);                                                        \
    }
    if (ppt <= 1) L(1, false) else if (ppt <= 2) L(2, false) else if (ppt <= 3) L(3, false) else if (ppt <= 4) L(4, false)
    else if (ppt <= 5) L(5, false) else if (ppt <= 6) L(6, false) else if (ppt <= 8) L(8, false) else if (ppt <= 10) L(10, false)
    else if (ppt <= 12) L(12, true) else if (ppt <= 16) L(16, true) else if (ppt <= 20) L(20, true) else if (ppt <= 24) L(24, true)
    else L(28, true)
#undef L
    ETCH_RETURN_LAST();
}

// replaces pointops_cuda.furthestsampling_cuda (external/pointops/src/sampling/sampling_cuda.cpp:8-16).
// `tmp` (the reference's 1e10-filled scratch) is accepted for signature parity and left untouched.
ETCH_API int etch_fps_packed(int b, int n_max, const float* xyz, const int* offset, const int* new_offset, float* tmp,
                             int* idx, cudaStream_t stream) {
    (void)tmp;
    if (!xyz || !offset || !new_offset || !idx || b <= 0 || n_max <= 0 || n_max > 28672) return ETCH_EINVAL;
    const int T = etch_opt_n_threads(n_max);
    if (n_max > FPS_CLUSTER_MIN) {
        const int cl = n_max <= 16384 ? 4 : 8;
        const int per = (n_max + cl - 1) / cl, cppt = (per + 1023) / 1024;
        const size_t sm = (size_t)per * 12;
#define LC(P)                                                                                                                      \
    return cl == 4 ? fps_cluster_launch(fps_cluster_kernel<P, true, 4>, b, 4, sm, stream, xyz, offset, new_offset, 0, 0, T, idx)    \
                   : fps_cluster_launch(fps_cluster_kernel<P, true, 8>, b, 8, sm, stream, xyz, offset, new_offset, 0, 0, T, idx);
        if (cppt <= 1) { LC(1) } else if (cppt <= 2) { LC(2) } else if (cppt <= 3) { LC(3) } else { LC(4) }
#undef LC
    }
    const int nt = n_max >= 1024 ? 1024 : ((n_max + 31) / 32) * 32;
    const int ppt = (n_max + nt - 1) / nt;
#define L(P, SX)                                                                                               \
    {                                                                                                          \
        auto kern = fps_packed_kernel<P, SX>;                                                                  \
        const size_t sm = (SX) ? (size_t)n_max * 8 : (size_t)n_max * 12;                                       \
        if (sm > 48 * 1024) ETCH_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
        kern<<<b, nt, sm, stream>>>(xyz, offset, new_offset, T, idx);                                          \
    }
    if (ppt <= 1) L(1, false) else if (ppt <= 2) L(2, false) else if (ppt <= 3) L(3, false) else if (ppt <= 4) L(4, false)
    else if (ppt <= 5) L(5, false) else if (ppt <= 6) L(6, false) else if (ppt <= 8) L(8, false) else if (ppt <= 10) L(10, false)
    else if (ppt <= 12) L(12, true) else if (ppt <= 16) L(16, true) else if (ppt <= 20) L(20, true) else if (ppt <= 24) L(24, true)
    else L(28, true)
#undef L
    ETCH_RETURN_LAST();
}

// replaces epn_grouping.ball_query (grouping_cuda.cpp:71-86); idx [B,m,nsample] is fully written
ETCH_API int etch_ball_query_bcn(const float* new_xyz, const float* xyz, int B, int m, int n, float radius, int nsample,
                                 int* idx, cudaStream_t stream) {
    if (!new_xyz || !xyz || !idx || B <= 0 || m <= 0 || n <= 0 || nsample <= 0) return ETCH_EINVAL;
    const float radius2 = radius * radius;
    dim3 grid(etch_cdiv(m, 8), B);
    ball_query_kernel<<<grid, 256, 0, stream>>>(new_xyz, xyz, n, m, radius2, nsample, idx);
    ETCH_RETURN_LAST();
}

// replaces epn_gathering.gather_points_forward (gathering_cuda.cpp:29-43)
ETCH_API int etch_gather_bcn(const float* points, const int* idx, int B, int C, int n, int m, float* out,
                             cudaStream_t stream) {
    if (!points || !idx || !out || B <= 0 || C <= 0 || n <= 0 || m <= 0) return ETCH_EINVAL;
    dim3 grid((unsigned)etch_cdiv((size_t)C * m, (size_t)256), B);
    gather_bcn_kernel<<<grid, 256, 0, stream>>>(points, idx, C, n, m, out);
    ETCH_RETURN_LAST();
}

// replaces pointops_cuda.knnquery_cuda (external/pointops/src/knnquery/knnquery_cuda.cpp:8-17); nbatch = len(offset)
ETCH_API int etch_knn_packed(int m, int nsample, const float* xyz, const float* new_xyz, const int* offset,
                             const int* new_offset, int nbatch, int* idx, float* dist2, cudaStream_t stream) {
    if (!xyz || !new_xyz || !offset || !new_offset || !idx || !dist2 || m <= 0 || nsample <= 0 || nsample > 100 || nbatch <= 0)
        return ETCH_EINVAL;
    const int grid = etch_cdiv(m, 64), wgrid = etch_cdiv(m, KNNW_WARPS);
    if (nsample == 3) knn_warp_kernel<3><<<wgrid, KNNW_WARPS * 32, 0, stream>>>(m, xyz, new_xyz, offset, new_offset, nbatch, idx, dist2);
    else if (nsample == 8) knn_warp_kernel<8><<<wgrid, KNNW_WARPS * 32, 0, stream>>>(m, xyz, new_xyz, offset, new_offset, nbatch, idx, dist2);
    else if (nsample == 16) knn_warp_kernel<16><<<wgrid, KNNW_WARPS * 32, 0, stream>>>(m, xyz, new_xyz, offset, new_offset, nbatch, idx, dist2);
    else knn_packed_kernel<0><<<grid, 64, 0, stream>>>(m, nsample, xyz, new_xyz, offset, new_offset, nbatch, idx, dist2);
    ETCH_RETURN_LAST();
}

// Grid-accelerated kNN with the results of etch_knn_packed (bit-exact, see knn_grid_query_kernel); nsample in {3, 8, 16}.
// n = rows of xyz.  scratch: caller-owned, etch_knn_grid_scratch_bytes(n, nbatch) bytes, 16-byte aligned.
ETCH_API long long etch_knn_grid_scratch_bytes(int n, int nbatch) {
    return (long long)nbatch * 64 + 2ll * nbatch * (KG_MAXCELLS + 1) * 4 + 16 + (long long)n * 16;
}
ETCH_API int etch_knn_grid(int m, int nsample, const float* xyz, int n, const float* new_xyz, const int* offset, const int* new_offset,
                           int nbatch, int* idx, float* dist2, void* scratch, cudaStream_t stream) {
    if (!xyz || !new_xyz || !offset || !new_offset || !idx || !dist2 || !scratch || m <= 0 || n <= 0 || nbatch <= 0) return ETCH_EINVAL;
    if (nsample != 3 && nsample != 8 && nsample != 16) return ETCH_EINVAL;
    unsigned char* sp = static_cast<unsigned char*>(scratch);
    KnnGrid* grids = reinterpret_cast<KnnGrid*>(sp);
    int* starts = reinterpret_cast<int*>(sp + (size_t)nbatch * 64);
    int* cursor = starts + (size_t)nbatch * (KG_MAXCELLS + 1);
    float4* sorted = reinterpret_cast<float4*>((reinterpret_cast<uintptr_t>(cursor + (size_t)nbatch * (KG_MAXCELLS + 1)) + 15) & ~(uintptr_t)15);
    ETCH_TRY(cudaMemsetAsync(starts, 0, (size_t)nbatch * (KG_MAXCELLS + 1) * 4, stream));
    knn_grid_setup_kernel<<<nbatch, 256, 0, stream>>>(xyz, offset, nsample, grids);
    knn_grid_bin_kernel<0><<<etch_cdiv(n, 256), 256, 0, stream>>>(xyz, offset, nbatch, n, grids, starts, sorted);
    knn_grid_scan_kernel<<<nbatch, 1024, 0, stream>>>(offset, grids, starts, cursor);
    knn_grid_bin_kernel<1><<<etch_cdiv(n, 256), 256, 0, stream>>>(xyz, offset, nbatch, n, grids, cursor, sorted);
    const int grid = etch_cdiv(m, 64);
    if (nsample == 3) knn_grid_query_kernel<3><<<grid, 64, 0, stream>>>(m, xyz, new_xyz, offset, new_offset, nbatch, grids, starts, sorted, idx, dist2);
    else if (nsample == 8) knn_grid_query_kernel<8><<<grid, 64, 0, stream>>>(m, xyz, new_xyz, offset, new_offset, nbatch, grids, starts, sorted, idx, dist2);
    else knn_grid_query_kernel<16><<<grid, 64, 0, stream>>>(m, xyz, new_xyz, offset, new_offset, nbatch, grids, starts, sorted, idx, dist2);
    ETCH_RETURN_LAST();
}

ETCH_API int etch_opt_threads(int work_size) { return etch_opt_n_threads(work_size); }

// probe (include/etch_b200_probes.h): reads and resets the kNN-grid diagnostic counters; they only count when the library is built
// with -DETCH_KNN_STATS (tools/knn_probe.py does that in a scratch copy), out[0] = fallback queries, out[1] = candidates visited
ETCH_API int etch_knn_grid_stats(unsigned long long* out_host) {
    if (!out_host) return ETCH_EINVAL;
    unsigned long long z[2] = {0ull, 0ull};
    ETCH_TRY(cudaMemcpyFromSymbol(&out_host[0], g_knn_fallbacks, sizeof(unsigned long long)));
    ETCH_TRY(cudaMemcpyFromSymbol(&out_host[1], g_knn_candidates, sizeof(unsigned long long)));
    ETCH_TRY(cudaMemcpyToSymbol(g_knn_fallbacks, &z[0], sizeof(unsigned long long)));
    ETCH_TRY(cudaMemcpyToSymbol(g_knn_candidates, &z[1], sizeof(unsigned long long)));
    return ETCH_OK;
}
