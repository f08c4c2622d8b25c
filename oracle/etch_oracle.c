/*
 * oracle/etch_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded CPU restatement of the five native index kernels on ETCH's
 * inference-and-fit hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product path
 * (etch_b200/) never links, imports or calls it.
 *
 * Every routine emulates the reference CUDA kernel *literally* (same thread striding,
 * same shared-memory tree reduction, same FP32 expression tree including the
 * FMUL/FFMA/FFMA contraction nvcc emits for  dx*dx + dy*dy + dz*dz  -- verified with
 * cuobjdump on the reference sources compiled for sm_100a), so ties resolve exactly as
 * on the reference.  Compile with -ffp-contract=off: all fused operations are explicit.
 *
 * Reference sources followed (file:line, relative to the ETCH tree):
 *   external/vgtk/vgtk/cuda/grouping_cuda_kernel.cu:29-33    opt_n_threads
 *   external/vgtk/vgtk/cuda/grouping_cuda_kernel.cu:67-113   ball_query_cuda_kernel
 *   external/vgtk/vgtk/cuda/grouping_cuda_kernel.cu:339-466  __update + furthest_point_sampling_cuda_kernel
 *   external/vgtk/vgtk/cuda/gathering_cuda_kernel.cu:42-68   gather_points_forward_kernel
 *   external/pointops/src/knnquery/knnquery_cuda_kernel.cu:21-108   reheap / heap_sort / knnquery_cuda_kernel
 *   external/pointops/src/sampling/sampling_cuda_kernel.cu:5-129    furthestsampling_cuda_kernel
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* grouping_cuda_kernel.cu:29-33 / pointops cuda_utils.h:10-13 */
int etch_oracle_opt_n_threads(int work_size) {
    const int pow_2 = (int)(log((double)work_size) / log(2.0));
    int t = 1 << pow_2;
    if (t > 1024) t = 1024;
    if (t < 1) t = 1;
    return t;
}

/* (a)*(a) + (b)*(b) + (c)*(c) as nvcc contracts it in the reference kernels: FMUL(dy,dy), FFMA(dx,dx,.), FFMA(dz,dz,.) --
 * the SECOND product is the plain multiply.  Read off the SASS of the reference's own .cu files built for sm_100a
 * (oracle/_ref) and pinned by running them against this emulation on the GPU box (tests/test_ref_kernels_gpu.py). */
static inline float sqdist3(float dx, float dy, float dz) {
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* tree reduction of grouping_cuda_kernel.cu:339-346,400-460 (and sampling_cuda_kernel.cu:5-10,66-123) */
static void tree_reduce(float *dists, int *dists_i, int T) {
    for (int s = T / 2; s >= 1; s >>= 1) {
        for (int tid = 0; tid < s; ++tid) {
            const float v1 = dists[tid], v2 = dists[tid + s];
            const int i1 = dists_i[tid], i2 = dists_i[tid + s];
            dists[tid] = v1 > v2 ? v1 : (v2 > v1 ? v2 : v1); /* max(v1,v2) */
            dists_i[tid] = v2 > v1 ? i2 : i1;
        }
    }
}

/* vgtk FPS. xyz [B,3,n] channel-major; idx [B,m]. grouping_cuda_kernel.cu:351-466, wrapper grouping_cuda.cpp:160-174 */
void etch_oracle_fps_bcn(const float *xyz, int B, int n, int m, int *idx) {
    const int T = etch_oracle_opt_n_threads(n);
    float *temp = (float *)malloc(sizeof(float) * (size_t)n);
    float dists[1024];
    int dists_i[1024];
    for (int b = 0; b < B; ++b) {
        const float *X = xyz + (size_t)b * 3 * n, *Y = X + n, *Z = Y + n;
        int *out = idx + (size_t)b * m;
        for (int k = 0; k < n; ++k) temp[k] = 1e10f;
        if (m <= 0) continue;
        int old = 0;
        out[0] = 0;
        for (int j = 1; j < m; ++j) {
            const float x1 = X[old], y1 = Y[old], z1 = Z[old];
            for (int tid = 0; tid < T; ++tid) {
                int besti = 0;
                float best = -1.0f;
                for (int k = tid; k < n; k += T) {
                    const float x2 = X[k], y2 = Y[k], z2 = Z[k];
                    const float mag = sqdist3(x2, y2, z2);
                    if ((double)mag <= 1e-3) continue; /* origin skip BEFORE the temp update (:385-387) */
                    const float d = sqdist3(x2 - x1, y2 - y1, z2 - z1);
                    const float d2 = d < temp[k] ? d : temp[k];
                    temp[k] = d2;
                    besti = d2 > best ? k : besti;
                    best = d2 > best ? d2 : best;
                }
                dists[tid] = best;
                dists_i[tid] = besti;
            }
            tree_reduce(dists, dists_i, T);
            old = dists_i[0];
            out[j] = old;
        }
    }
    free(temp);
}

/* pointops FPS on packed rows. xyz [n,3]; offset/new_offset [B] cumulative; tmp [n] in/out; idx [m_total].
 * sampling_cuda_kernel.cu:15-129; block size from n_max (launcher :131-171). */
void etch_oracle_fps_packed(int B, int n_max, const float *xyz, const int *offset, const int *new_offset,
                            float *tmp, int *idx) {
    const int T = etch_oracle_opt_n_threads(n_max);
    float dists[1024];
    int dists_i[1024];
    for (int bid = 0; bid < B; ++bid) {
        int start_n, end_n, start_m, end_m, old;
        if (bid == 0) { start_n = 0; end_n = offset[0]; start_m = 0; end_m = new_offset[0]; old = 0; }
        else { start_n = offset[bid - 1]; end_n = offset[bid]; start_m = new_offset[bid - 1]; end_m = new_offset[bid]; old = offset[bid - 1]; }
        idx[start_m] = start_n; /* written unconditionally by tid 0 (:44), even if the segment is empty */
        for (int j = start_m + 1; j < end_m; ++j) {
            const float x1 = xyz[old * 3 + 0], y1 = xyz[old * 3 + 1], z1 = xyz[old * 3 + 2];
            for (int tid = 0; tid < T; ++tid) {
                int besti = start_n;
                float best = -1.0f;
                for (int k = start_n + tid; k < end_n; k += T) {
                    const float x2 = xyz[k * 3 + 0], y2 = xyz[k * 3 + 1], z2 = xyz[k * 3 + 2];
                    const float d = sqdist3(x2 - x1, y2 - y1, z2 - z1);
                    const float d2 = d < tmp[k] ? d : tmp[k];
                    tmp[k] = d2;
                    besti = d2 > best ? k : besti;
                    best = d2 > best ? d2 : best;
                }
                dists[tid] = best;
                dists_i[tid] = besti;
            }
            tree_reduce(dists, dists_i, T);
            old = dists_i[0];
            idx[j] = old;
        }
    }
}

/* vgtk ball query. new_xyz [B,3,m], xyz [B,3,n] -> idx [B,m,nsample] (zero-filled first: grouping_cuda.cpp:80-82).
 * grouping_cuda_kernel.cu:67-113 */
void etch_oracle_ball_query_bcn(const float *new_xyz, const float *xyz, int B, int m, int n, float radius,
                                int nsample, int *idx) {
    memset(idx, 0, sizeof(int) * (size_t)B * m * nsample);
    const float radius2 = radius * radius;
    for (int b = 0; b < B; ++b) {
        const float *X = xyz + (size_t)b * 3 * n, *Y = X + n, *Z = Y + n;
        const float *QX = new_xyz + (size_t)b * 3 * m, *QY = QX + m, *QZ = QY + m;
        int *out = idx + (size_t)b * m * nsample;
        for (int j = 0; j < m; ++j) {
            const float nx = QX[j], ny = QY[j], nz = QZ[j];
            int cnt = 0;
            for (int k = 0; k < n && cnt < nsample; ++k) {
                const float d2 = sqdist3(nx - X[k], ny - Y[k], nz - Z[k]);
                if (d2 < radius2) { out[j * nsample + cnt] = k; ++cnt; }
            }
            if (cnt < nsample - 1) {
                for (int k = 0; k + cnt < nsample; ++k) out[j * nsample + k + cnt] = out[j * nsample + k];
            }
        }
    }
}

/* gather_points_forward: out[b,c,j] = points[b,c,idx[b,j]]. gathering_cuda_kernel.cu:42-68 */
void etch_oracle_gather_bcn(const float *points, int B, int C, int n, const int *idx, int m, float *out) {
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int j = 0; j < m; ++j)
                out[((size_t)b * C + c) * m + j] = points[((size_t)b * C + c) * n + idx[(size_t)b * m + j]];
}

/* knnquery_cuda_kernel.cu:21-36 */
static void reheap(float *dist, int *idx, int k) {
    int root = 0, child = 1;
    while (child < k) {
        if (child + 1 < k && dist[child + 1] > dist[child]) child++;
        if (dist[root] > dist[child]) return;
        float tf = dist[root]; dist[root] = dist[child]; dist[child] = tf;
        int ti = idx[root]; idx[root] = idx[child]; idx[child] = ti;
        root = child;
        child = root * 2 + 1;
    }
}

/* knnquery_cuda_kernel.cu:39-48 */
static void heap_sort(float *dist, int *idx, int k) {
    for (int i = k - 1; i > 0; i--) {
        float tf = dist[0]; dist[0] = dist[i]; dist[i] = tf;
        int ti = idx[0]; idx[0] = idx[i]; idx[i] = ti;
        reheap(dist, idx, i);
    }
}

/* pointops kNN on packed rows: xyz [n,3], new_xyz [m,3] -> idx [m,nsample], dist2 [m,nsample].
 * knnquery_cuda_kernel.cu:51-108 */
void etch_oracle_knn_packed(int m, int nsample, const float *xyz, const float *new_xyz, const int *offset,
                            const int *new_offset, int *idx, float *dist2) {
    float best_dist[100];
    int best_idx[100];
    for (int pt = 0; pt < m; ++pt) {
        int bt = 0;
        while (!(pt < new_offset[bt])) bt++;
        const int start = bt == 0 ? 0 : offset[bt - 1];
        const int end = offset[bt];
        const float nx = new_xyz[pt * 3 + 0], ny = new_xyz[pt * 3 + 1], nz = new_xyz[pt * 3 + 2];
        for (int i = 0; i < nsample; ++i) { best_dist[i] = 1e10f; best_idx[i] = start; }
        for (int i = start; i < end; ++i) {
            const float d2 = sqdist3(nx - xyz[i * 3 + 0], ny - xyz[i * 3 + 1], nz - xyz[i * 3 + 2]);
            if (d2 < best_dist[0]) {
                best_dist[0] = d2;
                best_idx[0] = i;
                reheap(best_dist, best_idx, nsample);
            }
        }
        heap_sort(best_dist, best_idx, nsample);
        for (int i = 0; i < nsample; ++i) {
            idx[(size_t)pt * nsample + i] = best_idx[i];
            dist2[(size_t)pt * nsample + i] = best_dist[i];
        }
    }
}
