/*
 * etch_b200.h -- C ABI of libetch_b200.so: the B200-native (sm_100a) implementation of ETCH's per-scan
 * inference-and-fit hot path.
 *
 * Conventions (every entry point):
 *   - extern "C", plain device pointers + sizes, no torch types; caller owns all memory; nothing is allocated inside;
 *   - returns 0 on success, a positive cudaError_t on a CUDA failure, ETCH_EINVAL (-1) on a bad argument;
 *   - work is enqueued on `stream` and the call returns immediately (no host synchronisation);
 *   - re-entrant and thread-safe; the ONE piece of process-wide state is the tuning knob etch_set_sm_budget (grid size of
 *     the persistent kernels; set it once at start-up, not per call); pointers must be device memory of the current device,
 *     fp32 tensors contiguous in the stated layout.
 * File:line citations refer to the reference tree (boqian-li/ETCH) and name the interface each entry replaces.
 * INTEGRATION.md shows the reference-side binding (pybind/ctypes stub) a maintainer would add.
 */
#ifndef ETCH_B200_H
#define ETCH_B200_H

#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ETCH_OK 0
#define ETCH_EINVAL (-1)
#define ETCH_EUNSUPPORTED (-2)

/* ---- native extension entry points of the reference (boundary B2) ------------------------- */

/* epn_grouping.furthest_point_sampling   external/vgtk/vgtk/cuda/grouping_cuda.cpp:160-174 (kernel grouping_cuda_kernel.cu:351-466)
 * xyz [B,3,n] -> idx [B,m]; start index 0, points with |p|^2 <= 1e-3 skipped, reference tie rule reproduced. n <= 28672. */
int etch_fps_bcn(const float* xyz, int B, int n, int m, int* idx, cudaStream_t stream);

/* pointops_cuda.furthestsampling_cuda    external/pointops/src/sampling/sampling_cuda.cpp:8-16 (kernel sampling_cuda_kernel.cu:15-129)
 * packed xyz [n,3], cumulative offset/new_offset [b] -> idx [new_offset[b-1]] (global row ids). tmp is unused (may be NULL). */
int etch_fps_packed(int b, int n_max, const float* xyz, const int* offset, const int* new_offset, float* tmp, int* idx,
                    cudaStream_t stream);

/* epn_grouping.ball_query                external/vgtk/vgtk/cuda/grouping_cuda.cpp:71-86 (kernel grouping_cuda_kernel.cu:67-113)
 * new_xyz [B,3,m], xyz [B,3,n] -> idx [B,m,nsample] (first-nsample-in-index-order, cyclic padding, zero last slot quirk). */
int etch_ball_query_bcn(const float* new_xyz, const float* xyz, int B, int m, int n, float radius, int nsample, int* idx,
                        cudaStream_t stream);

/* epn_gathering.gather_points_forward    external/vgtk/vgtk/cuda/gathering_cuda.cpp:29-43 (kernel gathering_cuda_kernel.cu:42-68)
 * points [B,C,n], idx [B,m] -> out [B,C,m]. */
int etch_gather_bcn(const float* points, const int* idx, int B, int C, int n, int m, float* out, cudaStream_t stream);

/* pointops_cuda.knnquery_cuda            external/pointops/src/knnquery/knnquery_cuda.cpp:8-17 (kernel knnquery_cuda_kernel.cu:65-108)
 * xyz [n,3], new_xyz [m,3], offsets [nbatch] -> idx [m,nsample], dist2 [m,nsample] ascending; nsample <= 100. */
int etch_knn_packed(int m, int nsample, const float* xyz, const float* new_xyz, const int* offset, const int* new_offset,
                    int nbatch, int* idx, float* dist2, cudaStream_t stream);

/* block-size rule of the reference launchers (grouping_cuda_kernel.cu:29-33): 2^floor(log2 n) capped at 1024 */
int etch_opt_threads(int work_size);

/* Number of SMs the persistent (one CTA per SM) kernels size their grids for; default 148.  The pipelined runtime lowers it by
 * the batch size so the LM fit of the previous batch (one CTA per scan, another stream) keeps its SMs. */
int etch_set_sm_budget(int sms);

/* ---- fused operators behind the Python operator API (boundary B1: models.models_pointcloud / models.fit_SMPL) ---- */
/* Encoder features are point-major: feat [B, P, 60, C].  `stats` buffers are [B][C][2] doubles (sum, sum of squares),
 * zeroed by the caller, accumulated by the producer and consumed (as InstanceNorm mean / rstd) by the next kernel.   */

/* InterSO3Conv for the first conv (c_in = 1, constant occupancy feature): fused kernel-weight generation + BasicSO3Conv.
 * vgtk/so3conv/modules.py:120-128,33-39; functional.py:286-324,61-67.  kr [60,24,3] = R_a k; Wt [24][cout]. */
int etch_so3_inter_conv_c1(const float* xyz, const int* sample_idx, const int* nbr, const float* kr, const float* Wt,
                           const float* bias, int B, int q, int P, int nn, int cout, float sigma, float* zraw,
                           double* stats, cudaStream_t stream);

/* InterSO3Conv, c_in in {32,64}: same fusion plus the neighbour-feature contraction (einsum 'bcpna,bpakn->bckpa').
 * krs [60,24,4] = {2/sigma * R_a k, |R_a k|^2 / sigma}; Wt [cin*24][cout] (row c*24+k). */
int etch_so3_inter_conv(const float* xyz, const float* feat, const int* sample_idx, const int* nbr, const float* krs,
                        const float* Wt, const float* bias, int B, int q, int P, int nn, int cin, int cout, float sigma,
                        float* zraw, double* stats, cudaStream_t stream);

/* IntraSO3Conv applied to leaky_relu(InstanceNorm(zin)) (norm folded into the load).  modules.py:131-153; functional.py:331-343;
 * src/models/so3conv.py:36-44.  intra_idx [60,12] int32; Wt [12][c][cout]. */
int etch_so3_intra_conv(const float* zin, const double* in_stats, const int* intra_idx, const float* Wt, const float* bias,
                        int B, int P, int c, int cout, float* zraw, double* stats, cudaStream_t stream);

/* skip branch: Conv2d 1x1 on feats[:, :, sample_idx] (src/models/so3conv.py:178-180).  Wt [cin][cout]; ident = arange(60). */
int etch_so3_skip_conv(const float* feat, const int* sample_idx, const int* ident, const float* Wt, const float* bias, int B,
                       int q, int P, int cin, int cout, float* zraw, double* stats, cudaStream_t stream);

/* block output = leaky_relu(IN(z_intra)) + leaky_relu(IN(z_skip))  (src/models/so3conv.py:181-182); z_skip may be NULL. */
int etch_so3_combine(const float* z_intra, const double* s_intra, const float* z_skip, const double* s_skip, int B, int P,
                     int c, float* out, cudaStream_t stream);

/* PointFeatPropagation's 3-NN search + weights (src/models/pointnet2_utils.py:45-74): fine [B,N,3], coarse [B,3,S]. */
int etch_upsample3(const float* fine_bn3, const float* coarse_b3s, int B, int N, int S, int* idx, float* w,
                   cudaStream_t stream);

/* feature propagation + anchor mean + decode_direction (models_pointcloud.py:111-126,181-184; direction_backbones.py:129-223;
 * src/models/so3conv.py:186-225).  Weight tensors are the host-folded forms documented in etch_b200/models/heads.py. */
int etch_direction_head(const float* feats, const int* up_idx, const float* up_w, const float* Wqkv1, const float* Wc1,
                        const float* bc1, const float* Wqkv2, const float* Wf, const float* bf, const float* vreg,
                        float creg, const float* anchors, int B, int N, int S, float* dir, float* inv, float* anc_w,
                        cudaStream_t stream);

/* nn.Linear / Conv1d(k=1) + eval BatchNorm1d + ReLU + residual (+ per-segment row bias) in one launch
 * (pointtransformer_seg.py:40-51,71-80,101-112,144-145).  Y = relu?((X Wt + seg) * scale + shift + R). */
int etch_linear(const float* X, int ldx, const float* Wt, int n, int ci, int co, const float* scale, const float* shift,
                const float* R, const float* seg, const int* seg_off, int nseg, int relu, float* Y, int ldy,
                cudaStream_t stream);

/* PointTransformerLayer.forward after the q/k/v projections + bn2 + ReLU (pointtransformer_seg.py:24-37,118). */
int etch_pt_attention(const float* p, const float* qkv, const int* idx, const float* P0, const float* p0b, const float* P3,
                      const float* p3b, const float* s0, const float* h0, const float* W1, const float* b1, const float* W2,
                      const float* b2, const float* so, const float* ho, int n, int ns, int c, float* out,
                      cudaStream_t stream);

/* TransitionDown (stride 4) grouping + BN + ReLU + MaxPool (pointtransformer_seg.py:59-66). */
int etch_pt_down_pool(const float* p, const float* new_p, const float* Yx, const int* idx, const float* Wp, const float* scale,
                      const float* shift, int m, int ns, int co, float* out, cudaStream_t stream);

/* TransitionUp: linear1(x1) + pointops.interpolation(...) (pointtransformer_seg.py:96-97; src/models/pointops.py:164-178). */
int etch_pt_interp_add(const float* a, const float* f, const int* idx, const float* d2, int n, int c, float* out,
                       cudaStream_t stream);

/* per-scan mean of packed rows (TransitionUp head, pointtransformer_seg.py:84-92). */
int etch_seg_mean(const float* x, const int* off, int nseg, int c, float* out, cudaStream_t stream);

/* row gather out[i] = x[idx[i]] (p[idx.long(), :], pointtransformer_seg.py:60). */
int etch_gather_rows(const float* x, const int* idx, int m, int c, float* out, cudaStream_t stream);

/* confi head + softmax(cls)-weighted confidence without the [B,11008,N] intermediate (pointtransformer_seg.py:145,184-189). */
int etch_conf_head(const float* x, const float* logits, const float* W0t, const float* b0, const float* w2, const float* b2,
                   int n, int K, float* conf, cudaStream_t stream);

/* labels = argmax, vec = dir*mag/scale, inner = p - vec (src/eval.py:103,116,183; src/inference_demo.py:52-59). */
int etch_postprocess(const float* p, const float* logits, const float* dir, const float* mag, int n, int K,
                     float scale_magnitude, long long* labels, float* vec, float* inner, cudaStream_t stream);

/* get_markers (src/models/fit_SMPL.py:17-62). */
int etch_markers_top3(const float* inner, const long long* labels, const float* conf, int B, int N, int M, float* markers,
                      unsigned char* valid, cudaStream_t stream);

/* two-stage Levenberg-Marquardt marker fit (src/models/fit_SMPL.py:157-255; theseus LevenbergMarquardt semantics).
 * params [B][85] = global_orient(3) | body_pose(69) | betas(10) | transl(3); iters/errs [B][2] per stage. */
int etch_lm_fit(const float* markers, const unsigned char* valid, const float* Tm, const float* Sm, const float* Pm,
                const float* Wm, const float* Jt, const float* Js, const int* parents, const unsigned* ancmask, int B, int M,
                int steps0, int steps1, float step0, float step1, float damp0, float damp1, float* params, int* iters,
                float* errs, cudaStream_t stream);

/* SMPL.forward for fitted parameters (external/smplx/smplx/lbs.py:153-248, body_models.py:386-414). */
int etch_lbs_forward(const float* params, const float* v_template, const float* shapedirs, const float* posedirs,
                     const float* weights, const float* Jt, const float* Js, const int* parents, const int* extra_vids, int B,
                     int V, float* verts, float* joints, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------------------
 * tcgen05 / TMEM variants of the GEMM-shaped stages (the product path; the CUDA-core entry points above remain as the
 * ETCH_B200_NO_TC=1 reference build of the same operators).  All GEMMs run as 3xTF32 (hi*hi + lo*hi + hi*lo, fp32
 * accumulate in TMEM), so results stay at fp32-level accuracy against the reference's fp32 cuBLAS/einsum path.  Weight
 * operands arrive pre-split into (hi, lo) TF32 pairs and pre-tiled in the canonical K-major UMMA layout
 * (element (r, k) of an [R x K] tile at byte (k/4)*(R*16) + r*16 + (k%4)*4); etch_b200/models/tc.py builds them once per
 * checkpoint. */

/* InterSO3Conv (vgtk/so3conv/functional.py:286-324, modules.py:19-39): neighbour contraction on the CUDA cores, channel
 * mixing on tcgen05.  Wc [cin/8*2][2][24][cout][4]: slabs of 96 K-columns, K'' = kgl*48 + c*6 + i. (cin,cout,nn) in
 * {(32,32,32),(32,64,64),(64,64,32)}. */
int etch_so3_inter_conv_tc(const float* xyz, const float* feat, const int* sample_idx, const int* nbr, const float* krs,
                           const float* Wc, const float* bias, int B, int q, int P, int nn, int cin, int cout, float sigma,
                           float* zraw, double* stats, cudaStream_t stream);

/* InterSO3Conv, one point per tile (the product path; replaces etch_so3_inter_conv_tc, which stays as an A/B reference):
 * neighbour rows by 2-D tiled TMA, contraction over the neighbours on the FP32 pipes with 96 accumulators per thread,
 * accumulators parked in TMEM and fed slab by slab to tcgen05 for the channel mixing.  Wc [cin/32*16][12][2*cout][4]: per
 * (pass, channel slot cc, kernel-point half hh) the 48-column slab K'' = o*12 + i <-> W[.][(32*pass + 8*o + cc)*24 + 12*hh + i],
 * rows [W_hi; W_lo].  g4: caller-owned scratch [B,P,nn,4] fp32 (relative neighbour positions, written by a pre-pass). */
int etch_so3_inter_conv_v3(const float* xyz, const float* feat, const int* sample_idx, const int* nbr, const float* krs,
                           const float* Wc, const float* bias, int B, int q, int P, int nn, int cin, int cout, float sigma,
                           float* g4, float* zraw, double* stats, cudaStream_t stream);

/* kNN with the results of etch_knn_packed (same indices and distances, bit for bit) through a uniform grid over the candidates
 * of every segment; nsample in {3, 8, 16}; n = rows of xyz.  scratch: caller-owned, etch_knn_grid_scratch_bytes(n, nbatch) bytes. */
long long etch_knn_grid_scratch_bytes(int n, int nbatch);
int etch_knn_grid(int m, int nsample, const float* xyz, int n, const float* new_xyz, const int* offset, const int* new_offset,
                  int nbatch, int* idx, float* dist2, void* scratch, cudaStream_t stream);

/* PointTransformerLayer (pointtransformer_seg.py:8-37) + bn2 + ReLU with the first attention linear as a GEMM over
 * (point, neighbour) rows on tcgen05.  chan [c][8] = {P3 row (3), p3b, s0, h0, so, ho}; W1c [c/64][2][16][max(c/8,16)][4]
 * (TF32 hi/lo canonical tiles of the BN-folded Linear(c, c/8), one per 64-channel chunk); ns in {8, 16}. */
int etch_pt_attention_tc(const float* p, const float* qkv, const int* idx, const float* P0, const float* p0b, const float* chan,
                         const float* W1c, const float* b1, const float* W2, const float* b2, int n, int ns, int c, float* out,
                         cudaStream_t stream);

/* IntraSO3Conv (functional.py:331-343, modules.py:131-153). Wc [12][2][c/4][cout][4]. */
int etch_so3_intra_conv_tc(const float* zin, const double* in_stats, const int* intra_idx, const float* Wc, const float* bias,
                           int B, int P, int c, int cout, float* zraw, double* stats, cudaStream_t stream);

/* skip 1x1 conv of a residual block (src/models/so3conv.py:178-180). Wc [1][2][cin/4][cout][4]. */
int etch_so3_skip_conv_tc(const float* feat, const int* sample_idx, const int* ident, const float* Wc, const float* bias,
                          int B, int q, int P, int cin, int cout, float* zraw, double* stats, cudaStream_t stream);

/* decode_direction (src/models/models_pointcloud.py:102-131; src/models/so3conv.py:186-225): 3-NN blend, 2 x MHSA over the 60
 * anchor tokens, fused MLP + so3_reg, chordal SO(3) mean.  wall [18][2][8][64][4] (etch_b200/models/heads.py::DirectionPlan);
 * scratch: B*S*64 + B*N*18 floats, caller-owned. */
int etch_direction_head_tc(const float* feats, const int* up_idx, const float* up_w, const float* wall, const float* bc1,
                           const float* bf, const float* vreg, float creg, const float* anchors, int B, int N, int S,
                           float* dir, float* inv, float* anc_w, float* scratch, cudaStream_t stream);

/* confi head (pointtransformer_seg.py:145,184-189). W0c [K*4][2][8][128][4]: per marker group and K-quarter a [128 x 32] slice. */
int etch_conf_head_tc(const float* x, const float* logits, const float* W0c, const float* b0, const float* w2, const float* b2,
                      int n, int K, float* conf, cudaStream_t stream);

/* etch_linear on tcgen05, weight-stationary. Wc [NG][2][Kpad/4][NB][4], Kpad = 32*ceil(ci/32), NG = ceil(co/NB). */
int etch_linear_tc(const float* X, int ldx, const float* Wc, int NB, int n, int ci, int co, const float* scale, const float* shift,
                   const float* R, const float* seg, const int* seg_off, int nseg, int relu, float* Y, int ldy,
                   cudaStream_t stream);

/* ---- mesh -> point cloud in front of the network ------------------------------------------ */

/* preprocess_scan   src/inference_demo.py:19-34: centre[3] = (min + max) / 2 over verts [V,3] (float64, as trimesh holds them);
 * centred [V,3] = verts - centre (may be NULL; may alias verts). */
int etch_mesh_center(const double* verts, int V, double* centre, double* centred, cudaStream_t stream);

/* trimesh.sample.sample_surface   src/inference_demo.py:36-39, src/data_utils/GT_dataloader.py:102 (trimesh itself is an
 * un-vendored dependency: its published algorithm, restated in float64).  faces [F,3] int32; u_face [count] and u_len [count,2] are
 * the uniform draws in trimesh's order; scratch [2*F] doubles.  Outputs (each optional): out64 / out32 [count,3], face_index [count]. */
int etch_mesh_sample(const double* verts, const int* faces, int V, int F, const double* u_face, const double* u_len, int count,
                     double* scratch, double* out64, float* out32, int* face_index, cudaStream_t stream);

/* ---- ground-truth tightness vectors of the evaluation dataset (the "VECTORS" block) --------------- */

/* trimesh.proximity.closest_point(smpl_mesh, sample_points)   src/data_utils/GT_dataloader.py:110 (trimesh un-vendored: brute force
 * over all faces with Ericson's point-triangle projection, float64).  scratch: chunks*n doubles + chunks*n ints with
 * chunks = etch_mesh_closest_chunks(F).  Outputs: closest [n,3]; dist [n], face [n] optional. */
int etch_mesh_closest_chunks(int F);
int etch_mesh_closest_point(const double* verts, const int* faces, int V, int F, const double* pts, int n, void* scratch,
                            double* closest, double* dist, int* face, cudaStream_t stream);

/* scipy cKDTree(ref).query(pts, k=1)   GT_dataloader.py:106-107: exact nearest neighbour, float64; ties -> lowest index. */
int etch_nearest_point(const double* ref, int m, const double* pts, int n, double* dist, int* index, cudaStream_t stream);

/* GT_dataloader.py:112-124: vectors = info_vectors[nn_idx] where nn_dist < threshold, else pts - closest. */
int etch_gt_vectors(const double* pts, const double* closest, const double* info_vectors, const double* nn_dist, const int* nn_idx,
                    int n, double threshold, double* vectors, cudaStream_t stream);

/* ---- evaluation-driver output writers: HOST pointers, no stream ------------------------------ */

/* utils.GT_utils.save_points_with_vector   src/utils/GT_utils.py:22-55 (src/eval.py:147-149): ASCII PLY of n points (red), the n
 * vector end points hit - vec (blue) and n edges; byte-identical to the reference's Python loop.  style = text form of a float32
 * inside an f-string: 0 = numpy >= 2 (repr of the widened double), 1 = numpy 1.x (shortest float32 form). */
int etch_write_points_vector_ply(const char* path, const float* hit_points, const float* vectors, int n, int style);

/* numpy.float32.__str__ for n values, newline separated (formatter self-test used by the CPU suite); returns bytes written, -1 = cap */
long long etch_format_np_float32(const float* x, int n, int style, char* out, long long cap);

#ifdef __cplusplus
}
#endif
#endif /* ETCH_B200_H */
