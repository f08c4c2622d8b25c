// SMPL marker fit: marker extraction, two-stage Levenberg-Marquardt with an ANALYTIC marker-only Jacobian, full-mesh LBS.
//
// Reference semantics restated:
//   src/models/fit_SMPL.py:17-62      get_markers (per label: top-3 confidences, weights conf^20, weighted centroid)
//   src/models/fit_SMPL.py:111-152    marker residual  r = mask * (pred_marker - V[marker_vid])
//   src/models/fit_SMPL.py:157-255    2 x theseus LevenbergMarquardt (30 its step .5 damping .01; 50 its step .2 damping 1e-3)
//   external/smplx/smplx/lbs.py:153-398  lbs / blend_shapes / vertices2joints / batch_rodrigues (+1e-8) / batch_rigid_transform
//   external/smplx/smplx/body_models.py:386-414, vertex_joint_selector.py:29-80   joints = 24 chain + 21 picked vertices
//
// B200 design: the reference differentiates 258 residuals through a full 6890-vertex LBS with autograd (~2 GFLOP per
// iteration and hundreds of launches).  Only 86 vertices matter, so one persistent CTA per scan keeps the whole solve in
// shared memory: forward kinematics, the 86 marker vertices, the closed-form Jacobian (258 x 85), J^T J + lambda I, a
// dense Cholesky and the update -- 80 iterations without leaving the SM; marker rows of the blend-shape bases are the
// only global reads (L2-resident, shared by all scans).  The full mesh is skinned once at the end.
#include "common.cuh"

namespace {

constexpr int NJ = 24;
constexpr int NM_MAX = 96;     // markers supported by the shared-memory layout
constexpr int DMAX = 85;       // 72 pose + 10 betas + 3 transl
constexpr int LDJ = 88;        // padded leading dimension of J
constexpr int LDA = 89;        // padded leading dimension of A = J^T J
constexpr int TS = 5;          // tile size of the register-resident blocked Cholesky (85 = 17 x 5)

// ------------------------------------------------------------------------------------------------ markers
// one warp per (scan, label): top-3 confidences among the points carrying that label (ties: lower point index first)
__global__ void __launch_bounds__(256) markers_kernel(const float* __restrict__ inner, const long long* __restrict__ labels,
                                                      const float* __restrict__ conf, int N, int M,
                                                      float* __restrict__ markers, unsigned char* __restrict__ valid) {
    const int b = blockIdx.y;
    const int label = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (label >= M) return;
    float c[3] = {-INFINITY, -INFINITY, -INFINITY};
    int id[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff};
    int cnt = 0;
    const long long* L = labels + (size_t)b * N;
    const float* C = conf + (size_t)b * N;
    for (int i = lane; i < N; i += 32) {
        if (L[i] == label) {
            ++cnt;
            const float v = C[i];
            // insert (v, i): larger v first, then smaller i
            if (v > c[2] || (v == c[2] && i < id[2])) {
                c[2] = v; id[2] = i;
                if (c[2] > c[1] || (c[2] == c[1] && id[2] < id[1])) { float t = c[1]; c[1] = c[2]; c[2] = t; int u = id[1]; id[1] = id[2]; id[2] = u; }
                if (c[1] > c[0] || (c[1] == c[0] && id[1] < id[0])) { float t = c[0]; c[0] = c[1]; c[1] = t; int u = id[0]; id[0] = id[1]; id[1] = u; }
            }
        }
    }
    // merge the 32 sorted triples: three rounds of warp arg-max extraction
    int total = cnt;
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    float tc[3]; int ti[3];
    int head = 0;
    for (int r = 0; r < 3; ++r) {
        float v = head < 3 ? c[head] : -INFINITY;
        int iv = head < 3 ? id[head] : 0x7fffffff;
        float bv = v; int bi = iv;
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        tc[r] = bv; ti[r] = bi;
        if (head < 3 && iv == bi && bi != 0x7fffffff) ++head;
    }
    if (lane == 0) {
        float* out = markers + ((size_t)b * M + label) * 3;
        if (total == 0) {
            out[0] = out[1] = out[2] = 0.f;
            valid[(size_t)b * M + label] = 0;
        } else {
            const int k = total < 3 ? total : 3;
            float ws = 0.f, sx = 0.f, sy = 0.f, sz = 0.f;
            for (int r = 0; r < k; ++r) {
                // fewer insertable candidates than labelled points: some confidence is NaN (never inserted above).  torch.topk
                // ranks NaN first, so the reference's marker is NaN here (fit_SMPL.py:38-51): emit NaN, never read out of bounds
                if (ti[r] == 0x7fffffff) { sx = sy = sz = ws = __int_as_float(0x7fc00000); break; }
                const float w = powf(tc[r], 20.0f);
                const float* p = inner + ((size_t)b * N + ti[r]) * 3;
                sx += p[0] * w; sy += p[1] * w; sz += p[2] * w; ws += w;
            }
            out[0] = sx / ws; out[1] = sy / ws; out[2] = sz / ws;
            valid[(size_t)b * M + label] = 1;
        }
    }
}

// ------------------------------------------------------------------------------------------------ body constants
struct BodyMarkers {           // marker-restricted SMPL tables (device pointers)
    const float* Tm;           // [M][3]      v_template[vid]
    const float* Sm;           // [M][3][10]  shapedirs[vid]
    const float* Pm;           // [207][M*3]  posedirs[:, vid*3+c]
    const float* Wm;           // [M][24]     lbs_weights[vid]
    const float* Jt;           // [24][3]     J_regressor v_template
    const float* Js;           // [24][3][10] J_regressor shapedirs
    const int* parents;        // [24]
    const unsigned* ancmask;   // [24]  bit j set iff joint j is an ancestor-or-self of the joint
    int M;
};

// Rodrigues exactly as lbs.batch_rodrigues (:295-330): angle = |r + 1e-8|, K = [r/angle]x, R = I + sin K + (1-cos) K^2,
// plus its analytic derivative wrt each component of r (valid through r = 0 thanks to the reference's own +1e-8).
__device__ void rodrigues_with_grad(const float* r, float* R, float* dR /*[3][9]*/) {
    const double rx = r[0], ry = r[1], rz = r[2];
    const double ax = rx + 1e-8, ay = ry + 1e-8, az = rz + 1e-8;
    const double th = sqrt(ax * ax + ay * ay + az * az);
    const double s = sin(th), c = cos(th), omc = 1.0 - c;
    const double rho[3] = {rx / th, ry / th, rz / th};
    double K[9] = {0, -rho[2], rho[1], rho[2], 0, -rho[0], -rho[1], rho[0], 0};
    double K2[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) K2[i * 3 + j] = K[i * 3] * K[j] + K[i * 3 + 1] * K[3 + j] + K[i * 3 + 2] * K[6 + j];
    for (int e = 0; e < 9; ++e) R[e] = (float)((e % 4 == 0 ? 1.0 : 0.0) + s * K[e] + omc * K2[e]);
    const double a3[3] = {ax, ay, az}, r3[3] = {rx, ry, rz};
    for (int i = 0; i < 3; ++i) {
        const double thi = a3[i] / th;  // d theta / d r_i
        double drho[3];
        for (int k = 0; k < 3; ++k) drho[k] = (k == i ? 1.0 / th : 0.0) - r3[k] * thi / (th * th);
        const double dK[9] = {0, -drho[2], drho[1], drho[2], 0, -drho[0], -drho[1], drho[0], 0};
        for (int a = 0; a < 3; ++a)
            for (int bb = 0; bb < 3; ++bb) {
                double dKK = 0.0;  // dK K + K dK
                for (int k = 0; k < 3; ++k) dKK += dK[a * 3 + k] * K[k * 3 + bb] + K[a * 3 + k] * dK[k * 3 + bb];
                dR[i * 9 + a * 3 + bb] = (float)(c * thi * K[a * 3 + bb] + s * dK[a * 3 + bb] + s * thi * K2[a * 3 + bb] + omc * dKK);
            }
    }
}

struct LmSmem {
    float J[3 * NM_MAX * LDJ];        // residual Jacobian (rows = 3M); during lm_eval it holds the raw pose-blend terms q
    float A[LDJ * LDA];               // [J^T J + lambda I ; -J^T r] (row D = right-hand side) -> Cholesky factor / y
    float cmk[NM_MAX * NJ * 3];       // marker transformed by bone k alone
    float S[NM_MAX * NJ * 3];         // sum over descendants of j of w (c_mk - t_j)
    float Tv[NM_MAX * 9];             // blended rotation per marker
    float vp[NM_MAX * 3];             // posed-template marker (before skinning)
    float vpb[NM_MAX * 3];            // pose-blend offsets
    float res[NM_MAX * 3];            // residual
    float tgt[NM_MAX * 3];
    float mask[NM_MAX];
    float R[NJ * 9], dR[NJ * 27], G[NJ * 12], ta[NJ * 3], Jr[NJ * 3], Om[NJ * 27];
    float dT[NJ * 30];                // d(posed joint k)/d beta_l
    float pf[207];
    float x[DMAX + 3];                // theta[72] | beta[10] | transl[3]
    float g[LDJ];                     // solution delta
    float red[32];
    float Ld[TS * TS];                // current diagonal tile (un-factored) of the blocked Cholesky
    float Lp[(DMAX / TS + 2) * TS * TS];  // factored column panel: one tile per block row (+ the rhs row)
    float Wm[NM_MAX * NJ];            // skinning weights of the marker vertices, COMPACTED: the nzc[m] non-zero ones first
    unsigned char nzk[NM_MAX * NJ];   // bone index of the i-th non-zero weight of marker m
    int nzc[NM_MAX];                  // number of non-zero weights of marker m
    float dinv[LDJ];                  // reciprocal diagonal of the Cholesky factor
    unsigned anc[NJ];
    float err;
};

// forward pass at the current parameters: kinematics, marker vertices, residual, error.  The single sweep over the
// marker rows of posedirs also leaves q[o][j*3+i] = sum_e dR_{j,i}[e] * P[(j-1)*9+e][o] in S.J for lm_jacobian.
__device__ void lm_eval(LmSmem& S, const BodyMarkers& Bm) {
    const int tid = threadIdx.x, M = Bm.M, M3 = Bm.M * 3;
    float* theta = S.x; float* beta = S.x + 72; float* transl = S.x + 82;
    if (tid < NJ) {
        rodrigues_with_grad(theta + tid * 3, S.R + tid * 9, S.dR + tid * 27);
        if (tid >= 1)
            for (int e = 0; e < 9; ++e) S.pf[(tid - 1) * 9 + e] = S.R[tid * 9 + e] - (e % 4 == 0 ? 1.f : 0.f);
    } else if (tid >= 32 && tid < 32 + NJ * 3) {
        const int o = tid - 32;
        float v = __ldg(Bm.Jt + o);
        for (int l = 0; l < 10; ++l) v = fmaf(__ldg(Bm.Js + o * 10 + l), beta[l], v);
        S.Jr[o] = v;
    }
    for (int o = tid; o < M3; o += blockDim.x) S.vpb[o] = 0.f;
    __syncthreads();
    if (tid < 32) {  // kinematic chain, one warp, 12 lanes = entries of [R|t]
        const int r = tid / 4, c = tid % 4;
        for (int j = 0; j < NJ; ++j) {
            if (tid < 12) {
                const int pa = __ldg(Bm.parents + j);
                if (j == 0) {
                    S.G[r * 4 + c] = c < 3 ? S.R[r * 3 + c] : S.Jr[r];
                } else {
                    const float* Gp = S.G + pa * 12;
                    float v = c == 3 ? Gp[r * 4 + 3] : 0.f;
                    for (int k = 0; k < 3; ++k) {
                        const float l = c < 3 ? S.R[j * 9 + k * 3 + c] : (S.Jr[j * 3 + k] - S.Jr[pa * 3 + k]);
                        v = fmaf(Gp[r * 4 + k], l, v);
                    }
                    S.G[j * 12 + r * 4 + c] = v;
                }
            }
            __syncwarp();
        }
    }
    // fused pose-blend sweep: task = (output o, group of joints); each posedirs value feeds the blend and 3 Jacobian terms
    {
        const int nthr = blockDim.x - 32;
        for (int t = tid - 32; t >= 0 && t < M3 * 4; t += nthr) {
            const int o = t % M3, jg = t / M3;
            const int j0 = 1 + jg * 6, j1 = min(NJ, j0 + 6);
            float acc = 0.f;
            for (int j = j0; j < j1; ++j) {
                float pv[9];
#pragma unroll
                for (int e = 0; e < 9; ++e) pv[e] = __ldg(Bm.Pm + (size_t)((j - 1) * 9 + e) * M3 + o);
                float q0 = 0.f, q1 = 0.f, q2 = 0.f;
#pragma unroll
                for (int e = 0; e < 9; ++e) {
                    acc = fmaf(S.pf[(j - 1) * 9 + e], pv[e], acc);
                    q0 = fmaf(S.dR[j * 27 + e], pv[e], q0);
                    q1 = fmaf(S.dR[j * 27 + 9 + e], pv[e], q1);
                    q2 = fmaf(S.dR[j * 27 + 18 + e], pv[e], q2);
                }
                S.J[o * LDJ + j * 3] = q0; S.J[o * LDJ + j * 3 + 1] = q1; S.J[o * LDJ + j * 3 + 2] = q2;
            }
            atomicAdd(&S.vpb[o], acc);
            if (jg == 0) { S.J[o * LDJ] = 0.f; S.J[o * LDJ + 1] = 0.f; S.J[o * LDJ + 2] = 0.f; }
        }
    }
    __syncthreads();
    if (tid < NJ * 3) {  // translation of the relative transform A_j = G_j - [0 | G_j J_j]
        const int j = tid / 3, c = tid % 3;
        const float* G = S.G + j * 12;
        S.ta[tid] = G[c * 4 + 3] - (G[c * 4] * S.Jr[j * 3] + G[c * 4 + 1] * S.Jr[j * 3 + 1] + G[c * 4 + 2] * S.Jr[j * 3 + 2]);
    }
    for (int o = tid; o < M3; o += blockDim.x) {
        float v = __ldg(Bm.Tm + o);
        for (int l = 0; l < 10; ++l) v = fmaf(__ldg(Bm.Sm + o * 10 + l), beta[l], v);
        S.vp[o] = v + S.vpb[o];
    }
    __syncthreads();
    for (int t = tid; t < M * NJ; t += blockDim.x) {   // entry i of marker m = its i-th non-zero bone
        const int m = t / NJ, i = t % NJ;
        if (i < S.nzc[m]) {
            const int k = S.nzk[t];
            const float* G = S.G + k * 12;
            const float x = S.vp[m * 3], y = S.vp[m * 3 + 1], z = S.vp[m * 3 + 2];
            for (int c = 0; c < 3; ++c) S.cmk[t * 3 + c] = fmaf(G[c * 4], x, fmaf(G[c * 4 + 1], y, fmaf(G[c * 4 + 2], z, S.ta[k * 3 + c])));
        }
    }
    for (int t = tid; t < M * 9; t += blockDim.x) {
        const int m = t / 9, e = t % 9;
        float v = 0.f;
        for (int i = 0; i < S.nzc[m]; ++i) v = fmaf(S.Wm[m * NJ + i], S.G[S.nzk[m * NJ + i] * 12 + (e / 3) * 4 + (e % 3)], v);
        S.Tv[t] = v;
    }
    __syncthreads();
    float e2 = 0.f;
    for (int o = tid; o < M3; o += blockDim.x) {
        const int m = o / 3, c = o % 3;
        float v = transl[c];
        for (int i = 0; i < S.nzc[m]; ++i) v = fmaf(S.Wm[m * NJ + i], S.cmk[(m * NJ + i) * 3 + c], v);
        const float r = S.mask[m] * (S.tgt[o] - v);
        S.res[o] = r;
        e2 = fmaf(r, r, e2);
    }
    for (int o = 16; o > 0; o >>= 1) e2 += __shfl_xor_sync(0xffffffffu, e2, o);
    if ((tid & 31) == 0) S.red[tid >> 5] = e2;
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += S.red[w];
        S.err = 0.5f * s;
    }
    __syncthreads();
}

// analytic Jacobian of the residual at the state left by lm_eval; columns: theta(72) | beta(nb) | transl(3)
__device__ void lm_jacobian(LmSmem& S, const BodyMarkers& Bm, int nb) {
    const int tid = threadIdx.x, M = Bm.M;
    // Omega_{j,i} = Rg_parent(j) dR_{j,i} Rg_j^T
    for (int t = tid; t < NJ * 3; t += blockDim.x) {
        const int j = t / 3;
        const int pa = __ldg(Bm.parents + j);
        float tmp[9];
        const float* dR = S.dR + t * 9;
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                float v = 0.f;
                for (int k = 0; k < 3; ++k) {
                    const float gp = j == 0 ? (a == k ? 1.f : 0.f) : S.G[pa * 12 + a * 4 + k];
                    v = fmaf(gp, dR[k * 3 + b], v);
                }
                tmp[a * 3 + b] = v;
            }
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                float v = 0.f;
                for (int k = 0; k < 3; ++k) v = fmaf(tmp[a * 3 + k], S.G[j * 12 + b * 4 + k], v);
                S.Om[t * 9 + a * 3 + b] = v;
            }
    }
    // d(posed joint k)/d beta_l : chain over k for each (l, c)
    for (int t = tid - 96; t >= 0 && t < nb * 3; t += blockDim.x) {
        const int l = t / 3, c = t % 3;
        for (int k = 0; k < NJ; ++k) {
            const int pa = __ldg(Bm.parents + k);
            float v;
            if (k == 0) v = __ldg(Bm.Js + (0 * 3 + c) * 10 + l);
            else {
                v = S.dT[pa * 30 + l * 3 + c];
                for (int q = 0; q < 3; ++q)
                    v = fmaf(S.G[pa * 12 + c * 4 + q], __ldg(Bm.Js + (k * 3 + q) * 10 + l) - __ldg(Bm.Js + (pa * 3 + q) * 10 + l), v);
            }
            S.dT[k * 30 + l * 3 + c] = v;
        }
    }
    // S[m][j] = sum_{k in desc(j)} w_mk (c_mk - t_j)
    for (int t = tid; t < M * NJ; t += blockDim.x) {
        const int m = t / NJ, j = t % NJ;
        float sx = 0.f, sy = 0.f, sz = 0.f;
        const float tx = S.G[j * 12 + 3], ty = S.G[j * 12 + 7], tz = S.G[j * 12 + 11];
        for (int i = 0; i < S.nzc[m]; ++i) {
            const int k = S.nzk[m * NJ + i];
            if ((S.anc[k] >> j) & 1u) {
                const float w = S.Wm[m * NJ + i];
                const float* c = S.cmk + (m * NJ + i) * 3;
                sx = fmaf(w, c[0] - tx, sx); sy = fmaf(w, c[1] - ty, sy); sz = fmaf(w, c[2] - tz, sz);
            }
        }
        S.S[t * 3] = sx; S.S[t * 3 + 1] = sy; S.S[t * 3 + 2] = sz;
    }
    __syncthreads();
    // theta columns (in place: the raw q terms are replaced by the Jacobian entries)
    for (int t = tid; t < M * NJ * 3; t += blockDim.x) {
        const int m = t / (NJ * 3), ji = t % (NJ * 3), j = ji / 3;
        const float mk = S.mask[m];
        float d[3] = {0.f, 0.f, 0.f};
        if (mk != 0.f) {
            const float* Om = S.Om + ji * 9;
            const float* sv = S.S + (m * NJ + j) * 3;
            for (int c = 0; c < 3; ++c) d[c] = Om[c * 3] * sv[0] + Om[c * 3 + 1] * sv[1] + Om[c * 3 + 2] * sv[2];
            if (j >= 1) {
                const float q0 = S.J[(m * 3) * LDJ + ji], q1 = S.J[(m * 3 + 1) * LDJ + ji], q2 = S.J[(m * 3 + 2) * LDJ + ji];
                const float* Tv = S.Tv + m * 9;
                for (int c = 0; c < 3; ++c) d[c] += Tv[c * 3] * q0 + Tv[c * 3 + 1] * q1 + Tv[c * 3 + 2] * q2;
            }
        }
        for (int c = 0; c < 3; ++c) S.J[(m * 3 + c) * LDJ + ji] = -mk * d[c];
    }
    // beta columns
    for (int t = tid; t < M * nb; t += blockDim.x) {
        const int m = t / nb, l = t % nb;
        const float mk = S.mask[m];
        float d[3] = {0.f, 0.f, 0.f};
        if (mk != 0.f) {
            const float s0 = __ldg(Bm.Sm + (m * 3 + 0) * 10 + l), s1 = __ldg(Bm.Sm + (m * 3 + 1) * 10 + l), s2 = __ldg(Bm.Sm + (m * 3 + 2) * 10 + l);
            for (int i = 0; i < S.nzc[m]; ++i) {
                const float w = S.Wm[m * NJ + i];
                const int k = S.nzk[m * NJ + i];
                const float a0 = s0 - __ldg(Bm.Js + (k * 3 + 0) * 10 + l), a1 = s1 - __ldg(Bm.Js + (k * 3 + 1) * 10 + l),
                            a2 = s2 - __ldg(Bm.Js + (k * 3 + 2) * 10 + l);
                const float* G = S.G + k * 12;
                for (int c = 0; c < 3; ++c)
                    d[c] = fmaf(w, G[c * 4] * a0 + G[c * 4 + 1] * a1 + G[c * 4 + 2] * a2 + S.dT[k * 30 + l * 3 + c], d[c]);
            }
        }
        for (int c = 0; c < 3; ++c) S.J[(m * 3 + c) * LDJ + 72 + l] = -mk * d[c];
    }
    // translation columns
    for (int t = tid; t < M * 9; t += blockDim.x) {
        const int m = t / 9, c = (t % 9) / 3, c2 = t % 3;
        S.J[(m * 3 + c) * LDJ + 72 + nb + c2] = c == c2 ? -S.mask[m] : 0.f;
    }
    __syncthreads();
}

// delta = (J^T J + lambda I)^-1 (-J^T r).  The normal matrix never touches shared memory un-factored: thread t owns one
// 5x5 tile of the lower triangle (plus 17 threads owning the 1x5 tiles of the right-hand side, appended as an extra block
// row so that y = L^-1 b falls out of the factorisation), accumulates it from J in registers, and the right-looking
// blocked Cholesky runs on those registers with two barriers per block column; a single warp then does the blocked
// back-substitution L^T x = y.
__device__ void lm_solve(LmSmem& S, int rows, int D, float lambda, long long* tprof) {
    const int tid = threadIdx.x;
    const long long tc0 = clock64();
    const int NT = (D + TS - 1) / TS;
    const int NTL = NT * (NT + 1) / 2;
    int ta = -1, tb = -1;      // block row / column of my tile; ta == NT marks a right-hand-side tile
    if (tid < NTL) {
        ta = (int)((sqrtf(8.f * tid + 1.f) - 1.f) * 0.5f);
        while ((ta + 1) * (ta + 2) / 2 <= tid) ++ta;
        while (ta * (ta + 1) / 2 > tid) --ta;
        tb = tid - ta * (ta + 1) / 2;
    } else if (tid < NTL + NT) {
        ta = NT; tb = tid - NTL;
    }
    float acc[TS][TS];
#pragma unroll
    for (int i = 0; i < TS; ++i)
#pragma unroll
        for (int j = 0; j < TS; ++j) acc[i][j] = 0.f;
    if (ta >= 0 && ta < NT) {
        for (int r = 0; r < rows; ++r) {
            const float* row = S.J + r * LDJ;
            float av[TS], bv[TS];
#pragma unroll
            for (int i = 0; i < TS; ++i) { av[i] = ta * TS + i < D ? row[ta * TS + i] : 0.f; bv[i] = tb * TS + i < D ? row[tb * TS + i] : 0.f; }
#pragma unroll
            for (int i = 0; i < TS; ++i)
#pragma unroll
                for (int j = 0; j < TS; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (ta == tb) {
#pragma unroll
            for (int i = 0; i < TS; ++i) acc[i][i] = ta * TS + i < D ? acc[i][i] + lambda : 1.0f;  // pad rows: identity
        }
    } else if (ta == NT) {
        for (int r = 0; r < rows; ++r) {
            const float rr = S.res[r];
            const float* row = S.J + r * LDJ;
#pragma unroll
            for (int j = 0; j < TS; ++j) acc[0][j] = fmaf(tb * TS + j < D ? row[tb * TS + j] : 0.f, -rr, acc[0][j]);
        }
    }
    __syncthreads();
    const long long tc1 = clock64();
    for (int kb = 0; kb < NT; ++kb) {
        if (ta == kb && tb == kb) {
#pragma unroll
            for (int i = 0; i < TS; ++i)
#pragma unroll
                for (int j = 0; j < TS; ++j) S.Ld[i * TS + j] = acc[i][j];
        }
        __syncthreads();
        if (tb == kb && ta >= kb) {
            float Lk[TS][TS];
#pragma unroll
            for (int i = 0; i < TS; ++i)
#pragma unroll
                for (int j = 0; j < TS; ++j) Lk[i][j] = j <= i ? S.Ld[i * TS + j] : 0.f;
            float dinv[TS];
#pragma unroll
            for (int c = 0; c < TS; ++c) {   // 5x5 Cholesky of the diagonal tile, redundantly in every panel thread
                float d = Lk[c][c];
#pragma unroll
                for (int e = 0; e < TS; ++e) if (e < c) d = fmaf(-Lk[c][e], Lk[c][e], d);
                const float inv = rsqrtf(d);
                dinv[c] = inv;
                Lk[c][c] = d * inv;
#pragma unroll
                for (int r = 0; r < TS; ++r) {
                    if (r > c) {
                        float v = Lk[r][c];
#pragma unroll
                        for (int e = 0; e < TS; ++e) if (e < c) v = fmaf(-Lk[r][e], Lk[c][e], v);
                        Lk[r][c] = v * inv;
                    }
                }
            }
            if (ta == kb) {
#pragma unroll
                for (int i = 0; i < TS; ++i) {
#pragma unroll
                    for (int j = 0; j < TS; ++j) acc[i][j] = Lk[i][j];
                    if (kb * TS + i < D) S.dinv[kb * TS + i] = dinv[i];
                }
            } else {   // X L_kk^T = A_ik  (row-wise forward substitution); rhs tiles only use row 0
#pragma unroll
                for (int i = 0; i < TS; ++i) {
#pragma unroll
                    for (int c = 0; c < TS; ++c) {
                        float v = acc[i][c];
#pragma unroll
                        for (int e = 0; e < TS; ++e) if (e < c) v = fmaf(-acc[i][e], Lk[c][e], v);
                        acc[i][c] = v * dinv[c];
                    }
                }
            }
            float* lp = S.Lp + ta * TS * TS;
#pragma unroll
            for (int i = 0; i < TS; ++i)
#pragma unroll
                for (int j = 0; j < TS; ++j) {
                    lp[i * TS + j] = acc[i][j];
                    const int gi = ta * TS + i, gj = kb * TS + j;
                    if (ta < NT) { if (gi < D && gj < D) S.A[gi * LDA + gj] = acc[i][j]; }
                    else if (i == 0 && gj < D) S.A[D * LDA + gj] = acc[0][j];
                }
        }
        __syncthreads();
        if (tb > kb && ta >= tb) {   // trailing update A_ij -= L_ik L_jk^T
            const float* li = S.Lp + ta * TS * TS;
            const float* lj = S.Lp + tb * TS * TS;
            float Lj[TS][TS];
#pragma unroll
            for (int j = 0; j < TS; ++j)
#pragma unroll
                for (int c = 0; c < TS; ++c) Lj[j][c] = lj[j * TS + c];
            const int ni = ta == NT ? 1 : TS;
#pragma unroll
            for (int i = 0; i < TS; ++i) {
                if (i < ni) {
                    float Li[TS];
#pragma unroll
                    for (int c = 0; c < TS; ++c) Li[c] = li[i * TS + c];
#pragma unroll
                    for (int j = 0; j < TS; ++j)
#pragma unroll
                        for (int c = 0; c < TS; ++c) acc[i][j] = fmaf(-Li[c], Lj[j][c], acc[i][j]);
                }
            }
        }
    }
    __syncthreads();
    const long long tc2 = clock64();
    // blocked back-substitution L^T x = y (y = row D of A), one warp; lane l accumulates the tiles of block rows kb+1+l, ...
    if (tid < 32) {
        for (int i = tid; i < NT * TS; i += 32) S.g[i] = i < D ? S.A[D * LDA + i] : 0.f;
        __syncwarp();
        for (int kb = NT - 1; kb >= 0; --kb) {
            float s[TS];
#pragma unroll
            for (int c = 0; c < TS; ++c) s[c] = 0.f;
            for (int ib = kb + 1 + tid; ib < NT; ib += 32) {
#pragma unroll
                for (int r = 0; r < TS; ++r) {
                    const int gi = ib * TS + r;
                    if (gi < D) {
                        const float xv = S.g[gi];
#pragma unroll
                        for (int c = 0; c < TS; ++c) { const int gc = kb * TS + c; if (gc < D) s[c] = fmaf(S.A[gi * LDA + gc], xv, s[c]); }
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < TS; ++c)
                for (int o = 16; o > 0; o >>= 1) s[c] += __shfl_xor_sync(0xffffffffu, s[c], o);
            if (tid == 0) {
                float xk[TS];
#pragma unroll
                for (int c = TS - 1; c >= 0; --c) {
                    const int gc = kb * TS + c;
                    if (gc < D) {
                        float v = S.g[gc] - s[c];
#pragma unroll
                        for (int e = TS - 1; e > c; --e) { const int ge = kb * TS + e; if (ge < D) v = fmaf(-S.A[ge * LDA + gc], xk[e], v); }
                        xk[c] = v * S.dinv[gc];
                        S.g[gc] = xk[c];
                    } else xk[c] = 0.f;
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (tprof && tid == 0) { tprof[0] += tc1 - tc0; tprof[1] += tc2 - tc1; tprof[2] += clock64() - tc2; }
}

__global__ void __launch_bounds__(256, 1) lm_fit_kernel(const float* __restrict__ markers, const unsigned char* __restrict__ valid,
                                                        BodyMarkers Bm, int it0, int it1, float step0, float step1,
                                                        float damp0, float damp1, float* __restrict__ params,
                                                        int* __restrict__ iters, float* __restrict__ errs,
                                                        long long* __restrict__ prof) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LmSmem& S = *reinterpret_cast<LmSmem*>(smem_raw);
    const int b = blockIdx.x, tid = threadIdx.x, M = Bm.M;
    for (int i = tid; i < DMAX + 3; i += blockDim.x) S.x[i] = 0.f;
    for (int i = tid; i < M * 3; i += blockDim.x) S.tgt[i] = __ldg(markers + (size_t)b * M * 3 + i);
    for (int i = tid; i < M; i += blockDim.x) S.mask[i] = valid[(size_t)b * M + i] ? 1.f : 0.f;
    for (int m = tid; m < M; m += blockDim.x) {   // compact the (sparse: <= 4 non-zeros for SMPL) skinning weights
        int cnt = 0;
        for (int k = 0; k < NJ; ++k) {
            const float w = __ldg(Bm.Wm + m * NJ + k);
            if (w != 0.f) { S.Wm[m * NJ + cnt] = w; S.nzk[m * NJ + cnt] = (unsigned char)k; ++cnt; }
        }
        S.nzc[m] = cnt;
    }
    for (int i = tid; i < NJ; i += blockDim.x) S.anc[i] = __ldg(Bm.ancmask + i);
    __syncthreads();
    for (int stage = 0; stage < 2; ++stage) {
        const int nb = stage == 0 ? 2 : 10;
        const int D = 72 + nb + 3;
        const int maxit = stage == 0 ? it0 : it1;
        const float step = stage == 0 ? step0 : step1, lambda = stage == 0 ? damp0 : damp1;
        lm_eval(S, Bm);
        float last = S.err;
        int done = 0;
        long long t_eval = 0, t_jac = 0, t_solve = 0;
        __shared__ long long tsolve[3];
        if (tid == 0) { tsolve[0] = tsolve[1] = tsolve[2] = 0; }
        for (int it = 0; it < maxit; ++it) {
            long long c0 = clock64();
            lm_jacobian(S, Bm, nb);
            long long c1 = clock64();
            lm_solve(S, M * 3, D, lambda, prof ? tsolve : nullptr);
            long long c2 = clock64();
            t_jac += c1 - c0; t_solve += c2 - c1;
            // x <- x + step * delta   (columns: theta | beta[0..nb) | transl)
            for (int a = tid; a < D; a += blockDim.x) {
                const int xi = a < 72 ? a : (a < 72 + nb ? a : 82 + (a - 72 - nb));
                S.x[xi] = fmaf(step, S.g[a], S.x[xi]);
            }
            __syncthreads();
            c0 = clock64();
            lm_eval(S, Bm);
            t_eval += clock64() - c0;
            const float err = S.err;
            done = it + 1;
            const float ae = fabsf(last - err);
            const bool conv = (fabsf(err) < 1e-10f) || (ae < 1e-10f) || (ae / last < 1e-8f);
            if (conv) break;
            last = err;
        }
        if (tid == 0) {
            iters[b * 2 + stage] = done; errs[b * 2 + stage] = S.err;
            if (prof) { prof[(b * 2 + stage) * 6] = t_eval; prof[(b * 2 + stage) * 6 + 1] = t_jac; prof[(b * 2 + stage) * 6 + 2] = t_solve;
                        prof[(b * 2 + stage) * 6 + 3] = tsolve[0]; prof[(b * 2 + stage) * 6 + 4] = tsolve[1]; prof[(b * 2 + stage) * 6 + 5] = tsolve[2]; }
        }
        __syncthreads();
    }
    for (int i = tid; i < DMAX; i += blockDim.x) params[(size_t)b * DMAX + i] = S.x[i];
}

// ------------------------------------------------------------------------------------------------ full-mesh LBS
struct BodyFull {
    const float* v_template;  // [V][3]
    const float* shapedirs;   // [V][3][10]
    const float* posedirs;    // [207][V*3]
    const float* weights;     // [V][24]
    const float* Jt;          // [24][3]
    const float* Js;          // [24][3][10]
    const int* parents;       // [24]
    const int* extra_vids;    // [21]
    int V;
};

// params [B][85] = theta(72: orient, pose) | beta(10) | transl(3)  ->  verts [B][V][3], joints [B][45][3]
// Test hook (include/etch_b200_probes.h): residual and analytic Jacobian of the marker cost at GIVEN parameters, written out so that
// tests/test_fit_gpu.py can compare them with the autodiff Jacobian of the oracle.  One CTA per scan;
// res [B][3M], jac [B][3M][85] with columns theta(72: orient | pose) | beta(10) | transl(3).
__global__ void __launch_bounds__(256, 1) lm_jacobian_dump_kernel(const float* __restrict__ params, const float* __restrict__ markers,
                                                                  const unsigned char* __restrict__ valid, BodyMarkers Bm,
                                                                  float* __restrict__ res, float* __restrict__ jac) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LmSmem& S = *reinterpret_cast<LmSmem*>(smem_raw);
    const int b = blockIdx.x, tid = threadIdx.x, M = Bm.M;
    for (int i = tid; i < DMAX + 3; i += blockDim.x) S.x[i] = i < DMAX ? __ldg(params + (size_t)b * DMAX + i) : 0.f;
    for (int i = tid; i < M * 3; i += blockDim.x) S.tgt[i] = __ldg(markers + (size_t)b * M * 3 + i);
    for (int i = tid; i < M; i += blockDim.x) S.mask[i] = valid[(size_t)b * M + i] ? 1.f : 0.f;
    for (int m = tid; m < M; m += blockDim.x) {
        int cnt = 0;
        for (int k = 0; k < NJ; ++k) {
            const float w = __ldg(Bm.Wm + m * NJ + k);
            if (w != 0.f) { S.Wm[m * NJ + cnt] = w; S.nzk[m * NJ + cnt] = (unsigned char)k; ++cnt; }
        }
        S.nzc[m] = cnt;
    }
    for (int i = tid; i < NJ; i += blockDim.x) S.anc[i] = __ldg(Bm.ancmask + i);
    __syncthreads();
    lm_eval(S, Bm);
    lm_jacobian(S, Bm, 10);
    for (int i = tid; i < M * 3; i += blockDim.x) res[(size_t)b * M * 3 + i] = S.res[i];
    for (int i = tid; i < M * 3 * DMAX; i += blockDim.x) jac[(size_t)b * M * 3 * DMAX + i] = S.J[(i / DMAX) * LDJ + i % DMAX];
}

__global__ void __launch_bounds__(256) lbs_kernel(const float* __restrict__ params, BodyFull Bf, float* __restrict__ verts,
                                                  float* __restrict__ joints) {
    __shared__ float sR[NJ * 9], sG[NJ * 12], sA[NJ * 12], sJ[NJ * 3], spf[207], sx[DMAX];
    const int b = blockIdx.y, tid = threadIdx.x;
    for (int i = tid; i < DMAX; i += 256) sx[i] = __ldg(params + (size_t)b * DMAX + i);
    __syncthreads();
    if (tid < NJ) {
        float dR[27];
        rodrigues_with_grad(sx + tid * 3, sR + tid * 9, dR);
        if (tid >= 1)
            for (int e = 0; e < 9; ++e) spf[(tid - 1) * 9 + e] = sR[tid * 9 + e] - (e % 4 == 0 ? 1.f : 0.f);
    } else if (tid >= 32 && tid < 32 + NJ * 3) {
        const int o = tid - 32;
        float v = __ldg(Bf.Jt + o);
        for (int l = 0; l < 10; ++l) v = fmaf(__ldg(Bf.Js + o * 10 + l), sx[72 + l], v);
        sJ[o] = v;
    }
    __syncthreads();
    if (tid == 0) {
        for (int j = 0; j < NJ; ++j) {
            const int pa = Bf.parents[j];
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 4; ++c) {
                    float v;
                    if (j == 0) v = c < 3 ? sR[r * 3 + c] : sJ[r];
                    else {
                        v = c == 3 ? sG[pa * 12 + r * 4 + 3] : 0.f;
                        for (int k = 0; k < 3; ++k) {
                            const float l = c < 3 ? sR[j * 9 + k * 3 + c] : (sJ[j * 3 + k] - sJ[pa * 3 + k]);
                            v = fmaf(sG[pa * 12 + r * 4 + k], l, v);
                        }
                    }
                    sG[j * 12 + r * 4 + c] = v;
                }
        }
        for (int j = 0; j < NJ; ++j)
            for (int r = 0; r < 3; ++r) {
                for (int c = 0; c < 3; ++c) sA[j * 12 + r * 4 + c] = sG[j * 12 + r * 4 + c];
                sA[j * 12 + r * 4 + 3] = sG[j * 12 + r * 4 + 3] -
                    (sG[j * 12 + r * 4] * sJ[j * 3] + sG[j * 12 + r * 4 + 1] * sJ[j * 3 + 1] + sG[j * 12 + r * 4 + 2] * sJ[j * 3 + 2]);
            }
    }
    __syncthreads();
    const int V = Bf.V;
    if (blockIdx.x == 0 && tid < NJ * 3)
        joints[((size_t)b * 45 + tid / 3) * 3 + tid % 3] = sG[(tid / 3) * 12 + (tid % 3) * 4 + 3] + sx[82 + tid % 3];
    for (int v = blockIdx.x * 256 + tid; v < V; v += gridDim.x * 256) {
        float vp[3];
        for (int c = 0; c < 3; ++c) {
            float a = __ldg(Bf.v_template + v * 3 + c);
            for (int l = 0; l < 10; ++l) a = fmaf(__ldg(Bf.shapedirs + (v * 3 + c) * 10 + l), sx[72 + l], a);
            float acc = 0.f;
            for (int k = 0; k < 207; ++k) acc = fmaf(spf[k], __ldg(Bf.posedirs + (size_t)k * V * 3 + v * 3 + c), acc);
            vp[c] = a + acc;
        }
        float T[12];
        for (int e = 0; e < 12; ++e) T[e] = 0.f;
        for (int k = 0; k < NJ; ++k) {
            const float w = __ldg(Bf.weights + v * NJ + k);
            if (w != 0.f)
                for (int e = 0; e < 12; ++e) T[e] = fmaf(w, sA[k * 12 + e], T[e]);
        }
        for (int c = 0; c < 3; ++c) {
            const float o = T[c * 4] * vp[0] + T[c * 4 + 1] * vp[1] + T[c * 4 + 2] * vp[2] + T[c * 4 + 3] + sx[82 + c];
            verts[((size_t)b * V + v) * 3 + c] = o;
        }
    }
}

__global__ void extra_joints_kernel(const float* __restrict__ verts, const int* __restrict__ extra_vids, int V,
                                    float* __restrict__ joints) {
    const int b = blockIdx.x, t = threadIdx.x;
    if (t < 21 * 3) joints[((size_t)b * 45 + 24 + t / 3) * 3 + t % 3] = verts[((size_t)b * V + extra_vids[t / 3]) * 3 + t % 3];
}

}  // namespace

// ================================================================================================ C ABI
// get_markers (src/models/fit_SMPL.py:17-62). labels int64 [B,N], conf [B,N], inner [B,N,3] -> markers [B,M,3], valid [B,M] u8
ETCH_API int etch_markers_top3(const float* inner, const long long* labels, const float* conf, int B, int N, int M,
                               float* markers, unsigned char* valid, cudaStream_t stream) {
    if (!inner || !labels || !conf || !markers || !valid || B <= 0 || N <= 0 || M <= 0) return ETCH_EINVAL;
    dim3 grid(etch_cdiv(M, 8), B);
    markers_kernel<<<grid, 256, 0, stream>>>(inner, labels, conf, N, M, markers, valid);
    ETCH_RETURN_LAST();
}

// Two-stage LM marker fit (fit_SMPL.py:157-255 with theseus semantics). params out: [B][85] = orient(3)|pose(69)|betas(10)|transl(3)
ETCH_API int etch_lm_fit(const float* markers, const unsigned char* valid, const float* Tm, const float* Sm, const float* Pm,
                         const float* Wm, const float* Jt, const float* Js, const int* parents, const unsigned* ancmask,
                         int B, int M, int steps0, int steps1, float step0, float step1, float damp0, float damp1,
                         float* params, int* iters, float* errs, cudaStream_t stream) {
    if (!markers || !valid || !Tm || !Sm || !Pm || !Wm || !Jt || !Js || !parents || !ancmask || !params || !iters || !errs ||
        B <= 0 || M <= 0 || M > NM_MAX)
        return ETCH_EINVAL;
    BodyMarkers Bm{Tm, Sm, Pm, Wm, Jt, Js, parents, ancmask, M};
    const size_t smem = sizeof(LmSmem);
    ETCH_TRY(cudaFuncSetAttribute(lm_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lm_fit_kernel<<<B, 256, smem, stream>>>(markers, valid, Bm, steps0, steps1, step0, step1, damp0, damp1, params, iters, errs, nullptr);
    ETCH_RETURN_LAST();
}

// same solve, additionally returning SM-clock totals per (scan, stage): [eval, jacobian, solve]  (profiling aid for tools/)
ETCH_API int etch_lm_fit_profile(const float* markers, const unsigned char* valid, const float* Tm, const float* Sm, const float* Pm,
                                 const float* Wm, const float* Jt, const float* Js, const int* parents, const unsigned* ancmask,
                                 int B, int M, int steps0, int steps1, float step0, float step1, float damp0, float damp1,
                                 float* params, int* iters, float* errs, long long* prof, cudaStream_t stream) {
    if (!markers || !valid || !params || !iters || !errs || !prof || B <= 0 || M <= 0 || M > NM_MAX) return ETCH_EINVAL;
    BodyMarkers Bm{Tm, Sm, Pm, Wm, Jt, Js, parents, ancmask, M};
    const size_t smem = sizeof(LmSmem);
    ETCH_TRY(cudaFuncSetAttribute(lm_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lm_fit_kernel<<<B, 256, smem, stream>>>(markers, valid, Bm, steps0, steps1, step0, step1, damp0, damp1, params, iters, errs, prof);
    ETCH_RETURN_LAST();
}

// SMPL forward for the fitted parameters (fit_SMPL.py:257-258): verts [B,V,3], joints [B,45,3] (translation applied)
ETCH_API int etch_lm_jacobian_dump(const float* params, const float* markers, const unsigned char* valid, const float* Tm, const float* Sm,
                                   const float* Pm, const float* Wm, const float* Jt, const float* Js, const int* parents,
                                   const unsigned* ancmask, int B, int M, float* res, float* jac, cudaStream_t stream) {
    if (!params || !markers || !valid || !Tm || !Sm || !Pm || !Wm || !Jt || !Js || !parents || !ancmask || !res || !jac || B <= 0 || M <= 0 || M > NM_MAX)
        return ETCH_EINVAL;
    BodyMarkers Bm{Tm, Sm, Pm, Wm, Jt, Js, parents, ancmask, M};
    ETCH_TRY(cudaFuncSetAttribute(lm_jacobian_dump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LmSmem)));
    lm_jacobian_dump_kernel<<<B, 256, sizeof(LmSmem), stream>>>(params, markers, valid, Bm, res, jac);
    ETCH_RETURN_LAST();
}

ETCH_API int etch_lbs_forward(const float* params, const float* v_template, const float* shapedirs, const float* posedirs,
                              const float* weights, const float* Jt, const float* Js, const int* parents,
                              const int* extra_vids, int B, int V, float* verts, float* joints, cudaStream_t stream) {
    if (!params || !v_template || !shapedirs || !posedirs || !weights || !Jt || !Js || !parents || !extra_vids || !verts || !joints)
        return ETCH_EINVAL;
    BodyFull Bf{v_template, shapedirs, posedirs, weights, Jt, Js, parents, extra_vids, V};
    dim3 grid(etch_cdiv(V, 256), B);
    lbs_kernel<<<grid, 256, 0, stream>>>(params, Bf, verts, joints);
    extra_joints_kernel<<<B, 64, 0, stream>>>(verts, extra_vids, V, joints);
    ETCH_RETURN_LAST();
}
