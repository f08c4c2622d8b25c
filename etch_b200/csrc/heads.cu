// 3-NN feature propagation + direction head (2-layer MHSA over the 60 anchors -> MLP -> so3 mean -> direction).
//
// Reference semantics restated:
//   src/models/pointnet2_utils.py:4-23,45-74      square_distance (-2ab + a^2 + b^2), full sort, 3-NN inverse-d^2 blend
//   src/models/models_pointcloud.py:161,181-184   channel layout c*60+a, reshape to [B,N,64,60], mean over anchors
//   src/models/direction_backbones.py:79-223      StackedMHSA(64, 128, 8 heads, 2 layers), BatchMLP
//   src/models/models_pointcloud.py:111-126       so3_reg, so3_mean, R * e_z
//   src/models/so3conv.py:186-225                 so3_mean: chordal mean via SVD, U diag(1,1,det(UV^T)) V^T
//
// B200 design: the reference materialises [B,N,3840] upsampled features (77 MB/scan) and runs 40k tiny bmm's per
// scan.  Here one CTA owns 2 points = 120 anchor tokens (padded to a 128-row GEMM tile), blends the three coarse
// feature rows straight from L2 into shared memory and keeps the whole token state (tokens, Q/K/V, attention output,
// MLP hidden) in the 227 KB of shared memory until the 3-vector direction comes out; only dir[3] + inv_feat[64] per
// point are written.  Algebraic fusions done on the host (exact in real arithmetic): 1/sqrt(d_k) folded into W_q,
// head_combine(layer 2) o Linear1 folded into one 64->128 map, Linear2 o so3_reg folded into one 128->1 map.
#include "common.cuh"

namespace {

constexpr int NA = 60;
constexpr int TOK = 128;   // padded tokens per tile (2 points x 60)
constexpr int LDT = 132;   // row stride of feature-major buffers [feat][token]
constexpr int LDQ = 196;   // row stride of the token-major QKV buffer [token][192]

// ------------------------------------------------------------------------------------------------ 3-NN
__global__ void __launch_bounds__(128) upsample3_kernel(const float* __restrict__ fine,    // [B,N,3]
                                                        const float* __restrict__ coarse,  // [B,3,S]
                                                        int N, int S, int* __restrict__ idx, float* __restrict__ w) {
    __shared__ float sx[512], sy[512], sz[512], sn[512];
    const int b = blockIdx.y;
    const int i = blockIdx.x * 128 + threadIdx.x;
    const bool act = i < N;
    float x = 0.f, y = 0.f, z = 0.f, s1 = 0.f;
    if (act) {
        const float* p = fine + ((size_t)b * N + i) * 3;
        x = __ldg(p); y = __ldg(p + 1); z = __ldg(p + 2);
        s1 = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
    }
    float d0 = INFINITY, d1 = INFINITY, d2 = INFINITY;
    int i0 = 0, i1 = 0, i2 = 0;
    const float* C = coarse + (size_t)b * 3 * S;
    for (int t0 = 0; t0 < S; t0 += 512) {
        const int cnt = min(512, S - t0);
        __syncthreads();
        for (int k = threadIdx.x; k < cnt; k += 128) {
            const float cx = __ldg(C + t0 + k), cy = __ldg(C + S + t0 + k), cz = __ldg(C + 2 * (size_t)S + t0 + k);
            sx[k] = cx; sy[k] = cy; sz[k] = cz;
            sn[k] = __fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz));
        }
        __syncthreads();
        if (act) {
            for (int k = 0; k < cnt; ++k) {
                const float dot = __fmaf_rn(z, sz[k], __fmaf_rn(y, sy[k], __fmul_rn(x, sx[k])));
                const float d = __fadd_rn(__fadd_rn(__fmul_rn(-2.f, dot), s1), sn[k]);
                if (d < d2) {
                    const int id = t0 + k;
                    if (d < d1) {
                        d2 = d1; i2 = i1;
                        if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = id; }
                        else { d1 = d; i1 = id; }
                    } else { d2 = d; i2 = id; }
                }
            }
        }
    }
    if (!act) return;
    const float r0 = 1.0f / (d0 + 1e-8f), r1 = 1.0f / (d1 + 1e-8f), r2 = 1.0f / (d2 + 1e-8f);
    const float nrm = (r0 + r1) + r2;
    int* oi = idx + ((size_t)b * N + i) * 3;
    float* ow = w + ((size_t)b * N + i) * 3;
    oi[0] = i0; oi[1] = i1; oi[2] = i2;
    ow[0] = r0 / nrm; ow[1] = r1 / nrm; ow[2] = r2 / nrm;
}

// ------------------------------------------------------------------------------------------------ block GEMM
// out[128 tokens][64 cols] = A[128][K] * B[K][n0..n0+64) (+bias, +relu, +residual).  A is feature-major in smem
// (At[k*LDT + m]); B is row-major in global memory with leading dimension ldb and is staged through s_B in 64-row
// chunks.  256 threads: thread (tr, tc) owns rows 4tr..4tr+3 and columns 8tc..8tc+7 of the 64-column chunk.
template <int K, bool OUT_TOKEN_MAJOR, bool RELU, bool RESIDUAL>
__device__ __forceinline__ void gemm_chunk(const float* __restrict__ At, const float* __restrict__ Bg, int ldb, int n0,
                                           const float* __restrict__ bias, float* __restrict__ out, int ldo,
                                           float* __restrict__ s_B) {
    const int tid = threadIdx.x, tr = tid & 31, tc = tid >> 5;
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += 64) {
        __syncthreads();
        for (int i = tid; i < 64 * 16; i += 256) {
            const int r = i >> 4, c4 = i & 15;
            reinterpret_cast<float4*>(s_B)[i] = __ldg(reinterpret_cast<const float4*>(Bg + (size_t)(k0 + r) * ldb + n0) + c4);
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 64; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(At + (k0 + k) * LDT + tr * 4);
            const float4 b0 = *reinterpret_cast<const float4*>(s_B + k * 64 + tc * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(s_B + k * 64 + tc * 8 + 4);
            const float a[4] = {av.x, av.y, av.z, av.w};
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float bv = bias ? __ldg(bias + n0 + tc * 8 + j) : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float v = acc[i][j] + bv;
            if (RELU) v = fmaxf(v, 0.f);
            acc[i][j] = v;
        }
    }
    if (OUT_TOKEN_MAJOR) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float* o = out + (tr * 4 + i) * ldo + n0 + tc * 8;
            *reinterpret_cast<float4*>(o) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            *reinterpret_cast<float4*>(o + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float* o = out + (n0 + tc * 8 + j) * ldo + tr * 4;
            float4 v = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
            if (RESIDUAL) {
                const float4 r = *reinterpret_cast<const float4*>(o);
                v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
            }
            *reinterpret_cast<float4*>(o) = v;
        }
    }
}

// one (point, query token, head) task: softmax(q K^T) V over the 60 anchor tokens of the point
__device__ __forceinline__ void attention_tile(const float* __restrict__ s_qkv, float* __restrict__ s_ot) {
    for (int t = threadIdx.x; t < 2 * 8 * NA; t += 256) {
        const int i = t % NA, rest = t / NA, h = rest & 7, pl = rest >> 3;
        const int tok = pl * NA + i;
        const float4 qa = *reinterpret_cast<const float4*>(s_qkv + tok * LDQ + h * 8);
        const float4 qb = *reinterpret_cast<const float4*>(s_qkv + tok * LDQ + h * 8 + 4);
        float s[NA];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < NA; ++j) {
            const float* kr = s_qkv + (pl * NA + j) * LDQ + 64 + h * 8;
            const float4 ka = *reinterpret_cast<const float4*>(kr);
            const float4 kb = *reinterpret_cast<const float4*>(kr + 4);
            float v = qa.x * ka.x;
            v = fmaf(qa.y, ka.y, v); v = fmaf(qa.z, ka.z, v); v = fmaf(qa.w, ka.w, v);
            v = fmaf(qb.x, kb.x, v); v = fmaf(qb.y, kb.y, v); v = fmaf(qb.z, kb.z, v); v = fmaf(qb.w, kb.w, v);
            s[j] = v;
            mx = fmaxf(mx, v);
        }
        float sum = 0.f;
        float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < NA; ++j) {
            const float p = expf(s[j] - mx);
            sum += p;
            const float* vr = s_qkv + (pl * NA + j) * LDQ + 128 + h * 8;
            const float4 va = *reinterpret_cast<const float4*>(vr);
            const float4 vb = *reinterpret_cast<const float4*>(vr + 4);
            o[0] = fmaf(p, va.x, o[0]); o[1] = fmaf(p, va.y, o[1]); o[2] = fmaf(p, va.z, o[2]); o[3] = fmaf(p, va.w, o[3]);
            o[4] = fmaf(p, vb.x, o[4]); o[5] = fmaf(p, vb.y, o[5]); o[6] = fmaf(p, vb.z, o[6]); o[7] = fmaf(p, vb.w, o[7]);
        }
        const float inv = 1.0f / sum;
#pragma unroll
        for (int d = 0; d < 8; ++d) s_ot[(h * 8 + d) * LDT + tok] = o[d] * inv;
    }
}

// symmetric 3x3 eigen-decomposition (cyclic Jacobi, double) -> eigenvectors in columns of V, eigenvalues in e
__device__ void jacobi3(double A[3][3], double V[3][3], double e[3]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) V[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (fabs(A[p][q]) < 1e-300) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    e[0] = A[0][0]; e[1] = A[1][1]; e[2] = A[2][2];
}

// third column of R = U diag(1,1,det(UV^T)) V^T for Ce = U S V^T; with (u0,u1,v0,v1) the two dominant singular
// pairs this equals u0 v0[2] + u1 v1[2] + (u0 x u1)(v0 x v1)[2], independent of the SVD sign conventions.
__device__ void so3_direction(const double Ce[3][3], float out[3]) {
    double A[3][3], V[3][3], e[3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A[i][j] = Ce[0][i] * Ce[0][j] + Ce[1][i] * Ce[1][j] + Ce[2][i] * Ce[2][j];
    jacobi3(A, V, e);
    int i0 = 0;
    if (e[1] > e[i0]) i0 = 1;
    if (e[2] > e[i0]) i0 = 2;
    int i1 = (i0 + 1) % 3, i2 = (i0 + 2) % 3;
    if (e[i2] > e[i1]) { const int t = i1; i1 = i2; i2 = t; }
    double v0[3] = {V[0][i0], V[1][i0], V[2][i0]}, v1[3] = {V[0][i1], V[1][i1], V[2][i1]};
    double u0[3], u1[3];
    for (int i = 0; i < 3; ++i) {
        u0[i] = Ce[i][0] * v0[0] + Ce[i][1] * v0[1] + Ce[i][2] * v0[2];
        u1[i] = Ce[i][0] * v1[0] + Ce[i][1] * v1[1] + Ce[i][2] * v1[2];
    }
    const double n0 = sqrt(u0[0] * u0[0] + u0[1] * u0[1] + u0[2] * u0[2]) + 1e-300;
    for (int i = 0; i < 3; ++i) u0[i] /= n0;
    const double d01 = u0[0] * u1[0] + u0[1] * u1[1] + u0[2] * u1[2];  // re-orthogonalise (exactly 0 in exact arithmetic)
    for (int i = 0; i < 3; ++i) u1[i] -= d01 * u0[i];
    const double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]) + 1e-300;
    for (int i = 0; i < 3; ++i) u1[i] /= n1;
    const double u2[3] = {u0[1] * u1[2] - u0[2] * u1[1], u0[2] * u1[0] - u0[0] * u1[2], u0[0] * u1[1] - u0[1] * u1[0]};
    const double v2z = v0[0] * v1[1] - v0[1] * v1[0];
    for (int i = 0; i < 3; ++i) out[i] = (float)(u0[i] * v0[2] + u1[i] * v1[2] + u2[i] * v2z);
}

struct DirWeights {
    const float* Wqkv1;  // [64][192]  (W_q/sqrt(8) | W_k | W_v)^T of layer 0
    const float* Wc1;    // [64][64]   head_combine^T layer 0
    const float* bc1;    // [64]
    const float* Wqkv2;  // [64][192]  layer 1
    const float* Wf;     // [64][128]  (Linear1 o head_combine(layer 1))^T
    const float* bf;     // [128]
    const float* vreg;   // [128]      Linear2^T so3_reg
    float creg;          //            so3_reg(bias of Linear2) + so3_reg bias
    const float* anchors;  // [60][9]
};

__global__ void __launch_bounds__(256, 1) direction_head_kernel(
    const float* __restrict__ feats,  // [B,S,60,64] coarse equivariant features
    const int* __restrict__ up_idx,   // [B,N,3]
    const float* __restrict__ up_w,   // [B,N,3]
    DirWeights W, int N, int S,
    float* __restrict__ dir,          // [B,N,3]
    float* __restrict__ inv,          // [B,N,64] anchor-mean ("invariant") feature
    float* __restrict__ anc_w)        // [B,N,60] raw anchor weights (nullable)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_t = reinterpret_cast<float*>(smem_raw);   // [64][LDT]   tokens, feature-major
    float* s_o = s_t + 64 * LDT;                        // [64][LDT]   attention output, feature-major
    float* s_qkv = s_o + 64 * LDT;                      // [128][LDQ]  token-major Q|K|V  (later: hidden [128][LDT])
    float* s_B = s_qkv + TOK * LDQ;                     // [64][64]    weight staging
    float* s_w = s_B + 64 * 64;                         // [128]       anchor weights
    float* s_anc = s_w + TOK;                           // [60][9]
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < NA * 9; i += 256) s_anc[i] = __ldg(W.anchors + i);
    const int ntiles = (N + 1) / 2;
    const float* F = feats + (size_t)b * S * NA * 64;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int p0 = tile * 2;
        __syncthreads();
        // ---- 1. blend the three coarse rows into the token tile (feature-major) ----
        for (int tok = warp; tok < TOK; tok += 8) {
            float v0 = 0.f, v1 = 0.f;
            if (tok < 2 * NA) {
                const int pl = tok / NA, a = tok % NA;
                const int p = min(p0 + pl, N - 1);
                const int* ip = up_idx + ((size_t)b * N + p) * 3;
                const float* wp = up_w + ((size_t)b * N + p) * 3;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float wk = __ldg(wp + k);
                    const float* row = F + ((size_t)__ldg(ip + k) * NA + a) * 64;
                    v0 = fmaf(__ldg(row + lane), wk, v0);
                    v1 = fmaf(__ldg(row + lane + 32), wk, v1);
                }
            }
            s_t[lane * LDT + tok] = v0;
            s_t[(lane + 32) * LDT + tok] = v1;
        }
        __syncthreads();
        if (tid < 128) {
            const int pl = tid >> 6, c = tid & 63;
            if (p0 + pl < N) {
                float s = 0.f;
                for (int a = 0; a < NA; ++a) s += s_t[c * LDT + pl * NA + a];
                inv[((size_t)b * N + p0 + pl) * 64 + c] = s / 60.0f;
            }
        }
        // ---- 2. layer 0: x += combine(attn(x)) ----
        for (int n0 = 0; n0 < 192; n0 += 64) gemm_chunk<64, true, false, false>(s_t, W.Wqkv1, 192, n0, nullptr, s_qkv, LDQ, s_B);
        __syncthreads();
        attention_tile(s_qkv, s_o);
        gemm_chunk<64, false, false, true>(s_o, W.Wc1, 64, 0, W.bc1, s_t, LDT, s_B);
        // ---- 3. layer 1: x = combine(attn(x)), fused with Linear1 + ReLU ----
        for (int n0 = 0; n0 < 192; n0 += 64) gemm_chunk<64, true, false, false>(s_t, W.Wqkv2, 192, n0, nullptr, s_qkv, LDQ, s_B);
        __syncthreads();
        attention_tile(s_qkv, s_o);
        float* s_h = s_qkv;  // [128][LDT] hidden, feature-major (QKV no longer needed)
        __syncthreads();
        for (int n0 = 0; n0 < 128; n0 += 64) gemm_chunk<64, false, true, false>(s_o, W.Wf, 128, n0, W.bf, s_h, LDT, s_B);
        __syncthreads();
        // ---- 4. anchor weights, chordal mean, direction ----
        if (tid < TOK) {
            float s = W.creg;
            for (int f = 0; f < 128; ++f) s = fmaf(s_h[f * LDT + tid], __ldg(W.vreg + f), s);
            s_w[tid] = s;
            if (anc_w && tid < 2 * NA && p0 + tid / NA < N) anc_w[((size_t)b * N + p0) * NA + tid] = s;
        }
        __syncthreads();
        if (tid < 2 && p0 + tid < N) {
            double Ce[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
            for (int a = 0; a < NA; ++a) {
                const float wa = s_w[tid * NA + a];
                for (int e = 0; e < 9; ++e) Ce[e / 3][e % 3] += (double)(wa * s_anc[a * 9 + e]);
            }
            float d[3];
            so3_direction(Ce, d);
            float* o = dir + ((size_t)b * N + p0 + tid) * 3;
            o[0] = d[0]; o[1] = d[1]; o[2] = d[2];
        }
    }
}

}  // namespace

// ================================================================================================ C ABI
// 3 nearest coarse points + normalised inverse-d^2 weights (PointFeatPropagation, src/models/pointnet2_utils.py:45-74)
ETCH_API int etch_upsample3(const float* fine_bn3, const float* coarse_b3s, int B, int N, int S, int* idx, float* w,
                            cudaStream_t stream) {
    if (!fine_bn3 || !coarse_b3s || !idx || !w || B <= 0 || N <= 0 || S < 3) return ETCH_EINVAL;
    dim3 grid(etch_cdiv(N, 128), B);
    upsample3_kernel<<<grid, 128, 0, stream>>>(fine_bn3, coarse_b3s, N, S, idx, w);
    ETCH_RETURN_LAST();
}

// Fused feature propagation + anchor mean + decode_direction (src/models/models_pointcloud.py:111-126,181-184)
ETCH_API int etch_direction_head(const float* feats, const int* up_idx, const float* up_w, const float* Wqkv1,
                                 const float* Wc1, const float* bc1, const float* Wqkv2, const float* Wf, const float* bf,
                                 const float* vreg, float creg, const float* anchors, int B, int N, int S, float* dir,
                                 float* inv, float* anc_w, cudaStream_t stream) {
    if (!feats || !up_idx || !up_w || !Wqkv1 || !Wc1 || !bc1 || !Wqkv2 || !Wf || !bf || !vreg || !anchors || !dir || !inv)
        return ETCH_EINVAL;
    DirWeights W{Wqkv1, Wc1, bc1, Wqkv2, Wf, bf, vreg, creg, anchors};
    constexpr size_t smem = (size_t)(2 * 64 * LDT + TOK * LDQ + 64 * 64 + TOK + NA * 9) * 4;
    static_assert(TOK * LDQ >= 128 * LDT, "hidden buffer must fit the QKV region");
    ETCH_TRY(cudaFuncSetAttribute(direction_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ntiles = (N + 1) / 2;
    int gx = etch_sm_budget() / B;  // persistent CTAs, 1 per SM (smem-bound): never more than one wave across the whole batch
    if (gx < 1) gx = 1;
    if (gx > ntiles) gx = ntiles;
    dim3 grid(gx, B);
    direction_head_kernel<<<grid, 256, smem, stream>>>(feats, up_idx, up_w, W, N, S, dir, inv, anc_w);
    ETCH_RETURN_LAST();
}
