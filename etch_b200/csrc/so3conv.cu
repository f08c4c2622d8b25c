// SO(3)-equivariant encoder kernels (EPN / SO3Conv with 60 icosahedral anchors, 24 kernel points).
//
// Reference semantics restated:
//   external/vgtk/vgtk/so3conv/functional.py:286-324  inter_so3conv_grouping_anchor  w = relu(1 - |g - R_a k|^2 / sigma)
//   external/vgtk/vgtk/so3conv/functional.py:61-67    inter_so3conv_feat_grouping    y[c,k,p,a] = sum_n f[c,nbr,a] w[p,a,k,n]
//   external/vgtk/vgtk/so3conv/modules.py:19-39       BasicSO3Conv                   z = W[o,(c,k)] y + bias
//   external/vgtk/vgtk/so3conv/functional.py:331-343  intra_so3conv_grouping         y[c,j,p,a] = x[c,p,intra_idx[a,j]]
//   src/models/so3conv.py:36-44,94-103,171-183        InstanceNorm2d(eps 1e-5, biased var) + leaky_relu(0.01), skip branch
//
// B200 design: features live point-major [B, P, 60, C] (one 7.7-15 KB contiguous row per point, so neighbour
// gathers are whole-sector reads that stay in the 126 MB L2).  The reference's inter_w [B,P,60,24,nn] (0.9 GB/scan)
// and grouped-feature tensors are never materialised: a CTA owns 2 output points x 60 anchors (= 120 (p,a) pairs,
// padded to 128 GEMM rows), regenerates the kernel weights in registers, contracts neighbour features into a
// [192 x 128] shared-memory tile per 8-channel chunk and immediately multiplies it with the matching slice of W.
// InstanceNorm statistics are accumulated in the epilogue (double atomics per CTA) and the normalisation itself is
// folded into the load path of the consumer kernel, so each activation makes one HBM round trip.
#include "common.cuh"

namespace {

constexpr int NA = 60;   // anchors
constexpr int NK = 24;   // kernel points
constexpr int TP = 2;    // output points per tile
constexpr int NPAIR = TP * NA;  // 120 valid GEMM rows
constexpr int MROWS = 128;      // padded

struct Stats {  // per-(batch, channel) running sums -> mean / rstd
    const double* sums;  // [B][C][2]
    double count;        // elements per (b,c) = P*60
};

__device__ __forceinline__ void stats_to_affine(const double* sums, int b, int C, double count, float* s_mean, float* s_rstd) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const double s = sums[((size_t)b * C + c) * 2], ss = sums[((size_t)b * C + c) * 2 + 1];
        const double mean = s / count;
        double var = ss / count - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[c] = (float)mean;
        s_rstd[c] = (float)(1.0 / sqrt(var + 1e-5));
    }
}

// ------------------------------------------------------------------------------------------------
// Layer b0.0: c_in = 1 and the input feature is identically 1  =>  y[a,k] = sum_n w[a,k,n];  z = W y + bias.
// 480 threads: thread t owns the (anchor, kernel point) pairs t, t + 480, t + 960 for every point it sees, so their rotated kernel
// points stay in registers ({2/sigma R_a k, |R_a k|^2/sigma}: the expanded form of the weight that the c_in > 1 kernel uses too,
// w = relu(g . kq.xyz + (1 - |g|^2/sigma) - kq.w): 3 FFMA + FADD + FMNMX per (pair, neighbour) instead of 12 instructions, and since
// round 2 two neighbours per packed FFMA2 / FADD2) and two broadcast LDS.128 per neighbour pair feed three pairs.
__device__ __forceinline__ uint64_t c1_pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void c1_unpk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t c1_fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t c1_add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

constexpr int C1_THREADS = 480;
template <int COUT>
__global__ void __launch_bounds__(C1_THREADS) inter_conv_c1_kernel(
    const float* __restrict__ xyz,       // [B,3,q] support points
    const int* __restrict__ sample_idx,  // [B,P]   centres (indices into q)
    const int* __restrict__ nbr,         // [B,P,nn]
    const float* __restrict__ kr,        // [60,24,3] rotated kernel points R_a k
    const float* __restrict__ Wt,        // [24][COUT]
    const float* __restrict__ bias,      // [COUT]
    int q, int P, int nn, float inv_sigma,
    float* __restrict__ zraw,            // [B,P,60,COUT]
    double* __restrict__ stats)          // [B][COUT][2]
{
    static_assert(NA * NK == 3 * C1_THREADS, "three pairs per thread");
    __shared__ __align__(16) float s_W[NK * COUT];
    __shared__ __align__(16) float s_bias[COUT];
    __shared__ float4 s_g[64];
    __shared__ float s_y[NA * NK];
    __shared__ __align__(16) float s_z[NA * COUT];
    const int b = blockIdx.y, tid = threadIdx.x;
    for (int i = tid; i < NK * COUT; i += C1_THREADS) s_W[i] = __ldg(Wt + i);
    if (tid < COUT) s_bias[tid] = __ldg(bias + tid);
    float4 kq[3];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
        const int t = tid + u * C1_THREADS;
        const float kx = __ldg(kr + t * 3), ky = __ldg(kr + t * 3 + 1), kz = __ldg(kr + t * 3 + 2);
        kq[u] = make_float4(2.0f * inv_sigma * kx, 2.0f * inv_sigma * ky, 2.0f * inv_sigma * kz, (kx * kx + ky * ky + kz * kz) * inv_sigma);
    }
    uint64_t kx2[3], ky2[3], kz2[3], kw2[3];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
        kx2[u] = c1_pk(kq[u].x, kq[u].x); ky2[u] = c1_pk(kq[u].y, kq[u].y); kz2[u] = c1_pk(kq[u].z, kq[u].z); kw2[u] = c1_pk(-kq[u].w, -kq[u].w);
    }
    const float* X = xyz + (size_t)b * 3 * q;
    double acc_s = 0.0, acc_ss = 0.0;
    for (int p = blockIdx.x; p < P; p += gridDim.x) {
        __syncthreads();
        if (tid < nn) {
            const int c = __ldg(sample_idx + (size_t)b * P + p);
            const int k = __ldg(nbr + ((size_t)b * P + p) * nn + tid);
            const float gx = __ldg(X + k) - __ldg(X + c), gy = __ldg(X + q + k) - __ldg(X + q + c);
            const float gz = __ldg(X + 2 * (size_t)q + k) - __ldg(X + 2 * (size_t)q + c);
            // pair layout: neighbours (2m, 2m+1) -> {gx0, gx1, gy0, gy1}, {gz0, gz1, gw0, gw1}
            float* gp = reinterpret_cast<float*>(s_g + (tid & ~1));
            const int e = tid & 1;
            gp[e] = gx; gp[2 + e] = gy; gp[4 + e] = gz; gp[6 + e] = 1.0f - (gx * gx + gy * gy + gz * gz) * inv_sigma;
        } else if (tid == nn && (nn & 1)) {   // odd neighbour count: the pad slot of the last pair contributes relu(-inf) = 0
            float* gp = reinterpret_cast<float*>(s_g + (tid & ~1));
            gp[1] = 0.f; gp[3] = 0.f; gp[5] = 0.f; gp[7] = -INFINITY;
        }
        __syncthreads();
        // two neighbours per packed fma.rn.f32x2 (same bits as the scalar expression, same summation order): 26 instead of 38
        // instructions per neighbour pair
        float sacc[3] = {0.f, 0.f, 0.f};
#pragma unroll 2
        for (int n2 = 0; n2 < (nn + 1) / 2; ++n2) {
            const ulonglong2 ga = *reinterpret_cast<const ulonglong2*>(s_g + 2 * n2);       // {gx0,gx1}, {gy0,gy1}
            const ulonglong2 gb = *reinterpret_cast<const ulonglong2*>(s_g + 2 * n2 + 1);   // {gz0,gz1}, {gw0,gw1}
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                uint64_t v = c1_add2(gb.y, kw2[u]);
                v = c1_fma2(gb.x, kz2[u], v); v = c1_fma2(ga.y, ky2[u], v); v = c1_fma2(ga.x, kx2[u], v);
                float w0, w1;
                c1_unpk(v, w0, w1);
                sacc[u] += fmaxf(w0, 0.f);
                sacc[u] += fmaxf(w1, 0.f);
            }
        }
        const float s0 = sacc[0], s1 = sacc[1], s2 = sacc[2];
        s_y[tid] = s0; s_y[tid + C1_THREADS] = s1; s_y[tid + 2 * C1_THREADS] = s2;
        __syncthreads();
        // channel mixing 24 -> COUT: thread (anchor, 4 consecutive output channels): one scalar y and one LDS.128 of W per 4 FMAs
        for (int t = tid; t < NA * (COUT / 4); t += C1_THREADS) {
            const int a = t / (COUT / 4), o4 = (t % (COUT / 4)) * 4;
            float4 z = *reinterpret_cast<const float4*>(s_bias + o4);
#pragma unroll
            for (int k = 0; k < NK; ++k) {
                const float y = s_y[a * NK + k];
                const float4 w = *reinterpret_cast<const float4*>(s_W + k * COUT + o4);
                z.x = fmaf(w.x, y, z.x); z.y = fmaf(w.y, y, z.y); z.z = fmaf(w.z, y, z.z); z.w = fmaf(w.w, y, z.w);
            }
            *reinterpret_cast<float4*>(s_z + a * COUT + o4) = z;
            *reinterpret_cast<float4*>(zraw + ((size_t)b * P + p) * NA * COUT + a * COUT + o4) = z;
        }
        __syncthreads();
        if (tid < COUT) {
            float s = 0.f, ss = 0.f;
            for (int a = 0; a < NA; ++a) { const float v = s_z[a * COUT + tid]; s += v; ss = fmaf(v, v, ss); }
            acc_s += (double)s; acc_ss += (double)ss;
        }
    }
    if (tid < COUT) {
        atomicAdd(stats + ((size_t)b * COUT + tid) * 2, acc_s);
        atomicAdd(stats + ((size_t)b * COUT + tid) * 2 + 1, acc_ss);
    }
}

// ------------------------------------------------------------------------------------------------
// General inter conv: fused weight generation + neighbour contraction (GEMM1) + channel mixing (GEMM2).
template <int CIN, int COUT, int NN>
__global__ void __launch_bounds__(256, 1) inter_conv_kernel(
    const float* __restrict__ xyz,       // [B,3,q]
    const float* __restrict__ feat,      // [B,q,60,CIN] finished features of the previous conv
    const int* __restrict__ sample_idx,  // [B,P]
    const int* __restrict__ nbr,         // [B,P,NN]
    const float4* __restrict__ krs,      // [60,24] {2/sigma * R_a k, |R_a k|^2/sigma}
    const float* __restrict__ Wt,        // [CIN*24][COUT]  (row = c*24+k)
    const float* __restrict__ bias,      // [COUT]
    int q, int P, float inv_sigma,
    float* __restrict__ zraw,            // [B,P,60,COUT]
    double* __restrict__ stats)
{
    constexpr int CC = 8;             // channels per chunk
    constexpr int KK = CC * NK;       // 192 GEMM2 k-rows per chunk
    constexpr int TN = COUT / 8;      // output columns per thread in GEMM2
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* s_krs = reinterpret_cast<float4*>(smem_raw);                    // 1440 float4 = 23040 B
    float4* s_g = s_krs + NA * NK;                                          // [TP][NN] {gx,gy,gz,1-|g|^2/sigma}
    int* s_off = reinterpret_cast<int*>(s_g + TP * NN);                     // [TP][NN] feature row offsets (in floats)
    float* s_y = reinterpret_cast<float*>(s_off + TP * NN);                 // [KK][MROWS]
    float* s_W = s_y + KK * MROWS;                                          // [KK][COUT]
    float* s_z = s_y;                                                       // epilogue reuse [MROWS][COUT]

    const int b = blockIdx.y, tid = threadIdx.x;
    const int tr = tid & 31, tc = tid >> 5;
    for (int i = tid; i < NA * NK; i += 256) s_krs[i] = __ldg(krs + i);
    const float* X = xyz + (size_t)b * 3 * q;
    const float* F = feat + (size_t)b * q * NA * CIN;
    double acc_s = 0.0, acc_ss = 0.0;
    const int ntiles = (P + TP - 1) / TP;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int p0 = tile * TP;
        const int npts = min(TP, P - p0);
        __syncthreads();
        for (int t = tid; t < TP * NN; t += 256) {
            const int pl = t / NN, n = t % NN;
            const int p = min(p0 + pl, P - 1);
            const int c = __ldg(sample_idx + (size_t)b * P + p);
            const int k = __ldg(nbr + ((size_t)b * P + p) * NN + n);
            const float gx = __ldg(X + k) - __ldg(X + c);
            const float gy = __ldg(X + q + k) - __ldg(X + q + c);
            const float gz = __ldg(X + 2 * (size_t)q + k) - __ldg(X + 2 * (size_t)q + c);
            s_g[t] = make_float4(gx, gy, gz, 1.0f - (gx * gx + gy * gy + gz * gz) * inv_sigma);
            s_off[t] = k * NA * CIN;
        }
        float zacc[4][TN];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) zacc[i][j] = 0.f;
        __syncthreads();

        for (int c0 = 0; c0 < CIN; c0 += CC) {
            // ---- GEMM1: y[(c,k)][pair] = sum_n f[nbr_n][a][c] * w[a][k][n], task = (pair, 6 kernel points)
            // task = (pair, group of 6 kernel points).  A warp covers 8 consecutive pairs x the 4 groups, so the 4 lanes
            // sharing a pair read the same feature sectors (one L1 wavefront instead of four) and a warp-wide LDG touches
            // 8 cache lines instead of 32 -- the kernel was L1tex-wavefront bound with the (pair-major) mapping.
            for (int grp = (tid >> 5); grp < NPAIR / 8; grp += 8) {
                const int pair = grp * 8 + (tid & 7), kg = (tid >> 3) & 3;
                const int pl = pair / NA, a = pair % NA;
                float4 kq[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) kq[i] = s_krs[a * NK + kg * 6 + i];
                float acc[CC][6];
#pragma unroll
                for (int c = 0; c < CC; ++c)
#pragma unroll
                    for (int i = 0; i < 6; ++i) acc[c][i] = 0.f;
                const float* fa = F + a * CIN + c0;
                // software pipeline: the feature row of neighbour n+1 is in flight while neighbour n is consumed
                float4 f0 = __ldg(reinterpret_cast<const float4*>(fa + s_off[pl * NN]));
                float4 f1 = __ldg(reinterpret_cast<const float4*>(fa + s_off[pl * NN]) + 1);
#pragma unroll 2
                for (int n = 0; n < NN; ++n) {
                    const float4 g = s_g[pl * NN + n];
                    const float fv[CC] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
                    if (n + 1 < NN) {
                        const float4* fp = reinterpret_cast<const float4*>(fa + s_off[pl * NN + n + 1]);
                        f0 = __ldg(fp); f1 = __ldg(fp + 1);
                    }
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        const float w = fmaxf(fmaf(g.x, kq[i].x, fmaf(g.y, kq[i].y, fmaf(g.z, kq[i].z, g.w - kq[i].w))), 0.f);
#pragma unroll
                        for (int c = 0; c < CC; ++c) acc[c][i] = fmaf(fv[c], w, acc[c][i]);
                    }
                }
#pragma unroll
                for (int c = 0; c < CC; ++c)
#pragma unroll
                    for (int i = 0; i < 6; ++i) s_y[(c * NK + kg * 6 + i) * MROWS + pair] = acc[c][i];
            }
            // ---- stage W chunk
            {
                const float4* src = reinterpret_cast<const float4*>(Wt + (size_t)c0 * NK * COUT);
                float4* dst = reinterpret_cast<float4*>(s_W);
                for (int i = tid; i < KK * COUT / 4; i += 256) dst[i] = __ldg(src + i);
            }
            __syncthreads();
            // ---- GEMM2: z[pair][o] += sum_kk y[kk][pair] * W[kk][o]
#pragma unroll 4
            for (int kk = 0; kk < KK; ++kk) {
                const float4 yv = *reinterpret_cast<const float4*>(s_y + kk * MROWS + tr * 4);
                const float ya[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
                for (int j4 = 0; j4 < TN / 4; ++j4) {
                    const float4 wv = *reinterpret_cast<const float4*>(s_W + kk * COUT + tc * TN + j4 * 4);
                    const float wa[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) zacc[i][j4 * 4 + j] = fmaf(ya[i], wa[j], zacc[i][j4 * 4 + j]);
                }
            }
            __syncthreads();
        }
        // ---- epilogue: bias, write, statistics
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) s_z[(tr * 4 + i) * COUT + tc * TN + j] = zacc[i][j] + __ldg(bias + tc * TN + j);
        __syncthreads();
        const int nvalid = npts * NA;
        float* dst = zraw + ((size_t)b * P + p0) * NA * COUT;
        for (int i = tid; i < nvalid * COUT / 4; i += 256)
            reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(s_z)[i];
        if (tid < COUT) {
            float s = 0.f, ss = 0.f;
            for (int r = 0; r < nvalid; ++r) { const float v = s_z[r * COUT + tid]; s += v; ss = fmaf(v, v, ss); }
            acc_s += (double)s; acc_ss += (double)ss;
        }
    }
    if (tid < COUT) {
        atomicAdd(stats + ((size_t)b * COUT + tid) * 2, acc_s);
        atomicAdd(stats + ((size_t)b * COUT + tid) * 2 + 1, acc_ss);
    }
}

// ------------------------------------------------------------------------------------------------
// Anchor-gather GEMM: z[p,a,o] = bias[o] + sum_{j<J} sum_c W[(j,c)][o] * x[src(p)][tab[a][j]][c]
//   intra conv : J = 12, tab = intra_idx, x = leaky_relu(InstanceNorm(zin)) applied on load, src(p) = p
//   skip conv  : J = 1,  tab[a][0] = a,  x = raw finished features, src(p) = sample_idx[p]
template <int CIN, int COUT, int J, bool NORM_IN>
__global__ void __launch_bounds__(256, 1) anchor_gemm_kernel(
    const float* __restrict__ xin,       // [B,Q,60,CIN]
    const int* __restrict__ src_idx,     // [B,P] or nullptr (identity)
    const int* __restrict__ tab,         // [60][J]
    const float* __restrict__ Wt,        // [J][CIN][COUT]
    const float* __restrict__ bias,      // [COUT]
    const double* __restrict__ in_stats, // [B][CIN][2] (NORM_IN)
    double in_count, int Q, int P,
    float* __restrict__ zraw,            // [B,P,60,COUT]
    double* __restrict__ stats)
{
    constexpr int LD = CIN + 4;        // padded row stride (floats) of the staged activations
    constexpr int TN = COUT / 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_x = reinterpret_cast<float*>(smem_raw);              // [TP*60][LD]
    float* s_W = s_x + TP * NA * LD;                               // [CIN][COUT]
    float* s_z = s_W + CIN * COUT;                                 // [MROWS][COUT]
    float* s_mean = s_z + MROWS * COUT;                            // [CIN]
    float* s_rstd = s_mean + CIN;
    int* s_tab = reinterpret_cast<int*>(s_rstd + CIN);             // [60][J]

    const int b = blockIdx.y, tid = threadIdx.x;
    const int tr = tid & 31, tc = tid >> 5;
    for (int i = tid; i < NA * J; i += 256) s_tab[i] = __ldg(tab + i);
    if (NORM_IN) stats_to_affine(in_stats, b, CIN, in_count, s_mean, s_rstd);
    double acc_s = 0.0, acc_ss = 0.0;
    const int ntiles = (P + TP - 1) / TP;
    int prow[4], pa[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int pair = min(tr * 4 + i, NPAIR - 1);
        prow[i] = (pair / NA) * NA;
        pa[i] = pair % NA;
    }

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int p0 = tile * TP;
        const int npts = min(TP, P - p0);
        __syncthreads();
        // stage (and normalise) the activations of the tile's points
        for (int t = tid; t < TP * NA * (CIN / 4); t += 256) {
            const int row = t / (CIN / 4), c4 = t % (CIN / 4);
            const int pl = row / NA;
            const int p = min(p0 + pl, P - 1);
            const int sp = src_idx ? __ldg(src_idx + (size_t)b * P + p) : p;
            float4 v = __ldg(reinterpret_cast<const float4*>(xin + (((size_t)b * Q + sp) * NA + (row % NA)) * CIN) + c4);
            if (NORM_IN) {
                const int c = c4 * 4;
                v.x = etch_lrelu((v.x - s_mean[c]) * s_rstd[c]);
                v.y = etch_lrelu((v.y - s_mean[c + 1]) * s_rstd[c + 1]);
                v.z = etch_lrelu((v.z - s_mean[c + 2]) * s_rstd[c + 2]);
                v.w = etch_lrelu((v.w - s_mean[c + 3]) * s_rstd[c + 3]);
            }
            *reinterpret_cast<float4*>(s_x + row * LD + c4 * 4) = v;
        }
        float zacc[4][TN];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) zacc[i][j] = 0.f;

        for (int j = 0; j < J; ++j) {
            __syncthreads();
            {
                const float4* src = reinterpret_cast<const float4*>(Wt + (size_t)j * CIN * COUT);
                float4* dst = reinterpret_cast<float4*>(s_W);
                for (int i = tid; i < CIN * COUT / 4; i += 256) dst[i] = __ldg(src + i);
            }
            __syncthreads();
            const float* xr[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) xr[i] = s_x + (prow[i] + s_tab[pa[i] * J + j]) * LD;
#pragma unroll 2
            for (int c = 0; c < CIN; c += 4) {
                float xa[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 v = *reinterpret_cast<const float4*>(xr[i] + c);
                    xa[i][0] = v.x; xa[i][1] = v.y; xa[i][2] = v.z; xa[i][3] = v.w;
                }
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
#pragma unroll
                    for (int j4 = 0; j4 < TN / 4; ++j4) {
                        const float4 wv = *reinterpret_cast<const float4*>(s_W + (c + cc) * COUT + tc * TN + j4 * 4);
                        const float wa[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) zacc[i][j4 * 4 + jj] = fmaf(xa[i][cc], wa[jj], zacc[i][j4 * 4 + jj]);
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) s_z[(tr * 4 + i) * COUT + tc * TN + j] = zacc[i][j] + __ldg(bias + tc * TN + j);
        __syncthreads();
        const int nvalid = npts * NA;
        float* dst = zraw + ((size_t)b * P + p0) * NA * COUT;
        for (int i = tid; i < nvalid * COUT / 4; i += 256)
            reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(s_z)[i];
        if (tid < COUT) {
            float s = 0.f, ss = 0.f;
            for (int r = 0; r < nvalid; ++r) { const float v = s_z[r * COUT + tid]; s += v; ss = fmaf(v, v, ss); }
            acc_s += (double)s; acc_ss += (double)ss;
        }
    }
    if (tid < COUT) {
        atomicAdd(stats + ((size_t)b * COUT + tid) * 2, acc_s);
        atomicAdd(stats + ((size_t)b * COUT + tid) * 2 + 1, acc_ss);
    }
}

// ------------------------------------------------------------------------------------------------
// Block output: out = leaky_relu(IN(z_intra)) + leaky_relu(IN(z_skip))   (z_skip == nullptr: skip branch is exactly 0)
__global__ void __launch_bounds__(256) so3_combine_kernel(const float* __restrict__ za, const double* __restrict__ sa,
                                                          const float* __restrict__ zb, const double* __restrict__ sb,
                                                          int C, size_t per_batch, double count, float* __restrict__ out) {
    __shared__ float ma[64], ra[64], mb[64], rb[64];
    const int b = blockIdx.y;
    stats_to_affine(sa, b, C, count, ma, ra);
    if (zb) stats_to_affine(sb, b, C, count, mb, rb);
    __syncthreads();
    const size_t n4 = per_batch / 4;
    const float4* A = reinterpret_cast<const float4*>(za + (size_t)b * per_batch);
    const float4* Bp = zb ? reinterpret_cast<const float4*>(zb + (size_t)b * per_batch) : nullptr;
    float4* O = reinterpret_cast<float4*>(out + (size_t)b * per_batch);
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
        const int c = (int)((i * 4) % C);
        const float4 a = __ldg(A + i);
        float4 r;
        r.x = etch_lrelu((a.x - ma[c]) * ra[c]);
        r.y = etch_lrelu((a.y - ma[c + 1]) * ra[c + 1]);
        r.z = etch_lrelu((a.z - ma[c + 2]) * ra[c + 2]);
        r.w = etch_lrelu((a.w - ma[c + 3]) * ra[c + 3]);
        if (Bp) {
            const float4 v = __ldg(Bp + i);
            r.x += etch_lrelu((v.x - mb[c]) * rb[c]);
            r.y += etch_lrelu((v.y - mb[c + 1]) * rb[c + 1]);
            r.z += etch_lrelu((v.z - mb[c + 2]) * rb[c + 2]);
            r.w += etch_lrelu((v.w - mb[c + 3]) * rb[c + 3]);
        }
        O[i] = r;
    }
}

int grid_for(int ntiles) { return ntiles < 148 * 4 ? ntiles : 148 * 4; }

}  // namespace

// ================================================================================================ C ABI
// Fused InterSO3Conv for c_in == 1 with the constant occupancy feature (first encoder conv).
// Replaces InterSO3Conv.forward's weight/feature grouping + BasicSO3Conv (modules.py:120-128, functional.py:286-324,61-67)
ETCH_API int etch_so3_inter_conv_c1(const float* xyz, const int* sample_idx, const int* nbr, const float* kr,
                                    const float* Wt, const float* bias, int B, int q, int P, int nn, int cout,
                                    float sigma, float* zraw, double* stats, cudaStream_t stream) {
    if (!xyz || !sample_idx || !nbr || !kr || !Wt || !bias || !zraw || !stats || nn > 64 || nn <= 0) return ETCH_EINVAL;
    const float inv_sigma = 1.0f / sigma;
    dim3 grid(min(P, 148 * 8), B);
    if (cout == 32) inter_conv_c1_kernel<32><<<grid, C1_THREADS, 0, stream>>>(xyz, sample_idx, nbr, kr, Wt, bias, q, P, nn, inv_sigma, zraw, stats);
    else if (cout == 64) inter_conv_c1_kernel<64><<<grid, C1_THREADS, 0, stream>>>(xyz, sample_idx, nbr, kr, Wt, bias, q, P, nn, inv_sigma, zraw, stats);
    else return ETCH_EINVAL;
    ETCH_RETURN_LAST();
}

template <int CIN, int COUT, int NN>
static int launch_inter(const float* xyz, const float* feat, const int* sample_idx, const int* nbr, const float* krs,
                        const float* Wt, const float* bias, int B, int q, int P, float sigma, float* zraw, double* stats,
                        cudaStream_t stream) {
    constexpr size_t smem = (size_t)NA * NK * 16 + (size_t)TP * NN * 16 + (size_t)TP * NN * 4 +
                            (size_t)8 * NK * MROWS * 4 + (size_t)8 * NK * COUT * 4;
    auto kern = inter_conv_kernel<CIN, COUT, NN>;
    ETCH_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(grid_for((P + TP - 1) / TP), B);
    kern<<<grid, 256, smem, stream>>>(xyz, feat, sample_idx, nbr, reinterpret_cast<const float4*>(krs), Wt, bias, q, P,
                                      1.0f / sigma, zraw, stats);
    ETCH_RETURN_LAST();
}

// Fused InterSO3Conv (c_in in {32,64}). krs = [60,24,4] {2/sigma*R_a k, |R_a k|^2/sigma}; Wt = W^T [c_in*24][c_out].
ETCH_API int etch_so3_inter_conv(const float* xyz, const float* feat, const int* sample_idx, const int* nbr,
                                 const float* krs, const float* Wt, const float* bias, int B, int q, int P, int nn,
                                 int cin, int cout, float sigma, float* zraw, double* stats, cudaStream_t stream) {
    if (!xyz || !feat || !sample_idx || !nbr || !krs || !Wt || !bias || !zraw || !stats) return ETCH_EINVAL;
#define CASE(ci, co, n) \
    if (cin == ci && cout == co && nn == n) return launch_inter<ci, co, n>(xyz, feat, sample_idx, nbr, krs, Wt, bias, B, q, P, sigma, zraw, stats, stream);
    CASE(32, 32, 32) CASE(32, 64, 64) CASE(64, 64, 32) CASE(32, 32, 64) CASE(64, 64, 64) CASE(32, 64, 32)
#undef CASE
    return ETCH_EINVAL;
}

template <int CIN, int COUT, int J, bool NORM>
static int launch_agemm(const float* xin, const int* src_idx, const int* tab, const float* Wt, const float* bias,
                        const double* in_stats, double in_count, int B, int Q, int P, float* zraw, double* stats,
                        cudaStream_t stream) {
    constexpr size_t smem = (size_t)TP * NA * (CIN + 4) * 4 + (size_t)CIN * COUT * 4 + (size_t)MROWS * COUT * 4 +
                            (size_t)2 * CIN * 4 + (size_t)NA * J * 4;
    auto kern = anchor_gemm_kernel<CIN, COUT, J, NORM>;
    ETCH_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(grid_for((P + TP - 1) / TP), B);
    kern<<<grid, 256, smem, stream>>>(xin, src_idx, tab, Wt, bias, in_stats, in_count, Q, P, zraw, stats);
    ETCH_RETURN_LAST();
}

// IntraSO3Conv on the InstanceNorm+leaky_relu of zin (normalisation folded into the load). Wt = [12][c][c_out].
ETCH_API int etch_so3_intra_conv(const float* zin, const double* in_stats, const int* intra_idx, const float* Wt,
                                 const float* bias, int B, int P, int c, int cout, float* zraw, double* stats,
                                 cudaStream_t stream) {
    if (!zin || !in_stats || !intra_idx || !Wt || !bias || !zraw || !stats) return ETCH_EINVAL;
    const double cnt = (double)P * NA;
    if (c == 32 && cout == 32) return launch_agemm<32, 32, 12, true>(zin, nullptr, intra_idx, Wt, bias, in_stats, cnt, B, P, P, zraw, stats, stream);
    if (c == 64 && cout == 64) return launch_agemm<64, 64, 12, true>(zin, nullptr, intra_idx, Wt, bias, in_stats, cnt, B, P, P, zraw, stats, stream);
    return ETCH_EINVAL;
}

// skip branch: 1x1 Conv2d on feats[:, :, sample_idx] (src/models/so3conv.py:178-180). Wt = [c_in][c_out]; ident = arange(60).
ETCH_API int etch_so3_skip_conv(const float* feat, const int* sample_idx, const int* ident, const float* Wt,
                                const float* bias, int B, int q, int P, int cin, int cout, float* zraw, double* stats,
                                cudaStream_t stream) {
    if (!feat || !ident || !Wt || !bias || !zraw || !stats) return ETCH_EINVAL;
    if (cin == 32 && cout == 32) return launch_agemm<32, 32, 1, false>(feat, sample_idx, ident, Wt, bias, nullptr, 1.0, B, q, P, zraw, stats, stream);
    if (cin == 32 && cout == 64) return launch_agemm<32, 64, 1, false>(feat, sample_idx, ident, Wt, bias, nullptr, 1.0, B, q, P, zraw, stats, stream);
    if (cin == 64 && cout == 64) return launch_agemm<64, 64, 1, false>(feat, sample_idx, ident, Wt, bias, nullptr, 1.0, B, q, P, zraw, stats, stream);
    return ETCH_EINVAL;
}

// out = lrelu(IN(z_intra)) + lrelu(IN(z_skip)); z_skip may be NULL (first conv: the skip input is constant => IN gives 0)
ETCH_API int etch_so3_combine(const float* z_intra, const double* s_intra, const float* z_skip, const double* s_skip,
                              int B, int P, int c, float* out, cudaStream_t stream) {
    if (!z_intra || !s_intra || !out || c > 64 || (c % 4)) return ETCH_EINVAL;
    const size_t per_batch = (size_t)P * NA * c;
    dim3 grid((unsigned)min((size_t)148 * 4, (per_batch / 4 + 255) / 256), B);
    so3_combine_kernel<<<grid, 256, 0, stream>>>(z_intra, s_intra, z_skip, s_skip, c, per_batch, (double)P * NA, out);
    ETCH_RETURN_LAST();
}
