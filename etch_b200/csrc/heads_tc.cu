// Tensor-core (tcgen05 + TMEM) confidence head.
//
//   conf[m] = sum_l softmax(cls logits[m])_l * ( b2[l] + sum_u w2[l][u] * relu(b0[l*128+u] + x[m] . W0[l*128+u]) )
//   (src/models/pointtransformer_seg.py:145,184-189: Conv1d(128 -> 128*K) + ReLU + grouped Conv1d(128*K -> K, groups=K),
//    weighted by softmax(cls)).  The [B, 11008, N] activation is never materialised.
//
// A 128-point tile of x (TF32 hi/lo split, canonical K-major layout) stays in shared memory; the 11008 x 128 weight matrix
// streams through a 2-deep ring of 32-column slices (cp.async.bulk on mbarriers, pre-split / pre-tiled on the host); one
// thread issues the 3xTF32 tcgen05.mma's into a double-buffered 32-column TMEM accumulator; four epilogue warps drain
// the accumulators (tcgen05.ld), apply bias + ReLU + the grouped-conv weights and fold each finished marker group
// into the softmax-weighted confidence -- MMA, weight streaming and epilogue overlap.
#include "common.cuh"
#include "umma.cuh"

namespace {

constexpr int CT_N = 128;   // weight rows per marker group (GEMM N)
constexpr int CT_KQ = 32;   // K columns per streamed slice (a group = 4 slices)
constexpr int CT_K = 128;   // feature dimension

__global__ void __launch_bounds__(256, 1) conf_head_tc_kernel(const float* __restrict__ x,       // [n][128]
                                                              const float* __restrict__ logits,  // [n][K]
                                                              const float* __restrict__ W0c,     // [K*4][2][8][128][4]
                                                              const float* __restrict__ b0,      // [K*128]
                                                              const float* __restrict__ w2,      // [K][128]
                                                              const float* __restrict__ b2,      // [K]
                                                              int n, int K, float* __restrict__ conf) {
    constexpr uint32_t A_BYTES = 128 * CT_K * 4;      // 64 KB per (hi|lo)
    constexpr uint32_t B_BYTES = CT_N * CT_KQ * 4;    // 16 KB per (hi|lo) slice [128 rows x 32 k]
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* s_A = smem_raw;                    // [hi | lo]
    unsigned char* s_B = s_A + 2 * A_BYTES;           // [2 buffers][hi | lo]
    __shared__ uint64_t b_full[2], b_empty[2], t_full[2], t_empty[2];
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) umma::tmem_alloc(&tmem_base, 256);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            umma::mbar_init(&b_full[i], 1); umma::mbar_init(&b_empty[i], 1);
            umma::mbar_init(&t_full[i], 1); umma::mbar_init(&t_empty[i], 128);
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = umma::uniform(tmem_base);
    const int nslice = K * 4;
    uint32_t gs = 0;  // global slice counter, gg = global group counter (barrier phases run across tiles)
    uint32_t gg = 0;

    for (int tile = blockIdx.x; tile * 128 < n; tile += gridDim.x) {
        const int m0 = tile * 128;
        {   // A tile: 2 threads per row, TF32 split, canonical layout
            const int r = tid >> 1, half = tid & 1;
            const bool ok = m0 + r < n;
            const float4* src = reinterpret_cast<const float4*>(x + (size_t)(ok ? m0 + r : 0) * CT_K + half * 64);
#pragma unroll 4
            for (int c4 = 0; c4 < 16; ++c4) {
                float4 v = ok ? __ldg(src + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
                float4 h, l;
                umma::split_tf32(v.x, h.x, l.x); umma::split_tf32(v.y, h.y, l.y);
                umma::split_tf32(v.z, h.z, l.z); umma::split_tf32(v.w, h.w, l.w);
                const int kc = half * 16 + c4;
                *reinterpret_cast<float4*>(s_A + kc * (128 * 16) + r * 16) = h;
                *reinterpret_cast<float4*>(s_A + A_BYTES + kc * (128 * 16) + r * 16) = l;
            }
        }
        umma::fence_async_smem();
        __syncthreads();

        if (warp == 0) {
            // ===== weight streaming + MMA issue (warp-collective) =====
            const uint32_t a_hi = umma::smem_u32(s_A), a_lo = a_hi + A_BYTES;
            {
                const uint32_t g = gs;
                if (g >= 2) umma::mbar_wait(&b_empty[g & 1], ((g - 2) >> 1) & 1);
                umma::bulk_load(s_B + (g & 1) * 2 * B_BYTES, W0c, 2 * B_BYTES, &b_full[g & 1]);
            }
            for (int i = 0; i < nslice; ++i) {
                const uint32_t g = gs + i, buf = g & 1;
                const int kq = i & 3;
                const uint32_t grp = gg + (i >> 2), acc = grp & 1;
                if (i + 1 < nslice) {
                    const uint32_t g1 = g + 1, nb = g1 & 1;
                    if (g1 >= 2) umma::mbar_wait(&b_empty[nb], ((g1 - 2) >> 1) & 1);
                    umma::bulk_load(s_B + nb * 2 * B_BYTES, W0c + (size_t)(i + 1) * 2 * CT_N * CT_KQ, 2 * B_BYTES, &b_full[nb]);
                }
                umma::mbar_wait(&b_full[buf], (g >> 1) & 1);
                if (kq == 0 && grp >= 2) umma::mbar_wait(&t_empty[acc], ((grp - 2) >> 1) & 1);
                umma::fence_after_sync();
                const uint32_t b_hi = umma::smem_u32(s_B + buf * 2 * B_BYTES), b_lo = b_hi + B_BYTES;
                // K quarter kq of the canonical A tile starts 8 k-chunks (8 * 2048 bytes) further per quarter
                umma::issue_gemm_3xtf32(tmem + acc * CT_N, a_hi + kq * 8 * 2048, a_lo + kq * 8 * 2048, b_hi, b_lo, CT_KQ, CT_N, kq > 0);
                umma::commit(&b_empty[buf]);
                if (kq == 3) umma::commit(&t_full[acc]);
            }
        } else if (warp >= 4) {
            // ===== epilogue: one thread per point, one marker group (128 columns) per accumulator =====
            const int q = warp & 3;
            const int row = q * 32 + lane;
            const int m = m0 + row;
            const bool ok = m < n;
            const float* lr = logits + (size_t)(ok ? m : 0) * K;
            float mx = -INFINITY;
            for (int l = 0; l < K; ++l) mx = fmaxf(mx, __ldg(lr + l));
            float den = 0.f;
            for (int l = 0; l < K; ++l) den += expf(__ldg(lr + l) - mx);
            const float inv_den = 1.0f / den;
            float acc_conf = 0.f;
            for (int l = 0; l < K; ++l) {
                const uint32_t grp = gg + l, acc = grp & 1;
                umma::mbar_wait(&t_full[acc], (grp >> 1) & 1);
                umma::fence_after_sync();
                float part = 0.f;
#pragma unroll 1
                for (int c0 = 0; c0 < CT_N; c0 += 32) {
                    float v[32];
                    umma::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + acc * CT_N + c0, v);
                    const float4* bp = reinterpret_cast<const float4*>(b0 + (size_t)l * CT_N + c0);
                    const float4* wp = reinterpret_cast<const float4*>(w2 + (size_t)l * CT_N + c0);
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 bb = __ldg(bp + j4), ww = __ldg(wp + j4);
                        part = fmaf(fmaxf(v[j4 * 4 + 0] + bb.x, 0.f), ww.x, part);
                        part = fmaf(fmaxf(v[j4 * 4 + 1] + bb.y, 0.f), ww.y, part);
                        part = fmaf(fmaxf(v[j4 * 4 + 2] + bb.z, 0.f), ww.z, part);
                        part = fmaf(fmaxf(v[j4 * 4 + 3] + bb.w, 0.f), ww.w, part);
                    }
                }
                umma::fence_before_sync();
                umma::mbar_arrive(&t_empty[acc]);
                const float p = expf(__ldg(lr + l) - mx) * inv_den;
                acc_conf = fmaf(p, part + __ldg(b2 + l), acc_conf);
            }
            if (ok) conf[m] = acc_conf;
        }
        gs += nslice;
        gg += K;
        umma::fence_before_sync();
        __syncthreads();   // A tile and accumulators are free for the next tile
        umma::fence_after_sync();
    }
    if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

}  // namespace

// Tensor-core confidence head. W0c = [K*4][2][8][128][4]: per marker group l and K-quarter kq the [128 rows x 32 k] slice of
// confi.0's weight as (hi, lo) canonical tiles (etch_b200/models/heads.py).
ETCH_API int etch_conf_head_tc(const float* x, const float* logits, const float* W0c, const float* b0, const float* w2,
                               const float* b2, int n, int K, float* conf, cudaStream_t stream) {
    if (!x || !logits || !W0c || !b0 || !w2 || !b2 || !conf || n <= 0 || K <= 0) return ETCH_EINVAL;
    const size_t smem = (size_t)2 * 128 * CT_K * 4 + (size_t)4 * CT_N * CT_KQ * 4 + 128;
    ETCH_TRY(cudaFuncSetAttribute(conf_head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = etch_cdiv(n, 128);
    if (grid > etch_sm_budget()) grid = etch_sm_budget();
    conf_head_tc_kernel<<<grid, 256, smem, stream>>>(x, logits, W0c, b0, w2, b2, n, K, conf);
    ETCH_RETURN_LAST();
}

// =====================================================================================================================
// Tensor-core direction head
// =====================================================================================================================
// Same math as direction_head_kernel (heads.cu): 3-NN blend of the coarse equivariant features -> 2 x MHSA over the 60
// anchor tokens -> fused (head_combine o Linear1)+ReLU -> fused (Linear2 o so3_reg) -> chordal SO(3) mean -> direction.
// Reference: src/models/models_pointcloud.py:111-126,181-184; direction_backbones.py:79-223; src/models/so3conv.py:186-225.
//
// One CTA = 2 points = 120 tokens (UMMA M = 128).  All projections run on tcgen05 (3xTF32, accumulators in TMEM):
//   D[0:192]   = X  [Wq|Wk|Wv]^T      (Q stays in TMEM and is read per (token, head) with tcgen05.ld;
//                                      K, V go to a token-major shared tile for the 60x60 attention)
//   D[192:256] = O  Wc^T              (layer 1: + bias + residual -> X, re-split in place)
//   D[256:384] = O  (W1 Wc2)^T        (layer 2 output folded with Linear1; ReLU and the 128->1 map are applied straight
//                                      from TMEM, the hidden layer is never stored)
// Activations live in shared memory as (hi, lo) TF32 pairs in the canonical UMMA layout (x == hi + lo exactly), weights
// stream as 18 pre-split [64 rows x 32 k] slices per tile (cp.async.bulk) into two dedicated slots plus storage that is dead at that
// point of the tile (DhSlots below).  The softmax attention itself is block-diagonal (60x60 per point and head) and stays on the
// CUDA cores (packed fma.rn.f32x2).  Overlaps inside a tile: the K and V blocks of a QKV projection are issued before the Q block,
// so K|V move to shared memory under the Q MMAs; the NEXT tile's tokens are blended into X (dead once the layer-2 QKV MMAs have
// completed) while the last two MMA blocks of the current tile run; the next tile's QKV slices travel under the epilogue.
namespace {

constexpr int DH_NA = 60;
constexpr int DH_CHUNK = 60;                       // keys per online-softmax chunk (must divide 60)
constexpr int DH_LDKV = 132;                       // floats per row of the K|V tile
constexpr uint32_t DH_XB = 128 * 64 * 4;           // bytes of one canonical [128 x 64] tile
constexpr uint32_t DH_WB = 64 * 32 * 4;            // bytes of one (hi or lo) weight slice [64 rows x 32 k]

__device__ void dh_jacobi3(double A[3][3], double V[3][3], double e[3]) {
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
                for (int k = 0; k < 3; ++k) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
                for (int k = 0; k < 3; ++k) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
                for (int k = 0; k < 3; ++k) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
            }
    }
    e[0] = A[0][0]; e[1] = A[1][1]; e[2] = A[2][2];
}

// third column of U diag(1,1,det(UV^T)) V^T for Ce = U S V^T (see heads.cu::so3_direction)
__device__ void dh_so3_direction(const double Ce[3][3], float out[3]) {
    double A[3][3], V[3][3], e[3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A[i][j] = Ce[0][i] * Ce[0][j] + Ce[1][i] * Ce[1][j] + Ce[2][i] * Ce[2][j];
    dh_jacobi3(A, V, e);
    int i0 = 0;
    if (e[1] > e[i0]) i0 = 1;
    if (e[2] > e[i0]) i0 = 2;
    int i1 = (i0 + 1) % 3, i2 = (i0 + 2) % 3;
    if (e[i2] > e[i1]) { const int t = i1; i1 = i2; i2 = t; }
    const double v0[3] = {V[0][i0], V[1][i0], V[2][i0]}, v1[3] = {V[0][i1], V[1][i1], V[2][i1]};
    double u0[3], u1[3];
    for (int i = 0; i < 3; ++i) {
        u0[i] = Ce[i][0] * v0[0] + Ce[i][1] * v0[1] + Ce[i][2] * v0[2];
        u1[i] = Ce[i][0] * v1[0] + Ce[i][1] * v1[1] + Ce[i][2] * v1[2];
    }
    const double n0 = sqrt(u0[0] * u0[0] + u0[1] * u0[1] + u0[2] * u0[2]) + 1e-300;
    for (int i = 0; i < 3; ++i) u0[i] /= n0;
    const double d01 = u0[0] * u1[0] + u0[1] * u1[1] + u0[2] * u1[2];
    for (int i = 0; i < 3; ++i) u1[i] -= d01 * u0[i];
    const double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]) + 1e-300;
    for (int i = 0; i < 3; ++i) u1[i] /= n1;
    const double u2[3] = {u0[1] * u1[2] - u0[2] * u1[1], u0[2] * u1[0] - u0[0] * u1[2], u0[0] * u1[1] - u0[1] * u1[0]};
    const double v2z = v0[0] * v1[1] - v0[1] * v1[0];
    for (int i = 0; i < 3; ++i) out[i] = (float)(u0[i] * v0[2] + u1[i] * v1[2] + u2[i] * v2z);
}

// Weight streaming of the MMA-issuing warp.  A tile consumes 18 [64 rows x 32 k] (hi|lo) slices in a fixed order (n = 0..17: K, V, Q
// blocks of layer 1, head-combine, K, V, Q of layer 2, the two blocks of the fused MLP).  Shared memory is full, so besides the two
// dedicated slots B0, B1 the slices land in storage that is dead at that point of the tile: the attention-output tile O (4 slots: dead
// from the end of the MLP MMAs until the next attention writes it) and the K|V tile (3 slots: dead between an attention and the next
// K|V copy).  Every load is issued at a program point where its slot is known to be free (after the mbarrier that tracks the MMAs which
// read it), so there are no "empty" barriers; slot s signals arrival on full[s], whose parity follows from the number of uses per tile.
//   n      : 0  1  2  3  4  5 | 6  7 | 8   9   10  11 12 13 | 14 15 16  17
//   slot   : O0 O1 O2 O3 B0 B1| B0 B1| KV0 KV1 KV2 B0 B1 O0 | B0 B1 KV0 KV1
struct DhSlots {
    const float* wall;      // [18][2][8][64][4]: 9 blocks of 64 output rows x 2 K-halves, (hi, lo) canonical tiles
    unsigned char *s_B, *s_O, *s_KV;
    uint64_t* full;         // [9]: B0 B1 O0 O1 O2 O3 KV0 KV1 KV2
    __device__ unsigned char* addr(int slot) const {
        return slot < 2 ? s_B + slot * 2 * DH_WB : slot < 6 ? s_O + (slot - 2) * 2 * DH_WB : s_KV + (slot - 6) * 2 * DH_WB;
    }
    // consumption index -> slice of `wall` (the K and V blocks of a QKV projection are consumed before its Q block)
    static __device__ int phys(int n) { return n < 6 ? (n < 4 ? n + 2 : n - 4) : (n >= 8 && n < 14 ? (n < 12 ? n + 2 : n - 4) : n); }
    __device__ void load(int n, int slot) const {
        umma::bulk_load(addr(slot), wall + (size_t)phys(n) * 2 * 64 * 32, 2 * DH_WB, &full[slot]);
    }
    // K-half kh of a 64-column output block: A = K-half kh of the canonical [128 x 64] tile ((32/4) k-chunks = 8 * 2048 bytes in)
    __device__ void mma(int slot, uint32_t parity, uint32_t a_hi, uint32_t a_lo, uint32_t tmem_d, int kh) const {
        umma::mbar_wait(&full[slot], parity);
        umma::fence_after_sync();
        const uint32_t b_hi = umma::smem_u32(addr(slot)), b_lo = b_hi + DH_WB;
        umma::issue_gemm_3xtf32(tmem_d, a_hi + kh * 8 * 2048, a_lo + kh * 8 * 2048, b_hi, b_lo, 32, 64, kh > 0);
    }
};
enum { SB0 = 0, SB1 = 1, SO0 = 2, SO1 = 3, SO2 = 4, SO3 = 5, SKV0 = 6, SKV1 = 7, SKV2 = 8 };

// K (warps 0-3) / V (warps 4-7) columns of the QKV accumulator -> token-major shared tile
__device__ __forceinline__ void dh_store_kv(uint32_t tmem, float* s_kv, int warp, int lane) {
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
        float v[32];
        umma::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 64 + half * 64 + c0, v);
        if (row < 2 * DH_NA) {
            float* dst = s_kv + row * DH_LDKV + half * 64 + c0;
#pragma unroll
            for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
    }
}

__device__ __forceinline__ float dh_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ uint64_t dh_pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void dh_unpk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t dh_fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t dh_mul2(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// softmax(q K^T) V for (token = my TMEM lane, my 4 heads); output written as (hi, lo) into the canonical O tile.
// Both products run as packed fma.rn.f32x2 (FFMA2: two lanes of the 8-wide head per instruction, the probability as a broadcast
// operand): 17 instead of 24 issued instructions per key, the FP32 pipe time (16 lane-FMAs per key) is unchanged -- round 2 A/B
// 5.09 -> 4.85 ms for the kernel.  The scores of a head are summed as (d0+d2+d4+d6) + (d1+d3+d5+d7).  Chunks of 20 / 30 / 60 keys
// measure the same within noise (4.87 / 4.88 / 4.85 ms); one head's 60 keys are unrolled, the heads are not (the 4 x 60-key body of
// the first version thrashed the instruction cache).
__device__ __forceinline__ void dh_attention(uint32_t tmem, const float* s_kv, unsigned char* s_O, int warp, int lane) {
    const int q = warp & 3, hq = (warp >> 2) * 4;
    const int row = q * 32 + lane;
    const bool valid = row < 2 * DH_NA;
    const int base = (valid && row >= DH_NA) ? DH_NA : 0;
#pragma unroll 1
    for (int hh = 0; hh < 4; ++hh) {
        const int h = hq + hh;
        float qv[8];
        umma::tmem_ld8(tmem + ((uint32_t)(q * 32) << 16) + h * 8, qv);   // warp-collective: every lane takes part
#pragma unroll
        for (int d = 0; d < 8; ++d) qv[d] *= 1.4426950408889634f;         // scores in log2 units: every exponential is one MUFU.EX2
        float mx = -INFINITY, sum = 0.f;
        const uint64_t q01 = dh_pk(qv[0], qv[1]), q23 = dh_pk(qv[2], qv[3]), q45 = dh_pk(qv[4], qv[5]), q67 = dh_pk(qv[6], qv[7]);
        uint64_t o2[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll 1
        for (int j0 = 0; j0 < DH_NA; j0 += DH_CHUNK) {
            float s[DH_CHUNK];
            float cm = -INFINITY;
#pragma unroll
            for (int j = 0; j < DH_CHUNK; ++j) {
                const float* kr = s_kv + (base + j0 + j) * DH_LDKV + h * 8;
                const ulonglong2 ka = *reinterpret_cast<const ulonglong2*>(kr);
                const ulonglong2 kb = *reinterpret_cast<const ulonglong2*>(kr + 4);
                uint64_t acc = dh_mul2(q01, ka.x);
                acc = dh_fma2(q23, ka.y, acc); acc = dh_fma2(q45, kb.x, acc); acc = dh_fma2(q67, kb.y, acc);
                float lo, hi;
                dh_unpk(acc, lo, hi);
                const float v = lo + hi;
                s[j] = v;
                cm = fmaxf(cm, v);
            }
            const float mn = fmaxf(mx, cm);
            const float sc = dh_ex2(mx - mn);   // first chunk: 2^(-inf) = 0
            mx = mn;
            sum *= sc;
            const uint64_t sc2 = dh_pk(sc, sc);
#pragma unroll
            for (int d = 0; d < 4; ++d) o2[d] = dh_mul2(o2[d], sc2);
#pragma unroll
            for (int j = 0; j < DH_CHUNK; ++j) {
                const float p = dh_ex2(s[j] - mn);
                sum += p;
                const uint64_t p2 = dh_pk(p, p);
                const float* vr = s_kv + (base + j0 + j) * DH_LDKV + 64 + h * 8;
                const ulonglong2 va = *reinterpret_cast<const ulonglong2*>(vr);
                const ulonglong2 vb = *reinterpret_cast<const ulonglong2*>(vr + 4);
                o2[0] = dh_fma2(p2, va.x, o2[0]); o2[1] = dh_fma2(p2, va.y, o2[1]);
                o2[2] = dh_fma2(p2, vb.x, o2[2]); o2[3] = dh_fma2(p2, vb.y, o2[3]);
            }
        }
        float o[8];
#pragma unroll
        for (int d = 0; d < 4; ++d) dh_unpk(o2[d], o[2 * d], o[2 * d + 1]);
        if (valid) {
            const float inv = 1.0f / sum;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                float4 hi, lo;
                umma::split_tf32(o[g * 4 + 0] * inv, hi.x, lo.x); umma::split_tf32(o[g * 4 + 1] * inv, hi.y, lo.y);
                umma::split_tf32(o[g * 4 + 2] * inv, hi.z, lo.z); umma::split_tf32(o[g * 4 + 3] * inv, hi.w, lo.w);
                const int kc = h * 2 + g;
                *reinterpret_cast<float4*>(s_O + kc * (128 * 16) + row * 16) = hi;
                *reinterpret_cast<float4*>(s_O + DH_XB + kc * (128 * 16) + row * 16) = lo;
            }
        }
    }
}

__global__ void __launch_bounds__(256, 1) direction_head_tc_kernel(
    const float* __restrict__ feats,   // [B,S,60,64]
    const int* __restrict__ up_idx, const float* __restrict__ up_w,   // [B,N,3]
    const float* __restrict__ wall,    // [18][2][8][64][4] weight slices
    const float* __restrict__ bc1,     // [64]
    const float* __restrict__ bf,      // [128]
    const float* __restrict__ vreg,    // [128]
    float creg, const float* __restrict__ anchors, int nscans, int N, int S,
    double* __restrict__ ce_out, float* __restrict__ anc_w)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* s_X = smem_raw;                         // [hi|lo] tokens
    unsigned char* s_O = s_X + 2 * DH_XB;                  // [hi|lo] attention output
    unsigned char* s_B = s_O + 2 * DH_XB;                  // [2][hi|lo] weight ring
    float* s_kv = reinterpret_cast<float*>(s_B + 4 * DH_WB);   // [120][132]
    float* s_part = s_kv + 2 * DH_NA * DH_LDKV;            // [2][128]
    float* s_w = s_part + 256;                             // [128]
    float* s_anc = s_w + 128;                              // [60][9]
    __shared__ __align__(16) float s_bc1[64], s_bf[128], s_vreg[128];
    __shared__ uint64_t s_full[9], bar_mma, bar_kv;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < DH_NA * 9; i += 256) s_anc[i] = __ldg(anchors + i);
    if (tid < 64) s_bc1[tid] = __ldg(bc1 + tid);
    if (tid < 128) { s_bf[tid] = __ldg(bf + tid); s_vreg[tid] = __ldg(vreg + tid); }
    if (warp == 0) umma::tmem_alloc(&tmem_base, 512);
    if (tid == 0) {
        for (int i = 0; i < 9; ++i) umma::mbar_init(&s_full[i], 1);
        umma::mbar_init(&bar_mma, 1); umma::mbar_init(&bar_kv, 1);
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = umma::uniform(tmem_base);
    const uint32_t x_hi = umma::smem_u32(s_X), x_lo = x_hi + DH_XB, o_hi = umma::smem_u32(s_O), o_lo = o_hi + DH_XB;
    const DhSlots iss{wall, s_B, s_O, reinterpret_cast<unsigned char*>(s_kv), s_full};
    static_assert(3 * 2 * DH_WB <= 2 * DH_NA * DH_LDKV * 4, "three slice slots fit the K|V tile");
    uint32_t n_mma = 0, n_kv = 0;
    const int ntiles = (N + 1) / 2;
    const int q = warp & 3, half = warp >> 2, row = q * 32 + lane;

    // Blend of a tile's tokens (3 coarse rows per token, interpolation weights from the 3-NN search) into X as (hi, lo).  The 12 loads
    // of a thread are issued back to back (one L2 round trip instead of three, the neighbour ids and weights were fetched a tile ahead)
    // and consumed later: the ncu source view had 17-22 % of the kernel's stall samples on these loads (LSU queue throttle).
    const int br = tid & 127, bhf = tid >> 7;             // token row, channel half
    const bool brow_ok = br < 2 * DH_NA;
    const int bpl = brow_ok ? br / DH_NA : 0, ba = brow_ok ? br % DH_NA : 0;
    int nb_idx[3] = {0, 0, 0};
    float nb_w[3] = {0.f, 0.f, 0.f};
    float4 bv[3][8];
    auto blend_ids = [&](int g) {
        const int b = g / ntiles, tile = g - b * ntiles;
        const int p = min(tile * 2 + bpl, N - 1);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            nb_idx[k] = __ldg(up_idx + ((size_t)b * N + p) * 3 + k);
            nb_w[k] = __ldg(up_w + ((size_t)b * N + p) * 3 + k);
        }
    };
    auto blend_load = [&](int g) {
        const float* F = feats + (size_t)(g / ntiles) * S * DH_NA * 64;
        if (brow_ok) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float* src = F + ((size_t)nb_idx[k] * DH_NA + ba) * 64 + bhf * 32;
                // every 32-byte sector of the coarse features is read exactly once per tile: 256-bit loads that do not allocate in L1
                // (A/B: 128-bit __ldg 4.50 ms, 256-bit 4.32 ms, 256-bit no-allocate 4.17 ms; coalescing the rows across lanes changed
                // nothing: the LSU queue counts instructions)
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                    float t8[8];
                    etch_ld256_na(src + i * 4, t8);
                    bv[k][i] = make_float4(t8[0], t8[1], t8[2], t8[3]);
                    bv[k][i + 1] = make_float4(t8[4], t8[5], t8[6], t8[7]);
                }
            }
        }
    };
    auto blend_store = [&]() {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (brow_ok) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    acc.x = fmaf(bv[k][i].x, nb_w[k], acc.x); acc.y = fmaf(bv[k][i].y, nb_w[k], acc.y);
                    acc.z = fmaf(bv[k][i].z, nb_w[k], acc.z); acc.w = fmaf(bv[k][i].w, nb_w[k], acc.w);
                }
            }
            float4 hi, lo;
            umma::split_tf32(acc.x, hi.x, lo.x); umma::split_tf32(acc.y, hi.y, lo.y);
            umma::split_tf32(acc.z, hi.z, lo.z); umma::split_tf32(acc.w, hi.w, lo.w);
            const int kc = bhf * 8 + i;
            *reinterpret_cast<float4*>(s_X + kc * (128 * 16) + br * 16) = hi;
            *reinterpret_cast<float4*>(s_X + DH_XB + kc * (128 * 16) + br * 16) = lo;
        }
    };
    const int total_tiles = ntiles * nscans;
    if ((int)blockIdx.x < total_tiles) {
        if (warp == 0) {   // slices of the first tile's QKV projection
            iss.load(0, SO0); iss.load(1, SO1); iss.load(2, SO2); iss.load(3, SO3); iss.load(4, SB0); iss.load(5, SB1);
        }
        blend_ids(blockIdx.x);
        blend_load(blockIdx.x);
        blend_store();
    }
    uint32_t tp = 0;   // parity of this CTA's tile counter (slots used once per tile)

    // scan-major tile sequence: the whole grid blends from one scan's coarse features at a time (19 MB, L2 resident) instead of
    // all B of them (the scan-parallel grid re-read them 4x from DRAM)
    for (int gt = blockIdx.x; gt < total_tiles; gt += gridDim.x) {
        const int b = gt / ntiles, tile = gt - b * ntiles;
        const int p0 = tile * 2;
        const bool has_next = gt + (int)gridDim.x < total_tiles;
        if (has_next) blend_ids(gt + gridDim.x);   // neighbour ids / weights of the next tile: in registers long before its blend
        for (int layer = 0; layer < 2; ++layer) {
            umma::fence_async_smem();
            __syncthreads();
            // ---- QKV projection: 6 slices -> D[0:192]; K and V first, they are copied to shared memory under the Q MMAs ----
            if (warp == 0) {
                umma::fence_after_sync();
                if (layer == 0) {
                    iss.mma(SO0, 0, x_hi, x_lo, tmem + 64, 0);   iss.mma(SO1, tp, x_hi, x_lo, tmem + 64, 1);
                    iss.mma(SO2, tp, x_hi, x_lo, tmem + 128, 0); iss.mma(SO3, tp, x_hi, x_lo, tmem + 128, 1);
                    umma::commit(&bar_kv);
                    iss.mma(SB0, 0, x_hi, x_lo, tmem, 0);        iss.mma(SB1, 0, x_hi, x_lo, tmem, 1);
                } else {
                    iss.mma(SKV0, 0, x_hi, x_lo, tmem + 64, 0);   iss.mma(SKV1, 0, x_hi, x_lo, tmem + 64, 1);
                    iss.mma(SKV2, tp, x_hi, x_lo, tmem + 128, 0); iss.mma(SB0, 0, x_hi, x_lo, tmem + 128, 1);
                    umma::commit(&bar_kv);
                    iss.mma(SB1, 0, x_hi, x_lo, tmem, 0);         iss.mma(SO0, 1, x_hi, x_lo, tmem, 1);
                }
                umma::commit(&bar_mma);
            }
            umma::mbar_wait(&bar_kv, n_kv & 1); ++n_kv;
            umma::fence_after_sync();
            dh_store_kv(tmem, s_kv, warp, lane);
            umma::mbar_wait(&bar_mma, n_mma & 1); ++n_mma;
            umma::fence_after_sync();
            if (warp == 0) {   // B0 / B1 are free: head-combine slices (layer 1) or the first MLP block (layer 2)
                if (layer == 0) { iss.load(6, SB0); iss.load(7, SB1); }
                else { iss.load(14, SB0); iss.load(15, SB1); }
            }
            __syncthreads();
            dh_attention(tmem, s_kv, s_O, warp, lane);
            umma::fence_before_sync();
            umma::fence_async_smem();
            __syncthreads();
            if (layer == 0) {
                // ---- head_combine + bias + residual -> X ----
                if (warp == 0) {
                    umma::fence_after_sync();
                    // the K|V tile is dead until the next K|V copy: it takes the first three slices of layer 2's projection
                    iss.load(8, SKV0); iss.load(9, SKV1); iss.load(10, SKV2);
                    iss.mma(SB0, 1, o_hi, o_lo, tmem + 192, 0); iss.mma(SB1, 1, o_hi, o_lo, tmem + 192, 1);
                    umma::commit(&bar_mma);
                }
                umma::mbar_wait(&bar_mma, n_mma & 1); ++n_mma;
                umma::fence_after_sync();
                if (warp == 0) { iss.load(11, SB0); iss.load(12, SB1); iss.load(13, SO0); }   // O is dead until the next attention
                float v[32];
                umma::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 192 + half * 32, v);
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const int kc = (half * 32 + i) >> 2;
                    float4* ph = reinterpret_cast<float4*>(s_X + kc * (128 * 16) + row * 16);
                    float4* pl_ = reinterpret_cast<float4*>(s_X + DH_XB + kc * (128 * 16) + row * 16);
                    const float4 xh = *ph, xl = *pl_;
                    const int n0 = half * 32 + i;
                    const float nx = (xh.x + xl.x) + (v[i] + s_bc1[n0]), ny = (xh.y + xl.y) + (v[i + 1] + s_bc1[n0 + 1]);
                    const float nz = (xh.z + xl.z) + (v[i + 2] + s_bc1[n0 + 2]), nw = (xh.w + xl.w) + (v[i + 3] + s_bc1[n0 + 3]);
                    float4 hi, lo;
                    umma::split_tf32(nx, hi.x, lo.x); umma::split_tf32(ny, hi.y, lo.y);
                    umma::split_tf32(nz, hi.z, lo.z); umma::split_tf32(nw, hi.w, lo.w);
                    *ph = hi; *pl_ = lo;
                }
                umma::fence_before_sync();
            }
        }
        // ---- fused (Linear1 o head_combine_2) + ReLU, then (so3_reg o Linear2): D[256:384] -> anchor weights ----
        if (warp == 0) {
            umma::fence_after_sync();
            iss.load(16, SKV0); iss.load(17, SKV1);   // second MLP block: the K|V tile is dead again
            iss.mma(SB0, 1, o_hi, o_lo, tmem + 256, 0);  iss.mma(SB1, 1, o_hi, o_lo, tmem + 256, 1);
            iss.mma(SKV0, 1, o_hi, o_lo, tmem + 320, 0); iss.mma(SKV1, 1, o_hi, o_lo, tmem + 320, 1);
            umma::commit(&bar_mma);
        }
        // X is dead since the layer-2 QKV MMAs completed: blend the next tile's tokens while the last MMAs run
        if (has_next) { blend_load(gt + gridDim.x); blend_store(); }
        umma::mbar_wait(&bar_mma, n_mma & 1); ++n_mma;
        umma::fence_after_sync();
        if (has_next && warp == 0) {   // O, B0, B1 are free: the next tile's QKV slices arrive under this tile's epilogue
            iss.load(0, SO0); iss.load(1, SO1); iss.load(2, SO2); iss.load(3, SO3); iss.load(4, SB0); iss.load(5, SB1);
        }
        tp ^= 1;
        {
            float part = 0.f;
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 32) {
                float v[32];
                umma::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 256 + half * 64 + c0, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int n0 = half * 64 + c0 + i;
                    part = fmaf(fmaxf(v[i] + s_bf[n0], 0.f), s_vreg[n0], part);
                }
            }
            s_part[half * 128 + row] = part;
        }
        umma::fence_before_sync();
        __syncthreads();
        if (tid < 128) {
            const float w = creg + s_part[tid] + s_part[128 + tid];
            s_w[tid] = w;
            if (anc_w && tid < 2 * DH_NA && p0 + tid / DH_NA < N) anc_w[((size_t)b * N + p0) * DH_NA + tid] = w;
        }
        __syncthreads();
        // chordal-mean matrix Ce = sum_a w_a R_a of the tile's two points (18 entries, one thread each); the 3x3 SVD that
        // turns it into a direction runs in so3_direction_kernel, off this kernel's critical path
        if (tid < 18 && p0 + tid / 9 < N) {
            const int pl = tid / 9, e = tid % 9;
            double acc = 0.0;
            for (int a = 0; a < DH_NA; ++a) acc += (double)(s_w[pl * DH_NA + a] * s_anc[a * 9 + e]);
            ce_out[((size_t)b * N + p0 + pl) * 9 + e] = acc;
        }
        // the next tile's blend overwrites X only; s_w / s_part are rewritten after further barriers
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

// direction[pt] = third column of the rotation closest to Ce[pt] (src/models/models_pointcloud.py so3_mean): one thread per point
__global__ void __launch_bounds__(128) so3_direction_kernel(const double* __restrict__ ce, int total_pts, float* __restrict__ dir) {
    const int pt = blockIdx.x * 128 + threadIdx.x;
    if (pt >= total_pts) return;
    double Ce[3][3];
#pragma unroll
    for (int e = 0; e < 9; ++e) Ce[e / 3][e % 3] = ce[(size_t)pt * 9 + e];
    float d[3];
    dh_so3_direction(Ce, d);
    dir[(size_t)pt * 3] = d[0]; dir[(size_t)pt * 3 + 1] = d[1]; dir[(size_t)pt * 3 + 2] = d[2];
}

// inv[b,n,:] = sum_k w_k * mean_a feats[b, idx_k, a, :]   (anchor mean commutes with the 3-NN blend)
__global__ void __launch_bounds__(256) anchor_mean_kernel(const float* __restrict__ feats, int total_pts, float* __restrict__ fmean) {
    const int pt = blockIdx.x * 4 + (threadIdx.x >> 6), c = threadIdx.x & 63;
    if (pt >= total_pts) return;
    const float* p = feats + (size_t)pt * DH_NA * 64 + c;
    float s = 0.f;
#pragma unroll 4
    for (int a = 0; a < DH_NA; ++a) s += __ldg(p + a * 64);
    fmean[(size_t)pt * 64 + c] = s / 60.0f;
}

__global__ void __launch_bounds__(256) interp_inv_kernel(const float* __restrict__ fmean, const int* __restrict__ up_idx,
                                                         const float* __restrict__ up_w, int N, int S, float* __restrict__ inv) {
    const int b = blockIdx.y;
    const int pt = blockIdx.x * 4 + (threadIdx.x >> 6), c = threadIdx.x & 63;
    if (pt >= N) return;
    const int* ip = up_idx + ((size_t)b * N + pt) * 3;
    const float* wp = up_w + ((size_t)b * N + pt) * 3;
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) v = fmaf(__ldg(fmean + ((size_t)b * S + __ldg(ip + k)) * 64 + c), __ldg(wp + k), v);
    inv[((size_t)b * N + pt) * 64 + c] = v;
}

}  // namespace

// Tensor-core decode_direction. wall = the 18 weight slices [18][2][8][64][4] (see etch_b200/models/heads.py::DirectionPlan).
// scratch: caller-owned, B*S*64 + B*N*18 floats (per-coarse-point anchor mean, then the [B,N,9] double chordal-mean matrices).
ETCH_API int etch_direction_head_tc(const float* feats, const int* up_idx, const float* up_w, const float* wall,
                                    const float* bc1, const float* bf, const float* vreg, float creg, const float* anchors,
                                    int B, int N, int S, float* dir, float* inv, float* anc_w, float* scratch,
                                    cudaStream_t stream) {
    if (!feats || !up_idx || !up_w || !wall || !bc1 || !bf || !vreg || !anchors || !dir || !inv || !scratch) return ETCH_EINVAL;
    float* fmean_scratch = scratch;
    double* ce = reinterpret_cast<double*>(scratch + (size_t)B * S * 64);
    anchor_mean_kernel<<<etch_cdiv(B * S, 4), 256, 0, stream>>>(feats, B * S, fmean_scratch);
    dim3 g2(etch_cdiv(N, 4), B);
    interp_inv_kernel<<<g2, 256, 0, stream>>>(fmean_scratch, up_idx, up_w, N, S, inv);
    const size_t smem = (size_t)4 * DH_XB + (size_t)4 * DH_WB + (size_t)(2 * DH_NA * DH_LDKV + 256 + 128 + DH_NA * 9) * 4 + 128;
    ETCH_TRY(cudaFuncSetAttribute(direction_head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ntiles = (N + 1) / 2;
    int gx = etch_sm_budget();   // persistent CTAs, 1 per SM (smem-bound): never more than one wave
    if (gx > ntiles * B) gx = ntiles * B;
    direction_head_tc_kernel<<<gx, 256, smem, stream>>>(feats, up_idx, up_w, wall, bc1, bf, vreg, creg, anchors, B, N, S, ce, anc_w);
    so3_direction_kernel<<<etch_cdiv(B * N, 128), 128, 0, stream>>>(ce, B * N, dir);
    ETCH_RETURN_LAST();
}
