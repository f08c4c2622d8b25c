// PointTransformerLayer (vector attention) with the attention MLP on the tensor cores.
//
//   w_in[(i,j)][ch] = relu( BN( k[nb_j][ch] - q[i][ch] + p_r[(i,j)][ch] ) )          p_r = Linear(3,c) o relu o BN o Linear(3,3) (p[nb_j] - p[i])
//   logit[(i,j)][u] = W2 relu( BN( W1 w_in ) ) + b2                                   (c -> c/8 -> c/8)
//   out[i][ch]      = relu( bn2( sum_j softmax_j(logit)[u = ch mod c/8] * (v[nb_j][ch] + p_r[(i,j)][ch]) ) )
//   (src/models/pointtransformer_seg.py:8-37 and the bn2 + ReLU of the enclosing block, :114-122)
//
// The first linear layer is a dense GEMM over (point, neighbour) rows: [n*ns x c] x [c x c/8].  A tile = 128 rows =
// 128/ns points.  The CTA builds w_in for 64 channels at a time straight into the canonical UMMA A tile as (hi, lo) TF32
// pairs (the k rows are gathered from L2 with 128-bit loads, one row per lane, conflict-free stores), warp 0 streams the
// matching [c/8 x 64] slice of W1 (cp.async.bulk, 2-deep ring) and issues the 3xTF32 tcgen05.mma's into a TMEM accumulator.
// The c/8 x c/8 second layer, the softmax over the ns neighbours and the aggregation stay on the CUDA cores (the second
// layer is 1/8 of the first; the aggregation is an L2 gather).  The previous version (pt.cu::pt_attn_kernel) walked over
// points with one CTA each and reduced every dot product with warp shuffles: 6 ms per step against ~1 ms here.
#include "common.cuh"
#include "umma.cuh"

namespace {

struct AttnTcW {
    const float* P0;    // [3][3] Linear(3,3) with BN(3) folded
    const float* p0b;   // [3]
    const float* chan;  // [c][8]  {P3x, P3y, P3z, p3b, s0, h0, so, ho}
    const float* W1c;   // [c/64][2][16][TP][4]  (hi, lo) canonical tiles of W1[:, 64-channel chunk], rows padded to TP
    const float* b1;    // [T]
    const float* W2;    // [T][T]
    const float* b2;    // [T]
};

constexpr int PT_ROWS = 128;
constexpr int PT_KC = 64;                           // channels per K chunk
constexpr uint32_t PT_A_BYTES = PT_ROWS * PT_KC * 4;   // one (hi or lo) A chunk tile

template <int C>
struct PtCfg {
    static constexpr int T = C / 8;
    static constexpr int TP = T < 16 ? 16 : T;      // UMMA N (multiple of 16 for M = 128)
    static constexpr int NCH = C / PT_KC;
    static constexpr int LDL = T + 1;               // logits row stride (floats)
    static constexpr uint32_t W_BYTES = TP * PT_KC * 4;      // one (hi or lo) W1 chunk tile
    static constexpr size_t smem = 2 * PT_A_BYTES + 4 * W_BYTES + (size_t)C * 32 + (size_t)(T * T + 2 * T) * 4 +
                                   (size_t)PT_ROWS * LDL * 4 + PT_ROWS * 16 + PT_ROWS * 4 + 128;
    static constexpr int TCOLS = TP <= 32 ? 32 : 64;
};

template <int C, int NS>
__global__ void __launch_bounds__(256) pt_attn_tc_kernel(const float* __restrict__ p, const float* __restrict__ qkv,
                                                         const int* __restrict__ idx, AttnTcW W, int n,
                                                         float* __restrict__ out) {
    constexpr int ns = NS;
    using Cfg = PtCfg<C>;
    constexpr int T = Cfg::T, TP = Cfg::TP, NCH = Cfg::NCH, LDL = Cfg::LDL;
    constexpr uint32_t W_BYTES = Cfg::W_BYTES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* s_A = smem_raw;                                    // [hi | lo]
    unsigned char* s_W = s_A + 2 * PT_A_BYTES;                        // [2 slots][hi | lo]
    float* s_chan = reinterpret_cast<float*>(s_W + 4 * W_BYTES);      // [C][8]
    float* s_W2 = s_chan + C * 8;                                     // [T][T]
    float* s_b1 = s_W2 + T * T;
    float* s_b2 = s_b1 + T;
    float* s_lg = s_b2 + T;                                           // [128][LDL] logits -> softmax weights
    float* s_e = s_lg + PT_ROWS * LDL;                                // [128][4]
    int* s_nb = reinterpret_cast<int*>(s_e + PT_ROWS * 4);            // [128]
    __shared__ uint64_t w_full[2], bar_mma;
    __shared__ uint32_t tmem_base;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < C * 8; i += 256) s_chan[i] = __ldg(W.chan + i);
    for (int i = tid; i < T * T; i += 256) s_W2[i] = __ldg(W.W2 + i);
    for (int i = tid; i < T; i += 256) { s_b1[i] = __ldg(W.b1 + i); s_b2[i] = __ldg(W.b2 + i); }
    if (warp == 0) umma::tmem_alloc(&tmem_base, Cfg::TCOLS);
    if (tid == 0) { umma::mbar_init(&w_full[0], 1); umma::mbar_init(&w_full[1], 1); umma::mbar_init(&bar_mma, 1); }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = umma::uniform(tmem_base);

    const int ppt = PT_ROWS / ns;                       // points per tile
    const int ntiles = (n + ppt - 1) / ppt;
    const int row = tid & 127, half = tid >> 7;
    uint32_t n_mma = 0, n_w = 0;                        // MMA groups committed / W chunks requested so far (uniform)
    const float p00 = __ldg(W.P0), p01 = __ldg(W.P0 + 1), p02 = __ldg(W.P0 + 2), p10 = __ldg(W.P0 + 3), p11 = __ldg(W.P0 + 4),
                p12 = __ldg(W.P0 + 5), p20 = __ldg(W.P0 + 6), p21 = __ldg(W.P0 + 7), p22 = __ldg(W.P0 + 8);
    const float pb0 = __ldg(W.p0b), pb1 = __ldg(W.p0b + 1), pb2 = __ldg(W.p0b + 2);

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int pt0 = tile * ppt;
        // W1 chunks 0 (and 1) of this tile: both ring slots are free (every MMA of the previous tile has been waited for)
        if (warp == 0) {
            umma::bulk_load(s_W + (n_w & 1) * 2 * W_BYTES, W.W1c, 2 * W_BYTES, &w_full[n_w & 1]);
            if (NCH > 1) umma::bulk_load(s_W + ((n_w + 1) & 1) * 2 * W_BYTES, W.W1c + (size_t)2 * TP * PT_KC, 2 * W_BYTES, &w_full[(n_w + 1) & 1]);
        }
        // ---- A. per-row geometry: neighbour id, relu(P0 d + b) ----
        if (tid < PT_ROWS) {
            const int i = pt0 + tid / ns;
            int nb = 0;
            float e0 = 0.f, e1 = 0.f, e2 = 0.f;
            if (i < n) {
                nb = __ldg(idx + (size_t)i * ns + (tid % ns));
                const float dx = __ldg(p + (size_t)nb * 3) - __ldg(p + (size_t)i * 3), dy = __ldg(p + (size_t)nb * 3 + 1) - __ldg(p + (size_t)i * 3 + 1),
                            dz = __ldg(p + (size_t)nb * 3 + 2) - __ldg(p + (size_t)i * 3 + 2);
                e0 = fmaxf(fmaf(p02, dz, fmaf(p01, dy, fmaf(p00, dx, pb0))), 0.f);
                e1 = fmaxf(fmaf(p12, dz, fmaf(p11, dy, fmaf(p10, dx, pb1))), 0.f);
                e2 = fmaxf(fmaf(p22, dz, fmaf(p21, dy, fmaf(p20, dx, pb2))), 0.f);
            }
            s_nb[tid] = nb;
            *reinterpret_cast<float4*>(s_e + tid * 4) = make_float4(e0, e1, e2, 0.f);
        }
        __syncthreads();
        // ---- B. logits layer 1 on the tensor cores, 64 channels per step ----
        {
            const int i = min(pt0 + row / ns, n - 1);
            const int nb = s_nb[row];
            const float4 ev = *reinterpret_cast<const float4*>(s_e + row * 4);
            const float* krow = qkv + (size_t)nb * 3 * C + C;
            const float* qrow = qkv + (size_t)i * 3 * C;
#pragma unroll 1
            for (int ch = 0; ch < NCH; ++ch) {
                if (ch > 0) {
                    // the MMAs of the previous chunk have read the A tile and their W slot: refill that slot
                    umma::mbar_wait(&bar_mma, (n_mma - 1) & 1);
                    umma::fence_after_sync();
                    if (warp == 0 && ch + 1 < NCH)
                        umma::bulk_load(s_W + ((n_w + ch + 1) & 1) * 2 * W_BYTES, W.W1c + (size_t)(ch + 1) * 2 * TP * PT_KC, 2 * W_BYTES,
                                        &w_full[(n_w + ch + 1) & 1]);
                }
                const int c0 = ch * PT_KC + half * 32;
                // the k row of a (point, neighbour) pair is read once: four 256-bit loads that do not allocate in L1, all in flight
                float k8[4][8];
#pragma unroll
                for (int g2 = 0; g2 < 4; ++g2) etch_ld256_na(krow + c0 + g2 * 8, k8[g2]);
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const float4 qv = __ldg(reinterpret_cast<const float4*>(qrow + c0) + g);
                    const float kk[4] = {k8[g >> 1][(g & 1) * 4], k8[g >> 1][(g & 1) * 4 + 1], k8[g >> 1][(g & 1) * 4 + 2], k8[g >> 1][(g & 1) * 4 + 3]};
                    const float qq[4] = {qv.x, qv.y, qv.z, qv.w};
                    float hi[4], lo[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float4 ca = *reinterpret_cast<const float4*>(s_chan + (c0 + g * 4 + u) * 8);
                        const float2 cb = *reinterpret_cast<const float2*>(s_chan + (c0 + g * 4 + u) * 8 + 4);
                        const float pr = fmaf(ca.z, ev.z, fmaf(ca.y, ev.y, fmaf(ca.x, ev.x, ca.w)));
                        const float wv = fmaxf(fmaf(kk[u] - qq[u] + pr, cb.x, cb.y), 0.f);
                        umma::split_tf32(wv, hi[u], lo[u]);
                    }
                    const int kc = half * 8 + g;
                    *reinterpret_cast<float4*>(s_A + kc * (PT_ROWS * 16) + row * 16) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(s_A + PT_A_BYTES + kc * (PT_ROWS * 16) + row * 16) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
                umma::fence_async_smem();
                __syncthreads();
                if (warp == 0) {
                    const uint32_t slot = (n_w + ch) & 1;
                    umma::mbar_wait(&w_full[slot], ((n_w + ch) >> 1) & 1);
                    umma::fence_after_sync();
                    const uint32_t a_hi = umma::smem_u32(s_A), a_lo = a_hi + PT_A_BYTES;
                    const uint32_t b_hi = umma::smem_u32(s_W + slot * 2 * W_BYTES), b_lo = b_hi + W_BYTES;
                    umma::issue_gemm_3xtf32(tmem, a_hi, a_lo, b_hi, b_lo, PT_KC, TP, ch > 0);
                    umma::commit(&bar_mma);
                }
                ++n_mma;
            }
            n_w += NCH;
        }
        umma::mbar_wait(&bar_mma, (n_mma - 1) & 1);
        umma::fence_after_sync();
        // ---- C. relu(+b1), second layer (T x T) on the CUDA cores: both warpgroups read the row, each does half of the outputs ----
        {
            float hh[T];
            const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll
            for (int t0 = 0; t0 < T; t0 += 8) {
                float v[8];
                umma::tmem_ld8(taddr + t0, v);
#pragma unroll
                for (int u = 0; u < 8; ++u) hh[t0 + u] = fmaxf(v[u] + s_b1[t0 + u], 0.f);
            }
            constexpr int UH = T / 2;
#pragma unroll 1
            for (int u0 = 0; u0 < UH; u0 += 4) {
                float acc[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[k] = s_b2[half * UH + u0 + k];
#pragma unroll
                for (int t = 0; t < T; t += 4) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float4 w = *reinterpret_cast<const float4*>(s_W2 + (half * UH + u0 + k) * T + t);
                        acc[k] = fmaf(w.x, hh[t], acc[k]); acc[k] = fmaf(w.y, hh[t + 1], acc[k]);
                        acc[k] = fmaf(w.z, hh[t + 2], acc[k]); acc[k] = fmaf(w.w, hh[t + 3], acc[k]);
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) s_lg[row * LDL + half * UH + u0 + k] = acc[k];
            }
        }
        umma::fence_before_sync();
        __syncthreads();
        // ---- D. softmax over the ns neighbours of each point, per shared plane ----
        for (int it = tid; it < ppt * T; it += 256) {
            const int pl = it / T, u = it % T;
            float* col = s_lg + (pl * ns) * LDL + u;
            float lg[NS];
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < NS; ++j) { lg[j] = col[j * LDL]; mx = fmaxf(mx, lg[j]); }
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < NS; ++j) { lg[j] = exp2f((lg[j] - mx) * 1.4426950408889634f); s += lg[j]; }
            const float inv = 1.0f / s;
#pragma unroll
            for (int j = 0; j < NS; ++j) col[j * LDL] = lg[j] * inv;
        }
        __syncthreads();
        // ---- E. aggregation (v + p_r) * weight, channel-parallel; bn2 + ReLU epilogue ----
        for (int o = tid; o < ppt * C; o += 256) {
            const int pl = o / C, chn = o % C;
            const int i = pt0 + pl;
            if (i >= n) continue;
            const float4 ca = *reinterpret_cast<const float4*>(s_chan + chn * 8);
            const float4 cb = *reinterpret_cast<const float4*>(s_chan + chn * 8 + 4);
            float vv[NS];
#pragma unroll
            for (int j = 0; j < NS; ++j) vv[j] = __ldg(qkv + (size_t)s_nb[pl * NS + j] * 3 * C + 2 * C + chn);   // NS gathers in flight
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                const int r = pl * NS + j;
                const float4 ev = *reinterpret_cast<const float4*>(s_e + r * 4);
                const float pr = fmaf(ca.z, ev.z, fmaf(ca.y, ev.y, fmaf(ca.x, ev.x, ca.w)));
                acc = fmaf(vv[j] + pr, s_lg[r * LDL + (chn % T)], acc);
            }
            out[(size_t)i * C + chn] = fmaxf(fmaf(acc, cb.z, cb.w), 0.f);
        }
        __syncthreads();
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, Cfg::TCOLS);
}

template <int C, int NS>
int launch_pt_attn_tc(const float* p, const float* qkv, const int* idx, const AttnTcW& W, int n, float* out, cudaStream_t stream) {
    using Cfg = PtCfg<C>;
    constexpr int ns = NS;
    static_assert(Cfg::smem <= 227 * 1024, "shared memory budget");
    auto kern = pt_attn_tc_kernel<C, NS>;
    ETCH_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem));
    const int ppt = PT_ROWS / ns;
    const int ntiles = (n + ppt - 1) / ppt;
    const int per_sm = Cfg::smem <= 110 * 1024 ? 2 : 1;
    int grid = etch_sm_budget() * per_sm;
    if (grid > ntiles) grid = ntiles;
    kern<<<grid, 256, Cfg::smem, stream>>>(p, qkv, idx, W, n, out);
    ETCH_RETURN_LAST();
}

}  // namespace

// PointTransformerLayer core + bn2 + ReLU of the enclosing block, first attention linear on tcgen05.
// qkv = [n][3c] (q | k | v), idx = self kNN [n][ns] with ns in {8, 16}; chan = [c][8] {P3 row, p3b, s0, h0, so, ho};
// W1c = [c/64][2][16][max(c/8,16)][4] (TF32 hi/lo, canonical tiles; etch_b200/models/heads.py::_Block).
ETCH_API int etch_pt_attention_tc(const float* p, const float* qkv, const int* idx, const float* P0, const float* p0b,
                                  const float* chan, const float* W1c, const float* b1, const float* W2, const float* b2, int n,
                                  int ns, int c, float* out, cudaStream_t stream) {
    if (!p || !qkv || !idx || !P0 || !p0b || !chan || !W1c || !b1 || !W2 || !b2 || !out || n <= 0) return ETCH_EINVAL;
    if (ns != 8 && ns != 16) return ETCH_EINVAL;
    AttnTcW W{P0, p0b, chan, W1c, b1, W2, b2};
#define CASE(cc, nn) if (c == cc && ns == nn) return launch_pt_attn_tc<cc, nn>(p, qkv, idx, W, n, out, stream);
    CASE(64, 8) CASE(64, 16) CASE(128, 8) CASE(128, 16) CASE(256, 8) CASE(256, 16) CASE(512, 8) CASE(512, 16)
#undef CASE
    return ETCH_EINVAL;
}
