// Tensor-core (tcgen05 + TMEM) versions of the dense-GEMM stages of the SO(3) encoder.
//
//   anchor_gemm_tc : IntraSO3Conv / skip 1x1 conv       z[(p,a), o] = sum_j sum_c x[p, tab[a][j], c] W_j[o, c] + bias[o]
//   (same semantics as anchor_gemm_kernel in so3conv.cu; reference: vgtk/so3conv/functional.py:331-343, modules.py:19-39,
//    131-153, src/models/so3conv.py:36-44,178-180)
//
// One CTA owns 2 points x 60 anchors = 120 GEMM rows padded to the UMMA M = 128.  For every anchor-neighbour slot j the
// threads gather the (normalised) activation rows into the canonical K-major shared-memory tile, split fp32 -> (hi, lo)
// TF32 pairs, and one thread issues hi*hi + lo*hi + hi*lo tcgen05.mma's that accumulate in TMEM (fp32-level accuracy).
// The weight slices arrive pre-split and pre-tiled from the host through cp.async.bulk (TMA 1-D) on an mbarrier,
// double-buffered so the copy of slice j+1 overlaps the MMAs of slice j.  Epilogue: tcgen05.ld -> bias -> shared tile ->
// coalesced store + InstanceNorm statistics, as in the CUDA-core version.
#include "common.cuh"
#include "umma.cuh"

namespace {

constexpr int NA = 60;
constexpr int TP = 2;
constexpr int NPAIR = TP * NA;
constexpr int MROWS = 128;

__device__ __forceinline__ void stats_to_affine_tc(const double* sums, int b, int C, double count, float* s_mean, float* s_rstd) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const double s = sums[((size_t)b * C + c) * 2], ss = sums[((size_t)b * C + c) * 2 + 1];
        const double mean = s / count;
        double var = ss / count - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[c] = (float)mean;
        s_rstd[c] = (float)(1.0 / sqrt(var + 1e-5));
    }
}

using umma::bulk_load;

__device__ __forceinline__ void compute_warps_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Warp-specialised pipeline: warps 0-7 (256 threads) stage the activations, gather + split one A tile per anchor-neighbour
// slot j into a 2-deep ring, and run the epilogue; warp 8 streams the weight slices (cp.async.bulk ring) and issues the
// tcgen05.mma's, so the fill of slot j+1 overlaps the MMAs of slot j and nobody waits for an MMA to retire except
// through the ring's mbarriers.
template <int CIN, int COUT, int J>
struct AgemmCfg {
    static constexpr int LD = CIN + 4;
    static constexpr uint32_t A_BYTES = MROWS * CIN * 4;     // one (hi or lo) A tile
    static constexpr uint32_t B_BYTES = COUT * CIN * 4;      // one (hi or lo) weight slice
    static constexpr int WR = (2 * B_BYTES <= 8192) ? 3 : 2;  // weight ring depth (3 keeps the c = 32 kernels at two CTAs per SM)
    static constexpr size_t smem = (size_t)4 * A_BYTES + (size_t)WR * 2 * B_BYTES + (size_t)NPAIR * LD * 4 + (size_t)2 * CIN * 4 +
                                   (size_t)((NA * J + 15) / 16) * 16 + 128;
    static constexpr int CTAS_PER_SM = smem <= 112 * 1024 ? 2 : 1;   // latency-bound fill: a second CTA per SM hides it
};

template <int CIN, int COUT, int J, bool NORM_IN>
__global__ void __launch_bounds__(288, (AgemmCfg<CIN, COUT, J>::CTAS_PER_SM)) anchor_gemm_tc_kernel(
    const float* __restrict__ xin,       // [B,Q,60,CIN]
    const int* __restrict__ src_idx,     // [B,P] or nullptr
    const int* __restrict__ tab,         // [60][J]
    const float* __restrict__ Wc,        // [J][2 (hi,lo)][CIN/4][COUT][4]  canonical K-major tiles, TF32-split on the host
    const float* __restrict__ bias,      // [COUT]
    const double* __restrict__ in_stats, double in_count, int Q, int P,
    float* __restrict__ zraw, double* __restrict__ stats)
{
    using Cfg = AgemmCfg<CIN, COUT, J>;
    constexpr int LD = Cfg::LD;
    constexpr uint32_t A_BYTES = Cfg::A_BYTES, B_BYTES = Cfg::B_BYTES;
    constexpr int WR = Cfg::WR;
    constexpr int TCOLS = COUT < 32 ? 32 : COUT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* s_A = smem_raw;                                   // [2 buffers][hi | lo]
    unsigned char* s_B = s_A + 4 * A_BYTES;                          // [WR buffers][hi | lo]
    float* s_x = reinterpret_cast<float*>(s_B + WR * 2 * B_BYTES);   // [120][LD]
    float* s_mean = s_x + NPAIR * LD;
    float* s_rstd = s_mean + CIN;
    unsigned char* s_tab = reinterpret_cast<unsigned char*>(s_rstd + CIN);   // [60][J]
    float* s_z = reinterpret_cast<float*>(s_A);                      // epilogue reuse [128][COUT]
    __shared__ uint64_t a_full[2], a_free[2], w_full[WR], w_free[WR], acc_full;
    __shared__ uint32_t tmem_base;

    const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < NA * J; i += 288) s_tab[i] = (unsigned char)__ldg(tab + i);
    if (NORM_IN) stats_to_affine_tc(in_stats, b, CIN, in_count, s_mean, s_rstd);
    if (warp == 8) umma::tmem_alloc(&tmem_base, TCOLS);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { umma::mbar_init(&a_full[i], 8); umma::mbar_init(&a_free[i], 1); }
        for (int i = 0; i < WR; ++i) { umma::mbar_init(&w_full[i], 1); umma::mbar_init(&w_free[i], 1); }
        umma::mbar_init(&acc_full, 1);
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = umma::uniform(tmem_base);
    const int ntiles = (P + TP - 1) / TP;
    const int my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == 8) {
        // ===== weight streaming + MMA issue (warp-collective) =====
        const uint32_t total = (uint32_t)my_tiles * J;
        uint32_t issued = 0;
        for (uint32_t g = 0; g < total; ++g) {
            while (issued < total && issued < g + WR) {
                const uint32_t wb = issued % WR, use = issued / WR;
                if (use > 0) umma::mbar_wait(&w_free[wb], (use - 1) & 1);
                bulk_load(s_B + wb * 2 * B_BYTES, Wc + (size_t)(issued % J) * 2 * COUT * CIN, 2 * B_BYTES, &w_full[wb]);
                ++issued;
            }
            const uint32_t wb = g % WR, ab = g & 1, j = g % J;
            umma::mbar_wait(&w_full[wb], (g / WR) & 1);
            umma::mbar_wait(&a_full[ab], (g >> 1) & 1);
            umma::fence_after_sync();
            const uint32_t a_hi = umma::smem_u32(s_A + ab * 2 * A_BYTES), a_lo = a_hi + A_BYTES;
            const uint32_t b_hi = umma::smem_u32(s_B + wb * 2 * B_BYTES), b_lo = b_hi + B_BYTES;
            umma::issue_gemm_3xtf32(tmem, a_hi, a_lo, b_hi, b_lo, CIN, COUT, j > 0);
            umma::commit(&a_free[ab]);
            umma::commit(&w_free[wb]);
            if (j == J - 1) umma::commit(&acc_full);
        }
    } else {
        // ===== activation staging, A-tile fills, epilogue =====
        constexpr int RG = 256 / COUT;                     // row groups of the statistics pass
        const int scol = tid % COUT, srg = tid / COUT;
        double acc_s = 0.0, acc_ss = 0.0;
        const int frow = tid & 127, fhalf = tid >> 7;      // fill mapping: row, channel half
        const bool fok = frow < NPAIR;
        const int fpl = fok ? frow / NA : 0, fa = fok ? frow % NA : 0;
        uint32_t g = 0, ntile_done = 0;
        // The tile's activations are fetched one tile ahead into registers: the global / L2 latency of the staging loads (17 % of the
        // kernel's warp time in the round-1 profile, more for the single-slot skip conv) hides behind the fills of the previous tile.
        constexpr int NV = (NPAIR * (CIN / 4) + 255) / 256;
        float4 pre[NV];
        auto fetch_tile = [&](int tile_) {
#pragma unroll
            for (int u = 0; u < NV; ++u) {
                const int t = tid + u * 256;
                if (t < NPAIR * (CIN / 4)) {
                    const int row = t / (CIN / 4), c4 = t % (CIN / 4);
                    const int p = min(tile_ * TP + row / NA, P - 1);
                    const int sp = src_idx ? __ldg(src_idx + (size_t)b * P + p) : p;
                    pre[u] = __ldg(reinterpret_cast<const float4*>(xin + (((size_t)b * Q + sp) * NA + (row % NA)) * CIN) + c4);
                }
            }
        };
        if ((int)blockIdx.x < ntiles) fetch_tile(blockIdx.x);
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int p0 = tile * TP;
            const int npts = min(TP, P - p0);
            // stage (and normalise) the activations of the tile's points
#pragma unroll
            for (int u = 0; u < NV; ++u) {
                const int t = tid + u * 256;
                if (t >= NPAIR * (CIN / 4)) break;
                const int row = t / (CIN / 4), c4 = t % (CIN / 4);
                float4 v = pre[u];
                if (NORM_IN) {
                    const int c = c4 * 4;
                    v.x = etch_lrelu((v.x - s_mean[c]) * s_rstd[c]);
                    v.y = etch_lrelu((v.y - s_mean[c + 1]) * s_rstd[c + 1]);
                    v.z = etch_lrelu((v.z - s_mean[c + 2]) * s_rstd[c + 2]);
                    v.w = etch_lrelu((v.w - s_mean[c + 3]) * s_rstd[c + 3]);
                }
                *reinterpret_cast<float4*>(s_x + row * LD + c4 * 4) = v;
            }
            compute_warps_sync();
            if (tile + (int)gridDim.x < ntiles) fetch_tile(tile + gridDim.x);

            for (int j = 0; j < J; ++j, ++g) {
                const uint32_t ab = g & 1;
                if (g >= 2) umma::mbar_wait(&a_free[ab], ((g >> 1) - 1) & 1);
                if (fok) {   // rows 120..127 are never written: their (garbage) products land in TMEM rows that are never read
                    const int srow = fpl * NA + s_tab[fa * J + j];
                    unsigned char* dh = s_A + ab * 2 * A_BYTES + frow * 16;
                    if constexpr (CIN == 32) {
                        // The 8 lanes of a quarter warp gather 8 arbitrary source rows; at the same column two rows that are congruent
                        // mod 8 share a 16-byte bank group (row stride = 9 groups), which doubled the wavefronts of this load (ncu:
                        // 29.4 M for 14.4 M ideal).  Lane l of thread half f therefore takes the chunk that lies in bank group
                        // (l + 4 f + i) mod 8 at step i: the quarter warp covers all 8 groups at every step and the two halves cover
                        // the 8 chunks of a row between them.
                        const float* xr = s_x + srow * LD;
                        const int rot = lane + 4 * fhalf - srow;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int c4 = (rot + i) & 7;
                            const float4 v = *reinterpret_cast<const float4*>(xr + c4 * 4);
                            float4 h, l;
                            umma::split_tf32(v.x, h.x, l.x); umma::split_tf32(v.y, h.y, l.y);
                            umma::split_tf32(v.z, h.z, l.z); umma::split_tf32(v.w, h.w, l.w);
                            *reinterpret_cast<float4*>(dh + c4 * (MROWS * 16)) = h;
                            *reinterpret_cast<float4*>(dh + A_BYTES + c4 * (MROWS * 16)) = l;
                        }
                    } else {
                        const float* xr = s_x + srow * LD + fhalf * (CIN / 2);
#pragma unroll
                        for (int c4 = 0; c4 < CIN / 8; ++c4) {
                            const float4 v = *reinterpret_cast<const float4*>(xr + c4 * 4);
                            float4 h, l;
                            umma::split_tf32(v.x, h.x, l.x); umma::split_tf32(v.y, h.y, l.y);
                            umma::split_tf32(v.z, h.z, l.z); umma::split_tf32(v.w, h.w, l.w);
                            const int kc = fhalf * (CIN / 8) + c4;
                            *reinterpret_cast<float4*>(dh + kc * (MROWS * 16)) = h;
                            *reinterpret_cast<float4*>(dh + A_BYTES + kc * (MROWS * 16)) = l;
                        }
                    }
                }
                umma::fence_async_smem();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&a_full[ab]);
            }
            // ---- epilogue: TMEM -> registers -> (+bias) -> shared tile (aliases the A ring: every MMA has retired) ----
            umma::mbar_wait(&acc_full, ntile_done & 1);
            ++ntile_done;
            umma::fence_after_sync();
            {
                const int q = warp & 3, half = warp >> 2;
                const int row = q * 32 + lane;
#pragma unroll
                for (int c0 = 0; c0 < COUT / 2; c0 += 8) {
                    float v[8];
                    const int col = half * (COUT / 2) + c0;
                    umma::tmem_ld8(tmem + ((uint32_t)(q * 32) << 16) + col, v);
                    float4 o0, o1;
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col)), b1 = __ldg(reinterpret_cast<const float4*>(bias + col + 4));
                    o0 = make_float4(v[0] + b0.x, v[1] + b0.y, v[2] + b0.z, v[3] + b0.w);
                    o1 = make_float4(v[4] + b1.x, v[5] + b1.y, v[6] + b1.z, v[7] + b1.w);
                    // 16-byte chunk ch of row r sits at chunk (ch ^ (r & 7)): the row stride is a multiple of 128 B, so the unswizzled
                    // stores of 8 consecutive rows all hit one bank group (ncu: 8-way conflicts on both stores)
                    *reinterpret_cast<float4*>(s_z + row * COUT + (((col >> 2) ^ (row & 7)) << 2)) = o0;
                    *reinterpret_cast<float4*>(s_z + row * COUT + ((((col >> 2) + 1) ^ (row & 7)) << 2)) = o1;
                }
            }
            umma::fence_before_sync();
            compute_warps_sync();
            const int nvalid = npts * NA;
            float* dst = zraw + ((size_t)b * P + p0) * NA * COUT;
            for (int i = tid; i < nvalid * COUT / 4; i += 256) {
                const int r = i / (COUT / 4), ch = i % (COUT / 4);
                reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(s_z)[r * (COUT / 4) + (ch ^ (r & 7))];
            }
            {
                float s = 0.f, ss = 0.f;
                for (int r = srg; r < nvalid; r += RG) { const float v = s_z[r * COUT + ((((scol >> 2) ^ (r & 7)) << 2) | (scol & 3))]; s += v; ss = fmaf(v, v, ss); }
                acc_s += (double)s; acc_ss += (double)ss;
            }
            compute_warps_sync();   // s_z (= A ring) and s_x are free for the next tile
        }
        atomicAdd(stats + ((size_t)b * COUT + scol) * 2, acc_s);
        atomicAdd(stats + ((size_t)b * COUT + scol) * 2 + 1, acc_ss);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 8) umma::tmem_dealloc(tmem, TCOLS);
}

// ---------------------------------------------------------------------------------------------------------------------
// InterSO3Conv with the channel-mixing GEMM on the tensor cores.
//   z[(p,a), o] = sum_{c,k} W[o, c*24+k] * ( sum_n f[nbr(p,n), a, c] * relu(1 - |g_n - R_a kappa_k|^2 / sigma) ) + bias[o]
// (vgtk/so3conv/functional.py:286-324,61-67; modules.py:19-39,120-128).  The neighbour contraction (150k tiny
// block-diagonal products per scan: not a dense GEMM) stays on the CUDA cores exactly as in inter_conv_kernel; its
// result y is produced in slabs of 8 channels x 12 kernel points = 96 K-columns for all 120 (p,a) rows, written straight
// into the canonical UMMA A tile as (hi, lo) TF32 pairs, and multiplied with the matching pre-split slab of W by
// tcgen05.mma (3xTF32) while the CUDA cores already work on the next slab.  Weight slabs stream through a 2-deep
// cp.async.bulk ring; the [128 x C_out] accumulator lives in TMEM for the whole tile.
template <int CIN, int COUT, int NN>
__global__ void __launch_bounds__(256, 1) inter_conv_tc_kernel(
    const float* __restrict__ xyz,       // [B,3,q]
    const float* __restrict__ feat,      // [B,q,60,CIN]
    const int* __restrict__ sample_idx,  // [B,P]
    const int* __restrict__ nbr,         // [B,P,NN]
    const float4* __restrict__ krs,      // [60,24] {2/sigma * R_a k, |R_a k|^2/sigma}
    const float* __restrict__ Wc,        // [CIN/8*2][2][24][COUT][4]  weight slabs (hi, lo), K'' = kgl*48 + c*6 + i
    const float* __restrict__ bias,
    int q, int P, float inv_sigma,
    float* __restrict__ zraw, double* __restrict__ stats)
{
    constexpr int NK = 24;
    constexpr int KS = 96;                            // K columns per slab
    constexpr int NSLAB = CIN / 8 * 2;
    constexpr uint32_t A_BYTES = MROWS * KS * 4;      // 49152
    constexpr uint32_t W_BYTES = COUT * KS * 4;       // one (hi|lo) weight slab
    constexpr int TCOLS = COUT < 32 ? 32 : COUT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* s_A = smem_raw;                                        // [hi | lo]
    unsigned char* s_W = s_A + 2 * A_BYTES;                               // [2 slots][hi | lo]
    float4* s_krs = reinterpret_cast<float4*>(s_W + 4 * W_BYTES);         // [60*24]
    float4* s_g = s_krs + NA * NK;                                        // [TP][NN]
    int* s_off = reinterpret_cast<int*>(s_g + TP * NN);                   // [TP][NN]
    float* s_z = reinterpret_cast<float*>(s_A);                           // epilogue reuse [128][COUT]
    __shared__ uint64_t bar_mma, bar_w[2];
    __shared__ uint32_t tmem_base;

    const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < NA * NK; i += 256) s_krs[i] = __ldg(krs + i);
    if (warp == 0) umma::tmem_alloc(&tmem_base, TCOLS);
    if (tid == 0) { umma::mbar_init(&bar_mma, 1); umma::mbar_init(&bar_w[0], 1); umma::mbar_init(&bar_w[1], 1); }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = umma::uniform(tmem_base);
    const float* X = xyz + (size_t)b * 3 * q;
    const float* F = feat + (size_t)b * q * NA * CIN;
    uint32_t n_mma = 0, n_w = 0;     // MMA groups issued / weight slabs requested so far (uniform)
    double acc_s = 0.0, acc_ss = 0.0;
    const int ntiles = (P + TP - 1) / TP;
    // GEMM1 task of this thread: pair = 16*warp + lane%16 (valid < 120), kernel-point half kgl = lane/16
    const int pair = warp * 16 + (lane & 15), kgl = lane >> 4;
    const bool task_ok = pair < NPAIR;
    const int pl = task_ok ? pair / NA : 0, a = task_ok ? pair % NA : 0;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int p0 = tile * TP;
        const int npts = min(TP, P - p0);
        if (warp == 0) { umma::bulk_load(s_W + (n_w & 1) * 2 * W_BYTES, Wc, 2 * W_BYTES, &bar_w[n_w & 1]); }
        ++n_w;
        for (int t = tid; t < TP * NN; t += 256) {
            const int ppl = t / NN, n = t % NN;
            const int p = min(p0 + ppl, P - 1);
            const int c = __ldg(sample_idx + (size_t)b * P + p);
            const int k = __ldg(nbr + ((size_t)b * P + p) * NN + n);
            const float gx = __ldg(X + k) - __ldg(X + c);
            const float gy = __ldg(X + q + k) - __ldg(X + q + c);
            const float gz = __ldg(X + 2 * (size_t)q + k) - __ldg(X + 2 * (size_t)q + c);
            s_g[t] = make_float4(gx, gy, gz, 1.0f - (gx * gx + gy * gy + gz * gz) * inv_sigma);
            s_off[t] = k * NA * CIN;
        }
        for (int t = tid; t < (MROWS - NPAIR) * (KS / 4) * 2; t += 256) {   // pad rows of the A tile (aliased by s_z)
            const int which = t / ((MROWS - NPAIR) * (KS / 4)), rem = t % ((MROWS - NPAIR) * (KS / 4));
            const int r = NPAIR + rem / (KS / 4), kc = rem % (KS / 4);
            *reinterpret_cast<float4*>(s_A + which * A_BYTES + kc * (MROWS * 16) + r * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();

        for (int sl = 0; sl < NSLAB; ++sl) {
            const int c0 = (sl >> 1) * 8, kg = (sl & 1) * 2 + kgl;
            // ---- GEMM1 for this slab: y[c][i] = sum_n f[nbr_n][a][c0+c] * w[a][kg*6+i][n] ----
            float acc[8][6];
#pragma unroll
            for (int c = 0; c < 8; ++c)
#pragma unroll
                for (int i = 0; i < 6; ++i) acc[c][i] = 0.f;
            if (task_ok) {
                float4 kq[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) kq[i] = s_krs[a * NK + kg * 6 + i];
                const float* fa = F + a * CIN + c0;
                // The kernel is bound by L1 wavefronts (one per distinct 128-B line per load instruction): fetch the 8 channels of a
                // neighbour with ONE 256-bit load, and keep PF neighbours in flight per thread.
                constexpr int PF = 4;
                static_assert(NN % PF == 0, "neighbour count must be a multiple of the prefetch depth");
                float fb[PF][8];
#pragma unroll
                for (int j = 0; j < PF; ++j) etch_ldg256(fa + s_off[pl * NN + j], fb[j]);
#pragma unroll 1
                for (int n0 = 0; n0 < NN; n0 += PF) {
#pragma unroll
                    for (int j = 0; j < PF; ++j) {
                        const int n = n0 + j;
                        const float4 g = s_g[pl * NN + n];
                        float fv[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) fv[c] = fb[j][c];
                        if (n + PF < NN) etch_ldg256(fa + s_off[pl * NN + n + PF], fb[j]);
#pragma unroll
                        for (int i = 0; i < 6; ++i) {
                            const float w = fmaxf(fmaf(g.x, kq[i].x, fmaf(g.y, kq[i].y, fmaf(g.z, kq[i].z, g.w - kq[i].w))), 0.f);
#pragma unroll
                            for (int c = 0; c < 8; ++c) acc[c][i] = fmaf(fv[c], w, acc[c][i]);
                        }
                    }
                }
            }
            // ---- the previous slab's MMAs must have finished reading the A tile; then request the next weight slab ----
            if (n_mma > 0) { umma::mbar_wait(&bar_mma, (n_mma - 1) & 1); umma::fence_after_sync(); }
            if (warp == 0 && sl + 1 < NSLAB) {
                umma::bulk_load(s_W + (n_w & 1) * 2 * W_BYTES, Wc + (size_t)(sl + 1) * 2 * COUT * KS, 2 * W_BYTES, &bar_w[n_w & 1]);
            }
            if (sl + 1 < NSLAB) ++n_w;
            // ---- y -> canonical A tile as (hi, lo): 48 consecutive K columns of this row = 12 float4 each ----
            if (task_ok) {
                const float* av = &acc[0][0];   // v = c*6 + i
#pragma unroll
                for (int v4 = 0; v4 < 12; ++v4) {
                    float4 h, l;
                    umma::split_tf32(av[v4 * 4 + 0], h.x, l.x); umma::split_tf32(av[v4 * 4 + 1], h.y, l.y);
                    umma::split_tf32(av[v4 * 4 + 2], h.z, l.z); umma::split_tf32(av[v4 * 4 + 3], h.w, l.w);
                    const int kc = kgl * 12 + v4;
                    *reinterpret_cast<float4*>(s_A + kc * (MROWS * 16) + pair * 16) = h;
                    *reinterpret_cast<float4*>(s_A + A_BYTES + kc * (MROWS * 16) + pair * 16) = l;
                }
            }
            umma::fence_async_smem();
            __syncthreads();
            if (warp == 0) {
                const uint32_t gw = n_w - (sl + 1 < NSLAB ? 2 : 1);      // index of the weight slab consumed now
                umma::mbar_wait(&bar_w[gw & 1], (gw >> 1) & 1);
                umma::fence_after_sync();
                const uint32_t a_hi = umma::smem_u32(s_A), a_lo = a_hi + A_BYTES;
                const uint32_t b_hi = umma::smem_u32(s_W + (gw & 1) * 2 * W_BYTES), b_lo = b_hi + W_BYTES;
                umma::issue_gemm_3xtf32(tmem, a_hi, a_lo, b_hi, b_lo, KS, COUT, sl > 0);
                umma::commit(&bar_mma);
            }
            ++n_mma;
        }
        // ---- epilogue ----
        umma::mbar_wait(&bar_mma, (n_mma - 1) & 1);
        umma::fence_after_sync();
        {
            const int qq = warp & 3, half = warp >> 2;
            const int row = qq * 32 + lane;
#pragma unroll
            for (int cc = 0; cc < COUT / 2; cc += 8) {
                float v[8];
                const int col = half * (COUT / 2) + cc;
                umma::tmem_ld8(tmem + ((uint32_t)(qq * 32) << 16) + col, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) s_z[row * COUT + col + i] = v[i] + __ldg(bias + col + i);
            }
        }
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();
        const int nvalid = npts * NA;
        float* dst = zraw + ((size_t)b * P + p0) * NA * COUT;
        for (int i = tid; i < nvalid * COUT / 4; i += 256)
            reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(s_z)[i];
        if (tid < COUT) {
            float s = 0.f, ss = 0.f;
            for (int r = 0; r < nvalid; ++r) { const float v = s_z[r * COUT + tid]; s += v; ss = fmaf(v, v, ss); }
            acc_s += (double)s; acc_ss += (double)ss;
        }
        __syncthreads();
    }
    if (tid < COUT) {
        atomicAdd(stats + ((size_t)b * COUT + tid) * 2, acc_s);
        atomicAdd(stats + ((size_t)b * COUT + tid) * 2 + 1, acc_ss);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, TCOLS);
}

int grid_for_tc(int ntiles, int B) {
    int g = etch_sm_budget() / B;   // one wave of persistent CTAs (1 CTA/SM)
    if (g < 1) g = 1;
    return ntiles < g ? ntiles : g;
}

template <int CIN, int COUT, int J, bool NORM>
int launch_agemm_tc(const float* xin, const int* src_idx, const int* tab, const float* Wc, const float* bias,
                    const double* in_stats, double in_count, int B, int Q, int P, float* zraw, double* stats, cudaStream_t stream) {
    using Cfg = AgemmCfg<CIN, COUT, J>;
    static_assert((size_t)MROWS * COUT * 4 <= (size_t)4 * Cfg::A_BYTES, "epilogue tile must fit the A ring");
    static_assert(Cfg::smem <= 227 * 1024, "shared memory budget");
    static_assert(256 % COUT == 0, "statistics pass mapping");
    auto kern = anchor_gemm_tc_kernel<CIN, COUT, J, NORM>;
    ETCH_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem));
    int gx = etch_sm_budget() * Cfg::CTAS_PER_SM / B;
    if (gx < 1) gx = 1;
    if (gx > (P + TP - 1) / TP) gx = (P + TP - 1) / TP;
    dim3 grid(gx, B);
    kern<<<grid, 288, Cfg::smem, stream>>>(xin, src_idx, tab, Wc, bias, in_stats, in_count, Q, P, zraw, stats);
    ETCH_RETURN_LAST();
}

template <int CIN, int COUT, int NN>
int launch_inter_tc(const float* xyz, const float* feat, const int* sample_idx, const int* nbr, const float* krs, const float* Wc,
                    const float* bias, int B, int q, int P, float sigma, float* zraw, double* stats, cudaStream_t stream) {
    constexpr size_t smem = (size_t)2 * MROWS * 96 * 4 + (size_t)4 * COUT * 96 * 4 + (size_t)NA * 24 * 16 + (size_t)TP * NN * 16 +
                            (size_t)TP * NN * 4 + 128;
    auto kern = inter_conv_tc_kernel<CIN, COUT, NN>;
    ETCH_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(grid_for_tc((P + TP - 1) / TP, B), B);
    kern<<<grid, 256, smem, stream>>>(xyz, feat, sample_idx, nbr, reinterpret_cast<const float4*>(krs), Wc, bias, q, P, 1.0f / sigma,
                                      zraw, stats);
    ETCH_RETURN_LAST();
}

}  // namespace

// Tensor-core InterSO3Conv (c_in in {32,64}).  Wc = [cin/8*2][2][24][cout][4]: weight slabs in (hi, lo) canonical tiles with
// K'' = kgl*48 + c*6 + i  <->  W[o][(c0+c)*24 + (2h+kgl)*6 + i]  (etch_b200/models/encoder.py).
ETCH_API int etch_so3_inter_conv_tc(const float* xyz, const float* feat, const int* sample_idx, const int* nbr, const float* krs,
                                    const float* Wc, const float* bias, int B, int q, int P, int nn, int cin, int cout,
                                    float sigma, float* zraw, double* stats, cudaStream_t stream) {
    if (!xyz || !feat || !sample_idx || !nbr || !krs || !Wc || !bias || !zraw || !stats) return ETCH_EINVAL;
#define CASE(ci, co, n) \
    if (cin == ci && cout == co && nn == n) return launch_inter_tc<ci, co, n>(xyz, feat, sample_idx, nbr, krs, Wc, bias, B, q, P, sigma, zraw, stats, stream);
    CASE(32, 32, 32) CASE(32, 64, 64) CASE(64, 64, 32)
#undef CASE
    return ETCH_EINVAL;
}

// Tensor-core IntraSO3Conv. Wc = [12][2][c/4][cout][4] (TF32 hi/lo split, canonical K-major tiles; see etch_b200/models/tc.py)
ETCH_API int etch_so3_intra_conv_tc(const float* zin, const double* in_stats, const int* intra_idx, const float* Wc,
                                    const float* bias, int B, int P, int c, int cout, float* zraw, double* stats,
                                    cudaStream_t stream) {
    if (!zin || !in_stats || !intra_idx || !Wc || !bias || !zraw || !stats) return ETCH_EINVAL;
    const double cnt = (double)P * NA;
    if (c == 32 && cout == 32) return launch_agemm_tc<32, 32, 12, true>(zin, nullptr, intra_idx, Wc, bias, in_stats, cnt, B, P, P, zraw, stats, stream);
    if (c == 64 && cout == 64) return launch_agemm_tc<64, 64, 12, true>(zin, nullptr, intra_idx, Wc, bias, in_stats, cnt, B, P, P, zraw, stats, stream);
    return ETCH_EINVAL;
}

// Tensor-core skip 1x1 conv. Wc = [1][2][cin/4][cout][4]
ETCH_API int etch_so3_skip_conv_tc(const float* feat, const int* sample_idx, const int* ident, const float* Wc,
                                   const float* bias, int B, int q, int P, int cin, int cout, float* zraw, double* stats,
                                   cudaStream_t stream) {
    if (!feat || !ident || !Wc || !bias || !zraw || !stats) return ETCH_EINVAL;
    if (cin == 32 && cout == 32) return launch_agemm_tc<32, 32, 1, false>(feat, sample_idx, ident, Wc, bias, nullptr, 1.0, B, q, P, zraw, stats, stream);
    if (cin == 32 && cout == 64) return launch_agemm_tc<32, 64, 1, false>(feat, sample_idx, ident, Wc, bias, nullptr, 1.0, B, q, P, zraw, stats, stream);
    if (cin == 64 && cout == 64) return launch_agemm_tc<64, 64, 1, false>(feat, sample_idx, ident, Wc, bias, nullptr, 1.0, B, q, P, zraw, stats, stream);
    return ETCH_EINVAL;
}
