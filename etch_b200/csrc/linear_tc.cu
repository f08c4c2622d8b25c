// Generic tensor-core fused linear layer for the PointTransformer heads:
//     Y[n, co] = relu?( (X[n, ci] W^T + seg[segment(row)]) * scale + shift + R )
// i.e. nn.Linear / Conv1d(k=1) + eval BatchNorm1d (folded) + residual + ReLU in one launch
// (src/models/pointtransformer_seg.py:40-51,71-80,101-112,144-145,222).  Same contract as etch_linear (pt.cu), but the
// GEMM runs on tcgen05 (3xTF32, fp32-level accuracy) with the accumulator in TMEM.
//
// WEIGHT-STATIONARY: a CTA owns NB output columns, bulk-copies their pre-split (hi, lo) weight tile [NB x Kpad] into
// shared memory once, and then streams 128-row tiles of X through it: every 32-wide K chunk of the tile is staged
// (TF32 split, canonical layout, double-buffered) by all threads while warp 0 issues the MMAs of the previous chunk.
// Weights are therefore read from L2 once per CTA instead of once per row tile, and X is read NG = co_pad/NB times.
#include "common.cuh"
#include "umma.cuh"

namespace {

constexpr uint32_t LT_A = 128 * 32 * 4;   // bytes of one (hi|lo) X chunk [128 rows x 32 k]

__global__ void __launch_bounds__(256, 1) linear_tc_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ Wc,
                                                           int n, int ci, int co, int Kpad, int NB,
                                                           const float* __restrict__ scale, const float* __restrict__ shift,
                                                           const float* __restrict__ R, const float* __restrict__ seg,
                                                           const int* __restrict__ seg_off, int nseg, int relu,
                                                           float* __restrict__ Y, int ldy, int tcols) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* s_A = smem_raw;                 // [2 buffers][hi|lo]
    unsigned char* s_W = s_A + 4 * LT_A;           // [hi|lo] [Kpad/4][NB][4]
    __shared__ __align__(16) float s_scale[256], s_shift[256];   // folded BatchNorm of this CTA's NB columns (1, 0 when absent)
    __shared__ uint64_t a_free[2], w_full, acc_full;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp_col0 = blockIdx.y * NB;
    const uint32_t W_BYTES = (uint32_t)NB * Kpad * 4;     // one (hi or lo) tile
    for (int i = tid; i < NB; i += 256) {
        const int gc = blockIdx.y * NB + i;
        s_scale[i] = (scale && gc < co) ? __ldg(scale + gc) : 1.0f;
        s_shift[i] = (shift && gc < co) ? __ldg(shift + gc) : 0.0f;
    }
    if (warp == 0) umma::tmem_alloc(&tmem_base, tcols);
    if (tid == 0) { umma::mbar_init(&a_free[0], 1); umma::mbar_init(&a_free[1], 1); umma::mbar_init(&w_full, 1); umma::mbar_init(&acc_full, 1); }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = umma::uniform(tmem_base);
    if (warp == 0) {   // weights of this column group: one bulk copy per (hi|lo) tile, chunked to stay under the tx-count limit
        umma::bulk_load_region(s_W, Wc + (size_t)blockIdx.y * 2 * NB * Kpad, 2 * W_BYTES, &w_full);
    }
    uint32_t ga = 0, nacc = 0;
    const int nchunks = Kpad / 32;
    const int frow = tid >> 1, fhalf = tid & 1;
    bool w_ready = false;

    // X[m0 + frow, kc*32 + fhalf*16 .. +16) -> 4 float4 registers (zero beyond n / ci); issued one chunk ahead of its use so
    // the global-load latency overlaps the barrier, the MMA issue and the ring wait of the current chunk
    auto load_chunk = [&](int tile, int kc, float4 (&v)[4]) {
        const int gm_row = tile * 128 + frow;
        const int k0 = kc * 32 + fhalf * 16;
        const float* src = X + (size_t)(gm_row < n ? gm_row : 0) * ldx + k0;
        if (gm_row < n && k0 + 15 < ci && ((reinterpret_cast<uintptr_t>(src) & 31) == 0)) {
            // the common case: two 256-bit loads that do not allocate in L1 (X is streamed once per CTA; see etch_ld256_na)
            float a8[8], b8[8];
            etch_ld256_na(src, a8);
            etch_ld256_na(src + 8, b8);
            v[0] = make_float4(a8[0], a8[1], a8[2], a8[3]); v[1] = make_float4(a8[4], a8[5], a8[6], a8[7]);
            v[2] = make_float4(b8[0], b8[1], b8[2], b8[3]); v[3] = make_float4(b8[4], b8[5], b8[6], b8[7]);
            return;
        }
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
            const int k = k0 + c4 * 4;
            v[c4] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gm_row < n) {
                if (k + 3 < ci && ((reinterpret_cast<uintptr_t>(src + c4 * 4) & 15) == 0)) v[c4] = __ldg(reinterpret_cast<const float4*>(src + c4 * 4));
                else {
                    if (k < ci) v[c4].x = __ldg(src + c4 * 4);
                    if (k + 1 < ci) v[c4].y = __ldg(src + c4 * 4 + 1);
                    if (k + 2 < ci) v[c4].z = __ldg(src + c4 * 4 + 2);
                    if (k + 3 < ci) v[c4].w = __ldg(src + c4 * 4 + 3);
                }
            }
        }
    };
    // The (tile, K-chunk) pairs of this CTA form one sequence g = 0 .. G-1; PF chunks of X are always in flight in registers
    // (each register set is refilled with chunk g + PF right after chunk g has been split into the A ring), so the L2 / HBM
    // latency of a row tile is hidden behind PF chunk steps instead of one.
    constexpr int PF = 3;
    const int my_tiles = (int)blockIdx.x * 128 < n ? ((n + 127) / 128 - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int G = my_tiles * nchunks;
    float4 xv0[4], xv1[4], xv2[4];
    auto prefetch = [&](int g, float4 (&v)[4]) {
        if (g < G) load_chunk((int)blockIdx.x + (g / nchunks) * (int)gridDim.x, g % nchunks, v);
    };
    prefetch(0, xv0); prefetch(1, xv1); prefetch(2, xv2);

    auto step = [&](int g, float4 (&xv)[4]) {
        const int tile = (int)blockIdx.x + (g / nchunks) * (int)gridDim.x, kc = g % nchunks;
        const int m0 = tile * 128;
        const uint32_t abuf = ga & 1;
        if (ga >= 2) umma::mbar_wait(&a_free[abuf], ((ga - 2) >> 1) & 1);
        {   // split + store the staged chunk: 2 threads per row, 16 floats each
            unsigned char* dh = s_A + abuf * 2 * LT_A;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const float4 v = xv[c4];
                float4 h, l;
                umma::split_tf32(v.x, h.x, l.x); umma::split_tf32(v.y, h.y, l.y);
                umma::split_tf32(v.z, h.z, l.z); umma::split_tf32(v.w, h.w, l.w);
                const int kq = fhalf * 4 + c4;
                *reinterpret_cast<float4*>(dh + kq * (128 * 16) + frow * 16) = h;
                *reinterpret_cast<float4*>(dh + LT_A + kq * (128 * 16) + frow * 16) = l;
            }
        }
        prefetch(g + PF, xv);
        umma::fence_async_smem();
        __syncthreads();
        if (warp == 0) {
            if (!w_ready) { umma::mbar_wait(&w_full, 0); w_ready = true; }
            umma::fence_after_sync();
            const uint32_t a_hi = umma::smem_u32(s_A + abuf * 2 * LT_A), a_lo = a_hi + LT_A;
            const uint32_t koff = (uint32_t)kc * 8 * NB * 16;     // 8 k-chunks of 4 floats, each NB*16 bytes
            const uint32_t b_hi = umma::smem_u32(s_W) + koff, b_lo = b_hi + W_BYTES;
            umma::issue_gemm_3xtf32(tmem, a_hi, a_lo, b_hi, b_lo, 32, NB, kc > 0);
            umma::commit(&a_free[abuf]);
            if (kc == nchunks - 1) umma::commit(&acc_full);
        }
        ++ga;
        if (kc != nchunks - 1) return;
        // ---- epilogue ----
        umma::mbar_wait(&acc_full, nacc & 1);
        ++nacc;
        umma::fence_after_sync();
        {
            const int q = warp & 3, grp = warp >> 2;
            const int row = q * 32 + lane;
            const int gmr = m0 + row;
            int sb = 0;
            if (seg && gmr < n) { while (sb < nseg - 1 && gmr >= __ldg(seg_off + sb)) ++sb; }
            for (int c = grp; c * 32 < NB; c += 2) {
                if (grp_col0 + c * 32 >= co) break;
                float v[32];
                umma::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c * 32, v);
                if (gmr < n) {
                    const int gc0 = grp_col0 + c * 32;
                    const bool full = gc0 + 32 <= co && c * 32 + 32 <= NB;
                    if (seg) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (gc0 + i < co) v[i] += __ldg(seg + (size_t)sb * co + gc0 + i);
                    }
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {   // folded BatchNorm from shared memory (warp-uniform addresses: broadcast)
                        const float4 sc = *reinterpret_cast<const float4*>(s_scale + c * 32 + i), sh = *reinterpret_cast<const float4*>(s_shift + c * 32 + i);
                        v[i] = fmaf(v[i], sc.x, sh.x); v[i + 1] = fmaf(v[i + 1], sc.y, sh.y);
                        v[i + 2] = fmaf(v[i + 2], sc.z, sh.z); v[i + 3] = fmaf(v[i + 3], sc.w, sh.w);
                    }
                    if (R) {
                        const float* rs = R + (size_t)gmr * co + gc0;
                        if (full && ((reinterpret_cast<uintptr_t>(rs) & 15) == 0)) {
#pragma unroll
                            for (int i = 0; i < 32; i += 4) {
                                const float4 r4 = __ldg(reinterpret_cast<const float4*>(rs + i));
                                v[i] += r4.x; v[i + 1] += r4.y; v[i + 2] += r4.z; v[i + 3] += r4.w;
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (gc0 + i < co && c * 32 + i < NB) v[i] += __ldg(rs + i);
                        }
                    }
                    if (relu) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                    }
                    float* dst = Y + (size_t)gmr * ldy + gc0;
                    if (gc0 + 32 <= co && c * 32 + 32 <= NB && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (gc0 + i < co && c * 32 + i < NB) dst[i] = v[i];
                    }
                }
            }
        }
        umma::fence_before_sync();
        __syncthreads();   // accumulator drained before the next tile's first MMA overwrites it
        umma::fence_after_sync();
    };
#pragma unroll 1
    for (int g = 0; g < G; g += PF) {
        step(g, xv0);
        if (g + 1 < G) step(g + 1, xv1);
        if (g + 2 < G) step(g + 2, xv2);
    }
    if (warp == 0 && !w_ready) {   // CTA had no tile: still drain the weight copies before exiting
        umma::mbar_wait(&w_full, 0);
    }
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, tcols);
}

}  // namespace

// Tensor-core fused linear (weight-stationary).  Wc = [NG][2][Kpad/4][NB][4]: for each group of NB output columns the
// (hi, lo) TF32 split of W[cols, :] zero-padded to Kpad = 32*ceil(ci/32), in canonical K-major tiles
// (etch_b200/models/tc.py::tc_linear_weights chooses NB so that the tile fits shared memory).
ETCH_API int etch_linear_tc(const float* X, int ldx, const float* Wc, int NB, int n, int ci, int co, const float* scale,
                            const float* shift, const float* R, const float* seg, const int* seg_off, int nseg, int relu,
                            float* Y, int ldy, cudaStream_t stream) {
    if (!X || !Wc || !Y || n <= 0 || ci <= 0 || co <= 0 || NB < 16 || NB > 256 || (NB % 16)) return ETCH_EINVAL;
    const int Kpad = ((ci + 31) / 32) * 32;
    const int NG = (co + NB - 1) / NB;
    const size_t wbytes = (size_t)2 * NB * Kpad * 4;
    const size_t smem = (size_t)4 * LT_A + wbytes + 128;
    if (smem > 227 * 1024) return ETCH_EINVAL;
    int tcols = 32;
    while (tcols < NB) tcols <<= 1;
    ETCH_TRY(cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ntiles = (n + 127) / 128;
    int gx = etch_sm_budget() / NG;
    if (gx < 1) gx = 1;
    if (gx > ntiles) gx = ntiles;
    dim3 grid(gx, NG);
    linear_tc_kernel<<<grid, 256, smem, stream>>>(X, ldx, Wc, n, ci, co, Kpad, NB, scale, shift, R, seg, seg_off, nseg, relu, Y, ldy, tcols);
    ETCH_RETURN_LAST();
}
