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

constexpr int CT_N = 32;    // weight columns (GEMM N) per chunk
constexpr int CT_K = 128;   // feature dimension

__global__ void __launch_bounds__(256, 1) conf_head_tc_kernel(const float* __restrict__ x,       // [n][128]
                                                              const float* __restrict__ logits,  // [n][K]
                                                              const float* __restrict__ W0c,     // [K*4][2][32][32][4]
                                                              const float* __restrict__ b0,      // [K*128]
                                                              const float* __restrict__ w2,      // [K][128]
                                                              const float* __restrict__ b2,      // [K]
                                                              int n, int K, float* __restrict__ conf) {
    constexpr uint32_t A_BYTES = 128 * CT_K * 4;      // 64 KB per (hi|lo)
    constexpr uint32_t B_BYTES = CT_N * CT_K * 4;     // 16 KB per (hi|lo)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* s_A = smem_raw;                    // [hi | lo]
    unsigned char* s_B = s_A + 2 * A_BYTES;           // [2 buffers][hi | lo]
    __shared__ uint64_t b_full[2], b_empty[2], t_full[2], t_empty[2];
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) umma::tmem_alloc(&tmem_base, 64);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            umma::mbar_init(&b_full[i], 1); umma::mbar_init(&b_empty[i], 1);
            umma::mbar_init(&t_full[i], 1); umma::mbar_init(&t_empty[i], 128);
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_base;
    const int nchunk = K * 4;
    uint32_t gi = 0;  // global chunk counter (barrier phases run across tiles)

    for (int tile = blockIdx.x; tile * 128 < n; tile += gridDim.x) {
        const int m0 = tile * 128;
        // ---- fill the A tile: 2 threads per row, TF32 split, canonical layout ----
        {
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

        if (tid == 0) {
            // ===== weight streaming + MMA issue =====
            const uint32_t a_hi = umma::smem_u32(s_A), a_lo = a_hi + A_BYTES;
            {   // first slice of the tile: its buffer was last read by MMA (gi-2), which completed before the tile barrier
                const uint32_t g = gi;
                if (g >= 2) umma::mbar_wait(&b_empty[g & 1], ((g - 2) >> 1) & 1);
                umma::bulk_load(s_B + (g & 1) * 2 * B_BYTES, W0c, 2 * B_BYTES, &b_full[g & 1]);
            }
            for (int i = 0; i < nchunk; ++i) {
                const uint32_t g = gi + i, buf = g & 1;
                if (i + 1 < nchunk) {
                    const uint32_t g1 = g + 1, nb = g1 & 1;
                    if (g1 >= 2) umma::mbar_wait(&b_empty[nb], ((g1 - 2) >> 1) & 1);
                    umma::bulk_load(s_B + nb * 2 * B_BYTES, W0c + (size_t)(i + 1) * 2 * CT_N * CT_K, 2 * B_BYTES, &b_full[nb]);
                }
                umma::mbar_wait(&b_full[buf], (g >> 1) & 1);
                if (g >= 2) umma::mbar_wait(&t_empty[buf], ((g - 2) >> 1) & 1);
                umma::fence_after_sync();
                const uint32_t b_hi = umma::smem_u32(s_B + buf * 2 * B_BYTES), b_lo = b_hi + B_BYTES;
                umma::issue_gemm_3xtf32(tmem + buf * CT_N, a_hi, a_lo, b_hi, b_lo, CT_K, CT_N, false);
                umma::commit(&b_empty[buf]);
                umma::commit(&t_full[buf]);
            }
        } else if (warp >= 4) {
            // ===== epilogue: one thread per point =====
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
            float acc_conf = 0.f, part = 0.f;
            for (int i = 0; i < nchunk; ++i) {
                const uint32_t g = gi + i, buf = g & 1;
                umma::mbar_wait(&t_full[buf], (g >> 1) & 1);
                umma::fence_after_sync();
                float v[32];
                umma::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + buf * CT_N, v);
                umma::fence_before_sync();
                umma::mbar_arrive(&t_empty[buf]);
                const float4* bp = reinterpret_cast<const float4*>(b0 + (size_t)i * CT_N);
                const float4* wp = reinterpret_cast<const float4*>(w2 + (size_t)i * CT_N);
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 bb = __ldg(bp + j4), ww = __ldg(wp + j4);
                    part = fmaf(fmaxf(v[j4 * 4 + 0] + bb.x, 0.f), ww.x, part);
                    part = fmaf(fmaxf(v[j4 * 4 + 1] + bb.y, 0.f), ww.y, part);
                    part = fmaf(fmaxf(v[j4 * 4 + 2] + bb.z, 0.f), ww.z, part);
                    part = fmaf(fmaxf(v[j4 * 4 + 3] + bb.w, 0.f), ww.w, part);
                }
                if ((i & 3) == 3) {  // marker group finished
                    const int l = i >> 2;
                    const float p = expf(__ldg(lr + l) - mx) * inv_den;
                    acc_conf = fmaf(p, part + __ldg(b2 + l), acc_conf);
                    part = 0.f;
                }
            }
            if (ok) conf[m] = acc_conf;
        }
        gi += nchunk;
        umma::fence_before_sync();
        __syncthreads();   // A tile and accumulators are free for the next tile
        umma::fence_after_sync();
    }
    if (warp == 0) umma::tmem_dealloc(tmem, 64);
}

}  // namespace

// Tensor-core confidence head. W0c = [K*4][2][32][32][4]: per 32-column slice of confi.0's weight, (hi, lo) canonical tiles.
ETCH_API int etch_conf_head_tc(const float* x, const float* logits, const float* W0c, const float* b0, const float* w2,
                               const float* b2, int n, int K, float* conf, cudaStream_t stream) {
    if (!x || !logits || !W0c || !b0 || !w2 || !b2 || !conf || n <= 0 || K <= 0) return ETCH_EINVAL;
    const size_t smem = (size_t)2 * 128 * CT_K * 4 + (size_t)4 * CT_N * CT_K * 4 + 128;
    ETCH_TRY(cudaFuncSetAttribute(conf_head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = etch_cdiv(n, 128);
    if (grid > 148) grid = 148;
    conf_head_tc_kernel<<<grid, 256, smem, stream>>>(x, logits, W0c, b0, w2, b2, n, K, conf);
    ETCH_RETURN_LAST();
}
