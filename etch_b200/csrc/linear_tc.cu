// Generic tensor-core fused linear layer for the PointTransformer heads:
//     Y[n, co] = relu?( (X[n, ci] W^T + seg[segment(row)]) * scale + shift + R )
// i.e. nn.Linear / Conv1d(k=1) + eval BatchNorm1d (folded) + residual + ReLU in one launch
// (src/models/pointtransformer_seg.py:40-51,71-80,101-112,144-145,222).  Same contract as etch_linear (pt.cu), but the
// GEMM runs on tcgen05 (3xTF32, fp32-level accuracy) with the [128 x co] accumulator living in TMEM.
//
// One CTA = 128 rows.  K is walked in 64-wide chunks: all threads stage the X chunk (TF32 hi/lo split, canonical
// layout, double-buffered), warp 0 streams the pre-split 64x64 weight blocks through a 2-deep cp.async.bulk ring and
// issues the MMAs for every 64-column block of the output, so X is read exactly once.  Epilogue straight from TMEM.
#include "common.cuh"
#include "umma.cuh"

namespace {

constexpr uint32_t LT_A = 128 * 64 * 4;   // bytes of one (hi|lo) X chunk  [128 x 64]
constexpr uint32_t LT_B = 64 * 64 * 4;    // bytes of one (hi|lo) weight block [64 x 64]

__global__ void __launch_bounds__(256, 1) linear_tc_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ Wc,
                                                           int n, int ci, int co, int KC, int NC,
                                                           const float* __restrict__ scale, const float* __restrict__ shift,
                                                           const float* __restrict__ R, const float* __restrict__ seg,
                                                           const int* __restrict__ seg_off, int nseg, int relu,
                                                           float* __restrict__ Y, int ldy, int tcols) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* s_A = smem_raw;                 // [2 buffers][hi|lo]
    unsigned char* s_B = s_A + 4 * LT_A;           // [2 slots][hi|lo]
    __shared__ uint64_t a_free[2], b_full[2], b_empty[2], acc_full;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) umma::tmem_alloc(&tmem_base, tcols);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { umma::mbar_init(&a_free[i], 1); umma::mbar_init(&b_full[i], 1); umma::mbar_init(&b_empty[i], 1); }
        umma::mbar_init(&acc_full, 1);
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = umma::uniform(tmem_base);
    uint32_t ga = 0;            // A chunks staged so far (buffer = ga & 1), uniform across threads
    uint32_t gl = 0, gm = 0;    // weight blocks loaded / consumed (warp 0 only)
    uint32_t nacc = 0;          // tiles finished
    const int nblk = KC * NC;
    const int frow = tid >> 1, fhalf = tid & 1;

    for (int tile = blockIdx.x; tile * 128 < n; tile += gridDim.x) {
        const int m0 = tile * 128;
        int ld = 0;             // weight blocks requested for this tile (warp 0)
        if (warp == 0) {
            for (int i = 0; i < 2 && ld < nblk; ++i) {
                const uint32_t slot = gl & 1;
                if (gl >= 2) umma::mbar_wait(&b_empty[slot], ((gl - 2) >> 1) & 1);
                umma::bulk_load(s_B + slot * 2 * LT_B, Wc + (size_t)ld * 2 * 64 * 64, 2 * LT_B, &b_full[slot]);
                ++ld; ++gl;
            }
        }
        for (int kc = 0; kc < KC; ++kc) {
            const uint32_t abuf = ga & 1;
            if (ga >= 2) umma::mbar_wait(&a_free[abuf], ((ga - 2) >> 1) & 1);
            // ---- stage X[m0:m0+128, kc*64:(kc+1)*64] ----
            {
                const int gm_row = m0 + frow;
                const float* src = X + (size_t)(gm_row < n ? gm_row : 0) * ldx + kc * 64 + fhalf * 32;
                unsigned char* dh = s_A + abuf * 2 * LT_A;
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const int k = kc * 64 + fhalf * 32 + c4 * 4;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (gm_row < n) {
                        if (k + 3 < ci && ((reinterpret_cast<uintptr_t>(src + c4 * 4) & 15) == 0)) v = __ldg(reinterpret_cast<const float4*>(src + c4 * 4));
                        else {
                            if (k < ci) v.x = __ldg(src + c4 * 4);
                            if (k + 1 < ci) v.y = __ldg(src + c4 * 4 + 1);
                            if (k + 2 < ci) v.z = __ldg(src + c4 * 4 + 2);
                            if (k + 3 < ci) v.w = __ldg(src + c4 * 4 + 3);
                        }
                    }
                    float4 h, l;
                    umma::split_tf32(v.x, h.x, l.x); umma::split_tf32(v.y, h.y, l.y);
                    umma::split_tf32(v.z, h.z, l.z); umma::split_tf32(v.w, h.w, l.w);
                    const int kq = fhalf * 8 + c4;
                    *reinterpret_cast<float4*>(dh + kq * (128 * 16) + frow * 16) = h;
                    *reinterpret_cast<float4*>(dh + LT_A + kq * (128 * 16) + frow * 16) = l;
                }
            }
            umma::fence_async_smem();
            __syncthreads();
            if (warp == 0) {
                umma::fence_after_sync();
                const uint32_t a_hi = umma::smem_u32(s_A + abuf * 2 * LT_A), a_lo = a_hi + LT_A;
                for (int nc = 0; nc < NC; ++nc) {
                    const uint32_t slot = gm & 1;
                    umma::mbar_wait(&b_full[slot], (gm >> 1) & 1);
                    umma::fence_after_sync();
                    const uint32_t b_hi = umma::smem_u32(s_B + slot * 2 * LT_B), b_lo = b_hi + LT_B;
                    umma::issue_gemm_3xtf32(tmem + nc * 64, a_hi, a_lo, b_hi, b_lo, 64, 64, kc > 0);
                    umma::commit(&b_empty[slot]);
                    ++gm;
                    if (ld < nblk) {
                        const uint32_t s2 = gl & 1;
                        umma::mbar_wait(&b_empty[s2], ((gl - 2) >> 1) & 1);
                        umma::bulk_load(s_B + s2 * 2 * LT_B, Wc + (size_t)ld * 2 * 64 * 64, 2 * LT_B, &b_full[s2]);
                        ++ld; ++gl;
                    }
                }
                umma::commit(&a_free[abuf]);
                if (kc == KC - 1) umma::commit(&acc_full);
            }
            ++ga;
        }
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
            const int nchunk = NC * 2;   // 32-column chunks
            for (int c = grp; c < nchunk; c += 2) {
                if (c * 32 >= co) break;
                float v[32];
                umma::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c * 32, v);
                if (gmr < n) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int gc = c * 32 + i;
                        if (gc < co) {
                            float t = v[i];
                            if (seg) t += __ldg(seg + (size_t)sb * co + gc);
                            if (scale) t *= __ldg(scale + gc);
                            if (shift) t += __ldg(shift + gc);
                            if (R) t += __ldg(R + (size_t)gmr * co + gc);
                            if (relu) t = fmaxf(t, 0.f);
                            v[i] = t;
                        }
                    }
                    float* dst = Y + (size_t)gmr * ldy + c * 32;
                    if (c * 32 + 32 <= co && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    } else {
                        for (int i = 0; i < 32 && c * 32 + i < co; ++i) dst[i] = v[i];
                    }
                }
            }
        }
        umma::fence_before_sync();
        __syncthreads();   // accumulator drained before the next tile's first MMA overwrites it
        umma::fence_after_sync();
    }
    if (warp == 0) umma::tmem_dealloc(tmem, tcols);
}

}  // namespace

// Tensor-core fused linear. Wc = [KC][NC][2][16][64][4]: 64x64 blocks of W (rows = outputs, zero padded), TF32 hi/lo split,
// canonical K-major tiles (etch_b200/models/tc.py::tc_linear_weights); KC = ceil(ci/64), NC = ceil(co/64) <= 8.
ETCH_API int etch_linear_tc(const float* X, int ldx, const float* Wc, int n, int ci, int co, const float* scale,
                            const float* shift, const float* R, const float* seg, const int* seg_off, int nseg, int relu,
                            float* Y, int ldy, cudaStream_t stream) {
    if (!X || !Wc || !Y || n <= 0 || ci <= 0 || co <= 0 || co > 512) return ETCH_EINVAL;
    const int KC = (ci + 63) / 64, NC = (co + 63) / 64;
    int tcols = 32;
    while (tcols < NC * 64) tcols <<= 1;
    const size_t smem = (size_t)4 * LT_A + (size_t)4 * LT_B + 128;
    ETCH_TRY(cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = (n + 127) / 128;
    if (grid > 148) grid = 148;
    linear_tc_kernel<<<grid, 256, smem, stream>>>(X, ldx, Wc, n, ci, co, KC, NC, scale, shift, R, seg, seg_off, nseg, relu, Y, ldy, tcols);
    ETCH_RETURN_LAST();
}
