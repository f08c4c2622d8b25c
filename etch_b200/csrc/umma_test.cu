// Self-test of the tcgen05 building blocks (umma.cuh): C[128 x N] = A[128 x K] * B[N x K]^T with the 3xTF32 split.
// Used by tests/test_umma_gpu.py to pin descriptor / layout / TMEM conventions against a float64 reference.
#include "common.cuh"
#include "umma.cuh"

namespace {

__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                               float* __restrict__ C, int K, int N) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    float* a_hi = reinterpret_cast<float*>(smem_raw);
    float* a_lo = a_hi + 128 * K;
    float* b_hi = a_lo + 128 * K;
    float* b_lo = b_hi + N * K;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * K; i += 128) {
        const int r = i / K, k = i % K;
        float h, l;
        umma::split_tf32(A[i], h, l);
        a_hi[umma::tile_off(r, k, 128) / 4] = h;
        a_lo[umma::tile_off(r, k, 128) / 4] = l;
    }
    for (int i = tid; i < N * K; i += 128) {
        const int r = i / K, k = i % K;
        float h, l;
        umma::split_tf32(B[i], h, l);
        b_hi[umma::tile_off(r, k, N) / 4] = h;
        b_lo[umma::tile_off(r, k, N) / 4] = l;
    }
    int ncols = 32;
    while (ncols < N) ncols <<= 1;
    if (warp == 0) umma::tmem_alloc(&tmem_base, ncols);
    if (tid == 0) umma::mbar_init(&bar, 1);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = umma::uniform(tmem_base);
    if (warp == 0) {
        umma::issue_gemm_3xtf32(tmem, umma::smem_u32(a_hi), umma::smem_u32(a_lo), umma::smem_u32(b_hi), umma::smem_u32(b_lo), K, N, false);
        umma::commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    for (int c = 0; c < N; c += 8) {
        float v[8];
        umma::tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) C[(size_t)tid * N + c + j] = v[j];
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, ncols);
}

// cycle counts of the primitives the tensor-core kernels are built from (single CTA, thread 0 timestamps)
__global__ void __launch_bounds__(128, 1) umma_latency_kernel(const float* __restrict__ src, long long* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) umma::tmem_alloc(&tmem_base, 64);
    if (tid == 0) { umma::mbar_init(&bar, 1); umma::mbar_init(&bar2, 1); }
    for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem_raw)[i] = 0.f;
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = umma::uniform(tmem_base);
    uint32_t ph = 0, ph2 = 0;
    int o = 0;
    for (int rep = 0; rep < 3; ++rep) {
        for (int kb = 8; kb <= 64; kb *= 2) {   // bulk load latency for 8..64 KB (second+ reps hit L2)
            long long t0 = clock64();
            if (warp == 0) { umma::bulk_load(smem_raw + 96 * 1024, src, kb * 1024, &bar); umma::mbar_wait(&bar, ph & 1); }
            ph++;
            __syncthreads();
            if (tid == 0) out[o] = clock64() - t0;
            o++;
        }
        for (int n = 32; n <= 64; n *= 2) {      // 24 MMAs (K=64, 3xTF32) + commit + wait
            long long t0 = clock64();
            if (warp == 0) {
                const uint32_t a = umma::smem_u32(smem_raw);
                umma::issue_gemm_3xtf32(tmem, a, a + 32768, a + 65536, a + 65536 + 16384, 64, n, false);
                umma::commit(&bar2);
                if (tid == 0) out[o] = clock64() - t0;   // issue cost only
                umma::mbar_wait(&bar2, ph2 & 1);
            }
            ph2++;
            __syncthreads();
            if (tid == 0) out[o + 1] = clock64() - t0;
            o += 2;
        }
        {
            long long t0 = clock64();
            umma::fence_async_smem();
            __syncthreads();
            if (tid == 0) out[o] = clock64() - t0;
            o++;
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 64);
}

}  // namespace

ETCH_API int etch_umma_latency(const float* src, long long* out, cudaStream_t stream) {
    if (!src || !out) return ETCH_EINVAL;
    ETCH_TRY(cudaFuncSetAttribute(umma_latency_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    umma_latency_kernel<<<1, 128, 160 * 1024, stream>>>(src, out);
    ETCH_RETURN_LAST();
}

// C[128,N] = A[128,K] B[N,K]^T on the 5th-gen tensor cores (3xTF32). K % 8 == 0, N % 16 == 0, N <= 256.
ETCH_API int etch_umma_selftest(const float* A, const float* B, float* C, int K, int N, cudaStream_t stream) {
    if (!A || !B || !C || K <= 0 || K % 8 || N < 16 || N % 16 || N > 256) return ETCH_EINVAL;
    const size_t smem = (size_t)(2 * 128 * K + 2 * N * K) * 4;
    if (smem > 200 * 1024) return ETCH_EINVAL;
    ETCH_TRY(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_selftest_kernel<<<1, 128, smem, stream>>>(A, B, C, K, N);
    ETCH_RETURN_LAST();
}
