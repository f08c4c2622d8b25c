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

// ---------------------------------------------------------------------------------------------------------------------
// Warp-level mma.sync (m16n8k8 TF32) issue-rate probe: every warp runs `iters` rounds of 8 independent accumulator
// chains; out[cta] = SM cycles of the CTA.  Used to size the per-(point, anchor) neighbour contraction of the inter conv.
namespace {
__global__ void __launch_bounds__(1024) mma_sync_rate_kernel(long long* __restrict__ out, int iters, float seed) {
    float c[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f; }
    uint32_t a[4], b[2];
    a[0] = __float_as_uint(seed + threadIdx.x); a[1] = a[0] ^ 0x100; a[2] = a[0] ^ 0x200; a[3] = a[0] ^ 0x300;
    b[0] = a[0] ^ 0x400; b[1] = a[0] ^ 0x500;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    __syncthreads();
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (s == 123.456f) out[blockIdx.x] = 0;
}
}  // namespace

ETCH_API int etch_mma_sync_rate(long long* out, int ctas, int warps, int iters, cudaStream_t stream) {
    if (!out || ctas <= 0 || warps <= 0 || warps > 32 || iters <= 0) return ETCH_EINVAL;
    mma_sync_rate_kernel<<<ctas, warps * 32, 0, stream>>>(out, iters, 1.0f);
    ETCH_RETURN_LAST();
}

// FP32 issue-rate probe: mode 0 = scalar FFMA (16 independent chains), mode 1 = packed fma.rn.f32x2 (16 chains of 2)
namespace {
template <int MODE>
__global__ void __launch_bounds__(1024) ffma_rate_kernel(long long* __restrict__ out, int iters, float seed) {
    float x = seed + threadIdx.x * 1e-3f, y = 1.0f - x * 1e-3f;
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = j * 0.5f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = fmaf(acc[j], x, y);
        } else {
            unsigned long long xx, yy;
            asm("mov.b64 %0, {%1, %2};" : "=l"(xx) : "f"(x), "f"(x));
            asm("mov.b64 %0, {%1, %2};" : "=l"(yy) : "f"(y), "f"(y));
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                unsigned long long a;
                asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(acc[j]), "f"(acc[j + 1]));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(xx), "l"(yy));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[j]), "=f"(acc[j + 1]) : "l"(a));
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) s += acc[j];
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (s == 123.456f) out[blockIdx.x] = 0;
}
}  // namespace

ETCH_API int etch_ffma_rate(long long* out, int ctas, int warps, int iters, int mode, cudaStream_t stream) {
    if (!out || ctas <= 0 || warps <= 0 || warps > 32 || iters <= 0) return ETCH_EINVAL;
    if (mode == 0) ffma_rate_kernel<0><<<ctas, warps * 32, 0, stream>>>(out, iters, 1.0f);
    else ffma_rate_kernel<1><<<ctas, warps * 32, 0, stream>>>(out, iters, 1.0f);
    ETCH_RETURN_LAST();
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Probe for DESIGN.md's "why the neighbour contraction stays on the FP32 pipes": the tensor-core cost of the proposed M = c_in mapping,
//   D_a[64 x 24] (+)= F_a[64 x 8] * W_a[24 x 8]^T   per (anchor, 8-neighbour K step), 3xTF32 = one N = 48 MMA (A_hi x [W_hi; W_lo])
//   + one N = 24 MMA (A_lo x W_hi),
// with every step reading a DIFFERENT operand tile from shared memory (a ring of `ring` (A_hi, A_lo, B) tile sets), as the real kernel
// would.  One CTA per SM; out[cta] = SM cycles for `steps` steps (both MMAs), measured from first issue to the commit's arrival.
// mode 0: M = 64, N = 48 + 24 (the proposal); mode 1: M = 128, N = 48 + 24 (two anchors' rows, if they could share B); mode 2: M = 64,
// N = 128 + 64 (the channel-mixing GEMM of the current kernel at c_out = 64, for scale).
namespace {
constexpr int CP_RING = 16;
template <int M, int N1, int N2>
__global__ void __launch_bounds__(128, 1) umma_contract_probe_kernel(long long* __restrict__ out, int steps) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr uint32_t a_bytes = (uint32_t)M * 8 * 4, b_bytes = (uint32_t)N1 * 8 * 4;
    constexpr uint32_t set_bytes = 2 * a_bytes + b_bytes;
    for (uint32_t i = tid; i < (uint32_t)CP_RING * set_bytes / 4; i += 128) reinterpret_cast<float*>(smem_raw)[i] = 1.0f / (float)(1 + (i & 1023));
    if (warp == 0) umma::tmem_alloc(&tmem_base, 512);
    if (tid == 0) umma::mbar_init(&bar, 1);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = umma::uniform(tmem_base);
    if (warp == 0) {
        constexpr uint32_t idesc1 = umma::make_idesc_tf32(M, N1), idesc2 = umma::make_idesc_tf32(M, N2);
        const uint32_t base = umma::smem_u32(smem_raw);
        // descriptors of tile set 0; a set is set_bytes further on: only the 14-bit start-address field moves (as in the product kernels)
        const uint64_t dah0 = umma::make_desc(base, (uint32_t)M * 16, 128), dal0 = umma::make_desc(base + a_bytes, (uint32_t)M * 16, 128);
        const uint64_t db0 = umma::make_desc(base + 2 * a_bytes, (uint32_t)N1 * 16, 128);
        constexpr uint64_t inc = set_bytes >> 4;
        const long long t0 = clock64();
#pragma unroll 8
        for (int s = 0; s < steps; ++s) {
            const uint64_t o = (uint64_t)(s & (CP_RING - 1)) * inc;
            const uint32_t d = tmem + (uint32_t)((s & 1) * 256);       // alternate accumulators so consecutive steps do not serialise on D
            umma::mma_tf32(d, dah0 + o, db0 + o, idesc1, s >= 2 ? 1u : 0u);
            umma::mma_tf32(d, dal0 + o, db0 + o, idesc2, 1u);
        }
        umma::commit(&bar);
        umma::mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (umma::elect_one()) out[blockIdx.x] = t1 - t0;
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

template <int M, int N1, int N2>
int launch_contract_probe(long long* out, int ctas, int steps, cudaStream_t stream) {
    constexpr size_t smem = (size_t)CP_RING * (2 * M * 32 + N1 * 32) + 1024;
    static_assert(smem <= 200 * 1024, "probe ring");
    auto kern = umma_contract_probe_kernel<M, N1, N2>;
    ETCH_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<ctas, 128, smem, stream>>>(out, steps);
    ETCH_RETURN_LAST();
}
}  // namespace

// mode 0: M = 64, N = 48 + 24 (the proposal); 1: M = 128, N = 48 + 24; 2: M = 64, N = 128 + 64 (channel-mixing GEMM of the shipping kernel,
// c_out = 64); 3: M = 128, N = 128 + 64; `ring` is fixed at 16 tile sets (argument kept for the ABI, must be 16)
ETCH_API int etch_umma_contract_probe(long long* out, int ctas, int steps, int ring, int mode, cudaStream_t stream) {
    if (!out || ctas <= 0 || steps <= 0 || ring != CP_RING || mode < 0 || mode > 3) return ETCH_EINVAL;
    if (mode == 0) return launch_contract_probe<64, 48, 24>(out, ctas, steps, stream);
    if (mode == 1) return launch_contract_probe<128, 48, 24>(out, ctas, steps, stream);
    if (mode == 2) return launch_contract_probe<64, 128, 64>(out, ctas, steps, stream);
    return launch_contract_probe<128, 128, 64>(out, ctas, steps, stream);
}
