// Shared helpers for the etch_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define ETCH_OK 0
#define ETCH_EINVAL (-1)       // bad argument (null pointer, size out of the supported range)
#define ETCH_EUNSUPPORTED (-2) // reference entry point that is not on the hot path (stub)

#define ETCH_API extern "C" __attribute__((visibility("default")))

#define ETCH_RETURN_LAST()                         \
    do {                                           \
        cudaError_t e__ = cudaGetLastError();      \
        return e__ == cudaSuccess ? ETCH_OK : (int)e__; \
    } while (0)

#define ETCH_TRY(x)                                 \
    do {                                            \
        cudaError_t e__ = (x);                      \
        if (e__ != cudaSuccess) return (int)e__;    \
    } while (0)

// dx*dx + dy*dy + dz*dz exactly as nvcc contracts the reference expression `(a)*(a) + (b)*(b) + (c)*(c)`: the SECOND product
// is the plain FMUL and the first is fused onto it -- FMUL(dy,dy), FFMA(dx,dx,.), FFMA(dz,dz,.).  Read off the SASS of the
// reference's own kernels built for sm_100a (oracle/_ref: knnquery_cuda_kernel.cu:94, sampling_cuda_kernel.cu:53,
// grouping_cuda_kernel.cu:93,389-390) and pinned by executing them (tests/test_ref_kernels_gpu.py).
__device__ __forceinline__ float etch_sqdist3(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__device__ __forceinline__ float etch_lrelu(float x) { return x > 0.f ? x : 0.01f * x; }

template <typename T>
__host__ __device__ __forceinline__ T etch_cdiv(T a, T b) { return (a + b - 1) / b; }

// Host: block size rule of the reference launchers (grouping_cuda_kernel.cu:29-33, pointops cuda_utils.h:10-13).
static inline int etch_opt_n_threads(int work_size) {
    const int pow_2 = (int)(log((double)work_size) / log(2.0));
    int t = 1 << pow_2;
    if (t > 1024) t = 1024;
    if (t < 1) t = 1;
    return t;
}

// SMs the persistent (1 CTA/SM) kernels may fill; 148 on B200 unless etch_set_sm_budget lowered it so that a concurrent
// small-grid kernel of another stream (the LM fit of the previous batch) keeps its SMs.  Defined in index.cu.
int etch_sm_budget();

// 256-bit read-only global load (sm_100: LDG.E.256); p must be 32-byte aligned
__device__ __forceinline__ void etch_ldg256(const float* p, float (&v)[8]) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}

// 256-bit global load that does not allocate in L1 (data read once per CTA: with a 200+ KB shared-memory carve-out the few L1
// lines that are left throttle the LSU queue; see the direction-head blend in heads_tc.cu)
__device__ __forceinline__ void etch_ld256_na(const float* p, float (&v)[8]) {
    asm volatile("ld.global.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
