/*
 * etch_b200_probes.h -- self-test and measurement entry points of libetch_b200.so.  NOT part of the product ABI (nothing on
 * the hot path calls them): tests/test_umma_gpu.py validates the tcgen05 building blocks through etch_umma_selftest, and
 * tools/microbench.py / tools/mma_probe.py use the probes whose numbers DESIGN.md quotes.  Same conventions as etch_b200.h.
 */
#ifndef ETCH_B200_PROBES_H
#define ETCH_B200_PROBES_H

#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

/* etch_lm_fit plus SM-clock totals per (scan, phase) in prof (profiling aid for tools/microbench.py). */
int etch_lm_fit_profile(const float* markers, const unsigned char* valid, const float* Tm, const float* Sm, const float* Pm,
                        const float* Wm, const float* Jt, const float* Js, const int* parents, const unsigned* ancmask, int B,
                        int M, int steps0, int steps1, float step0, float step1, float damp0, float damp1, float* params,
                        int* iters, float* errs, long long* prof, cudaStream_t stream);

/* tcgen05 building-block self test: C[128,N] = A[128,K] B[N,K]^T (3xTF32); and issue/copy latency probe (tests/, tools/). */
int etch_umma_selftest(const float* A, const float* B, float* C, int K, int N, cudaStream_t stream);
int etch_umma_latency(const float* src, long long* out, cudaStream_t stream);
/* probe: SM cycles for `iters` x 8 warp-level mma.sync.m16n8k8 TF32 per warp, `warps` warps per CTA; out[ctas] */
int etch_mma_sync_rate(long long* out, int ctas, int warps, int iters, cudaStream_t stream);
/* probe: SM cycles for `iters` x 32 FP32 FMAs per thread; mode 0 = scalar FFMA, 1 = packed fma.rn.f32x2 */
int etch_ffma_rate(long long* out, int ctas, int warps, int iters, int mode, cudaStream_t stream);

/* probe: SM cycles for `steps` x (N = 48 + N = 24) tf32 MMAs at M = 64 reading a ring of distinct shared-memory operand tiles -- the tensor-core
 * cost of the M = c_in mapping of the InterSO3Conv neighbour contraction that DESIGN.md argues against (mode 1: M = 128; mode 2: N = 128 + 64) */
int etch_umma_contract_probe(long long* out, int ctas, int steps, int ring, int mode, cudaStream_t stream);

/* test hook: residual res [B,3M] and analytic Jacobian jac [B,3M,85] (columns orient 3 | pose 69 | betas 10 | transl 3) of the marker cost that
 * etch_lm_fit minimises, at the given packed parameters [B,85] (tests/test_fit_gpu.py compares them with the oracle's autodiff Jacobian) */
int etch_lm_jacobian_dump(const float* params, const float* markers, const unsigned char* valid, const float* Tm, const float* Sm,
                          const float* Pm, const float* Wm, const float* Jt, const float* Js, const int* parents,
                          const unsigned* ancmask, int B, int M, float* res, float* jac, cudaStream_t stream);

/* probe: reads and resets the diagnostic counters of etch_knn_grid (HOST pointer to 2 values: fallback queries, candidates visited); the
 * counters only tick in a library built with -DETCH_KNN_STATS */
int etch_knn_grid_stats(unsigned long long* out_host);

#ifdef __cplusplus
}
#endif
#endif /* ETCH_B200_PROBES_H */
