// TEST INFRASTRUCTURE (oracle/): raw-pointer C launchers over the REFERENCE's own, unmodified CUDA kernels.
// oracle/build_ref.sh compiles the reference sources where they lie under /root/reference (with oracle/refshim standing in
// for the two ATen headers) and links them with this file into oracle/_ref/libetch_ref_kernels.so.  Only tests/ and the
// reference-kernel microbench (tools/ref_kernel_bench.py) load that library -- as the checker and as "the number to beat",
// never on the product path.  Each wrapper restates the allocation/initialisation its pybind caller does:
//   ref_ball_query              <- epn_grouping.ball_query              grouping_cuda.cpp:71-86   (idx zero-filled)
//   ref_furthest_point_sampling <- epn_grouping.furthest_point_sampling grouping_cuda.cpp:160-174 (temp = 1e10, idx = 0)
//   ref_gather_points_forward   <- epn_gathering.gather_points_forward  gathering_cuda.cpp:29-43  (out zero-filled)
//   ref_knnquery                <- pointops_cuda.knnquery_cuda          knnquery_cuda.cpp:8-17
//   ref_furthestsampling        <- pointops_cuda.furthestsampling_cuda  sampling_cuda.cpp:8-16    (tmp = 1e10 by the caller, pointops.py:21-23)
// All kernels run on the legacy default stream, as in the reference; every wrapper returns cudaDeviceSynchronize()'s status.
#include <ATen/ATen.h>
#include <cuda_runtime.h>

at::Tensor ball_query_cuda(int b, int n, int m, float radius, int nsample, at::Tensor new_xyz, at::Tensor xyz, at::Tensor idx);
at::Tensor furthest_point_sampling_cuda(at::Tensor source, at::Tensor temp, at::Tensor sampled_idx, const int m);
at::Tensor gather_points_forward_cuda(at::Tensor support_points, at::Tensor grouped_indices, at::Tensor grouped_points);
extern "C" void knnquery_cuda_launcher(int m, int nsample, const float* xyz, const float* new_xyz, const int* offset,
                                       const int* new_offset, int* idx, float* dist2);
extern "C" void furthestsampling_cuda_launcher(int b, int n, const float* xyz, const int* offset, const int* new_offset,
                                               float* tmp, int* idx);

namespace {
__global__ void fill_f32(float* p, long n, float v) {
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
int done() {
  cudaError_t e = cudaGetLastError();
  cudaError_t s = cudaDeviceSynchronize();
  return e != cudaSuccess ? (int)e : (int)s;
}
}  // namespace

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API int ref_ball_query(const float* new_xyz, const float* xyz, int b, int m, int n, float radius, int nsample, int* idx) {
  cudaMemset(idx, 0, sizeof(int) * (size_t)b * m * nsample);
  ball_query_cuda(b, n, m, radius, nsample, at::Tensor((void*)new_xyz, {b, 3, m}), at::Tensor((void*)xyz, {b, 3, n}),
                  at::Tensor(idx, {b, m, nsample}));
  return done();
}

REF_API int ref_furthest_point_sampling(const float* xyz, int b, int n, int m, float* temp, int* idx) {
  long tot = (long)b * n;
  fill_f32<<<(unsigned)((tot + 255) / 256), 256>>>(temp, tot, 1e10f);
  cudaMemset(idx, 0, sizeof(int) * (size_t)b * m);
  furthest_point_sampling_cuda(at::Tensor((void*)xyz, {b, 3, n}), at::Tensor(temp, {b, n}), at::Tensor(idx, {b, m}), m);
  return done();
}

REF_API int ref_gather_points_forward(const float* points, const int* idx, int b, int c, int n, int m, float* out) {
  cudaMemset(out, 0, sizeof(float) * (size_t)b * c * m);
  gather_points_forward_cuda(at::Tensor((void*)points, {b, c, n}), at::Tensor((void*)idx, {b, m}), at::Tensor(out, {b, c, m}));
  return done();
}

REF_API int ref_knnquery(int m, int nsample, const float* xyz, const float* new_xyz, const int* offset, const int* new_offset,
                         int* idx, float* dist2) {
  knnquery_cuda_launcher(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2);
  return done();
}

REF_API int ref_furthestsampling(int b, int n_max, int n_total, const float* xyz, const int* offset, const int* new_offset,
                                 float* tmp, int* idx) {
  fill_f32<<<(unsigned)((n_total + 255) / 256), 256>>>(tmp, n_total, 1e10f);
  furthestsampling_cuda_launcher(b, n_max, xyz, offset, new_offset, tmp, idx);
  return done();
}

// Timing variants for the per-kernel "reference vs ours" table: launch `reps` times back to back between two CUDA events on
// the default stream (initialisation included, as the reference wrapper pays it on every call) and return ms per call.
REF_API float ref_time_ball_query(const float* new_xyz, const float* xyz, int b, int m, int n, float radius, int nsample,
                                  int* idx, int reps) {
  cudaEvent_t a, z; cudaEventCreate(&a); cudaEventCreate(&z);
  cudaEventRecord(a);
  for (int r = 0; r < reps; ++r) {
    cudaMemsetAsync(idx, 0, sizeof(int) * (size_t)b * m * nsample);
    ball_query_cuda(b, n, m, radius, nsample, at::Tensor((void*)new_xyz, {b, 3, m}), at::Tensor((void*)xyz, {b, 3, n}),
                    at::Tensor(idx, {b, m, nsample}));
  }
  cudaEventRecord(z); cudaEventSynchronize(z);
  float ms = 0; cudaEventElapsedTime(&ms, a, z); cudaEventDestroy(a); cudaEventDestroy(z);
  return ms / reps;
}

REF_API float ref_time_furthest_point_sampling(const float* xyz, int b, int n, int m, float* temp, int* idx, int reps) {
  cudaEvent_t a, z; cudaEventCreate(&a); cudaEventCreate(&z);
  long tot = (long)b * n;
  cudaEventRecord(a);
  for (int r = 0; r < reps; ++r) {
    fill_f32<<<(unsigned)((tot + 255) / 256), 256>>>(temp, tot, 1e10f);
    cudaMemsetAsync(idx, 0, sizeof(int) * (size_t)b * m);
    furthest_point_sampling_cuda(at::Tensor((void*)xyz, {b, 3, n}), at::Tensor(temp, {b, n}), at::Tensor(idx, {b, m}), m);
  }
  cudaEventRecord(z); cudaEventSynchronize(z);
  float ms = 0; cudaEventElapsedTime(&ms, a, z); cudaEventDestroy(a); cudaEventDestroy(z);
  return ms / reps;
}

REF_API float ref_time_knnquery(int m, int nsample, const float* xyz, const float* new_xyz, const int* offset,
                                const int* new_offset, int* idx, float* dist2, int reps) {
  cudaEvent_t a, z; cudaEventCreate(&a); cudaEventCreate(&z);
  cudaEventRecord(a);
  for (int r = 0; r < reps; ++r) knnquery_cuda_launcher(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2);
  cudaEventRecord(z); cudaEventSynchronize(z);
  float ms = 0; cudaEventElapsedTime(&ms, a, z); cudaEventDestroy(a); cudaEventDestroy(z);
  return ms / reps;
}

REF_API float ref_time_furthestsampling(int b, int n_max, int n_total, const float* xyz, const int* offset,
                                        const int* new_offset, float* tmp, int* idx, int reps) {
  cudaEvent_t a, z; cudaEventCreate(&a); cudaEventCreate(&z);
  cudaEventRecord(a);
  for (int r = 0; r < reps; ++r) {
    fill_f32<<<(unsigned)((n_total + 255) / 256), 256>>>(tmp, n_total, 1e10f);
    furthestsampling_cuda_launcher(b, n_max, xyz, offset, new_offset, tmp, idx);
  }
  cudaEventRecord(z); cudaEventSynchronize(z);
  float ms = 0; cudaEventElapsedTime(&ms, a, z); cudaEventDestroy(a); cudaEventDestroy(z);
  return ms / reps;
}
