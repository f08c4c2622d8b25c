// TEST INFRASTRUCTURE (oracle/): a 40-line stand-in for the part of <ATen/ATen.h> that the reference's CUDA sources touch,
// so that oracle/build_ref.sh can compile them UNMODIFIED, from where they lie under /root/reference, without linking torch:
//   at::Tensor::{data<T>(), size(i), type()} and AT_DISPATCH_FLOATING_TYPES  (grouping_cuda_kernel.cu:471-739,
//   gathering_cuda_kernel.cu:100-166).  A "tensor" here is a raw device pointer plus its sizes; every tensor is float32 or
//   int32, so the dispatch macro instantiates the lambda with scalar_t = float (the only type the hot path ever uses).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>
namespace at {
struct DeprecatedTypeProperties {};
class Tensor {
 public:
  Tensor() : p_(nullptr), nd_(0) { s_[0] = s_[1] = s_[2] = s_[3] = s_[4] = 0; }
  Tensor(void* p, std::initializer_list<int64_t> sizes) : p_(p), nd_(0) {
    for (int64_t v : sizes) s_[nd_++] = v;
  }
  template <typename T> T* data() const { return static_cast<T*>(p_); }
  template <typename T> T* data_ptr() const { return static_cast<T*>(p_); }
  int64_t size(int i) const { return s_[i]; }
  DeprecatedTypeProperties type() const { return DeprecatedTypeProperties(); }
 private:
  void* p_;
  int nd_;
  int64_t s_[5];
};
}  // namespace at
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...) \
  do { (void)(TYPE); using scalar_t = float; __VA_ARGS__(); } while (0)
