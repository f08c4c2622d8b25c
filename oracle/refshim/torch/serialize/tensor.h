// TEST INFRASTRUCTURE (oracle/): stand-in for <torch/serialize/tensor.h>; the pointops kernel headers only need the name
// at::Tensor to declare (never define or call) their torch-facing wrappers (knnquery_cuda_kernel.h:7, sampling_cuda_kernel.h:7).
#pragma once
#include <ATen/ATen.h>
