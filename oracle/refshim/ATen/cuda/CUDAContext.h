// TEST INFRASTRUCTURE (oracle/): stand-in for <ATen/cuda/CUDAContext.h> (included, unused, by the pointops kernel headers).
#pragma once
#include <cuda_runtime.h>
