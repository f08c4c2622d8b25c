"""etch_b200 -- B200-native (sm_100a) implementation of ETCH's per-scan inference-and-fit hot path.

Layout:
  csrc/            hand-written CUDA kernels + the C ABI (include/etch_b200.h) -> libetch_b200.so
  _lib.py          ctypes binding of the C ABI (fails loudly if the library is missing; there is NO CPU fallback)
  ext/             drop-in modules named like the reference's native extensions
                   (epn_grouping, epn_gathering, epn_zpconv, pointops_cuda)
  models/          host-side mirror of the reference operator API (src/models): GT_network_equiv, fit_smpl
  smpl_model.py    SMPL parameter loading (real pkl) / synthetic SMPL-shaped bodies
  synth.py         seeded synthetic scans and reference-layout checkpoints for tests and bench
"""
__version__ = "0.1.0"
