"""Drop-in for the reference extension ``epn_gathering`` (external/vgtk/vgtk/cuda/gathering_cuda.cpp:62-65)."""
import torch

import os
import sys

try:
    from etch_b200 import _lib as L
except ImportError:   # only this directory is on sys.path (B2 drop-in use): add the repository root
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from etch_b200 import _lib as L


def gather_points_forward(points, idx):
    """points [B,C,n] f32, idx [B,m] i32 -> [B,C,m] (gathering_cuda.cpp:29-43)."""
    if not points.is_cuda or not idx.is_cuda:
        raise RuntimeError("points/idx must be CUDA tensors")
    if not points.is_contiguous() or not idx.is_contiguous():
        raise RuntimeError("points/idx must be contiguous")
    B, C, n = points.shape
    m = idx.shape[1]
    out = torch.empty(B, C, m, dtype=torch.float32, device=points.device)
    L.call("gather_bcn", L.ptr(points), L.ptr(idx.int()), B, C, n, m, L.ptr(out))
    return out


def gather_points_backward(*args, **kwargs):
    raise RuntimeError("epn_gathering.gather_points_backward is training-only (not provided by etch_b200)")
