"""Drop-in for the reference extension ``epn_grouping`` (external/vgtk/vgtk/cuda/grouping_cuda.cpp:176-181)."""
import torch

import os
import sys

try:
    from etch_b200 import _lib as L
except ImportError:   # only this directory is on sys.path (B2 drop-in use): add the repository root
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from etch_b200 import _lib as L


def _check(x, name):
    # CHECK_INPUT of the reference (grouping_cuda.cpp:66-68) -> RuntimeError
    if not x.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if not x.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)
    if x.dtype != torch.float32:
        raise RuntimeError("%s must be float32" % name)


def ball_query(new_xyz, xyz, radius, nsample):
    """new_xyz [B,3,m], xyz [B,3,n] -> idx [B,m,nsample] int32 (grouping_cuda.cpp:71-86)."""
    _check(new_xyz, "new_xyz")
    _check(xyz, "xyz")
    B, _, m = new_xyz.shape
    n = xyz.shape[2]
    idx = torch.empty(B, m, int(nsample), dtype=torch.int32, device=xyz.device)
    L.call("ball_query_bcn", L.ptr(new_xyz), L.ptr(xyz), B, m, n, L.f32(radius), int(nsample), L.ptr(idx))
    return idx


def furthest_point_sampling(source_xyz, m):
    """source_xyz [B,3,n] -> idx [B,m] int32 (grouping_cuda.cpp:160-174)."""
    _check(source_xyz, "source_xyz")
    B, _, n = source_xyz.shape
    idx = torch.zeros(B, int(m), dtype=torch.int32, device=source_xyz.device)
    L.call("fps_bcn", L.ptr(source_xyz), B, n, int(m), L.ptr(idx))
    return idx


def anchor_query(*args, **kwargs):
    raise RuntimeError("epn_grouping.anchor_query is not on the ETCH inference path (not provided by etch_b200)")


def initial_anchor_query(*args, **kwargs):
    raise RuntimeError("epn_grouping.initial_anchor_query is not on the ETCH inference path (not provided by etch_b200)")
