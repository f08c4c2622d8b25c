"""ctypes front-end of oracle/_ref/libetch_ref_kernels.so -- TEST INFRASTRUCTURE (see oracle/__init__.py).

The library holds the REFERENCE's own CUDA kernels, compiled unmodified for sm_100a by oracle/build_ref.sh from
/root/reference (external/vgtk/vgtk/cuda/{grouping,gathering}_cuda_kernel.cu,
external/pointops/src/{knnquery,sampling}/*_cuda_kernel.cu).  GPU tests use it to pin (a) the product's index kernels and
(b) the C emulation in oracle/etch_oracle.c against what the reference really computes; tools/ref_kernel_bench.py uses the
timing entry points for the per-kernel "reference vs ours" table.  Never imported by etch_b200/.

Functions take and return torch CUDA tensors (device memory plumbing only).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libetch_ref_kernels.so")
_LIB = None


def available():
    return os.path.exists(SO)


def lib():
    global _LIB
    if _LIB is None:
        if not available():
            raise RuntimeError("oracle/_ref/libetch_ref_kernels.so missing: run oracle/build_ref.sh where /root/reference exists")
        _LIB = ctypes.CDLL(SO)
        for n in ("ref_time_ball_query", "ref_time_furthest_point_sampling", "ref_time_knnquery", "ref_time_furthestsampling"):
            getattr(_LIB, n).restype = ctypes.c_float
    return _LIB


def _p(t):
    assert t.is_cuda and t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError("reference kernel %s failed: cudaError %d" % (what, rc))


def ball_query(new_xyz, xyz, radius, nsample):
    """new_xyz [B,3,m], xyz [B,3,n] f32 -> idx [B,m,nsample] i32 (epn_grouping.ball_query)."""
    B, _, m = new_xyz.shape
    n = xyz.shape[2]
    idx = torch.empty(B, m, nsample, dtype=torch.int32, device=xyz.device)
    torch.cuda.synchronize()
    _chk(lib().ref_ball_query(_p(new_xyz), _p(xyz), B, m, n, ctypes.c_float(radius), int(nsample), _p(idx)), "ball_query")
    return idx


def furthest_point_sampling(xyz, m):
    """xyz [B,3,n] f32 -> idx [B,m] i32 (epn_grouping.furthest_point_sampling)."""
    B, _, n = xyz.shape
    temp = torch.empty(B, n, dtype=torch.float32, device=xyz.device)
    idx = torch.empty(B, m, dtype=torch.int32, device=xyz.device)
    torch.cuda.synchronize()
    _chk(lib().ref_furthest_point_sampling(_p(xyz), B, n, int(m), _p(temp), _p(idx)), "furthest_point_sampling")
    return idx


def gather_points_forward(points, idx):
    """points [B,C,n] f32, idx [B,m] i32 -> [B,C,m] (epn_gathering.gather_points_forward)."""
    B, C, n = points.shape
    m = idx.shape[1]
    out = torch.empty(B, C, m, dtype=torch.float32, device=points.device)
    torch.cuda.synchronize()
    _chk(lib().ref_gather_points_forward(_p(points), _p(idx), B, C, n, m, _p(out)), "gather_points_forward")
    return out


def knnquery(nsample, xyz, new_xyz, offset, new_offset):
    """pointops_cuda.knnquery_cuda: -> (idx [m,nsample] i32, dist2 [m,nsample] f32)."""
    m = new_xyz.shape[0]
    idx = torch.zeros(m, nsample, dtype=torch.int32, device=xyz.device)       # pointops.py:40-42
    d2 = torch.zeros(m, nsample, dtype=torch.float32, device=xyz.device)
    torch.cuda.synchronize()
    _chk(lib().ref_knnquery(m, int(nsample), _p(xyz), _p(new_xyz), _p(offset), _p(new_offset), _p(idx), _p(d2)), "knnquery")
    return idx, d2


def furthestsampling(xyz, offset, new_offset):
    """pointops_cuda.furthestsampling_cuda on packed rows: -> idx [new_offset[-1]] i32 (pointops.py:10-27)."""
    off = offset.cpu().tolist()
    n_max = max(b - a for a, b in zip([0] + off[:-1], off))
    n = xyz.shape[0]
    tmp = torch.empty(n, dtype=torch.float32, device=xyz.device)
    idx = torch.zeros(int(new_offset[-1].item()), dtype=torch.int32, device=xyz.device)
    torch.cuda.synchronize()
    _chk(lib().ref_furthestsampling(len(off), int(n_max), n, _p(xyz), _p(offset), _p(new_offset), _p(tmp), _p(idx)), "furthestsampling")
    return idx
