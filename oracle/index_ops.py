"""ctypes front-end of oracle/libetch_oracle.so (TEST INFRASTRUCTURE, see oracle/__init__.py).

Each function mirrors one native entry point of the reference:
  fps_bcn        <- epn_grouping.furthest_point_sampling   (external/vgtk/vgtk/cuda/grouping_cuda.cpp:160-174)
  ball_query_bcn <- epn_grouping.ball_query                (grouping_cuda.cpp:71-86)
  gather_bcn     <- epn_gathering.gather_points_forward    (gathering_cuda.cpp:29-43)
  knn_packed     <- pointops_cuda.knnquery_cuda            (external/pointops/src/knnquery/knnquery_cuda.cpp:8-17)
  fps_packed     <- pointops_cuda.furthestsampling_cuda    (external/pointops/src/sampling/sampling_cuda.cpp:8-16)
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libetch_oracle.so")
        if not os.path.exists(so):
            build()
        _LIB = ctypes.CDLL(so)
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


_F, _I = ctypes.c_float, ctypes.c_int


def opt_n_threads(n):
    return int(lib().etch_oracle_opt_n_threads(int(n)))


def fps_bcn(xyz, m):
    xyz = np.ascontiguousarray(xyz, np.float32)
    B, _, n = xyz.shape
    idx = np.zeros((B, m), np.int32)
    lib().etch_oracle_fps_bcn(_p(xyz, _F), B, n, int(m), _p(idx, _I))
    return idx


def ball_query_bcn(new_xyz, xyz, radius, nsample):
    new_xyz = np.ascontiguousarray(new_xyz, np.float32)
    xyz = np.ascontiguousarray(xyz, np.float32)
    B, _, m = new_xyz.shape
    n = xyz.shape[2]
    idx = np.zeros((B, m, nsample), np.int32)
    lib().etch_oracle_ball_query_bcn(_p(new_xyz, _F), _p(xyz, _F), B, m, n, _F(float(radius)), int(nsample), _p(idx, _I))
    return idx


def gather_bcn(points, idx):
    points = np.ascontiguousarray(points, np.float32)
    idx = np.ascontiguousarray(idx, np.int32)
    B, C, n = points.shape
    m = idx.shape[1]
    out = np.zeros((B, C, m), np.float32)
    lib().etch_oracle_gather_bcn(_p(points, _F), B, C, n, _p(idx, _I), m, _p(out, _F))
    return out


def knn_packed(nsample, xyz, new_xyz, offset, new_offset):
    xyz = np.ascontiguousarray(xyz, np.float32)
    new_xyz = np.ascontiguousarray(new_xyz, np.float32)
    offset = np.ascontiguousarray(offset, np.int32)
    new_offset = np.ascontiguousarray(new_offset, np.int32)
    m = new_xyz.shape[0]
    idx = np.zeros((m, nsample), np.int32)
    d2 = np.zeros((m, nsample), np.float32)
    lib().etch_oracle_knn_packed(m, int(nsample), _p(xyz, _F), _p(new_xyz, _F), _p(offset, _I), _p(new_offset, _I),
                                 _p(idx, _I), _p(d2, _F))
    return idx, d2


def fps_packed(xyz, offset, new_offset):
    xyz = np.ascontiguousarray(xyz, np.float32)
    offset = np.ascontiguousarray(offset, np.int32)
    new_offset = np.ascontiguousarray(new_offset, np.int32)
    B = offset.shape[0]
    seg = np.diff(np.concatenate([[0], offset]))
    n_max = int(seg.max())
    tmp = np.full((xyz.shape[0],), 1e10, np.float32)
    idx = np.zeros((int(new_offset[-1]),), np.int32)
    lib().etch_oracle_fps_packed(B, n_max, _p(xyz, _F), _p(offset, _I), _p(new_offset, _I), _p(tmp, _F), _p(idx, _I))
    return idx
