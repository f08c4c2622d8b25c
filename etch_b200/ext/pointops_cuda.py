"""Drop-in for the reference extension ``pointops_cuda`` (external/pointops/src/pointops_api.cpp:12-23).

Caller allocates the outputs, exactly as src/models/pointops.py:21-23,40-42 does."""
import os
import sys

try:
    from etch_b200 import _lib as L
except ImportError:   # only this directory is on sys.path (B2 drop-in use): add the repository root
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from etch_b200 import _lib as L


def knnquery_cuda(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2):
    L.call("knn_packed", int(m), int(nsample), L.ptr(xyz), L.ptr(new_xyz), L.ptr(offset), L.ptr(new_offset),
           int(offset.shape[0]), L.ptr(idx), L.ptr(dist2))


def furthestsampling_cuda(b, n_max, xyz, offset, new_offset, tmp, idx):
    L.call("fps_packed", int(b), int(n_max), L.ptr(xyz), L.ptr(offset), L.ptr(new_offset), L.ptr(tmp), L.ptr(idx))


def _unsupported(*args, **kwargs):
    raise RuntimeError("this pointops_cuda entry point is unused by the ETCH inference path (not provided by etch_b200)")


grouping_forward_cuda = grouping_backward_cuda = _unsupported
interpolation_forward_cuda = interpolation_backward_cuda = _unsupported
subtraction_forward_cuda = subtraction_backward_cuda = _unsupported
aggregation_forward_cuda = aggregation_backward_cuda = _unsupported
