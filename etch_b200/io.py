"""Output writers of the evaluation driver (src/eval.py:125-179, src/utils/GT_utils.py:9-55).

  save_points_with_vector(hit_points, vectors, file_path)   same name / arguments / bytes as utils.GT_utils.save_points_with_vector
                                                            (the reference spends ~35 ms per file in a Python loop; this is one C call)
  results_to_host(fit)                                      ONE pinned device->host copy per batch of everything eval.py stores per scan
                                                            (the reference issues ~20 .detach().cpu().numpy() calls per scan)
The colour PLYs and OBJ exports of eval.py go through trimesh / matplotlib (un-vendored third-party code) and stay the caller's."""
import ctypes

import numpy as np
import torch

from . import _lib as L


def _fstring_style():
    """how `f"{numpy.float32}"` prints in this interpreter: 0 = numpy >= 2 (repr of the widened double), 1 = numpy 1.x (short form)"""
    return 1 if f"{np.float32(0.1)}" == "0.1" else 0


def save_points_with_vector(hit_points, vectors, file_path, style=None):
    hit = np.ascontiguousarray(hit_points, np.float32)
    vec = np.ascontiguousarray(vectors, np.float32)
    assert len(hit) == len(vec)                       # GT_utils.py:25
    fn = L.lib().etch_write_points_vector_ply
    fn.restype = ctypes.c_int
    rc = fn(str(file_path).encode(), hit.ctypes.data_as(ctypes.c_void_p), vec.ctypes.data_as(ctypes.c_void_p), int(len(hit)),
            _fstring_style() if style is None else int(style))
    if rc != 0:
        raise OSError(rc, "etch_write_points_vector_ply failed for %s" % file_path)


def format_np_float32(x, style):
    """text of every element: style 0 = repr(float(numpy.float32)), style 1 = str(numpy.float32) (used by the formatter tests)"""
    x = np.ascontiguousarray(x, np.float32).ravel()
    cap = 40 * x.size + 16
    buf = ctypes.create_string_buffer(cap)
    fn = L.lib().etch_format_np_float32
    fn.restype = ctypes.c_longlong
    n = fn(x.ctypes.data_as(ctypes.c_void_p), int(x.size), int(style), buf, ctypes.c_longlong(cap))
    if n < 0:
        raise RuntimeError("formatter buffer too small")
    return buf.raw[:n].decode().split("\n")[:-1]


_KEYS = ("labels", "tightness", "inner", "confidences", "markers", "valid", "vertices", "joints", "params")


def results_to_host(fit, stream=None):
    """dict of CUDA tensors (etch_b200.runtime.ScanFitter output) -> dict of numpy arrays through ONE packed pinned copy."""
    parts = [fit[k].reshape(fit[k].shape[0], -1).to(torch.float32) for k in _KEYS]
    widths = [p.shape[1] for p in parts]
    packed = torch.cat(parts, 1)
    host = torch.empty(packed.shape, dtype=torch.float32, pin_memory=True)
    host.copy_(packed, non_blocking=True)
    torch.cuda.current_stream(packed.device).synchronize()
    out, o = {}, 0
    arr = host.numpy()
    for k, w in zip(_KEYS, widths):
        v = arr[:, o:o + w].reshape(tuple(fit[k].shape))
        out[k] = v.astype(np.int64) if k == "labels" else (v.astype(bool) if k == "valid" else v.copy())
        o += w
    return out
