"""ctypes binding of libetch_b200.so (the C ABI declared in include/etch_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("ETCH_B200_LIB") or os.path.join(_HERE, "libetch_b200.so")   # override: A/B builds of the same ABI (tools/)
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise RuntimeError("etch_b200: %s not found -- run `python -m etch_b200.build` (no CPU fallback exists)" % _SO)
        _lib = ctypes.CDLL(_SO)
    return _lib


class DevPtr(ctypes.c_void_p):
    """A device pointer that remembers which CUDA device owns it (call() launches on that device and its current stream)."""
    device_index = None


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (or NULL for None)."""
    if t is None:
        return ctypes.c_void_p(0)
    if not t.is_cuda:
        raise RuntimeError("etch_b200: expected a CUDA tensor, got device %s" % t.device)
    if not t.is_contiguous():
        raise RuntimeError("etch_b200: tensor must be contiguous")
    p = DevPtr(t.data_ptr())
    p.device_index = t.device.index
    return p


def stream(device_index=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device_index).cuda_stream)


def _device_of(args):
    """The one CUDA device all pointer arguments live on (None when a call carries no tensors); mixed devices raise."""
    dev = None
    for a in args:
        if isinstance(a, DevPtr):
            if dev is None:
                dev = a.device_index
            elif a.device_index != dev:
                raise RuntimeError("etch_b200: arguments live on different CUDA devices (cuda:%d and cuda:%d)" % (dev, a.device_index))
    return dev


def f32(x):
    return ctypes.c_float(float(x))


# kernel launches issued per C-ABI call (for bench.py's "gpu_launches" claim); everything not listed launches once
_LAUNCHES_PER_CALL = {"lbs_forward": 2, "direction_head_tc": 4, "so3_inter_conv_v3": 2, "knn_grid": 5}
launch_count = 0
_profile = None  # when a dict: name -> [(start_event, end_event), ...]
# ETCH_B200_NVTX=1: every C-ABI call is wrapped in an NVTX range named after the entry point (shows up in ncu / nsys timelines; the
# reference has no tracing at all)
_NVTX = os.environ.get("ETCH_B200_NVTX", "0") == "1"


def start_profile():
    global _profile
    _profile = {}


def stop_profile():
    """-> {name: (calls, total_ms)} measured with CUDA events on the launching stream."""
    global _profile
    torch.cuda.synchronize()
    out = {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in _profile.items()}
    _profile = None
    return out


def call(name, *args):
    """Invoke ``int etch_<name>(..., cudaStream_t)`` on the device that owns the tensor arguments, on that device's current
    torch stream (like the reference's device-guarded native ops); raise on a non-zero status."""
    global launch_count
    fn = getattr(lib(), "etch_" + name)
    fn.restype = ctypes.c_int
    launch_count += _LAUNCHES_PER_CALL.get(name, 1)
    dev = _device_of(args)
    if dev is not None and dev != torch.cuda.current_device():
        with torch.cuda.device(dev):
            return _call_on_current_device(fn, name, args, dev)
    return _call_on_current_device(fn, name, args, dev)


def _call_on_current_device(fn, name, args, dev):
    if _NVTX:
        torch.cuda.nvtx.range_push("etch_" + name)
        try:
            rc = fn(*args, stream(dev))
        finally:
            torch.cuda.nvtx.range_pop()
    elif _profile is not None:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st = torch.cuda.current_stream(dev)
        a.record(st)
        rc = fn(*args, stream(dev))
        b.record(st)
        _profile.setdefault(name, []).append((a, b))
    else:
        rc = fn(*args, stream(dev))
    if rc != 0:
        raise RuntimeError("etch_b200: etch_%s failed with status %d (%s)" % (
            name, rc, "invalid argument" if rc == -1 else "unsupported" if rc == -2 else "cudaError"))
    return rc


def exported_symbols():
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", _SO], capture_output=True, text=True).stdout
    return sorted(l.split()[-1] for l in out.splitlines() if " T etch_" in l)
