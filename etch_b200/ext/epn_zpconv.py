"""Import-time stand-in for ``epn_zpconv`` (external/vgtk/vgtk/cuda/zpconv_cuda.cpp:113-118).

``import vgtk`` imports this module (vgtk/zpconv/functional.py:17-19) but the ETCH inference path never calls it."""


def _unsupported(*args, **kwargs):
    raise RuntimeError("epn_zpconv kernels are legacy ZPConv ops, not on the ETCH inference path")


inter_zpconv_forward = inter_zpconv_backward = intra_zpconv_forward = intra_zpconv_backward = _unsupported
