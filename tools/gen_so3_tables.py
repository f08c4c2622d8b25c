"""Regenerates etch_b200/data/so3_tables.npz from the reference's own table builders.

Runs ONLY in the build container (needs /root/reference).  It executes the reference's
``vgtk.functional.rotation.icosahedron_so3_trimesh`` (external/vgtk/vgtk/functional/rotation.py:237-345, via
``vgtk.so3conv.functional.get_anchors / get_intra_idx``, functional.py:384-405) and reads the reference *data*
file ``vgtk/data/anchors/kpsphere24.ply`` (functional.py:146-157), with the tooling shims of tools/ref_shim.py.
Outputs (constants, also present as persistent buffers in any reference checkpoint):
    anchors   float32 [60,3,3]   icosahedral rotation group, anchors[29] == I
    intra_idx int64   [60,12]    SO(3) neighbourhood table of IntraSO3Conv
    kpsphere24 float32 [24,3]    raw (un-normalised) spherical kernel points, kp[0] == 0
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_shim  # noqa: E402

ref_shim.install()
import vgtk  # noqa: E402,F401
import vgtk.so3conv.functional as L  # noqa: E402
from vgtk import pc  # noqa: E402

anchors = np.ascontiguousarray(L.get_anchors(60)).astype(np.float32)
intra_idx = np.ascontiguousarray(L.get_intra_idx()).astype(np.int64)
kp = pc.load_ply(os.path.join(vgtk.__path__[0], "data", "anchors", "kpsphere24.ply")).astype("float32")
assert anchors.shape == (60, 3, 3) and intra_idx.shape == (60, 12) and kp.shape == (24, 3)
assert np.abs(anchors[29] - np.eye(3)).max() == 0
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "etch_b200", "data", "so3_tables.npz")
np.savez(out, anchors=anchors, intra_idx=intra_idx, kpsphere24=kp)
print("wrote", out)
