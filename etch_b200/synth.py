"""Seeded synthetic inputs: clothed-human scans and reference-layout checkpoints (no network, no datasets here).

  sample_scan(n, seed)        -> [n,3] f32 human-scale "clothed scan": points on a capsule humanoid (1.7 m), pushed
                                 outward along the normal by U(0,3cm), jittered by N(0,1mm), rotated about the vertical
                                 and bbox-centred the way src/inference_demo.py:25-28 centres a scan.
  sample_real_scans(B, n, seed) -> [B,n,3] f32 clouds drawn from the area-weighted sample pool of the reference's in-tree
                                 4D-Dress sample (etch_b200/data/scan_pool.npz, made by tools/gen_scan_pool.py): even scans
                                 are the real clothed scan, odd scans the SMPL body under it pushed out along the normal
                                 by U(0,3cm) ("synthetic clothed"); bbox-centred, random SO(3) rotation, N(0,1mm) jitter
                                 (SURVEY.md section 8d).
  make_state_dict(seed)       -> state_dict with the reference key layout: reference initialisation
                                 (BasicSO3Conv xavier_normal + bias 1e-3, then the global xavier_uniform_ of
                                 models_pointcloud.py:72-77) + non-trivial BatchNorm statistics + sharpened attention, so
                                 every code path (BN folding, softmax, polar decomposition) is exercised non-degenerately;
                                 seed 1 (the checkpoint of tests and bench) additionally carries BatchNorm running statistics
                                 calibrated on real-scan clouds (data/synth_calibration.npz, tools/gen_calibration.py) so that
                                 all 86 marker labels and sensible confidences occur, as with a trained checkpoint.
"""
import numpy as np
import torch

from . import smpl_model
from .models import spec


def sample_scan(n, seed=0, centre=True):
    rng = np.random.default_rng(1000 + seed)
    pts, nrm, _ = smpl_model.sample_capsule_surface(rng, n)
    pts = pts + nrm * rng.uniform(0.0, 0.03, size=(n, 1)) + rng.normal(0.0, 0.001, size=(n, 3))
    a = rng.uniform(0, 2 * np.pi)
    R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    pts = pts @ R.T
    if centre:
        pts = pts - (pts.max(0) + pts.min(0)) / 2
    return pts.astype(np.float32)


def sample_scans(B, n, seed=0):
    return np.stack([sample_scan(n, seed * 1000 + b) for b in range(B)], 0)


_POOL = None


def _scan_pool():
    global _POOL
    if _POOL is None:
        import os
        _POOL = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "scan_pool.npz")))
    return _POOL


def _random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def sample_real_scan(n, seed=0, kind=None, rotate=True):
    """n <= 32768 points of the in-tree 4D-Dress sample; kind "scan" | "body" (default: by seed parity)."""
    pool = _scan_pool()
    rng = np.random.default_rng(1000 + seed)
    kind = kind or ("scan" if seed % 2 == 0 else "body")
    sel = rng.permutation(pool["scan"].shape[0])[:n]
    if kind == "scan":
        pts = pool["scan"][sel].astype(np.float64)
    else:
        pts = pool["body"][sel].astype(np.float64) + pool["body_n"][sel].astype(np.float64) * rng.uniform(0.0, 0.03, size=(n, 1))
    pts = pts - (pts.max(0) + pts.min(0)) / 2          # src/inference_demo.py:25-28
    if rotate:
        pts = pts @ _random_rotation(rng).T
    pts = pts + rng.normal(0.0, 0.001, size=(n, 3))
    return pts.astype(np.float32)


def sample_real_scans(B, n, seed=0):
    return np.stack([sample_real_scan(n, seed * 1000 + b) for b in range(B)], 0)


def make_state_dict(seed=1, n_markers=86, attention_gain=2.0, calibrated=True, residual_gain=0.3):
    g = torch.Generator().manual_seed(seed)
    torch_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        tree = spec.ParamTree(spec.network_spec(n_markers))
        spec.xavier_reset_(tree)
    finally:
        torch.random.set_rng_state(torch_state)
    sd = tree.state_dict()
    for k, v in sd.items():
        if k.endswith("running_mean"):
            v.copy_(0.1 * torch.randn(v.shape, generator=g))
        elif k.endswith("running_var"):
            v.copy_(0.7 + 0.6 * torch.rand(v.shape, generator=g))
        elif v.dim() == 1 and v.is_floating_point() and (k.endswith(".bias") or k.endswith(".weight")):
            # BN affine terms and linear biases: small, non-zero
            if k.endswith(".weight"):
                v.copy_(1.0 + 0.2 * torch.randn(v.shape, generator=g))
            else:
                v.copy_(0.05 * torch.randn(v.shape, generator=g))
    # Residual-branch gains: a stack of unit-gain random residual blocks behind BatchNorm is chaotic (summation-order noise is
    # amplified ~30x per level and unseen scans collapse onto a handful of marker labels); trained networks are not.  Scale the
    # last BatchNorm of every PointTransformerBlock and of every TransitionUp's upsampled branch by 0.3.
    for k, v in sd.items():
        if k.endswith("bn3.weight") or (".0.linear2.1.weight" in k and ".dec" in k):
            v.mul_(residual_gain)
    for l in (0, 1):
        for nm in ("key_transform", "query_transform"):
            sd["direction_encoder.self_attention_layers.%d.%s.weight" % (l, nm)].mul_(attention_gain)
    if calibrated and seed == 1 and n_markers == 86:
        # BatchNorm running statistics measured on real-scan clouds (tools/gen_calibration.py) -- what training leaves in a real
        # checkpoint; without them the random PointTransformers collapse onto 3 marker labels and ~0.01 confidences
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "synth_calibration.npz")
        with np.load(path) as cal:
            for k in cal.files:
                sd[k].copy_(torch.from_numpy(cal[k]).reshape(sd[k].shape))
    return {k: v.clone() for k, v in sd.items()}
