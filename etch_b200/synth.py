"""Seeded synthetic inputs: clothed-human scans and reference-layout checkpoints (no network, no datasets here).

  sample_scan(n, seed)        -> [n,3] f32 human-scale "clothed scan": points on a capsule humanoid (1.7 m), pushed
                                 outward along the normal by U(0,3cm), jittered by N(0,1mm), rotated about the vertical
                                 and bbox-centred the way src/inference_demo.py:25-28 centres a scan.
  make_state_dict(seed)       -> state_dict with the reference key layout: reference initialisation
                                 (BasicSO3Conv xavier_normal + bias 1e-3, then the global xavier_uniform_ of
                                 models_pointcloud.py:72-77) + non-trivial BatchNorm statistics + sharpened attention, so
                                 every code path (BN folding, softmax, polar decomposition) is exercised non-degenerately.
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


def make_state_dict(seed=1, n_markers=86, attention_gain=2.0):
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
    for l in (0, 1):
        for nm in ("key_transform", "query_transform"):
            sd["direction_encoder.self_attention_layers.%d.%s.weight" % (l, nm)].mul_(attention_gain)
    return {k: v.clone() for k, v in sd.items()}
