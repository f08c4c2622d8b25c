"""GPU parity of marker extraction, SMPL LBS and the two-stage LM fit against the oracle / reference goldens.

Tolerances: LBS vertices 2e-5 m; markers 1e-5 m; fitted mesh: vertex-to-vertex (V2V, mean over 6890 vertices) between the
CUDA fit and the oracle fit < 1 mm (BASELINE target), typically ~1e-2 mm."""
import json
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _args(body=None):
    ms = json.load(open(os.path.join(ROOT, "etch_b200", "data", "superset_smpl.json")))
    return types.SimpleNamespace(markerset=ms, smpl_model=body, device="cuda:0")


def _tmodel(body):
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in body.items()}


def test_lbs_matches_reference_golden(cuda):
    from etch_b200 import smpl_model
    from etch_b200.models import fit_SMPL as F
    g = np.load(os.path.join(GOLD, "golden_lbs.npz"))
    body = smpl_model.synthetic_body(0)
    T = F.body_tables(_args(body), "neutral", cuda)
    params = torch.from_numpy(np.concatenate([g["pose"], g["betas"], g["transl"]], 1)).to(cuda)
    v, j = F.smpl_forward(T, params)
    assert np.abs(v.cpu().numpy() - g["verts"]).max() < 2e-5
    assert np.abs(j.cpu().numpy() - g["joints"]).max() < 2e-5


def test_markers_match_reference_golden(cuda):
    from etch_b200.models import fit_SMPL as F
    g = np.load(os.path.join(GOLD, "golden_markers.npz"))
    mk, valid = F.get_markers(_args(), torch.from_numpy(g["inner"]).to(cuda), torch.from_numpy(g["labels"]).to(cuda),
                              torch.from_numpy(g["conf"]).to(cuda))
    assert (valid.cpu().numpy() == g["valid"]).all()
    assert np.abs(mk.cpu().numpy() - g["markers"]).max() < 1e-5


def test_markers_ragged_and_empty(cuda):
    from etch_b200.models import fit_SMPL as F
    from oracle import lm as olm
    g = torch.Generator().manual_seed(3)
    B, N = 3, 5000
    inner = torch.randn(B, N, 3, generator=g)
    labels = torch.randint(0, 86, (B, N), generator=g)
    labels[2] = 5  # one scan with a single label everywhere -> 85 invalid markers
    conf = torch.rand(B, N, 1, generator=g)
    mk, valid = F.get_markers(_args(), inner.to(cuda), labels.to(cuda), conf.to(cuda))
    rm, rv = olm.get_markers(inner, labels, conf, 86)
    assert (valid.cpu() == rv).all() and int(valid[2].sum()) == 1
    assert (mk.cpu() - rm).abs().max() < 1e-5


def _synthetic_targets(body, B, seed, noise=0.003, drop=6):
    from oracle import smpl as osmpl
    g = torch.Generator().manual_seed(seed)
    model = _tmodel(body)
    betas = 0.8 * torch.randn(B, 10, generator=g)
    pose = 0.25 * torch.randn(B, 69, generator=g)
    orient = 0.3 * torch.randn(B, 3, generator=g)
    transl = 0.2 * torch.randn(B, 3, generator=g)
    v, _ = osmpl.smpl_forward(model, orient, pose, betas, transl)
    vids = torch.tensor(list(_args().markerset.values()))
    target = v[:, vids] + noise * torch.randn(B, len(vids), 3, generator=g)
    mask = torch.ones(B, len(vids), dtype=torch.bool)
    for b in range(B):
        mask[b, torch.randperm(len(vids), generator=g)[:drop]] = False
    return model, vids, target, mask, v


def test_lm_fit_matches_oracle(cuda):
    from etch_b200 import smpl_model
    from etch_b200.models import fit_SMPL as F
    from oracle import lm as olm
    body = smpl_model.synthetic_body(0)
    model, vids, target, mask, v_true = _synthetic_targets(body, 2, 5)
    ref = olm.fit(model, vids, target, mask)
    T = F.body_tables(_args(body), "neutral", cuda)
    out = F.lm_fit(T, target.to(cuda), mask.to(cuda))
    torch.cuda.synchronize()
    v2v = (out["vertices"].cpu() - ref["vertices"]).norm(dim=-1).mean(dim=1) * 1000.0
    print("V2V cuda-vs-oracle (mm):", v2v.tolist(), "iters:", out["iters"].tolist(), "errs:", out["errs"].tolist())
    assert v2v.max().item() < 1.0
    for k in ("pose", "betas", "orient", "transl"):
        assert (out[k].cpu() - ref[k]).abs().max() < 2e-2, k
    assert (out["joints"].cpu() - ref["joints"]).abs().max() < 2e-3
    # and the fit is a good fit: close to the generating mesh despite 3 mm marker noise
    assert ((out["vertices"].cpu() - v_true).norm(dim=-1).mean(dim=1) * 1000.0).max() < 15.0


def test_lm_noise_free_recovers_parameters(cuda):
    """size-independent property: from noise-free markers the fit lands on the generating mesh (a few mm: the
    reference's fixed-damping LM with step sizes 0.5/0.2 is not run to convergence in 30+50 iterations)."""
    from etch_b200 import smpl_model
    from etch_b200.models import fit_SMPL as F
    body = smpl_model.synthetic_body(0)
    model, vids, target, mask, v_true = _synthetic_targets(body, 8, 9, noise=0.0, drop=0)
    T = F.body_tables(_args(body), "neutral", cuda)
    out = F.lm_fit(T, target.to(cuda), mask.to(cuda))
    v2v = (out["vertices"].cpu() - v_true).norm(dim=-1).mean(dim=1) * 1000.0
    assert v2v.max().item() < 5.0, v2v.tolist()


def test_fit_smpl_contract(cuda):
    from etch_b200 import smpl_model
    from etch_b200.models.fit_SMPL import fit_smpl
    body = smpl_model.synthetic_body(0)
    args = _args(body)
    g = torch.Generator().manual_seed(1)
    B, N = 2, 3000
    inner = 0.3 * torch.randn(B, N, 3, generator=g)
    labels = torch.randint(0, 86, (B, N), generator=g)
    conf = torch.rand(B, N, 1, generator=g)
    meshes, markers, valid, info = fit_smpl(args, inner.to(cuda), labels.to(cuda), conf.to(cuda), "neutral")
    assert len(meshes) == B and np.asarray(meshes[0].vertices).shape == (6890, 3)
    assert markers.shape == (B, 86, 3) and valid.dtype == torch.bool
    assert [a.shape for a in info] == [(B, 23, 3), (B, 10), (B, 3), (B, 3), (B, 45, 3)]
    with pytest.raises(ValueError):
        fit_smpl(args, inner.to(cuda), labels.to(cuda), conf.to(cuda), "robot")


def test_analytic_jacobian_matches_autodiff(cuda):
    """the closed-form marker-only Jacobian inside etch_lm_fit against torch.func.jacrev through the
    oracle's full-mesh LBS (what theseus' AutoDiffCostFunction differentiates, fit_SMPL.py:111-152,179,230), at random poses / shapes,
    at the zero pose the solve starts from (Rodrigues singularity) and with some markers masked out."""
    from etch_b200 import _lib as L, smpl_model
    from etch_b200.models import fit_SMPL as F
    from oracle import lm as olm
    body = smpl_model.synthetic_body(0)
    model, vids, target, mask, _ = _synthetic_targets(body, 3, 5)
    mask = mask.clone()
    mask[1, ::7] = False
    g = torch.Generator().manual_seed(3)
    B, M = 3, target.shape[1]
    params = torch.zeros(B, 85)
    params[1:, 0:72] = 0.4 * torch.randn(2, 72, generator=g)        # orient | pose   (row 0 stays at the zero pose)
    params[1:, 72:82] = torch.randn(2, 10, generator=g)             # betas
    params[:, 82:85] = 0.2 * torch.randn(B, 3, generator=g)         # transl
    T = F.body_tables(_args(body), "neutral", cuda)
    res = torch.empty(B, M * 3, device=cuda)
    jac = torch.empty(B, M * 3, 85, device=cuda)
    dp, dt, dm = params.to(cuda), target.to(cuda).contiguous(), mask.to(torch.uint8).to(cuda).contiguous()
    L.call("lm_jacobian_dump", L.ptr(dp), L.ptr(dt), L.ptr(dm), L.ptr(T.Tm), L.ptr(T.Sm), L.ptr(T.Pm), L.ptr(T.Wm), L.ptr(T.Jt), L.ptr(T.Js),
           L.ptr(T.parents), L.ptr(T.ancmask), B, M, L.ptr(res), L.ptr(jac))
    torch.cuda.synchronize()
    vid_t = torch.as_tensor(vids, dtype=torch.long)
    for b in range(B):
        # oracle variable order: pose 69 | betas 10 | orient 3 | transl 3   ->   kernel column order: orient 3 | pose 69 | betas 10 | transl 3
        x = torch.cat([params[b, 3:72], params[b, 72:82], params[b, 0:3], params[b, 82:85]])
        f = lambda xi: olm._residual(xi, 10, model, vid_t, target[b], mask[b].float())      # noqa: E731
        r_ref = f(x)
        J_ref = torch.func.jacrev(f)(x)
        J_ref = torch.cat([J_ref[:, 79:82], J_ref[:, 0:69], J_ref[:, 69:79], J_ref[:, 82:85]], 1)
        r_err = (res[b].cpu() - r_ref).abs().max().item()
        J_err = (jac[b].cpu() - J_ref).abs().max().item()
        print("scan %d: residual max err %.2e, Jacobian max err %.2e (|J| max %.2f)" % (b, r_err, J_err, J_ref.abs().max().item()))
        assert r_err < 2e-6 and J_err < 5e-6 * max(1.0, J_ref.abs().max().item())      # measured on B200: 4.8e-7
    assert (jac[1].cpu().view(M, 3, 85)[::7] == 0).all()            # masked markers contribute nothing
