"""Parity WHERE THE METRIC IS QUOTED: the CUDA path against the CPU oracle at BASELINE sizes (5000 and 10000 points per
scan, clouds sampled from the reference's in-tree 4D-Dress scan) and end to end (points -> fitted SMPL vertices).

Stated tolerances (fp32 arithmetic on both sides; differences = summation order, folded BatchNorm, fused affine maps,
3xTF32 products with fp32 accumulation):
  FPS / ball-query / kNN indices ........ identical
  encoder activations per layer ......... max |err| <= 4e-4 x tensor scale
  invariant feature, logits, confidence,
  magnitude ............................. max |err| <= 1e-3 x tensor scale
  marker-label argmax ................... identical wherever the oracle's top-2 logit gap > 1e-3 (flips counted, < 1 %)
  tightness vectors (ALL points) ........ histogram printed; median <= 2e-5 m, 99 % <= 1e-4 m (1e-4 m) --
                                          the tail beyond that is the chordal mean of the seeded RANDOM checkpoint where
                                          sum_a w_a R_a is nearly singular (its polar factor is discontinuous there); every
                                          point whose sigma_2/sigma_1 > 1e-2 must be within 1e-3 m
  markers ............................... <= 1e-4 m for every marker whose top-3 points are the same on both sides
  fitted SMPL ........................... V2V <= 1 mm per scan (BASELINE.json target), joints <= 1 mm
"""
import json
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _markerset():
    return json.load(open(os.path.join(ROOT, "etch_b200", "data", "superset_smpl.json")))


def _model(cuda):
    from etch_b200 import synth
    from etch_b200.models.models_pointcloud import GT_network_equiv
    sd = synth.make_state_dict(1)
    opt = types.SimpleNamespace(output_folder=None, EPN_input_radius=0.4, EPN_layer_num=2, markerset=_markerset())
    net = GT_network_equiv(opt)
    net.load_state_dict(sd)
    return net.to(cuda).eval(), sd


def _rel(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    return ((got - ref).abs().max() / (ref.abs().max() + 1e-12)).item()


def _conditioning(anc_w):
    from etch_b200.models import spec
    anchors = torch.from_numpy(spec.so3_tables()["anchors"])
    Ce = torch.einsum("...a,aij->...ij", anc_w.cpu(), anchors)
    sv = torch.linalg.svdvals(Ce)
    return (sv[..., 1] / sv[..., 0].clamp_min(1e-20)).numpy()


def _hist(name, err, edges=(1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1)):
    err = np.asarray(err).ravel()
    parts = ["%s: n=%d median=%.2e p99=%.2e max=%.2e |" % (name, err.size, np.median(err), np.quantile(err, 0.99), err.max())]
    lo = 0.0
    for e in edges:
        parts.append("<%.0e:%d" % (e, int(((err >= lo) & (err < e)).sum())))
        lo = e
    parts.append(">=%.0e:%d" % (lo, int((err >= lo).sum())))
    print(" ".join(parts))


@pytest.mark.parametrize("N,kind", [(5000, "real"), (5000, "capsule"), (10000, "real")])
def test_forward_parity_at_baseline_sizes(cuda, N, kind):
    from etch_b200 import synth
    from etch_b200.models import encoder, spec
    from oracle import net as onet
    net, sd = _model(cuda)
    pts = torch.from_numpy(synth.sample_real_scans(1, N, 7) if kind == "real" else synth.sample_scans(1, N, 7))
    with torch.no_grad():
        ref, rt = onet.forward(pts, sd, spec.so3_tables(), return_trace=True)
    tr = {}
    out, _ = net(pts.to(cuda), ["confidence", "direction", "magnitude"], "standard_vector", _trace=tr)
    torch.cuda.synchronize()
    # encoder, layer by layer: indices identical, activations within 4e-4 of the tensor's scale
    for li, (g, r) in enumerate(zip(tr["enc"], rt["enc"])):
        np.testing.assert_array_equal(g["sample_idx"].cpu().numpy(), r["sample_idx"].numpy(), err_msg="sample_idx L%d" % li)
        np.testing.assert_array_equal(g["ball_idx"].cpu().numpy(), r["ball_idx"].numpy(), err_msg="ball_idx L%d" % li)
        e = _rel(encoder.to_reference_layout(g["out"]), r["out"])
        print("encoder out L%d rel err %.2e" % (li, e))
        assert e < 4e-4, (li, e)
    # 13 kNN graphs + 4 FPS levels of the PointTransformer hierarchy
    geo = tr["geo"]
    assert len(geo.p) == len(rt["pt_points"]) == 5
    for lvl, (gp, rp) in enumerate(zip(geo.p, rt["pt_points"])):      # identical FPS picks <=> identical level coordinates
        np.testing.assert_array_equal(gp.cpu().numpy(), rp.numpy(), err_msg="pointops FPS, level %d" % lvl)
    assert _rel(tr["inv"], rt["inv"]) < 1e-3
    assert _rel(tr["anc_w"], rt["anc_w"]) < 2e-3
    for k in ("part_labels", "confidences", "magnitude"):
        e = _rel(out[k], ref[k])
        print("%s rel err %.2e" % (k, e))
        assert e < 1e-3, (k, e)
    top2 = ref["part_labels"].topk(2, dim=-1).values
    gap = (top2[..., 0] - top2[..., 1]).numpy()
    flips = (out["part_labels"].argmax(-1).cpu() != ref["part_labels"].argmax(-1)).numpy()
    print("argmax flips: %d of %d (all with oracle top-2 gap < 1e-3: %s)" % (flips.sum(), flips.size, not (flips & (gap > 1e-3)).any()))
    assert not (flips & (gap > 1e-3)).any() and flips.mean() < 0.01
    # tightness vectors, ALL points
    vec = (out["direction"] * out["magnitude"] / 10.0).cpu().numpy()
    vref = (ref["direction"] * ref["magnitude"] / 10.0).numpy()
    verr = np.linalg.norm(vec - vref, axis=-1).ravel()
    cond = _conditioning(rt["anc_w"]).ravel()
    _hist("tightness vector |err| (m), all points", verr)
    _hist("tightness vector |err| (m), sigma2/sigma1 > 1e-2 (%.1f%% of points)" % (100 * (cond > 1e-2).mean()), verr[cond > 1e-2])
    assert np.median(verr) <= 2e-5 and np.quantile(verr, 0.99) <= 1e-4, (np.median(verr), np.quantile(verr, 0.99))
    assert verr[cond > 1e-2].max() <= 1e-3


def _oracle_pipeline(pts, sd, body, ms):
    from etch_b200.models import spec
    from oracle import lm as olm
    from oracle import net as onet
    with torch.no_grad():
        out = onet.forward(pts, sd, spec.so3_tables())
        labels, vec, inner = onet.postprocess(pts, out)
        mk, valid = olm.get_markers(inner, labels, out["confidences"], len(ms))
    body_t = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in body.items()}
    fit = olm.fit(body_t, list(ms.values()), mk, valid)
    return out, labels, vec, inner, mk, valid, fit


@pytest.mark.parametrize("B,N", [(2, 5000), (1, 10000)])
def test_points_to_vertices_matches_oracle(cuda, B, N):
    """The whole hot path on 2 scans x 5000 points (one real-scan cloud, one synthetic-clothed body cloud; BASELINE configs[1]) and on
    one 10000-point scan (configs[2], "V2V/MPJPE checked vs reference"):
    GPU net -> GPU markers -> GPU LM fit  vs  oracle net -> oracle markers -> oracle LM fit, full 30+50 iterations."""
    from etch_b200 import smpl_model, synth
    from etch_b200.models import fit_SMPL
    ms = _markerset()
    net, sd = _model(cuda)
    body = smpl_model.synthetic_body(0)
    pts = torch.from_numpy(synth.sample_real_scans(B, N, 3))
    o_out, o_labels, o_vec, o_inner, o_mk, o_valid, o_fit = _oracle_pipeline(pts, sd, body, ms)
    d = pts.to(cuda)
    out, _ = net(d, ["confidence", "direction", "magnitude"], "standard_vector")
    labels, vec, inner = net.postprocess(d, out)
    args = types.SimpleNamespace(markerset=ms, smpl_model=body, device="cuda:0")
    meshes, markers, valid, info = fit_SMPL.fit_smpl(args, inner, labels, out["confidences"], "neutral")
    torch.cuda.synchronize()
    top2 = o_out["part_labels"].topk(2, dim=-1).values
    gap = (top2[..., 0] - top2[..., 1]).numpy()
    flips = (labels.cpu() != o_labels).numpy()
    print("label flips %d / %d" % (flips.sum(), flips.size))
    assert not (flips & (gap > 1e-3)).any()
    np.testing.assert_array_equal(valid.cpu().numpy(), o_valid.numpy())
    merr = np.linalg.norm(markers.cpu().numpy() - o_mk.numpy(), axis=-1)[o_valid.numpy()]
    _hist("marker |err| (m)", merr)
    for b in range(B):
        v = np.asarray(meshes[b].vertices)
        v2v = 1000.0 * np.linalg.norm(v - o_fit["vertices"][b].numpy(), axis=-1).mean()
        jd = 1000.0 * np.linalg.norm(info[4][b] - o_fit["joints"][b].numpy(), axis=-1)
        jerr, mpjpe = jd.max(), jd[:22].mean()       # MPJPE over the first 22 joints (scripts/experiment_scripts/compute_mpjpe_error.py:23-24)
        print("scan %d: V2V vs oracle %.4f mm, MPJPE(22) %.4f mm, max joint err %.4f mm" % (b, v2v, mpjpe, jerr))
        if np.isfinite(o_fit["vertices"][b].numpy()).all():
            assert v2v <= 1.0 and jerr <= 1.0 and mpjpe <= 1.0, (b, v2v, mpjpe, jerr)
        else:   # a NaN marker (conf**20 underflow) makes the reference's fit NaN as well
            assert not np.isfinite(v).all()
    assert np.quantile(merr, 0.9) <= 1e-4 and merr.max() <= 1e-3, (np.quantile(merr, 0.9), merr.max())


def test_bench_configuration_in_flight_matches_eager(cuda):
    """The benchmarked configuration itself: B = 8 x 5000, five graph copies in flight on five streams, pinned-host input.
    Every ticket's vertices/labels must equal an eager (no graph, one stream) run of the same batch."""
    from etch_b200 import smpl_model, synth
    from etch_b200.runtime import ScanFitter
    net, _ = _model(cuda)
    args = types.SimpleNamespace(markerset=_markerset(), smpl_model=smpl_model.synthetic_body(0), device="cuda:0")
    eager = ScanFitter(net, args, use_graph=False)
    piped = ScanFitter(net, args, use_graph=True, in_flight=5)
    batches = [torch.from_numpy(synth.sample_real_scans(8, 5000, 40 + i)).pin_memory() for i in range(7)]
    tickets = [piped.submit(b, device=cuda) for b in batches[:5]]
    got = []
    for i, t in enumerate(tickets):
        r = t.result()
        got.append({k: r[k].clone() for k in ("vertices", "labels", "markers")})
    for b in batches[5:]:      # slots 0 and 1 come round again after their results were cloned on the caller's stream
        r = piped.submit(b, device=cuda).result()
        got.append({k: r[k].clone() for k in ("vertices", "labels", "markers")})
    torch.cuda.synchronize()
    for i, b in enumerate(batches):
        ref = eager(b.to(cuda))
        torch.cuda.synchronize()
        assert (got[i]["labels"] == ref["labels"]).all(), i
        fin = torch.isfinite(ref["vertices"]).all(-1).all(-1)
        v2v = ((got[i]["vertices"] - ref["vertices"]).norm(dim=-1).mean(-1))[fin].max().item() * 1000.0
        assert v2v < 0.05, (i, v2v)     # mm; only the double-atomic InstanceNorm statistics are order dependent
        assert (torch.isfinite(got[i]["vertices"]).all(-1).all(-1) == fin).all()
