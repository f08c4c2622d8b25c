"""A/B parity of the kernel generations on the GPU and the BASELINE.json scan sizes beyond 5k points.

The fp32 CUDA-core kernels of the first build stay in the library as same-operator references (ETCH_B200_NO_TC path):
  * InterSO3Conv "v3" (one point per tile, TMA-fed, TMEM-parked accumulators)  vs  etch_so3_inter_conv (fp32 FMA only)
  * PointTransformer attention on tcgen05                                       vs  etch_pt_attention (CTA per point)
Tolerance: 3xTF32 products + a different summation order => max abs err <= 1e-4 x the tensor's scale (observed 1e-5).
"""
import json
import os
import types

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rel(got, ref):
    return ((got.float() - ref.float()).abs().max() / (ref.float().abs().max() + 1e-12)).item()


@pytest.mark.parametrize("B,N,sm_budget", [(2, 1024, 8), (1, 1531, 148), (1, 10000, 148)])
def test_inter_conv_v3_matches_fp32_kernel(cuda, B, N, sm_budget):
    """sm_budget=8 leaves 4 CTAs per scan: every CTA walks over ~100 points (ring wrap-around, parked accumulators, the flush
    iteration); N=1531 gives odd point counts at both levels; N=10000 is BASELINE configs[2]'s scan size."""
    from etch_b200 import _lib as L, synth
    from etch_b200.models import encoder
    sd = synth.make_state_dict(1)
    plan = encoder.EncoderPlan(sd, cuda)
    pts = torch.from_numpy(synth.sample_scans(B, N, 7)).permute(0, 2, 1).contiguous().to(cuda)
    saved = (encoder.USE_TC, encoder.INTER_VARIANT)
    L.lib().etch_set_sm_budget(sm_budget)
    try:
        traces = {}
        for name, use_tc, variant in (("fp32", False, "v2"), ("v3", True, "v3")):
            encoder.USE_TC, encoder.INTER_VARIANT = use_tc, variant
            tr = []
            encoder.run_encoder(plan, pts, tr)
            torch.cuda.synchronize()
            traces[name] = tr
    finally:
        encoder.USE_TC, encoder.INTER_VARIANT = saved
        L.lib().etch_set_sm_budget(148)
    for li, (a, b) in enumerate(zip(traces["v3"], traces["fp32"])):
        assert (a["ball_idx"] == b["ball_idx"]).all()
        assert _rel(a["inter_z"], b["inter_z"]) < 1e-4, "inter_z layer %d" % li
        assert _rel(a["out"], b["out"]) < 1e-4, "block output layer %d" % li


def _net(cuda):
    from etch_b200 import synth
    from etch_b200.models.models_pointcloud import GT_network_equiv
    ms = json.load(open(os.path.join(ROOT, "etch_b200", "data", "superset_smpl.json")))
    net = GT_network_equiv(types.SimpleNamespace(output_folder=None, EPN_input_radius=0.4, EPN_layer_num=2, markerset=ms))
    net.load_state_dict(synth.make_state_dict(1))
    return net.to(cuda).eval(), ms


def test_pt_attention_tc_matches_cta_kernel(cuda):
    from etch_b200 import synth
    from etch_b200.models import heads
    net, _ = _net(cuda)
    pts = torch.from_numpy(synth.sample_scans(2, 1024, 11)).to(cuda)
    saved = heads.PT_ATTN_TC
    try:
        outs = {}
        for flag in (False, True):
            heads.PT_ATTN_TC = flag
            o, _ = net(pts, ["confidence", "magnitude"])
            torch.cuda.synchronize()
            outs[flag] = {k: v.clone() for k, v in o.items()}
    finally:
        heads.PT_ATTN_TC = saved
    for k in ("part_labels", "confidences", "magnitude"):
        assert _rel(outs[True][k], outs[False][k]) < 1e-4, k


@pytest.mark.parametrize("B,N", [(2, 10000), (2, 20000)])
def test_large_scans_through_the_whole_path(cuda, B, N):
    """BASELINE configs[2] / configs[3] scan sizes (10k and 20k points; the batch is cut to 2 scans to keep the test short):
    network forward + post-processing + markers + LM fit; size-independent properties + graph replay == eager."""
    from etch_b200 import smpl_model, synth
    from etch_b200.runtime import ScanFitter
    net, ms = _net(cuda)
    args = types.SimpleNamespace(markerset=ms, smpl_model=smpl_model.synthetic_body(0), device="cuda:0")
    pts = torch.from_numpy(synth.sample_scans(B, N, 21)).to(cuda)
    eager = ScanFitter(net, args, use_graph=False)
    ref = {k: v.clone() for k, v in eager(pts).items()}
    assert ref["labels"].shape == (B, N) and int(ref["labels"].min()) >= 0 and int(ref["labels"].max()) < len(ms)
    for k in ("vertices", "joints", "params", "tightness", "inner", "confidences"):
        assert torch.isfinite(ref[k]).all(), k
    assert ref["vertices"].shape == (B, 6890, 3)
    # tightness vectors point from the inner (body) point to the scan point: inner + vec == scan
    assert (ref["inner"] + ref["tightness"] - pts).abs().max() < 1e-5
    graphed = ScanFitter(net, args, use_graph=True)
    out = graphed(pts)
    torch.cuda.synchronize()
    assert (out["labels"] == ref["labels"]).float().mean() > 0.999
    assert (out["vertices"] - ref["vertices"]).norm(dim=-1).mean().item() * 1000.0 < 0.05   # mm


def test_mixed_size_stream_matches_per_scan_runs(cuda):
    """BASELINE configs[4] (mixed stream): scans of three sizes, batched by size and run through the in-flight graph slots,
    give each scan the result of running it alone."""
    from etch_b200 import smpl_model, stream, synth
    from etch_b200.runtime import ScanFitter
    net, ms = _net(cuda)
    args = types.SimpleNamespace(markerset=ms, smpl_model=smpl_model.synthetic_body(0), device="cuda:0")
    sizes = [1024, 2048, 1024, 1531, 2048, 1024, 1024]
    scans = {i: torch.from_numpy(synth.sample_scan(n, 100 + i)) for i, n in enumerate(sizes)}
    plan = stream.plan_stream(sizes, 1, 2)
    assert sum(len(ids) for _, ids in plan[0]) == len(sizes)
    fitter = ScanFitter(net, args, use_graph=True, in_flight=2)
    got = stream.run_stream(fitter, scans, plan[0], cuda, pad_to=2)
    eager = ScanFitter(net, args, use_graph=False)
    for i in (0, 3, 4, 6):
        ref = eager(scans[i][None].to(cuda))
        assert (got[i]["labels"] == ref["labels"][0]).float().mean() > 0.999
        assert (got[i]["vertices"] - ref["vertices"][0]).norm(dim=-1).mean().item() * 1000.0 < 0.05   # mm


@pytest.mark.parametrize("c,ns,n", [(64, 8, 1000), (128, 8, 517), (128, 16, 2048), (256, 16, 333), (512, 16, 19), (256, 8, 130)])
def test_pt_attention_tc_unit(cuda, c, ns, n):
    """etch_pt_attention_tc against etch_pt_attention on random operands: every channel width of the two PointTransformers,
    both neighbour counts, row counts that are not a multiple of the 128-row tile (ragged last tile, a single partial tile)."""
    from etch_b200 import _lib as L
    from etch_b200.models import tc
    g = torch.Generator().manual_seed(c * 1000 + ns + n)
    T = c // 8
    f = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).to(cuda).contiguous()  # noqa: E731
    p, qkv = f(n, 3, scale=0.3), f(n, 3 * c)
    idx = torch.randint(0, n, (n, ns), generator=g, dtype=torch.int32).to(cuda)
    P0, p0b, P3, p3b = f(3, 3), f(3, scale=0.1), f(c, 3), f(c, scale=0.1)
    s0, h0, so, ho = 1 + f(c, scale=0.1), f(c, scale=0.1), 1 + f(c, scale=0.1), f(c, scale=0.1)
    W1, b1, W2, b2 = f(T, c, scale=c ** -0.5), f(T, scale=0.1), f(T, T, scale=T ** -0.5), f(T, scale=0.1)
    ref = torch.empty(n, c, device=cuda)
    L.call("pt_attention", L.ptr(p), L.ptr(qkv), L.ptr(idx), L.ptr(P0), L.ptr(p0b), L.ptr(P3), L.ptr(p3b), L.ptr(s0), L.ptr(h0),
           L.ptr(W1), L.ptr(b1), L.ptr(W2), L.ptr(b2), L.ptr(so), L.ptr(ho), n, ns, c, L.ptr(ref))
    chan = torch.cat([P3, p3b[:, None], s0[:, None], h0[:, None], so[:, None], ho[:, None]], 1).contiguous()
    tp = max(T, 16)
    wa = torch.zeros(tp, c)
    wa[:T] = W1.cpu()
    W1c = torch.stack([tc.tc_operand(wa[:, k:k + 64].contiguous(), "cpu") for k in range(0, c, 64)], 0).contiguous().to(cuda)
    out = torch.empty(n, c, device=cuda)
    L.call("pt_attention_tc", L.ptr(p), L.ptr(qkv), L.ptr(idx), L.ptr(P0), L.ptr(p0b), L.ptr(chan), L.ptr(W1c), L.ptr(b1), L.ptr(W2),
           L.ptr(b2), n, ns, c, L.ptr(out))
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    assert _rel(out, ref) < 1e-4
