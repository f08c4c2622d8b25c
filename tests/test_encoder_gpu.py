"""GPU parity of the fused SO(3) encoder kernels against the torch-CPU oracle (same seeded weights and scans).

Indices (FPS order, ball-query lists) must be identical; activations are compared layer by layer with the tolerance
stated below (fp32 everywhere; the only differences are summation order and the expanded form of the kernel weight)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL, ATOL = 2e-4, 2e-4  # activations are O(1) after InstanceNorm; pre-norm z values are compared relative to their scale


def _close(got, ref, name):
    got, ref = got.float().cpu(), ref.float().cpu()
    scale = ref.abs().max().item() + 1e-12
    err = (got - ref).abs().max().item()
    assert err <= ATOL * scale + RTOL * scale, "%s: max abs err %.3e vs scale %.3e" % (name, err, scale)


@pytest.mark.parametrize("B,N", [(2, 1024), (1, 1531)])
def test_encoder_layers_match_oracle(cuda, B, N):
    from etch_b200 import synth
    from etch_b200.models import encoder, spec
    from oracle import net as onet

    sd = synth.make_state_dict(1)
    pts = torch.from_numpy(synth.sample_scans(B, N, 7))
    ref_trace = []
    with torch.no_grad():
        rxyz, rfeats, _ = onet.encoder(pts, sd, spec.so3_tables(), trace=ref_trace)
    plan = encoder.EncoderPlan(sd, cuda)
    trace = []
    xyz, feats = encoder.run_encoder(plan, pts.permute(0, 2, 1).contiguous().to(cuda), trace)
    torch.cuda.synchronize()
    for li, (g, r) in enumerate(zip(trace, ref_trace)):
        np.testing.assert_array_equal(g["sample_idx"].cpu().numpy(), r["sample_idx"].numpy(), err_msg="sample_idx L%d" % li)
        np.testing.assert_array_equal(g["ball_idx"].cpu().numpy(), r["ball_idx"].numpy(), err_msg="ball_idx L%d" % li)
        _close(g["xyz"], r["xyz"], "xyz L%d" % li)
        _close(encoder.to_reference_layout(g["inter_z"]), r["inter_z"], "inter_z L%d" % li)
        _close(encoder.to_reference_layout(g["intra_z"]), r["intra_z"], "intra_z L%d" % li)
        if g["skip_z"] is not None:
            _close(encoder.to_reference_layout(g["skip_z"]), r["skip_z"], "skip_z L%d" % li)
        _close(encoder.to_reference_layout(g["out"]), r["out"], "out L%d" % li)
    _close(encoder.to_reference_layout(feats), rfeats, "encoder output")
