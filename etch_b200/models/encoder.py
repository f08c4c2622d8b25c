"""Host-side driver of the SO(3)-equivariant encoder (EquivBackbone, src/models/so3net.py:11-33).

Only orchestration lives here: weight re-layout at load time, buffer allocation, kernel launches through the C ABI.
Feature tensors are point-major ``[B, P, 60, C]`` (the reference keeps ``[B, C, P, 60]``); use ``to_reference_layout``
when comparing.
"""
import os

import torch

from .. import _lib as L
from . import spec, tc


# tensor-core (tcgen05) kernels for the dense-GEMM stages; ETCH_B200_NO_TC=1 selects the fp32 CUDA-core versions (A/B tests)
USE_TC = os.environ.get("ETCH_B200_NO_TC", "0") != "1"
# InterSO3Conv variant: "v3" (one point per tile, TMA-fed, TMEM-parked accumulators: the product path) or "v2" (the
# 2-point slab kernel etch_so3_inter_conv_tc, kept for A/B runs)
INTER_VARIANT = os.environ.get("ETCH_B200_INTER", "v3")


def to_reference_layout(feats_bpac):
    """[B,P,60,C] -> the reference's [B,C,P,60]."""
    return feats_bpac.permute(0, 3, 1, 2).contiguous()


def _inter_slabs(W, ci, co):
    """BasicSO3Conv weight [co, ci*24] -> [ci/8*2, 2, 24, co, 4]: per (8-channel chunk, kernel-point half) the [co x 96] slab in
    the column order K'' = kgl*48 + c*6 + i used by inter_conv_tc_kernel, TF32 hi/lo split, canonical tiles."""
    W3 = W.view(co, ci, 24)
    slabs = []
    for sl in range(ci // 8 * 2):
        c0, h = (sl // 2) * 8, sl % 2
        blk = W3[:, c0:c0 + 8, h * 12:(h + 1) * 12].reshape(co, 8, 2, 6).permute(0, 2, 1, 3).reshape(co, 96)  # [o][kgl][c][i]
        slabs.append(tc.tc_operand(blk.contiguous(), "cpu"))
    return torch.stack(slabs, 0).contiguous()


def _inter_slabs_v3(W, ci, co):
    """BasicSO3Conv weight [co, ci*24] -> [ci/32*16, 12, 2*co, 4] for etch_so3_inter_conv_v3: per (pass, channel slot cc,
    kernel-point half hh) the 48-column slab K'' = o*12 + i <-> W[:, (32*pass + 8*o + cc)*24 + 12*hh + i]; rows are
    [W_hi; W_lo] (TF32 split) so that one MMA with N = 2*co forms A_hi*W_hi and A_hi*W_lo; canonical K-major tiles."""
    W3 = W.view(co, ci, 24)
    slabs = []
    for ps in range(ci // 32):
        for cc in range(8):
            for hh in range(2):
                chans = [32 * ps + 8 * o + cc for o in range(4)]
                blk = W3[:, chans, 12 * hh:12 * hh + 12].reshape(co, 48).contiguous()
                hi, lo = tc.split_tf32(blk)
                slabs.append(tc.canonical(torch.cat([hi, lo], 0)))
    return torch.stack(slabs, 0).contiguous()


class EncoderPlan:
    """Device-resident, kernel-friendly copies of the encoder weights (built once per state_dict/device)."""

    def __init__(self, sd, device, input_radius=0.4, n_layers=2):
        self.layers = []
        self.device = device
        f32 = dict(dtype=torch.float32, device="cpu")   # computed on the host, uploaded once at the end
        sd = {k: v.detach().cpu() for k, v in sd.items()} if any(v.is_cuda for v in sd.values()) else sd
        for lp in spec.epn_layers(input_radius, n_layers):
            pre = "encoder.backbone.%d.blocks.%d." % (lp["block"], lp["conv"])
            ci, co = lp["dim_in"], lp["dim_out"]
            anchors = sd[pre + "inter_conv.conv.anchors"].to(**f32)
            kernels = sd[pre + "inter_conv.conv.kernels"].to(**f32)
            # rotated kernel points R_a k  (functional.py:296): [60,24,3]
            kr = torch.matmul(anchors, kernels.t()).permute(0, 2, 1).contiguous()
            sigma = float(lp["sigma"])
            krs = torch.cat([kr * (2.0 / sigma), (kr * kr).sum(-1, keepdim=True) / sigma], -1).contiguous()
            W = sd[pre + "inter_conv.conv.basic_conv.W"].to(**f32)
            Wi = sd[pre + "intra_conv.conv.basic_conv.W"].to(**f32)
            d = dict(lp)
            d.update(
                kr=kr, krs=krs,
                Wt_inter=W.t().contiguous(),  # [(c,k)][o]
                b_inter=sd[pre + "inter_conv.conv.basic_conv.bias"].to(**f32).reshape(-1).contiguous(),
                Wt_intra=Wi.view(co, co, 12).permute(2, 1, 0).contiguous(),  # [j][c][o]
                b_intra=sd[pre + "intra_conv.conv.basic_conv.bias"].to(**f32).reshape(-1).contiguous(),
                intra_idx=sd[pre + "intra_conv.conv.intra_idx"].to(dtype=torch.int32).contiguous(),
                Wt_skip=sd[pre + "skip_conv.weight"].to(**f32).view(co, ci).t().contiguous(),  # [c][o]
                # tensor-core operands: per anchor-neighbour slot j the [c_out x c] slice, TF32-split, canonical tiles
                Wc_intra=torch.stack([tc.tc_operand(Wi.view(co, co, 12)[:, :, j].cpu(), "cpu") for j in range(12)], 0).contiguous(),
                Wc_inter=(_inter_slabs(W.cpu(), ci, co) if ci > 1 else None),
                Wc_inter3=(_inter_slabs_v3(W.cpu(), ci, co) if ci > 1 else None),
                Wc_skip=(tc.tc_operand(sd[pre + "skip_conv.weight"].view(co, ci).cpu(), "cpu")[None].contiguous() if ci > 1 else None),
                b_skip=sd[pre + "skip_conv.bias"].to(**f32).contiguous(),
            )
            self.layers.append({k: (v.to(device) if torch.is_tensor(v) else v) for k, v in d.items()})
        self.anchors = sd["encoder.backbone.0.blocks.0.inter_conv.conv.anchors"].to(**f32).contiguous().to(device)
        self.ident = torch.arange(60, dtype=torch.int32, device=device)


def run_encoder(plan, xyz_bcn, trace=None):
    """xyz_bcn [B,3,N] f32 cuda -> (xyz [B,3,P2], feats [B,P2,60,64]).  Mirrors BasicSO3ConvBlock/SeparableSO3ConvBlock
    (src/models/so3conv.py:125-183): inter conv -> IN+lrelu -> intra conv -> IN+lrelu, plus the skip branch."""
    dev = xyz_bcn.device
    B = xyz_bcn.shape[0]
    xyz = xyz_bcn.contiguous()
    feats = None
    for lp in plan.layers:
        q = xyz.shape[2]
        ci, co, nn_ = lp["dim_in"], lp["dim_out"], lp["n_neighbor"]
        P = -(-q // lp["stride"])
        if q == P or lp["lazy_sample"]:  # pc/sample.py:75-79
            sidx = torch.arange(P, dtype=torch.int32, device=dev).view(1, -1).expand(B, -1).contiguous()
        else:
            sidx = torch.empty(B, P, dtype=torch.int32, device=dev)
            L.call("fps_bcn", L.ptr(xyz), B, q, P, L.ptr(sidx))
        new_xyz = torch.empty(B, 3, P, dtype=torch.float32, device=dev)
        L.call("gather_bcn", L.ptr(xyz), L.ptr(sidx), B, 3, q, P, L.ptr(new_xyz))
        nbr = torch.empty(B, P, nn_, dtype=torch.int32, device=dev)
        L.call("ball_query_bcn", L.ptr(new_xyz), L.ptr(xyz), B, P, q, L.f32(lp["radius"]), nn_, L.ptr(nbr))
        stats = torch.zeros(3, B, co, 2, dtype=torch.float64, device=dev)
        z1 = torch.empty(B, P, 60, co, dtype=torch.float32, device=dev)
        if ci == 1:
            L.call("so3_inter_conv_c1", L.ptr(xyz), L.ptr(sidx), L.ptr(nbr), L.ptr(lp["kr"]), L.ptr(lp["Wt_inter"]),
                   L.ptr(lp["b_inter"]), B, q, P, nn_, co, L.f32(lp["sigma"]), L.ptr(z1), L.ptr(stats[0]))
        elif USE_TC and INTER_VARIANT == "v3":
            g4 = torch.empty(B, P, nn_, 4, dtype=torch.float32, device=dev)
            L.call("so3_inter_conv_v3", L.ptr(xyz), L.ptr(feats), L.ptr(sidx), L.ptr(nbr), L.ptr(lp["krs"]), L.ptr(lp["Wc_inter3"]),
                   L.ptr(lp["b_inter"]), B, q, P, nn_, ci, co, L.f32(lp["sigma"]), L.ptr(g4), L.ptr(z1), L.ptr(stats[0]))
        else:
            L.call("so3_inter_conv_tc" if USE_TC else "so3_inter_conv", L.ptr(xyz), L.ptr(feats), L.ptr(sidx), L.ptr(nbr), L.ptr(lp["krs"]),
                   L.ptr(lp["Wc_inter"] if USE_TC else lp["Wt_inter"]), L.ptr(lp["b_inter"]), B, q, P, nn_, ci, co, L.f32(lp["sigma"]),
                   L.ptr(z1), L.ptr(stats[0]))
        z2 = torch.empty_like(z1)
        if USE_TC:
            L.call("so3_intra_conv_tc", L.ptr(z1), L.ptr(stats[0]), L.ptr(lp["intra_idx"]), L.ptr(lp["Wc_intra"]),
                   L.ptr(lp["b_intra"]), B, P, co, co, L.ptr(z2), L.ptr(stats[1]))
        else:
            L.call("so3_intra_conv", L.ptr(z1), L.ptr(stats[0]), L.ptr(lp["intra_idx"]), L.ptr(lp["Wt_intra"]),
                   L.ptr(lp["b_intra"]), B, P, co, co, L.ptr(z2), L.ptr(stats[1]))
        out = torch.empty_like(z1)
        if ci == 1:
            # skip input is the constant occupancy feature: Conv1x1 gives a per-channel constant, whose InstanceNorm is 0
            z3 = None
            L.call("so3_combine", L.ptr(z2), L.ptr(stats[1]), L.ptr(None), L.ptr(None), B, P, co, L.ptr(out))
        else:
            z3 = torch.empty_like(z1)
            L.call("so3_skip_conv_tc" if USE_TC else "so3_skip_conv", L.ptr(feats), L.ptr(sidx), L.ptr(plan.ident),
                   L.ptr(lp["Wc_skip"] if USE_TC else lp["Wt_skip"]), L.ptr(lp["b_skip"]), B, q, P, ci, co, L.ptr(z3), L.ptr(stats[2]))
            L.call("so3_combine", L.ptr(z2), L.ptr(stats[1]), L.ptr(z3), L.ptr(stats[2]), B, P, co, L.ptr(out))
        if trace is not None:
            trace.append(dict(sample_idx=sidx, ball_idx=nbr, xyz=new_xyz, inter_z=z1, intra_z=z2, skip_z=z3, out=out))
        xyz, feats = new_xyz, out
    return xyz, feats
