"""Parameter/buffer layout of ETCH's network in the reference's state-dict key layout.

Checkpoints of the reference are plain ``state_dict``s, so the key names, shapes and dtypes below ARE the drop-in
contract (src/eval.py:309, src/inference_demo.py:15).  ``network_spec`` returns a nested dict whose leaves are
``("param"|"buffer", shape, dtype, init)``; ``ParamTree`` turns it into an ``nn.Module`` container tree.

Reference constructors mirrored (names only; the arithmetic lives in csrc/):
  src/models/models_pointcloud.py:18-70, src/models/so3net.py:11-21, src/models/so3conv.py:19-183,
  external/vgtk/vgtk/so3conv/modules.py:19-31,92-118,131-148, src/models/direction_backbones.py:129-209,
  src/models/pointtransformer_seg.py:8-22,40-51,71-80,101-112,125-161,199-239.
"""
import math
import os

import numpy as np
import torch
from torch import nn

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data")

PT_BLOCKS = [2, 3, 4, 6, 3]
PT_STRIDE = [1, 4, 4, 4, 4]
PT_NSAMPLE = [8, 16, 16, 16, 16]
PT_PLANES = {"confidence_encoder": [128, 128, 256, 256, 512], "magnitude_encoder": [64, 128, 256, 256, 512]}
SHARE_PLANES = 8


def so3_tables():
    """anchors [60,3,3] f32, intra_idx [60,12] i64, kpsphere24 [24,3] f32 (tools/gen_so3_tables.py)."""
    t = np.load(os.path.join(_DATA, "so3_tables.npz"))
    return {k: t[k] for k in t.files}


def epn_layers(input_radius=0.4, n_layers=2):
    """Per-conv hyper-parameters, replaying build_model (src/models/so3net.py:36-152); input_num stays 1024."""
    mlps = [[32, 32], [64, 64], [128, 128], [256, 256]][:n_layers]
    strides = [2, 2, 2, 2][:n_layers]
    input_num, r0, samp_ratio, samp_density, sig_ratio = 1024, 0.2, 0.8, 0.5, 0.5
    mult = [2 ** i for i in range(n_layers + 1)]
    centers = [int(input_num / m) for m in mult]
    ratio = [r0 * m ** samp_density for m in mult]
    radii = [r * input_radius for r in ratio]
    sigma = [sig_ratio * radii[0] ** 2]
    for i, s in enumerate(strides):
        sigma.append(sigma[i] * s)
    out, dim_in = [], 1
    for i, block in enumerate(mlps):
        for j, dim_out in enumerate(block):
            nn_ = int(samp_ratio * centers[i] * ratio[i] ** (1 / samp_density))
            if i == 0 and j == 0:
                nn_ *= int(input_num / 1024)
            if j == 0:
                stride, nidx = strides[i], (i if i == 0 else i + 1)
                nn_ *= 2
            else:
                stride, nidx = 1, i + 1
            out.append(dict(block=i, conv=j, dim_in=dim_in, dim_out=dim_out, stride=stride, radius=radii[nidx],
                            sigma=sigma[nidx], n_neighbor=nn_, lazy_sample=(i != 0 or j != 0), kernel_size=1,
                            kanchor=60, activation="leaky_relu"))
            dim_in = dim_out
    return out


def scaled_kernel_points(kp24, radius):
    """kpsphere24 rescaled so that the farthest point sits at 0.7*radius (functional.py:146-157, modules.py:13,99)."""
    pc = kp24.astype("float32")
    r = np.sqrt((pc ** 2).sum(1).max())
    return pc * (0.7 * radius) / r


# ----------------------------------------------------------------------------- leaf helpers
def P(*shape, init="default", fan=None):
    return ("param", tuple(shape), torch.float32, (init, fan))


def Bf(value):
    return ("buffer", tuple(value.shape), value.dtype, ("value", value))


def _linear(cin, cout, bias=True):
    d = {"weight": P(cout, cin, init="linear", fan=cin)}
    if bias:
        d["bias"] = P(cout, init="linear_bias", fan=cin)
    return d


def _conv1d(cin, cout, groups=1):
    return {"weight": P(cout, cin // groups, 1, init="linear", fan=cin // groups),
            "bias": P(cout, init="linear_bias", fan=cin // groups)}


def _bn(c):
    return {"weight": P(c, init="ones"), "bias": P(c, init="zeros"),
            "running_mean": Bf(torch.zeros(c)), "running_var": Bf(torch.ones(c)),
            "num_batches_tracked": Bf(torch.tensor(0, dtype=torch.long))}


def _pt_layer(c, share=SHARE_PLANES):
    return {"linear_q": _linear(c, c), "linear_k": _linear(c, c), "linear_v": _linear(c, c),
            "linear_p": {"0": _linear(3, 3), "1": _bn(3), "3": _linear(3, c)},
            "linear_w": {"0": _bn(c), "2": _linear(c, c // share), "3": _bn(c // share), "5": _linear(c // share, c // share)}}


def _pt_block(c):
    return {"linear1": _linear(c, c, False), "bn1": _bn(c), "transformer2": _pt_layer(c), "bn2": _bn(c),
            "linear3": _linear(c, c, False), "bn3": _bn(c)}


def _pt_net(c_in, planes, head):
    d, cur = {}, c_in
    for lvl in range(5):
        enc = {"0": {"linear": _linear((3 + cur) if PT_STRIDE[lvl] != 1 else cur, planes[lvl], False), "bn": _bn(planes[lvl])}}
        cur = planes[lvl]
        for b in range(1, PT_BLOCKS[lvl]):
            enc[str(b)] = _pt_block(cur)
        d["enc%d" % (lvl + 1)] = enc
    for lvl in range(4, -1, -1):
        if lvl == 4:
            up = {"linear1": {"0": _linear(2 * cur, cur), "1": _bn(cur)}, "linear2": {"0": _linear(cur, cur)}}
        else:
            up = {"linear1": {"0": _linear(planes[lvl], planes[lvl]), "1": _bn(planes[lvl])},
                  "linear2": {"0": _linear(cur, planes[lvl]), "1": _bn(planes[lvl])}}
        cur = planes[lvl]
        d["dec%d" % (lvl + 1)] = {"0": up, "1": _pt_block(cur)}
    d.update(head)
    return d


def network_spec(n_markers=86, input_radius=0.4, n_layers=2):
    tabs = so3_tables()
    anchors = torch.from_numpy(tabs["anchors"])
    intra_idx = torch.from_numpy(tabs["intra_idx"]).long()
    backbone = {}
    for lp in epn_layers(input_radius, n_layers):
        kern = torch.from_numpy(scaled_kernel_points(tabs["kpsphere24"], lp["radius"]))
        ci, co = lp["dim_in"], lp["dim_out"]
        blk = {"inter_conv": {"conv": {"basic_conv": {"W": P(co, ci * 24, init="so3", fan=(ci, co, 24)),
                                                       "bias": P(1, co, 1, init="const1e-3")},
                                       "anchors": Bf(anchors.clone()), "kernels": Bf(kern)}},
               "intra_conv": {"conv": {"basic_conv": {"W": P(co, co * 12, init="so3", fan=(co, co, 12)),
                                                       "bias": P(1, co, 1, init="const1e-3")},
                                       "anchors": Bf(anchors.clone()), "intra_idx": Bf(intra_idx.clone())}},
               "skip_conv": {"weight": P(co, ci, 1, 1, init="linear", fan=ci), "bias": P(co, init="linear_bias", fan=ci)}}
        backbone.setdefault(str(lp["block"]), {"blocks": {}})["blocks"][str(lp["conv"])] = blk
    feat = [32, 64, 128, 256][n_layers - 1]

    def mha(value_dim):
        return {"key_transform": _linear(feat, feat, False), "query_transform": _linear(feat, feat, False),
                "value_transform": _linear(feat, feat, False), "head_combine": _linear(feat, value_dim)}

    k = n_markers
    return {
        "encoder": {"backbone": backbone},
        "direction_encoder": {"self_attention_layers": {"0": mha(feat), "1": mha(128)}},
        "direction_predictor": {"net": {"0": _linear(128, 128), "2": _linear(128, 128)}},
        "so3_reg": _conv1d(128, 1),
        "magnitude_encoder": _pt_net(feat + 3, PT_PLANES["magnitude_encoder"], {
            "final_layer": {"0": _linear(64, 64), "1": _bn(64), "3": _linear(64, 1)}}),
        "confidence_encoder": _pt_net(feat + 3, PT_PLANES["confidence_encoder"], {
            "cls": {"0": _conv1d(128, 128), "1": _bn(128), "3": _conv1d(128, k)},
            "confi": {"0": _conv1d(128, 128 * k), "2": _conv1d(128 * k, k, groups=k)}}),
    }


def _init_tensor(shape, dtype, init):
    kind, fan = init
    t = torch.empty(shape, dtype=dtype)
    if kind == "ones":
        return t.fill_(1.0)
    if kind == "zeros":
        return t.zero_()
    if kind == "const1e-3":
        return t.fill_(1e-3)
    if kind == "so3":  # BasicSO3Conv: xavier_normal_(gain=relu) on [c_out, c_in, ks] then flattened (modules.py:24-27)
        ci, co, ks = fan
        std = math.sqrt(2.0) * math.sqrt(2.0 / float(ci * ks + co * ks))
        return t.normal_(0.0, std)
    if kind == "linear":  # kaiming_uniform_(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in))
        b = 1.0 / math.sqrt(fan)
        return t.uniform_(-b, b)
    if kind == "linear_bias":
        b = 1.0 / math.sqrt(fan)
        return t.uniform_(-b, b)
    raise ValueError(kind)


class ParamTree(nn.Module):
    """Pure container: registers parameters/buffers/children exactly under the names of the spec."""

    def __init__(self, spec):
        super().__init__()
        for name, leaf in spec.items():
            if isinstance(leaf, dict):
                self.add_module(name, ParamTree(leaf))
            elif leaf[0] == "param":
                self.register_parameter(name, nn.Parameter(_init_tensor(leaf[1], leaf[2], leaf[3]), requires_grad=False))
            else:
                self.register_buffer(name, leaf[3][1].clone())


def xavier_reset_(module):
    """GT_network_equiv._reset_parameters (models_pointcloud.py:72-77): xavier_uniform_ on every >1-D parameter."""
    for p in module.parameters():
        if p.dim() > 1:
            nn.init.xavier_uniform_(p)
