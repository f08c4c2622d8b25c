"""torch-CPU (fp32) restatement of ETCH's network forward -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Takes a ``state_dict`` in the reference key layout and reproduces
``GT_network_equiv.forward`` (src/models/models_pointcloud.py:146-221) stage by stage, returning every
intermediate the CUDA path is checked against.  Native index kernels come from oracle/index_ops.py (C).
Pinned against goldens produced by importing the unmodified reference Python (tools/gen_golden.py).

Each function cites the reference lines it restates.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import index_ops as ops

EPS_IN = 1e-5  # nn.InstanceNorm2d default (src/models/so3conv.py:27,85,168)
EPS_BN = 1e-5  # nn.BatchNorm1d default


# ----------------------------------------------------------------------------- layer constants
def epn_layer_params(input_radius=0.4, n_layers=2):
    """Replays build_model (src/models/so3net.py:36-152) for mlps [[32,32],[64,64]], strides [2,2], input_num 1024."""
    mlps = [[32, 32], [64, 64], [128, 128], [256, 256]][:n_layers]
    strides = [2, 2, 2, 2][:n_layers]
    input_num, initial_radius_ratio, sampling_ratio, sampling_density, sigma_ratio = 1024, 0.2, 0.8, 0.5, 0.5
    stride_mult = [1]
    for _ in range(len(mlps)):
        stride_mult.append(stride_mult[-1] * 2)
    num_centers = [int(input_num / m) for m in stride_mult]
    radius_ratio = [initial_radius_ratio * m ** sampling_density for m in stride_mult]
    radii = [r * input_radius for r in radius_ratio]
    weighted_sigma = [sigma_ratio * radii[0] ** 2]
    for i, s in enumerate(strides):
        weighted_sigma.append(weighted_sigma[i] * s)
    layers = []
    dim_in = 1
    for i, block in enumerate(mlps):
        for j, dim_out in enumerate(block):
            lazy = i != 0 or j != 0
            neighbor = int(sampling_ratio * num_centers[i] * radius_ratio[i] ** (1 / sampling_density))
            if i == 0 and j == 0:
                neighbor *= int(input_num / 1024)
            if j == 0:
                stride, nidx = strides[i], (i if i == 0 else i + 1)
                neighbor *= 2
            else:
                stride, nidx = 1, i + 1
            layers.append(dict(block=i, conv=j, dim_in=dim_in, dim_out=dim_out, stride=stride, radius=radii[nidx],
                               sigma=weighted_sigma[nidx], n_neighbor=neighbor, lazy_sample=lazy))
            dim_in = dim_out
    return layers


def scaled_kernels(kpsphere24, radius):
    """get_sphereical_kernel_points_from_ply(0.7*radius, 1) (vgtk/so3conv/functional.py:146-157; modules.py:13,99)."""
    pc = kpsphere24.astype("float32")
    r = np.sqrt((pc ** 2).sum(1).max())
    return pc * (0.7 * radius) / r


# ----------------------------------------------------------------------------- encoder
def _instnorm_lrelu(x):
    """InstanceNorm2d(affine=False) + leaky_relu(0.01) on [B,C,P,A] (src/models/so3conv.py:38-40,96-99,181)."""
    mean = x.mean(dim=(2, 3), keepdim=True)
    var = x.var(dim=(2, 3), unbiased=False, keepdim=True)
    return F.leaky_relu((x - mean) / torch.sqrt(var + EPS_IN), 0.01)


def inter_conv(xyz, feats, lp, anchors, kernels, W, bias):
    """InterSO3Conv.forward (vgtk/so3conv/modules.py:120-128; functional.py:176-185,224-324,61-67; modules.py:33-39).
    xyz [B,3,q]; feats [B,c,q,60] -> (new_xyz [B,3,P], z [B,c_out,P,60] pre-norm, sample_idx, ball_idx)."""
    B, _, q = xyz.shape
    P = math.ceil(q / lp["stride"])
    if q == P or lp["lazy_sample"]:
        sample_idx = torch.arange(P, dtype=torch.int32).view(1, -1).expand(B, -1).contiguous()
    else:
        sample_idx = torch.from_numpy(ops.fps_bcn(xyz.numpy(), P))
    new_xyz = torch.from_numpy(ops.gather_bcn(xyz.numpy(), sample_idx.numpy()))
    ball_idx = torch.from_numpy(ops.ball_query_bcn(new_xyz.numpy(), xyz.numpy(), lp["radius"], lp["n_neighbor"]))
    nn_ = lp["n_neighbor"]
    li = ball_idx.long()
    grouped = torch.gather(xyz.unsqueeze(2).expand(-1, -1, P, -1), 3, li.unsqueeze(1).expand(-1, 3, -1, -1))
    grouped = grouped - new_xyz.unsqueeze(3)  # [B,3,P,nn]
    rk = torch.matmul(anchors, kernels.t()).permute(1, 0, 2).contiguous()  # [3,60,24]
    d = ((grouped[..., None, None, :] - rk[None, :, None, :, :, None]) ** 2).sum(1)  # [B,P,60,24,nn]
    w = F.relu(1.0 - d / lp["sigma"])
    c = feats.shape[1]
    nf = torch.gather(feats, 2, li.reshape(B, 1, P * nn_, 1).expand(-1, c, -1, 60)).view(B, c, P, nn_, 60)
    y = torch.einsum("bcpna,bpakn->bckpa", nf, w).contiguous()  # [B,c,24,P,60]
    z = torch.matmul(W, y.view(B, c * 24, P * 60)) + bias
    return new_xyz, z.view(B, -1, P, 60), sample_idx, ball_idx


def intra_conv(x, intra_idx, W, bias):
    """IntraSO3Conv.forward (modules.py:150-153; functional.py:331-343)."""
    B, c, P, A = x.shape
    g = x.index_select(3, intra_idx.view(-1)).view(B, c, P, A, 12).permute(0, 1, 4, 2, 3).contiguous()
    z = torch.matmul(W, g.view(B, c * 12, P * A)) + bias
    return z.view(B, -1, P, A)


def encoder(hitpts, sd, tables, input_radius=0.4, n_layers=2, trace=None):
    """EquivBackbone.forward (src/models/so3net.py:23-33) + SeparableSO3ConvBlock (src/models/so3conv.py:171-183)."""
    B, N, _ = hitpts.shape
    xyz = hitpts.permute(0, 2, 1).contiguous()
    feats = torch.ones(B, 1, N, 60)
    anchors = torch.from_numpy(tables["anchors"])
    intra_idx = torch.from_numpy(tables["intra_idx"]).long()
    for lp in epn_layer_params(input_radius, n_layers):
        pre = "encoder.backbone.%d.blocks.%d." % (lp["block"], lp["conv"])
        kernels = torch.from_numpy(scaled_kernels(tables["kpsphere24"], lp["radius"]))
        skip = feats
        xyz, z, sample_idx, ball_idx = inter_conv(xyz, feats, lp, anchors, kernels, sd[pre + "inter_conv.conv.basic_conv.W"],
                                                  sd[pre + "inter_conv.conv.basic_conv.bias"])
        x = _instnorm_lrelu(z)
        z2 = intra_conv(x, intra_idx, sd[pre + "intra_conv.conv.basic_conv.W"], sd[pre + "intra_conv.conv.basic_conv.bias"])
        x2 = _instnorm_lrelu(z2)
        if lp["stride"] > 1:
            skip = torch.gather(skip, 2, sample_idx.long()[:, None, :, None].expand(-1, skip.shape[1], -1, 60))
        s = F.conv2d(skip, sd[pre + "skip_conv.weight"], sd[pre + "skip_conv.bias"])
        feats = x2 + _instnorm_lrelu(s)
        if trace is not None:
            trace.append(dict(sample_idx=sample_idx, ball_idx=ball_idx, xyz=xyz, inter_z=z, intra_z=z2, skip_z=s, out=feats))
    return xyz, feats, anchors


# ----------------------------------------------------------------------------- upsample + direction head
def upsample_3nn(xyz1, xyz2, points2):
    """PointFeatPropagation (src/models/pointnet2_utils.py:45-74,4-23). xyz1 [B,N,3], xyz2 [B,S,3], points2 [B,S,D]."""
    d = -2 * torch.matmul(xyz1, xyz2.permute(0, 2, 1))
    d += torch.sum(xyz1 ** 2, -1).unsqueeze(-1)
    d += torch.sum(xyz2 ** 2, -1).unsqueeze(1)
    d, idx = d.sort(dim=-1)
    d, idx = d[:, :, :3], idx[:, :, :3]
    rec = 1.0 / (d + 1e-8)
    w = rec / rec.sum(2, keepdim=True)
    B = xyz1.shape[0]
    g = points2[torch.arange(B).view(B, 1, 1), idx]  # [B,N,3,D]
    return (g * w.unsqueeze(-1)).sum(2), idx, w


def mhsa(x, sd, pre, heads=8):
    """MultiHeadAttention.forward (src/models/direction_backbones.py:129-194,79-126). x [T,60,64]."""
    k = x @ sd[pre + "key_transform.weight"].t()
    q = x @ sd[pre + "query_transform.weight"].t()
    v = x @ sd[pre + "value_transform.weight"].t()
    T, L, E = k.shape
    hs = E // heads

    def split(o):
        return o.view(T, L, heads, hs).permute(2, 0, 1, 3).reshape(heads * T, L, hs)

    k, q, v = split(k), split(q), split(v)
    att = torch.softmax(torch.bmm(q, k.transpose(1, 2)) / np.sqrt(hs), dim=-1)
    o = torch.bmm(att, v).view(heads, T, L, hs).permute(1, 2, 0, 3).reshape(T, L, E)
    return o @ sd[pre + "head_combine.weight"].t() + sd[pre + "head_combine.bias"]


def so3_mean(anchors, w):
    """so3_mean (src/models/so3conv.py:186-225): chordal mean via SVD. anchors [60,3,3], w [T,60] -> [T,3,3]."""
    Ce = torch.einsum("ta,aij->tij", w, anchors)
    U, S, V = torch.svd(Ce)
    Vt = V.transpose(1, 2)
    det = torch.det(U @ Vt)
    D = torch.eye(3).repeat(Ce.shape[0], 1, 1)
    D[:, 2, 2] = det
    return U @ D @ Vt


def direction_head(point_equiv, anchors, sd):
    """decode_direction (src/models/models_pointcloud.py:111-126). point_equiv [B,N,64,60] -> dir [B,N,3], w [B*N,60]."""
    B, N, C, A = point_equiv.shape
    x = point_equiv.permute(0, 1, 3, 2).reshape(-1, A, C)
    x = x + mhsa(x, sd, "direction_encoder.self_attention_layers.0.")
    x = mhsa(x, sd, "direction_encoder.self_attention_layers.1.")
    x = F.relu(x @ sd["direction_predictor.net.0.weight"].t() + sd["direction_predictor.net.0.bias"])
    x = x @ sd["direction_predictor.net.2.weight"].t() + sd["direction_predictor.net.2.bias"]
    w = (x @ sd["so3_reg.weight"].view(1, -1).t()).squeeze(-1) + sd["so3_reg.bias"]
    R = so3_mean(anchors, w)
    return R[:, :, 2].reshape(B, N, 3), w


# ----------------------------------------------------------------------------- PointTransformer
def _bn(x, sd, pre):
    return (x - sd[pre + "running_mean"]) / torch.sqrt(sd[pre + "running_var"] + EPS_BN) * sd[pre + "weight"] + sd[pre + "bias"]


def _lin(x, sd, pre, bias=True):
    y = x @ sd[pre + "weight"].t()
    return y + sd[pre + "bias"] if bias else y


def _knn(k, p_src, p_query, o_src, o_query):
    idx, d2 = ops.knn_packed(k, p_src.numpy(), p_query.numpy(), o_src.numpy(), o_query.numpy())
    return torch.from_numpy(idx).long(), torch.from_numpy(d2)


def pt_layer(p, x, o, sd, pre, nsample, share=8):
    """PointTransformerLayer.forward (src/models/pointtransformer_seg.py:24-37)."""
    q, k, v = _lin(x, sd, pre + "linear_q."), _lin(x, sd, pre + "linear_k."), _lin(x, sd, pre + "linear_v.")
    idx, _ = _knn(nsample, p, p, o, o)
    n, c = x.shape
    p_r = p[idx.view(-1)].view(n, nsample, 3) - p.unsqueeze(1)
    xk, xv = k[idx.view(-1)].view(n, nsample, c), v[idx.view(-1)].view(n, nsample, c)
    p_r = _lin(p_r, sd, pre + "linear_p.0.")
    p_r = F.relu(_bn(p_r, sd, pre + "linear_p.1."))
    p_r = _lin(p_r, sd, pre + "linear_p.3.")
    w = xk - q.unsqueeze(1) + p_r
    w = F.relu(_bn(w, sd, pre + "linear_w.0."))
    w = _lin(w, sd, pre + "linear_w.2.")
    w = F.relu(_bn(w, sd, pre + "linear_w.3."))
    w = _lin(w, sd, pre + "linear_w.5.")
    w = torch.softmax(w, dim=1)
    return ((xv + p_r).view(n, nsample, share, c // share) * w.unsqueeze(2)).sum(1).view(n, c)


def pt_block(p, x, o, sd, pre, nsample):
    """PointTransformerBlock.forward (pointtransformer_seg.py:114-122)."""
    y = F.relu(_bn(_lin(x, sd, pre + "linear1.", False), sd, pre + "bn1."))
    y = F.relu(_bn(pt_layer(p, y, o, sd, pre + "transformer2.", nsample), sd, pre + "bn2."))
    y = _bn(_lin(y, sd, pre + "linear3.", False), sd, pre + "bn3.")
    return F.relu(y + x)


def pt_down(p, x, o, sd, pre, stride, nsample):
    """TransitionDown.forward (pointtransformer_seg.py:52-68; src/models/pointops.py:10-27,79-100)."""
    if stride == 1:
        return p, F.relu(_bn(_lin(x, sd, pre + "linear.", False), sd, pre + "bn.")), o
    on = o.numpy()
    seg = np.diff(np.concatenate([[0], on]))
    n_o = torch.from_numpy(np.cumsum(seg // stride).astype(np.int32))
    idx = torch.from_numpy(ops.fps_packed(p.numpy(), on, n_o.numpy())).long()
    n_p = p[idx]
    kidx, _ = _knn(nsample, p, n_p, o, n_o)
    m = n_p.shape[0]
    g = torch.cat([p[kidx.view(-1)].view(m, nsample, 3) - n_p.unsqueeze(1), x[kidx.view(-1)].view(m, nsample, -1)], -1)
    y = F.relu(_bn(_lin(g, sd, pre + "linear.", False), sd, pre + "bn."))
    return n_p, y.max(1)[0], n_o


def pt_up(pxo1, pxo2, sd, pre):
    """TransitionUp.forward (pointtransformer_seg.py:81-98; pointops.interpolation src/models/pointops.py:164-178)."""
    if pxo2 is None:
        _, x, o = pxo1
        outs, s = [], 0
        for e in o.tolist():
            xb = x[s:e]
            g = F.relu(_lin(xb.sum(0, True) / (e - s), sd, pre + "linear2.0."))
            outs.append(torch.cat([xb, g.repeat(e - s, 1)], 1))
            s = e
        x = torch.cat(outs, 0)
        return F.relu(_bn(_lin(x, sd, pre + "linear1.0."), sd, pre + "linear1.1."))
    p1, x1, o1 = pxo1
    p2, x2, o2 = pxo2
    a = F.relu(_bn(_lin(x1, sd, pre + "linear1.0."), sd, pre + "linear1.1."))
    f = F.relu(_bn(_lin(x2, sd, pre + "linear2.0."), sd, pre + "linear2.1."))
    idx, d2 = _knn(3, p2, p1, o2, o1)
    rec = 1.0 / (torch.sqrt(d2) + 1e-8)
    w = rec / rec.sum(1, keepdim=True)
    out = torch.zeros(p1.shape[0], f.shape[1])
    for i in range(3):
        out += f[idx[:, i]] * w[:, i].unsqueeze(-1)
    return a + out


PT_BLOCKS = [2, 3, 4, 6, 3]
PT_STRIDE = [1, 4, 4, 4, 4]
PT_NSAMPLE = [8, 16, 16, 16, 16]


def point_transformer_body(p0, x0, o0, sd, pre):
    """PointTransformer_{confidence,magnitude}.forward up to x1 (pointtransformer_seg.py:163-179,241-254)."""
    x = torch.cat([p0, x0], 1)
    p, o = p0, o0
    enc = []
    for lvl in range(5):
        e = pre + "enc%d." % (lvl + 1)
        p, x, o = pt_down(p, x, o, sd, e + "0.", PT_STRIDE[lvl], PT_NSAMPLE[lvl])
        for b in range(1, PT_BLOCKS[lvl]):
            x = pt_block(p, x, o, sd, e + "%d." % b, PT_NSAMPLE[lvl])
        enc.append((p, x, o))
    p5, x5, o5 = enc[4]
    x5 = pt_block(p5, pt_up((p5, x5, o5), None, sd, pre + "dec5.0."), o5, sd, pre + "dec5.1.", PT_NSAMPLE[4])
    prev = (p5, x5, o5)
    for lvl in range(3, -1, -1):
        pl, xl, ol = enc[lvl]
        d = pre + "dec%d." % (lvl + 1)
        xl = pt_block(pl, pt_up((pl, xl, ol), prev, sd, d + "0."), ol, sd, d + "1.", PT_NSAMPLE[lvl])
        prev = (pl, xl, ol)
    return prev[1], [e[0] for e in enc]


def confidence_head(x1, sd, pre="confidence_encoder.", k=86):
    """cls / confi heads (pointtransformer_seg.py:144-145,181-192). x1 [n,128] -> labels [n,86], conf [n]."""
    h = F.relu(_bn(_lin(x1, _sq(sd, pre + "cls.0."), ""), sd, pre + "cls.1."))
    logits = _lin(h, _sq(sd, pre + "cls.3."), "")
    sm = torch.softmax(logits, dim=1)
    g = F.relu(_lin(x1, _sq(sd, pre + "confi.0."), ""))  # [n, 128*k]
    w2 = sd[pre + "confi.2.weight"].view(k, -1)  # grouped conv: [k,128]
    confi = (g.view(-1, k, w2.shape[1]) * w2.unsqueeze(0)).sum(-1) + sd[pre + "confi.2.bias"]
    return logits, (confi * sm).sum(1)


def _sq(sd, pre):
    return {"weight": sd[pre + "weight"].squeeze(-1), "bias": sd[pre + "bias"]}


def magnitude_head(x1, sd, pre="magnitude_encoder."):
    """final_layer (pointtransformer_seg.py:222,256)."""
    h = F.relu(_bn(_lin(x1, sd, pre + "final_layer.0."), sd, pre + "final_layer.1."))
    return _lin(h, sd, pre + "final_layer.3.")


# ----------------------------------------------------------------------------- full forward
def forward(hitpts, sd, tables, input_radius=0.4, n_layers=2, return_trace=False):
    """GT_network_equiv.forward(hitpts, ['confidence','direction','magnitude'], 'standard_vector')
    (src/models/models_pointcloud.py:146-221)."""
    sd = {k: v.float() if v.is_floating_point() else v for k, v in sd.items()}
    B, N, _ = hitpts.shape
    trace = {"enc": []}
    xyz2, feats, anchors = encoder(hitpts, sd, tables, input_radius, n_layers, trace["enc"])
    S = xyz2.shape[-1]
    equiv = feats.permute(0, 1, 3, 2).reshape(B, -1, S)  # [B,3840,S], channel = c*60+a
    up, up_idx, up_w = upsample_3nn(hitpts, xyz2.permute(0, 2, 1), equiv.permute(0, 2, 1))
    point_equiv = up.reshape(B, N, -1, 60)
    inv = point_equiv.mean(-1)
    p = hitpts.reshape(-1, 3).contiguous()
    x = inv.reshape(B * N, -1).contiguous()
    o = torch.tensor([N * (i + 1) for i in range(B)], dtype=torch.int32)
    xc, pt_points = point_transformer_body(p, x, o, sd, "confidence_encoder.")
    logits, conf = confidence_head(xc, sd)
    direction, anc_w = direction_head(point_equiv, anchors, sd)
    xm, _ = point_transformer_body(p, x, o, sd, "magnitude_encoder.")
    mag = magnitude_head(xm, sd)
    out = {"confidences": conf.view(B, N, 1), "part_labels": logits.view(B, N, -1), "direction": direction,
           "magnitude": mag.view(B, N, 1)}
    if return_trace:
        trace.update(up_idx=up_idx, up_w=up_w, inv=inv, anc_w=anc_w.view(B, N, 60), xc=xc, xm=xm, xyz2=xyz2, feats=feats,
                     pt_points=pt_points)
        return out, trace
    return out


def postprocess(hitpts, out, scale_magnitude=10.0):
    """eval.py:103,116,183 / inference_demo.py:52-59."""
    labels = out["part_labels"].argmax(-1)
    vec = out["direction"] * out["magnitude"] / scale_magnitude
    return labels, vec, hitpts - vec
