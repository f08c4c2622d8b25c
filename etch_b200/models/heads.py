"""Host-side drivers of the three decoding heads: direction (MHSA over anchors), magnitude and marker confidence
(PointTransformer U-Nets).  Weight folding (eval BatchNorm, algebraic fusions) happens once per state_dict; the forward
functions only allocate buffers and launch kernels through the C ABI.

Reference modules restated: src/models/direction_backbones.py:129-223, src/models/pointtransformer_seg.py (all),
src/models/models_pointcloud.py:94-126.
"""
import ctypes
import math
import os

import torch

from .. import _lib as L
from . import tc
from .spec import PT_BLOCKS, PT_NSAMPLE, PT_STRIDE

EPS_BN = 1e-5
USE_TC = os.environ.get("ETCH_B200_NO_TC", "0") != "1"  # tcgen05 kernels (default) vs fp32 CUDA-core versions
KNN_GRID = os.environ.get("ETCH_B200_KNN", "grid") == "grid"   # kNN through a uniform grid (default) or the brute-force scan
PT_ATTN_TC = os.environ.get("ETCH_B200_PT_ATTN", "tc") == "tc"  # vector attention: tensor-core tiles (default) or the CTA-per-point kernel


def _upload(obj, device):
    """Move every tensor reachable from a plan object (attributes, dicts, lists, tuples) to `device`, in place."""
    if torch.is_tensor(obj):
        return obj.to(device)
    if isinstance(obj, dict):
        for k in obj:
            obj[k] = _upload(obj[k], device)
        return obj
    if isinstance(obj, list):
        for i in range(len(obj)):
            obj[i] = _upload(obj[i], device)
        return obj
    if isinstance(obj, tuple):
        return tuple(_upload(o, device) for o in obj)
    if hasattr(obj, "__dict__") and not isinstance(obj, type):
        for k, v in vars(obj).items():
            setattr(obj, k, _upload(v, device))
    return obj


_HOST = dict(dtype=torch.float32, device="cpu")   # plans are computed on the host and uploaded once (_upload)


# ----------------------------------------------------------------------------- direction head
class DirectionPlan:
    def __init__(self, sd, device, anchors):
        f64 = lambda k: sd[k].detach().cpu().to(torch.float64)  # noqa: E731
        dev = _HOST

        def qkv(layer):
            pre = "direction_encoder.self_attention_layers.%d." % layer
            wq, wk, wv = f64(pre + "query_transform.weight"), f64(pre + "key_transform.weight"), f64(pre + "value_transform.weight")
            head = wq.shape[0] // 8
            return torch.cat([wq / math.sqrt(head), wk, wv], 0).t().contiguous()  # [64][192]

        self.Wqkv1 = qkv(0).to(**dev)
        self.Wc1 = f64("direction_encoder.self_attention_layers.0.head_combine.weight").t().contiguous().to(**dev)
        self.bc1 = f64("direction_encoder.self_attention_layers.0.head_combine.bias").to(**dev)
        self.Wqkv2 = qkv(1).to(**dev)
        wc2, bc2 = f64("direction_encoder.self_attention_layers.1.head_combine.weight"), f64("direction_encoder.self_attention_layers.1.head_combine.bias")
        w1, b1 = f64("direction_predictor.net.0.weight"), f64("direction_predictor.net.0.bias")
        w2, b2 = f64("direction_predictor.net.2.weight"), f64("direction_predictor.net.2.bias")
        wr, br = f64("so3_reg.weight").reshape(-1), f64("so3_reg.bias").reshape(-1)
        self.Wf = (w1 @ wc2).t().contiguous().to(**dev)           # [64][128]  Linear1 o head_combine
        self.bf = (w1 @ bc2 + b1).to(**dev)
        self.vreg = (w2.t() @ wr).contiguous().to(**dev)          # [128]      so3_reg o Linear2
        self.creg = float(b2 @ wr + br[0])
        self.anchors = anchors.detach().cpu().reshape(60, 9).contiguous().to(**dev)
        # tensor-core operands: 9 blocks of 64 output rows x 64 inputs, in consumption order
        #   layer 0: Wq/sqrt(dk), Wk, Wv (2 slices each), head_combine (2); layer 1: Wq/sqrt(dk), Wk, Wv (6); Linear1 o head_combine (4)
        def rows(layer):
            pre = "direction_encoder.self_attention_layers.%d." % layer
            wq, wk, wv = f64(pre + "query_transform.weight"), f64(pre + "key_transform.weight"), f64(pre + "value_transform.weight")
            return [wq / math.sqrt(wq.shape[0] // 8), wk, wv]
        mats = rows(0) + [f64("direction_encoder.self_attention_layers.0.head_combine.weight")] + rows(1) + [w1 @ wc2]
        allrows = torch.cat(mats, 0).float()          # [576, 64]
        assert allrows.shape == (576, 64)
        # 9 blocks of 64 output rows, each split in two K halves (32 inputs): [18][2][8][64][4]
        self.wall = torch.stack([tc.tc_operand(allrows[(i // 2) * 64:(i // 2 + 1) * 64, (i % 2) * 32:(i % 2 + 1) * 32], "cpu")
                                 for i in range(18)], 0).contiguous()
        _upload(self, device)


def run_direction(plan, hitpts, xyz2_b3s, feats2, want_anchor_weights=False):
    """hitpts [B,N,3], xyz2 [B,3,S], feats2 [B,S,60,64] -> direction [B,N,3], inv_feat [B,N,64] (+ anchor weights)."""
    B, N, _ = hitpts.shape
    S = xyz2_b3s.shape[2]
    dev = hitpts.device
    up_idx = torch.empty(B, N, 3, dtype=torch.int32, device=dev)
    up_w = torch.empty(B, N, 3, dtype=torch.float32, device=dev)
    L.call("upsample3", L.ptr(hitpts), L.ptr(xyz2_b3s), B, N, S, L.ptr(up_idx), L.ptr(up_w))
    direction = torch.empty(B, N, 3, dtype=torch.float32, device=dev)
    inv = torch.empty(B, N, 64, dtype=torch.float32, device=dev)
    anc_w = torch.empty(B, N, 60, dtype=torch.float32, device=dev) if want_anchor_weights else None
    if USE_TC:
        fmean = torch.empty(B * S * 64 + B * N * 18, dtype=torch.float32, device=dev)   # anchor means + [B,N,9] double Ce
        L.call("direction_head_tc", L.ptr(feats2), L.ptr(up_idx), L.ptr(up_w), L.ptr(plan.wall), L.ptr(plan.bc1), L.ptr(plan.bf),
               L.ptr(plan.vreg), L.f32(plan.creg), L.ptr(plan.anchors), B, N, S, L.ptr(direction), L.ptr(inv), L.ptr(anc_w), L.ptr(fmean))
    else:
        L.call("direction_head", L.ptr(feats2), L.ptr(up_idx), L.ptr(up_w), L.ptr(plan.Wqkv1), L.ptr(plan.Wc1), L.ptr(plan.bc1),
               L.ptr(plan.Wqkv2), L.ptr(plan.Wf), L.ptr(plan.bf), L.ptr(plan.vreg), L.f32(plan.creg), L.ptr(plan.anchors),
               B, N, S, L.ptr(direction), L.ptr(inv), L.ptr(anc_w))
    return direction, inv, anc_w, up_idx, up_w


# ----------------------------------------------------------------------------- PointTransformer
def _fold_bn(sd, pre, dev, lin_bias=None):
    """eval BatchNorm1d -> (scale, shift); an optional preceding linear bias is absorbed into the shift."""
    g, b = sd[pre + "weight"].double(), sd[pre + "bias"].double()
    m, v = sd[pre + "running_mean"].double(), sd[pre + "running_var"].double()
    scale = g / torch.sqrt(v + EPS_BN)
    shift = b - m * scale
    if lin_bias is not None:
        shift = shift + lin_bias.double() * scale
    return scale.to(**dev).contiguous(), shift.to(**dev).contiguous()


def _wt(w, dev):
    return w.detach().to(**dev).t().contiguous()


def _host_sd(sd):
    return {k: v.detach().cpu() for k, v in sd.items()} if any(v.is_cuda for v in sd.values()) else sd


class _Block:
    def __init__(self, sd, pre, dev):
        d = dev
        self.W1t = _wt(sd[pre + "linear1.weight"], d)
        self.s1, self.h1 = _fold_bn(sd, pre + "bn1.", d)
        t = pre + "transformer2."
        self.Wqkv = torch.cat([sd[t + "linear_q.weight"], sd[t + "linear_k.weight"], sd[t + "linear_v.weight"]], 0).to(**d).t().contiguous()
        self.bqkv = torch.cat([sd[t + "linear_q.bias"], sd[t + "linear_k.bias"], sd[t + "linear_v.bias"]], 0).to(**d).contiguous()
        ps, ph = _fold_bn(sd, t + "linear_p.1.", d, sd[t + "linear_p.0.bias"])
        self.P0 = (sd[t + "linear_p.0.weight"].to(**d) * ps[:, None]).contiguous()
        self.p0b = ph
        self.P3 = sd[t + "linear_p.3.weight"].to(**d).contiguous()
        self.p3b = sd[t + "linear_p.3.bias"].to(**d).contiguous()
        self.s0, self.h0 = _fold_bn(sd, t + "linear_w.0.", d)
        s3, h3 = _fold_bn(sd, t + "linear_w.3.", d, sd[t + "linear_w.2.bias"])
        self.Wa = (sd[t + "linear_w.2.weight"].to(**d) * s3[:, None]).contiguous()  # [T][c]
        self.ba = h3
        self.Wb = sd[t + "linear_w.5.weight"].to(**d).contiguous()                   # [T][T]
        self.bb = sd[t + "linear_w.5.bias"].to(**d).contiguous()
        self.so, self.ho = _fold_bn(sd, pre + "bn2.", d)
        self.W3t = _wt(sd[pre + "linear3.weight"], d)
        self.s3, self.h3 = _fold_bn(sd, pre + "bn3.", d)
        self.c = self.W1t.shape[0]
        # tensor-core attention (etch_pt_attention_tc): per-channel constants packed [c][8] and W1 as (hi, lo) canonical tiles
        # per 64-channel chunk, rows padded to the UMMA N granularity
        c = self.Wa.shape[1]
        T = self.Wa.shape[0]
        tp = max(T, 16)
        self.chan = torch.cat([self.P3, self.p3b[:, None], self.s0[:, None], self.h0[:, None], self.so[:, None], self.ho[:, None]], 1).contiguous()
        wa = torch.zeros(tp, c)
        wa[:T] = self.Wa.detach().float().cpu()
        self.Wa_c = torch.stack([tc.tc_operand(wa[:, k:k + 64].contiguous(), "cpu") for k in range(0, c, 64)], 0).contiguous()


class PTPlan:
    """Folded weights of one PointTransformer (prefix = 'confidence_encoder.' or 'magnitude_encoder.')."""

    def __init__(self, sd, prefix, device):
        d = _HOST
        sd = _host_sd(sd)
        self.enc, self.dec = [], []
        for lvl in range(5):
            e = prefix + "enc%d." % (lvl + 1)
            W = sd[e + "0.linear.weight"].to(**d)
            s, h = _fold_bn(sd, e + "0.bn.", d)
            if PT_STRIDE[lvl] == 1:
                down = dict(Wt=W.t().contiguous(), s=s, h=h)
            else:
                down = dict(Wp=W[:, :3].contiguous(), Wxt=W[:, 3:].t().contiguous(), s=s, h=h)
            blocks = [_Block(sd, e + "%d." % b, d) for b in range(1, PT_BLOCKS[lvl])]
            self.enc.append((down, blocks))
        for lvl in range(4, -1, -1):
            u = prefix + "dec%d.0." % (lvl + 1)
            if lvl == 4:
                W1 = sd[u + "linear1.0.weight"].to(**d)
                c = W1.shape[0]
                s, h = _fold_bn(sd, u + "linear1.1.", d)
                up = dict(head=True, W1at=W1[:, :c].t().contiguous(), W1bt=W1[:, c:].t().contiguous(),
                          b1=sd[u + "linear1.0.bias"].to(**d).contiguous(), s=s, h=h,
                          W2t=_wt(sd[u + "linear2.0.weight"], d), b2=sd[u + "linear2.0.bias"].to(**d).contiguous())
            else:
                s1, h1 = _fold_bn(sd, u + "linear1.1.", d, sd[u + "linear1.0.bias"])
                s2, h2 = _fold_bn(sd, u + "linear2.1.", d, sd[u + "linear2.0.bias"])
                up = dict(head=False, W1t=_wt(sd[u + "linear1.0.weight"], d), s1=s1, h1=h1,
                          W2t=_wt(sd[u + "linear2.0.weight"], d), s2=s2, h2=h2)
            self.dec.append((lvl, up, _Block(sd, prefix + "dec%d.1." % (lvl + 1), d)))
        if prefix.startswith("confidence"):
            s, h = _fold_bn(sd, prefix + "cls.1.", d, sd[prefix + "cls.0.bias"])
            k = sd[prefix + "cls.3.weight"].shape[0]
            w0 = sd[prefix + "confi.0.weight"].squeeze(-1).float().cpu()  # [K*128, 128]
            hi, lo = tc.split_tf32(w0)
            # [group l][row u (128)][K quarter kq (4)][k/4 (8)][4] -> [l][kq][k/4][u][4]: one [128 x 32] slice per (group, quarter)
            tile = lambda t: t.view(k, 128, 4, 8, 4).permute(0, 2, 3, 1, 4).reshape(k * 4, 8, 128, 4)  # noqa: E731
            W0c = torch.stack([tile(hi), tile(lo)], 1).contiguous()
            self.head = dict(kind="conf", K=k, W0c=W0c, Wc0t=_wt(sd[prefix + "cls.0.weight"].squeeze(-1), d), sc=s, hc=h,
                             Wc3t=_wt(sd[prefix + "cls.3.weight"].squeeze(-1), d), bc3=sd[prefix + "cls.3.bias"].to(**d).contiguous(),
                             W0t=_wt(sd[prefix + "confi.0.weight"].squeeze(-1), d), b0=sd[prefix + "confi.0.bias"].to(**d).contiguous(),
                             w2=sd[prefix + "confi.2.weight"].to(**d).reshape(k, -1).contiguous(),
                             b2=sd[prefix + "confi.2.bias"].to(**d).contiguous())
        else:
            s, h = _fold_bn(sd, prefix + "final_layer.1.", d, sd[prefix + "final_layer.0.bias"])
            self.head = dict(kind="mag", W0t=_wt(sd[prefix + "final_layer.0.weight"], d), s=s, h=h,
                             W3t=_wt(sd[prefix + "final_layer.3.weight"], d), b3=sd[prefix + "final_layer.3.bias"].to(**d).contiguous())
        _upload(self, device)


class PTGeometry:
    """Point hierarchy + the 13 neighbour graphs of one batch (shared by both PointTransformers: FPS and kNN depend
    only on the coordinates).  The reference recomputes them 88 times per forward (pointtransformer_seg.py:28-29,59-61,95)."""

    _OFFSETS = {}   # (sizes, device) -> cumulative-offset tensor, built once (no H2D copies inside a CUDA-graph capture)

    @classmethod
    def _offsets(cls, sizes, dev):
        key = (tuple(sizes), str(dev))
        t = cls._OFFSETS.get(key)
        if t is None:
            t = cls._OFFSETS[key] = torch.tensor([sum(sizes[:i + 1]) for i in range(len(sizes))], dtype=torch.int32, device=dev)
        return t

    def __init__(self, p0, B, N):
        dev = p0.device
        self.p, self.off, self.n = [p0], [], []
        sizes = [N] * B
        off = self._offsets(sizes, dev)
        self.off.append(off)
        self.n.append(sum(sizes))
        self.seg = [sizes]
        self.fps_idx, self.knn_self, self.knn_down, self.knn_up = [], [], [None], [None]
        for lvl in range(5):
            if lvl > 0:
                prev_p, prev_off, prev_sizes = self.p[lvl - 1], self.off[lvl - 1], self.seg[lvl - 1]
                sizes = [s // PT_STRIDE[lvl] for s in prev_sizes]
                m = sum(sizes)
                off = self._offsets(sizes, dev)
                idx = torch.empty(m, dtype=torch.int32, device=dev)
                L.call("fps_packed", B, max(prev_sizes), L.ptr(prev_p), L.ptr(prev_off), L.ptr(off), L.ptr(None), L.ptr(idx))
                newp = torch.empty(m, 3, dtype=torch.float32, device=dev)
                L.call("gather_rows", L.ptr(prev_p), L.ptr(idx), m, 3, L.ptr(newp))
                self.p.append(newp); self.off.append(off); self.n.append(m); self.seg.append(sizes)
                self.fps_idx.append(idx)
                self.knn_down.append(self._knn(PT_NSAMPLE[lvl], prev_p, newp, prev_off, off))       # new points in old cloud
                self.knn_up.append(self._knn(3, newp, prev_p, off, prev_off))                        # fine points in coarse cloud
            self.knn_self.append(self._knn(PT_NSAMPLE[lvl], self.p[lvl], self.p[lvl], self.off[lvl], self.off[lvl]))

    @staticmethod
    def _knn(k, src, qry, o_src, o_qry):
        m, n, nb = qry.shape[0], src.shape[0], int(o_src.shape[0])
        idx = torch.empty(m, k, dtype=torch.int32, device=src.device)
        d2 = torch.empty(m, k, dtype=torch.float32, device=src.device)
        if KNN_GRID and k in (3, 8, 16):
            # uniform-grid search with the brute-force results (bit-exact; ties fall back to the literal heap emulation)
            fn = L.lib().etch_knn_grid_scratch_bytes
            fn.restype = ctypes.c_longlong
            scratch = torch.empty(int(fn(n, nb)), dtype=torch.uint8, device=src.device)
            L.call("knn_grid", m, k, L.ptr(src), n, L.ptr(qry), L.ptr(o_src), L.ptr(o_qry), nb, L.ptr(idx), L.ptr(d2), L.ptr(scratch))
        else:
            L.call("knn_packed", m, k, L.ptr(src), L.ptr(qry), L.ptr(o_src), L.ptr(o_qry), nb, L.ptr(idx), L.ptr(d2))
        return idx, d2


def _tc_layout(Wt):
    """tensor-core block layout of a transposed weight, built on first use and kept ON the plan's tensor (freed and rebuilt
    with the plan: no global cache keyed on pointers)."""
    ent = getattr(Wt, "_etch_tc", None)
    if ent is None:
        ent = Wt._etch_tc = tc.tc_linear_weights(Wt, Wt.device)
    return ent


def _linear(X, Wt, scale=None, shift=None, R=None, relu=False, seg=None, seg_off=None):
    n, ci = X.shape
    co = Wt.shape[1]
    Y = torch.empty(n, co, dtype=torch.float32, device=X.device)
    if USE_TC and ci >= 16 and co >= 16 and ci <= 1024:
        ent = _tc_layout(Wt)
        L.call("linear_tc", L.ptr(X), ci, L.ptr(ent[0]), ent[1], n, ci, co, L.ptr(scale), L.ptr(shift), L.ptr(R), L.ptr(seg),
               L.ptr(seg_off), 0 if seg_off is None else int(seg_off.shape[0]), 1 if relu else 0, L.ptr(Y), co)
        return Y
    L.call("linear", L.ptr(X), ci, L.ptr(Wt), n, ci, co, L.ptr(scale), L.ptr(shift), L.ptr(R), L.ptr(seg), L.ptr(seg_off),
           0 if seg_off is None else int(seg_off.shape[0]), 1 if relu else 0, L.ptr(Y), co)
    return Y


def _block(blk, geo, lvl, x):
    """PointTransformerBlock (pointtransformer_seg.py:114-122)."""
    n, c = x.shape
    y = _linear(x, blk.W1t, blk.s1, blk.h1, relu=True)
    qkv = _linear(y, blk.Wqkv, None, blk.bqkv)
    a = torch.empty(n, c, dtype=torch.float32, device=x.device)
    idx = geo.knn_self[lvl][0]
    if USE_TC and PT_ATTN_TC:
        L.call("pt_attention_tc", L.ptr(geo.p[lvl]), L.ptr(qkv), L.ptr(idx), L.ptr(blk.P0), L.ptr(blk.p0b), L.ptr(blk.chan), L.ptr(blk.Wa_c),
               L.ptr(blk.ba), L.ptr(blk.Wb), L.ptr(blk.bb), n, int(idx.shape[1]), c, L.ptr(a))
    else:
        L.call("pt_attention", L.ptr(geo.p[lvl]), L.ptr(qkv), L.ptr(idx), L.ptr(blk.P0), L.ptr(blk.p0b), L.ptr(blk.P3), L.ptr(blk.p3b),
               L.ptr(blk.s0), L.ptr(blk.h0), L.ptr(blk.Wa), L.ptr(blk.ba), L.ptr(blk.Wb), L.ptr(blk.bb), L.ptr(blk.so), L.ptr(blk.ho),
               n, int(idx.shape[1]), c, L.ptr(a))
    return _linear(a, blk.W3t, blk.s3, blk.h3, R=x, relu=True)


def run_point_transformer(plan, geo, inv_feat_packed):
    """PointTransformer_{confidence,magnitude}.forward (pointtransformer_seg.py:163-195,241-260).
    inv_feat_packed [B*N,64]; returns x1 [B*N, planes[0]] and the head outputs."""
    p0 = geo.p[0]
    x = torch.cat([p0, inv_feat_packed], 1).contiguous()
    feats = []
    for lvl in range(5):
        down, blocks = plan.enc[lvl]
        if PT_STRIDE[lvl] == 1:
            x = _linear(x, down["Wt"], down["s"], down["h"], relu=True)
        else:
            yx = _linear(x, down["Wxt"])
            kidx = geo.knn_down[lvl][0]
            m, co = geo.n[lvl], yx.shape[1]
            out = torch.empty(m, co, dtype=torch.float32, device=x.device)
            L.call("pt_down_pool", L.ptr(geo.p[lvl - 1]), L.ptr(geo.p[lvl]), L.ptr(yx), L.ptr(kidx), L.ptr(down["Wp"]),
                   L.ptr(down["s"]), L.ptr(down["h"]), m, int(kidx.shape[1]), co, L.ptr(out))
            x = out
        for blk in blocks:
            x = _block(blk, geo, lvl, x)
        feats.append(x)
    prev = None
    for lvl, up, blk in plan.dec:
        xl = feats[lvl]
        if up["head"]:
            B = len(geo.seg[lvl])
            c = xl.shape[1]
            mean = torch.empty(B, c, dtype=torch.float32, device=xl.device)
            L.call("seg_mean", L.ptr(xl), L.ptr(geo.off[lvl]), B, c, L.ptr(mean))
            g = _linear(mean, up["W2t"], None, up["b2"], relu=True)
            segb = _linear(g, up["W1bt"], None, up["b1"])
            y = _linear(xl, up["W1at"], up["s"], up["h"], relu=True, seg=segb, seg_off=geo.off[lvl])
        else:
            a = _linear(xl, up["W1t"], up["s1"], up["h1"], relu=True)
            f = _linear(prev, up["W2t"], up["s2"], up["h2"], relu=True)
            idx, d2 = geo.knn_up[lvl + 1]
            y = torch.empty_like(a)
            L.call("pt_interp_add", L.ptr(a), L.ptr(f), L.ptr(idx), L.ptr(d2), a.shape[0], a.shape[1], L.ptr(y))
        prev = _block(blk, geo, lvl, y)
    x1 = prev
    h = plan.head
    if h["kind"] == "conf":
        hid = _linear(x1, h["Wc0t"], h["sc"], h["hc"], relu=True)
        logits = _linear(hid, h["Wc3t"], None, h["bc3"])
        conf = torch.empty(x1.shape[0], dtype=torch.float32, device=x1.device)
        if USE_TC:
            L.call("conf_head_tc", L.ptr(x1), L.ptr(logits), L.ptr(h["W0c"]), L.ptr(h["b0"]), L.ptr(h["w2"]), L.ptr(h["b2"]),
                   x1.shape[0], h["K"], L.ptr(conf))
        else:
            L.call("conf_head", L.ptr(x1), L.ptr(logits), L.ptr(h["W0t"]), L.ptr(h["b0"]), L.ptr(h["w2"]), L.ptr(h["b2"]),
                   x1.shape[0], h["K"], L.ptr(conf))
        return x1, logits, conf
    hid = _linear(x1, h["W0t"], h["s"], h["h"], relu=True)
    mag = _linear(hid, h["W3t"], None, h["b3"])
    return x1, mag
