"""Host-side operand layouts of the tensor-core kernels (pure CPU): the TF32 split, the canonical UMMA tiles and the weight
slabs of etch_so3_inter_conv_v3 must reproduce the reference weight exactly (hi + lo == W, bit for bit)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_split_tf32_is_exact_and_tf32_representable():
    from etch_b200.models import tc
    g = torch.Generator().manual_seed(0)
    w = torch.randn(257, 64, generator=g) * torch.logspace(-6, 3, 64)
    hi, lo = tc.split_tf32(w)
    assert torch.equal(hi + lo, w)                                   # the split loses nothing
    assert (hi.numpy().view(np.uint32) & 0x1FFF).max() == 0          # hi has a 10-bit mantissa
    rel = (lo.abs() / w.abs().clamp_min(1e-30)).max().item()
    assert rel <= 2.0 ** -11 + 1e-9                                  # round to nearest: |lo| <= half a TF32 ulp


def test_canonical_tile_layout():
    from etch_b200.models import tc
    w = torch.arange(6 * 8, dtype=torch.float32).view(6, 8)
    t = tc.canonical(w)                                              # [K/4, N, 4]
    assert t.shape == (2, 6, 4)
    for n in range(6):
        for k in range(8):
            assert t[k // 4, n, k % 4] == w[n, k]                    # element (n, k) at (k/4) * (N*4) + n*4 + k%4


def test_inter_conv_v3_slabs_cover_the_weight_exactly():
    from etch_b200.models import encoder
    g = torch.Generator().manual_seed(1)
    for ci, co in ((32, 32), (32, 64), (64, 64)):
        W = torch.randn(co, ci * 24, generator=g)
        slabs = encoder._inter_slabs_v3(W, ci, co)                   # [ci/32*16, 12, 2*co, 4]
        assert slabs.shape == (ci // 32 * 16, 12, 2 * co, 4)
        W3 = W.view(co, ci, 24)
        seen = torch.zeros(ci, 24, dtype=torch.bool)
        for ps in range(ci // 32):
            for cc in range(8):
                for hh in range(2):
                    t = slabs[(ps * 8 + cc) * 2 + hh]                # canonical [K/4][rows][4] -> [rows][48]
                    rows = t.permute(1, 0, 2).reshape(2 * co, 48)
                    rec = rows[:co] + rows[co:]                      # W_hi + W_lo
                    for o in range(4):
                        ch = 32 * ps + 8 * o + cc
                        assert torch.equal(rec[:, o * 12:(o + 1) * 12], W3[:, ch, 12 * hh:12 * hh + 12])
                        seen[ch, 12 * hh:12 * hh + 12] = True
        assert seen.all()                                            # every (channel, kernel point) column appears once


def test_fold_bn_matches_eval_batchnorm():
    """heads._fold_bn: (scale, shift) of an eval-mode BatchNorm1d, with an optional preceding linear bias absorbed."""
    from etch_b200.models import heads
    g = torch.Generator().manual_seed(2)
    c = 37
    bn = torch.nn.BatchNorm1d(c).eval()
    with torch.no_grad():
        bn.weight.copy_(1 + 0.3 * torch.randn(c, generator=g)); bn.bias.copy_(0.2 * torch.randn(c, generator=g))
        bn.running_mean.copy_(0.5 * torch.randn(c, generator=g)); bn.running_var.copy_(0.5 + torch.rand(c, generator=g))
    sd = {"bn." + k: v for k, v in bn.state_dict().items()}
    lin_bias = 0.1 * torch.randn(c, generator=g)
    x = torch.randn(11, c, generator=g)
    dev = dict(dtype=torch.float32, device="cpu")
    s, h = heads._fold_bn(sd, "bn.", dev)
    assert torch.allclose(x * s + h, bn(x), atol=1e-6)
    s, h = heads._fold_bn(sd, "bn.", dev, lin_bias)
    assert torch.allclose(x * s + h, bn(x + lin_bias), atol=1e-6)


def test_pt_block_packs_attention_operands_as_documented():
    """_Block: chan[c] = {P3 row, p3b, s0, h0, so, ho}; Wa_c = (hi, lo) canonical tiles of the BN-folded Linear(c, c/8) per
    64-channel chunk, rows padded to 16 (include/etch_b200.h: etch_pt_attention_tc)."""
    from etch_b200 import synth
    from etch_b200.models import heads
    sd = synth.make_state_dict(1)
    dev = dict(dtype=torch.float32, device="cpu")
    for pre, c in (("magnitude_encoder.enc1.1.", 64), ("confidence_encoder.enc3.1.", 256)):
        blk = heads._Block(sd, pre, dev)
        T, tp = c // 8, max(c // 8, 16)
        assert blk.chan.shape == (c, 8)
        assert torch.equal(blk.chan[:, :3], blk.P3) and torch.equal(blk.chan[:, 3], blk.p3b)
        assert torch.equal(blk.chan[:, 4], blk.s0) and torch.equal(blk.chan[:, 5], blk.h0)
        assert torch.equal(blk.chan[:, 6], blk.so) and torch.equal(blk.chan[:, 7], blk.ho)
        assert blk.Wa_c.shape == (c // 64, 2, 16, tp, 4)
        for k in range(c // 64):
            t = blk.Wa_c[k]                                          # [2][16][tp][4] -> hi + lo as [tp][64]
            rec = (t[0] + t[1]).permute(1, 0, 2).reshape(tp, 64)
            assert torch.equal(rec[:T], blk.Wa[:, 64 * k:64 * k + 64])
            assert (rec[T:] == 0).all()
