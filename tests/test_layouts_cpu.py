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
