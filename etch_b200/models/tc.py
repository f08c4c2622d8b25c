"""Host-side operand preparation for the tcgen05 (tensor-core) kernels: TF32 hi/lo split and the canonical K-major
no-swizzle tile layout documented in csrc/umma.cuh:  element (n, k) of an [N x K] operand lives at float index
(k // 4) * (N * 4) + n * 4 + (k % 4)."""
import numpy as np
import torch


def split_tf32(w):
    """fp32 tensor -> (hi, lo): hi = round-to-nearest(ties away) TF32 (== cvt.rna.tf32.f32), lo = w - hi (exact)."""
    a = np.ascontiguousarray(w.detach().cpu().to(torch.float32).numpy())
    u = a.view(np.uint32).astype(np.uint64)
    hi = ((u + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)
    lo = (a - hi).astype(np.float32)
    return torch.from_numpy(hi.copy()), torch.from_numpy(lo)


def canonical(w_nk):
    """[N,K] (K % 4 == 0) -> [K/4, N, 4] contiguous."""
    n, k = w_nk.shape
    return w_nk.reshape(n, k // 4, 4).permute(1, 0, 2).contiguous()


def tc_operand(w_nk, device):
    """[N,K] fp32 -> [2, K/4, N, 4] (hi tile, lo tile) on `device`."""
    hi, lo = split_tf32(w_nk)
    return torch.stack([canonical(hi), canonical(lo)], 0).contiguous().to(device)


def tc_operand_chunks(w_nk, kchunk, device):
    """[N,K] -> [K/kchunk, 2, kchunk/4, N, 4]: one (hi, lo) tile pair per K-chunk, each contiguous for a bulk copy."""
    n, k = w_nk.shape
    return torch.stack([tc_operand(w_nk[:, c:c + kchunk], "cpu") for c in range(0, k, kchunk)], 0).contiguous().to(device)


def tc_linear_weights(wt_ci_co, device):
    """Wt [ci, co] (the transposed nn.Linear weight used by etch_linear) -> [KC, NC, 2, 16, 64, 4] blocks for etch_linear_tc."""
    w = wt_ci_co.detach().float().cpu().t().contiguous()  # [co, ci]
    co, ci = w.shape
    KC, NC = (ci + 63) // 64, (co + 63) // 64
    wp = torch.zeros(NC * 64, KC * 64)
    wp[:co, :ci] = w
    blocks = [[tc_operand(wp[nc * 64:(nc + 1) * 64, kc * 64:(kc + 1) * 64], "cpu") for nc in range(NC)] for kc in range(KC)]
    return torch.stack([torch.stack(b, 0) for b in blocks], 0).contiguous().to(device)
