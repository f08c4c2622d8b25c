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


def tc_linear_weights(wt_ci_co, device, smem_budget=150 * 1024):
    """Wt [ci, co] (the transposed nn.Linear weight used by etch_linear) -> (Wc [NG, 2, Kpad/4, NB, 4], NB) for
    etch_linear_tc: groups of NB output columns whose (hi, lo) tile fits `smem_budget` bytes of shared memory."""
    w = wt_ci_co.detach().float().cpu().t().contiguous()  # [co, ci]
    co, ci = w.shape
    kpad = ((ci + 31) // 32) * 32
    nb = min(256, ((co + 15) // 16) * 16, (smem_budget // (8 * kpad)) // 16 * 16)
    if nb < 16:
        raise ValueError("layer too wide for the weight-stationary kernel: ci=%d" % ci)
    ng = (co + nb - 1) // nb
    wp = torch.zeros(ng * nb, kpad)
    wp[:co, :ci] = w
    groups = [tc_operand(wp[g * nb:(g + 1) * nb], "cpu") for g in range(ng)]   # each [2, kpad/4, nb, 4]
    return torch.stack(groups, 0).contiguous().to(device), nb
