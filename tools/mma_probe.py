"""mma.sync TF32 issue rate on this GPU (MAC / clk / SM) for a few occupancies."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from etch_b200 import _lib as L
dev = torch.device("cuda:0")
out = torch.zeros(1024, dtype=torch.int64, device=dev)
iters = 2000
for warps in (4, 8, 16, 32):
    for _ in range(2):
        L.call("mma_sync_rate", L.ptr(out), 148, warps, iters)
    torch.cuda.synchronize()
    cyc = out[:148].float().mean().item()
    macs = warps * iters * 8 * 16 * 8 * 8
    print("mma.sync m16n8k8 tf32: %2d warps/SM  %.0f cycles  -> %.0f MAC/clk/SM (%.1f cyc per mma per SM)" % (warps, cyc, macs / cyc, cyc / (warps * iters * 8)))
for mode, nm in ((0, "FFMA scalar"), (1, "fma.rn.f32x2")):
    for warps in (4, 8, 16, 32):
        for _ in range(2):
            L.call("ffma_rate", L.ptr(out), 148, warps, iters, mode)
        torch.cuda.synchronize()
        cyc = out[:148].float().mean().item()
        fmas = warps * 32 * iters * 32
        print("%s: %2d warps/SM  %.0f cycles -> %.1f FMA/clk/SM" % (nm, warps, cyc, fmas / cyc))
