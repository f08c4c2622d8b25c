"""Measures, on the B200, the tensor-core cost of the M = c_in mapping of the InterSO3Conv neighbour contraction (DESIGN.md section 3,
"Why the neighbour contraction stays on the FP32 pipes"): cycles per (anchor, 8-neighbour K step) of the 3xTF32 product
D[64 x 24] += F[64 x 8] W[24 x 8]^T issued as one N = 48 and one N = 24 tcgen05.mma whose operands come from a ring of distinct
shared-memory tiles, and what that makes per point and per layer against the FP32 loop of the shipping kernel.
    python tools/contract_probe.py [out.md]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from etch_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
out = torch.zeros(148, dtype=torch.int64, device=dev)
steps = 4000
lines = ["# tcgen05 cost of the proposed contraction mapping (measured, 148 CTAs, %d steps each, ring of 16 operand tile sets)" % steps, "",
         "| shape per step | cycles / step (mean over SMs) | per point, nn = 32 (240 steps) | per point, nn = 64 (480 steps) |", "|---|---|---|---|"]
res = {}
for mode, name in ((0, "M = 64: N = 48 + N = 24 (the proposal)"), (1, "M = 128: N = 48 + N = 24"), (2, "M = 64: N = 128 + N = 64 (channel-mixing GEMM, c_out = 64)"),
                   (3, "M = 128: N = 128 + N = 64")):
    for _ in range(2):
        L.call("umma_contract_probe", L.ptr(out), 148, steps, 16, mode)
    torch.cuda.synchronize()
    cyc = out.float().mean().item() / steps
    res[mode] = cyc
    lines.append("| %s | %.1f | %.0f | %.0f |" % (name, cyc, cyc * 240, cyc * 480))
clk = 1.965e9
pts = {"b0.1 (c 32, nn 32, 20000 points)": (20000, 240), "b1.0 (c 32, nn 64, 10000 points)": (10000, 480), "b1.1 (c 64, nn 32, 10000 points)": (10000, 240)}
lines += ["", "Tensor-pipe floor of the contraction alone per launch at B = 8 (148 SMs, 1.965 GHz), against the whole shipping launch:", "",
          "| layer | contraction on tcgen05 (tensor floor only) | shipping kernel, whole launch (r02b ncu) |", "|---|---|---|"]
ship = {"b0.1 (c 32, nn 32, 20000 points)": 3.15, "b1.0 (c 32, nn 64, 10000 points)": 2.95, "b1.1 (c 64, nn 32, 10000 points)": 3.36}
for k, (n, st) in pts.items():
    lines.append("| %s | %.2f ms | %.2f ms |" % (k, n / 148.0 * st * res[0] / clk * 1e3, ship[k]))
lines += ["", "On top of that floor the mapping needs (per point) 0.98 MB of hi/lo feature tiles written by 1920 gather4 TMA instructions, 0.37-0.74 MB of",
          "generated-weight stores, and three 20-anchor TMEM batches whose drain feeds an M = 64 / N = 24 channel-mixing GEMM that re-streams its weights",
          "three times per point (DESIGN.md section 3)."]
text = "\n".join(lines) + "\n"
print(text)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
