"""Per-kernel time of the REFERENCE's own CUDA kernels (oracle/_ref, built unmodified for sm_100a) against this repo's kernels on
the same box, at the shapes of one 8 x 5000-point step (and at 10k / 20k points): the "number to beat" BASELINE.md section 3 asks for.

    python tools/ref_kernel_bench.py [out.md]

Reference timings use the library's own event-timed loops (default stream, initialisation included as the pybind wrappers pay
it); ours are CUDA events around the C-ABI call on the current stream.  Results are checked equal before timing."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from etch_b200 import _lib as L, synth  # noqa: E402
from etch_b200.ext import epn_grouping, pointops_cuda  # noqa: E402
from oracle import ref_kernels as R  # noqa: E402

dev = torch.device("cuda:0")
REPS = 5
rows = []
# bring the clocks up before the first measurement (a fresh box idles at low SM clocks: the first rows were 3x too slow without it)
_w = torch.randn(4096, 4096, device=dev)
_t = torch.cuda.Event(enable_timing=True); _u = torch.cuda.Event(enable_timing=True)
_t.record()
for _ in range(400):
    _w = (_w @ _w).clamp_(-1, 1)
_u.record(); torch.cuda.synchronize()


def ours(fn, reps=REPS):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def p(t):
    return ctypes.c_void_p(t.data_ptr())


for B, N in ((8, 5000), (16, 10000), (8, 20000)):
    pts = synth.sample_real_scans(B, N, 50)
    x = torch.from_numpy(np.ascontiguousarray(pts.transpose(0, 2, 1))).to(dev)
    m = N // 2
    # ---- vgtk FPS N -> N/2
    ref_idx = R.furthest_point_sampling(x, m)
    got = epn_grouping.furthest_point_sampling(x, m)
    assert torch.equal(ref_idx, got)
    temp = torch.empty(B, N, device=dev)
    idx = torch.empty(B, m, dtype=torch.int32, device=dev)
    t_ref = R.lib().ref_time_furthest_point_sampling(p(x), B, N, m, p(temp), p(idx), 2)
    t_our = ours(lambda: epn_grouping.furthest_point_sampling(x, m), 3)
    rows.append(("vgtk FPS %d -> %d" % (N, m), B, t_ref, t_our))
    # ---- ball query, first layer shape (m queries in N points, radius 0.08, 64 samples)
    q = torch.gather(x, 2, ref_idx.long().unsqueeze(1).expand(-1, 3, -1)).contiguous()
    ref_b = R.ball_query(q, x, 0.08, 64)
    assert torch.equal(ref_b, epn_grouping.ball_query(q, x, 0.08, 64))
    bi = torch.empty(B, m, 64, dtype=torch.int32, device=dev)
    t_ref = R.lib().ref_time_ball_query(p(q), p(x), B, m, N, ctypes.c_float(0.08), 64, p(bi), 3)
    t_our = ours(lambda: epn_grouping.ball_query(q, x, 0.08, 64))
    rows.append(("ball query %d in %d, r=0.08, 64" % (m, N), B, t_ref, t_our))
    # ---- pointops FPS level 1 (N -> N/4) and kNN graphs of the PointTransformer hierarchy
    xyz = torch.from_numpy(np.ascontiguousarray(pts.reshape(-1, 3))).to(dev)
    off = torch.tensor([N * (i + 1) for i in range(B)], dtype=torch.int32, device=dev)
    noff = torch.tensor([(N // 4) * (i + 1) for i in range(B)], dtype=torch.int32, device=dev)
    ref_f = R.furthestsampling(xyz, off, noff)
    tmp = torch.full((B * N,), 1e10, device=dev)
    fi = torch.zeros(B * (N // 4), dtype=torch.int32, device=dev)
    pointops_cuda.furthestsampling_cuda(B, N, xyz, off, noff, tmp, fi)
    assert torch.equal(ref_f, fi)
    t_ref = R.lib().ref_time_furthestsampling(B, N, B * N, p(xyz), p(off), p(noff), p(tmp), p(fi), 2)
    t_our = ours(lambda: pointops_cuda.furthestsampling_cuda(B, N, xyz, off, noff, tmp, fi), 3)
    rows.append(("pointops FPS %d -> %d" % (N, N // 4), B, t_ref, t_our))
    for k in (8, 16):
        ri, rd = R.knnquery(k, xyz, xyz, off, off)
        ki = torch.zeros(B * N, k, dtype=torch.int32, device=dev)
        kd = torch.zeros(B * N, k, device=dev)
        fn = L.lib().etch_knn_grid_scratch_bytes
        fn.restype = ctypes.c_longlong
        scratch = torch.empty(int(fn(B * N, B)), dtype=torch.uint8, device=dev)

        def grid():
            L.call("knn_grid", B * N, k, L.ptr(xyz), B * N, L.ptr(xyz), L.ptr(off), L.ptr(off), B, L.ptr(ki), L.ptr(kd), L.ptr(scratch))
        grid()
        assert torch.equal(ri, ki) and torch.equal(rd, kd)
        t_ref = R.lib().ref_time_knnquery(B * N, k, p(xyz), p(xyz), p(off), p(off), p(ki), p(kd), 3)
        t_grid = ours(grid)
        def brute():
            L.call("knn_packed", B * N, k, L.ptr(xyz), L.ptr(xyz), L.ptr(off), L.ptr(off), B, L.ptr(ki), L.ptr(kd))
        t_bf = ours(brute)
        rows.append(("kNN self-graph k=%d, %d points/scan (etch_knn_grid: product path and pointops_cuda binding)" % (k, N), B, t_ref, t_grid))
        rows.append(("kNN self-graph k=%d, %d points/scan (etch_knn_packed: warp-per-query brute force)" % (k, N), B, t_ref, t_bf))

out = ["# Reference CUDA kernels (unmodified, nvcc -O2, sm_100a) vs etch_b200 on the same B200", "",
       "Identical outputs asserted before timing (bit-exact indices and squared distances).  ms per call.", "",
       "| kernel / shape | scans | reference ms | etch_b200 ms | speed-up |", "|---|---|---|---|---|"]
for name, B, a, b in rows:
    out.append("| %s | %d | %.3f | %.3f | %.1fx |" % (name, B, a, b, a / b))
text = "\n".join(out) + "\n"
print(text)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
