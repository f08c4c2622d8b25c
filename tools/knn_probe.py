"""Why is etch_knn_grid slower per query on 8 x 5000 points than on 16 x 10000?  Builds a scratch copy of the library with
-DETCH_KNN_STATS (diagnostic counters), then times the self-graph queries for a few sizes and prints fallback / candidate counts.
    python tools/knn_probe.py"""
import ctypes
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from etch_b200 import synth  # noqa: E402

tmp = tempfile.mkdtemp()
so = os.path.join(tmp, "libknn_stats.so")
subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                       "--expt-relaxed-constexpr", "-DETCH_KNN_STATS", "-I", os.path.join(ROOT, "include"), "-shared", "-o", so,
                       os.path.join(ROOT, "etch_b200", "csrc", "index.cu"), "-lcudart"])
lib = ctypes.CDLL(so)
lib.etch_knn_grid_scratch_bytes.restype = ctypes.c_longlong
dev = torch.device("cuda:0")
w = torch.randn(4096, 4096, device=dev)
for _ in range(200):
    w = (w @ w).clamp_(-1, 1)
torch.cuda.synchronize()
p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
st = ctypes.c_void_p(0)
for kind in ("real", "capsule"):
    for B, N in ((8, 2500), (8, 5000), (8, 7500), (8, 10000), (16, 5000), (16, 10000), (32, 5000)):
        pts = synth.sample_real_scans(B, N, 50) if kind == "real" else synth.sample_scans(B, N, 50)
        xyz = torch.from_numpy(np.ascontiguousarray(pts.reshape(-1, 3))).to(dev)
        off = torch.tensor([N * (i + 1) for i in range(B)], dtype=torch.int32, device=dev)
        for k in (8, 16):
            idx = torch.zeros(B * N, k, dtype=torch.int32, device=dev)
            d2 = torch.zeros(B * N, k, device=dev)
            scratch = torch.empty(int(lib.etch_knn_grid_scratch_bytes(B * N, B)), dtype=torch.uint8, device=dev)
            stats = (ctypes.c_ulonglong * 2)()

            def run():
                rc = lib.etch_knn_grid(B * N, k, p(xyz), B * N, p(xyz), p(off), p(off), B, p(idx), p(d2), p(scratch), st)
                assert rc == 0, rc
            run(); torch.cuda.synchronize(); lib.etch_knn_grid_stats(stats)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                run()
            b.record(); torch.cuda.synchronize()
            lib.etch_knn_grid_stats(stats)
            print("%-7s B=%2d N=%5d k=%2d: %.3f ms/call  %.1f ns/query  fallbacks/call %d  candidates/query %.0f" % (
                kind, B, N, k, a.elapsed_time(b) / 5, a.elapsed_time(b) / 5 * 1e6 / (B * N), stats[0] // 5, stats[1] / 5 / (B * N)), flush=True)
shutil.rmtree(tmp)
