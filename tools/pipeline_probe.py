"""Throughput of S whole-step CUDA graphs in flight on S streams (batch i+1's network overlapping batch i's LM fit), for
several SM budgets of the persistent kernels.  Usage: python tools/pipeline_probe.py [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from etch_b200 import _lib, synth  # noqa: E402
from etch_b200.runtime import ScanFitter  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 12
B, N = 8, 5000
dev = torch.device("cuda:0")
pipe = bench.Pipeline(dev, use_graph=False)
pts = [torch.from_numpy(synth.sample_scans(B, N, 50 + i)).to(dev) for i in range(4)]

for budget in (148, 140, 136, 132):
    _lib.lib().etch_set_sm_budget(budget)
    for nstream in (1, 2, 3):
        fitters = [ScanFitter(pipe.net, pipe.args, "neutral", use_graph=True) for _ in range(nstream)]
        streams = [torch.cuda.Stream(device=dev) for _ in range(nstream)]
        for f, s in zip(fitters, streams):
            with torch.cuda.stream(s):
                f(pts[0])
                f(pts[1])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams:
            s.wait_event(e0)
        for i in range(K):
            with torch.cuda.stream(streams[i % nstream]):
                fitters[i % nstream](pts[i % 4])
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print("budget %d streams %d: %.2f ms/step  %.1f scans/s" % (budget, nstream, ms, B * 1000.0 / ms), flush=True)
        del fitters
_lib.lib().etch_set_sm_budget(148)
