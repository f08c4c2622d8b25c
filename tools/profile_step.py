"""One hot-path step between cudaProfilerStart/Stop, for `ncu --profile-from-start off` (see profiles/README.md)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from etch_b200 import synth  # noqa: E402

B = int(os.environ.get("ETCH_PROFILE_BATCH", "8"))
N = int(os.environ.get("ETCH_PROFILE_POINTS", "5000"))
dev = torch.device("cuda:0")
pipe = bench.Pipeline(dev, use_graph=False)
pts = torch.from_numpy(synth.sample_scans(B, N, 50)).to(dev)
for _ in range(2):
    pipe.step(pts)
torch.cuda.synchronize()
torch.cuda.profiler.start()
pipe.step(pts)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step: B=%d N=%d" % (B, N))
