"""Generates tests/golden/*.npz by running the UNMODIFIED reference code (from /root/reference) in the build container.

  golden_net_b2_n600.npz  -- GT_network_equiv.forward (src/models/models_pointcloud.py:146-221) of the reference, imported
                             through tools/ref_shim.py, with the seeded checkpoint etch_b200.synth.make_state_dict(1)
                             loaded by load_state_dict(strict=True) (which also pins the state-dict key layout), on
                             etch_b200.synth.sample_scans(2, 600, seed=5).
  golden_net_b1_n5000.npz -- the same forward on etch_b200.synth.sample_real_scans(1, 5000, seed=7) (a 5000-point cloud of the
                             in-tree 4D-Dress scan: the size the metric is quoted on); logits reduced to arg-max/top-2/lse.
  golden_lbs.npz          -- external/smplx/smplx/lbs.py::lbs (+ translation and the 21 extra joints) on the seeded
                             synthetic body etch_b200.smpl_model.synthetic_body(0) with seeded random parameters.
  golden_markers.npz      -- src/models/fit_SMPL.py::get_markers on seeded labels / confidences.
/root/reference does not exist on the GPU box; only the fixtures travel.  Re-run: python tools/gen_golden.py
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_shim  # noqa: E402

ref_shim.install()
from etch_b200 import smpl_model, synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
markerset = json.load(open(os.path.join(ROOT, "etch_b200", "data", "superset_smpl.json")))

# ---------------------------------------------------------------- network
from models.models_pointcloud import GT_network_equiv  # noqa: E402  (the reference's)

net = GT_network_equiv(ref_shim.make_option(markerset)).eval()
sd = synth.make_state_dict(1)
net.load_state_dict(sd, strict=True)
pts = torch.from_numpy(synth.sample_scans(2, 600, 5))
with torch.inference_mode():
    res, sel = net(pts, ["confidence", "direction", "magnitude"], "standard_vector")
np.savez_compressed(os.path.join(OUT, "golden_net_b2_n600.npz"), pts=pts.numpy(),
                    **{k: v.numpy().astype(np.float32) for k, v in res.items()}, selected=sel.numpy()[:, :4])
print("net golden written", {k: tuple(v.shape) for k, v in res.items()})

# the same at the size the metric is quoted on (BASELINE configs[1]: 5000 points), on a cloud of the in-tree real scan; the
# 86 logits per point are reduced to their arg-max, top-2 values and log-sum-exp to keep the fixture small
pts5k = torch.from_numpy(synth.sample_real_scans(1, 5000, 7))
with torch.inference_mode():
    res5k, _ = net(pts5k, ["confidence", "direction", "magnitude"], "standard_vector")
lg = res5k["part_labels"][0]
top2 = lg.topk(2, dim=-1)
np.savez_compressed(os.path.join(OUT, "golden_net_b1_n5000.npz"), pts=pts5k.numpy(), labels=top2.indices[:, 0].numpy().astype(np.int16),
                    top2=top2.values.numpy().astype(np.float32), lse=torch.logsumexp(lg, -1).numpy().astype(np.float32),
                    confidences=res5k["confidences"][0, :, 0].numpy().astype(np.float32),
                    magnitude=res5k["magnitude"][0, :, 0].numpy().astype(np.float32),
                    direction=res5k["direction"][0].numpy().astype(np.float32))
print("net golden (1 x 5000, real scan) written; labels used:", int(np.unique(top2.indices[:, 0].numpy()).size))

# ---------------------------------------------------------------- LBS
sys.path.insert(0, os.path.join(ref_shim.REF, "external", "smplx"))
import importlib.util  # noqa: E402

spec_ = importlib.util.spec_from_file_location("ref_lbs", os.path.join(ref_shim.REF, "external", "smplx", "smplx", "lbs.py"),
                                               submodule_search_locations=None)
# lbs.py does `from .utils import ...`: give it a minimal parent package
pkg = types.ModuleType("refsmplx")
pkg.__path__ = [os.path.join(ref_shim.REF, "external", "smplx", "smplx")]
sys.modules["refsmplx"] = pkg
spec_ = importlib.util.spec_from_file_location("refsmplx.lbs", os.path.join(pkg.__path__[0], "lbs.py"))
ref_lbs = importlib.util.module_from_spec(spec_)
sys.modules["refsmplx.lbs"] = ref_lbs
spec_.loader.exec_module(ref_lbs)

body = smpl_model.synthetic_body(0)
g = torch.Generator().manual_seed(7)
Bn = 3
betas = torch.randn(Bn, 10, generator=g)
pose = 0.4 * torch.randn(Bn, 72, generator=g)
pose[0] = 0.0  # exactly the Rodrigues singularity the LM starts from
transl = 0.3 * torch.randn(Bn, 3, generator=g)
tb = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in body.items()}
parents = tb["parents"].clone()
verts, joints, _ = ref_lbs.lbs(betas, pose, tb["v_template"], tb["shapedirs"], tb["posedirs"], tb["J_regressor"], parents,
                               tb["lbs_weights"], pose2rot=True)
extra = verts[:, torch.from_numpy(smpl_model.EXTRA_JOINT_VIDS)]
joints = torch.cat([joints, extra], 1) + transl[:, None]
verts = verts + transl[:, None]
np.savez_compressed(os.path.join(OUT, "golden_lbs.npz"), betas=betas.numpy(), pose=pose.numpy(), transl=transl.numpy(),
                    verts=verts.numpy(), joints=joints.numpy())
print("lbs golden written", tuple(verts.shape), tuple(joints.shape))

# ---------------------------------------------------------------- get_markers
for name in ["sklearn", "sklearn.linear_model", "sklearn.cluster", "tqdm", "utils", "utils.GT_utils", "utils.prior"]:
    if name not in sys.modules:
        sys.modules[name] = types.ModuleType(name)
sys.modules["sklearn.linear_model"].RANSACRegressor = object
sys.modules["sklearn.cluster"].DBSCAN = object
sys.modules["tqdm"].tqdm = lambda x, **k: x
sys.modules["utils.GT_utils"].save_points_with_vector = None
sys.modules["utils.GT_utils"].save_points_with_color = None
sys.modules["utils.prior"].MaxMixturePrior = object
from models.fit_SMPL import get_markers as ref_get_markers  # noqa: E402

g = torch.Generator().manual_seed(11)
B, N = 2, 700
inner = torch.randn(B, N, 3, generator=g)
labels = torch.randint(0, 86, (B, N), generator=g)
labels[0][labels[0] == 17] = 3  # an absent label -> invalid marker
labels[1, :5] = 40
labels[1][5:][labels[1][5:] == 40] = 41
labels[1, 5:][labels[1, 5:] == 60] = 61
labels[1, 100] = 60  # a label with a single point
conf = torch.rand(B, N, 1, generator=g) * 0.5 + 0.5
args = types.SimpleNamespace(markerset=markerset)
mk, valid = ref_get_markers(args, inner, labels, conf)
np.savez_compressed(os.path.join(OUT, "golden_markers.npz"), inner=inner.numpy(), labels=labels.numpy(), conf=conf.numpy(),
                    markers=mk.numpy(), valid=valid.numpy())
print("markers golden written", tuple(mk.shape), int(valid.sum()))

# ---------------------------------------------------------------- save_points_with_vector (eval.py output writer)
# the reference's own function, unmodified, run under THIS interpreter's numpy (its f-string text form of float32 depends on the
# numpy major version; numpy.__version__ is stored next to the bytes)
import importlib.util as _ilu  # noqa: E402

_spec = _ilu.spec_from_file_location("ref_gt_utils", os.path.join(ref_shim.REF, "src", "utils", "GT_utils.py"))
ref_gt_utils = _ilu.module_from_spec(_spec)
_spec.loader.exec_module(ref_gt_utils)
g = np.random.default_rng(21)
hit = (g.normal(size=(300, 3)) * np.array([0.3, 0.9, 0.2])).astype(np.float32)
vec = (g.normal(size=(300, 3)) * 0.03).astype(np.float32)
hit[0] = [0.0, 1.0, -100000.0]
vec[0] = [1e-5, -2.5e-7, 0.0]
hit[1] = [1234567.0, 1e-4, 3.0e16]
ply = os.path.join(OUT, "golden_points_vector.ply")
ref_gt_utils.save_points_with_vector(hit, vec, ply)
np.savez_compressed(os.path.join(OUT, "golden_points_vector_in.npz"), hit=hit, vec=vec, numpy_version=np.__version__)
print("vector PLY golden written", os.path.getsize(ply), "bytes under numpy", np.__version__)
