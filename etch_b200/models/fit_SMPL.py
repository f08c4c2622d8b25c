"""Drop-in for the reference's ``models.fit_SMPL`` (src/models/fit_SMPL.py:17-269): ``get_markers`` and ``fit_smpl``.

Same signature, same outputs; the work is three kernel launches (segmented top-3 marker extraction, persistent two-stage
Levenberg-Marquardt with an analytic marker-only Jacobian, one full-mesh LBS) instead of theseus + autograd + smplx.
Body-model parameters: the SMPL pickle the reference expects (fit_SMPL.py:92-99), or ``args.smpl_model`` -- a dict in
the layout of ``etch_b200.smpl_model`` -- when the caller supplies one (tests/bench use a synthetic SMPL-shaped body
because the licensed SMPL files cannot be redistributed).
"""
import os

import numpy as np
import torch

from .. import _lib as L
from .. import smpl_model as SM

_BODY_PATHS = {  # fit_SMPL.py:92-99
    "neutral": "datafolder/body_models/smpl/neutral/SMPL_NEUTRAL_10pc_rmchumpy.pkl",
    "female": "datafolder/body_models/smpl/female/SMPL_FEMALE_10pc.pkl",
    "male": "datafolder/body_models/smpl/male/SMPL_MALE_10pc.pkl",
}
_CACHE = {}


class SimpleMesh:
    """Stand-in for trimesh.Trimesh(process=False) when trimesh is not installed: .vertices / .faces / .copy() / .export(obj)
    (what src/inference_demo.py:108-116 and src/eval.py:213-230 do with the returned meshes)."""

    def __init__(self, vertices, faces):
        self.vertices, self.faces = vertices, faces

    def copy(self):
        return SimpleMesh(self.vertices.copy(), self.faces.copy())

    def export(self, path):
        with open(path, "w") as fh:
            for v in self.vertices:
                fh.write("v %.6f %.6f %.6f\n" % tuple(v))
            for f in self.faces:
                fh.write("f %d %d %d\n" % tuple(int(i) + 1 for i in f))


def _make_mesh(v, f):
    try:
        import trimesh
        return trimesh.Trimesh(v, f, process=False, maintain_order=True)
    except ImportError:
        return SimpleMesh(v, f)


class BodyTables:
    """Device-resident SMPL tables: full (for the final mesh) and restricted to the marker vertices (for the solve)."""

    def __init__(self, model, marker_vids, device):
        f = dict(dtype=torch.float32, device=device)
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a))  # noqa: E731
        vt, sd, pd = t(model["v_template"]).double(), t(model["shapedirs"]).double(), t(model["posedirs"]).float()
        Jr, W = t(model["J_regressor"]).double(), t(model["lbs_weights"]).float()
        parents = np.asarray(model["parents"], np.int64).copy()
        self.V = vt.shape[0]
        self.faces = np.asarray(model["faces"])
        vids = torch.as_tensor(np.asarray(marker_vids, np.int64))
        self.M = int(vids.numel())
        self.v_template = vt.to(**f).contiguous()
        self.shapedirs = sd.to(**f).contiguous()            # [V][3][10]
        self.posedirs = pd.to(**f).contiguous()             # [207][V*3]
        self.weights = W.to(**f).contiguous()
        self.Jt = (Jr @ vt).to(**f).contiguous()            # [24][3]
        self.Js = torch.einsum("jv,vcl->jcl", Jr, sd).to(**f).contiguous()  # [24][3][10]
        par = parents.copy()
        par[0] = 0
        self.parents = torch.as_tensor(par.astype(np.int32)).to(device)
        anc = np.zeros(24, np.uint32)
        for k in range(24):
            j = k
            while True:
                anc[k] |= np.uint32(1) << np.uint32(j)
                if parents[j] < 0 or j == 0:
                    break
                j = int(parents[j])
        self.ancmask = torch.as_tensor(anc.astype(np.int64)).to(torch.int32).to(device)  # bit pattern fits 24 bits
        self.extra = torch.as_tensor(SM.EXTRA_JOINT_VIDS.astype(np.int32)).to(device)
        self.Tm = vt[vids].to(**f).contiguous()
        self.Sm = sd[vids].to(**f).contiguous()
        pdv = pd.view(207, self.V, 3)[:, vids, :].reshape(207, self.M * 3)
        self.Pm = pdv.to(**f).contiguous()
        self.Wm = W[vids].to(**f).contiguous()


def body_tables(args, gender, device):
    model = getattr(args, "smpl_model", None)
    vids = tuple(int(v) for v in args.markerset.values())
    if model is None:
        if gender not in _BODY_PATHS:
            raise ValueError(f"Unexpected gender: {gender}")
        path = _BODY_PATHS[gender]
        key = (path, vids, str(device))
        if key not in _CACHE:
            if not os.path.exists(path):
                raise FileNotFoundError("SMPL body model %s not found (and args.smpl_model not given)" % path)
            _CACHE[key] = BodyTables(SM.load_smpl_pkl(path), vids, device)
        return _CACHE[key]
    if gender not in _BODY_PATHS:
        raise ValueError(f"Unexpected gender: {gender}")
    # caller-supplied model: the entry keeps a reference to the dict it was built from, so a recycled id() can never hand
    # another model the wrong tables
    key = (id(model), vids, str(device))
    ent = _CACHE.get(key)
    if ent is None or ent[0] is not model:
        ent = _CACHE[key] = (model, BodyTables(model, vids, device))
    return ent[1]


def get_markers(args, inner_points, part_labels, confidences):
    """inner_points [B,K,3], part_labels [B,K] int64, confidences [B,K,1] -> markers [B,M,3], valid [B,M] bool."""
    B, N, _ = inner_points.shape
    M = len(args.markerset)
    dev = inner_points.device
    markers = torch.empty(B, M, 3, dtype=torch.float32, device=dev)
    valid = torch.empty(B, M, dtype=torch.uint8, device=dev)
    L.call("markers_top3", L.ptr(inner_points.float().contiguous()), L.ptr(part_labels.to(torch.int64).contiguous()),
           L.ptr(confidences.float().reshape(B, N).contiguous()), B, N, M, L.ptr(markers), L.ptr(valid))
    return markers, valid.bool()


def lm_fit(tables, markers, valid, steps_stage0=30, steps_stage1=50, lr_stage0=0.5, lr_stage1=0.2, damping0=0.01,
           damping1=1e-3):
    """Two-stage LM + final SMPL forward on device. Returns dict of CUDA tensors."""
    B = markers.shape[0]
    dev = markers.device
    params = torch.empty(B, 85, dtype=torch.float32, device=dev)
    iters = torch.empty(B, 2, dtype=torch.int32, device=dev)
    errs = torch.empty(B, 2, dtype=torch.float32, device=dev)
    T = tables
    L.call("lm_fit", L.ptr(markers.contiguous()), L.ptr(valid.to(torch.uint8).contiguous()), L.ptr(T.Tm), L.ptr(T.Sm), L.ptr(T.Pm),
           L.ptr(T.Wm), L.ptr(T.Jt), L.ptr(T.Js), L.ptr(T.parents), L.ptr(T.ancmask), B, T.M, int(steps_stage0), int(steps_stage1),
           L.f32(lr_stage0), L.f32(lr_stage1), L.f32(damping0), L.f32(damping1), L.ptr(params), L.ptr(iters), L.ptr(errs))
    verts = torch.empty(B, T.V, 3, dtype=torch.float32, device=dev)
    joints = torch.empty(B, 45, 3, dtype=torch.float32, device=dev)
    L.call("lbs_forward", L.ptr(params), L.ptr(T.v_template), L.ptr(T.shapedirs), L.ptr(T.posedirs), L.ptr(T.weights),
           L.ptr(T.Jt), L.ptr(T.Js), L.ptr(T.parents), L.ptr(T.extra), B, T.V, L.ptr(verts), L.ptr(joints))
    return dict(params=params, orient=params[:, 0:3], pose=params[:, 3:72], betas=params[:, 72:82], transl=params[:, 82:85],
                vertices=verts, joints=joints, iters=iters, errs=errs)


def smpl_forward(tables, params):
    """SMPL forward for packed params [B,85] = orient|pose|betas|transl -> (vertices [B,V,3], joints [B,45,3])."""
    B = params.shape[0]
    T = tables
    verts = torch.empty(B, T.V, 3, dtype=torch.float32, device=params.device)
    joints = torch.empty(B, 45, 3, dtype=torch.float32, device=params.device)
    L.call("lbs_forward", L.ptr(params.contiguous()), L.ptr(T.v_template), L.ptr(T.shapedirs), L.ptr(T.posedirs), L.ptr(T.weights),
           L.ptr(T.Jt), L.ptr(T.Js), L.ptr(T.parents), L.ptr(T.extra), B, T.V, L.ptr(verts), L.ptr(joints))
    return verts, joints


def fit_smpl(args, inner_points, part_labels, confidences, gender, steps_stage0=30, steps_stage1=50, lr_stage0=5e-1,
             lr_stage1=2e-1):
    """Fit SMPL to the predicted markers (same contract as src/models/fit_SMPL.py:68-269).

    Returns (final_mesh_list, pred_markers_position [B,M,3], valid_mask [B,M] bool,
             [pose (B,23,3), shape (B,10), global_orient (B,3), transl (B,3), joints (B,45,3)] as numpy)."""
    if gender not in _BODY_PATHS:
        raise ValueError(f"Unexpected gender: {gender}")
    tables = body_tables(args, gender, inner_points.device)
    markers, valid = get_markers(args, inner_points, part_labels, confidences)
    out = lm_fit(tables, markers, valid, steps_stage0, steps_stage1, lr_stage0, lr_stage1)
    B = markers.shape[0]
    verts = out["vertices"].cpu().numpy()
    meshes = [_make_mesh(verts[b], tables.faces) for b in range(B)]
    info = [out["pose"].cpu().numpy().reshape(B, 23, 3), out["betas"].cpu().numpy(), out["orient"].cpu().numpy(),
            out["transl"].cpu().numpy(), out["joints"].cpu().numpy()]
    return meshes, markers, valid, info
