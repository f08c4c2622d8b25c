"""SMPL body-model parameters for the marker fit.

``load_smpl_pkl`` reads a real SMPL pickle (the files ``fit_smpl`` expects under datafolder/body_models/smpl/...,
src/models/fit_SMPL.py:92-99); ``synthetic_body`` builds a seeded SMPL-SHAPED model (6890 vertices, 10 betas, 207 pose
blend shapes, the real 24-joint kinematic tree) for tests and benchmarks when no licensed SMPL file is available.

Both return a dict of numpy arrays:
  v_template [V,3] f32, shapedirs [V,3,10] f32, posedirs [207,V*3] f32 (smplx layout: posedirs.reshape(V*3,207).T),
  J_regressor [24,V] f32, parents [24] i64 (parents[0] = -1), lbs_weights [V,24] f32, faces [F,3] i64
"""
import pickle

import numpy as np

SMPL_PARENTS = np.array([-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21], np.int64)
NUM_VERTS = 6890
# vertex_ids['smplh'] picked by pip smplx's VertexJointSelector for SMPL (vertex_joint_selector.py:38-68)
EXTRA_JOINT_VIDS = np.array([332, 6260, 2800, 4071, 583, 3216, 3226, 3387, 6617, 6624, 6787,
                             2746, 2319, 2445, 2556, 2673, 6191, 5782, 5905, 6016, 6133], np.int64)

# approximate SMPL T-pose joint locations (metres), y up
_JOINTS = np.array([
    [0.00, -0.24, 0.03], [0.07, -0.33, 0.02], [-0.07, -0.33, 0.02], [0.00, -0.12, 0.00],
    [0.10, -0.71, 0.02], [-0.10, -0.71, 0.02], [0.00, 0.02, 0.02], [0.09, -1.11, -0.02],
    [-0.09, -1.11, -0.02], [0.00, 0.08, 0.04], [0.12, -1.17, 0.10], [-0.12, -1.17, 0.10],
    [0.00, 0.29, 0.00], [0.08, 0.20, 0.00], [-0.08, 0.20, 0.00], [0.00, 0.38, 0.04],
    [0.18, 0.23, -0.01], [-0.18, 0.23, -0.01], [0.44, 0.22, -0.03], [-0.44, 0.22, -0.03],
    [0.70, 0.22, -0.03], [-0.70, 0.22, -0.03], [0.79, 0.21, -0.04], [-0.79, 0.21, -0.04]], np.float64)
_RADIUS = np.array([0.14, 0.085, 0.085, 0.13, 0.06, 0.06, 0.13, 0.045, 0.045, 0.13, 0.04, 0.04,
                    0.06, 0.07, 0.07, 0.10, 0.055, 0.055, 0.045, 0.045, 0.035, 0.035, 0.03, 0.03], np.float64)


def _bone_segments():
    """One capsule per joint: from the joint to the mean of its children (or a short stub for leaves)."""
    segs = []
    for j in range(24):
        kids = np.where(SMPL_PARENTS == j)[0]
        if len(kids):
            end = _JOINTS[kids].mean(0)
        else:
            end = _JOINTS[j] + (_JOINTS[j] - _JOINTS[SMPL_PARENTS[j]]) * 0.8
        segs.append((_JOINTS[j], end))
    return segs


def sample_capsule_surface(rng, n, inflate=0.0):
    """n points on the union-of-capsules humanoid (area weighted), with outward normals and owning bone."""
    segs = _bone_segments()
    lens = np.array([np.linalg.norm(e - s) for s, e in segs])
    area = (lens + 2 * _RADIUS) * _RADIUS
    bone = rng.choice(24, size=n, p=area / area.sum())
    t = rng.uniform(-0.15, 1.15, size=n)
    ang = rng.uniform(0, 2 * np.pi, size=n)
    pts = np.zeros((n, 3))
    nrm = np.zeros((n, 3))
    for j in range(24):
        m = bone == j
        if not m.any():
            continue
        s, e = segs[j]
        ax = (e - s) / (lens[j] + 1e-12)
        ref = np.array([0.0, 0.0, 1.0]) if abs(ax[2]) < 0.9 else np.array([1.0, 0.0, 0.0])
        u = np.cross(ax, ref)
        u /= np.linalg.norm(u)
        v = np.cross(ax, u)
        tt = np.clip(t[m], 0.0, 1.0)
        over = (t[m] - tt) * lens[j]  # beyond the segment ends -> hemispherical caps
        r = _RADIUS[j] + inflate
        rad = np.sqrt(np.clip(r * r - np.minimum(np.abs(over), r) ** 2, 0.0, None))
        ring = np.cos(ang[m])[:, None] * u + np.sin(ang[m])[:, None] * v
        centre = s + tt[:, None] * (e - s)
        pts[m] = centre + ring * rad[:, None] + ax * np.clip(over, -r, r)[:, None]
        d = pts[m] - centre
        nrm[m] = d / (np.linalg.norm(d, axis=1, keepdims=True) + 1e-12)
    return pts, nrm, bone


def synthetic_body(seed=0):
    rng = np.random.default_rng(seed)
    V = NUM_VERTS
    v, _, _ = sample_capsule_surface(rng, V)
    segs = _bone_segments()
    # skinning weights: softmin over distance to bone segments, top-4
    d = np.zeros((V, 24))
    for j, (s, e) in enumerate(segs):
        ab = e - s
        t = np.clip(((v - s) @ ab) / (ab @ ab + 1e-12), 0, 1)
        d[:, j] = np.linalg.norm(v - (s + t[:, None] * ab), axis=1)
    w = np.exp(-(d / 0.05) ** 2)
    order = np.argsort(-w, axis=1)
    keep = np.zeros_like(w, dtype=bool)
    np.put_along_axis(keep, order[:, :4], True, axis=1)
    w = np.where(keep, w, 0.0) + 1e-12 * keep
    w /= w.sum(1, keepdims=True)
    # joint regressor: the 40 vertices nearest to each joint, inverse-distance weighted, sums to 1
    Jr = np.zeros((24, V))
    for j in range(24):
        dj = np.linalg.norm(v - _JOINTS[j], axis=1)
        nn_ = np.argsort(dj)[:40]
        ww = 1.0 / (dj[nn_] + 0.02)
        Jr[j, nn_] = ww / ww.sum()
    # shape directions: low-frequency fields (height, girth, limb length ...) + smooth noise, ~1-3 cm per unit beta
    sd = np.zeros((V, 3, 10))
    c = v - v.mean(0)
    freq = rng.normal(size=(10, 3)) * 3.0
    phase = rng.uniform(0, 2 * np.pi, size=10)
    for l in range(10):
        axis_scale = rng.normal(size=3) * (0.03 if l < 2 else 0.012)
        sd[:, :, l] = c * axis_scale + 0.004 * np.sin(c @ freq[l] + phase[l])[:, None] * rng.normal(size=3)
    # pose directions: localised by the skinning weight of the driving joint, ~mm scale
    pd = np.zeros((207, V, 3))
    for j in range(1, 24):
        for e in range(9):
            pd[(j - 1) * 9 + e] = 0.01 * w[:, j:j + 1] * rng.normal(size=(1, 3)) * np.sin(c @ rng.normal(size=3) * 4.0)[:, None]
    faces = np.stack([np.arange(V - 2), np.arange(1, V - 1), np.arange(2, V)], 1)
    faces = np.concatenate([faces, faces[: 13776 - len(faces)]], 0)[:13776].astype(np.int64)
    return dict(v_template=v.astype(np.float32), shapedirs=sd.astype(np.float32),
                posedirs=pd.reshape(207, V * 3).astype(np.float32), J_regressor=Jr.astype(np.float32),
                parents=SMPL_PARENTS.copy(), lbs_weights=w.astype(np.float32), faces=faces)


def load_smpl_pkl(path, num_betas=10):
    """Reads an SMPL .pkl the way pip smplx's SMPL.__init__ does (body_models.py: v_template, shapedirs[:, :, :10],
    posedirs reshaped to [207, V*3], J_regressor, kintree_table[0], weights, f)."""
    with open(path, "rb") as fh:
        data = pickle.load(fh, encoding="latin1")

    def arr(x):
        if hasattr(x, "todense"):
            x = x.todense()
        if hasattr(x, "r"):  # chumpy
            x = x.r
        return np.asarray(x)

    v_template = arr(data["v_template"]).astype(np.float32)
    V = v_template.shape[0]
    shapedirs = arr(data["shapedirs"])[:, :, :num_betas].astype(np.float32)
    posedirs = arr(data["posedirs"]).reshape(V * 3, -1).T.astype(np.float32)
    parents = arr(data["kintree_table"])[0].astype(np.int64).copy()
    parents[0] = -1
    return dict(v_template=v_template, shapedirs=shapedirs, posedirs=np.ascontiguousarray(posedirs),
                J_regressor=arr(data["J_regressor"]).astype(np.float32), parents=parents,
                lbs_weights=arr(data["weights"]).astype(np.float32), faces=arr(data["f"]).astype(np.int64))
