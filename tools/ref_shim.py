"""Shims that let the UNMODIFIED reference Python (``/root/reference/src/models``, ``vgtk.so3conv``,
``vgtk.pc``) be imported and run on CPU inside the build container.

Golden-fixture tooling only: used by tools/gen_so3_tables.py and tools/gen_golden.py.  It is never imported
by the product (etch_b200/), by tests, or on the GPU box (``/root/reference`` does not exist there).

What is shimmed:
  * native extensions ``epn_grouping`` / ``epn_gathering`` / ``epn_zpconv`` / ``pointops_cuda`` -> backed by the
    bit-faithful C restatements in oracle/etch_oracle.c;
  * missing pure-Python deps: ``trimesh`` (5 calls used by vgtk/functional/rotation.py), ``plyfile`` (ascii PLY
    reader), ``yacs``; empty placeholders for wandb / pytorch3d / theseus / smplx / colour / matplotlib;
  * hard-coded ``.cuda()`` / ``torch.device('cuda:0')`` / ``torch.cuda.{Int,Float}Tensor`` -> CPU.
All PyTorch-level arithmetic (InstanceNorm, einsum, MHSA, PointTransformer, so3_mean ...) stays the reference's own.
"""
import os
import struct
import sys
import types

import numpy as np
import torch

REF = os.environ.get("ETCH_REFERENCE_ROOT", "/root/reference")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)


# ----------------------------------------------------------------------------- trimesh stand-in
class _Mesh:
    def __init__(self, vertices, faces):
        self.vertices = np.asarray(vertices, np.float64)
        self.faces = np.asarray(faces, np.int64)

    def fix_normals(self):
        # sphere12.ply is already consistently outward-wound (checked below); trimesh would be a no-op.
        c = self.vertices[self.faces].mean(1)
        assert (np.einsum("ij,ij->i", self.face_normals, c) > 0).all(), "winding not outward"

    @property
    def face_normals(self):
        t = self.vertices[self.faces]
        n = np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0])
        return n / np.linalg.norm(n, axis=1, keepdims=True)

    @property
    def face_adjacency(self):
        # trimesh.graph.face_adjacency: edges sorted per row, grouped; pairs ordered by the sorted unique edge.
        f = self.faces
        edges = np.stack([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 1).reshape(-1, 2)
        edges.sort(axis=1)
        face_of = np.repeat(np.arange(len(f)), 3)
        order = np.lexsort((edges[:, 1], edges[:, 0]))
        es, fs = edges[order], face_of[order]
        pairs = []
        i = 0
        while i < len(es) - 1:
            if (es[i] == es[i + 1]).all():
                a, b = fs[i], fs[i + 1]
                pairs.append((min(a, b), max(a, b)))
                i += 2
            else:
                i += 1
        return np.asarray(pairs, np.int64)


def _load_binary_ply_mesh(path, **_kw):
    with open(path, "rb") as fh:
        raw = fh.read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    header = raw[:end].decode("ascii").splitlines()
    assert "format binary_little_endian 1.0" in header
    nv = int([h for h in header if h.startswith("element vertex")][0].split()[-1])
    nf = int([h for h in header if h.startswith("element face")][0].split()[-1])
    off = end
    verts = []
    for _ in range(nv):  # float x,y,z + 4 uchar
        verts.append(struct.unpack_from("<fff", raw, off))
        off += 12 + 4
    faces = []
    for _ in range(nf):  # list uchar int vertex_indices; list uchar float texcoord; 4 uchar
        k = raw[off]
        off += 1
        faces.append(struct.unpack_from("<%di" % k, raw, off))
        off += 4 * k
        k2 = raw[off]
        off += 1 + 4 * k2
        off += 4
    return _Mesh(verts, faces)


def _install_trimesh():
    m = types.ModuleType("trimesh")
    m.load_mesh = _load_binary_ply_mesh
    m.load = _load_binary_ply_mesh

    class Trimesh:  # only constructed by fit_SMPL / eval (not used by the golden generator)
        def __init__(self, vertices=None, faces=None, **kw):
            self.vertices, self.faces = vertices, faces

    m.Trimesh = Trimesh
    sys.modules["trimesh"] = m


# ----------------------------------------------------------------------------- plyfile stand-in (ascii)
def _install_plyfile():
    m = types.ModuleType("plyfile")

    class PlyData(dict):
        @staticmethod
        def read(path):
            with open(path, "r") as fh:
                lines = fh.read().splitlines()
            assert lines[1].startswith("format ascii")
            nv = int([l for l in lines if l.startswith("element vertex")][0].split()[-1])
            props = [l.split()[-1] for l in lines if l.startswith("property")]
            body = lines[lines.index("end_header") + 1:][:nv]
            arr = np.array([[float(t) for t in l.split()] for l in body], np.float64)
            d = PlyData()
            d["vertex"] = {p: arr[:, i] for i, p in enumerate(props)}
            return d

    m.PlyData = PlyData
    m.PlyElement = type("PlyElement", (), {})
    sys.modules["plyfile"] = m


# ----------------------------------------------------------------------------- yacs stand-in
def _install_yacs():
    class CfgNode(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        def __setattr__(self, k, v):
            self[k] = v

    yacs = types.ModuleType("yacs")
    cfg = types.ModuleType("yacs.config")
    cfg.CfgNode = CfgNode
    yacs.config = cfg
    sys.modules["yacs"] = yacs
    sys.modules["yacs.config"] = cfg


def _install_placeholders():
    for name in ["wandb", "colour", "theseus", "pytorch3d", "pytorch3d.structures", "smplx", "matplotlib",
                 "matplotlib.pyplot", "potpourri3d"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["pytorch3d.structures"].Meshes = object
    sys.modules["pytorch3d.structures"].Pointclouds = object
    sys.modules["smplx"].SMPL = object


# ----------------------------------------------------------------------------- native extension stand-ins
def _install_native():
    from oracle import index_ops as ops

    g = types.ModuleType("epn_grouping")

    def ball_query(new_xyz, xyz, radius, nsample):
        assert new_xyz.is_contiguous() and xyz.is_contiguous()
        return torch.from_numpy(ops.ball_query_bcn(new_xyz.numpy(), xyz.numpy(), radius, nsample))

    def furthest_point_sampling(xyz, m):
        assert xyz.is_contiguous()
        return torch.from_numpy(ops.fps_bcn(xyz.numpy(), m))

    def _unused(*a, **k):
        raise RuntimeError("not on the hot path")

    g.ball_query, g.furthest_point_sampling = ball_query, furthest_point_sampling
    g.anchor_query = g.initial_anchor_query = _unused
    sys.modules["epn_grouping"] = g

    ga = types.ModuleType("epn_gathering")

    def gather_points_forward(points, idx):
        return torch.from_numpy(ops.gather_bcn(points.contiguous().numpy(), idx.contiguous().numpy()))

    ga.gather_points_forward = gather_points_forward
    ga.gather_points_backward = _unused
    sys.modules["epn_gathering"] = ga
    sys.modules["epn_zpconv"] = types.ModuleType("epn_zpconv")

    p = types.ModuleType("pointops_cuda")

    def knnquery_cuda(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2):
        i, d = ops.knn_packed(nsample, xyz.numpy(), new_xyz.numpy(), offset.numpy(), new_offset.numpy())
        idx.copy_(torch.from_numpy(i))
        dist2.copy_(torch.from_numpy(d))

    def furthestsampling_cuda(b, n_max, xyz, offset, new_offset, tmp, idx):
        idx.copy_(torch.from_numpy(ops.fps_packed(xyz.numpy(), offset.numpy(), new_offset.numpy())))

    p.knnquery_cuda, p.furthestsampling_cuda = knnquery_cuda, furthestsampling_cuda
    sys.modules["pointops_cuda"] = p


# ----------------------------------------------------------------------------- cuda -> cpu
def _patch_torch():
    torch.Tensor.cuda = lambda self, *a, **k: self
    _orig_mod_to = torch.nn.Module.to

    def mod_to(self, *a, **k):
        a = tuple(x for x in a if not (isinstance(x, torch.device) and x.type == "cuda") and not (
            isinstance(x, str) and x.startswith("cuda")))
        if not a and not k:
            return self
        return _orig_mod_to(self, *a, **k)

    torch.nn.Module.to = mod_to
    torch.nn.Module.cuda = lambda self, *a, **k: self
    _orig_t_to = torch.Tensor.to

    def t_to(self, *a, **k):
        a = tuple(torch.device("cpu") if ((isinstance(x, torch.device) and x.type == "cuda") or (
            isinstance(x, str) and x.startswith("cuda"))) else x for x in a)
        return _orig_t_to(self, *a, **k)

    torch.Tensor.to = t_to

    class _IntT:
        def __new__(cls, *a):
            if len(a) == 1 and isinstance(a[0], (list, tuple)):
                return torch.tensor(a[0], dtype=torch.int32)
            return torch.empty(*[int(x) for x in a], dtype=torch.int32)

    class _FloatT:
        def __new__(cls, *a):
            if len(a) == 1 and isinstance(a[0], (list, tuple)):
                return torch.tensor(a[0], dtype=torch.float32)
            return torch.empty(*[int(x) for x in a], dtype=torch.float32)

    torch.cuda.IntTensor = _IntT
    torch.cuda.FloatTensor = _FloatT


_INSTALLED = False


def install():
    global _INSTALLED
    if _INSTALLED:
        return
    _install_trimesh()
    _install_plyfile()
    _install_yacs()
    _install_placeholders()
    _install_native()
    _patch_torch()
    for p in [os.path.join(REF, "src"), os.path.join(REF, "external", "vgtk")]:
        if p not in sys.path:
            sys.path.insert(0, p)
    _INSTALLED = True


def make_option(markerset, output_folder="/tmp/etch_ref_out", radius=0.4, layers=2):
    os.makedirs(output_folder, exist_ok=True)
    return types.SimpleNamespace(output_folder=output_folder, EPN_input_radius=radius, EPN_layer_num=layers,
                                 markerset=markerset, device="cpu", scale_magnitude=10)
