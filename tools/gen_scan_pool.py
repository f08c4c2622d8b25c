"""Generates etch_b200/data/scan_pool.npz: area-weighted surface samples of the reference's in-tree 4D-Dress sample
(SURVEY.md section 8d "concrete synthetic inputs"), so that tests and bench run on point clouds with the ball-query density
regimes of a real clothed scan although /root/reference does not travel to the GPU box.

  scan   [32768,3] f32   samples of datafolder/4D-DRESS/data_processed/model/00122_Inner_Take2_00011/*.obj (the clothed scan)
  body   [32768,3] f32   samples of .../smplh/00122_Inner_Take2_00011/mesh_smpl_*.obj (the SMPL(-H) fit under the clothes)
  body_n [32768,3] f16   unit face normals at the body samples (for the "synthetic clothed" displacement along the normal)
Also writes tests/golden/mesh_smpl_00122.npz (vertices float64 as parsed from the OBJ, faces int32) for the mesh tests.
Sampling = face ~ area, then uniform barycentric (u,v) with the sqrt trick (what trimesh.sample.sample_surface does,
src/inference_demo.py:36-39), numpy.random.default_rng(20240917).  Re-run: python tools/gen_scan_pool.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ETCH_REFERENCE", "/root/reference")
BASE = os.path.join(REF, "datafolder", "4D-DRESS", "data_processed")
NAME = "00122_Inner_Take2_00011"
POOL = 32768


def load_obj(path):
    v, f = [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith("v "):
                v.append([float(x) for x in line.split()[1:4]])
            elif line.startswith("f "):
                f.append([int(t.split("/")[0]) - 1 for t in line.split()[1:4]])
    return np.asarray(v, np.float64), np.asarray(f, np.int64)


def sample_surface(v, f, n, rng):
    a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    cr = np.cross(b - a, c - a)
    area = 0.5 * np.linalg.norm(cr, axis=1)
    fi = rng.choice(len(f), size=n, p=area / area.sum())
    r1, r2 = np.sqrt(rng.random(n)), rng.random(n)
    w = np.stack([1 - r1, r1 * (1 - r2), r1 * r2], 1)
    pts = w[:, :1] * a[fi] + w[:, 1:2] * b[fi] + w[:, 2:] * c[fi]
    nrm = cr[fi] / (2 * area[fi, None] + 1e-20)
    return pts, nrm


def main():
    rng = np.random.default_rng(20240917)
    sv, sf = load_obj(os.path.join(BASE, "model", NAME, NAME + ".obj"))
    bv, bf = load_obj(os.path.join(BASE, "smplh", NAME, "mesh_smpl_" + NAME + ".obj"))
    sp, _ = sample_surface(sv, sf, POOL, rng)
    bp, bn = sample_surface(bv, bf, POOL, rng)
    out = os.path.join(ROOT, "etch_b200", "data", "scan_pool.npz")
    np.savez_compressed(out, scan=sp.astype(np.float32), body=bp.astype(np.float32), body_n=bn.astype(np.float16))
    print(out, os.path.getsize(out), "bytes; scan bbox", sp.min(0), sp.max(0))
    # the SMPL(-H) body mesh of the sample as a mesh fixture for the mesh -> point-cloud tests (6890 vertices, 13776 faces)
    mesh = os.path.join(ROOT, "tests", "golden", "mesh_smpl_00122.npz")
    np.savez_compressed(mesh, vertices=bv.astype(np.float64), faces=bf.astype(np.int32))
    print(mesh, os.path.getsize(mesh), "bytes")


if __name__ == "__main__":
    sys.exit(main())
