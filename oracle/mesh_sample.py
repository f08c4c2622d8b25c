"""numpy restatement of the reference's mesh -> point-cloud step -- TEST INFRASTRUCTURE (see oracle/__init__.py).

  preprocess_scan   src/inference_demo.py:19-34   (bbox centre of the vertices, vertices -= centre)
  sample_surface    trimesh.sample.sample_surface as called at src/inference_demo.py:36-39 and
                    src/data_utils/GT_dataloader.py:102 (seed = self.seed + 15)

PARITY UNPINNED for sample_surface: trimesh is a third-party dependency of the reference (environment.yml:21), not vendored,
not version-pinned and not installed in the build container, so no golden vector could be generated from it.  The algorithm
below restates trimesh 4.x as published (trimesh/sample.py::sample_surface, trimesh/triangles.py::area / cross):
  area    = sqrt((cross(v1 - v0, v2 - v1) ** 2).sum(axis=1)) / 2          (np.diff of each triangle, then np.cross)
  cum     = np.cumsum(area);  pick = random(count) * cum[-1];  face = np.searchsorted(cum, pick)
  lengths = random((count, 2, 1));  lengths[lengths.sum(axis=1) > 1] -= 1;  lengths = abs(lengths)
  sample  = ((v1 - v0, v2 - v0) * lengths).sum(axis=1) + v0
with random = np.random.default_rng(seed).random (or np.random.random when seed is None).  preprocess_scan is the reference's
own code and is pinned by construction.
"""
import numpy as np


def preprocess_scan(vertices):
    vertices = np.asarray(vertices, np.float64)
    centre = (vertices.min(axis=0) + vertices.max(axis=0)) / 2.0
    return vertices - centre, centre


def draws(count, seed=None):
    """the uniform draws in trimesh's order: `count` for the face pick, then (count, 2, 1) for the barycentric lengths"""
    random = np.random.random if seed is None else np.random.default_rng(seed).random
    u_face = random(count)
    u_len = random((count, 2, 1))
    return u_face, u_len.reshape(count, 2)


def face_areas(vertices, faces):
    tri = np.asarray(vertices, np.float64)[np.asarray(faces)]
    vec = np.diff(tri, axis=1)
    crosses = np.cross(vec[:, 0], vec[:, 1])
    return np.sqrt((crosses ** 2).sum(axis=1)) / 2.0


def sample_surface(vertices, faces, u_face, u_len):
    vertices = np.asarray(vertices, np.float64)
    faces = np.asarray(faces)
    cum = np.cumsum(face_areas(vertices, faces))
    face_index = np.searchsorted(cum, u_face * cum[-1])
    origins = vertices[faces[:, 0]]
    vectors = vertices[faces[:, 1:]].copy()
    vectors -= np.tile(origins, (1, 2)).reshape((-1, 2, 3))
    origins, vectors = origins[face_index], vectors[face_index]
    lengths = np.array(u_len, np.float64).reshape(-1, 2, 1)
    test = lengths.sum(axis=1).reshape(-1) > 1.0
    lengths[test] -= 1.0
    lengths = np.abs(lengths)
    return (vectors * lengths).sum(axis=1) + origins, face_index


# ---------------------------------------------------------------------------------------------------------------------------------
# "VECTORS" block of the evaluation dataset (src/data_utils/GT_dataloader.py:104-124).  PARITY UNPINNED for the two third-party
# calls (scipy cKDTree.query and trimesh.proximity.closest_point: the latter is not installed and walks an r-tree before calling
# trimesh.triangles.closest_point); both are exact searches, restated here as brute force with the same point-triangle projection
# (Ericson, Real-Time Collision Detection 5.1.5) evaluated in float64 with unfused ((x + y) + z) dot products.
def _dot(a, b):
    return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]


def closest_on_triangles(p, a, b, c):
    """p [3]; a, b, c [F,3] -> closest point on every triangle [F,3] (region tests in Ericson's order)."""
    ab, ac, ap = b - a, c - a, p - a
    d1, d2 = _dot(ab, ap), _dot(ac, ap)
    bp = p - b
    d3, d4 = _dot(ab, bp), _dot(ac, bp)
    cp = p - c
    d5, d6 = _dot(ab, cp), _dot(ac, cp)
    vc = d1 * d4 - d3 * d2
    vb = d5 * d2 - d1 * d6
    va = d3 * d6 - d5 * d4
    e43, e56 = d4 - d3, d5 - d6
    with np.errstate(divide="ignore", invalid="ignore"):
        out = a + ab * (vb * (1.0 / ((va + vb) + vc)))[:, None]
        out = out + ac * (vc * (1.0 / ((va + vb) + vc)))[:, None]
        conds = [((d1 <= 0) & (d2 <= 0), a),
                 ((d3 >= 0) & (d4 <= d3), b),
                 ((vc <= 0) & (d1 >= 0) & (d3 <= 0), a + ab * (d1 / (d1 - d3))[:, None]),
                 ((d6 >= 0) & (d5 <= d6), c),
                 ((vb <= 0) & (d2 >= 0) & (d6 <= 0), a + ac * (d2 / (d2 - d6))[:, None]),
                 ((va <= 0) & (e43 >= 0) & (e56 >= 0), b + (c - b) * (e43 / (e43 + e56))[:, None])]
    done = np.zeros(len(a), bool)
    for m, val in conds:
        take = m & ~done
        out[take] = val[take]
        done |= m
    return out


def closest_point(vertices, faces, points):
    """-> (closest [n,3], distance [n], face index [n]); ties -> lowest face index."""
    v = np.asarray(vertices, np.float64)
    f = np.asarray(faces)
    a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    n = len(points)
    out, dist, face = np.zeros((n, 3)), np.zeros(n), np.zeros(n, np.int64)
    for i, p in enumerate(np.asarray(points, np.float64)):
        q = closest_on_triangles(p, a, b, c)
        d = p - q
        dd = _dot(d, d)
        k = int(np.argmin(dd))
        out[i], dist[i], face[i] = q[k], np.sqrt(dd[k]), k
    return out, dist, face


def nearest_point(ref, points):
    """cKDTree(ref).query(points, k=1) as brute force: -> (distance [n], index [n]); ties -> lowest index."""
    ref = np.asarray(ref, np.float64)
    dist, idx = np.zeros(len(points)), np.zeros(len(points), np.int64)
    for i, p in enumerate(np.asarray(points, np.float64)):
        d = p - ref
        dd = _dot(d, d)
        k = int(np.argmin(dd))
        dist[i], idx[i] = np.sqrt(dd[k]), k
    return dist, idx


def gt_vectors(sample_points, info_points, info_vectors, smpl_vertices, smpl_faces, threshold=0.01):
    """GT_dataloader.py:104-124."""
    dists, indices = nearest_point(info_points, sample_points)
    closest, _, _ = closest_point(smpl_vertices, smpl_faces, sample_points)
    vectors = np.zeros((len(sample_points), 3), np.float64)
    cond = dists < threshold
    vectors[cond] = np.asarray(info_vectors, np.float64)[indices[cond]]
    vectors[~cond] = np.asarray(sample_points, np.float64)[~cond] - closest[~cond]
    return vectors
