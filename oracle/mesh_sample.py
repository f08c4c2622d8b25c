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
