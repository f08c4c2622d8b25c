"""Mesh -> point cloud on the GPU: the step in front of the network in both reference callers.

  load_obj(path)                       Wavefront OBJ -> (vertices [V,3] float64, faces [F,3] int32) in file order, like
                                       trimesh.load_mesh(path, process=False, maintain_order=True) (src/inference_demo.py:21)
  preprocess_scan(vertices, device)    src/inference_demo.py:19-34: -> (centred vertices on the device, centre as numpy)
  sample_surface(v, f, count, seed)    trimesh.sample.sample_surface (src/inference_demo.py:36-39; GT_dataloader.py:102):
                                       -> (points [count,3] float64 cuda, face_index [count] int32 cuda); the uniform draws come
                                       from numpy's generator exactly as trimesh draws them, the rest runs in libetch_b200.so
  scan_to_points(path, n, device)      the whole inference_demo.py:19-39,45 sequence -> (points [1,n,3] float32 cuda, centre)
There is no CPU path: a missing library or a CPU `device` raises."""
import numpy as np
import torch

from . import _lib as L


def load_obj(path):
    v, f = [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith("v "):
                p = line.split()
                v.append((float(p[1]), float(p[2]), float(p[3])))
            elif line.startswith("f "):
                p = line.split()[1:]
                idx = [int(t.split("/")[0]) for t in p]
                for k in range(1, len(idx) - 1):      # fan-triangulate polygons (triangle files pass through unchanged)
                    f.append((idx[0], idx[k], idx[k + 1]))
    v = np.asarray(v, np.float64).reshape(-1, 3)
    f = np.asarray(f, np.int64).reshape(-1, 3)
    f = np.where(f < 0, f + len(v), f - 1)            # OBJ indices are 1-based; negative = relative to the end
    return v, f.astype(np.int32)


def _dev(device):
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("etch_b200 has no CPU path: device must be a CUDA device")
    return dev


def preprocess_scan(vertices, device):
    """vertices [V,3] (numpy or tensor) -> (centred [V,3] float64 on `device`, centre [3] numpy float64)."""
    dev = _dev(device)
    v = torch.as_tensor(vertices, dtype=torch.float64).to(dev).contiguous()
    centre = torch.empty(3, dtype=torch.float64, device=dev)
    out = torch.empty_like(v)
    L.call("mesh_center", L.ptr(v), int(v.shape[0]), L.ptr(centre), L.ptr(out))
    return out, centre.cpu().numpy()


def draws(count, seed=None):
    """trimesh's draw order (trimesh/sample.py): random(count), then random((count, 2, 1))."""
    random = np.random.random if seed is None else np.random.default_rng(seed).random
    u_face = random(count)
    u_len = random((count, 2, 1)).reshape(count, 2)
    return u_face, u_len


def sample_surface(vertices, faces, count, seed=None, u=None, want_float32=False):
    """vertices [V,3] float64 cuda, faces [F,3] int32 cuda -> (points [count,3] float64, face_index [count] int32) on the device
    (+ the float32 copy the network consumes when want_float32)."""
    if not vertices.is_cuda:
        raise RuntimeError("etch_b200 has no CPU path: vertices must be a CUDA tensor")
    dev = vertices.device
    v = vertices.to(torch.float64).contiguous()
    f = faces.to(device=dev, dtype=torch.int32).contiguous()
    u_face, u_len = u if u is not None else draws(count, seed)
    uf = torch.from_numpy(np.ascontiguousarray(u_face, np.float64)).to(dev)
    ul = torch.from_numpy(np.ascontiguousarray(u_len, np.float64)).to(dev)
    F = int(f.shape[0])
    scratch = torch.empty(2 * F, dtype=torch.float64, device=dev)
    pts = torch.empty(count, 3, dtype=torch.float64, device=dev)
    pts32 = torch.empty(count, 3, dtype=torch.float32, device=dev) if want_float32 else None
    fidx = torch.empty(count, dtype=torch.int32, device=dev)
    L.call("mesh_sample", L.ptr(v), L.ptr(f), int(v.shape[0]), F, L.ptr(uf), L.ptr(ul), int(count), L.ptr(scratch), L.ptr(pts),
           L.ptr(pts32), L.ptr(fidx))
    return (pts, fidx, pts32) if want_float32 else (pts, fidx)


def scan_to_points(path, num_point, device, seed=None):
    """src/inference_demo.py:19-39 + the tensor conversion of :45 -> (points_tensor [1,num_point,3] float32 cuda, centre numpy)."""
    v, f = load_obj(path)
    cv, centre = preprocess_scan(v, device)
    _, _, p32 = sample_surface(cv, torch.from_numpy(f).to(cv.device), num_point, seed, want_float32=True)
    return p32.unsqueeze(0), centre


# ---- ground-truth tightness vectors of the evaluation dataset (src/data_utils/GT_dataloader.py:104-124) ----
def closest_point(vertices, faces, points):
    """trimesh.proximity.closest_point(mesh, points) on device buffers: -> (closest [n,3] f64, distance [n] f64, face [n] i32)."""
    import ctypes
    dev = vertices.device
    v = vertices.to(torch.float64).contiguous()
    f = faces.to(device=dev, dtype=torch.int32).contiguous()
    p = points.to(device=dev, dtype=torch.float64).contiguous()
    n, F = int(p.shape[0]), int(f.shape[0])
    chunks = int(L.lib().etch_mesh_closest_chunks(F))
    scratch = torch.empty(chunks * n * 12, dtype=torch.uint8, device=dev)
    closest = torch.empty(n, 3, dtype=torch.float64, device=dev)
    dist = torch.empty(n, dtype=torch.float64, device=dev)
    face = torch.empty(n, dtype=torch.int32, device=dev)
    L.call("mesh_closest_point", L.ptr(v), L.ptr(f), int(v.shape[0]), F, L.ptr(p), n, L.ptr(scratch), L.ptr(closest), L.ptr(dist), L.ptr(face))
    return closest, dist, face


def nearest_point(ref, points):
    """scipy cKDTree(ref).query(points, k=1): -> (distance [n] f64, index [n] i32)."""
    dev = ref.device
    r = ref.to(torch.float64).contiguous()
    p = points.to(device=dev, dtype=torch.float64).contiguous()
    n = int(p.shape[0])
    dist = torch.empty(n, dtype=torch.float64, device=dev)
    idx = torch.empty(n, dtype=torch.int32, device=dev)
    L.call("nearest_point", L.ptr(r), int(r.shape[0]), L.ptr(p), n, L.ptr(dist), L.ptr(idx))
    return dist, idx


def gt_vectors(sample_points, info_points, info_vectors, smpl_vertices, smpl_faces, threshold=0.01):
    """GT_dataloader.py:104-124 on the device (all float64): the info vector of the nearest info point when it is closer than
    `threshold`, else sample point - closest point on the SMPL mesh.  -> vectors [n,3] f64."""
    import ctypes
    p = sample_points.to(torch.float64).contiguous()
    dists, indices = nearest_point(info_points.to(p.device), p)
    closest, _, _ = closest_point(smpl_vertices.to(p.device), smpl_faces, p)
    iv = info_vectors.to(device=p.device, dtype=torch.float64).contiguous()
    out = torch.empty_like(p)
    L.call("gt_vectors", L.ptr(p), L.ptr(closest), L.ptr(iv), L.ptr(dists), L.ptr(indices), int(p.shape[0]), ctypes.c_double(threshold), L.ptr(out))
    return out
