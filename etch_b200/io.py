"""Output writers of the evaluation driver (src/eval.py:125-179, src/utils/GT_utils.py:9-55).

  save_points_with_vector(hit_points, vectors, file_path)   same name / arguments / bytes as utils.GT_utils.save_points_with_vector
                                                            (the reference spends ~35 ms per file in a Python loop; this is one C call)
  results_to_host(fit)                                      ONE pinned device->host copy per batch of everything eval.py stores per scan
                                                            (the reference issues ~20 .detach().cpu().numpy() calls per scan)
  save_tightness_vectors_info / save_output_smpl_info        the two .npz files eval.py writes per scan (same file names, keys, slices)
  v2v_score / append_v2v_score / append_v2v_summary          the V2V number of eval.py and the lines of v2v_score.txt
The colour PLYs and OBJ exports of eval.py go through trimesh / matplotlib (un-vendored third-party code) and stay the caller's."""
import ctypes
import os

import numpy as np
import torch

from . import _lib as L


def _fstring_style():
    """how `f"{numpy.float32}"` prints in this interpreter: 0 = numpy >= 2 (repr of the widened double), 1 = numpy 1.x (short form)"""
    return 1 if f"{np.float32(0.1)}" == "0.1" else 0


def save_points_with_vector(hit_points, vectors, file_path, style=None):
    hit = np.ascontiguousarray(hit_points, np.float32)
    vec = np.ascontiguousarray(vectors, np.float32)
    assert len(hit) == len(vec)                       # GT_utils.py:25
    fn = L.lib().etch_write_points_vector_ply
    fn.restype = ctypes.c_int
    rc = fn(str(file_path).encode(), hit.ctypes.data_as(ctypes.c_void_p), vec.ctypes.data_as(ctypes.c_void_p), int(len(hit)),
            _fstring_style() if style is None else int(style))
    if rc != 0:
        raise OSError(rc, "etch_write_points_vector_ply failed for %s" % file_path)


def format_np_float32(x, style):
    """text of every element: style 0 = repr(float(numpy.float32)), style 1 = str(numpy.float32) (used by the formatter tests)"""
    x = np.ascontiguousarray(x, np.float32).ravel()
    cap = 40 * x.size + 16
    buf = ctypes.create_string_buffer(cap)
    fn = L.lib().etch_format_np_float32
    fn.restype = ctypes.c_longlong
    n = fn(x.ctypes.data_as(ctypes.c_void_p), int(x.size), int(style), buf, ctypes.c_longlong(cap))
    if n < 0:
        raise RuntimeError("formatter buffer too small")
    return buf.raw[:n].decode().split("\n")[:-1]


_KEYS = ("labels", "tightness", "inner", "confidences", "markers", "valid", "vertices", "joints", "params")


def results_to_host(fit, stream=None):
    """dict of CUDA tensors (etch_b200.runtime.ScanFitter output) -> dict of numpy arrays through ONE packed pinned copy."""
    parts = [fit[k].reshape(fit[k].shape[0], -1).to(torch.float32) for k in _KEYS]
    widths = [p.shape[1] for p in parts]
    packed = torch.cat(parts, 1)
    host = torch.empty(packed.shape, dtype=torch.float32, pin_memory=True)
    host.copy_(packed, non_blocking=True)
    torch.cuda.current_stream(packed.device).synchronize()
    out, o = {}, 0
    arr = host.numpy()
    for k, w in zip(_KEYS, widths):
        v = arr[:, o:o + w].reshape(tuple(fit[k].shape))
        out[k] = v.astype(np.int64) if k == "labels" else (v.astype(bool) if k == "valid" else v.copy())
        o += w
    return out


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def save_tightness_vectors_info(output_folder, id_, hitpts, pred_vectors, pred_part_labels, pred_confidences, gt_vectors, gt_labels,
                                gt_confidences):
    """src/eval.py:136-145: <output_folder>/<id>/tightness_vectors_info_<id>.npz with the reference's seven keys (one scan)."""
    os.makedirs(os.path.join(output_folder, f"{id_}"), exist_ok=True)
    path = os.path.join(output_folder, f"{id_}", f"tightness_vectors_info_{id_}.npz")
    np.savez(path, hitpts=_np(hitpts), pred_vectors=_np(pred_vectors), pred_part_labels=_np(pred_part_labels),
             pred_confidences=_np(pred_confidences), gt_vectors=_np(gt_vectors), gt_labels=_np(gt_labels),
             gt_confidences=_np(gt_confidences))
    return path


def save_output_smpl_info(output_folder, id_, output_smpl_info, j):
    """src/eval.py:239-246: <output_folder>/<id>/output_smpl_info_<id>.npz from fit_smpl's 5-list, sample j (21 body + 2 hand joints)."""
    os.makedirs(os.path.join(output_folder, f"{id_}"), exist_ok=True)
    path = os.path.join(output_folder, f"{id_}", f"output_smpl_info_{id_}.npz")
    np.savez(path, body_pose=output_smpl_info[0][j][:21, :], hand_pose=output_smpl_info[0][j][21:23, :], betas=output_smpl_info[1][j],
             global_orient=output_smpl_info[2][j], transl=output_smpl_info[3][j], joints=output_smpl_info[4][j])
    return path


def v2v_score(gt_vertices, pred_vertices):
    """src/eval.py:233-236: mean Euclidean vertex distance, in float64 like the trimesh vertex arrays the reference subtracts."""
    g, q = np.asarray(_np(gt_vertices), np.float64), np.asarray(_np(pred_vertices), np.float64)
    return np.mean(np.linalg.norm(g - q, axis=1))


def append_v2v_score(output_folder, id_, v2v, valid_mask_row):
    """src/eval.py:251-252: one line of <output_folder>/v2v_score.txt, flagged when a marker of the scan was invalid."""
    m = _np(valid_mask_row)
    note = "  attention, the valid mask is not full" if int(m.sum()) != int(m.shape[0]) else ""
    with open(os.path.join(output_folder, "v2v_score.txt"), "a") as f:
        f.write(f"{id_}: {v2v}{note}\n")


def append_v2v_summary(output_folder, total_v2v, sample_num):
    """src/eval.py:258-262: the closing block of v2v_score.txt."""
    with open(os.path.join(output_folder, "v2v_score.txt"), "a") as f:
        f.write("==========\n")
        f.write(f"average v2v: {total_v2v / sample_num}\n")
        f.write(f"total v2v: {total_v2v}\n")
        f.write(f"sample num: {sample_num}\n")
