"""torch-CPU restatement of SMPL linear blend skinning -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows external/smplx/smplx/lbs.py (the arithmetic pip ``smplx`` shares):
  lbs :153-248, vertices2joints :251-268, blend_shapes :271-292, batch_rodrigues :295-330 (note the +1e-8 at :313),
  transform_mat :333-342, batch_rigid_transform :345-398; SMPL.forward body_models.py:331-424;
  VertexJointSelector vertex_joint_selector.py:29-80 with vertex_ids['smplh'] (vertex_ids.py:24-46).
``model`` is the dict produced by etch_b200.smpl_model (v_template [V,3], shapedirs [V,3,10], posedirs [207,V*3],
J_regressor [24,V], parents [24], lbs_weights [V,24]) as torch tensors.
"""
import torch
import torch.nn.functional as F

EXTRA_JOINT_VIDS = [332, 6260, 2800, 4071, 583, 3216, 3226, 3387, 6617, 6624, 6787,
                    2746, 2319, 2445, 2556, 2673, 6191, 5782, 5905, 6016, 6133]


def batch_rodrigues(rot_vecs):
    n = rot_vecs.shape[0]
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    rot_dir = rot_vecs / angle
    cos = torch.cos(angle).unsqueeze(1)
    sin = torch.sin(angle).unsqueeze(1)
    rx, ry, rz = torch.split(rot_dir, 1, dim=1)
    zeros = torch.zeros((n, 1), dtype=rot_vecs.dtype)
    K = torch.cat([zeros, -rz, ry, rz, zeros, -rx, -ry, rx, zeros], dim=1).view(n, 3, 3)
    ident = torch.eye(3, dtype=rot_vecs.dtype).unsqueeze(0)
    return ident + sin * K + (1 - cos) * torch.bmm(K, K)


def lbs(betas, pose, model):
    """betas [B,10], pose [B,72] -> verts [B,V,3], joints [B,24,3] (no translation)."""
    B = betas.shape[0]
    v_shaped = model["v_template"] + torch.einsum("bl,mkl->bmk", betas, model["shapedirs"])
    J = torch.einsum("bik,ji->bjk", v_shaped, model["J_regressor"])
    rot = batch_rodrigues(pose.reshape(-1, 3)).view(B, -1, 3, 3)
    pose_feature = (rot[:, 1:] - torch.eye(3)).reshape(B, -1)
    v_posed = v_shaped + torch.matmul(pose_feature, model["posedirs"]).view(B, -1, 3)
    parents = model["parents"]
    joints = J.unsqueeze(-1)
    rel = joints.clone()
    rel[:, 1:] = rel[:, 1:] - joints[:, parents[1:]]
    tm = torch.cat([F.pad(rot.reshape(-1, 3, 3), [0, 0, 0, 1]), F.pad(rel.reshape(-1, 3, 1), [0, 0, 0, 1], value=1)],
                   dim=2).reshape(B, -1, 4, 4)
    chain = [tm[:, 0]]
    for i in range(1, parents.shape[0]):
        chain.append(torch.matmul(chain[int(parents[i])], tm[:, i]))
    T = torch.stack(chain, dim=1)
    posed_joints = T[:, :, :3, 3]
    jh = F.pad(joints, [0, 0, 0, 1])
    A = T - F.pad(torch.matmul(T, jh), [3, 0, 0, 0, 0, 0, 0, 0])
    Tv = torch.matmul(model["lbs_weights"].unsqueeze(0).expand(B, -1, -1), A.view(B, -1, 16)).view(B, -1, 4, 4)
    vh = torch.cat([v_posed, torch.ones(B, v_posed.shape[1], 1)], dim=2)
    verts = torch.matmul(Tv, vh.unsqueeze(-1))[:, :, :3, 0]
    return verts, posed_joints


def smpl_forward(model, global_orient, body_pose, betas, transl):
    """SMPL.forward(...) -> (vertices [B,V,3], joints [B,45,3]) incl. translation (body_models.py:386-414)."""
    verts, joints = lbs(betas, torch.cat([global_orient, body_pose], dim=1), model)
    extra = verts[:, torch.tensor(EXTRA_JOINT_VIDS)]
    joints = torch.cat([joints, extra], dim=1)
    return verts + transl.unsqueeze(1), joints + transl.unsqueeze(1)
