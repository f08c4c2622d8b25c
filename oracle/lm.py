"""torch-CPU restatement of get_markers + the two-stage Levenberg-Marquardt SMPL marker fit -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference delegates the solve to ``theseus`` (src/models/fit_SMPL.py:2,14,157-255), which is not
vendored, not installed here and not version-pinned (README.md:48 clones HEAD).  The semantics below restate upstream
theseus' ``LevenbergMarquardt`` + ``CholeskyDenseSolver`` + ``NonlinearOptimizer._optimize_loop`` as published:
  * dense linearisation J = d(err)/d(vars) by autodiff (AutoDiffCostFunction, ScaleCostWeight(1.0));
  * delta = solve(J^T J + damping*I, -J^T err) by dense Cholesky; fixed damping (0.01 stage 0 via optimizer_kwargs,
    default 1e-3 stage 1); no adaptive damping, no step rejection;
  * x <- x + step_size*delta for samples not yet converged (Euclidean th.Vector retract, batch_ignore_mask);
  * error metric e = 0.5*||err||^2; a sample is converged when |e_prev-e| < 1e-10 or |e_prev-e|/e_prev < 1e-8, or for
    all samples when mean(|e|) < 1e-10; the loop stops at max_iterations or when all samples converged.
Structure follows src/models/fit_SMPL.py: get_markers :17-62; stage 0 :157-206 (pose 69 + betas[:2] + orient + transl,
30 its, step 0.5); stage 1 :210-255 (all 10 betas, 50 its, step 0.2); final forward :257-269.
"""
import torch

from . import smpl as osmpl


def get_markers(inner_points, part_labels, confidences, n_markers):
    """fit_SMPL.py:17-62 (loop form kept; ties in topk resolved by first occurrence)."""
    B = inner_points.shape[0]
    valid = torch.zeros(B, n_markers, dtype=torch.bool)
    pos = torch.zeros(B, n_markers, 3)
    for b in range(B):
        for label in range(n_markers):
            mask = part_labels[b] == label
            cnt = int(mask.sum())
            if cnt == 0:
                continue
            pts = inner_points[b][mask]
            conf = confidences[b][mask].reshape(-1)
            k = min(cnt, 3)
            _, ind = torch.topk(conf, k, largest=True)
            w = conf[ind] ** 20
            pos[b, label] = (pts[ind] * w.unsqueeze(-1)).sum(0) / w.sum()
            valid[b, label] = True
    return pos, valid


def _residual(x, n_betas, model, marker_vids, target, mask):
    """marker_error_fn_0/1 (fit_SMPL.py:111-152) for ONE sample; x = [pose69, betas n_betas, orient3, transl3]."""
    pose, betas, orient, transl = x[:69], x[69:69 + n_betas], x[69 + n_betas:72 + n_betas], x[72 + n_betas:]
    betas = torch.cat([betas, torch.zeros(10 - n_betas)])
    v, _ = osmpl.smpl_forward(model, orient[None], pose[None], betas[None], transl[None])
    err = (target - v[0, marker_vids]) * mask.unsqueeze(-1)
    return err.reshape(-1)


def lm_stage(x, n_betas, model, marker_vids, target, mask, iters, step, damping, history=None):
    """One theseus LevenbergMarquardt run on a batch. x [B,D]."""
    B, D = x.shape
    f = lambda xi, ti, mi: _residual(xi, n_betas, model, marker_vids, ti, mi)  # noqa: E731
    jac = torch.func.vmap(torch.func.jacrev(f))
    res = torch.func.vmap(f)
    x = x.clone()
    last = 0.5 * (res(x, target, mask) ** 2).sum(1)
    converged = torch.zeros(B, dtype=torch.bool)
    for it in range(iters):
        r = res(x, target, mask)
        J = jac(x, target, mask)
        AtA = J.transpose(1, 2) @ J + damping * torch.eye(D)
        Atb = -(J.transpose(1, 2) @ r.unsqueeze(-1))
        delta = torch.cholesky_solve(Atb, torch.linalg.cholesky(AtA)).squeeze(-1)
        x = torch.where(converged.unsqueeze(-1), x, x + step * delta)
        err = 0.5 * (res(x, target, mask) ** 2).sum(1)
        if history is not None:
            history.append(err.clone())
        if err.abs().mean() < 1e-10:
            converged = torch.ones(B, dtype=torch.bool)
        else:
            ae = (last - err).abs()
            converged = (ae < 1e-10) | (ae / last < 1e-8)
        if converged.all():
            break
        last = err
    return x


def fit(model, marker_vids, target, mask, steps0=30, steps1=50, lr0=0.5, lr1=0.2, history=None):
    """fit_smpl after get_markers (fit_SMPL.py:157-269). target [B,M,3], mask [B,M] float/bool.
    Returns dict(pose [B,69], betas [B,10], orient [B,3], transl [B,3], vertices [B,V,3], joints [B,45,3])."""
    B = target.shape[0]
    mask = mask.float()
    vids = torch.as_tensor(marker_vids, dtype=torch.long)
    x0 = torch.zeros(B, 69 + 2 + 6)
    x0 = lm_stage(x0, 2, model, vids, target, mask, steps0, lr0, 0.01, history)
    x1 = torch.cat([x0[:, :71], torch.zeros(B, 8), x0[:, 71:]], dim=1)
    x1 = lm_stage(x1, 10, model, vids, target, mask, steps1, lr1, 1e-3, history)
    pose, betas, orient, transl = x1[:, :69], x1[:, 69:79], x1[:, 79:82], x1[:, 82:85]
    v, j = osmpl.smpl_forward(model, orient, pose, betas, transl)
    return dict(pose=pose, betas=betas, orient=orient, transl=transl, vertices=v, joints=j)
