"""Per-waypoint constraint evaluations of the reference's trajectory planner (planning.jl) on the B200
backend.  Waypoints are independent on this path (planning.jl:59-67), so (problem, waypoint) pairs are
flattened into the batch axis and the block-diagonal constraint Jacobian is stored as its blocks only
(the reference materialises a dense (n_dof n_wp) x (n_coll n_wp) matrix, planning.jl:50,64-65).
The solvers (NLopt SLSQP / Ipopt / scipy, planning.jl:332-401) are third-party and out of scope."""
from __future__ import annotations

import numpy as np

from . import lib as _lib
from .collision import compute_coll_dists_and_grads
from .mechanism import set_joint_angles


def create_straight_trajectory(q_start, q_goal, n_wp):
    """planning.jl:304-308.  Vectors -> the flat xi (n_dof * n_wp,) of the reference;
    (P, n_dof) tensors -> (P, n_wp, n_dof) on the device of the inputs."""
    import torch
    if isinstance(q_start, torch.Tensor) and q_start.dim() == 2:
        interval = (q_goal - q_start) / (n_wp - 1)
        steps = torch.arange(n_wp, dtype=q_start.dtype, device=q_start.device)
        return q_start[:, None, :] + interval[:, None, :] * steps[None, :, None]
    q_start, q_goal = np.asarray(q_start, dtype=np.float64), np.asarray(q_goal, dtype=np.float64)
    interval = (q_goal - q_start) / (n_wp - 1)
    return np.concatenate([q_start + interval * i for i in range(n_wp)])


class IneqConst:
    """planning.jl:32-68: per-waypoint collision stack, ``val = dists - margin`` with
    ``truncation_dist = margin + 0.05``."""

    def __init__(self, sscc, joints, sdf, n_wp, margin):
        self.sscc, self.joints, self.sdf = sscc, list(joints), sdf
        self.n_wp, self.margin = int(n_wp), float(margin)
        self.n_dof = len(joints) + (3 if sscc.mech.with_base else 0)
        self.n_coll = len(sscc.sphere_links)
        self.n_cons = self.n_coll * self.n_wp

    def __call__(self, xi, grad_mode=_lib.GRAD_FD, scratch_mode=_lib.SCRATCH_REFERENCE):
        """xi: the reference's flat vector (n_dof * n_wp,) -> (val_vec (n_cons,), blocks (n_wp, n_dof, n_coll)) as
        numpy; or a CUDA tensor (P, n_wp, n_dof) of P problems -> tensors (P, n_wp, n_coll), (P, n_wp, n_dof, n_coll)."""
        import torch
        single = not (isinstance(xi, torch.Tensor) and xi.dim() == 3)
        if single:
            X = torch.as_tensor(np.asarray(xi, dtype=np.float64).reshape(1, self.n_wp, self.n_dof), device="cuda")
        else:
            X = xi
        P = X.shape[0]
        set_joint_angles(self.sscc.mech, self.joints, X.reshape(P * self.n_wp, self.n_dof))
        # evaluated in the SoA layout (coalesced stores); the batch-first views below hide the storage order
        vals, grads = compute_coll_dists_and_grads(self.sscc, self.joints, self.sdf, truncation_dist=self.margin + 0.05,
                                                   grad_mode=grad_mode, scratch_mode=scratch_mode, vals_offset=self.margin,
                                                   layout=_lib.SOA)
        vals = vals.reshape(P, self.n_wp, self.n_coll)
        grads = grads.reshape(P, self.n_wp, self.n_dof, self.n_coll)
        if single:
            return vals[0].reshape(-1).cpu().numpy(), grads[0].cpu().numpy()
        return vals, grads

    def dense(self, blocks):
        """The reference's dense jac_mat (n_dof n_wp, n_coll n_wp) from the diagonal blocks of one problem."""
        J = np.zeros((self.n_dof * self.n_wp, self.n_cons))
        for i in range(self.n_wp):
            J[i * self.n_dof:(i + 1) * self.n_dof, i * self.n_coll:(i + 1) * self.n_coll] = blocks[i]
        return J


def nloptize(cons):
    """planning.jl:178-185: NLopt's sign convention (constraint <= 0)."""
    def inner(xi):
        val, jac = cons(xi)
        return -val, -jac
    return inner


def pose_constraint(m, link, joints, target, with_rot=True):
    """One (link, target) pair of ``PoseConstraint`` (planning.jl:114-138) at the configuration(s) of the last
    set_joint_angles: val = [p - p_t; rpy - rpy_t] (dim = 3 | 6), jac_T (n_dof, dim) = transpose of the
    Euler-rate Jacobian.  single -> ndarrays (dim,), (n_dof, dim); batch -> tensors (N, dim), (N, n_dof, dim)."""
    from .inverse_kinematics import _pose_residual
    val, jt = _pose_residual(m, link, joints, target, with_rot, _lib.POSE_CONSTRAINT)
    if m._single:
        return val[0].double().cpu().numpy(), jt[0].double().cpu().numpy()
    return val, jt


class PoseConstraint:
    """planning.jl:90-138."""

    def __init__(self, idx_wp, n_dof, move_links, target_poses, with_rots, mech, joints):
        if not isinstance(move_links, (list, tuple)):
            move_links, target_poses, with_rots = [move_links], [target_poses], [with_rots]
        self.idx_wp, self.n_dof = idx_wp, n_dof
        self.move_links, self.target_poses, self.with_rots = list(move_links), list(target_poses), list(with_rots)
        self.mech, self.joints = mech, list(joints)
        self.n_cons = sum(6 if w else 3 for w in self.with_rots)

    def __call__(self, q):
        """q: (n_dof,) -> (val (n_cons,), jac_rows (n_dof, n_cons)) -- the rows j_start:j_end of the reference's jac_mat."""
        set_joint_angles(self.mech, self.joints, np.asarray(q, dtype=np.float64))
        vals, jacs = [], []
        for link, tp, wr in zip(self.move_links, self.target_poses, self.with_rots):
            v, jt = pose_constraint(self.mech, link, self.joints, tp, wr)
            vals.append(v)
            jacs.append(jt)
        return np.concatenate(vals), np.concatenate(jacs, axis=1)


class ConfigurationConstraint:
    """planning.jl:72-88 (trivial; host side)."""

    def __init__(self, idx_wp, n_dof, q_const):
        self.idx_wp, self.n_dof, self.n_cons = idx_wp, n_dof, n_dof
        self.q_const = np.asarray(q_const, dtype=np.float64)

    def __call__(self, q):
        return self.q_const - np.asarray(q), -np.eye(self.n_dof)


def smoothness_objective(xi, n_wp, weights):
    """``Objective`` of planning.jl:1-28: xi' A xi with A = kron(acceleration stencil, diag(w^2)); off the hot
    path (cheap), batched with torch.  xi (P, n_wp, n_dof) -> (val (P,), grad (P, n_wp, n_dof))."""
    import torch
    w2 = torch.as_tensor(np.asarray(weights, dtype=np.float64) ** 2, device=xi.device, dtype=xi.dtype)
    acc = xi[:, :-2] - 2 * xi[:, 1:-1] + xi[:, 2:]                 # (P, n_wp-2, n_dof)
    val = (acc * acc * w2).sum(dim=(1, 2))
    grad = torch.zeros_like(xi)
    g = 2 * acc * w2
    grad[:, :-2] += g
    grad[:, 1:-1] -= 2 * g
    grad[:, 2:] += g
    return val, grad


def gather_stacked(vals, grads, group=None):
    """Config 5: all-gather the per-rank slabs of stacked constraint values / Jacobian blocks so
    that every rank holds all problems.  vals (P_local, n_wp, n_coll), grads (P_local, n_wp, n_dof, n_coll);
    NCCL on GPUs (gloo in the CPU tests).  Returns the concatenation over ranks along the problem axis."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return vals, grads
    W = dist.get_world_size(group)
    # equal shard sizes (P divisible by the world size) are required, as in config 5 (4096 problems)
    v_all = torch.empty((W * vals.shape[0],) + tuple(vals.shape[1:]), dtype=vals.dtype, device=vals.device)
    g_all = torch.empty((W * grads.shape[0],) + tuple(grads.shape[1:]), dtype=grads.dtype, device=grads.device)
    dist.all_gather(list(v_all.chunk(W)), vals.contiguous(), group=group)
    dist.all_gather(list(g_all.chunk(W)), grads.contiguous(), group=group)
    return v_all, g_all


def shard_range(n_total, rank, world):
    """Contiguous batch shard of SURVEY 8e: rank r gets [r N / G, (r + 1) N / G)."""
    return (n_total * rank) // world, (n_total * (rank + 1)) // world
