"""get_transform / get_jacobian / get_jacobian! (algorithm.jl:1-114) on the B200 backend.

State follows the reference: ``set_joint_angles(m, joints, angles)`` stores the configuration (a
vector, or a batch ``(N, n_dof)``), the getters evaluate it.  With a vector they return what the
reference returns (a ``Transform`` / a ``rows x cols`` matrix); with a batch they return CUDA tensors
with the batch axis first."""
from __future__ import annotations

import numpy as np

from . import lib as _lib
from .device import current_q, device_model, evaluate
from .mechanism import Link, Mechanism
from .transform import Transform


def _as_list(x):
    return (list(x), True) if isinstance(x, (list, tuple)) else ([x], False)


def _to_transform(T34):
    M = np.eye(4)
    M[:3, :] = T34
    return Transform(M)


def get_transform(m: Mechanism, link, layout=None, dtype=None):
    """algorithm.jl:1-37.  ``link`` is a Link or a list of Links.
    single configuration -> Transform (or list of); batch -> tensor (N, 3, 4) or (N, n_links, 3, 4)."""
    links, many = _as_list(link)
    dm = device_model(m)
    Q, ql, N = current_q(m, dtype)
    T = evaluate(dm, Q, ql, N, layout=layout, fk_links=[l.id for l in links])["T"]
    if m._single:
        Th = T[0].double().cpu().numpy()
        res = [_to_transform(Th[i]) for i in range(len(links))]
        return res if many else res[0]
    return T if many else T[:, 0]


def get_jacobian(m: Mechanism, link, joints, with_rot: bool, rpy_jac=False, layout=None, dtype=None):
    """algorithm.jl:108-114 (zero-initialised result).  ``joints`` must be the joints of the last
    ``set_joint_angles`` call (they define the columns).
    single -> ndarray (rows, cols); batch -> tensor (N, rows, cols) or (N, n_links, rows, cols)."""
    links, many = _as_list(link)
    _check_joints(m, joints)
    dm = device_model(m)
    Q, ql, N = current_q(m, dtype)
    J = evaluate(dm, Q, ql, N, layout=layout, jac_links=[l.id for l in links], with_rot=with_rot, rpy_jac=rpy_jac)["J"]
    if m._single:
        Jh = J[0].double().cpu().numpy()
        return [Jh[i] for i in range(len(links))] if many else Jh[0]
    return J if many else J[:, 0]


def get_jacobian_(m: Mechanism, link: Link, joints, with_rot: bool, mat_out, rpy_jac=False):
    """``get_jacobian!`` (algorithm.jl:83-106): writes ONLY the columns of joints that move ``link`` (plus
    the base columns) into the caller's matrix and leaves the others untouched.
    ``mat_out``: ndarray (rows, cols) for a single configuration, or a CUDA tensor whose memory is the
    AoS block ``(N, cols, rows)`` / SoA block ``(cols, rows, N)`` presented as (N, rows, cols)."""
    import torch
    _check_joints(m, joints)
    dm = device_model(m)
    Q, ql, N = current_q(m)
    rows = 6 if with_rot else 3
    if isinstance(mat_out, np.ndarray):
        assert m._single and mat_out.shape == (rows, dm.n_dof)
        store = torch.as_tensor(np.ascontiguousarray(mat_out.T), device="cuda").reshape(1, 1, dm.n_dof, rows)
        evaluate(dm, Q, ql, N, layout=_lib.AOS, jac_links=[link.id], with_rot=with_rot, rpy_jac=rpy_jac,
                 keep_irrelevant=True, J_into=store)
        mat_out[...] = store[0, 0].cpu().numpy().T
        return
    assert mat_out.shape == (N, rows, dm.n_dof)
    if not mat_out.is_cuda:
        raise ValueError("get_jacobian_: a batched mat_out must be a CUDA tensor")
    if mat_out.dtype != Q.dtype:             # the kernel writes elements of q's type into the caller's buffer
        Q, ql, N = current_q(m, mat_out.dtype)
    if mat_out.permute(0, 2, 1).is_contiguous():
        layout, store = _lib.AOS, mat_out.permute(0, 2, 1).unsqueeze(1)
    elif mat_out.permute(2, 1, 0).is_contiguous():
        layout, store = _lib.SOA, mat_out.permute(2, 1, 0).unsqueeze(0)
    else:
        raise ValueError("get_jacobian_: mat_out must be an AoS (N, cols, rows) or SoA (cols, rows, N) block")
    evaluate(dm, Q, ql, N, layout=layout, jac_links=[link.id], with_rot=with_rot, rpy_jac=rpy_jac,
             keep_irrelevant=True, J_into=store)


def _check_joints(m: Mechanism, joints):
    ids = tuple(j.id for j in joints)
    if ids != tuple(m._ctrl):
        raise ValueError("the `joints` argument must be the joints of the last set_joint_angles call "
                         "(they define the configuration columns)")
