"""CPU ORACLE (test infrastructure, NOT product code) -- Python front end.

Loads a URDF the way the reference does (``load_urdf.jl:20-80`` on top of
scikit-robot 0.0.15's ``URDF.load``), hands the joint/link tables to the C
restatement in ``kin_oracle.c`` and exposes the reference's operator names
(``get_transform``, ``get_jacobian``, ``compute_coll_dists_and_grads`` ...) one
configuration at a time, plus batch drivers used by the tests and by
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU-baseline legs may
import this module.  It shares NO code with the product package: the URDF
reader, the id assignment and the box / sphere bookkeeping are restated here
independently so that a flattener bug in the product shows up as a parity
failure.

scikit-robot is not installed here; what it contributes to the path is restated
from its published behaviour (it embeds urdfpy's parser):
  * links and joints are enumerated in XML document order (ids = 1-based
    position, ``load_urdf.jl:22-32``);
  * ``origin`` = 4x4 with R = Rz(yaw) Ry(pitch) Rx(roll) from ``rpy`` and t from
    ``xyz`` (both default 0);
  * ``axis`` defaults to (1,0,0) and is normalised;
  * a ``<collision><geometry><box size=...>`` gives metadata ``extents`` = size
    and ``origin`` = the collision origin (``load_urdf.jl:1-18``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import xml.etree.ElementTree as ET

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FIXED, REVOLUTE, PRISMATIC = 0, 1, 2
GRAD_FD, GRAD_ANALYTIC = 0, 1
SCRATCH_REFERENCE, SCRATCH_CLEAN = 0, 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    """Compile kin_oracle.c (oracle/Makefile)."""
    so = os.path.join(_HERE, "libkin_oracle.so")
    src = os.path.join(_HERE, "kin_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "CC=gcc"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.or_mech_create.restype = C.c_void_p
        L.or_mech_create.argtypes = [C.c_int, C.c_int, _ip, _ip, _ip, _dp, _dp, C.c_int]
        L.or_mech_destroy.argtypes = [C.c_void_p]
        L.or_add_new_link.restype = C.c_int
        L.or_add_new_link.argtypes = [C.c_void_p, C.c_int, _dp]
        L.or_set_joint_angles.argtypes = [C.c_void_p, _ip, C.c_int, _dp]
        L.or_is_relevant.restype = C.c_int
        L.or_is_relevant.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.or_get_transform.argtypes = [C.c_void_p, C.c_int, _dp]
        L.or_rpy.argtypes = [_dp, _dp]
        L.or_get_jacobian.argtypes = [C.c_void_p, C.c_int, _ip, C.c_int, C.c_int, C.c_int, _dp]
        L.or_get_jacobian_inplace.argtypes = L.or_get_jacobian.argtypes
        L.or_sdf_create.restype = C.c_void_p
        L.or_sdf_create.argtypes = [C.c_int, _dp, _dp]
        L.or_sdf_create_prims.restype = C.c_void_p
        L.or_sdf_create_prims.argtypes = [C.c_int, _ip, _dp, _dp]
        L.or_sdf_destroy.argtypes = [C.c_void_p]
        L.or_sdf_eval.restype = C.c_double
        L.or_sdf_eval.argtypes = [C.c_void_p, _dp]
        L.or_sdf_argmin.restype = C.c_int
        L.or_sdf_argmin.argtypes = [C.c_void_p]
        L.or_sdf_gradient.argtypes = [C.c_void_p, _dp, _dp]
        L.or_sdf_gradient_analytic.argtypes = [C.c_void_p, _dp, _dp]
        L.or_compute_coll_dists.argtypes = [C.c_void_p, _ip, _dp, C.c_int, C.c_void_p, _dp, _ip]
        L.or_compute_coll_dists_and_grads.argtypes = [
            C.c_void_p, _ip, C.c_int, _ip, _dp, C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_int, _dp, _dp, _ip]
        L.or_ik_objective.restype = C.c_double
        L.or_ik_objective.argtypes = [C.c_void_p, C.c_int, _ip, C.c_int, _dp, _dp, C.c_int, _dp]
        L.or_pose_constraint.argtypes = [C.c_void_p, C.c_int, _ip, C.c_int, _dp, _dp, C.c_int, _dp, _dp]
        L.or_ineq_const.argtypes = [C.c_void_p, _ip, C.c_int, _ip, _dp, C.c_int, C.c_void_p, _dp, C.c_int,
                                    C.c_double, C.c_int, C.c_int, _dp, _dp]
        L.or_batch_fk.argtypes = [C.c_void_p, _ip, C.c_int, _dp, C.c_long, _ip, C.c_int, _dp, C.c_int]
        L.or_batch_jacobian.argtypes = [C.c_void_p, _ip, C.c_int, _dp, C.c_long, _ip, C.c_int, C.c_int, C.c_int,
                                        _dp, C.c_int]
        L.or_batch_collision.argtypes = [C.c_void_p, _ip, C.c_int, _dp, C.c_long, _ip, _dp, C.c_int, C.c_void_p,
                                         C.c_double, C.c_int, C.c_int, _dp, _dp, _ip, C.c_int]
        L.or_batch_fused.argtypes = [C.c_void_p, _ip, C.c_int, _dp, C.c_long, _ip, C.c_int, C.c_int, C.c_int,
                                     C.c_int, _ip, _dp, C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_int,
                                     _dp, _dp, _dp, _dp, C.c_int]
        L.or_max_threads.restype = C.c_int
        L.or_tf_mul_count.restype = C.c_long
        L.or_tf_mul_count.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _ints(xs):
    return np.ascontiguousarray(np.asarray(xs, dtype=np.int32))


def _dbl(xs):
    return np.ascontiguousarray(np.asarray(xs, dtype=np.float64))


# --------------------------------------------------------------------------
# URDF -> tables (restates scikit-robot/urdfpy parsing + load_urdf.jl wiring)
# --------------------------------------------------------------------------
def rpy_to_matrix(rpy):
    """urdfpy ``rpy_to_matrix``: R = Rz(yaw) Ry(pitch) Rx(roll)."""
    c3, c2, c1 = np.cos(rpy)
    s3, s2, s1 = np.sin(rpy)
    return np.array([
        [c1 * c2, (c1 * s2 * s3) - (c3 * s1), (s1 * s3) + (c1 * c3 * s2)],
        [c2 * s1, (c1 * c3) + (s1 * s2 * s3), (c3 * s1 * s2) - (c1 * s3)],
        [-s2, c2 * s3, c2 * c3]], dtype=np.float64)


def _origin(node):
    T = np.eye(4)
    o = node.find("origin") if node is not None else None
    if o is not None:
        xyz = np.array([float(v) for v in o.get("xyz", "0 0 0").split()])
        rpy = np.array([float(v) for v in o.get("rpy", "0 0 0").split()])
        T[:3, :3] = rpy_to_matrix(rpy)
        T[:3, 3] = xyz
    return T


class RefLink:
    def __init__(self, name, id_, box=None):
        self.name, self.id = name, id_
        self.box = box          # (extents[3], origin 4x4) or None  -- BoxMetaData, mechanism.jl:3-6
        self.prim = None        # extension: (kind, size[3], origin 4x4) of a sphere / cylinder collision primitive
        self.has_meta = False   # any collision geometry at all


class RefJoint:
    def __init__(self, name, id_, plink_id, clink_id, pose, jtype, axis, lower, upper):
        self.name, self.id = name, id_
        self.plink_id, self.clink_id = plink_id, clink_id
        self.pose, self.type, self.axis = pose, jtype, axis
        self.lower, self.upper = lower, upper


class RefMechanism:
    """mechanism.jl:147-181 front end; the state lives in the C object."""

    def __init__(self, links, joints, with_base):
        self.links, self.joints, self.with_base = links, joints, with_base
        self.linkid_map = {l.name: l.id for l in links}
        self.jointid_map = {j.name: j.id for j in joints}
        L = lib()
        nj = len(joints)
        pl = _ints([j.plink_id for j in joints])
        cl = _ints([j.clink_id for j in joints])
        ty = _ints([j.type for j in joints])
        # column-major 4x4 per joint
        poses = _dbl(np.stack([j.pose.T.reshape(-1) for j in joints]) if nj else np.zeros((0, 16)))
        axes = _dbl(np.stack([j.axis for j in joints]) if nj else np.zeros((0, 3)))
        self.h = L.or_mech_create(len(links), nj, _i(pl), _i(cl), _i(ty), _d(poses), _d(axes), int(with_base))

    def __del__(self):
        try:
            lib().or_mech_destroy(self.h)
        except Exception:
            pass

    @property
    def n_dof_extra(self):
        return 3 if self.with_base else 0


def parse_urdf(urdf_path, with_base=False):
    """load_urdf.jl:20-80."""
    root = ET.parse(urdf_path).getroot()
    link_nodes = root.findall("link")
    joint_nodes = root.findall("joint")
    linkid = {n.get("name"): i + 1 for i, n in enumerate(link_nodes)}
    links = []
    for n in link_nodes:
        l = RefLink(n.get("name"), linkid[n.get("name")])
        col = n.find("collision")
        if col is not None:
            l.has_meta = True
            geom = col.find("geometry")
            box = geom.find("box") if geom is not None else None
            if box is not None:
                ext = np.array([float(v) for v in box.get("size").split()])
                l.box = (ext, _origin(col))
            # extension (no reference counterpart): sphere / cylinder collision primitives, kept as (kind, size, origin)
            sph = geom.find("sphere") if geom is not None else None
            cyl = geom.find("cylinder") if geom is not None else None
            if sph is not None:
                l.prim = (1, np.array([float(sph.get("radius")), 0.0, 0.0]), _origin(col))
            if cyl is not None:
                l.prim = (2, np.array([float(cyl.get("radius")), float(cyl.get("length")), 0.0]), _origin(col))
        links.append(l)
    joints = []
    for i, n in enumerate(joint_nodes):
        t = n.get("type")
        ax = n.find("axis")
        axis = np.array([float(v) for v in ax.get("xyz").split()]) if ax is not None else np.array([1.0, 0, 0])
        nrm = np.linalg.norm(axis)
        axis = axis / nrm if nrm > 0 else axis
        lim = n.find("limit")
        lo = float(lim.get("lower", "0")) if lim is not None else 0.0
        hi = float(lim.get("upper", "0")) if lim is not None else 0.0
        if t == "revolute":
            jt = REVOLUTE
        elif t == "continuous":
            jt, lo, hi = REVOLUTE, -np.inf, np.inf
        elif t == "prismatic":
            jt = PRISMATIC
        elif t == "fixed":
            jt, lo, hi = FIXED, -np.inf, np.inf
        else:
            raise ValueError("unknown joint type " + t)     # load_urdf.jl:62
        joints.append(RefJoint(n.get("name"), i + 1, linkid[n.find("parent").get("link")],
                               linkid[n.find("child").get("link")], _origin(n), jt, axis, lo, hi))
    return RefMechanism(links, joints, with_base)


def find_link(m, name):
    return m.links[m.linkid_map[name] - 1]


def find_joint(m, name):
    return m.joints[m.jointid_map[name] - 1]


def add_new_link(m, name, parent, pose):
    """mechanism.jl:233-267. ``pose`` is a 3-vector (position) or a 4x4."""
    pose = np.asarray(pose, dtype=np.float64)
    if pose.shape == (3,):
        T = np.eye(4)
        T[:3, 3] = pose
        pose = T
    colmajor = _dbl(pose.T.reshape(-1))
    id_ = lib().or_add_new_link(m.h, parent.id, _d(colmajor))
    l = RefLink(name, id_)
    m.links.append(l)
    m.linkid_map[name] = id_
    j = RefJoint(name + "_joint", len(m.joints) + 1, parent.id, id_, pose, FIXED, np.zeros(3), -np.inf, np.inf)
    m.joints.append(j)
    m.jointid_map[j.name] = j.id
    return l


def is_relevant(m, joint, link):
    return bool(lib().or_is_relevant(m.h, joint.id, link.id))


def set_joint_angles(m, joints, angles):
    ids = _ints([j.id for j in joints])
    a = _dbl(angles)
    assert a.shape[0] == len(joints) + m.n_dof_extra
    lib().or_set_joint_angles(m.h, _i(ids), len(joints), _d(a))


def get_transform(m, link):
    """4x4 numpy (row/col indexable like the Julia SMatrix)."""
    out = np.zeros(16)
    lib().or_get_transform(m.h, link.id, _d(out))
    return out.reshape(4, 4).T.copy()


def rpy(T):
    out = np.zeros(3)
    lib().or_rpy(_d(_dbl(np.asarray(T).T.reshape(-1))), _d(out))
    return out


def get_jacobian(m, link, joints, with_rot, rpy_jac=False):
    rows, cols = (6 if with_rot else 3), len(joints) + m.n_dof_extra
    out = np.zeros(rows * cols)
    ids = _ints([j.id for j in joints])
    lib().or_get_jacobian(m.h, link.id, _i(ids), len(joints), int(with_rot), int(rpy_jac), _d(out))
    return out.reshape(cols, rows).T.copy()


def get_jacobian_inplace(m, link, joints, with_rot, mat, rpy_jac=False):
    """get_jacobian! -- ``mat`` is a (rows, cols) F-ordered array that is NOT cleared."""
    assert mat.flags["F_CONTIGUOUS"]
    ids = _ints([j.id for j in joints])
    lib().or_get_jacobian_inplace(m.h, link.id, _i(ids), len(joints), int(with_rot), int(rpy_jac), _d(mat))


# --------------------------------------------------------------------------
# SDFs (sdf.jl)
# --------------------------------------------------------------------------
class RefSDF:
    """UnionSDF over boxes with fixed world poses (a single BoxSDF is a union of one)."""

    def __init__(self, poses, widths, kinds=None):
        """kinds (extension beyond the reference, which has boxes only): 0 box (width = extents), 1 sphere
        (width[0] = radius), 2 cylinder along local z (width[0] = radius, width[1] = length)."""
        self.poses = [np.asarray(p, dtype=np.float64) for p in poses]
        self.widths = [np.asarray(w, dtype=np.float64) for w in widths]
        self.kinds = [0] * len(self.poses) if kinds is None else [int(k) for k in kinds]
        P = _dbl(np.stack([p.T.reshape(-1) for p in self.poses]))
        W = _dbl(np.stack(self.widths))
        self.h = lib().or_sdf_create_prims(len(self.poses), _i(_ints(self.kinds)), _d(P), _d(W))

    def __del__(self):
        try:
            lib().or_sdf_destroy(self.h)
        except Exception:
            pass

    def __call__(self, p):
        return lib().or_sdf_eval(self.h, _d(_dbl(p)))

    @property
    def argmin(self):
        return lib().or_sdf_argmin(self.h)

    def gradient(self, p, analytic=False):
        g = np.zeros(3)
        (lib().or_sdf_gradient_analytic if analytic else lib().or_sdf_gradient)(self.h, _d(_dbl(p)), _d(g))
        return g


def BoxSDF(pose, width):
    return RefSDF([pose], [width])


def SphereSDF(pose, radius):
    return RefSDF([pose], [[radius, 0.0, 0.0]], [1])


def CylinderSDF(pose, radius, length):
    return RefSDF([pose], [[radius, length, 0.0]], [2])


def UnionSDF(mech_or_sdfs, primitives=False):
    """sdf.jl:82-97: one box per link that carries box collision metadata, in
    ``mech.links`` order, world pose = get_transform(link) * meta.origin evaluated
    at the obstacle mechanism's CURRENT joint angles / base pose (sdf.jl:14-32).
    ``primitives=True`` (extension) also takes sphere / cylinder collision primitives."""
    if isinstance(mech_or_sdfs, RefMechanism):
        m = mech_or_sdfs
        poses, widths, kinds = [], [], []
        for l in list(m.links):
            if l.box is not None:
                ext, origin = l.box
                poses.append(get_transform(m, l) @ origin)
                widths.append(ext)
                kinds.append(0)
            elif primitives and l.prim is not None:
                kind, size, origin = l.prim
                poses.append(get_transform(m, l) @ origin)
                widths.append(size)
                kinds.append(kind)
        return RefSDF(poses, widths, kinds)
    poses = [p for s in mech_or_sdfs for p in s.poses]
    widths = [w for s in mech_or_sdfs for w in s.widths]
    kinds = [k for s in mech_or_sdfs for k in s.kinds]
    return RefSDF(poses, widths, kinds)


# --------------------------------------------------------------------------
# collision.jl
# --------------------------------------------------------------------------
class SweptSphereCollisionChecker:
    def __init__(self, mech):
        self.mech, self.sphere_links, self.sphere_radii = mech, [], []


def add_coll_links(sscc, coll_link, centers, radii):
    """collision.jl:39-49 with the sphere table given explicitly (the reference
    obtains it from scikit-robot's compute_swept_sphere on a mesh that is not
    available here -- sphere placement parity is unpinned)."""
    for k, (c, r) in enumerate(zip(centers, radii)):
        l = add_new_link(sscc.mech, "sphere_%s_%d" % (coll_link.name, len(sscc.sphere_links)), coll_link, c)
        sscc.sphere_links.append(l)
        sscc.sphere_radii.append(float(r))


def compute_coll_dists(sscc, joints, sdf):
    S = len(sscc.sphere_links)
    ids, rad = _ints([l.id for l in sscc.sphere_links]), _dbl(sscc.sphere_radii)
    vals, am = np.zeros(S), np.zeros(S, dtype=np.int32)
    lib().or_compute_coll_dists(sscc.mech.h, _i(ids), _d(rad), S, sdf.h, _d(vals), _i(am))
    return vals


def compute_coll_dists_and_grads(sscc, joints, sdf, truncation_dist=np.inf, grad_mode=GRAD_FD,
                                 scratch_mode=SCRATCH_REFERENCE, return_argmin=False):
    S = len(sscc.sphere_links)
    n_dof = len(joints) + sscc.mech.n_dof_extra
    ids, rad = _ints([l.id for l in sscc.sphere_links]), _dbl(sscc.sphere_radii)
    jids = _ints([j.id for j in joints])
    vals, grads, am = np.zeros(S), np.zeros(S * n_dof), np.zeros(S, dtype=np.int32)
    lib().or_compute_coll_dists_and_grads(sscc.mech.h, _i(jids), len(joints), _i(ids), _d(rad), S, sdf.h,
                                          float(truncation_dist), grad_mode, scratch_mode, _d(vals), _d(grads), _i(am))
    grads = grads.reshape(S, n_dof).T.copy()       # (n_dof, n_coll) like the reference
    return (vals, grads, am) if return_argmin else (vals, grads)


# --------------------------------------------------------------------------
# callers
# --------------------------------------------------------------------------
def ik_objective(m, link, joints, angles, target, with_rot=True):
    """inverse_kinematics.jl:38-50 -> (f, grad)."""
    jids = _ints([j.id for j in joints])
    n_dof = len(joints) + m.n_dof_extra
    g = np.zeros(n_dof)
    f = lib().or_ik_objective(m.h, link.id, _i(jids), len(joints), _d(_dbl(angles)),
                              _d(_dbl(np.asarray(target).T.reshape(-1))), int(with_rot), _d(g))
    return f, g


def pose_constraint(m, link, joints, q, target, with_rot=True):
    """planning.jl:114-138 for one link -> (val[dim], jac_T (n_dof, dim))."""
    jids = _ints([j.id for j in joints])
    n_dof, dim = len(joints) + m.n_dof_extra, (6 if with_rot else 3)
    val, jt = np.zeros(dim), np.zeros(dim * n_dof)
    lib().or_pose_constraint(m.h, link.id, _i(jids), len(joints), _d(_dbl(q)),
                             _d(_dbl(np.asarray(target).T.reshape(-1))), int(with_rot), _d(val), _d(jt))
    return val, jt.reshape(dim, n_dof).T.copy()


def ineq_const(sscc, joints, sdf, xi, n_wp, margin, grad_mode=GRAD_FD, scratch_mode=SCRATCH_REFERENCE):
    """planning.jl:55-68 -> (val_vec[n_coll*n_wp], blocks[n_wp, n_dof, n_coll])."""
    S = len(sscc.sphere_links)
    n_dof = len(joints) + sscc.mech.n_dof_extra
    ids, rad = _ints([l.id for l in sscc.sphere_links]), _dbl(sscc.sphere_radii)
    jids = _ints([j.id for j in joints])
    xi = _dbl(xi)
    val, blocks = np.zeros(S * n_wp), np.zeros(n_wp * S * n_dof)
    lib().or_ineq_const(sscc.mech.h, _i(jids), len(joints), _i(ids), _d(rad), S, sdf.h, _d(xi), n_wp,
                        float(margin), grad_mode, scratch_mode, _d(val), _d(blocks))
    return val, blocks.reshape(n_wp, S, n_dof).transpose(0, 2, 1).copy()


def objective_matrix(n_wp, weights):
    """planning.jl:7-20: A = kron(A_sub, Diagonal(weights.^2)), A_sub = sum of the 3x3 acceleration blocks
    [1 -2 1; -2 4 -2; 1 -2 1] placed at rows/cols i-1:i+1 for i in 2:n_wp-1 (dense here)."""
    acc_block = np.array([[1.0, -2.0, 1.0], [-2.0, 4.0, -2.0], [1.0, -2.0, 1.0]])
    A_sub = np.zeros((n_wp, n_wp))
    for i in range(2, n_wp):                       # Julia 2:n_wp-1, 1-based
        A_sub[i - 2:i + 1, i - 2:i + 1] += acc_block
    return np.kron(A_sub, np.diag(np.asarray(weights, dtype=np.float64) ** 2))


def objective(A, xi):
    """planning.jl:22-28 -> (val, grad) = (xi' A xi, 2 A xi)."""
    tmp = A @ xi
    return float(xi @ tmp), 2.0 * tmp


def eq_const(m, joints, xi, n_wp, cons_arr):
    """planning.jl:162-176 with the partial constraints of :72-138.  cons_arr: list of
    ("config", idx_wp, q_const) | ("pose", idx_wp, [links], [targets 4x4], [with_rots]); idx_wp is 1-based.
    -> (val_vec[n_cons], jac_mat (n_dof n_wp, n_cons))."""
    n_dof = len(joints) + m.n_dof_extra
    X = np.asarray(xi, dtype=np.float64).reshape(n_wp, n_dof)      # xi_reshaped[:, idx_wp] of the reference
    n_cons = sum(n_dof if c[0] == "config" else sum(6 if w else 3 for w in c[4]) for c in cons_arr)
    val_vec, jac_mat = np.zeros(n_cons), np.zeros((n_dof * n_wp, n_cons))
    i_end = 0
    for c in cons_arr:
        idx_wp = c[1]
        q = X[idx_wp - 1]
        j0 = (idx_wp - 1) * n_dof
        if c[0] == "config":                        # planning.jl:83-88
            i_start, i_end = i_end, i_end + n_dof
            jac_mat[j0:j0 + n_dof, i_start:i_end] = -np.eye(n_dof)
            val_vec[i_start:i_end] = np.asarray(c[2]) - q
        else:                                       # planning.jl:114-138
            for link, target, with_rot in zip(c[2], c[3], c[4]):
                dim = 6 if with_rot else 3
                i_start, i_end = i_end, i_end + dim
                v, jt = pose_constraint(m, link, joints, q, target, with_rot)
                val_vec[i_start:i_end] = v
                jac_mat[j0:j0 + n_dof, i_start:i_end] = jt
    return val_vec, jac_mat


# --------------------------------------------------------------------------
# batch drivers: q is (N, n_dof) C-contiguous
# --------------------------------------------------------------------------
def batch_fk(m, joints, q, links, n_threads=1):
    """-> T[N, n_req, 4, 4] (row, col indexable)."""
    q = _dbl(q)
    N = q.shape[0]
    jids, lids = _ints([j.id for j in joints]), _ints([l.id for l in links])
    out = np.zeros((N, len(links), 16))
    lib().or_batch_fk(m.h, _i(jids), len(joints), _d(q), N, _i(lids), len(links), _d(out), n_threads)
    return out.reshape(N, len(links), 4, 4).transpose(0, 1, 3, 2)


def batch_jacobian(m, joints, q, links, with_rot, rpy_jac=False, n_threads=1):
    """-> J[N, n_req, rows, cols]."""
    q = _dbl(q)
    N = q.shape[0]
    rows, cols = (6 if with_rot else 3), len(joints) + m.n_dof_extra
    jids, lids = _ints([j.id for j in joints]), _ints([l.id for l in links])
    out = np.zeros((N, len(links), cols, rows))
    lib().or_batch_jacobian(m.h, _i(jids), len(joints), _d(q), N, _i(lids), len(links), int(with_rot),
                            int(rpy_jac), _d(out), n_threads)
    return out.transpose(0, 1, 3, 2)


def batch_collision(sscc, joints, sdf, q, truncation_dist=np.inf, grad_mode=GRAD_FD,
                    scratch_mode=SCRATCH_REFERENCE, with_grads=True, n_threads=1):
    """-> vals[N, S], grads[N, S, n_dof] (or None), argmin[N, S] (1-based)."""
    q = _dbl(q)
    N, S = q.shape[0], len(sscc.sphere_links)
    n_dof = len(joints) + sscc.mech.n_dof_extra
    ids, rad = _ints([l.id for l in sscc.sphere_links]), _dbl(sscc.sphere_radii)
    jids = _ints([j.id for j in joints])
    vals, am = np.zeros((N, S)), np.zeros((N, S), dtype=np.int32)
    grads = np.zeros((N, S, n_dof)) if with_grads else None
    lib().or_batch_collision(sscc.mech.h, _i(jids), len(joints), _d(q), N, _i(ids), _d(rad), S, sdf.h,
                             float(truncation_dist), grad_mode, scratch_mode, _d(vals),
                             _d(grads) if with_grads else None, _i(am), n_threads)
    return vals, grads, am


def batch_fused(sscc, joints, sdf, q, links, jac_link, with_rot=True, rpy_jac=False, truncation_dist=np.inf,
                grad_mode=GRAD_FD, scratch_mode=SCRATCH_REFERENCE, n_threads=1, keep_outputs=True):
    """North-star unit of work. Returns (T, J, vals, grads) or None when keep_outputs is False
    (timing runs: results are computed and dropped)."""
    q = _dbl(q)
    m = sscc.mech
    N, S = q.shape[0], len(sscc.sphere_links)
    n_dof = len(joints) + m.n_dof_extra
    rows = 6 if with_rot else 3
    ids, rad = _ints([l.id for l in sscc.sphere_links]), _dbl(sscc.sphere_radii)
    jids, lids = _ints([j.id for j in joints]), _ints([l.id for l in links])
    if keep_outputs:
        T = np.zeros((N, len(links), 16))
        J = np.zeros((N, n_dof, rows))
        vals, grads = np.zeros((N, S)), np.zeros((N, S, n_dof))
        args = (_d(T), _d(J), _d(vals), _d(grads))
    else:
        args = (None, None, None, None)
    lib().or_batch_fused(m.h, _i(jids), len(joints), _d(q), N, _i(lids), len(links),
                         jac_link.id if jac_link is not None else 0, int(with_rot), int(rpy_jac),
                         _i(ids), _d(rad), S, sdf.h if sdf is not None else None, float(truncation_dist),
                         grad_mode, scratch_mode, *args, n_threads)
    if not keep_outputs:
        return None
    return (T.reshape(N, len(links), 4, 4).transpose(0, 1, 3, 2), J.transpose(0, 2, 1), vals, grads)


def max_threads():
    return lib().or_max_threads()
