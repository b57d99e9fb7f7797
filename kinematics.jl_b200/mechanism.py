"""Host-side mirror of mechanism.jl: the link/joint tree and its (mutable) joint state.  The tree
is metadata; the numbers are produced by libkin_b200 from the flattened tables of device.py."""
from __future__ import annotations

import numpy as np

from . import lib as _lib
from .transform import Transform

FIXED, REVOLUTE, PRISMATIC = _lib.FIXED, _lib.REVOLUTE, _lib.PRISMATIC


class BoxMetaData:                         # mechanism.jl:3-6
    def __init__(self, extents, origin: Transform):
        self.extents, self.origin = np.asarray(extents, dtype=np.float64), origin


class SphereMetaData:                      # mechanism.jl:8-11
    def __init__(self, radius, origin: Transform):
        self.radius, self.origin = float(radius), origin


class CylinderMetaData:                    # extension: URDF <cylinder> collision primitive (the reference skips it,
    def __init__(self, radius, length, origin: Transform):     # load_urdf.jl:13-15); only used to fit swept spheres
        self.radius, self.length, self.origin = float(radius), float(length), origin


class MeshMetaData:                        # mechanism.jl:13-16
    def __init__(self, file_path, origin: Transform):
        self.file_path, self.origin = file_path, origin


class Link:                                # mechanism.jl:35-49
    def __init__(self, name, link_type="URDF", geometric_meta_data=None):
        self.link_type, self.name = link_type, name
        self.id = self.pjoint_id = self.plink_id = -1
        self.cjoint_ids, self.clink_ids = [], []
        self.geometric_meta_data = geometric_meta_data
        self.data = {}

    def __repr__(self):
        return "Link(%s, id=%d)" % (self.name, self.id)


User = "User"                              # mechanism.jl:33: Link(User, "name")


class Joint:                               # mechanism.jl:74-88 with the JointType folded in (:51-72)
    def __init__(self, name, id_, plink_id, clink_id, pose: Transform, jtype, axis=(0.0, 0.0, 0.0),
                 lower_limit=-np.inf, upper_limit=np.inf):
        self.name, self.id, self.plink_id, self.clink_id = name, id_, plink_id, clink_id
        self.pose, self.type = pose, jtype
        self.axis = np.asarray(axis, dtype=np.float64)
        self.lower_limit, self.upper_limit = float(lower_limit), float(upper_limit)

    def __repr__(self):
        return "Joint(%s, id=%d)" % (self.name, self.id)


def lower_limit(j: Joint):
    return j.lower_limit


def upper_limit(j: Joint):
    return j.upper_limit


class Mechanism:                           # mechanism.jl:147-181
    def __init__(self, links, joints, linkid_map, jointid_map, with_base, robot_type="basic"):
        self.robot_type = robot_type
        self.links, self.joints = links, joints
        self.linkid_map, self.jointid_map = linkid_map, jointid_map
        self.angles = np.zeros(len(joints))
        self.base_pose = np.zeros(3)
        self.with_base = bool(with_base)
        # device-side bookkeeping (device.py)
        self._structure_version = 0        # bumped by add_new_link
        self._state_version = 0            # bumped by every set_joint_angle(s) / set_base_pose
        self._ctrl = ()                    # control joints of the current configuration (batch)
        self._Q = None                     # torch (N, n_dof) on the device, or None => from angles/base_pose
        self._single = True
        self._models = {}                  # cache of device models

    @property
    def rptable(self):                     # create_rptable, mechanism.jl:117-139 (joint id x link id)
        table = np.zeros((len(self.joints), len(self.links)), dtype=bool)
        for j in self.joints:
            stack = [j.clink_id]
            while stack:
                l = stack.pop()
                table[j.id - 1, l - 1] = True
                stack.extend(self.links[l - 1].clink_ids)
        return table


def parent_link(m: Mechanism, x):          # mechanism.jl:183,186
    return m.links[x.plink_id - 1]


def child_link(m: Mechanism, joint: Joint):
    return m.links[joint.clink_id - 1]


def child_links(m: Mechanism, link: Link):
    return [m.links[i - 1] for i in link.clink_ids]


def parent_joint(m: Mechanism, link: Link):
    return m.joints[link.pjoint_id - 1]


def child_joints(m: Mechanism, link: Link):
    return [m.joints[i - 1] for i in link.cjoint_ids]


def find_joint(m: Mechanism, name):
    return m.joints[m.jointid_map[name] - 1]


def find_link(m: Mechanism, name):
    return m.links[m.linkid_map[name] - 1]


def isroot(link: Link):
    return link.plink_id == -1


def isleaf(link: Link):
    return len(link.clink_ids) == 0


def joint_angle(m: Mechanism, joint: Joint):
    return float(m.angles[joint.id - 1])


def is_relevant(m: Mechanism, joint: Joint, link: Link):   # mechanism.jl:277
    l = link
    while True:
        if l.pjoint_id == joint.id:
            return True
        if l.plink_id == -1:
            return False
        l = m.links[l.plink_id - 1]


def _touch(m: Mechanism):
    m._state_version += 1


def set_joint_angle(m: Mechanism, joint, angle):           # mechanism.jl:199-200
    jid = joint if isinstance(joint, int) else joint.id
    m.angles[jid - 1] = float(angle)
    m._Q = None
    m._single = True
    _touch(m)


def set_base_pose(m: Mechanism, vec):                      # mechanism.jl:201
    m.base_pose = np.asarray(vec, dtype=np.float64).copy()
    m._Q = None
    m._single = True
    _touch(m)


def set_joint_angles(m: Mechanism, joints, angles):
    """mechanism.jl:223-231, extended to batches.

    ``angles`` is the reference's vector (``len(joints)`` values, then x, y, theta of the base when
    ``with_base``) or a batch of them: an ``(N, n_dof)`` array / CUDA tensor.  A batch whose memory is
    ``(n_dof, N)`` row-major (``tensor.t()`` of a contiguous ``(n_dof, N)``) is consumed as the SoA layout
    without a copy."""
    import torch
    n_dof = len(joints) + (3 if m.with_base else 0)
    m._ctrl = tuple(j.id for j in joints)
    if isinstance(angles, torch.Tensor) and angles.dim() == 2 or (not isinstance(angles, torch.Tensor) and np.ndim(angles) == 2):
        if angles.shape[1] != n_dof:
            raise ValueError("set_joint_angles: expected %d columns, got %d" % (n_dof, angles.shape[1]))
        m._Q = angles
        m._single = False
    else:
        a = np.asarray(angles.detach().cpu().numpy() if isinstance(angles, torch.Tensor) else angles, dtype=np.float64)
        if a.shape != (n_dof,):
            raise AssertionError("set_joint_angles: length(joints) + base dofs != length(angles)")   # :224
        for j, v in zip(joints, a[:len(joints)]):
            m.angles[j.id - 1] = v
        if m.with_base:
            m.base_pose = a[-3:].copy()
        m._Q = None
        m._single = True
    _touch(m)


def get_joint_angles(m: Mechanism, joints):                # mechanism.jl:203-221
    out = [m.angles[j.id - 1] for j in joints]
    if m.with_base:
        out.extend(m.base_pose)
    return np.array(out, dtype=np.float64)


def add_new_link(m: Mechanism, new_link, parent: Link, pose, name=None):
    """mechanism.jl:233-267: append ``new_link`` under ``parent`` through a Fixed joint.
    ``pose`` is a position 3-vector or a Transform."""
    if isinstance(new_link, str):          # convenience: add_new_link(m, "name", parent, pos)
        new_link = Link(new_link, link_type=User)
    if not isinstance(pose, Transform):
        pose = Transform(np.asarray(pose, dtype=np.float64))
    hlink_id = len(m.links) + 1
    parent.clink_ids.append(hlink_id)
    joint_id = len(m.joints) + 1
    j = Joint(new_link.name + "_joint", joint_id, parent.id, hlink_id, pose, FIXED)
    # (the reference pushes the new id to parent.clink_ids only, not to parent.cjoint_ids)
    new_link.id, new_link.pjoint_id, new_link.plink_id = hlink_id, joint_id, parent.id
    new_link.cjoint_ids, new_link.clink_ids, new_link.data = [], [], {}
    m.links.append(new_link)
    m.joints.append(j)
    m.linkid_map[new_link.name] = hlink_id
    m.jointid_map[j.name] = joint_id
    m.angles = np.append(m.angles, 0.0)
    m._structure_version += 1
    _touch(m)
    return new_link
