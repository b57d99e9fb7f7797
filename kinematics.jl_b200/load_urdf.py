"""parse_urdf (load_urdf.jl:20-80) without scikit-robot: the URDF is read with xml.etree following
the conventions of the parser the reference delegates to (urdfpy inside scikit-robot 0.0.15):
document-order enumeration, origin = Rz(yaw) Ry(pitch) Rx(roll) + xyz, normalised axis, box collision
geometry -> BoxMetaData(extents, origin)."""
from __future__ import annotations

import xml.etree.ElementTree as ET

import numpy as np

from .mechanism import (FIXED, PRISMATIC, REVOLUTE, BoxMetaData, CylinderMetaData, Joint, Link, Mechanism, MeshMetaData,
                        SphereMetaData)
from .transform import Transform


def _floats(s):
    return np.array([float(v) for v in s.split()], dtype=np.float64)


def _origin(node):
    T = np.eye(4)
    o = node.find("origin") if node is not None else None
    if o is not None:
        r, p, y = _floats(o.get("rpy", "0 0 0"))
        cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
        T[:3, :3] = [[cy * cp, cy * sp * sr - cr * sy, sy * sr + cy * cr * sp],
                     [cp * sy, cy * cr + sy * sp * sr, cr * sy * sp - cy * sr],
                     [-sp, cp * sr, cp * cr]]
        T[:3, 3] = _floats(o.get("xyz", "0 0 0"))
    return Transform(T)


def _geometry_meta(link_node):             # load_urdf.jl:1-18
    col = link_node.find("collision")
    if col is None:
        return None
    geom = col.find("geometry")
    if geom is None:
        return None
    box = geom.find("box")
    if box is not None:
        return BoxMetaData(_floats(box.get("size")), _origin(col))
    mesh = geom.find("mesh")
    if mesh is not None:
        return MeshMetaData(mesh.get("filename"), _origin(col))
    # "primitive type other than box is not supported yet" (load_urdf.jl:14): the reference keeps no meta data for
    # these, so they never become SDFs (sdf.py only takes BoxMetaData); they are kept here only as a vertex source
    # for the swept-sphere fit (swept_sphere.py)
    cyl = geom.find("cylinder")
    if cyl is not None:
        return CylinderMetaData(float(cyl.get("radius")), float(cyl.get("length")), _origin(col))
    sph = geom.find("sphere")
    if sph is not None:
        return SphereMetaData(float(sph.get("radius")), _origin(col))
    return None


def parse_urdf(urdf_path, with_base=False, robot_type="basic") -> Mechanism:
    root = ET.parse(urdf_path).getroot()
    link_nodes, joint_nodes = root.findall("link"), root.findall("joint")
    linkid_map = {n.get("name"): i + 1 for i, n in enumerate(link_nodes)}       # :29-32
    jointid_map = {n.get("name"): i + 1 for i, n in enumerate(joint_nodes)}     # :23-26
    links = []
    for n in link_nodes:
        l = Link(n.get("name"), geometric_meta_data=_geometry_meta(n))
        l.id = linkid_map[l.name]
        links.append(l)
    joints = []
    for n in joint_nodes:
        t = n.get("type")
        ax = n.find("axis")
        axis = _floats(ax.get("xyz")) if ax is not None else np.array([1.0, 0.0, 0.0])
        nrm = np.linalg.norm(axis)
        axis = axis / nrm if nrm > 0 else axis
        lim = n.find("limit")
        lo = float(lim.get("lower", "0")) if lim is not None else 0.0
        hi = float(lim.get("upper", "0")) if lim is not None else 0.0
        if t == "revolute":                # :48-64
            jt = REVOLUTE
        elif t == "continuous":
            jt, lo, hi = REVOLUTE, -np.inf, np.inf
        elif t == "prismatic":
            jt = PRISMATIC
        elif t == "fixed":
            jt, lo, hi = FIXED, -np.inf, np.inf
        else:
            raise ValueError("unsupported joint type %r" % t)
        j = Joint(n.get("name"), jointid_map[n.get("name")], linkid_map[n.find("parent").get("link")],
                  linkid_map[n.find("child").get("link")], _origin(n), jt, axis, lo, hi)
        joints.append(j)
        p, c = links[j.plink_id - 1], links[j.clink_id - 1]      # :69-75
        p.cjoint_ids.append(j.id)
        p.clink_ids.append(j.clink_id)
        c.pjoint_id, c.plink_id = j.id, j.plink_id
    return Mechanism(links, joints, linkid_map, jointid_map, with_base, robot_type)
