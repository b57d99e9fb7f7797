"""Swept-sphere generation (collision.jl:16-30 delegates it to scikit-robot, which is not available: parity with
skrobot's numbers is UNPINNED).  The mesh-free restatement in kinematics.jl_b200/swept_sphere.py is checked through
the properties the published algorithm guarantees: centres collinear on the principal (PCA) axis and evenly spaced, one
common radius = 1.01 x the largest distance of a vertex from that axis, every vertex within tol x radius of the union,
invariance under rigid motions of the vertex cloud; plus the STL reader and the primitive samplers."""
import os
import struct

import numpy as np
import pytest

from kinematics_jl_b200 import swept_sphere as SS
from kinematics_jl_b200.load_urdf import parse_urdf
from kinematics_jl_b200.mechanism import BoxMetaData, find_link
from conftest import DATA


def capsule_cloud(length, radius, n=400, seed=0):
    rng = np.random.default_rng(seed)
    t = rng.uniform(-length / 2, length / 2, n)
    a = rng.uniform(0, 2 * np.pi, n)
    r = radius * np.sqrt(rng.uniform(0, 1, n))
    r[:40] = radius                                     # some points on the hull
    return np.stack([r * np.cos(a), r * np.sin(a), t], axis=1)


def test_fit_properties_on_a_capsule_like_cloud():
    v = capsule_cloud(0.6, 0.05)
    c, r = SS.compute_swept_sphere(v)
    # the radius is 1.01 x the largest distance of a vertex from the fitted axis (through the mean, along the
    # eigenvector of the largest eigenvalue of the scatter matrix) -- recomputed here independently
    mean = v.mean(axis=0)
    w, U = np.linalg.eigh((v - mean).T @ (v - mean))
    ax = U[:, np.argmax(w)]
    dist = np.linalg.norm((v - mean) - np.outer((v - mean) @ ax, ax), axis=1)
    assert r == pytest.approx(dist.max() * SS.MARGIN_FACTOR, rel=1e-9) and 0.05 < r < 0.06
    assert len(c) >= 2
    d = c - c.mean(axis=0)
    u, s, vt = np.linalg.svd(d)
    assert s[1] < 1e-12 * max(1.0, s[0])                                      # centres collinear ...
    assert abs(abs(vt[0] @ np.array([0, 0, 1.0])) - 1) < 1e-3                 # ... along the long (PCA) axis
    steps = np.linalg.norm(np.diff(c, axis=0), axis=1)
    np.testing.assert_allclose(steps, steps[0], rtol=1e-9)                    # evenly spaced
    assert SS.max_jut_ratio(v, c, r) < 0.1                                    # tol of the reference call
    # fewer spheres with a looser tolerance, an explicit count is honoured
    c2, r2 = SS.compute_swept_sphere(v, tol=0.5)
    assert len(c2) <= len(c) and r2 == r
    c3, _ = SS.compute_swept_sphere(v, n_sphere=7)
    assert len(c3) == 7


def test_fit_is_equivariant_under_rigid_motions():
    v = capsule_cloud(0.4, 0.08, seed=1)
    c, r = SS.compute_swept_sphere(v)
    a = 0.7
    Rm = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]]) @ np.array([[1, 0, 0], [0, 0, -1.0], [0, 1, 0]])
    t = np.array([0.3, -0.2, 1.1])
    c2, r2 = SS.compute_swept_sphere(v @ Rm.T + t)
    assert r2 == pytest.approx(r, rel=1e-9)
    # the set of centres is the same up to the order along the axis
    want = c @ Rm.T + t
    got = c2 if np.linalg.norm(c2[0] - want[0]) < np.linalg.norm(c2[-1] - want[0]) else c2[::-1]
    np.testing.assert_allclose(got, want, atol=1e-9)


def test_stl_readers_and_primitive_samplers(tmp_path):
    tri = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]], [[0, 0, 1], [1, 0, 1], [0, 1, 1]]], dtype=np.float32)
    binary = tmp_path / "b.stl"
    with open(binary, "wb") as f:
        f.write(b"\0" * 80 + struct.pack("<I", len(tri)))
        for t in tri:
            f.write(struct.pack("<3f", 0, 0, 1) + t.tobytes() + b"\0\0")
    np.testing.assert_array_equal(SS.load_stl_vertices(str(binary)), tri.reshape(-1, 3))
    ascii_ = tmp_path / "a.stl"
    ascii_.write_text("solid x\n" + "".join("facet normal 0 0 1\n outer loop\n" + "".join("  vertex %g %g %g\n" % tuple(p) for p in t) +
                                             " endloop\nendfacet\n" for t in tri) + "endsolid x\n")
    np.testing.assert_array_equal(SS.load_stl_vertices(str(ascii_)), tri.reshape(-1, 3))
    box = SS.primitive_vertices("box", [0.2, 0.4, 1.0])
    assert len(box) == 26 and np.abs(box).max(axis=0).tolist() == [0.1, 0.2, 0.5]
    cyl = SS.primitive_vertices("cylinder", (0.28, 0.33))
    np.testing.assert_allclose(np.hypot(cyl[:, 0], cyl[:, 1]), 0.28)
    assert set(np.round(cyl[:, 2], 6)) == {0.165, -0.165}
    sph = SS.primitive_vertices("sphere", 0.065)
    np.testing.assert_allclose(np.linalg.norm(sph, axis=1), 0.065)
    # a box along z -> spheres along z, radius = half diagonal of the cross-section x 1.01
    c, r = SS.compute_swept_sphere(box)
    assert r == pytest.approx(np.hypot(0.1, 0.2) * 1.01) and np.abs(c[:, :2]).max() < 1e-12


def test_spheres_from_urdf_box_primitives():
    """data/fridge.urdf has box collision primitives: its links get swept spheres without any mesh."""
    f = parse_urdf(os.path.join(DATA, "fridge.urdf"))
    link = find_link(f, "door_link")
    assert isinstance(link.geometric_meta_data, BoxMetaData)
    v = SS.link_vertices(link)
    c, r = SS.compute_swept_sphere(v)
    assert SS.max_jut_ratio(v, c, r) < 0.1 and len(c) >= 2
    # the Fetch meshes are not shipped: no vertices, and add_coll_links says so instead of inventing spheres
    import kinematics_jl_b200 as K
    m = parse_urdf(os.path.join(DATA, "fetch.urdf"))
    assert SS.link_vertices(find_link(m, "wrist_flex_link")) is None
    sscc = K.SweptSphereCollisionChecker(m)
    with pytest.raises(K.KinError):
        K.add_coll_links(sscc, find_link(m, "wrist_flex_link"))
    n0 = len(m.links)
    K.add_coll_links(sscc, find_link(m, "wrist_flex_link"), vertices=capsule_cloud(0.2, 0.06))
    assert len(sscc.sphere_links) == len(m.links) - n0 >= 2 and all(r_ == sscc.sphere_radii[0] for r_ in sscc.sphere_radii)
