"""Sphere / cylinder SDF primitives (EXTENSION, SURVEY 8 f4: the reference has boxes only, load_urdf.jl:10-15,
sdf.jl:92-94) -- the oracle's restatement of the textbook formulas, pinned here by known answers and by the
properties a signed distance has, because no reference vector exists for them ("parity unpinned" for these two
primitive kinds; the box path stays pinned by test_sdf.jl's KATs in test_oracle_golden.py)."""
import os

import numpy as np

from oracle import ref_model as R
from conftest import GOLDEN


def _pose(t, Rm=None):
    T = np.eye(4)
    T[:3, 3] = t
    if Rm is not None:
        T[:3, :3] = Rm
    return T


def _rot(axis, a):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * K @ K


def test_sphere_known_answers():
    s = R.SphereSDF(_pose([1.0, 2.0, 3.0]), 0.5)
    assert s([1.0, 2.0, 3.0]) == -0.5
    assert abs(s([1.0, 2.0, 3.5])) < 1e-15
    assert abs(s([4.0, 6.0, 3.0]) - 4.5) < 1e-15            # 3-4-5 triangle
    np.testing.assert_allclose(s.gradient([4.0, 6.0, 3.0], analytic=True), [0.6, 0.8, 0.0], atol=1e-15)


def test_cylinder_known_answers():
    c = R.CylinderSDF(_pose([0.0, 0.0, 1.0]), 0.5, 2.0)      # z in [0, 2], radius 0.5
    assert c([0.0, 0.0, 1.0]) == -0.5                         # centre: nearest surface is the side
    assert abs(c([0.0, 0.0, 1.9]) + 0.1) < 1e-15              # near the top cap
    assert abs(c([1.5, 0.0, 1.0]) - 1.0) < 1e-15              # radially outside
    assert abs(c([0.0, 0.0, 3.0]) - 1.0) < 1e-15              # above the cap
    assert abs(c([0.5 + 0.3, 0.0, 2.0 + 0.4]) - 0.5) < 1e-15  # off the rim: 3-4-5
    np.testing.assert_allclose(c.gradient([0.8, 0.0, 2.4], analytic=True), [0.6, 0.0, 0.8], atol=1e-15)
    # rotated: axis along world x
    c2 = R.CylinderSDF(_pose([0.0, 0.0, 0.0], _rot([0, 1, 0], np.pi / 2)), 0.25, 1.0)
    assert abs(c2([1.0, 0.0, 0.0]) - 0.5) < 1e-12 and abs(c2([0.0, 0.0, 0.75]) - 0.5) < 1e-12


def test_forward_difference_gradient_agrees_with_closed_form_and_is_unit():
    rng = np.random.default_rng(3)
    u = R.UnionSDF([R.BoxSDF(_pose([0.3, 0.0, 0.2], _rot([1, 2, 3], 0.4)), [0.4, 0.2, 0.6]),
                    R.SphereSDF(_pose([-0.4, 0.3, 0.0]), 0.25),
                    R.CylinderSDF(_pose([0.0, -0.5, 0.1], _rot([1, 0, 1], 1.1)), 0.15, 0.7)])
    assert u.kinds == [0, 1, 2]
    seen = set()
    for p in (rng.random((3000, 3)) - 0.5) * 2.0:
        d = u(p)
        seen.add((u.argmin, d < 0))
        ga, gf = u.gradient(p, analytic=True), u.gradient(p)
        assert abs(np.linalg.norm(ga) - 1.0) < 1e-12
        if np.abs(gf - ga).max() > 1e-5:            # only next to a kink of the distance field
            q = p + 2e-7 * ga
            u(q)
            assert np.abs(u.gradient(q, analytic=True) - ga).max() > 1e-6 or abs(d) < 1e-6
    assert {k for k, _ in seen} == {1, 2, 3} and any(inside for _, inside in seen)


def test_urdf_primitives_become_union_members():
    m = R.parse_urdf(os.path.join(GOLDEN, "prims_obstacle.urdf"), with_base=True)
    R.set_joint_angles(m, [R.find_joint(m, "arm_joint")], [0.7, 0.5, -0.2, 0.3])
    assert R.UnionSDF(m).kinds == [0]                          # the reference's behaviour: boxes only
    u = R.UnionSDF(m, primitives=True)
    assert u.kinds == [0, 2, 2, 1]
    ball = R.get_transform(m, R.find_link(m, "ball"))[:3, 3]
    assert abs(u(ball) + 0.09) < 1e-12 and u.argmin == 4
