"""The synthetic branching mechanism (tests/golden/synthetic_tree.urdf) built with the product mirror and
with the oracle, with collision spheres on three branches and a two-box obstacle."""
import os

import numpy as np

import kinematics_jl_b200 as K
from oracle import ref_model as R
from conftest import GOLDEN

URDF = os.path.join(GOLDEN, "synthetic_tree.urdf")
JOINTS = ["j_c2", "j_trunk", "j_a1", "j_b2", "j_a2", "j_a3", "j_b1", "j_b3", "j_c1"]        # deliberately shuffled
SPHERES = [("a3", [[0, 0, 0], [0.1, 0, 0.05]], 0.05), ("trunk", [[0, 0, 0.2]], 0.12),
           ("b2_frame", [[0.02, 0, 0], [0, 0.03, 0.04], [0, 0, 0.1]], 0.04), ("c2", [[0, 0, 0.1]], 0.06),
           ("a_tool", [[0.01, 0.02, 0.03]], 0.03)]
BOX_POSES = [np.array([[0.36, -0.48, 0.8, 0.5], [0.8, 0.6, 0.0, 0.2], [-0.48, 0.64, 0.6, 0.4], [0, 0, 0, 1.0]]),
             np.array([[1.0, 0, 0, -0.3], [0, 1, 0, -0.2], [0, 0, 1, 0.6], [0, 0, 0, 1.0]])]
BOX_WIDTHS = [[0.3, 0.2, 0.25], [0.4, 0.4, 0.1]]


def product(with_base, n_ctrl=None):
    m = K.parse_urdf(URDF, with_base=with_base)
    joints = [K.find_joint(m, n) for n in JOINTS[:n_ctrl]]
    sscc = K.SweptSphereCollisionChecker(m)
    for link, centers, r in SPHERES:
        K.add_coll_links(sscc, K.find_link(m, link), centers, r)
    sdf = K.UnionSDF([K.BoxSDF(K.Transform(p), w) for p, w in zip(BOX_POSES, BOX_WIDTHS)])
    return m, joints, sscc, sdf


def oracle(with_base, n_ctrl=None):
    m = R.parse_urdf(URDF, with_base=with_base)
    joints = [R.find_joint(m, n) for n in JOINTS[:n_ctrl]]
    sscc = R.SweptSphereCollisionChecker(m)
    for link, centers, r in SPHERES:
        R.add_coll_links(sscc, R.find_link(m, link), centers, [r] * len(centers))
    sdf = R.UnionSDF([R.BoxSDF(p, w) for p, w in zip(BOX_POSES, BOX_WIDTHS)])
    return m, joints, sscc, sdf


def random_q(joints_ref, N, with_base, seed):
    rng = np.random.default_rng(seed)
    lo = np.array([j.lower if np.isfinite(j.lower) else -np.pi for j in joints_ref])
    hi = np.array([j.upper if np.isfinite(j.upper) else np.pi for j in joints_ref])
    q = lo + (hi - lo) * rng.random((N, len(joints_ref)))
    if with_base:
        q = np.concatenate([q, rng.uniform(-1, 1, (N, 2)), rng.uniform(-np.pi, np.pi, (N, 1))], axis=1)
    q[::17] = 0.0
    return np.ascontiguousarray(q)
