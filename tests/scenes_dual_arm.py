"""A dual-arm mechanism (tests/golden/dual_arm.urdf: torso + 2 x 7 joints, the size class of the PR2 that the
reference's fridge_demo.jl drives) built with the product mirror and with the oracle: 15 configuration columns,
18 with the planar base, 19 collision spheres on both arms and the torso, a three-box obstacle."""
import os

import numpy as np

import kinematics_jl_b200 as K
from oracle import ref_model as R
from conftest import GOLDEN

URDF = os.path.join(GOLDEN, "dual_arm.urdf")
JOINTS = ["torso_joint"] + ["%s_joint%d" % (s, i) for s in "lr" for i in range(1, 8)]
SPHERES = [("%s_link%d" % (s, i), [[0.05, 0, 0]] if i % 2 else [[0.03, 0, 0], [0.1, 0, 0.01]], 0.05)
           for s in "lr" for i in range(2, 8)] + [("torso", [[0, 0, 0.2]], 0.15)]
BOX_POSES = [np.array([[1.0, 0, 0, 0.8], [0, 1, 0, 0.0], [0, 0, 1, 0.9], [0, 0, 0, 1.0]]),
             np.array([[0.36, -0.48, 0.8, 0.5], [0.8, 0.6, 0.0, 0.5], [-0.48, 0.64, 0.6, 1.0], [0, 0, 0, 1.0]]),
             np.array([[1.0, 0, 0, 0.5], [0, 1, 0, -0.6], [0, 0, 1, 0.6], [0, 0, 0, 1.0]])]
BOX_WIDTHS = [[0.3, 0.8, 0.05], [0.2, 0.2, 0.6], [0.4, 0.2, 0.3]]


def product(with_base):
    m = K.parse_urdf(URDF, with_base=with_base)
    joints = [K.find_joint(m, n) for n in JOINTS]
    sscc = K.SweptSphereCollisionChecker(m)
    for link, centers, r in SPHERES:
        K.add_coll_links(sscc, K.find_link(m, link), centers, r)
    sdf = K.UnionSDF([K.BoxSDF(K.Transform(p), w) for p, w in zip(BOX_POSES, BOX_WIDTHS)])
    return m, joints, sscc, sdf


def oracle(with_base):
    m = R.parse_urdf(URDF, with_base=with_base)
    joints = [R.find_joint(m, n) for n in JOINTS]
    sscc = R.SweptSphereCollisionChecker(m)
    for link, centers, r in SPHERES:
        R.add_coll_links(sscc, R.find_link(m, link), centers, [r] * len(centers))
    sdf = R.UnionSDF([R.BoxSDF(p, w) for p, w in zip(BOX_POSES, BOX_WIDTHS)])
    return m, joints, sscc, sdf
