"""The benchmark / smoke scene of SURVEY 8(d), built with the PRODUCT's host mirror only (no oracle import):
Fetch (data/fetch.urdf, 8 control joints) + the 16-sphere fixture (data/fetch_spheres.json; swept-sphere
placement is an input fixture, parity unpinned -- SURVEY 8c) + the fridge of fridge_demo.jl:28 as a UnionSDF."""
import json
import os

import numpy as np

import kinematics_jl_b200 as K

ROOT = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(ROOT, "data")

FETCH_JOINT_NAMES = [
    "torso_lift_joint", "shoulder_pan_joint", "shoulder_lift_joint", "upperarm_roll_joint",
    "elbow_flex_joint", "forearm_roll_joint", "wrist_flex_joint", "wrist_roll_joint"]
FRIDGE_STATE = [2.0, 1.2, 0.0, 0.0]       # door angle, base x, y, theta (fridge_demo.jl:28)


def sphere_fixture():
    return json.load(open(os.path.join(DATA, "fetch_spheres.json")))["links"]


def product_fetch(with_base=False, sphere_links=None):
    """-> (Mechanism, control joints, SweptSphereCollisionChecker with the fixture's spheres)."""
    m = K.parse_urdf(os.path.join(DATA, "fetch.urdf"), with_base=with_base)
    joints = [K.find_joint(m, n) for n in FETCH_JOINT_NAMES]
    sscc = K.SweptSphereCollisionChecker(m)
    for s in sphere_fixture():
        if sphere_links is None or s["link"] in sphere_links:
            K.add_coll_links(sscc, K.find_link(m, s["link"]), s["centers"], s["radius"])
    return m, joints, sscc


def product_fridge_sdf():
    """UnionSDF of data/fridge.urdf at base (1.2, 0, 0), door 2.0 (the obstacle FK runs on the GPU)."""
    fridge = K.parse_urdf(os.path.join(DATA, "fridge.urdf"), with_base=True)
    K.set_joint_angles(fridge, [K.find_joint(fridge, "door_joint")], FRIDGE_STATE)
    return K.UnionSDF(fridge)


def joint_limits(joints):
    """(lo, hi) with continuous joints mapped to [-pi, pi]."""
    lo = np.array([j.lower_limit if np.isfinite(j.lower_limit) else -np.pi for j in joints])
    hi = np.array([j.upper_limit if np.isfinite(j.upper_limit) else np.pi for j in joints])
    return lo, hi


# ---- a dual-arm mechanism (data/dual_arm.urdf: torso + 2 x 7 joints; 15 configuration columns, 18 with the planar base:
#      the size class of the PR2 that the reference's fridge_demo.jl and test_inverse_kinematics.jl:26-88 drive) ----
DUAL_ARM_JOINT_NAMES = ["torso_joint"] + ["%s_joint%d" % (s, i) for s in "lr" for i in range(1, 8)]
DUAL_ARM_SPHERES = [("%s_link%d" % (s, i), [[0.05, 0, 0]] if i % 2 else [[0.03, 0, 0], [0.1, 0, 0.01]], 0.05)
                    for s in "lr" for i in range(2, 8)] + [("torso", [[0, 0, 0.2]], 0.15)]
DUAL_ARM_BOX_POSES = [np.array([[1.0, 0, 0, 0.8], [0, 1, 0, 0.0], [0, 0, 1, 0.9], [0, 0, 0, 1.0]]),
                      np.array([[0.36, -0.48, 0.8, 0.5], [0.8, 0.6, 0.0, 0.5], [-0.48, 0.64, 0.6, 1.0], [0, 0, 0, 1.0]]),
                      np.array([[1.0, 0, 0, 0.5], [0, 1, 0, -0.6], [0, 0, 1, 0.6], [0, 0, 0, 1.0]])]
DUAL_ARM_BOX_WIDTHS = [[0.3, 0.8, 0.05], [0.2, 0.2, 0.6], [0.4, 0.2, 0.3]]


def product_dual_arm(with_base=False):
    """-> (Mechanism, the 15 control joints, checker with 19 spheres on both arms and the torso, three-box UnionSDF)."""
    m = K.parse_urdf(os.path.join(DATA, "dual_arm.urdf"), with_base=with_base)
    joints = [K.find_joint(m, n) for n in DUAL_ARM_JOINT_NAMES]
    sscc = K.SweptSphereCollisionChecker(m)
    for link, centers, r in DUAL_ARM_SPHERES:
        K.add_coll_links(sscc, K.find_link(m, link), centers, r)
    sdf = K.UnionSDF([K.BoxSDF(K.Transform(p), w) for p, w in zip(DUAL_ARM_BOX_POSES, DUAL_ARM_BOX_WIDTHS)])
    return m, joints, sscc, sdf
