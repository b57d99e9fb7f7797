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
