"""Shared scene builders for the tests: the same Fetch + sphere fixture + fridge scene built twice,
once with the product's host mirror (kinematics_jl_b200) and once with the oracle (oracle.ref_model)."""
import json
import os

import numpy as np

from oracle import ref_model as R
from conftest import DATA, FETCH_JOINT_NAMES
from scene_fetch import FRIDGE_STATE, product_fetch, sphere_fixture  # noqa: F401  (the product-side builders)


def random_configs(joints_ref, N, with_base, seed=0, zeros_every=0):
    """Uniform in the joint limits (continuous joints: [-pi, pi]); base ~ U[-1,1]^2 x U[-pi,pi]."""
    rng = np.random.default_rng(seed)
    lo = np.array([j.lower if np.isfinite(j.lower) else -np.pi for j in joints_ref])
    hi = np.array([j.upper if np.isfinite(j.upper) else np.pi for j in joints_ref])
    q = lo + (hi - lo) * rng.random((N, len(joints_ref)))
    if with_base:
        b = np.concatenate([rng.uniform(-1, 1, (N, 2)), rng.uniform(-np.pi, np.pi, (N, 1))], axis=1)
        q = np.concatenate([q, b], axis=1)
    if zeros_every:
        q[::zeros_every] = 0.0              # exercises the a == 0.0 short-cut (mechanism.jl:95,101)
    return np.ascontiguousarray(q)


def oracle_fetch(with_base=False, sphere_links=None):
    m = R.parse_urdf(os.path.join(DATA, "fetch.urdf"), with_base=with_base)
    joints = [R.find_joint(m, n) for n in FETCH_JOINT_NAMES]
    sscc = R.SweptSphereCollisionChecker(m)
    for s in sphere_fixture():
        if sphere_links is None or s["link"] in sphere_links:
            R.add_coll_links(sscc, R.find_link(m, s["link"]), s["centers"], [s["radius"]] * len(s["centers"]))
    return m, joints, sscc


def oracle_fridge_sdf():
    f = R.parse_urdf(os.path.join(DATA, "fridge.urdf"), with_base=True)
    R.set_joint_angles(f, [R.find_joint(f, "door_joint")], FRIDGE_STATE)
    return R.UnionSDF(f)


def fridge_boxes_host():
    """World box poses / widths of the fridge scene computed WITHOUT a GPU (numpy) -- only for the CPU
    flattener tests; the product's UnionSDF(mech) obtains them from the GPU FK."""
    sdf = oracle_fridge_sdf()
    return np.stack(sdf.poses), np.stack(sdf.widths)
