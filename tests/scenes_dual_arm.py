"""A dual-arm mechanism (data/dual_arm.urdf: torso + 2 x 7 joints, the size class of the PR2 that the reference's
fridge_demo.jl drives) built with the product mirror (scene_fetch.product_dual_arm) and with the oracle: 15 configuration
columns, 18 with the planar base, 19 collision spheres on both arms and the torso, a three-box obstacle."""
import os

from oracle import ref_model as R
from conftest import DATA
import scene_fetch

URDF = os.path.join(DATA, "dual_arm.urdf")
JOINTS = scene_fetch.DUAL_ARM_JOINT_NAMES
SPHERES = scene_fetch.DUAL_ARM_SPHERES
BOX_POSES = scene_fetch.DUAL_ARM_BOX_POSES
BOX_WIDTHS = scene_fetch.DUAL_ARM_BOX_WIDTHS

product = scene_fetch.product_dual_arm


def oracle(with_base):
    m = R.parse_urdf(URDF, with_base=with_base)
    joints = [R.find_joint(m, n) for n in JOINTS]
    sscc = R.SweptSphereCollisionChecker(m)
    for link, centers, r in SPHERES:
        R.add_coll_links(sscc, R.find_link(m, link), centers, [r] * len(centers))
    sdf = R.UnionSDF([R.BoxSDF(p, w) for p, w in zip(BOX_POSES, BOX_WIDTHS)])
    return m, joints, sscc, sdf
