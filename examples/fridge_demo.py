"""The reference's fridge_demo.jl on the B200 backend, with the robot whose URDF ships with the reference's data
(Fetch with the planar base instead of PR2, whose assets are not available offline; no visualiser).

Same sequence of calls as fridge_demo.jl:9-41: the fridge as a UnionSDF at base pose (1.2, 0, 0) with the door opened to
2.0 rad, a swept-sphere collision checker on the arm links, a target pose 1.2 m above the fridge's base frame (inside
the cabinet), inverse kinematics without and then with the collision constraint, and a 10-waypoint trajectory from the
tucked arm to the IK solution (the straight line between the two passes through the cabinet wall) with margin 0.03 (SLSQP: scipy, the reference's own SCIPY back-end; NLopt is not
installed here).  Every evaluation -- link transforms, Jacobians, sphere-vs-SDF distances and gradients -- runs in
libkin_b200 on the GPU.

    python examples/fridge_demo.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kinematics_jl_b200 as K  # noqa: E402
import scene_fetch  # noqa: E402


def main():
    fridge = K.parse_urdf(os.path.join(ROOT, "data", "fridge.urdf"), with_base=True)            # fridge_demo.jl:9-11
    sdf = K.UnionSDF(fridge)
    robot, joints, sscc = scene_fetch.product_fetch(with_base=True)                               # :13-19 (Fetch + sphere fixture)
    K.set_joint_angles(fridge, [K.find_joint(fridge, "door_joint")], [2.0, 1.2, 0.0, 0.0])      # :26-28
    tf_fridge = K.get_transform(fridge, K.find_link(fridge, "base_link"))
    tf_target = K.Transform(np.array([0.0, 0.0, 1.2])) * tf_fridge                                # :30-31
    q_start = np.array([0.0, 1.32, 1.4, -0.2, 1.72, 0.0, 1.66, 0.0, 0.0, 0.0, 0.0])               # Fetch's tucked arm, base at the origin
    q_seed = np.array([0.2, 0.0, 0.0, 0.0, 0.5, 0.0, 0.5, 0.0, 0.0, 0.0, 0.0])                    # IK seed: arm forward (reset_manip_pose, :27)
    K.set_joint_angles(robot, joints, q_seed)
    link = K.find_link(robot, "gripper_link")
    t0 = time.perf_counter()
    K.inverse_kinematics(robot, link, joints, tf_target, with_rot=True, ftol=1e-8)                # :36 pre-solve
    q_goal, res = K.inverse_kinematics(robot, link, joints, tf_target, sscc, sdf, with_rot=True, ftol=1e-8)   # :37
    t_ik = time.perf_counter() - t0
    K.set_joint_angles(robot, joints, q_goal)
    pose = K.get_transform(robot, link)
    d_goal = K.compute_coll_dists(sscc, joints, sdf)
    print("IK: success=%s  position error %.2e  rpy error %.2e  min sphere distance %.4f  (%.2f s)"
          % (res.success, np.abs(K.translation(pose) - K.translation(tf_target)).max(), np.abs(K.rpy(pose) - K.rpy(tf_target)).max(),
             d_goal.min(), t_ik))
    n_wp, margin = 10, 0.03                                                                       # :39-41
    t0 = time.perf_counter()
    q_seq, ret = K.plan_trajectory(sscc, joints, sdf, q_start, q_goal, n_wp, ftol_abs=1e-4, solver="SCIPY", margin=margin)
    t_plan = time.perf_counter() - t0
    import torch
    K.set_joint_angles(robot, joints, torch.as_tensor(q_seq, device="cuda"))
    d = K.compute_coll_dists(sscc, joints, sdf).cpu().numpy()
    K.set_joint_angles(robot, joints, torch.as_tensor(K.create_straight_trajectory(q_start, q_goal, n_wp).reshape(n_wp, -1), device="cuda"))
    d_line = K.compute_coll_dists(sscc, joints, sdf).cpu().numpy()
    print("plan_trajectory: success=%s  iterations %d  objective %.5f  min sphere distance over the %d waypoints %.4f "
          "(straight line: %.4f)  (%.2f s)" % (ret.success, ret.nit, ret.fun, n_wp, d.min(), d_line.min(), t_plan))
    return res, ret, d_goal, d, d_line


if __name__ == "__main__":
    main()
