"""GPU parity of the caller-side evaluations: IK objective,
PoseConstraint, the per-waypoint IneqConst stack, straight-line trajectories and the batched IK driver."""
import os

import numpy as np
import pytest
import torch

import kinematics_jl_b200 as K
from oracle import ref_model as R
from conftest import DATA
import scenes

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")


def target_T(xyz, rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    T = np.eye(4)
    T[:3, :3] = [[cy * cp, cy * sp * sr - cr * sy, sy * sr + cy * cr * sp],
                 [cp * sy, cy * cr + sy * sp * sr, cr * sy * sp - cy * sr],
                 [-sp, cp * sr, cp * cr]]
    T[:3, 3] = xyz
    return T


@pytest.mark.parametrize("with_base", [False, True])
@pytest.mark.parametrize("with_rot", [True, False])
def test_ik_objective_and_pose_constraint_vs_oracle(with_base, with_rot):
    m, joints, _ = scenes.product_fetch(with_base)
    mo, jo, _ = scenes.oracle_fetch(with_base)
    link, link_o = K.find_link(m, "gripper_link"), R.find_link(mo, "gripper_link")
    q = scenes.random_configs(jo, 300, with_base, seed=51)
    Tt = target_T([0.3, -0.4, 1.2], [0.2, -0.1, 0.4])
    for Q in (dev(q), dev(q).t().contiguous().t()):                 # AoS and SoA
        K.set_joint_angles(m, joints, Q)
        f, g = K.ik_objective(m, link, joints, K.Transform(Tt), with_rot)
        v, jt = K.pose_constraint(m, link, joints, K.Transform(Tt), with_rot)
        f, g, v, jt = (x.cpu().numpy() for x in (f, g, v, jt))
        for n in range(0, 300, 7):
            fo, go = R.ik_objective(mo, link_o, jo, q[n], Tt, with_rot)
            vo, jto = R.pose_constraint(mo, link_o, jo, q[n], Tt, with_rot)
            np.testing.assert_allclose(f[n], fo, rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(g[n], go, rtol=1e-11, atol=1e-11)
            np.testing.assert_allclose(v[n], vo, rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(jt[n], jto, rtol=1e-11, atol=1e-11)
    # single configuration, reference-style call; per-configuration targets
    K.set_joint_angles(m, joints, q[0])
    f1, g1 = K.ik_objective(m, link, joints, K.Transform(Tt), with_rot)
    fo, go = R.ik_objective(mo, link_o, jo, q[0], Tt, with_rot)
    np.testing.assert_allclose(f1, fo, rtol=1e-12)
    np.testing.assert_allclose(g1, go, rtol=1e-11, atol=1e-11)
    tg = np.tile(np.concatenate([Tt[:3, 3], K.rpy(K.Transform(Tt))]), (300, 1))
    K.set_joint_angles(m, joints, dev(q))
    f2, _ = K.ik_objective(m, link, joints, dev(tg), with_rot)
    np.testing.assert_allclose(f2.cpu().numpy(), f, rtol=1e-13)


@pytest.mark.parametrize("with_base", [False, True])
def test_ineq_const_vs_oracle(with_base):
    """planning.jl:55-68 with the box of test/test_planning.jl:23-25, n_wp = 10."""
    m, joints, sscc = scenes.product_fetch(with_base)
    mo, jo, so = scenes.oracle_fetch(with_base)
    pose = np.eye(4)
    pose[:3, 3] = [0.4, -0.25, 0.7]
    box, box_o = K.BoxSDF(K.Transform(pose), [0.05, 0.05, 0.5]), R.BoxSDF(pose, [0.05, 0.05, 0.5])
    n_wp, margin = 10, 0.02
    nd = 8 + (3 if with_base else 0)
    q0 = np.zeros(nd)
    q1 = scenes.random_configs(jo, 1, with_base, seed=61)[0]
    xi = K.create_straight_trajectory(q0, q1, n_wp)
    assert xi.shape == (nd * n_wp,)
    np.testing.assert_allclose(xi.reshape(n_wp, nd)[3], q0 + (q1 - q0) / (n_wp - 1) * 3, rtol=1e-15)
    G = K.IneqConst(sscc, joints, box, n_wp, margin)
    val, blocks = G(xi)
    val_o, blocks_o = R.ineq_const(so, jo, box_o, xi, n_wp, margin)
    np.testing.assert_allclose(val, val_o, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(blocks, blocks_o, rtol=0, atol=1e-7)
    assert np.array_equal(val == 0.05, val_o == 0.05)            # truncated entries: (margin + 0.05) - margin
    dense = G.dense(blocks)
    assert dense.shape == (nd * n_wp, 16 * n_wp) and np.count_nonzero(dense[:nd, 16:]) == 0
    nv, nj = K.nloptize(G)(xi)
    assert np.array_equal(nv, -val)
    # batched problems: P straight lines, flattened (problem, waypoint) batch
    qs = dev(scenes.random_configs(jo, 6, with_base, seed=62))
    qg = dev(scenes.random_configs(jo, 6, with_base, seed=63))
    X = K.create_straight_trajectory(qs, qg, n_wp)
    assert X.shape == (6, n_wp, nd)
    V, B = G(X)
    for p in range(6):
        vo, bo = R.ineq_const(so, jo, box_o, X[p].reshape(-1).cpu().numpy(), n_wp, margin)
        np.testing.assert_allclose(V[p].reshape(-1).cpu().numpy(), vo, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(B[p].cpu().numpy(), bo, rtol=0, atol=1e-7)


def test_batched_ik_reaches_reachable_targets():
    """Config 4 acceptance (test_inverse_kinematics.jl:16-23 criterion: pose within 1e-3)."""
    m, joints, _ = scenes.product_fetch(False)
    mo, jo, _ = scenes.oracle_fetch(False)
    link = K.find_link(m, "gripper_link")
    N = 2048
    q_true = scenes.random_configs(jo, N, False, seed=71)
    K.set_joint_angles(m, joints, dev(q_true))
    T = K.get_transform(m, link).cpu().numpy()
    tg = np.zeros((N, 6))
    for n in range(N):
        M = np.eye(4)
        M[:3] = T[n]
        tg[n, :3], tg[n, 3:] = M[:3, 3], K.rpy(K.Transform(M))
    # the reference's own test target first: (0.3, -0.4, 1.2), identity rotation
    tg[0] = [0.3, -0.4, 1.2, 0, 0, 0]
    q0 = np.tile(np.array([0.2, 0, 0, 0, 0.5, 0, 0.5, 0]), (N, 1))      # a non-singular seed
    q, f = K.inverse_kinematics_batch(m, link, joints, dev(tg), dev(q0), with_rot=True, iters=100)
    K.set_joint_angles(m, joints, q)
    v, _ = K.pose_constraint(m, link, joints, dev(tg), True)
    v[:, 3:] = torch.remainder(v[:, 3:] + np.pi, 2 * np.pi) - np.pi
    err = v.abs().amax(dim=1).cpu().numpy()
    print("batched IK: %.1f %% of %d targets within 1e-3" % (100 * (err < 1e-3).mean(), N))
    assert err[0] < 1e-3
    assert (err < 1e-3).mean() > 0.9
    lo = np.array([j.lower_limit for j in joints])
    hi = np.array([j.upper_limit for j in joints])
    qn = q.cpu().numpy()
    assert np.all(qn >= lo - 1e-12) and np.all(qn <= hi + 1e-12)


def test_batched_collision_aware_ik():
    """Config 4 with the collision constraint of the reference's two-stage IK (inverse_kinematics.jl:1-21,
    margin 0.02) against the thin box of test/test_planning.jl:23-25: about a quarter of the unconstrained
    solutions run an arm sphere through the box; the penalised solve clears all of them.  (A penalty
    trades pose error for clearance where the margin is active, so somewhat fewer targets end within 1e-3
    than with SLSQP's hard constraint; the solver itself is out of scope, the evaluations are what is tested.)"""
    m, joints, sscc = scenes.product_fetch(False)
    link = K.find_link(m, "gripper_link")
    pose = np.eye(4)
    pose[:3, 3] = [0.4, -0.25, 0.8]
    box = K.BoxSDF(K.Transform(pose), [0.05, 0.05, 0.5])
    N = 1024
    rng = np.random.default_rng(5)
    tg = np.zeros((N, 6))
    tg[:, 0], tg[:, 1], tg[:, 2] = rng.uniform(0.55, 0.8, N), rng.uniform(-0.3, 0.3, N), rng.uniform(0.7, 1.1, N)
    q0 = np.tile(np.array([0.2, 0, 0, 0, 0.5, 0, 0.5, 0]), (N, 1))

    def outcome(q):
        K.set_joint_angles(m, joints, q)
        d = K.compute_coll_dists(sscc, joints, box)
        v, _ = K.pose_constraint(m, link, joints, dev(tg), True)
        v[:, 3:] = torch.remainder(v[:, 3:] + np.pi, 2 * np.pi) - np.pi
        return v.abs().amax(dim=1) < 1e-3, d.amin(dim=1) > -1e-3

    q_free, _ = K.inverse_kinematics_batch(m, link, joints, dev(tg), dev(q0), with_rot=True, iters=80)
    r0, c0 = outcome(q_free)
    q, _ = K.inverse_kinematics_batch(m, link, joints, dev(tg), dev(q0), with_rot=True, iters=80, sscc=sscc, sdf=box, margin=0.02)
    r1, c1 = outcome(q)
    f = lambda t: 100 * float(t.double().mean())
    print("IK vs thin box: unconstrained reached %.1f %% / clear %.1f %% / both %.1f %%; constrained reached %.1f %% / clear %.1f %% / both %.1f %%"
          % (f(r0), f(c0), f(r0 & c0), f(r1), f(c1), f(r1 & c1)))
    assert f(c0) < 90.0                      # the scenario does exercise the constraint
    assert f(c1) > 99.0
    assert f(r1 & c1) > 55.0
