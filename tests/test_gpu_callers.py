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


def _thin_box():
    pose = np.eye(4)
    pose[:3, 3] = [0.4, -0.25, 0.8]
    return pose, [0.05, 0.05, 0.5]                         # the pillar of test/test_planning.jl:23-25


def test_batched_collision_aware_ik(monkeypatch):
    """Config 4 with the reference's HARD collision constraint (inverse_kinematics.jl:14-19: dists - 0.02 >= 0 after the
    collision-free warm start) solved on the device for the whole batch (kin_ik_solve with collision = 1).

    Targets are feasible by construction: gripper poses of random configurations that clear the pillar of
    test/test_planning.jl:23-25 by 0.03, with the gripper behind / beside the pillar.  The pose-only solutions from the
    common seed violate the margin for part of them (the scenario does exercise the constraint); the constrained
    solve must end within 1e-3 of the pose (test_inverse_kinematics.jl:22-23) AND with every sphere at >= margin,
    for >= 90 % of the targets from the one seed and >= 97 % with three re-seeded restarts.  The reported objective
    and smallest distance are checked against the oracle at the returned configurations."""
    m, joints, sscc = scenes.product_fetch(False)
    mo, jo, so = scenes.oracle_fetch(False)
    link, link_o = K.find_link(m, "gripper_link"), R.find_link(mo, "gripper_link")
    pose, width = _thin_box()
    box, box_o = K.BoxSDF(K.Transform(pose), width), R.BoxSDF(pose, width)
    margin, N = 0.02, 1024
    qc = scenes.random_configs(jo, 60 * N, False, seed=3)
    K.set_joint_angles(m, joints, dev(qc))
    d = K.compute_coll_dists(sscc, joints, box).amin(dim=1).cpu().numpy()
    p = K.get_transform(m, link)[:, :, 3].cpu().numpy()
    keep = (d > 0.03) & (p[:, 0] > 0.5) & (np.abs(p[:, 1] + 0.25) < 0.3) & (p[:, 2] > 0.5) & (p[:, 2] < 1.1)
    idx = np.nonzero(keep)[0][:N]
    assert idx.size == N
    tg = _pose_targets(m, joints, link, qc[idx])
    q0 = np.tile(np.array([0.2, 0, 0, 0, 0.5, 0, 0.5, 0]), (N, 1))

    def outcome(q):
        err = _pose_err(m, joints, link, q, tg)
        K.set_joint_angles(m, joints, q)
        dm = K.compute_coll_dists(sscc, joints, box).amin(dim=1)
        return (err < 1e-3).cpu().numpy(), (dm >= margin - 1e-5).cpu().numpy(), dm.cpu().numpy()

    q_free, _ = K.inverse_kinematics_batch(m, link, joints, dev(tg), dev(q0), with_rot=True, iters=40)
    r0, c0, _ = outcome(q_free)
    lib = K.load_library()
    n0 = lib.kin_launch_count()
    q1, f1, its, dmin1 = K.ik_solve_device(m, link, joints, dev(tg), q_free, with_rot=True, iters=60, sscc=sscc, sdf=box, margin=margin)
    torch.cuda.synchronize()
    # init + 61 x (kin_eval, step) + the final distance evaluation + finish: no other launches, nothing read back
    assert lib.kin_launch_count() - n0 == 1 + 61 * 2 + 2
    r1, c1, dm1 = outcome(q1)
    np.testing.assert_allclose(dmin1.cpu().numpy(), dm1, rtol=1e-12, atol=1e-12)
    # the one-warp-per-problem step kernel (opt-in; an independent parallelisation of the step) on the same problems: the same iterates, bit for bit
    monkeypatch.setenv("KIN_IK_STEP", "warp")
    q1w, f1w, itsw, dmin1w = K.ik_solve_device(m, link, joints, dev(tg), q_free, with_rot=True, iters=60, sscc=sscc, sdf=box, margin=margin)
    monkeypatch.delenv("KIN_IK_STEP")
    assert torch.equal(q1w, q1) and torch.equal(f1w, f1) and torch.equal(itsw, its) and torch.equal(dmin1w, dmin1)
    q2, f2, dmin2 = K.inverse_kinematics_batch(m, link, joints, dev(tg), dev(q0), with_rot=True, iters=40, sscc=sscc, sdf=box, margin=margin,
                                               restarts=3, return_dmin=True)
    r2, c2, dm2 = outcome(q2)
    pc = lambda t: 100 * float(np.mean(t))
    print("IK vs thin box (feasible targets): pose-only reached %.1f %% / margin kept %.1f %% / both %.1f %%; constrained, one seed: "
          "%.1f / %.1f / %.1f %% (mean %.1f iterations); with 3 restarts: %.1f / %.1f / %.1f %%"
          % (pc(r0), pc(c0), pc(r0 & c0), pc(r1), pc(c1), pc(r1 & c1), float(its.double().mean()), pc(r2), pc(c2), pc(r2 & c2)))
    assert pc(c0) < 97.0                     # the scenario does exercise the constraint
    assert pc(c1) > 99.0 and pc(r1 & c1) >= 90.0
    assert pc(r2 & c2) >= 97.0
    lo = np.array([j.lower_limit for j in joints])
    hi = np.array([j.upper_limit for j in joints])
    qn = q2.cpu().numpy()
    assert np.all(qn >= lo - 1e-12) and np.all(qn <= hi + 1e-12)
    # the oracle at the returned configurations: objective, distances, constraint
    d_o, _, _ = R.batch_collision(so, jo, box_o, qn, with_grads=False)
    np.testing.assert_allclose(dmin2.cpu().numpy(), d_o.min(axis=1), rtol=1e-12, atol=1e-12)
    good = r2 & c2
    assert np.all(d_o.min(axis=1)[good] >= margin - 1e-5)
    fn = f2.cpu().numpy()
    for n in range(0, N, 97):
        fo, _ = R.ik_objective(mo, link_o, jo, qn[n], target_T(tg[n, :3], tg[n, 3:]), True)
        if fn[n] < 1e-2:
            np.testing.assert_allclose(fn[n], fo, rtol=1e-9, atol=1e-18)


def test_batched_collision_aware_ik_arbitrary_targets():
    """The same solve on targets drawn without regard to feasibility (a box of positions behind the pillar, identity
    rotation; about 30 % of them cannot be reached with every sphere 0.02 clear -- repeated re-seeding of the
    prototype profiles/proto_ik_coll.py saturates at 70 %): an infeasible problem must still end with the constraint
    satisfied (the pose error is what gives), and the reached-and-clear share must be near that ceiling."""
    m, joints, sscc = scenes.product_fetch(False)
    link = K.find_link(m, "gripper_link")
    pose, width = _thin_box()
    box = K.BoxSDF(K.Transform(pose), width)
    N = 1024
    rng = np.random.default_rng(5)
    tg = np.zeros((N, 6))
    tg[:, 0], tg[:, 1], tg[:, 2] = rng.uniform(0.55, 0.8, N), rng.uniform(-0.3, 0.3, N), rng.uniform(0.7, 1.1, N)
    q0 = np.tile(np.array([0.2, 0, 0, 0, 0.5, 0, 0.5, 0]), (N, 1))
    q, f, dmin = K.inverse_kinematics_batch(m, link, joints, dev(tg), dev(q0), with_rot=True, iters=40, sscc=sscc, sdf=box, margin=0.02,
                                            restarts=2, return_dmin=True)
    reached = (_pose_err(m, joints, link, q, tg) < 1e-3).cpu().numpy()
    clear = (dmin >= 0.02 - 1e-3).cpu().numpy()
    print("IK vs thin box (arbitrary targets): reached %.1f %% / clear %.1f %% / both %.1f %%"
          % (100 * reached.mean(), 100 * clear.mean(), 100 * (reached & clear).mean()))
    assert clear.mean() > 0.97
    assert (reached & clear).mean() > 0.62


# ------------------------------------------------------------------------------------------------
# PoseConstraint with several (link, target, with_rot) triples in ONE library call; EqConst stacking
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("with_base", [False, True])
def test_pose_constraint_multi_link_one_call_and_eq_const(with_base):
    """planning.jl:124-137 loops over (link, target, with_rot); here it is one kin_pose_residual_multi call
    (two kernel launches whatever the number of links); EqConst (planning.jl:140-176) places the rows."""
    m, joints, _ = scenes.product_fetch(with_base)
    mo, jo, _ = scenes.oracle_fetch(with_base)
    names, rots = ["gripper_link", "elbow_flex_link", "wrist_roll_link"], [True, False, True]
    links, links_o = [K.find_link(m, n) for n in names], [R.find_link(mo, n) for n in names]
    Tts = [target_T([0.3, -0.4, 1.2], [0.2, -0.1, 0.4]), target_T([0.1, 0.2, 0.9], [0, 0, 0]), target_T([0.5, 0.1, 1.0], [-0.3, 0.2, 0.1])]
    nd = 8 + (3 if with_base else 0)
    n_cons = 6 + 3 + 6
    q = scenes.random_configs(jo, 200, with_base, seed=91)
    lib = K.load_library()
    for Q in (dev(q), dev(q).t().contiguous().t()):                 # AoS and SoA
        K.set_joint_angles(m, joints, Q)
        K.pose_constraint(m, links, joints, [K.Transform(T) for T in Tts], rots)      # builds the program
        n0 = lib.kin_launch_count()
        v, jt = K.pose_constraint(m, links, joints, [K.Transform(T) for T in Tts], rots)
        assert lib.kin_launch_count() - n0 == 2                       # kin_eval + residual kernel, not 2 per link
        assert v.shape == (200, n_cons) and jt.shape == (200, nd, n_cons)
        v, jt = v.cpu().numpy(), jt.cpu().numpy()
        for n in range(0, 200, 11):
            c0 = 0
            for lo_, T, w in zip(links_o, Tts, rots):
                vo, jto = R.pose_constraint(mo, lo_, jo, q[n], T, w)
                dim = 6 if w else 3
                np.testing.assert_allclose(v[n, c0:c0 + dim], vo, rtol=1e-12, atol=1e-12)
                np.testing.assert_allclose(jt[n, :, c0:c0 + dim], jto, rtol=1e-11, atol=1e-11)
                c0 += dim
    # EqConst: start / goal configuration rows + a pose constraint at waypoint 4, reference-style flat xi
    n_wp = 7
    xi = scenes.random_configs(jo, n_wp, with_base, seed=92).reshape(-1)
    qs, qg = xi[:nd] + 0.01, xi[-nd:] - 0.02
    H = K.EqConst(n_wp, [K.ConfigurationConstraint(1, nd, qs), K.ConfigurationConstraint(n_wp, nd, qg),
                         K.PoseConstraint(4, nd, links, [K.Transform(T) for T in Tts], rots, m, joints)])
    val, jac = H(xi)
    val_o, jac_o = R.eq_const(mo, jo, xi, n_wp, [("config", 1, qs), ("config", n_wp, qg), ("pose", 4, links_o, Tts, rots)])
    assert jac.shape == (nd * n_wp, 2 * nd + n_cons)
    np.testing.assert_allclose(val, val_o, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(jac, jac_o, rtol=1e-11, atol=1e-11)
    nz = np.abs(jac).max(axis=1) > 0                                  # rows of waypoints 1, 4 and n_wp only (planning.jl:83-86,121-122)
    assert set(np.nonzero(nz)[0] // nd) <= {0, 3, n_wp - 1}
    h, dh = K.scipynize(H)
    assert np.array_equal(h(xi), val) and np.array_equal(dh(xi), jac.T)
    # batched problems: (P, n_wp, n_dof) tensor -> values + row blocks, expanded to the dense matrices
    Xb = dev(scenes.random_configs(jo, 5 * n_wp, with_base, seed=93).reshape(5, n_wp, nd))
    Hb = K.EqConst(n_wp, [K.ConfigurationConstraint(1, nd, dev(qs)), K.ConfigurationConstraint(n_wp, nd, dev(qg)),
                          K.PoseConstraint(4, nd, links, [K.Transform(T) for T in Tts], rots, m, joints)])
    vb, blocks = Hb(Xb)
    dense = Hb.dense_batch(blocks)
    for p in range(5):
        vo, jo_ = R.eq_const(mo, jo, Xb[p].reshape(-1).cpu().numpy(), n_wp,
                             [("config", 1, qs), ("config", n_wp, qg), ("pose", 4, links_o, Tts, rots)])
        np.testing.assert_allclose(vb[p].cpu().numpy(), vo, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(dense[p].cpu().numpy(), jo_, rtol=1e-11, atol=1e-11)


# ------------------------------------------------------------------------------------------------
# config 5 at its real size: 4096 problems x 64 waypoints, margin 0.03, fridge SDF
# ------------------------------------------------------------------------------------------------
def test_ineq_const_config5_full_size():
    """BASELINE.json configs[4]: the IneqConst stack (planning.jl:55-68, truncation = margin + 0.05) of 4096
    straight-line problems x 64 waypoints in one batched call; 32 whole problems (2048 waypoints) sampled
    across the batch are checked against the oracle, the rest through size-independent properties."""
    m, joints, sscc = scenes.product_fetch(False)
    mo, jo, so = scenes.oracle_fetch(False)
    import scene_fetch
    sdf, sdf_o = scene_fetch.product_fridge_sdf(), scenes.oracle_fridge_sdf()
    P_, n_wp, margin, nd, S = 4096, 64, 0.03, 8, 16
    qs = dev(scenes.random_configs(jo, P_, False, seed=101))
    qg = dev(scenes.random_configs(jo, P_, False, seed=102))
    X = K.create_straight_trajectory(qs, qg, n_wp)
    G = K.IneqConst(sscc, joints, sdf, n_wp, margin)
    V, B = G(X)
    assert V.shape == (P_, n_wp, S) and B.shape == (P_, n_wp, nd, S)
    # properties over all 262 144 waypoints: truncated entries are exactly (margin + 0.05) - margin with zero
    # gradient blocks (collision.jl:84-86), nothing exceeds the truncation value, everything is finite
    trunc_val = (margin + 0.05) - margin
    assert bool(torch.isfinite(V).all()) and bool(torch.isfinite(B).all())
    assert float(V.max()) <= trunc_val
    is_tr = V == trunc_val
    assert 0.05 < float(is_tr.double().mean()) < 0.999          # the scene exercises both branches
    assert float(B.permute(0, 1, 3, 2)[is_tr].abs().max()) == 0.0
    # waypoint 1 / n_wp of every problem are the start / goal configurations
    K.set_joint_angles(m, joints, qs)
    d0 = K.compute_coll_dists(sscc, joints, sdf)
    np.testing.assert_allclose(torch.minimum(d0, torch.tensor(margin + 0.05, device="cuda", dtype=torch.float64)).sub(margin).cpu().numpy(),
                               V[:, 0].cpu().numpy(), rtol=1e-12, atol=1e-12)
    # 32 problems (2048 waypoints) against the oracle
    idx = np.linspace(0, P_ - 1, 32).astype(int)
    n_tr = 0
    for p in idx:
        vo, bo = R.ineq_const(so, jo, sdf_o, X[p].reshape(-1).cpu().numpy(), n_wp, margin)
        np.testing.assert_allclose(V[p].reshape(-1).cpu().numpy(), vo, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(B[p].cpu().numpy(), bo, rtol=0, atol=1e-7)
        assert np.array_equal(V[p].reshape(-1).cpu().numpy() == trunc_val, vo == trunc_val)
        n_tr += int((vo == trunc_val).sum())
    assert 0 < n_tr < 32 * n_wp * S


# ------------------------------------------------------------------------------------------------
# the reference's solver-level callers with the solver that exists here (scipy SLSQP = planning.jl:388-394)
# ------------------------------------------------------------------------------------------------
def _oracle_plan(so, jo, sdf_o, q_start, q_goal, n_wp, margin, ftol_abs):
    """plan_trajectory's SCIPY back-end (planning.jl:388-394) driven by the ORACLE's evaluations."""
    from scipy.optimize import minimize
    nd = len(q_start)
    A = R.objective_matrix(n_wp, np.ones(nd))
    lo = [j.lower for j in jo] + [-np.inf] * (nd - len(jo))
    hi = [j.upper for j in jo] + [np.inf] * (nd - len(jo))
    bounds = [(a if np.isfinite(a) else None, b if np.isfinite(b) else None) for a, b in zip(lo, hi)] * n_wp
    S = len(so.sphere_links)

    def g(xi):
        return R.ineq_const(so, jo, sdf_o, xi, n_wp, margin)[0]

    def dg(xi):
        blocks = R.ineq_const(so, jo, sdf_o, xi, n_wp, margin)[1]
        J = np.zeros((nd * n_wp, S * n_wp))
        for i in range(n_wp):
            J[i * nd:(i + 1) * nd, i * S:(i + 1) * S] = blocks[i]
        return J.T
    cons = [("config", 1, q_start), ("config", n_wp, q_goal)]
    xi0 = np.concatenate([q_start + (q_goal - q_start) / (n_wp - 1) * i for i in range(n_wp)])
    return minimize(lambda x: R.objective(A, x)[0], xi0, jac=lambda x: R.objective(A, x)[1], method="SLSQP", bounds=bounds,
                    options={"ftol": ftol_abs, "maxiter": 200},
                    constraints=[{"type": "ineq", "fun": g, "jac": dg},
                                 {"type": "eq", "fun": lambda x: R.eq_const(so.mech, jo, x, n_wp, cons)[0],
                                  "jac": lambda x: R.eq_const(so.mech, jo, x, n_wp, cons)[1].T}])


@pytest.mark.parametrize("with_base", [False, True])
def test_inverse_kinematics_and_plan_trajectory_slsqp(with_base):
    """test/test_inverse_kinematics.jl:16-23 and test/test_planning.jl:3-46 with SLSQP (scipy: the reference's own
    SCIPY back-end; NLopt is not installed): IK reaches (0.3, -0.4, 1.2) within 1e-3, the planned trajectory keeps
    every sphere of every waypoint above -1e-2, and agrees with the same solver driven by the oracle."""
    m, joints, sscc = scenes.product_fetch(with_base)
    mo, jo, so = scenes.oracle_fetch(with_base)
    nd = 8 + (3 if with_base else 0)
    link = K.find_link(m, "gripper_link")
    target = K.Transform(np.array([0.3, -0.4, 1.2]))
    q_start = np.zeros(nd)
    K.set_joint_angles(m, joints, q_start)
    # ftol: the reference's default 1e-5 is NLopt's ftol_abs; scipy's SLSQP stops on the same quantity a little earlier
    # (|f - f_prev| < ftol with f ~ err^2), so the 1e-3 pose tolerance of the reference test needs ftol = 1e-8 here
    q_goal, res = K.inverse_kinematics(m, link, joints, target, with_rot=True, ftol=1e-8)
    assert res.success
    K.set_joint_angles(m, joints, q_goal)
    pose = K.get_transform(m, link)
    np.testing.assert_allclose(K.translation(pose), [0.3, -0.4, 1.2], atol=1e-3)
    np.testing.assert_allclose(K.rpy(pose), [0, 0, 0], atol=1e-3)
    # planning scenario of test_planning.jl: thin box, n_wp = 10, goal from a position-only IK
    K.set_joint_angles(m, joints, q_start)
    q_goal, res = K.inverse_kinematics(m, link, joints, target, with_rot=False, ftol=1e-8)
    assert res.success
    pose_b = np.eye(4)
    pose_b[:3, 3] = [0.4, -0.25, 0.7]
    box, box_o = K.BoxSDF(K.Transform(pose_b), [0.05, 0.05, 0.5]), R.BoxSDF(pose_b, [0.05, 0.05, 0.5])
    n_wp = 10
    q_seq, ret = K.plan_trajectory(sscc, joints, box, q_start, q_goal, n_wp, ftol_abs=1e-5, solver="SCIPY")
    assert ret.success, ret.message
    K.set_joint_angles(m, joints, dev(q_seq))
    d = K.compute_coll_dists(sscc, joints, box).cpu().numpy()
    assert d.shape == (n_wp, 16) and np.all(d > -1e-2)              # test_planning.jl:41-45
    np.testing.assert_allclose(q_seq[0], q_start, atol=1e-6)
    np.testing.assert_allclose(q_seq[-1], q_goal, atol=1e-6)
    # the same solver driven by the oracle's evaluations.  With ftol_abs = 1e-5 SLSQP stops on a flat part of the
    # objective, so the two runs (whose gradients differ by the ~1e-9 noise of the FD quotient) are compared on the
    # objective and on feasibility, and the iterates to the accuracy that stopping rule supports; tightened to
    # ftol_abs = 1e-10 they must agree closely
    ret_o = _oracle_plan(so, jo, box_o, q_start, q_goal, n_wp, 0.02, 1e-5)
    assert ret_o.success
    np.testing.assert_allclose(ret.fun, ret_o.fun, rtol=1e-3, atol=1e-5)
    print("plan_trajectory (ftol 1e-5): max |q - q_oracle| = %.2e, f = %.6f vs %.6f, iterations %d vs %d"
          % (np.abs(q_seq.reshape(-1) - ret_o.x).max(), ret.fun, ret_o.fun, ret.nit, ret_o.nit))
    assert np.abs(q_seq.reshape(-1) - ret_o.x).max() < 5e-2
    q_seq2, ret2 = K.plan_trajectory(sscc, joints, box, q_start, q_goal, n_wp, ftol_abs=1e-10, solver="SCIPY")
    ret_o2 = _oracle_plan(so, jo, box_o, q_start, q_goal, n_wp, 0.02, 1e-10)
    print("plan_trajectory (ftol 1e-10): max |q - q_oracle| = %.2e, f = %.8f vs %.8f, iterations %d vs %d"
          % (np.abs(q_seq2.reshape(-1) - ret_o2.x).max(), ret2.fun, ret_o2.fun, ret2.nit, ret_o2.nit))
    np.testing.assert_allclose(ret2.fun, ret_o2.fun, rtol=1e-6, atol=1e-8)
    assert np.abs(q_seq2.reshape(-1) - ret_o2.x).max() < 2e-3
    with pytest.raises(K.KinError):
        K.plan_trajectory(sscc, joints, box, q_start, q_goal, n_wp, solver="NLOPT")


def test_collision_aware_inverse_kinematics_slsqp_hard_constraint():
    """inverse_kinematics.jl:1-21: the two-stage solve with the HARD constraint dists - 0.02 >= 0 (tol 1e-8 in the
    reference).  A target behind the thin box of test_planning.jl: the unconstrained solution runs the arm
    through it, the constrained one keeps every sphere >= margin (to the solver's tolerance)."""
    m, joints, sscc = scenes.product_fetch(False)
    link = K.find_link(m, "gripper_link")
    pose_b = np.eye(4)
    pose_b[:3, 3] = [0.4, -0.25, 0.8]
    box = K.BoxSDF(K.Transform(pose_b), [0.05, 0.05, 0.5])
    q0 = np.array([0.2, 0, 0, 0, 0.5, 0, 0.5, 0])
    found = 0
    # targets low behind the pillar: SLSQP's unconstrained solution puts a forearm / wrist sphere inside the margin there
    # (found with the oracle-driven solver); the constrained solve must clear it and still reach the target
    for x in (0.35, 0.4, 0.45, 0.5):
        for y in (-0.5, -0.45, -0.4):
            tgt = K.Transform(np.array([x, y, 0.6]))
            K.set_joint_angles(m, joints, q0)
            q_free, r_free = K.inverse_kinematics(m, link, joints, tgt, with_rot=False)
            d_free = K.compute_coll_dists(sscc, joints, box)
            if d_free.min() > 0.02:
                continue
            K.set_joint_angles(m, joints, q0)
            q_c, r_c = K.inverse_kinematics(m, link, joints, tgt, sscc=sscc, sdf=box, with_rot=False)
            d_c = K.compute_coll_dists(sscc, joints, box)
            if r_c.success:
                found += 1
                assert d_c.min() >= 0.02 - 1e-5                       # the constraint holds at the solution (SLSQP's own tolerance)
                np.testing.assert_allclose(K.translation(K.get_transform(m, link)), K.translation(tgt), atol=5e-3)
    print("collision-aware SLSQP IK: %d of 12 targets needed and satisfied the constraint" % found)
    assert found >= 6


def _oracle_ik(mo, jo, link, target, x0, with_rot, sscc=None, sdf=None, ftol=1e-8, margin=0.02):
    """The solver of K.inverse_kinematics (scipy SLSQP, same options, same two stages) driven by the ORACLE's evaluations."""
    from scipy.optimize import minimize
    nb = mo.n_dof_extra
    lo, hi = [j.lower for j in jo] + [-np.inf] * nb, [j.upper for j in jo] + [np.inf] * nb
    bounds = [(a if np.isfinite(a) else None, b if np.isfinite(b) else None) for a, b in zip(lo, hi)]

    def solve(x0, cons):
        return minimize(lambda x: R.ik_objective(mo, link, jo, x, target, with_rot), x0, jac=True, method="SLSQP", bounds=bounds,
                        constraints=cons, options={"ftol": ftol, "maxiter": 200})
    if sscc is None:
        return solve(x0, ())
    x0 = solve(x0, ()).x
    return solve(x0, [{"type": "ineq", "fun": lambda x: R.ineq_const(sscc, jo, sdf, x, 1, margin)[0],
                       "jac": lambda x: R.ineq_const(sscc, jo, sdf, x, 1, margin)[1][0].T}])


def _dual_arm_target(mo, jo, so, sdf_o, with_base, trial):
    """A reachable tool pose: FK of an in-limit configuration (fixed seed; `trial` picks one of the draws)."""
    rng = np.random.default_rng(5)
    lo, hi = np.array([j.lower for j in jo]), np.array([j.upper for j in jo])
    for _ in range(trial + 1):
        qt = lo + (hi - lo) * (0.25 + 0.5 * rng.random(len(jo)))
    if with_base:
        qt = np.concatenate([qt, [0.2, -0.1, 0.3]])
    R.set_joint_angles(mo, jo, qt)
    return R.get_transform(mo, R.find_link(mo, "l_tool")).copy()


@pytest.mark.parametrize("with_base", [False, True])
def test_inverse_kinematics_dual_arm_as_the_pr2_test(with_base):
    """test/test_inverse_kinematics.jl:26-88 ("inverse kinematics_pr2") on the dual-arm fixture (the PR2 URDF and its
    meshes are not available; tests/scenes_dual_arm.py has the same shape: both arms = 14 joints, 17 columns with the
    planar base, the left tool frame as the target link, collision spheres on both arms).  "no collision": with and
    without rotation, status success and the pose within 1e-3; "with collision" (with_base, with_rot, bistage): the
    pose within 1e-3 and every sphere clear (the reference asserts > -1e-5; the constraint is dists - 0.02 >= 0).
    Each solve is also run with the oracle driving the same SLSQP: same iteration count, same solution."""
    import scenes_dual_arm as DA
    m, joints, sscc, sdf = DA.product(with_base)
    mo, jo, so, sdf_o = DA.oracle(with_base)
    joints, jo = joints[1:], jo[1:]                      # vcat(rarm_joints, larm_joints): the torso stays put
    nd = len(joints) + (3 if with_base else 0)
    link, link_o = K.find_link(m, "l_tool"), R.find_link(mo, "l_tool")
    tgt = _dual_arm_target(mo, jo, so, sdf_o, with_base, 0)
    for with_rot in (False, True):
        K.set_joint_angles(m, joints, np.zeros(nd))
        q, res = K.inverse_kinematics(m, link, joints, K.Transform(tgt), with_rot=with_rot, ftol=1e-8)
        assert res.success
        pose = K.get_transform(m, link)
        assert np.linalg.norm(K.translation(pose) - tgt[:3, 3]) < 1e-3
        if with_rot:
            np.testing.assert_allclose(K.rpy(pose), K.rpy(K.Transform(tgt)), atol=1e-3)
        R.set_joint_angles(mo, jo, np.zeros(nd))
        res_o = _oracle_ik(mo, jo, link_o, tgt, np.zeros(nd), with_rot)
        assert res.nit == res_o.nit
        np.testing.assert_allclose(q, res_o.x, atol=1e-6)
    if not with_base:
        return
    n_active = 0
    for trial in (0, 2):            # unconstrained solutions come within 0.011 / 0.006 of a box (oracle-driven prototype)
        tgt = _dual_arm_target(mo, jo, so, sdf_o, True, trial)
        K.set_joint_angles(m, joints, np.zeros(nd))
        K.inverse_kinematics(m, link, joints, K.Transform(tgt), with_rot=True, ftol=1e-8)
        d_free = float(K.compute_coll_dists(sscc, joints, sdf).min())
        K.set_joint_angles(m, joints, np.zeros(nd))
        q, res = K.inverse_kinematics(m, link, joints, K.Transform(tgt), sscc=sscc, sdf=sdf, with_rot=True, use_bistage=True, ftol=1e-8)
        assert res.success
        pose = K.get_transform(m, link)
        assert np.linalg.norm(K.translation(pose) - tgt[:3, 3]) < 1e-3
        np.testing.assert_allclose(K.rpy(pose), K.rpy(K.Transform(tgt)), atol=1e-3)
        d = np.asarray(K.compute_coll_dists(sscc, joints, sdf))          # one configuration: a host vector
        assert np.all(d > -1e-5) and d.min() >= 0.02 - 1e-5
        n_active += d_free < 0.02
        res_o = _oracle_ik(mo, jo, link_o, tgt, np.zeros(nd), True, so, sdf_o)
        assert res_o.success and abs(res.nit - res_o.nit) <= 1
        np.testing.assert_allclose(q, res_o.x, atol=1e-4)
    assert n_active == 2            # the constraint was needed in both cases


# ------------------------------------------------------------------------------------------------
# the device-resident batched IK solve (one kernel launch: kin_ik_solve)
# ------------------------------------------------------------------------------------------------
def _pose_targets(m, joints, link, q_true):
    K.set_joint_angles(m, joints, dev(q_true))
    T = K.get_transform(m, link).cpu().numpy()
    tg = np.zeros((len(q_true), 6))
    for n in range(len(q_true)):
        M = np.eye(4)
        M[:3] = T[n]
        tg[n, :3], tg[n, 3:] = M[:3, 3], K.rpy(K.Transform(M))
    return tg


def _pose_err(m, joints, link, q, tg, with_rot=True):
    K.set_joint_angles(m, joints, q)
    v, _ = K.pose_constraint(m, link, joints, dev(tg), with_rot)
    if with_rot:
        v[:, 3:] = torch.remainder(v[:, 3:] + np.pi, 2 * np.pi) - np.pi
    return v.abs().amax(dim=1)


@pytest.mark.parametrize("with_base", [False, True])
def test_device_resident_ik_solve(with_base, monkeypatch):
    """Config 4: the whole Levenberg-Marquardt solve in one launch of the generated kernel.  Acceptance as in
    test/test_inverse_kinematics.jl:19-23 (pose within 1e-3): >= 99 % of reachable targets within 40 iterations per
    solve when failed problems are re-seeded twice; iterates stay inside the joint limits; the objective the kernel
    reports is the reference's f_objective at the returned configuration (checked against the oracle); and the result
    agrees in quality with the multi-kernel Levenberg-Marquardt path it replaces."""
    m, joints, _ = scenes.product_fetch(with_base)
    mo, jo, _ = scenes.oracle_fetch(with_base)
    link, link_o = K.find_link(m, "gripper_link"), R.find_link(mo, "gripper_link")
    N, nd = 4096, 8 + (3 if with_base else 0)
    q_true = scenes.random_configs(jo, N, with_base, seed=81)
    tg = _pose_targets(m, joints, link, q_true)
    tg[0] = [0.3, -0.4, 1.2, 0, 0, 0]                      # the reference's own test target
    seed = np.array([0.2, 0, 0, 0, 0.5, 0, 0.5, 0] + [0.0] * (nd - 8))
    q0 = np.tile(seed, (N, 1))
    lib = K.load_library()
    n0 = lib.kin_launch_count()
    q, f, its = K.ik_solve_device(m, link, joints, dev(tg), dev(q0), with_rot=True, iters=40)
    torch.cuda.synchronize()
    assert lib.kin_launch_count() - n0 == 1                # ONE launch for the whole solve
    err = _pose_err(m, joints, link, q, tg).cpu().numpy()
    ok1 = (err < 1e-3).mean()
    assert err[0] < 1e-3 and ok1 > 0.93
    assert int(its.max()) <= 40 and int(its.min()) >= 1 and float(its.double().mean()) < 30      # most stop early on f < ftol
    lo = np.array([j.lower_limit for j in joints] + [-np.inf] * (nd - 8))
    hi = np.array([j.upper_limit for j in joints] + [np.inf] * (nd - 8))
    qn = q.cpu().numpy()
    assert np.all(qn >= lo - 1e-12) and np.all(qn <= hi + 1e-12)
    # the reported objective is f_objective of the oracle at the returned configuration (angles wrapped)
    fn = f.cpu().numpy()
    for n in range(0, N, 173):
        Tt = target_T(tg[n, :3], tg[n, 3:])
        fo, _ = R.ik_objective(mo, link_o, jo, qn[n], Tt, True)
        if err[n] < 1e-1:                                    # no 2 pi wrap in play
            np.testing.assert_allclose(fn[n], fo, rtol=1e-9, atol=1e-18)
    # two re-seeded restarts for the problems that did not converge
    q2, f2 = K.inverse_kinematics_batch(m, link, joints, dev(tg), dev(q0), with_rot=True, iters=40, restarts=2)
    err2 = _pose_err(m, joints, link, q2, tg).cpu().numpy()
    print("device IK (with_base=%s): %.2f %% within 1e-3 after one solve, %.2f %% with two restarts, mean iterations %.1f"
          % (with_base, 100 * ok1, 100 * (err2 < 1e-3).mean(), float(its.double().mean())))
    assert (err2 < 1e-3).mean() >= 0.99
    # without the run-time compiler the same method runs as (kin_eval, step kernel) pairs: same quality
    monkeypatch.setenv("KIN_DISABLE_JIT", "1")
    q3, _ = K.inverse_kinematics_batch(m, link, joints, dev(tg[:1024]), dev(q0[:1024]), with_rot=True, iters=40)
    monkeypatch.delenv("KIN_DISABLE_JIT")
    err3 = _pose_err(m, joints, link, q3, tg[:1024]).cpu().numpy()
    assert abs((err3 < 1e-3).mean() - (err[:1024] < 1e-3).mean()) < 0.03
    # position-only targets (3 rows)
    q4, f4, _ = K.ik_solve_device(m, link, joints, dev(tg), dev(q0), with_rot=False, iters=40)
    err4 = _pose_err(m, joints, link, q4, tg, with_rot=False).cpu().numpy()
    assert (err4 < 1e-3).mean() > 0.97


@pytest.mark.parametrize("with_base", [False, True])
def test_device_resident_ik_solve_dual_arm(with_base, monkeypatch):
    """kin_ik_solve on a model beyond 12 configuration columns (tests/scenes_dual_arm.py: 15, 18 with the planar base -- the
    shape of the reference's PR2 inverse-kinematics test, test_inverse_kinematics.jl:26-88): the generated one-launch
    kernel for the pose-only solve, and the collision-constrained solve whose step kernel runs its run-time-sized
    instance (normal equations in local memory).  Pose within 1e-3 as the reference asserts, iterates inside the limits,
    the reported objective = the oracle's f_objective at the returned configuration, the reported minimum distance = the
    oracle's at the returned configuration, the constraint dists - 0.02 >= 0 kept."""
    import scenes_dual_arm as DA
    import scenes_synthetic as SS
    m, joints, sscc, sdf = DA.product(with_base)
    mo, jo, so, sdf_o = DA.oracle(with_base)
    link, link_o = K.find_link(m, "l_tool"), R.find_link(mo, "l_tool")
    N, nd = 2048, 15 + (3 if with_base else 0)
    q_true = SS.random_q(jo, N, with_base, seed=91)
    tg = _pose_targets(m, joints, link, q_true)
    q0 = np.zeros((N, nd))
    q, f, its = K.ik_solve_device(m, link, joints, dev(tg), dev(q0), with_rot=True, iters=40)
    err = _pose_err(m, joints, link, q, tg).cpu().numpy()
    lo = np.array([j.lower_limit for j in joints] + [-np.inf] * (nd - 15))
    hi = np.array([j.upper_limit for j in joints] + [np.inf] * (nd - 15))
    qn, fn = q.cpu().numpy(), f.cpu().numpy()
    assert np.all(qn >= lo - 1e-12) and np.all(qn <= hi + 1e-12)
    for n in range(0, N, 101):
        fo, _ = R.ik_objective(mo, link_o, jo, qn[n], target_T(tg[n, :3], tg[n, 3:]), True)
        if err[n] < 1e-1:                                    # no 2 pi wrap in play
            np.testing.assert_allclose(fn[n], fo, rtol=1e-9, atol=1e-18)
    q2, f2 = K.inverse_kinematics_batch(m, link, joints, dev(tg), dev(q0), with_rot=True, iters=40, restarts=2)
    err2 = _pose_err(m, joints, link, q2, tg).cpu().numpy()
    print("dual-arm device IK (%d columns): %.2f %% within 1e-3 after one solve, %.2f %% with two restarts, mean iterations %.1f"
          % (nd, 100 * (err < 1e-3).mean(), 100 * (err2 < 1e-3).mean(), float(its.double().mean())))
    # fixed base: only the torso and the left arm move the tool, and the targets are drawn over the FULL joint ranges (many
    # sit at a limit): 83 % / 93 % measured from the all-zero seed; with the planar base every target is an easy one
    assert (err < 1e-3).mean() > (0.95 if with_base else 0.78) and (err2 < 1e-3).mean() >= (0.99 if with_base else 0.90)
    # under the hard collision constraint against the three boxes
    q3, f3, dmin = K.inverse_kinematics_batch(m, link, joints, dev(tg), dev(q0), with_rot=True, iters=40, sscc=sscc, sdf=sdf, margin=0.02,
                                              restarts=2, return_dmin=True)
    reached = (_pose_err(m, joints, link, q3, tg) < 1e-3).cpu().numpy()
    clear = (dmin >= 0.02 - 1e-3).cpu().numpy()
    q3n, dn = q3.cpu().numpy(), dmin.cpu().numpy()
    assert np.all(q3n >= lo - 1e-12) and np.all(q3n <= hi + 1e-12)
    for n in range(0, N, 101):
        R.set_joint_angles(mo, jo, q3n[n])
        np.testing.assert_allclose(dn[n], R.compute_coll_dists(so, jo, sdf_o).min(), rtol=1e-10, atol=1e-12)
    print("dual-arm constrained IK (%d columns): reached %.1f %% / clear %.1f %% / both %.1f %%"
          % (nd, 100 * reached.mean(), 100 * clear.mean(), 100 * (reached & clear).mean()))
    assert clear.mean() > 0.97 and (reached & clear).mean() > (0.90 if with_base else 0.60)
    # the step kernel above 12 columns is the run-time-sized one-thread-per-problem instance (normal equations in shared
    # memory); the one-warp-per-problem kernel, an independent parallelisation of the same step, gives the same iterates
    monkeypatch.setenv("KIN_IK_STEP", "warp")
    q4, f4, dmin4 = K.inverse_kinematics_batch(m, link, joints, dev(tg), dev(q0), with_rot=True, iters=40, sscc=sscc, sdf=sdf, margin=0.02,
                                               restarts=2, return_dmin=True)
    monkeypatch.delenv("KIN_IK_STEP")
    assert torch.equal(q4, q3) and torch.equal(f4, f3) and torch.equal(dmin4, dmin)
    # position-only targets (3 rows of residual) through both step kernels
    outs = []
    for mode in ("warp", "thread"):
        monkeypatch.setenv("KIN_IK_STEP", mode)
        outs.append(K.ik_solve_device(m, link, joints, dev(tg[:512]), dev(q0[:512]), with_rot=False, iters=30, sscc=sscc, sdf=sdf, margin=0.02))
        monkeypatch.delenv("KIN_IK_STEP")
    for x, y in zip(*outs):
        assert torch.equal(x, y)


def test_staged_ik_solve_is_bitwise_the_single_launch(monkeypatch):
    """Large batches run the device-resident solve in stages over the still-running problems (kin_b200.cu: STAGES):
    a later stage restarts from the best point and the damping of the one before and takes exactly the steps the single
    launch would have taken, so configurations, objectives and iteration counts are identical bit for bit."""
    m, joints, _ = scenes.product_fetch(False)
    mo, jo, _ = scenes.oracle_fetch(False)
    link = K.find_link(m, "gripper_link")
    N = 40000
    tg = _pose_targets(m, joints, link, scenes.random_configs(jo, N, False, seed=83))
    q0 = np.tile(np.array([0.2, 0, 0, 0, 0.5, 0, 0.5, 0]), (N, 1))
    lib = K.load_library()
    monkeypatch.setenv("KIN_IK_STAGES", "0")
    n0 = lib.kin_launch_count()
    qa, fa, ia = K.ik_solve_device(m, link, joints, dev(tg), dev(q0), with_rot=True, iters=40)
    torch.cuda.synchronize()
    assert lib.kin_launch_count() - n0 == 1
    for spec in (None, "3,5,7,9", "39"):
        if spec is None:
            monkeypatch.delenv("KIN_IK_STAGES")
        else:
            monkeypatch.setenv("KIN_IK_STAGES", spec)
        n0 = lib.kin_launch_count()
        qb, fb, ib = K.ik_solve_device(m, link, joints, dev(tg), dev(q0), with_rot=True, iters=40)
        torch.cuda.synchronize()
        assert lib.kin_launch_count() - n0 > 1                          # stages + compactions
        assert torch.equal(qa, qb) and torch.equal(fa, fb) and torch.equal(ia, ib), spec
    assert float((fa < 1e-10).double().mean()) > 0.9 and int(ia.max()) == 40


def test_fridge_demo_example():
    """examples/fridge_demo.py = the reference's fridge_demo.jl (same call sequence, Fetch with the planar base instead of
    PR2): collision-constrained IK into the cabinet, then a 10-waypoint trajectory with margin 0.03 whose straight-line
    initialisation passes through the cabinet wall.  The plan must keep the margin and agree with the same SLSQP driven
    by the oracle's evaluations."""
    import importlib.util
    import sys
    spec = importlib.util.spec_from_file_location("fridge_demo", os.path.join(os.path.dirname(DATA), "examples", "fridge_demo.py"))
    demo = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(demo)
    res, ret, d_goal, d, d_line = demo.main()
    assert res.success and ret.success
    assert d_goal.min() > 0.02 - 1e-6                       # the IK's hard constraint (inverse_kinematics.jl:14-19)
    assert d.min() > 0.03 - 1e-3 and d_line.min() < -0.05   # planned: margin kept; straight line: through the wall
    mo, jo, so = scenes.oracle_fetch(True)
    q_seq = ret.x.reshape(10, 11)
    ret_o = _oracle_plan(so, jo, scenes.oracle_fridge_sdf(), q_seq[0], q_seq[-1], 10, 0.03, 1e-4)
    assert ret_o.success
    np.testing.assert_allclose(ret.fun, ret_o.fun, rtol=2e-3, atol=1e-5)
    print("fridge demo: objective %.6f vs oracle-driven %.6f, iterations %d vs %d" % (ret.fun, ret_o.fun, ret.nit, ret_o.nit))
