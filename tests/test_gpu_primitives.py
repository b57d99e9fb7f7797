"""GPU parity of the sphere / cylinder SDF primitives (EXTENSION; the reference has boxes only) against the
oracle's restatement: point queries, the fused collision call through every kernel (interpreting, model-specialised,
one-warp-per-configuration; all three layouts), and the URDF route.  Tolerances as for boxes: values 1e-12, argmin
exact, closed-form gradient 1e-12, forward-difference gradient 1e-7 absolute (FD noise)."""
import os

import numpy as np
import pytest
import torch

import kinematics_jl_b200 as K
from kinematics_jl_b200 import lib as L
from kinematics_jl_b200.device import current_q, device_model, evaluate
from oracle import ref_model as R
from conftest import DATA, GOLDEN
import scenes
from test_gpu_parity import ATOL, RTOL, dev, host
from test_primitives_cpu import _pose, _rot

pytestmark = pytest.mark.gpu


def _mixed_union():
    specs = [("box", _pose([0.9, 0.1, 0.6], _rot([1, 2, 3], 0.4)), [0.4, 0.3, 0.8]),
             ("sphere", _pose([0.7, -0.35, 0.9]), [0.2]),
             ("cylinder", _pose([0.6, 0.3, 0.5], _rot([1, 0, 1], 1.1)), [0.12, 0.9]),
             ("cylinder", _pose([0.85, -0.1, 1.25]), [0.3, 0.06])]
    mk = {"box": K.BoxSDF, "sphere": K.SphereSDF, "cylinder": K.CylinderSDF}
    mo = {"box": R.BoxSDF, "sphere": R.SphereSDF, "cylinder": R.CylinderSDF}
    return (K.UnionSDF([mk[k](K.Transform(T), *sz) if k != "box" else K.BoxSDF(K.Transform(T), sz) for k, T, sz in specs]),
            R.UnionSDF([mo[k](T, *sz) if k != "box" else R.BoxSDF(T, sz) for k, T, sz in specs]))


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_primitive_points_vs_oracle(dtype):
    sdf, so = _mixed_union()
    rng = np.random.default_rng(5)
    pts = np.array([0.75, 0.0, 0.8]) + (rng.random((6000, 3)) - 0.5) * 1.6
    # special places: the sphere's centre (gradient undefined: both sides return the primitive's +x axis), points on
    # the axes through it, the disc-shaped cylinder's centre, its axis inside and outside, a point off its rim
    special = [[0.7, -0.35, 0.9], [0.7, -0.35, 1.15], [0.95, -0.35, 0.9], [0.85, -0.1, 1.25], [0.85, -0.1, 1.5], [1.3, -0.1, 1.4],
               [0.85, -0.1, 1.27]]
    if dtype == torch.float64:
        pts = np.concatenate([pts, np.array(special)])
    vals, am = sdf(dev(pts, dtype), return_argmin=True)
    g_fd = sdf.gradient(dev(pts, dtype))
    g_an = sdf.gradient(dev(pts, dtype), grad_mode=K.GRAD_ANALYTIC)
    v_ref, am_ref = np.zeros(len(pts)), np.zeros(len(pts), dtype=np.int32)
    gf_ref, ga_ref = np.zeros_like(pts), np.zeros_like(pts)
    for i, p in enumerate(pts):
        v_ref[i] = so(p)
        am_ref[i] = so.argmin
        gf_ref[i], ga_ref[i] = so.gradient(p), so.gradient(p, analytic=True)
    assert set(am_ref) == {1, 2, 3, 4} and (v_ref < 0).sum() > 100          # every kind wins somewhere, inside and outside
    if dtype == torch.float64:
        np.testing.assert_allclose(host(vals), v_ref, rtol=RTOL, atol=1e-14)
        assert np.array_equal(am.cpu().numpy(), am_ref)
        np.testing.assert_allclose(host(g_an), ga_ref, rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(host(g_fd), gf_ref, rtol=0, atol=1e-7)
    else:                                                                     # FP32 mode: 1e-5 (north_star)
        np.testing.assert_allclose(host(vals), v_ref, rtol=0, atol=1e-5)
        same = am.cpu().numpy() == am_ref
        assert same.mean() > 0.995
        np.testing.assert_allclose(host(g_an)[same], ga_ref[same], rtol=0, atol=2e-4)     # next to a kink the face can flip


def _scene():
    m, joints, sscc = scenes.product_fetch(False)
    mo, jo, so = scenes.oracle_fetch(False)
    sdf, sdf_o = _mixed_union()
    return m, joints, sscc, sdf, mo, jo, so, sdf_o


@pytest.mark.parametrize("layout", ["aos", "soa"])
@pytest.mark.parametrize("truncation", [np.inf, 0.08])
def test_collision_against_mixed_primitives_vs_oracle(layout, truncation):
    m, joints, sscc, sdf, mo, jo, so, sdf_o = _scene()
    q = scenes.random_configs(jo, 555, False, seed=41, zeros_every=100)
    Q = dev(q)
    K.set_joint_angles(m, joints, Q.t().contiguous().t() if layout == "soa" else Q)
    for scratch, scratch_o in ((K.SCRATCH_REFERENCE, R.SCRATCH_REFERENCE), (K.SCRATCH_CLEAN, R.SCRATCH_CLEAN)):
        for gm, gm_o, tol in ((K.GRAD_FD, R.GRAD_FD, 1e-7), (K.GRAD_ANALYTIC, R.GRAD_ANALYTIC, 1e-11)):
            vals, grads, am = K.compute_coll_dists_and_grads(sscc, joints, sdf, truncation_dist=truncation, grad_mode=gm,
                                                             scratch_mode=scratch, return_argmin=True)
            v_ref, g_ref, am_ref = R.batch_collision(so, jo, sdf_o, q, truncation, gm_o, scratch_o)
            assert set(np.unique(am_ref)) == {1, 2, 3, 4}
            np.testing.assert_allclose(host(vals), v_ref, rtol=RTOL, atol=ATOL)
            assert np.array_equal(am.cpu().numpy(), am_ref)
            assert np.array_equal(host(vals) == truncation, v_ref == truncation)
            np.testing.assert_allclose(host(grads), g_ref.transpose(0, 2, 1), rtol=tol if tol < 1e-9 else 0, atol=tol)


@pytest.mark.parametrize("layout", [L.SOA, L.TILED32, L.AOS])
def test_every_kernel_agrees_bitwise_on_mixed_primitives(layout, monkeypatch):
    """Interpreting kernel, model-specialised kernel (KPRIMS = 1 in its generated configuration) and the
    one-warp-per-configuration kernel evaluate a table with sphere / cylinder rows to the same bits."""
    m, joints, sscc, sdf, mo, jo, so, sdf_o = _scene()
    gl = K.find_link(m, "gripper_link").id

    def run(n, mode, warp_max=None, **kw):
        for k in ("KIN_DISABLE_JIT", "KIN_FORCE_JIT", "KIN_JIT_WARP_MAX"):
            monkeypatch.delenv(k, raising=False)
        if mode:
            monkeypatch.setenv(mode, "1")
        if warp_max is not None:
            monkeypatch.setenv("KIN_JIT_WARP_MAX", str(warp_max))
        q = scenes.random_configs(jo, n, False, seed=43, zeros_every=50)
        K.set_joint_angles(m, joints, dev(q))
        K.compute_coll_dists(sscc, joints, sdf)
        Qc, ql, N = current_q(m)
        o = evaluate(device_model(m), Qc, ql, N, layout=layout, fk_links=[gl], jac_links=[gl], collision=True, want_argmin=True,
                     launch_info=True, **kw)
        torch.cuda.synchronize()
        return q, o

    for kw in (dict(truncation_dist=np.inf, grad_mode=L.GRAD_FD, scratch_mode=L.SCRATCH_REFERENCE),
               dict(truncation_dist=0.08, grad_mode=L.GRAD_ANALYTIC, scratch_mode=L.SCRATCH_CLEAN, vals_offset=0.03)):
        q, a = run(40013, "KIN_DISABLE_JIT", **kw)
        _, b = run(40013, None, **kw)
        assert a["launch"]["block"] > 0 and b["launch"]["block"] < 0, L.lib().kin_jit_status()
        for key in ("T", "J", "vals", "grads", "argmin"):
            assert torch.equal(a[key], b[key]), (key, kw)
        _, c = run(77, "KIN_DISABLE_JIT", **kw)
        _, d = run(77, "KIN_FORCE_JIT", **kw)                      # small batch, one thread per configuration
        _, e = run(77, "KIN_FORCE_JIT", warp_max=2048, **kw)       # small batch, one warp per configuration
        assert c["launch"]["block"] > 0 and d["launch"]["block"] < 0 and e["launch"]["block"] < 0
        assert d["launch"]["grid"] == 1 and e["launch"]["grid"] == 20
        for key in ("T", "J", "vals", "grads", "argmin"):
            assert torch.equal(c[key], d[key]) and torch.equal(c[key], e[key]), (key, kw)
    v_ref, g_ref, am_ref = R.batch_collision(so, jo, sdf_o, q[:300], 0.08, R.GRAD_ANALYTIC, R.SCRATCH_CLEAN)
    np.testing.assert_allclose(host(b["vals"][:300]) + 0.03, v_ref, rtol=RTOL, atol=ATOL)
    assert np.array_equal(b["argmin"][:300].cpu().numpy(), am_ref)


def test_primitives_from_urdf_and_box_only_tables_keep_their_kernel():
    obst = K.parse_urdf(os.path.join(GOLDEN, "prims_obstacle.urdf"), with_base=True)
    oo = R.parse_urdf(os.path.join(GOLDEN, "prims_obstacle.urdf"), with_base=True)
    state = [0.7, 0.9, -0.2, 0.3]
    K.set_joint_angles(obst, [K.find_joint(obst, "arm_joint")], state)
    R.set_joint_angles(oo, [R.find_joint(oo, "arm_joint")], state)
    assert len(K.UnionSDF(K.parse_urdf(os.path.join(GOLDEN, "prims_obstacle.urdf"))).sdfs) == 1     # reference behaviour
    sdf, sdf_o = K.UnionSDF(obst, primitives=True), R.UnionSDF(oo, primitives=True)
    poses, sizes, kinds = sdf.world_primitives()
    assert list(kinds) == [0, 2, 2, 1] == sdf_o.kinds
    np.testing.assert_allclose(poses, np.stack(sdf_o.poses), rtol=0, atol=1e-15)
    m, joints, sscc = scenes.product_fetch(False)
    mo, jo, so = scenes.oracle_fetch(False)
    q = scenes.random_configs(jo, 300, False, seed=47)
    K.set_joint_angles(m, joints, dev(q))
    vals, grads, am = K.compute_coll_dists_and_grads(sscc, joints, sdf, return_argmin=True)
    v_ref, g_ref, am_ref = R.batch_collision(so, jo, sdf_o, q)
    np.testing.assert_allclose(host(vals), v_ref, rtol=RTOL, atol=ATOL)
    assert np.array_equal(am.cpu().numpy(), am_ref)
    np.testing.assert_allclose(host(grads), g_ref.transpose(0, 2, 1), rtol=0, atol=1e-7)
    # moving the obstacle keeps the table shape: rows are rewritten in place (kin_model_set_primitives)
    K.set_joint_angles(obst, [K.find_joint(obst, "arm_joint")], [-1.1, 0.9, -0.2, 0.3])
    R.set_joint_angles(oo, [R.find_joint(oo, "arm_joint")], [-1.1, 0.9, -0.2, 0.3])
    vals2 = K.compute_coll_dists(sscc, joints, sdf)
    v_ref2, _, _ = R.batch_collision(so, jo, R.UnionSDF(oo, primitives=True), q, with_grads=False)
    np.testing.assert_allclose(host(vals2), v_ref2, rtol=RTOL, atol=ATOL)
    assert np.abs(v_ref2 - v_ref).max() > 1e-3
