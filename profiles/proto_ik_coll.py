"""CPU prototype (numpy + the oracle's batch evaluators) of the device-resident collision-constrained batched IK:
augmented-Lagrangian Levenberg-Marquardt on  min |e_pose(q)|^2  s.t.  d_s(q) - margin >= 0, lo <= q <= hi.
Used to choose the algorithm's constants before writing kin_ik_coll_step_kernel; scenario of
tests/test_gpu_callers.py::test_batched_collision_aware_ik (thin box of test/test_planning.jl:23-25)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_model as R   # noqa: E402
import scenes                        # noqa: E402

NT = R.max_threads()


def evaluate(mo, jo, so, sdf, link, q, tg, margin, trunc):
    T = R.batch_fk(mo, jo, q, [link], n_threads=NT)[:, 0]
    J = R.batch_jacobian(mo, jo, q, [link], True, rpy_jac=True, n_threads=NT)[:, 0]      # (N, 6, nd)
    d, G, _ = R.batch_collision(so, jo, sdf, q, truncation_dist=trunc, scratch_mode=R.SCRATCH_CLEAN, n_threads=NT)
    N = q.shape[0]
    e = np.zeros((N, 6))
    e[:, :3] = T[:, :3, 3] - tg[:, :3]
    for n in range(N):
        e[n, 3:] = R.rpy(T[n]) - tg[n, 3:]
    e[:, 3:] = np.remainder(e[:, 3:] + np.pi, 2 * np.pi) - np.pi
    return e, J, d, G


def solve(mo, jo, so, sdf, link, tg, q0, lo, hi, margin=0.02, iters=60, mu=1e3, mult_every=1, verbose=False):
    N, nd = q0.shape
    S = len(so.sphere_links)
    trunc = margin + 0.05
    q = q0.copy()
    lam_m = np.zeros((N, S))                     # multipliers
    damp = np.full(N, 1e-2)

    def merit_parts(e, J, d, G, lam_m):
        psi = np.maximum(0.0, (margin - d) + lam_m / mu)            # (N, S)
        phi = (e * e).sum(1) + mu * (psi * psi).sum(1)
        act = psi > 0
        H = np.einsum("nri,nrj->nij", J, J) + mu * np.einsum("ns,nsi,nsj->nij", act.astype(float), G, G)
        g = np.einsum("nri,nr->ni", J, e) - mu * np.einsum("ns,nsi->ni", psi, G)
        return phi, H, g, psi

    e, J, d, G = evaluate(mo, jo, so, sdf, link, q, tg, margin, trunc)
    phi, H, g, psi = merit_parts(e, J, d, G, lam_m)
    fpose = (e * e).sum(1)
    dmin = d.min(1)
    for it in range(iters):
        fr = ~(((q <= lo + 1e-12) & (g > 0)) | ((q >= hi - 1e-12) & (g < 0)))
        x = np.zeros((N, nd))
        for n in range(N):
            f = fr[n]
            Hn = H[n][np.ix_(f, f)]
            A = Hn + damp[n] * (np.eye(f.sum()) + np.diag(np.diag(Hn)))
            x[n, f] = np.linalg.solve(A, g[n, f])
        qt = np.clip(q - x, lo, hi)
        e_t, J_t, d_t, G_t = evaluate(mo, jo, so, sdf, link, qt, tg, margin, trunc)
        phi_t, H_t, g_t, psi_t = merit_parts(e_t, J_t, d_t, G_t, lam_m)
        ok = phi_t < phi
        damp = np.clip(np.where(ok, damp * 0.3, damp * 4.0), 1e-9, 1e4)
        q[ok] = qt[ok]
        fpose[ok] = (e_t[ok] ** 2).sum(1)
        dmin[ok] = d_t[ok].min(1)
        # multiplier update at the accepted point (inexact AL), then the merit / normal equations under the new multipliers
        upd = ok & ((it % mult_every) == mult_every - 1)
        lam_new = np.where(upd[:, None], mu * psi_t, lam_m)
        phi_n, H_n, g_n, psi_n = merit_parts(e_t, J_t, d_t, G_t, lam_new)
        lam_m = lam_new
        for arr, new in ((phi, phi_n), (H, H_n), (g, g_n)):
            arr[ok] = new[ok]
        if verbose and it % 10 == 9:
            print(it, "reached %.3f clear %.3f both %.3f" % ((fpose < 1e-6).mean(), (dmin > margin - 1e-3).mean(),
                                                              ((fpose < 1e-6) & (dmin > margin - 1e-3)).mean()))
    return q, fpose, dmin


def main():
    mo, jo, so = scenes.oracle_fetch(False)
    link = R.find_link(mo, "gripper_link")
    pose = np.eye(4)
    pose[:3, 3] = [0.4, -0.25, 0.8]
    box = R.BoxSDF(pose, [0.05, 0.05, 0.5])
    N = int(os.environ.get("N", 256))
    rng = np.random.default_rng(5)
    tg = np.zeros((N, 6))
    tg[:, 0], tg[:, 1], tg[:, 2] = rng.uniform(0.55, 0.8, N), rng.uniform(-0.3, 0.3, N), rng.uniform(0.7, 1.1, N)
    q0 = np.tile(np.array([0.2, 0, 0, 0, 0.5, 0, 0.5, 0]), (N, 1))
    lo = np.array([j.lower for j in jo])
    hi = np.array([j.upper for j in jo])
    lo_s = np.where(np.isfinite(lo), lo, -np.pi)
    hi_s = np.where(np.isfinite(hi), hi, np.pi)

    def outcome(q):
        e, J, d, G = evaluate(mo, jo, so, box, link, q, tg, 0.02, np.inf)
        return np.abs(e).max(1) < 1e-3, d.min(1) > -1e-3, d.min(1) > 0.02 - 1e-6

    for mu in [float(x) for x in os.environ.get("MU", "1e2,1e3,1e4").split(",")]:
        # stage 1: pose only (mu irrelevant when no sphere is near: use an empty constraint via huge negative margin)
        q1, _, _ = solve(mo, jo, so, box, link, tg, q0, lo, hi, margin=-10.0, iters=40, mu=mu)
        r0, c0, m0 = outcome(q1)
        q2, fp, dm = solve(mo, jo, so, box, link, tg, q1, lo, hi, margin=0.02, iters=int(os.environ.get("ITERS", 60)), mu=mu,
                           verbose=bool(os.environ.get("V")))
        r1, c1, m1 = outcome(q2)
        print("mu %g: free reached %.3f clear %.3f both %.3f | constrained reached %.3f clear(-1e-3) %.3f margin-ok %.3f reached&clear %.3f reached&margin %.3f"
              % (mu, r0.mean(), c0.mean(), (r0 & c0).mean(), r1.mean(), c1.mean(), m1.mean(), (r1 & c1).mean(), (r1 & m1).mean()))
        # restarts of the failures from random seeds
        rng2 = np.random.default_rng(1)
        good = r1 & m1
        qbest = q2.copy()
        for rs in range(int(os.environ.get("RESTARTS", 3))):
            bad = np.nonzero(~good)[0]
            if bad.size == 0:
                break
            qs = lo_s + (hi_s - lo_s) * rng2.random((bad.size, len(jo)))
            qa, _, _ = solve(mo, jo, so, box, link, tg[bad], qs, lo, hi, margin=-10.0, iters=40, mu=mu)
            qb, _, _ = solve(mo, jo, so, box, link, tg[bad], qa, lo, hi, margin=0.02, iters=int(os.environ.get("ITERS", 60)), mu=mu)
            e, J, d, G = evaluate(mo, jo, so, box, link, qb, tg[bad], 0.02, np.inf)
            okb = (np.abs(e).max(1) < 1e-3) & (d.min(1) > 0.02 - 1e-6)
            qbest[bad[okb]] = qb[okb]
            good[bad[okb]] = True
            print("   restart %d: reached & margin-ok %.3f" % (rs + 1, good.mean()))




def diag():
    mo, jo, so = scenes.oracle_fetch(False)
    link = R.find_link(mo, "gripper_link")
    pose = np.eye(4)
    pose[:3, 3] = [0.4, -0.25, 0.8]
    box = R.BoxSDF(pose, [0.05, 0.05, 0.5])
    N = 256
    rng = np.random.default_rng(5)
    tg = np.zeros((N, 6))
    tg[:, 0], tg[:, 1], tg[:, 2] = rng.uniform(0.55, 0.8, N), rng.uniform(-0.3, 0.3, N), rng.uniform(0.7, 1.1, N)
    q0 = np.tile(np.array([0.2, 0, 0, 0, 0.5, 0, 0.5, 0]), (N, 1))
    lo = np.array([j.lower for j in jo]); hi = np.array([j.upper for j in jo])
    mu = float(os.environ.get("MU", "1e3"))
    q1, _, _ = solve(mo, jo, so, box, link, tg, q0, lo, hi, margin=-10.0, iters=40, mu=mu)
    e, J, d, G = evaluate(mo, jo, so, box, link, q1, tg, 0.02, np.inf)
    err1 = np.abs(e).max(1)
    print("free: err percentiles", np.percentile(err1, [50, 80, 90, 95, 99]))
    q2, fp, dm = solve(mo, jo, so, box, link, tg, q1, lo, hi, margin=0.02, iters=int(os.environ.get("ITERS", 60)), mu=mu)
    e, J, d, G = evaluate(mo, jo, so, box, link, q2, tg, 0.02, np.inf)
    err2 = np.abs(e).max(1)
    print("cons: err percentiles", np.percentile(err2, [30, 50, 60, 70, 80, 90, 95, 99]))
    print("cons: dmin percentiles", np.percentile(d.min(1), [1, 5, 10, 50]))
    bad = np.nonzero((err2 > 1e-3) & (err1 < 1e-3))[0]
    print("lost by the constraint stage:", bad.size, "err:", err2[bad][:20], "dmin", d.min(1)[bad][:20])
    print("at limit:", ((q2[bad] <= lo + 1e-9) | (q2[bad] >= hi - 1e-9)).sum(1)[:20])


if os.environ.get("DIAG"):
    diag()

if __name__ == "__main__" and not os.environ.get("DIAG"):
    main()
