"""CPU ORACLE (test infrastructure, NOT product code) -- the reference's solver-level IK caller timed on host
cores for bench.py's ``cpu_baseline`` leg of the batched-IK row (config 4 of BASELINE.json).

``inverse_kinematics!`` (inverse_kinematics.jl:23-63) = SLSQP on ``f_objective`` with joint-limit bounds.  The
reference uses NLopt's LD_SLSQP; NLopt is not installed, so scipy's SLSQP (the same Kraft routine, and the
reference's own SCIPY back-end in planning.jl:388-394) drives the oracle's ``or_ik_objective``.

``run_ik_baseline`` solves a list of targets on ``n_procs`` worker PROCESSES (``python oracle/callers_cpu.py
<job.npz>``; one mechanism per process because the reference's scratch is not shareable).  Workers
are plain subprocesses (no fork of a CUDA-holding parent, no multiprocessing pool)."""
from __future__ import annotations

import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(job_path):
    sys.path.insert(0, os.path.dirname(_HERE))
    from scipy.optimize import minimize
    from oracle import ref_model as R
    job = np.load(job_path, allow_pickle=True)
    m = R.parse_urdf(str(job["urdf"]), with_base=False)
    joints = [R.find_joint(m, str(n)) for n in job["joint_names"]]
    link = R.find_link(m, str(job["link_name"]))
    with_rot, ftol, q0, targets = bool(job["with_rot"]), float(job["ftol"]), job["q0"], job["targets"]
    lo, hi = [j.lower for j in joints], [j.upper for j in joints]
    bounds = [(a if np.isfinite(a) else None, b if np.isfinite(b) else None) for a, b in zip(lo, hi)]

    collision = "sphere_fixture" in job.files and str(job["sphere_fixture"]) != ""
    if collision:
        # the two-stage driver of inverse_kinematics.jl:1-21: collision-free warm start, then the same problem under
        # IneqConst(sscc, joints, sdf, 1, margin) (planning.jl:55-68), every evaluation by the oracle
        margin = float(job["margin"])
        sscc = R.SweptSphereCollisionChecker(m)
        for sp in json.load(open(str(job["sphere_fixture"])))["links"]:
            R.add_coll_links(sscc, R.find_link(m, sp["link"]), sp["centers"], [sp["radius"]] * len(sp["centers"]))
        fr = R.parse_urdf(str(job["obstacle_urdf"]), with_base=True)
        R.set_joint_angles(fr, [R.find_joint(fr, "door_joint")], [float(v) for v in job["obstacle_state"]])
        sdf = R.UnionSDF(fr)
        cons = [{"type": "ineq", "fun": lambda x: R.ineq_const(sscc, joints, sdf, x, 1, margin)[0],
                 "jac": lambda x: R.ineq_const(sscc, joints, sdf, x, 1, margin)[1][0].T}]

    def solve(T):
        f = lambda x: R.ik_objective(m, link, joints, x, T, with_rot)
        opts = {"ftol": ftol, "maxiter": 200}
        r = minimize(f, q0, jac=True, method="SLSQP", bounds=bounds, options=opts)
        if collision:
            r1 = minimize(f, r.x, jac=True, method="SLSQP", bounds=bounds, constraints=cons, options=opts)
            r1.nfev += r.nfev
            r1.dmin = float(R.ineq_const(sscc, joints, sdf, r1.x, 1, margin)[0].min() + margin)
            return r1
        r.dmin = float("inf")
        return r
    solve(targets[0])                                   # warm-up
    t0 = time.perf_counter()
    res = [solve(T) for T in targets]
    dt = time.perf_counter() - t0
    print(json.dumps({"seconds": dt, "n": len(targets), "f": [float(r.fun) for r in res], "nfev": [int(r.nfev) for r in res],
                      "dmin": [r.dmin for r in res]}))


def run_ik_baseline(urdf, joint_names, link_name, targets, q0, with_rot=True, ftol=1e-10, n_procs=None, timeout=300,
                    sphere_fixture="", obstacle_urdf="", obstacle_state=(), margin=0.02):
    """targets: (n, 4, 4).  -> dict(targets_per_s, seconds (slowest worker), n, procs, fraction_objective_below_1e-6, mean_evals).
    With ``sphere_fixture`` (data/fetch_spheres.json) and ``obstacle_urdf`` (+ its state: door angle, base x y theta) the
    collision-constrained two-stage solve of inverse_kinematics.jl:1-21 runs instead (margin as there)."""
    n_procs = max(1, min(n_procs or os.cpu_count() or 1, len(targets)))
    targets = np.asarray(targets, dtype=np.float64)
    chunks = np.array_split(np.arange(len(targets)), n_procs)
    procs = []
    with tempfile.TemporaryDirectory() as tmp:
        for k, idx in enumerate(chunks):
            path = os.path.join(tmp, "job%d.npz" % k)
            np.savez(path, urdf=urdf, joint_names=np.array(joint_names), link_name=link_name, with_rot=with_rot, ftol=ftol,
                     q0=np.asarray(q0, dtype=np.float64), targets=targets[idx], sphere_fixture=sphere_fixture,
                     obstacle_urdf=obstacle_urdf, obstacle_state=np.asarray(obstacle_state, dtype=np.float64), margin=margin)
            procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), path], stdout=subprocess.PIPE,
                                          stderr=subprocess.PIPE, text=True, env=dict(os.environ, OMP_NUM_THREADS="1")))
        outs = []
        for p in procs:
            try:
                so, se = p.communicate(timeout=timeout)
            except subprocess.TimeoutExpired:
                p.kill()
                raise RuntimeError("IK baseline worker timed out")
            if p.returncode != 0:
                raise RuntimeError("IK baseline worker failed:\n" + se[-2000:])
            outs.append(json.loads(so.strip().splitlines()[-1]))
    dt = max(o["seconds"] for o in outs)
    f = np.concatenate([o["f"] for o in outs])
    ev = np.concatenate([o["nfev"] for o in outs])
    dmin = np.concatenate([o["dmin"] for o in outs])
    return {"targets_per_s": len(targets) / dt, "seconds": dt, "n": int(len(targets)), "procs": n_procs,
            "fraction_objective_below_1e-6": float((f < 1e-6).mean()), "mean_evals": float(ev.mean()),
            "fraction_reached_and_margin_kept": float(((f < 1e-6) & (dmin >= margin - 1e-5)).mean())}


if __name__ == "__main__":
    _worker(sys.argv[1])
