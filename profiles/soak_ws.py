#!/usr/bin/env python
"""Soak: warp-specialised kernel vs kin_eval_kernel, bitwise, over random batch sizes / modes / layouts, with
back-to-back launches re-using the pool's ring memory.  Usage: python profiles/soak_ws.py [iterations]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import kinematics_jl_b200 as K
from kinematics_jl_b200 import lib as L
from kinematics_jl_b200.device import device_model, current_q, evaluate
import scenes

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(12345)
os.environ["KIN_FORCE_WS"] = "1"
bad = 0
for it in range(iters):
    with_base = bool(rng.integers(0, 2))
    m, joints, sscc = scenes.product_fetch(with_base)
    fridge = K.parse_urdf(os.path.join(scenes.DATA, "fridge.urdf"), with_base=True)
    K.set_joint_angles(fridge, [K.find_joint(fridge, "door_joint")], [float(rng.uniform(0, 2)), 1.2 + float(rng.uniform(-.3, .3)), 0.0, float(rng.uniform(-.5, .5))])
    sdf = K.UnionSDF(fridge)
    mo, jo, so = scenes.oracle_fetch(with_base)
    n = int(rng.choice([1, 127, 128, 129, 18943, 18944, 18945, 37889, int(rng.integers(1, 400000)), int(rng.integers(400000, 3000000))]))
    q = torch.as_tensor(scenes.random_configs(jo, n, with_base, seed=int(rng.integers(1 << 30))), device="cuda")
    K.set_joint_angles(m, joints, q)
    K.compute_coll_dists(sscc, joints, sdf)
    dm = device_model(m)
    Q, ql, N = current_q(m)
    kw = dict(layout=[L.SOA, L.TILED32][int(rng.integers(0, 2))], collision=True, launch_info=True,
              truncation_dist=[np.inf, 0.05, 0.3][int(rng.integers(0, 3))],
              grad_mode=[K.GRAD_FD, K.GRAD_ANALYTIC, K.GRAD_FD_DIRECT][int(rng.integers(0, 3))],
              scratch_mode=[K.SCRATCH_REFERENCE, K.SCRATCH_CLEAN][int(rng.integers(0, 2))], want_argmin=True)
    if rng.integers(0, 3) > 0:
        kw.update(fk_links=[l.id for l in m.links[:25]], jac_links=[K.find_link(m, "gripper_link").id],
                  with_rot=bool(rng.integers(0, 2)), rpy_jac=bool(rng.integers(0, 2)))
    os.environ.pop("KIN_DISABLE_WS", None)
    outs = [evaluate(dm, Q, ql, N, **kw) for _ in range(3)]          # back to back: ring memory is re-used
    junk = torch.empty(int(rng.integers(1, 1 << 22)), device="cuda").normal_()
    torch.cuda.synchronize()
    assert outs[0]["launch"]["block"] == 384, outs[0]["launch"]
    os.environ["KIN_DISABLE_WS"] = "1"
    ref = evaluate(dm, Q, ql, N, **kw)
    torch.cuda.synchronize()
    assert ref["launch"]["block"] != 384
    for o in outs:
        for k in ref:
            if k != "launch" and not torch.equal(o[k].contiguous(), ref[k].contiguous()):
                bad += 1
                print("MISMATCH", it, k, n, with_base, {a: b for a, b in kw.items() if a not in ("fk_links", "jac_links")})
    print("iter %d ok: n=%d base=%s layout=%d smem=%d" % (it, n, with_base, kw["layout"], outs[0]["launch"]["smem_bytes"]), flush=True)
print("mismatches:", bad)
sys.exit(1 if bad else 0)
