#!/usr/bin/env python
"""Fused FK-all + gripper Jacobian + collision on the Fetch WITH the planar base (11 columns), SoA FP64:
warp-specialised kernel vs kin_eval_kernel (KIN_DISABLE_WS=1).  Usage: python profiles/bench_with_base.py [N]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import kinematics_jl_b200 as K
from kinematics_jl_b200 import lib as L
from kinematics_jl_b200.device import device_model, current_q, evaluate
import scenes

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
m, joints, sscc = scenes.product_fetch(True)
fridge = K.parse_urdf(os.path.join(scenes.DATA, "fridge.urdf"), with_base=True)
K.set_joint_angles(fridge, [K.find_joint(fridge, "door_joint")], scenes.FRIDGE_STATE)
sdf = K.UnionSDF(fridge)
mo, jo, so = scenes.oracle_fetch(True)
q = torch.as_tensor(scenes.random_configs(jo, N, True, seed=3), device="cuda")
K.set_joint_angles(m, joints, q)
K.compute_coll_dists(sscc, joints, sdf)
dm = device_model(m)
Q, ql, n = current_q(m)
kw = dict(layout=L.SOA, fk_links=[l.id for l in m.links[:25]], jac_links=[K.find_link(m, "gripper_link").id],
          with_rot=True, collision=True, launch_info=True)
for label, env in (("warp-specialised", None), ("kin_eval_kernel", "1")):
    if env: os.environ["KIN_DISABLE_WS"] = env
    else: os.environ.pop("KIN_DISABLE_WS", None)
    out = evaluate(dm, Q, ql, n, **kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = evaluate(dm, Q, ql, n, **kw); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    byts = 8 * (11 + 12 * 25 + 6 * 11 + 16 + 16 * 11)
    print("%-18s block %d  %.3f ms per %d configs  %.3g configs/s  %.0f GB/s algorithmic" %
          (label, out["launch"]["block"], min(ts), n, n / (min(ts) * 1e-3), byts * n / (min(ts) * 1e-3) / 1e9))
