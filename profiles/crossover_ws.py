#!/usr/bin/env python
"""Batch-size crossover between kin_eval_kernel and the warp-specialised kernel (fused, SoA, FP64)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import kinematics_jl_b200 as K
from kinematics_jl_b200 import lib as L
from kinematics_jl_b200.device import device_model, current_q, evaluate
import scenes
m, joints, sscc = scenes.product_fetch(False)
fridge = K.parse_urdf(os.path.join(scenes.DATA, "fridge.urdf"), with_base=True)
K.set_joint_angles(fridge, [K.find_joint(fridge, "door_joint")], scenes.FRIDGE_STATE)
sdf = K.UnionSDF(fridge)
mo, jo, so = scenes.oracle_fetch(False)
for n in (8192, 16384, 32768, 49152, 65536, 131072, 262144, 1048576):
    q = torch.as_tensor(scenes.random_configs(jo, n, False, seed=3), device="cuda")
    K.set_joint_angles(m, joints, q)
    K.compute_coll_dists(sscc, joints, sdf)
    dm = device_model(m)
    Q, ql, N = current_q(m)
    kw = dict(layout=L.SOA, fk_links=[l.id for l in m.links[:25]], jac_links=[K.find_link(m, "gripper_link").id],
              with_rot=True, collision=True, launch_info=True)
    res = {}
    for label in ("ws", "classic"):
        os.environ.pop("KIN_DISABLE_WS", None); os.environ.pop("KIN_FORCE_WS", None)
        os.environ["KIN_FORCE_WS" if label == "ws" else "KIN_DISABLE_WS"] = "1"
        out = evaluate(dm, Q, ql, N, **kw); torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out = evaluate(dm, Q, ql, N, **kw); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res[label] = (np.median(ts), out["launch"]["block"])
    print("n=%8d  ws %.4f ms (block %d)   classic %.4f ms (block %d)" % (n, res["ws"][0], res["ws"][1], res["classic"][0], res["classic"][1]), flush=True)
