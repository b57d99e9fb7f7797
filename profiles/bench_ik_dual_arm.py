"""Batched IK on the dual-arm model (data/dual_arm.urdf; 15 columns, 18 with the planar base): pose-only solve
(generated one-launch kernel, staged) and the collision-constrained solve (kin_eval + run-time-sized step kernel pairs)
over 2^k reachable left-tool targets.   python profiles/bench_ik_dual_arm.py [log2 n]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kinematics_jl_b200 as K  # noqa: E402
import scene_fetch  # noqa: E402

N = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 18)
dev = torch.device("cuda", 0)
for with_base in ((True,) if os.environ.get("IK_BASE_ONLY") else (False, True)):
    m, joints, sscc, sdf = scene_fetch.product_dual_arm(with_base)
    link = K.find_link(m, "l_tool")
    nd = len(joints) + (3 if with_base else 0)
    lo = torch.tensor([j.lower_limit for j in joints], device=dev, dtype=torch.float64)
    hi = torch.tensor([j.upper_limit for j in joints], device=dev, dtype=torch.float64)
    g = torch.Generator(device=dev).manual_seed(3)
    qt = lo + (hi - lo) * (0.2 + 0.6 * torch.rand((N, len(joints)), generator=g, device=dev, dtype=torch.float64))
    if with_base:
        qt = torch.cat([qt, 2 * torch.rand((N, 3), generator=g, device=dev, dtype=torch.float64) - 1], dim=1)
    K.set_joint_angles(m, joints, qt)
    T = K.get_transform(m, link)                               # (N, 3, 4)
    Rm = T[:, :, :3]
    yaw = torch.atan2(Rm[:, 1, 0], Rm[:, 0, 0])
    pitch = torch.atan2(-Rm[:, 2, 0], torch.sqrt(Rm[:, 2, 1] ** 2 + Rm[:, 2, 2] ** 2))
    roll = torch.atan2(Rm[:, 2, 1], Rm[:, 2, 2])
    tg = torch.cat([T[:, :, 3], roll[:, None], pitch[:, None], yaw[:, None]], dim=1).contiguous()
    q0 = torch.zeros((N, nd), device=dev, dtype=torch.float64)
    for name, kw in (("pose only", {}), ("collision-constrained (margin 0.02)", dict(sscc=sscc, sdf=sdf, margin=0.02, return_dmin=True))):
        for rep in range(int(os.environ.get("IK_REPS", 3))):                                   # the first repetition compiles the kernels
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = K.inverse_kinematics_batch(m, link, joints, tg, q0, with_rot=True, iters=40, **kw)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        f = out[1]
        good = f <= 1e-6
        if kw:
            good &= out[2] >= 0.02 - 1e-5
        print("%d columns, %s: 2^%d targets in %.1f ms (%.2e targets/s), %.1f %% solved" %
              (nd, name, int(np.log2(N)), dt * 1e3, N / dt, 100 * float(good.double().mean())), flush=True)
