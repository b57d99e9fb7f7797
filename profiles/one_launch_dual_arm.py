"""A few launches of one large fused kin_eval call on the dual-arm model with the planar base (18 columns, 37 links, 19
spheres, 3 boxes; bench.py's dual_arm row), for ncu:  python profiles/one_launch_dual_arm.py [log2 N] [coll]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kinematics_jl_b200 as K  # noqa: E402
from kinematics_jl_b200 import lib as L  # noqa: E402
from kinematics_jl_b200.device import device_model  # noqa: E402
import scene_fetch  # noqa: E402

N = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
coll_only = len(sys.argv) > 2 and sys.argv[2] == "coll"
dev = torch.device("cuda", 0)
m, joints, sscc, sdf = scene_fetch.product_dual_arm(True)
nd, S, nl = len(joints) + 3, len(sscc.sphere_radii), len(m.links)
K.set_joint_angles(m, joints, torch.zeros((1, nd), dtype=torch.float64, device=dev))
K.compute_coll_dists(sscc, joints, sdf)
dm = device_model(m)
lib = L.lib()
g = torch.Generator(device=dev).manual_seed(0)
Q = 2.0 * torch.rand((nd, N), generator=g, device=dev, dtype=torch.float64) - 1.0
T = torch.empty((12 * nl, N), dtype=torch.float64, device=dev)
V = torch.empty((S, N), dtype=torch.float64, device=dev)
G = torch.empty((S * nd, N), dtype=torch.float64, device=dev)
fk = np.array([l.id for l in m.links], dtype=np.int32)
c = L.KinCall()
c.precision, c.layout, c.n, c.q = L.F64, L.SOA, N, Q.data_ptr()
if not coll_only:
    c.n_fk_links, c.fk_links, c.T_out = nl, fk.ctypes.data_as(C.POINTER(C.c_int32)), T.data_ptr()
c.truncation_dist = float("inf")
c.vals_out, c.grads_out = V.data_ptr(), G.data_ptr()
c.stream = torch.cuda.current_stream(dev).cuda_stream
for _ in range(3):
    L.check(lib.kin_eval(dm.h, C.byref(c)))
torch.cuda.synchronize()
print("ok", float(V.sum()))
