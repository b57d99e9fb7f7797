"""A few launches of one large kin_eval call, for ncu:  python profiles/one_launch.py {soa|aos|tiled} {fkj|fused|coll} [log2 N] [f32]
(f32: the optional FP32 mode with the analytic gradient, as bench.py's fused_fp32_mode row)"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kinematics_jl_b200 as K  # noqa: E402
from kinematics_jl_b200 import lib as L  # noqa: E402
from kinematics_jl_b200.device import device_model, tile32  # noqa: E402
import scene_fetch  # noqa: E402

layout = {"soa": L.SOA, "aos": L.AOS, "tiled": L.TILED32}[sys.argv[1]]
what = sys.argv[2]
N = 1 << (int(sys.argv[3]) if len(sys.argv) > 3 else 22)
f32 = len(sys.argv) > 4 and sys.argv[4] == "f32"
dt = torch.float32 if f32 else torch.float64
dev = torch.device("cuda", 0)
m, joints, sscc = scene_fetch.product_fetch(False)
sdf = scene_fetch.product_fridge_sdf()
K.set_joint_angles(m, joints, torch.zeros((1, 8), dtype=torch.float64, device=dev))
K.compute_coll_dists(sscc, joints, sdf)
dm = device_model(m)
lib = L.lib()
lo, hi = scene_fetch.joint_limits(joints)
g = torch.Generator(device=dev).manual_seed(0)
Qa = torch.tensor(lo, device=dev) + torch.tensor(hi - lo, device=dev) * torch.rand((N, 8), generator=g, device=dev, dtype=torch.float64)
Qa = Qa.to(dt)
Q = Qa if layout == L.AOS else (Qa.t().contiguous() if layout == L.SOA else tile32(Qa))
T = torch.empty(300 * N, dtype=dt, device=dev)
J = torch.empty(48 * N, dtype=dt, device=dev)
V = torch.empty(16 * N, dtype=dt, device=dev)
G = torch.empty(128 * N, dtype=dt, device=dev)
fk = np.arange(1, 26, dtype=np.int32)
jac = np.array([K.find_link(m, "gripper_link").id], dtype=np.int32)
ip = C.POINTER(C.c_int32)
c = L.KinCall()
c.precision, c.layout, c.n, c.q = (L.F32 if f32 else L.F64), layout, N, Q.data_ptr()
if f32:
    c.grad_mode = L.GRAD_ANALYTIC
if what != "coll":
    c.n_fk_links, c.fk_links, c.T_out = 25, fk.ctypes.data_as(ip), T.data_ptr()
    c.n_jac_links, c.jac_links, c.J_out, c.with_rot = 1, jac.ctypes.data_as(ip), J.data_ptr(), 1
c.truncation_dist = float("inf")
if what != "fkj":
    c.vals_out, c.grads_out = V.data_ptr(), G.data_ptr()
c.stream = torch.cuda.current_stream(dev).cuda_stream
for _ in range(5):
    L.check(lib.kin_eval(dm.h, C.byref(c)))
torch.cuda.synchronize()
print("done", sys.argv[1:], N)
