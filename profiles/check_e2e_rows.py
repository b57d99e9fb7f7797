"""kin_eval_host against kin_eval on the same inputs, every element: python profiles/check_e2e_rows.py [log2 N]"""
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

N = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 21)
dev = torch.device("cuda", 0)
m, joints, sscc = scene_fetch.product_fetch(False)
K.set_joint_angles(m, joints, torch.zeros((1, 8), dtype=torch.float64, device=dev))
dm = device_model(m)
lib = L.lib()
lo, hi = scene_fetch.joint_limits(joints)
g = torch.Generator(device=dev).manual_seed(0)
Q = torch.tensor(lo, device=dev)[:, None] + torch.tensor(hi - lo, device=dev)[:, None] * torch.rand((8, N), generator=g, device=dev, dtype=torch.float64)
T = torch.empty((300, N), dtype=torch.float64, device=dev)
J = torch.empty((48, N), dtype=torch.float64, device=dev)
fk = np.arange(1, 26, dtype=np.int32)
jac = np.array([K.find_link(m, "gripper_link").id], dtype=np.int32)
ip = C.POINTER(C.c_int32)


def call(q, t, j):
    c = L.KinCall()
    c.precision, c.layout, c.n, c.q = L.F64, L.SOA, N, q
    c.n_fk_links, c.fk_links, c.T_out = 25, fk.ctypes.data_as(ip), t
    c.n_jac_links, c.jac_links, c.J_out, c.with_rot = 1, jac.ctypes.data_as(ip), j, 1
    c.truncation_dist = float("inf")
    c.stream = torch.cuda.current_stream(dev).cuda_stream
    return c


L.check(lib.kin_eval(dm.h, C.byref(call(Q.data_ptr(), T.data_ptr(), J.data_ptr()))))
torch.cuda.synchronize()
qh = Q.cpu().pin_memory()
Th = torch.empty((300, N), dtype=torch.float64).pin_memory()
Jh = torch.empty((48, N), dtype=torch.float64).pin_memory()
for rep in range(4):
    Th.fill_(-1.0)
    Jh.fill_(-1.0)
    L.check(lib.kin_eval_host(dm.h, C.byref(call(qh.data_ptr(), Th.data_ptr(), Jh.data_ptr()))))
    for name, a, b in (("T", Th, T.cpu()), ("J", Jh, J.cpu())):
        bad = (a != b)
        if bad.any():
            rows = torch.nonzero(bad.any(dim=1)).flatten().tolist()
            cols = torch.nonzero(bad.any(dim=0)).flatten()
            print("rep %d %s: %d mismatching elements, rows %s, columns %d..%d (%d distinct), sample host %r device %r" % (
                rep, name, int(bad.sum()), rows[:20], int(cols.min()), int(cols.max()), cols.numel(),
                a[rows[0], int(cols[0])].item(), b[rows[0], int(cols[0])].item()))
        else:
            print("rep %d %s: identical" % (rep, name))
