"""Where does the input batching of the FK / Jacobian kernels (grid-wide barriers) start to pay?   python profiles/sweep_qbatch_min.py"""
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

dev = torch.device("cuda", 0)
m, joints, sscc = scene_fetch.product_fetch(False)
K.set_joint_angles(m, joints, torch.zeros((1, 8), dtype=torch.float64, device=dev))
dm = device_model(m)
lib = L.lib()
lo, hi = scene_fetch.joint_limits(joints)
NMAX = 1 << 23
g = torch.Generator(device=dev).manual_seed(0)
Q = torch.tensor(lo, device=dev)[:, None] + torch.tensor(hi - lo, device=dev)[:, None] * torch.rand((8, NMAX), generator=g, device=dev, dtype=torch.float64)
T = torch.empty((300, NMAX), dtype=torch.float64, device=dev)
J = torch.empty((48, NMAX), dtype=torch.float64, device=dev)
fk = np.arange(1, 26, dtype=np.int32)
jac = np.array([K.find_link(m, "gripper_link").id], dtype=np.int32)
ip = C.POINTER(C.c_int32)
stream = torch.cuda.current_stream(dev)
for n in (1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22, 1 << 23):
    row = []
    for label, min_n in (("with input batching", "0"), ("without", str(1 << 40))):
        os.environ["KIN_JIT_QBATCH_MIN_N"] = min_n
        c = L.KinCall()
        c.precision, c.layout, c.n, c.q, c.batch_stride = L.F64, L.SOA, n, Q.data_ptr(), NMAX
        c.n_fk_links, c.fk_links, c.T_out = 25, fk.ctypes.data_as(ip), T.data_ptr()
        c.n_jac_links, c.jac_links, c.J_out, c.with_rot = 1, jac.ctypes.data_as(ip), J.data_ptr(), 1
        c.truncation_dist = float("inf")
        c.stream = stream.cuda_stream
        for _ in range(5):
            L.check(lib.kin_eval(dm.h, C.byref(c)))
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        a.record(stream)
        for _ in range(reps):
            L.check(lib.kin_eval(dm.h, C.byref(c)))
        b.record(stream)
        torch.cuda.synchronize()
        us = a.elapsed_time(b) / reps * 1e3
        row.append("%s %9.1f us (%.3f of HBM)" % (label, us, 2848 * n / us / 1e3 / 6553.6))
    print("n 2^%d: %s" % (n.bit_length() - 1, "   ".join(row)), flush=True)
