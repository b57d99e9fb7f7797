"""Mid-size batches (2 Ki .. 32 Ki configurations, what a planner with a few thousand waypoints or the active list of the
batched IK evaluates): the interpreting kernel against the specialised one.   python profiles/sweep_midsize.py"""
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
sdf = scene_fetch.product_fridge_sdf()
K.set_joint_angles(m, joints, torch.zeros((1, 8), dtype=torch.float64, device=dev))
K.compute_coll_dists(sscc, joints, sdf)
dm = device_model(m)
lib = L.lib()
lo, hi = scene_fetch.joint_limits(joints)
NMAX = 1 << 16
g = torch.Generator(device=dev).manual_seed(0)
Q = torch.tensor(lo, device=dev)[:, None] + torch.tensor(hi - lo, device=dev)[:, None] * torch.rand((8, NMAX), generator=g, device=dev, dtype=torch.float64)
T = torch.empty((300, NMAX), dtype=torch.float64, device=dev)
J = torch.empty((48, NMAX), dtype=torch.float64, device=dev)
V = torch.empty((16, NMAX), dtype=torch.float64, device=dev)
G = torch.empty((128, NMAX), dtype=torch.float64, device=dev)
fk = np.arange(1, 26, dtype=np.int32)
jac = np.array([K.find_link(m, "gripper_link").id], dtype=np.int32)
ip = C.POINTER(C.c_int32)
stream = torch.cuda.current_stream(dev)


def call(n, fused, trunc):
    c = L.KinCall()
    c.precision, c.layout, c.n, c.q, c.batch_stride = L.F64, L.SOA, n, Q.data_ptr(), NMAX
    c.n_fk_links, c.fk_links, c.T_out = 25, fk.ctypes.data_as(ip), T.data_ptr()
    c.n_jac_links, c.jac_links, c.J_out, c.with_rot = 1, jac.ctypes.data_as(ip), J.data_ptr(), 1
    c.truncation_dist = trunc
    if fused:
        c.vals_out, c.grads_out = V.data_ptr(), G.data_ptr()
    c.stream = stream.cuda_stream
    return c


def timed(c, reps=50):
    for _ in range(8):
        L.check(lib.kin_eval(dm.h, C.byref(c)))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        L.check(lib.kin_eval(dm.h, C.byref(c)))
    b.record(stream)
    torch.cuda.synchronize()
    blk = C.c_int32()
    L.check(lib.kin_query_launch(dm.h, C.byref(c), None, None, C.byref(blk), None))
    return a.elapsed_time(b) / reps * 1e3, blk.value


modes = (("interpreting", {"KIN_DISABLE_JIT": "1"}),
         ("warp/config", {"KIN_FORCE_JIT": "1", "KIN_JIT_WARP_MAX": "1000000"}),
         ("thread/config", {"KIN_FORCE_JIT": "1", "KIN_JIT_WARP_MAX": "0"}),
         ("thread/config, no input batching", {"KIN_FORCE_JIT": "1", "KIN_JIT_WARP_MAX": "0", "KIN_JIT_QBATCH": "0"}))
for fused, trunc, name in ((True, float("inf"), "fused"), (False, float("inf"), "fk+jac")):
    for n in (1, 64, 512, 1024, 2048, 4096, 16384, 32768, 65536):
        row = []
        for label, env in modes:
            if label.startswith("warp") and n > 4096:
                continue
            if label.endswith("batching") and fused:
                continue
            for k in ("KIN_DISABLE_JIT", "KIN_FORCE_JIT", "KIN_JIT_WARP_MAX", "KIN_JIT_QBATCH"):
                os.environ.pop(k, None)
            os.environ.update(env)
            us, blk = timed(call(n, fused, trunc))
            row.append("%s %6.1f us" % (label, us))
        print("%-7s n %6d: %s" % (name, n, "   ".join(row)), flush=True)
