"""AoS (the reference-native (n_dof, N) layout) through the model-specialised large-batch kernel, whose outputs are staged
through shared memory and written record-wise by whole warps (kin_gen_skeleton.cuh: aos_flush), against the
interpreting kernel (KIN_DISABLE_JIT) and against SoA.   python profiles/sweep_aos.py [log2 N]"""
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

N = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 22)
dev = torch.device("cuda", 0)
m, joints, sscc = scene_fetch.product_fetch(False)
sdf = scene_fetch.product_fridge_sdf()
K.set_joint_angles(m, joints, torch.zeros((1, 8), dtype=torch.float64, device=dev))
K.compute_coll_dists(sscc, joints, sdf)
dm = device_model(m)
lib = L.lib()
lo, hi = scene_fetch.joint_limits(joints)
g = torch.Generator(device=dev).manual_seed(0)
Qs = torch.tensor(lo, device=dev)[:, None] + torch.tensor(hi - lo, device=dev)[:, None] * torch.rand((8, N), generator=g, device=dev, dtype=torch.float64)
Qa = Qs.t().contiguous()
T = torch.empty(300 * N, dtype=torch.float64, device=dev)
J = torch.empty(48 * N, dtype=torch.float64, device=dev)
V = torch.empty(16 * N, dtype=torch.float64, device=dev)
G = torch.empty(128 * N, dtype=torch.float64, device=dev)
fk = np.arange(1, 26, dtype=np.int32)
jac = np.array([K.find_link(m, "gripper_link").id], dtype=np.int32)
ip = C.POINTER(C.c_int32)
stream = torch.cuda.current_stream(dev)


def call(layout, fused, coll_only=False):
    c = L.KinCall()
    c.precision, c.layout, c.n, c.q = L.F64, layout, N, (Qa if layout == L.AOS else Qs).data_ptr()
    if not coll_only:
        c.n_fk_links, c.fk_links, c.T_out = 25, fk.ctypes.data_as(ip), T.data_ptr()
        c.n_jac_links, c.jac_links, c.J_out, c.with_rot = 1, jac.ctypes.data_as(ip), J.data_ptr(), 1
    c.truncation_dist = float("inf")
    if fused:
        c.vals_out, c.grads_out = V.data_ptr(), G.data_ptr()
    c.stream = stream.cuda_stream
    return c


def timed(c, reps=10):
    for _ in range(3):
        L.check(lib.kin_eval(dm.h, C.byref(c)))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        L.check(lib.kin_eval(dm.h, C.byref(c)))
    b.record(stream)
    torch.cuda.synchronize()
    regs, smem, block, grid = (C.c_int32() for _ in range(4))
    L.check(lib.kin_query_launch(dm.h, C.byref(c), C.byref(regs), C.byref(smem), C.byref(block), C.byref(grid)))
    return a.elapsed_time(b) / reps, regs.value, smem.value, block.value, grid.value


for name, fused, coll_only, bytes_cfg in (("fkj", False, False, 2848), ("fused", True, False, 4000), ("coll", True, True, 1216)):
    for lname, layout, env in (("soa", L.SOA, {}), ("aos", L.AOS, {}), ("aos-interp", L.AOS, {"KIN_DISABLE_JIT": "1"}),
                               ("aos 256x1", L.AOS, {"KIN_JIT_BLOCK": "256", "KIN_JIT_MINB": "1"}),
                               ("aos nosync", L.AOS, {"KIN_JIT_KSYNC": "0"})):
        for k, v in env.items():
            os.environ[k] = v
        try:
            ms, regs, smem, block, grid = timed(call(layout, fused, coll_only))
            print("%-6s %-11s %7.3f ms  %6.1f GB/s (%.3f of 6553.6)  regs %3d smem %6d launch block %d grid %d"
                  % (name, lname, ms, bytes_cfg * N / ms / 1e6, bytes_cfg * N / ms / 1e6 / 6553.6, regs, smem, block, grid), flush=True)
        except Exception as e:
            print(name, lname, str(e)[:100])
        for k in env:
            del os.environ[k]
