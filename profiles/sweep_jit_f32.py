"""Launch-shape sweep of the FP32 mode of the specialised kernels (block, min CTAs/SM): an FP32 thread holds half the
state of an FP64 one, so more CTAs fit per SM if the register allocation is bounded accordingly.
python profiles/sweep_jit_f32.py [log2 N]"""
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
Q = (torch.tensor(lo, device=dev)[:, None] + torch.tensor(hi - lo, device=dev)[:, None] * torch.rand((8, N), generator=g, device=dev, dtype=torch.float64)).float()
Qt = Q.t().reshape(N // 32, 32, 8).permute(0, 2, 1).contiguous()
T = torch.empty((300, N), dtype=torch.float32, device=dev)
J = torch.empty((48, N), dtype=torch.float32, device=dev)
V = torch.empty((16, N), dtype=torch.float32, device=dev)
G = torch.empty((128, N), dtype=torch.float32, device=dev)
fk = np.arange(1, 26, dtype=np.int32)
jac = np.array([K.find_link(m, "gripper_link").id], dtype=np.int32)
ip = C.POINTER(C.c_int32)
stream = torch.cuda.current_stream(dev)


def call(layout, fused, grad_mode):
    c = L.KinCall()
    c.precision, c.layout, c.n, c.q = L.F32, layout, N, (Qt if layout == L.TILED32 else Q).data_ptr()
    c.n_fk_links, c.fk_links, c.T_out = 25, fk.ctypes.data_as(ip), T.data_ptr()
    c.n_jac_links, c.jac_links, c.J_out, c.with_rot = 1, jac.ctypes.data_as(ip), J.data_ptr(), 1
    c.truncation_dist = float("inf")
    c.grad_mode = grad_mode
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


for gm, gname in ((L.GRAD_ANALYTIC, "analytic"), (L.GRAD_FD, "fd")):
    for layout, lname in ((L.SOA, "soa"), (L.TILED32, "tiled")):
        for blk, minb, ksync in ((128, 2, 1), (128, 3, 1), (128, 4, 1), (128, 4, 0), (256, 2, 1), (256, 2, 0), (128, 5, 1), (192, 3, 1), (64, 8, 1)):
            os.environ["KIN_JIT_BLOCK"], os.environ["KIN_JIT_MINB"], os.environ["KIN_JIT_KSYNC"] = str(blk), str(minb), str(ksync)
            try:
                ms, regs, smem, block, grid = timed(call(layout, True, gm))
            except Exception as e:
                print("f32 fused %-8s %-6s block %4d minb %d ksync %d: %s" % (gname, lname, blk, minb, ksync, str(e)[:80]))
                continue
            print("f32 fused %-8s %-6s block %4d minb %d ksync %d: %7.3f ms  %6.1f GB/s (%.3f of 6553.6)  regs %3d smem %6d launch block %d grid %d"
                  % (gname, lname, blk, minb, ksync, ms, 2000 * N / ms / 1e6, 2000 * N / ms / 1e6 / 6553.6, regs, smem, block, grid), flush=True)
