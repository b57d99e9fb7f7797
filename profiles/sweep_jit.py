"""Launch-shape sweep of the model-specialised kernels: (block, min CTAs/SM) for the headline FK + Jacobian call and the
fused call, SoA and tiled.  KIN_JIT_BLOCK / KIN_JIT_MINB are read when a kernel is first built and are part of its
cache key, so one process can sweep them.   python profiles/sweep_jit.py [log2 N] [fkj|fused|all|fkg|coll]
(fkg: FK of the gripper link + its Jacobian, 544 B per configuration; coll: collision cost + gradient only, 1216 B)"""
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
what = sys.argv[2] if len(sys.argv) > 2 else "all"
dev = torch.device("cuda", 0)
m, joints, sscc = scene_fetch.product_fetch(False)
sdf = scene_fetch.product_fridge_sdf()
K.set_joint_angles(m, joints, torch.zeros((1, 8), dtype=torch.float64, device=dev))
K.compute_coll_dists(sscc, joints, sdf)
dm = device_model(m)
lib = L.lib()
lo, hi = scene_fetch.joint_limits(joints)
g = torch.Generator(device=dev).manual_seed(0)
Q = torch.tensor(lo, device=dev)[:, None] + torch.tensor(hi - lo, device=dev)[:, None] * torch.rand((8, N), generator=g, device=dev, dtype=torch.float64)
Qt = Q.t().reshape(N // 32, 32, 8).permute(0, 2, 1).contiguous()
T = torch.empty((300, N), dtype=torch.float64, device=dev)
J = torch.empty((48, N), dtype=torch.float64, device=dev)
V = torch.empty((16, N), dtype=torch.float64, device=dev)
G = torch.empty((128, N), dtype=torch.float64, device=dev)
fk = np.arange(1, 26, dtype=np.int32)
jac = np.array([K.find_link(m, "gripper_link").id], dtype=np.int32)
ip = C.POINTER(C.c_int32)
stream = torch.cuda.current_stream(dev)


def call(layout, fused):
    c = L.KinCall()
    c.precision, c.layout, c.n, c.q = L.F64, layout, N, (Qt if layout == L.TILED32 else Q).data_ptr()
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


if what in ("fkg", "coll"):
    fkg = np.array([jac[0]], dtype=np.int32)
    shapes = [(128, 1, 12), (128, 1, 0), (128, 2, 0), (128, 4, 0), (256, 1, 6), (256, 2, 0), (256, 3, 0), (512, 1, 3), (512, 2, 0), (256, 2, 3), (128, 4, 3)] \
        if what == "fkg" else [(128, 2, 0), (128, 3, 0), (256, 1, 0), (160, 2, 0), (192, 2, 0), (96, 3, 0), (64, 4, 0)]
    bytes_cfg = 544 if what == "fkg" else 1216
    for layout, lname in ((L.SOA, "soa"), (L.TILED32, "tiled")):
        for blk, minb, qb in shapes:
            os.environ["KIN_JIT_BLOCK"], os.environ["KIN_JIT_MINB"], os.environ["KIN_JIT_QBATCH"] = str(blk), str(minb), str(qb)
            c = call(layout, what == "coll")
            if what == "fkg":
                c.n_fk_links, c.fk_links = 1, fkg.ctypes.data_as(ip)
            else:
                c.n_fk_links = c.n_jac_links = 0
                c.T_out = c.J_out = None
            try:
                ms, regs, smem, block, grid = timed(c)
            except Exception as e:
                print("%-6s %-6s block %4d minb %d qbatch %2d: %s" % (what, lname, blk, minb, qb, str(e)[:80]))
                continue
            print("%-6s %-6s block %4d minb %d qbatch %2d: %7.3f ms  %6.1f GB/s (%.3f of 6553.6)  regs %3d smem %6d launch block %d grid %d"
                  % (what, lname, blk, minb, qb, ms, bytes_cfg * N / ms / 1e6, bytes_cfg * N / ms / 1e6 / 6553.6, regs, smem, block, grid), flush=True)
    sys.exit(0)

shapes_fkj = [(128, 3, 0), (256, 1, 0), (128, 1, 12), (256, 1, 6), (256, 1, 4), (256, 1, 2), (384, 1, 4), (512, 1, 3), (128, 1, 6)]
shapes_fused = [(128, 2, 0), (256, 1, 0), (288, 1, 0), (320, 1, 0), (352, 1, 0), (160, 2, 0), (176, 2, 0)]
for fused in ([False, True] if what == "all" else [what == "fused"]):
    bytes_cfg = 4000 if fused else 2848
    for layout, lname in ((L.SOA, "soa"), (L.TILED32, "tiled")):
        for blk, minb, qb in (shapes_fused if fused else shapes_fkj):
            os.environ["KIN_JIT_BLOCK"], os.environ["KIN_JIT_MINB"], os.environ["KIN_JIT_QBATCH"] = str(blk), str(minb), str(qb)
            try:
                ms, regs, smem, block, grid = timed(call(layout, fused))
            except Exception as e:
                print("%-6s %-6s block %4d minb %d qbatch %2d: %s" % ("fused" if fused else "fkj", lname, blk, minb, qb, str(e)[:80]))
                continue
            print("%-6s %-6s block %4d minb %d qbatch %2d: %7.3f ms  %6.1f GB/s (%.3f of 6553.6)  regs %3d smem %6d launch block %d grid %d"
                  % ("fused" if fused else "fkj", lname, blk, minb, qb, ms, bytes_cfg * N / ms / 1e6, bytes_cfg * N / ms / 1e6 / 6553.6, regs, smem, block, grid))
