"""Tiled FK-all + gripper Jacobian with plain st.global.cs stores against cp.async.bulk (TMA) stores from per-warp stages
(KIN_JIT_BULK), a few launch shapes.   python profiles/sweep_bulk.py [log2 N]"""
import os
import subprocess
import sys

N = sys.argv[1] if len(sys.argv) > 1 else "24"
here = os.path.dirname(os.path.abspath(__file__))
code = r'''
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(%r))
import kinematics_jl_b200 as K
from kinematics_jl_b200 import lib as L
from kinematics_jl_b200.device import device_model
import scene_fetch
N = 1 << int(%r)
dev = torch.device("cuda", 0)
m, joints, sscc = scene_fetch.product_fetch(False)
K.set_joint_angles(m, joints, torch.zeros((1, 8), dtype=torch.float64, device=dev))
dm = device_model(m); lib = L.lib()
lo, hi = scene_fetch.joint_limits(joints)
g = torch.Generator(device=dev).manual_seed(0)
Q = torch.tensor(lo, device=dev)[:, None] + torch.tensor(hi - lo, device=dev)[:, None] * torch.rand((8, N), generator=g, device=dev, dtype=torch.float64)
Qt = Q.t().reshape(N // 32, 32, 8).permute(0, 2, 1).contiguous()
T = torch.empty((300, N), dtype=torch.float64, device=dev); J = torch.empty((48, N), dtype=torch.float64, device=dev)
fk = np.arange(1, 26, dtype=np.int32); jac = np.array([K.find_link(m, "gripper_link").id], dtype=np.int32)
ip = C.POINTER(C.c_int32); stream = torch.cuda.current_stream(dev)
c = L.KinCall()
c.precision, c.layout, c.n, c.q = L.F64, L.TILED32, N, Qt.data_ptr()
c.n_fk_links, c.fk_links, c.T_out = 25, fk.ctypes.data_as(ip), T.data_ptr()
c.n_jac_links, c.jac_links, c.J_out, c.with_rot = 1, jac.ctypes.data_as(ip), J.data_ptr(), 1
c.truncation_dist = float("inf"); c.stream = stream.cuda_stream
for _ in range(3): L.check(lib.kin_eval(dm.h, C.byref(c)))
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(stream)
for _ in range(10): L.check(lib.kin_eval(dm.h, C.byref(c)))
b.record(stream); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
regs, smem, block, grid = (C.c_int32() for _ in range(4))
L.check(lib.kin_query_launch(dm.h, C.byref(c), C.byref(regs), C.byref(smem), C.byref(block), C.byref(grid)))
print("%%7.3f ms  %%6.1f GB/s (%%.3f of 6553.6)  regs %%3d smem %%6d block %%d grid %%d  checksum %%.6f" %% (ms, 2848 * N / ms / 1e6, 2848 * N / ms / 1e6 / 6553.6, regs.value, smem.value, block.value, grid.value, float(T[:, ::4097].sum() + J[:, ::4097].sum())))
''' % (here + "/x", N)
for bulk, blk, minb, qb in ((0, 128, 1, 12), (1, 128, 1, 11), (1, 128, 1, 8), (1, 256, 1, 5), (1, 128, 2, 5), (1, 128, 3, 0), (1, 256, 2, 0), (0, 128, 3, 0)):
    env = dict(os.environ, KIN_JIT_BULK=str(bulk), KIN_JIT_BLOCK=str(blk), KIN_JIT_MINB=str(minb), KIN_JIT_QBATCH=str(qb))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print("bulk %d block %3d minb %d qbatch %2d: %s" % (bulk, blk, minb, qb, (r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1]), flush=True)
