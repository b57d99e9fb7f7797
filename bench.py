#!/usr/bin/env python
"""bench.py -- throughput of the Kinematics.jl hot path on B200 (driver contract in the task brief).

Default workload (BASELINE.json configs[1]): Fetch (data/fetch.urdf), 8 control joints, FP64, SoA:
    per configuration: get_transform for all 25 links + 6x8 geometric Jacobian of gripper_link,
    N = 2^24 random in-limit configurations per GPU (batch-sharded, weak scaling, no collective).
A "step" is one kin_eval over the whole batch (one kernel launch).  Inputs (1 GiB) and outputs
(46.7 GB) are far larger than the 126 MB L2, so nothing is served from cache between steps.

The same JSON line also carries `north_star` -- the fused FK-all + gripper Jacobian + 16-sphere /
fridge-SDF collision cost and gradient step (BASELINE.json north_star, configs[1]+[2] in one pass).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FETCH_JOINTS = ["torso_lift_joint", "shoulder_pan_joint", "shoulder_lift_joint", "upperarm_roll_joint",
                "elbow_flex_joint", "forearm_roll_joint", "wrist_flex_joint", "wrist_roll_joint"]
N_LINKS, N_DOF, N_SPH = 25, 8, 16
# algorithmic bytes per configuration (SURVEY 8d / DESIGN.md): q in, 3x4 per link out, 6x8 Jacobian,
# S distances + S x n_dof gradients
BYTES_FKJ = 8 * N_DOF + 8 * 12 * N_LINKS + 8 * 6 * N_DOF                   # 2848
BYTES_FUSED = BYTES_FKJ + 8 * N_SPH + 8 * N_SPH * N_DOF                     # 4000
# DRAM bytes per configuration measured by ncu (dram__bytes_read.sum + dram__bytes_write.sum of one
# `--set full` capture of a 4 194 304-configuration launch, profiles/r01b_fkj_ncu_summary.txt and
# profiles/r01c_fused_ws_ncu_summary.txt)
NCU_DRAM_BYTES_PER_CONFIG_FKJ = (268579072 + 11617420000) / 4194304          # 2833.8
NCU_DRAM_BYTES_PER_CONFIG_FUSED = (272353792 + 16609201000) / 4194304        # 4024.9 (warp-specialised kernel)
METRIC = "fetch_fk_jacobian_configs_per_s"
UNIT = "configs/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 7 and r[3 + k].lower().startswith("active") for r in self.rows)]
        # the median under load: ignore idle samples (low power) when there are loaded ones
        loaded = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] if pw else sm
        return {"sm_mhz": float(np.median(loaded)) if loaded else None, "sm_max_mhz": mx[0] if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference, all host threads, bounded sample
# ---------------------------------------------------------------------------------------------------
def cpu_arm(workload, target_seconds, fused):
    """Times the restated reference CPU path (oracle/, kind = "port": the Julia reference cannot run here).
    Returns dict(value=configs/s, cores, sample, seconds)."""
    from oracle import ref_model as R
    import scenes
    mo, jo, so = scenes.oracle_fetch(False)
    sdf_o = scenes.oracle_fridge_sdf() if fused else None
    threads = R.max_threads()
    links = mo.links[:N_LINKS]
    gl = R.find_link(mo, "gripper_link")

    def run(n, seed):
        q = scenes.random_configs(jo, n, False, seed=seed)
        t0 = time.perf_counter()
        R.batch_fused(so, jo, sdf_o, q, links, gl, with_rot=True, rpy_jac=False, n_threads=threads,
                      keep_outputs=False)
        return time.perf_counter() - t0

    run(20000, 1)                                   # warm-up (page-in, thread start)
    probe_n = 200000
    rate = probe_n / run(probe_n, 2)
    n = int(max(probe_n, min(rate * target_seconds, 5e7)))
    dt = run(n, 3)
    return {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d random in-limit Fetch configurations of the %s workload, one pass, %.1f s, %d pthreads, "
                      "one mechanism clone per thread" % (n, workload, dt, threads)}


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    workload = "fk_all_links+gripper_jacobian"
    vals, last = [], None
    per_step = max(2.0, min(20.0, 90.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        last = cpu_arm(workload, per_step, fused=False)
        if i >= args.warmup:
            vals.append(last["value"])
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "model": "data/fetch.urdf", "n_links": N_LINKS, "n_dof": N_DOF,
                       "note": "reference = CPU restatement of the Julia path (oracle/, Julia + scikit-robot are not installable here)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": last["cores"], "kind": "port", "sample": last["sample"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--configs", dest="n", type=int, default=1 << 24, help="configurations per GPU")
    ap.add_argument("--configs-e2e", dest="n_e2e", type=int, default=1 << 21, help="configurations per e2e step (host buffers)")
    ap.add_argument("--skip-cpu", dest="no_cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--skip-north-star", dest="no_north_star", action="store_true")
    ap.add_argument("--skip-variants", dest="no_variants", action="store_true", help="skip the other BASELINE.md rows / layouts")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_main(args)

    import torch
    import kinematics_jl_b200 as K
    from kinematics_jl_b200 import lib as L
    from kinematics_jl_b200.device import device_model
    import scenes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W = max(3, args.warmup)
    N = args.n
    dev = torch.device("cuda", local)

    # ---- model: Fetch + sphere fixture + fridge boxes (fridge_demo.jl:28) ----
    m, joints, sscc = scenes.product_fetch(False)
    fridge = K.parse_urdf(os.path.join(ROOT, "data", "fridge.urdf"), with_base=True)
    K.set_joint_angles(fridge, [K.find_joint(fridge, "door_joint")], scenes.FRIDGE_STATE)
    sdf = K.UnionSDF(fridge)
    K.set_joint_angles(m, joints, torch.zeros((1, N_DOF), dtype=torch.float64, device=dev))
    K.compute_coll_dists(sscc, joints, sdf)          # builds the device model and uploads sphere / box tables
    dm = device_model(m)
    lib = L.lib()

    # ---- synthetic inputs, resident in HBM (SoA: q[d][n]) ----
    g = torch.Generator(device=dev).manual_seed(rank)
    lo = torch.tensor([j.lower_limit if np.isfinite(j.lower_limit) else -np.pi for j in joints], device=dev, dtype=torch.float64)
    hi = torch.tensor([j.upper_limit if np.isfinite(j.upper_limit) else np.pi for j in joints], device=dev, dtype=torch.float64)
    Q = lo[:, None] + (hi - lo)[:, None] * torch.rand((N_DOF, N), generator=g, device=dev, dtype=torch.float64)
    T = torch.empty((N_LINKS * 12, N), dtype=torch.float64, device=dev)
    J = torch.empty((6 * N_DOF, N), dtype=torch.float64, device=dev)
    fk_ids = np.arange(1, N_LINKS + 1, dtype=np.int32)
    jac_ids = np.array([K.find_link(m, "gripper_link").id], dtype=np.int32)
    ip = C.POINTER(C.c_int32)
    stream = torch.cuda.current_stream(dev)

    def make_call(n, q, T_, J_, V_=None, G_=None, layout=L.SOA):
        c = L.KinCall()
        c.precision, c.layout, c.n, c.q = L.F64, layout, n, q
        c.n_fk_links, c.fk_links, c.T_out = N_LINKS, fk_ids.ctypes.data_as(ip), T_
        c.n_jac_links, c.jac_links, c.J_out, c.with_rot = 1, jac_ids.ctypes.data_as(ip), J_, 1
        c.truncation_dist = float("inf")
        c.grad_mode, c.scratch_mode = L.GRAD_FD, L.SCRATCH_REFERENCE
        c.vals_out, c.grads_out = V_, G_
        c.stream = stream.cuda_stream
        return c

    def timed(call, steps, warm):
        for _ in range(warm):
            L.check(lib.kin_eval(dm.h, C.byref(call)))
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        n0 = lib.kin_launch_count()
        evs[0].record(stream)
        for i in range(steps):
            L.check(lib.kin_eval(dm.h, C.byref(call)))
            evs[i + 1].record(stream)
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)
        launches = lib.kin_launch_count() - n0
        total_ms = evs[0].elapsed_time(evs[-1])
        per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
        if dist is not None:
            t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms, per, launches

    def launch_info(call):
        regs, smem, block, grid = (C.c_int32() for _ in range(4))
        L.check(lib.kin_query_launch(dm.h, C.byref(call), C.byref(regs), C.byref(smem), C.byref(block), C.byref(grid)))
        return {"regs": regs.value, "smem_bytes": smem.value, "block": block.value, "grid": grid.value}

    peak, peak_src = peaks()
    sampler = ClockSampler(local)

    # ---- headline: FK all links + gripper Jacobian ----
    call = make_call(N, Q.data_ptr(), T.data_ptr(), J.data_ptr())
    if rank == 0:
        sampler.start()
    total_ms, per, launches = timed(call, args.steps, W)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = world * N / (ms_per_step * 1e-3)
    kern_ms = float(np.mean(per))
    achieved = BYTES_FKJ * N / (kern_ms * 1e-3) / 1e9
    info = launch_info(call)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": NCU_DRAM_BYTES_PER_CONFIG_FKJ * N, "traffic_source": "ncu dram__bytes_read+write per configuration "
                "(profiles/r01b_fkj_ncu_summary.txt, 2^22-configuration launch) x configurations per launch",
                "peak_source": peak_src, "kernel": "kin_eval_kernel<double,SoA>",
                "algorithmic_bytes_per_config": BYTES_FKJ, "launch_ms": kern_ms, "launch": info}

    # ---- north star: fused FK-all + Jacobian + collision cost/grad ----
    north = None
    if not args.no_north_star:
        V = torch.empty((N_SPH, N), dtype=torch.float64, device=dev)
        G = torch.empty((N_SPH * N_DOF, N), dtype=torch.float64, device=dev)
        callf = make_call(N, Q.data_ptr(), T.data_ptr(), J.data_ptr(), V.data_ptr(), G.data_ptr())
        tms, perf_, _ = timed(callf, args.steps, W)
        ms_f = tms / args.steps
        ach_f = BYTES_FUSED * N / (float(np.mean(perf_)) * 1e-3) / 1e9
        north = {"workload": "fk_all_links+gripper_jacobian+collision_cost_grad(S=16,B=7,fd,reference-scratch)",
                 "value": world * N / (ms_f * 1e-3), "unit": UNIT, "ms_per_step": ms_f,
                 "roofline": {"bound": "hbm-or-fp64 (see DESIGN.md)", "achieved": ach_f, "peak": peak, "unit": "GB/s",
                              "frac": ach_f / peak, "algorithmic_bytes_per_config": BYTES_FUSED,
                              "traffic": NCU_DRAM_BYTES_PER_CONFIG_FUSED * N,
                              "kernel": "kin_eval_ws_kernel<SoA> (block 384) when launch.block == 384, else kin_eval_kernel",
                              "launch": launch_info(callf)}}
        del V, G

    # ---- other rows of BASELINE.md section 3, same batch (reported, not the headline) ----
    variants = {}
    if not args.no_variants:
        def variant(name, bytes_per_cfg, call_):
            tms_, per_, _ = timed(call_, max(3, args.steps // 2), W)
            ms_ = float(np.mean(per_))
            variants[name] = {"value": world * N / (ms_ * 1e-3), "unit": UNIT, "ms_per_step": ms_,
                              "algorithmic_bytes_per_config": bytes_per_cfg,
                              "hbm_frac": bytes_per_cfg * N / (ms_ * 1e-3) / 1e9 / peak, "launch": launch_info(call_)}
        cg = make_call(N, Q.data_ptr(), T.data_ptr(), J.data_ptr())
        cg.n_fk_links, cg.fk_links = 1, jac_ids.ctypes.data_as(ip)
        variant("fk_gripper+gripper_jacobian", 8 * N_DOF + 96 + 384, cg)
        V = torch.empty((N_SPH, N), dtype=torch.float64, device=dev)
        G = torch.empty((N_SPH * N_DOF, N), dtype=torch.float64, device=dev)
        cc = make_call(N, Q.data_ptr(), None, None, V.data_ptr(), G.data_ptr())
        cc.n_fk_links = cc.n_jac_links = 0
        variant("collision_cost_grad(S=16,B=7,fd,reference-scratch)", 8 * N_DOF + 8 * N_SPH + 8 * N_SPH * N_DOF, cc)
        cc2 = make_call(N, Q.data_ptr(), None, None, V.data_ptr(), G.data_ptr())
        cc2.n_fk_links = cc2.n_jac_links = 0
        cc2.grad_mode, cc2.scratch_mode = L.GRAD_ANALYTIC, L.SCRATCH_CLEAN
        variant("collision_cost_grad(S=16,B=7,analytic,clean-scratch)", 8 * N_DOF + 8 * N_SPH + 8 * N_SPH * N_DOF, cc2)
        del V, G
        # the tiled (AoSoA-32) layout: one contiguous block per warp
        Qt = Q.t().reshape(N // 32, 32, N_DOF).permute(0, 2, 1).contiguous()
        Tt = torch.empty((N // 32, N_LINKS * 12, 32), dtype=torch.float64, device=dev)
        Jt = torch.empty((N // 32, 6 * N_DOF, 32), dtype=torch.float64, device=dev)
        ct = make_call(N, Qt.data_ptr(), Tt.data_ptr(), Jt.data_ptr(), layout=L.TILED32)
        variant("fk_all_links+gripper_jacobian_tiled32_layout", BYTES_FKJ, ct)
        Vt = torch.empty((N // 32, N_SPH, 32), dtype=torch.float64, device=dev)
        Gt = torch.empty((N // 32, N_SPH * N_DOF, 32), dtype=torch.float64, device=dev)
        ctf = make_call(N, Qt.data_ptr(), Tt.data_ptr(), Jt.data_ptr(), Vt.data_ptr(), Gt.data_ptr(), layout=L.TILED32)
        variant("fused_tiled32_layout", BYTES_FUSED, ctf)
        del Qt, Tt, Jt, Vt, Gt
        # the optional FP32 mode (1e-5 tolerance), fused step, SoA
        Q32 = Q.float()
        T32 = torch.empty((N_LINKS * 12, N), dtype=torch.float32, device=dev)
        J32 = torch.empty((6 * N_DOF, N), dtype=torch.float32, device=dev)
        V32 = torch.empty((N_SPH, N), dtype=torch.float32, device=dev)
        G32 = torch.empty((N_SPH * N_DOF, N), dtype=torch.float32, device=dev)
        c32 = make_call(N, Q32.data_ptr(), T32.data_ptr(), J32.data_ptr(), V32.data_ptr(), G32.data_ptr())
        c32.precision = L.F32
        c32.grad_mode = L.GRAD_ANALYTIC          # a 1e-7 forward difference is meaningless in FP32
        variant("fused_fp32_mode(analytic gradient)", BYTES_FUSED // 2, c32)
        del Q32, T32, J32, V32, G32
        # the fused step in the AoS layout (one contiguous record per configuration, planning.jl:58)
        Na = min(N, 1 << 22)
        Qa = Q[:, :Na].t().contiguous()
        Ta = torch.empty((Na, N_LINKS * 12), dtype=torch.float64, device=dev)
        Ja = torch.empty((Na, 6 * N_DOF), dtype=torch.float64, device=dev)
        Va = torch.empty((Na, N_SPH), dtype=torch.float64, device=dev)
        Ga = torch.empty((Na, N_SPH * N_DOF), dtype=torch.float64, device=dev)
        ca = make_call(Na, Qa.data_ptr(), Ta.data_ptr(), Ja.data_ptr(), Va.data_ptr(), Ga.data_ptr(), layout=L.AOS)
        tms_, per_, _ = timed(ca, max(3, args.steps // 2), W)
        ms_ = float(np.mean(per_))
        variants["fused_aos_layout"] = {"value": world * Na / (ms_ * 1e-3), "unit": UNIT, "ms_per_step": ms_, "configs": Na,
                                        "hbm_frac": BYTES_FUSED * Na / (ms_ * 1e-3) / 1e9 / peak, "launch": launch_info(ca)}
        del Qa, Ta, Ja, Va, Ga
        # the fused step on the Fetch WITH the planar base (11 columns: SURVEY 8d variant), SoA
        mb, jb, sb = scenes.product_fetch(True)
        Nb = min(N, 1 << 22)
        K.set_joint_angles(mb, jb, torch.zeros((1, N_DOF + 3), dtype=torch.float64, device=dev))
        K.compute_coll_dists(sb, jb, sdf)
        dmb = device_model(mb)
        gb = torch.Generator(device=dev).manual_seed(1000 + rank)
        lob = torch.cat([lo, torch.tensor([-1.0, -1.0, -np.pi], device=dev, dtype=torch.float64)])
        hib = torch.cat([hi, torch.tensor([1.0, 1.0, np.pi], device=dev, dtype=torch.float64)])
        Qb = lob[:, None] + (hib - lob)[:, None] * torch.rand((N_DOF + 3, Nb), generator=gb, device=dev, dtype=torch.float64)
        Tb = torch.empty((N_LINKS * 12, Nb), dtype=torch.float64, device=dev)
        Jb = torch.empty((6 * (N_DOF + 3), Nb), dtype=torch.float64, device=dev)
        Vb = torch.empty((N_SPH, Nb), dtype=torch.float64, device=dev)
        Gb = torch.empty((N_SPH * (N_DOF + 3), Nb), dtype=torch.float64, device=dev)
        cb = make_call(Nb, Qb.data_ptr(), Tb.data_ptr(), Jb.data_ptr(), Vb.data_ptr(), Gb.data_ptr())
        jac_b = np.array([K.find_link(mb, "gripper_link").id], dtype=np.int32)
        cb.jac_links = jac_b.ctypes.data_as(ip)
        dm_main, dm = dm, dmb                      # timed() / launch_info() use `dm`
        try:
            tms_, per_, _ = timed(cb, max(3, args.steps // 2), W)
            ms_ = float(np.mean(per_))
            bytes_b = 8 * (N_DOF + 3) + 8 * 12 * N_LINKS + 8 * 6 * (N_DOF + 3) + 8 * N_SPH + 8 * N_SPH * (N_DOF + 3)
            variants["fused_with_planar_base(11 columns)"] = {
                "value": world * Nb / (ms_ * 1e-3), "unit": UNIT, "ms_per_step": ms_, "configs": Nb,
                "algorithmic_bytes_per_config": bytes_b, "hbm_frac": bytes_b * Nb / (ms_ * 1e-3) / 1e9 / peak,
                "launch": launch_info(cb)}
        finally:
            dm = dm_main
        del Qb, Tb, Jb, Vb, Gb

    # ---- configs 4 and 5 of BASELINE.json (caller-side rows of SURVEY 8f), through the host mirror ----
    callers = {}
    if not args.no_variants:
        def ev_time(fn, reps):
            fn()
            torch.cuda.synchronize(dev)
            if dist is not None:
                dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(reps):
                r = fn()
            b.record(stream)
            torch.cuda.synchronize(dev)
            ms = a.elapsed_time(b) / reps
            if dist is not None:
                t = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            return ms, r
        # config 5: 4096 problems x 64 waypoints, straight-line initial trajectories, IneqConst stack
        # (margin 0.03, truncation 0.08), problems sharded over the ranks, stacked outputs all-gathered
        P, n_wp = 4096, 64
        p0, p1 = K.shard_range(P, rank, world)
        gq = torch.Generator(device=dev).manual_seed(1234)
        qs = lo + (hi - lo) * torch.rand((P, N_DOF), generator=gq, device=dev, dtype=torch.float64)
        qg = lo + (hi - lo) * torch.rand((P, N_DOF), generator=gq, device=dev, dtype=torch.float64)
        X = K.create_straight_trajectory(qs[p0:p1], qg[p0:p1], n_wp)
        G5 = K.IneqConst(sscc, joints, sdf, n_wp, 0.03)
        ms_eval, (v5, g5) = ev_time(lambda: G5(X), 5)
        ms_all, _ = ev_time(lambda: K.gather_stacked(*G5(X)), 5)
        callers["trajectory_stack(4096x64)"] = {
            "waypoint_configs_per_s_eval_only": P * n_wp / (ms_eval * 1e-3), "waypoint_configs_per_s_with_allgather": P * n_wp / (ms_all * 1e-3),
            "ms_eval": ms_eval, "ms_eval_plus_allgather": ms_all, "n_gpus": world,
            "gathered_bytes": 8 * P * n_wp * (N_SPH + N_SPH * N_DOF), "collective": "NCCL all_gather of vals/grads slabs" if world > 1 else "none (1 rank)"}
        del X, v5, g5
        # config 4: 2^20 independent gripper pose targets (FK of random in-limit configurations), LM iterations
        Nik = 1 << 20
        qt = lo + (hi - lo) * torch.rand((Nik, N_DOF), generator=gq, device=dev, dtype=torch.float64)
        K.set_joint_angles(m, joints, qt)
        gl = K.find_link(m, "gripper_link")
        Tg = K.get_transform(m, gl)
        c1 = torch.cos(torch.atan2(Tg[:, 1, 0], Tg[:, 0, 0]))
        yaw = torch.atan2(Tg[:, 1, 0], Tg[:, 0, 0])
        pitch = torch.atan2(-Tg[:, 2, 0], torch.sqrt(Tg[:, 2, 1] ** 2 + Tg[:, 2, 2] ** 2))
        roll = torch.atan2(Tg[:, 0, 2] * torch.sin(yaw) - Tg[:, 1, 2] * c1, Tg[:, 1, 1] * c1 - Tg[:, 0, 1] * torch.sin(yaw))
        tg = torch.cat([Tg[:, :, 3], roll[:, None], pitch[:, None], yaw[:, None]], dim=1).contiguous()
        q0 = torch.tensor([0.2, 0, 0, 0, 0.5, 0, 0.5, 0], device=dev, dtype=torch.float64).repeat(Nik, 1)
        K.set_joint_angles(m, joints, q0)
        ms_it, _ = ev_time(lambda: K.pose_constraint(m, gl, joints, tg, True), 5)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        qsol, fsol = K.inverse_kinematics_batch(m, gl, joints, tg, q0, with_rot=True, iters=40)
        torch.cuda.synchronize(dev)
        t_solve = time.perf_counter() - t0
        K.set_joint_angles(m, joints, qsol)
        vv, _ = K.pose_constraint(m, gl, joints, tg, True)
        vv[:, 3:] = torch.remainder(vv[:, 3:] + np.pi, 2 * np.pi) - np.pi
        ok = float((vv.abs().amax(dim=1) < 1e-3).double().mean())
        callers["batched_ik(2^20 targets)"] = {"residual_and_jacobian_evals_per_s": world * Nik / (ms_it * 1e-3), "ms_per_evaluation": ms_it,
                                               "solve_seconds_40_lm_iterations": t_solve, "targets_per_s": world * Nik / t_solve,
                                               "fraction_within_1e-3": ok, "n_gpus": world}
        del qt, Tg, tg, q0, qsol, fsol, vv
        K.set_joint_angles(m, joints, torch.zeros((1, N_DOF), dtype=torch.float64, device=dev))

    # ---- e2e: the C-ABI call with HOST buffers (pinned), H2D + D2H inside the timed region ----
    Ne = min(args.n_e2e, N)
    qh = torch.empty((N_DOF, Ne), dtype=torch.float64).pin_memory()
    qh.copy_(Q[:, :Ne].cpu())
    Th = torch.empty((N_LINKS * 12, Ne), dtype=torch.float64).pin_memory()
    Jh = torch.empty((6 * N_DOF, Ne), dtype=torch.float64).pin_memory()
    calle = make_call(Ne, qh.data_ptr(), Th.data_ptr(), Jh.data_ptr())
    for _ in range(2):
        L.check(lib.kin_eval_host(dm.h, C.byref(calle)))
    if dist is not None:
        dist.barrier()
    e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e_steps):
        L.check(lib.kin_eval_host(dm.h, C.byref(calle)))     # returns when the outputs are in host memory
    e_dt = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e_dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_dt = float(t.item())
    e2e = {"value": world * Ne * e_steps / e_dt, "unit": UNIT, "h2d_bytes_per_step": 8 * N_DOF * Ne,
           "d2h_bytes_per_step": 8 * (N_LINKS * 12 + 6 * N_DOF) * Ne, "configs_per_step": Ne, "steps": e_steps,
           "api": "kin_eval_host (C ABI, pinned host q / T / J, chunked H2D -> kernel -> D2H on 3 streams)",
           "check": float(Th[9, 0])}

    # ---- CPU baseline (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_arm("fk_all_links+gripper_jacobian", 12.0, fused=False)
        if north is not None:
            north["cpu_baseline"] = cpu_arm(north["workload"], 12.0, fused=True)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": "fk_all_links+gripper_jacobian", "model": "data/fetch.urdf", "n_links": N_LINKS,
                           "n_dof": N_DOF, "configs_per_gpu": N, "layout": "soa", "parallelism": "batch-shard x%d, no collective" % world,
                           "l2": "inputs+outputs per step (%.1f GB) exceed the 126 MB L2" % (BYTES_FKJ * N / 1e9)},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "north_star": north, "variants": variants, "callers": callers}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
