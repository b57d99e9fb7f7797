"""ctypes binding of libkin_b200.so (include/kin_b200.h).  There is NO fallback: if the shared
library is missing or there is no CUDA device, every operator raises."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libkin_b200.so")
CSRC = os.path.join(HERE, "csrc")

F64, F32 = 0, 1
SOA, AOS, TILED32 = 0, 1, 2
FIXED, REVOLUTE, PRISMATIC = 0, 1, 2
GRAD_FD, GRAD_ANALYTIC, GRAD_FD_DIRECT = 0, 1, 2
SCRATCH_REFERENCE, SCRATCH_CLEAN = 0, 1
POSE_IK_OBJECTIVE, POSE_CONSTRAINT = 0, 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared", "-cudart", "static"]
SOURCES = ["kin_b200.cu", "kin_model.cpp"]
HEADERS = ["kin_kernels.cuh", "kin_kernels_ws.cuh", "kin_program.h", "kin_model.hpp"]


class KinError(RuntimeError):
    pass


class KinModelDesc(C.Structure):
    _fields_ = [("n_links", C.c_int32), ("parent_link", _ip), ("joint_type", _ip), ("joint_pose", _dp),
                ("joint_axis", _dp), ("q_index", _ip), ("default_angle", _dp), ("n_joints", C.c_int32),
                ("with_base", C.c_int32), ("n_spheres", C.c_int32), ("sphere_link", _ip), ("sphere_center", _dp),
                ("sphere_radius", _dp), ("n_boxes", C.c_int32), ("box_pose", _dp), ("box_width", _dp)]


class KinCall(C.Structure):
    _fields_ = [("precision", C.c_int32), ("layout", C.c_int32), ("n", C.c_int64), ("batch_stride", C.c_int64),
                ("q", C.c_void_p),
                ("n_fk_links", C.c_int32), ("fk_links", _ip), ("T_out", C.c_void_p),
                ("n_jac_links", C.c_int32), ("jac_links", _ip), ("with_rot", C.c_int32), ("rpy_jac", C.c_int32),
                ("keep_irrelevant", C.c_int32), ("J_out", C.c_void_p),
                ("truncation_dist", C.c_double), ("grad_mode", C.c_int32), ("scratch_mode", C.c_int32),
                ("vals_out", C.c_void_p), ("grads_out", C.c_void_p), ("argmin_out", C.c_void_p),
                ("vals_offset", C.c_double), ("stream", C.c_void_p)]


EXPORTS = ["kin_last_error", "kin_abi_version", "kin_model_create", "kin_model_destroy", "kin_model_set_spheres",
           "kin_model_set_boxes", "kin_model_n_dof", "kin_model_n_spheres", "kin_model_n_boxes", "kin_eval",
           "kin_eval_host", "kin_fk_links", "kin_fk_jacobian", "kin_collision", "kin_launch_count",
           "kin_query_launch", "kin_sdf_points", "kin_program_dump", "kin_pose_residual", "kin_lm_step", "kin_lm_accept"]


def needs_build() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(os.path.dirname(HERE), "include", "kin_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> kinematics.jl_b200/libkin_b200.so (in-tree)."""
    if force or needs_build():
        cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO_PATH] + \
              [os.path.join(CSRC, s) for s in SOURCES]
        out = subprocess.run(cmd, capture_output=True, text=True)
        if out.returncode != 0:
            raise KinError("nvcc failed:\n" + out.stdout + out.stderr)
        if verbose:
            print(out.stderr)
    return SO_PATH


_LIB = None


def lib():
    """Load libkin_b200.so; raises KinError when it has not been built (no fallback path exists)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise KinError("libkin_b200.so is missing (%s): run `python -c 'import __graft_entry__ as g; g.build()'`; "
                           "there is no CPU fallback" % SO_PATH)
        L = C.CDLL(SO_PATH)
        L.kin_last_error.restype = C.c_char_p
        L.kin_abi_version.restype = C.c_int
        L.kin_model_create.argtypes = [C.POINTER(KinModelDesc), C.POINTER(C.c_void_p)]
        L.kin_model_destroy.argtypes = [C.c_void_p]
        L.kin_model_set_spheres.argtypes = [C.c_void_p, C.c_int32, _ip, _dp, _dp]
        L.kin_model_set_boxes.argtypes = [C.c_void_p, C.c_int32, _dp, _dp]
        for f in (L.kin_model_n_dof, L.kin_model_n_spheres, L.kin_model_n_boxes):
            f.argtypes = [C.c_void_p]
        L.kin_eval.argtypes = [C.c_void_p, C.POINTER(KinCall)]
        L.kin_eval_host.argtypes = [C.c_void_p, C.POINTER(KinCall)]
        L.kin_launch_count.restype = C.c_int64
        L.kin_query_launch.argtypes = [C.c_void_p, C.POINTER(KinCall), _ip, _ip, _ip, _ip]
        L.kin_sdf_points.argtypes = [C.c_int32, _dp, _dp, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.kin_pose_residual.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                        C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.kin_lm_step.argtypes = [C.c_int64, C.c_int32, C.c_int32] + [C.c_void_p] * 8
        L.kin_lm_accept.argtypes = [C.c_int64, C.c_int32, C.c_int32] + [C.c_void_p] * 10
        L.kin_program_dump.argtypes = [C.POINTER(KinModelDesc), _ip, C.c_int32, _ip, C.c_int32, C.c_int32, C.c_int32,
                                       _ip, C.c_int32, _ip, C.c_int32, _dp, C.c_int32]
        _LIB = L
    return _LIB


def check(rc: int):
    if rc != 0:
        raise KinError("libkin_b200 error %d: %s" % (rc, lib().kin_last_error().decode()))
