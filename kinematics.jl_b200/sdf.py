"""BoxSDF / UnionSDF (sdf.jl:1-119) on the B200 backend.  An SDF object is a description (box poses
and widths); evaluating it at points or against a robot's collision spheres runs in libkin_b200.

SphereSDF / CylinderSDF are an EXTENSION: the reference supports boxes only
(load_urdf.jl:10-15, sdf.jl:92-94).  They honour the same AbstractSDF contract -- ``sdf(p)``, the generic
forward-difference ``gradient!`` (sdf.jl:34-41), membership in a UnionSDF with first-minimum argmin."""
from __future__ import annotations

import ctypes as C
import uuid

import numpy as np

from . import lib as _lib
from .mechanism import BoxMetaData, CylinderMetaData, Link, Mechanism, SphereMetaData, add_new_link
from .transform import Transform


class AbstractSDF:
    def world_boxes(self):
        """-> (poses (B, 4, 4), widths (B, 3)) in UnionSDF.sdfs order."""
        raise NotImplementedError

    def world_primitives(self):
        """-> (poses (B, 4, 4), sizes (B, 3), kinds (B,) int32: lib.PRIM_BOX / PRIM_SPHERE / PRIM_CYLINDER)."""
        poses, widths = self.world_boxes()
        return poses, widths, np.zeros(len(poses), dtype=np.int32)

    # ---- sdf(p) and gradient!(sdf, p, out) -------------------------------------------------------
    def __call__(self, p, return_argmin=False):
        """sdf.jl:67-74 / 108-114.  ``p``: 3-vector -> float, or points (N, 3) -> tensor (N,)."""
        vals, _, am = self._points(p, want_grad=False, want_argmin=return_argmin)
        return (vals, am) if return_argmin else vals

    def gradient(self, p, grad_mode=_lib.GRAD_FD):
        """``gradient!`` (sdf.jl:34-41, 116-119): forward difference (eps 1e-7) on the argmin box."""
        return self._points(p, want_grad=True, grad_mode=grad_mode)[1]

    def _points(self, p, want_grad, want_argmin=False, grad_mode=_lib.GRAD_FD):
        import torch
        if not torch.cuda.is_available():
            raise _lib.KinError("no CUDA device: the kinematics.jl_b200 operators have no CPU fallback")
        single = not isinstance(p, torch.Tensor) and np.ndim(p) == 1
        P = p if isinstance(p, torch.Tensor) else torch.as_tensor(np.atleast_2d(np.asarray(p, dtype=np.float64)))
        P = P.cuda()
        if P.dtype not in (torch.float32, torch.float64):
            P = P.double()
        P = P.contiguous()
        N = P.shape[0]
        poses, widths, kinds = self.world_primitives()
        poses_cm = np.ascontiguousarray(poses.transpose(0, 2, 1)).reshape(-1, 16)
        widths = np.ascontiguousarray(widths, dtype=np.float64)
        kinds = np.ascontiguousarray(kinds, dtype=np.int32)
        vals = torch.empty(N, dtype=P.dtype, device=P.device)
        grads = torch.empty((N, 3), dtype=P.dtype, device=P.device) if want_grad else None
        am = torch.empty(N, dtype=torch.int32, device=P.device) if want_argmin else None
        dp = C.POINTER(C.c_double)
        _lib.check(_lib.lib().kin_sdf_points_prims(
            len(poses_cm), kinds.ctypes.data_as(C.POINTER(C.c_int32)), poses_cm.ctypes.data_as(dp), widths.ctypes.data_as(dp),
            _lib.F32 if P.dtype == torch.float32 else _lib.F64, _lib.AOS, P.data_ptr(), N, grad_mode,
            vals.data_ptr(), grads.data_ptr() if want_grad else None, am.data_ptr() if want_argmin else None,
            torch.cuda.current_stream(P.device).cuda_stream))
        if single:
            return (float(vals[0]), grads[0].double().cpu().numpy() if want_grad else None,
                    int(am[0]) if want_argmin else None)
        return vals, grads, am


class BoxSDF(AbstractSDF):
    """sdf.jl:48-65: ``BoxSDF(pose, width)`` stand-alone, or attached to a link of a mechanism."""

    def __init__(self, pose, width=None, attach=None):
        if isinstance(pose, BoxMetaData):
            pose, width = pose.origin, pose.extents
        self.pose = pose if isinstance(pose, Transform) else Transform(pose)
        self.width = np.asarray(width, dtype=np.float64)
        self.attach = attach               # (mech, link) for IsAttached (sdf.jl:3-6)

    def world_pose(self):
        if self.attach is None:
            return self.pose.mat
        from .algorithm import get_transform
        mech, link = self.attach           # sdf.jl:14-32: pose of the SDF link in the obstacle mechanism
        return _obstacle_transform(mech, link)

    def world_boxes(self):
        return self.world_pose()[None], self.width[None]


class SphereSDF(BoxSDF):
    """Extension: ``SphereSDF(pose, radius)``; d = |p - c| - r."""
    kind = _lib.PRIM_SPHERE

    def __init__(self, pose, radius=None, attach=None):
        if isinstance(pose, SphereMetaData):
            pose, radius = pose.origin, pose.radius
        BoxSDF.__init__(self, pose, [radius, 0.0, 0.0], attach)
        self.radius = float(radius)

    def world_boxes(self):
        raise _lib.KinError("a SphereSDF is not a box: use world_primitives()")

    def world_primitives(self):
        return self.world_pose()[None], self.width[None], np.array([self.kind], dtype=np.int32)


class CylinderSDF(SphereSDF):
    """Extension: ``CylinderSDF(pose, radius, length)``, axis = local z, centred on the pose (URDF <cylinder>)."""
    kind = _lib.PRIM_CYLINDER

    def __init__(self, pose, radius=None, length=None, attach=None):
        if isinstance(pose, CylinderMetaData):
            pose, radius, length = pose.origin, pose.radius, pose.length
        BoxSDF.__init__(self, pose, [radius, length, 0.0], attach)
        self.radius, self.length = float(radius), float(length)


def _obstacle_transform(mech: Mechanism, link: Link):
    """World pose of an attached SDF link at the obstacle mechanism's current state (FK runs on the GPU,
    N = 1), memoised on the mechanism per state version like the reference memoises inv_pose (sdf.jl:14-20)."""
    from .algorithm import get_transform
    version = (mech._state_version, mech._structure_version)
    cache = getattr(mech, "_sdf_pose_cache", None)
    if cache is None or cache[0] != version:
        cache = mech._sdf_pose_cache = (version, {})
    if link.id not in cache[1]:
        assert mech._single, "an obstacle mechanism must hold a single configuration"
        cache[1][link.id] = get_transform(mech, link).mat
    return cache[1][link.id]


class UnionSDF(AbstractSDF):
    """sdf.jl:76-114.  ``UnionSDF(mech)`` makes one box per link of ``mech`` that carries box collision
    metadata, in ``mech.links`` order, attached through a new link placed at the collision origin
    (sdf.jl:82-97); ``UnionSDF([sdf, ...])`` unions existing SDFs."""

    def __init__(self, arg, primitives=False):
        """``primitives=True`` (extension) also turns the URDF's <sphere> / <cylinder> collision geometry into
        SphereSDF / CylinderSDF members; the default skips them as the reference does (sdf.jl:92-94)."""
        if isinstance(arg, Mechanism):
            mech, sdfs = arg, []
            for link in list(mech.links):
                meta = link.geometric_meta_data
                cls = BoxSDF if isinstance(meta, BoxMetaData) else None
                if primitives and link.link_type == "URDF":
                    cls = SphereSDF if isinstance(meta, SphereMetaData) else CylinderSDF if isinstance(meta, CylinderMetaData) else cls
                if cls is not None:
                    new_link = Link("boxsdf_" + str(uuid.uuid1()), link_type="SdfLinkType")
                    add_new_link(mech, new_link, link, meta.origin)
                    sdfs.append(cls(meta, attach=(mech, new_link)))
            self.sdfs = sdfs
        else:
            self.sdfs = list(arg)

    def world_boxes(self):
        parts = [s.world_boxes() for s in self.sdfs]
        return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])

    def world_primitives(self):
        parts = [s.world_primitives() for s in self.sdfs]
        return tuple(np.concatenate([p[i] for p in parts]) for i in range(3))
