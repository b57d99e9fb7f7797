"""Host-side mirror of the reference's ``Transform`` (transform.jl:1-65): a 4x4 homogeneous matrix.
Only used to describe models and to hand single-configuration results back; the batched arithmetic
runs in libkin_b200."""
from __future__ import annotations

import numpy as np


class Transform:
    """transform.jl:3-31.  ``Transform(trans, rot)``, ``Transform(rot)``, ``Transform(trans)``, ``Transform(mat4x4)``."""

    __slots__ = ("mat",)

    def __init__(self, a=None, b=None):
        m = np.eye(4)
        if a is not None:
            a = np.asarray(a, dtype=np.float64)
            if a.shape == (4, 4):
                m = a.copy()
            elif a.shape == (3, 3):
                m[:3, :3] = a
            elif a.shape == (3,):
                m[:3, 3] = a
                if b is not None:
                    m[:3, :3] = np.asarray(b, dtype=np.float64)
            else:
                raise ValueError("Transform: expected a 4x4, a 3x3 rotation or a 3-vector")
        self.mat = m

    def __mul__(self, other):              # transform.jl:39-41, 58-60
        if isinstance(other, Transform):
            return Transform(self.mat @ other.mat)
        p = np.asarray(other, dtype=np.float64)
        return self.mat[:3, 3] + self.mat[:3, :3] @ p

    def inv(self):                         # transform.jl:62-65
        R = self.mat[:3, :3].T
        return Transform(-R @ self.mat[:3, 3], R)

    def __repr__(self):
        return "Transform(%r)" % (self.mat,)


def rotation(t: Transform):                # transform.jl:42
    return t.mat[:3, :3].copy()


def translation(t: Transform):             # transform.jl:43
    return t.mat[:3, 3].copy()


def rpy(t: Transform):                     # transform.jl:45-48 (RotZYX): [roll, pitch, yaw]
    R = t.mat
    t1 = np.arctan2(R[1, 0], R[0, 0])
    s1, c1 = np.sin(t1), np.cos(t1)
    t2 = np.arctan2(-R[2, 0], np.sqrt(R[2, 1] ** 2 + R[2, 2] ** 2))
    t3 = np.arctan2(R[0, 2] * s1 - R[1, 2] * c1, R[1, 1] * c1 - R[0, 1] * s1)
    return np.array([t3, t2, t1])


def rotz(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
