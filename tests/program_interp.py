"""Test infrastructure: a numpy interpreter of the kinematic program that libkin_b200 compiles
(csrc/kin_program.h), following kin_kernels.cuh statement by statement but vectorised over the batch.
It lets the CPU suite check the HOST half of the product (URDF -> tables -> program) against the
oracle without a GPU.  It is never imported by the product."""
import ctypes as C

import numpy as np

from kinematics_jl_b200 import lib as L
from kinematics_jl_b200.device import make_desc

NODE_INTS, NODE_REALS, ATT_INTS, ATT_REALS, SPH_REALS, BOX_REALS = 12, 16, 4, 12, 4, 18
HEADER_FIELDS = ["n_nodes", "n_att", "n_sph", "n_box", "n_joints", "with_base", "n_dof", "n_fk", "n_jac",
                 "io_node", "io_att", "io_sph_order", "io_sph_mask", "io_col_type", "n_int",
                 "ro_node", "ro_att", "ro_sph", "ro_box", "n_real",
                 "so_q", "so_save", "so_jf", "so_cent", "so_stale", "n_slots", "so_q2"]


def dump_program(mech, ctrl_ids, fk_ids, jac_ids, spheres=None, boxes=None, want_stale=True):
    d, keep = make_desc(mech, ctrl_ids, spheres, boxes)
    fk = np.ascontiguousarray(fk_ids, dtype=np.int32)
    jac = np.ascontiguousarray(jac_ids, dtype=np.int32)
    hdr = np.zeros(64, dtype=np.int32)
    ints = np.zeros(1 << 16, dtype=np.int32)
    reals = np.zeros(1 << 16, dtype=np.float64)
    ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    want_coll = spheres is not None and boxes is not None
    L.check(L.lib().kin_program_dump(C.byref(d), fk.ctypes.data_as(ip), len(fk), jac.ctypes.data_as(ip), len(jac),
                                     int(want_coll), int(want_stale and want_coll), hdr.ctypes.data_as(ip), len(hdr),
                                     ints.ctypes.data_as(ip), len(ints), reals.ctypes.data_as(dp), len(reals)))
    h = dict(zip(HEADER_FIELDS, hdr[:len(HEADER_FIELDS)].tolist()))
    return h, ints[:h["n_int"]].copy(), reals[:h["n_real"]].copy()


def _mul_const(R, p, c, r_identity):
    Rc, tc = c[:9].reshape(3, 3), c[9:12]
    po = p + R @ tc
    return (R.copy() if r_identity else R @ Rc), po


def _box_sdf(b, P):
    Ri, ti, half = b[:9].reshape(3, 3), b[9:12], b[12:15]
    l = P @ Ri.T + ti
    q = np.abs(l) - half
    return np.linalg.norm(np.maximum(q, 0.0), axis=-1) + np.minimum(q.max(axis=-1), 0.0)


def run_program(h, ti, tr, Q, with_rot=True, rpy_jac=False, truncation=np.inf, grad_mode=0, scratch_ref=True,
                want_grads=True):
    """Q: (N, n_dof).  Returns dict with T (N, n_fk, 3, 4), J (N, n_jac, rows, cols), vals (N, S),
    grads (N, n_dof, S), argmin (N, S)."""
    N = Q.shape[0]
    ND = h["n_dof"]
    D = ND          # the planar base is compiled into three ordinary nodes: every column is a joint column
    rows = 6 if with_rot else 3
    T_out = np.zeros((N, h["n_fk"], 3, 4))
    J_out = np.zeros((N, h["n_jac"], rows, ND))
    save = {}
    jf_o, jf_a = np.zeros((N, max(D, 1), 3)), np.zeros((N, max(D, 1), 3))
    cent = np.zeros((N, max(h["n_sph"], 1), 3))
    col_type = ti[h["io_col_type"]:h["io_col_type"] + D]
    R = p = None
    for node in range(h["n_nodes"]):
        ni = ti[h["io_node"] + node * NODE_INTS: h["io_node"] + (node + 1) * NODE_INTS]
        nr = tr[h["ro_node"] + node * NODE_REALS: h["ro_node"] + (node + 1) * NODE_REALS]
        psrc, jtype, flags, qcol, save_slot, a0, a1, s0, s1, relmask = ni[:10]
        if jtype == 3:
            R = np.tile(np.eye(3), (N, 1, 1))
            p = np.zeros((N, 3))
        else:
            if psrc >= 0:
                R, p = save[psrc]
            Ra = R.copy() if flags & 1 else R @ nr[:9].reshape(3, 3)
            pa = p + R @ nr[9:12]
            code = (flags >> 1) & 7
            axis = nr[12:15]
            if code:
                k = (code - 1) % 3
                aw = Ra[:, :, k] * (-1.0 if code >= 4 else 1.0)
            else:
                aw = Ra @ axis
            jf_o[:, qcol], jf_a[:, qcol] = pa, aw
            qa = Q[:, qcol]
            if jtype == 2:
                R, p = Ra, pa + aw * qa[:, None]
            else:
                c, s = np.cos(qa), np.sin(qa)
                K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
                rot = (c[:, None, None] * np.eye(3) + s[:, None, None] * K
                       + (1 - c)[:, None, None] * np.outer(axis, axis))
                R, p = Ra @ rot, pa
        if save_slot >= 0:
            save[save_slot] = (R.copy(), p.copy())
        for a in range(a0, a1):
            ai = ti[h["io_att"] + a * ATT_INTS: h["io_att"] + (a + 1) * ATT_INTS]
            ar = tr[h["ro_att"] + a * ATT_REALS: h["ro_att"] + (a + 1) * ATT_REALS]
            Rl, pl = _mul_const(R, p, ar, ai[1] & 1)
            if ai[0] >= 0:
                T_out[:, ai[0], :, :3], T_out[:, ai[0], :, 3] = Rl, pl
            if ai[2] >= 0:
                mask = int(ai[3]) & 0xFFFFFFFF
                if with_rot and rpy_jac:
                    yaw = np.arctan2(Rl[:, 1, 0], Rl[:, 0, 0])
                    pitch = np.arctan2(-Rl[:, 2, 0], np.sqrt(Rl[:, 2, 1] ** 2 + Rl[:, 2, 2] ** 2))
                    s2, c2, s3, c3 = np.sin(-pitch), np.cos(-pitch), np.sin(-yaw), np.cos(-yaw)
                for j in range(D):
                    if not (mask >> j) & 1:
                        continue
                    a_w = jf_a[:, j]
                    if col_type[j] == 1:
                        J_out[:, ai[2], :3, j] = np.cross(a_w, pl - jf_o[:, j])
                        if with_rot:
                            if rpy_jac:
                                x, y, z = a_w[:, 0], a_w[:, 1], a_w[:, 2]
                                J_out[:, ai[2], 3, j] = c3 / c2 * x - s3 / c2 * y
                                J_out[:, ai[2], 4, j] = s3 * x + c3 * y
                                J_out[:, ai[2], 5, j] = -c3 * s2 / c2 * x + s3 * s2 / c2 * y + z
                            else:
                                J_out[:, ai[2], 3:, j] = a_w
                    else:
                        J_out[:, ai[2], :3, j] = a_w
        for k in range(s0, s1):
            s = ti[h["io_sph_order"] + k]
            sr = tr[h["ro_sph"] + s * SPH_REALS: h["ro_sph"] + (s + 1) * SPH_REALS]
            cent[:, s] = p + R @ sr[:3]
    out = {"T": T_out, "J": J_out}
    S, B = h["n_sph"], h["n_box"]
    if S and B:
        vals, grads, argmin = np.zeros((N, S)), np.zeros((N, ND, S)), np.zeros((N, S), dtype=np.int32)
        stale = np.zeros((N, max(D, 1), 3))
        boxes = [tr[h["ro_box"] + b * BOX_REALS: h["ro_box"] + (b + 1) * BOX_REALS] for b in range(B)]
        idx = np.arange(N)
        for s in range(S):
            P = cent[:, s]
            d_all = np.stack([_box_sdf(b, P) for b in boxes], axis=1)
            kmin = d_all.argmin(axis=1)              # first minimum
            dmin = d_all[idx, kmin]
            dist0 = dmin - tr[h["ro_sph"] + s * SPH_REALS + 3]
            argmin[:, s] = kmin + 1
            trunc = dist0 > truncation
            vals[:, s] = np.where(trunc, truncation, dist0)
            if not want_grads:
                continue
            g = np.zeros((N, 3))
            for b in range(B):
                sel = (kmin == b) & ~trunc
                if not sel.any():
                    continue
                if grad_mode == 0:
                    for i in range(3):
                        Pe = P[sel].copy()
                        Pe[:, i] += 1e-7
                        g[sel, i] = (_box_sdf(boxes[b], Pe) - dmin[sel]) / 1e-7
                else:
                    raise NotImplementedError
            mask = int(ti[h["io_sph_mask"] + s]) & 0xFFFFFFFF
            live = ~trunc
            for j in range(D):
                if (mask >> j) & 1:
                    colv = np.cross(jf_a[:, j], P - jf_o[:, j]) if col_type[j] == 1 else jf_a[:, j]
                    if scratch_ref:
                        stale[live, j] = colv[live]
                elif scratch_ref:
                    colv = stale[:, j]
                else:
                    colv = np.zeros((N, 3))
                grads[live, j, s] = np.einsum("ni,ni->n", g, colv)[live]
        out.update(vals=vals, grads=grads, argmin=argmin)
    return out
