// kin_program.h -- the "kinematic program" a KinModel is compiled into, shared by the host-side
// compiler (kin_model.cpp) and the kernels (kin_kernels.cu).
//
// Reference data structure it replaces: Mechanism{links, joints, rptable, tf_cache, axis_cache,
// tf_stack} (mechanism.jl:147-164).  The reference walks link -> root through boxed Link/Joint
// objects and memoises 4x4 transforms per link (algorithm.jl:1-37).  Here the tree is compiled
// once per (model, requested outputs) into a flat table:
//
//   * DYNAMIC NODES: a virtual root plus one node per link whose parent joint is driven by a
//     column of the configuration.  They form a small tree, stored in DFS pre-order so that a
//     node's parent transform is either still in registers (first child) or in a per-thread
//     scratch slot (later children of a branching node).
//   * ATTACHMENTS: every requested link is `T_node * C` with C a constant transform composed on
//     the host from the chain of fixed joints (and un-controlled joints frozen at their default
//     angle, honouring the a == 0.0 short-cut of mechanism.jl:95,101) between the link and its
//     nearest dynamic ancestor.
//   * SPHERES: collision spheres (collision.jl:39-49) are pure-translation attachments; their
//     centre is pre-composed into the frame of the dynamic node.
//   * BOXES: world inverse pose + half extents of each BoxSDF of the UnionSDF (sdf.jl:48-74).
//
// The blob is [int32 section][real section]; the real section is emitted in f64 or f32.
#pragma once
#include <stdint.h>

namespace kin {

enum { NODE_ROOT = 3 };                 // jtype of the virtual root node (others: KinJointType)
enum { PARENT_CUR = -2, PARENT_NONE = -1 };

// node flags
enum {
    NF_OFF_R_IDENTITY = 1,              // joint.pose has identity rotation
    NF_AXIS_SHIFT = 1, NF_AXIS_MASK = 7 // (flags >> 1) & 7: 0 general, 1..3 = +x,+y,+z, 4..6 = -x,-y,-z
};
// attachment flags
enum { AF_R_IDENTITY = 1, AF_T_ZERO = 2 };

constexpr int NODE_INTS = 12;   // parent_src, jtype, flags, qcol, save_slot, att_begin, att_end,
                                // sph_begin, sph_end, relmask, pad, pad
constexpr int NODE_REALS = 16;  // R_off[9] row-major, t_off[3], axis[3], pad
constexpr int ATT_INTS = 4;     // fk_index, flags, jac_index, relmask
constexpr int ATT_REALS = 12;   // C: R[9] row-major, t[3]
constexpr int SPH_REALS = 4;    // centre in node frame [3], radius
constexpr int BOX_REALS = 18;   // inv_R[9] row-major, inv_t[3], half[3], kind, rounding radius, pad.  Even stride from an even offset: an
                                // FP64 row is 8 aligned 16-byte pairs (8 LDS.128 instead of 15 LDS.64); when lanes
                                // read different rows in the gradient pass, row b starts at bank 4b mod 32, so up to 8
                                // boxes are conflict-free

// Passed BY VALUE as a kernel parameter (lives in the constant bank).
struct ProgHeader {
    int n_nodes, n_att, n_sph, n_box;
    int n_joints, with_base, n_dof;
    int n_fk, n_jac;
    // int32 offsets into the int section
    int io_node, io_att, io_sph_order, io_sph_mask, io_col_type;
    int n_int;                  // ints in the int section (padded to a multiple of 4)
    // element offsets into the real section
    int ro_node, ro_att, ro_sph, ro_box;
    int n_real;
    // per-thread scratch slots (in reals)
    int so_q, so_save, so_jf, so_cent, so_stale, n_slots;
    int so_q2;                  // second configuration buffer (the next tile is prefetched with cp.async)
};

}  // namespace kin
