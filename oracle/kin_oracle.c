/*
 * kin_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C, one-configuration-at-a-time restatement of the hot path of
 * HiroIshida/Kinematics.jl, written so that it follows the reference's own
 * control flow (lazy per-link memo, explicit stack walk link->root, full 4x4
 * products, forward-difference SDF gradient, one shared Jacobian scratch per
 * collision call).  Every function cites the reference file:line it follows
 * (paths relative to /root/reference/src).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (libkin_b200.so)
 * never links, imports or calls it.
 *
 * Parity pin: the reference itself (Julia + scikit-robot) cannot run in the
 * build container, so the oracle is pinned by the reference's own golden
 * vectors and known-answer tests: data/ground_truth.json (FK, test_kinematics.jl
 * :18-39), the box/union SDF KATs of test_sdf.jl:16-36 and the Jacobian /
 * collision-gradient finite-difference self checks of test_kinematics.jl:45-72,
 * test_collision.jl:33-43 -- see tests/test_oracle_golden.py.
 *
 * Arithmetic that lives in un-vendored third-party Julia packages and is
 * restated here from their published algorithm:
 *   Rotations.jl 1.0.2  : UnitQuaternion(w,x,y,z) (normalising ctor),
 *                         quaternion -> 3x3, RotZYX(::RotMatrix)
 *   StaticArrays 1.0.1  : 4x4 product, cross, norm
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -pthread).
 * -ffp-contract=off because Julia never fuses a*b+c unless asked to.
 *
 * Conventions: ids are 1-based exactly as load_urdf.jl:22-32 /
 * mechanism.jl:239-243 assign them; matrices are column-major like Julia.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define OR_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------ */
/* transform.jl                                                             */
/* ------------------------------------------------------------------------ */

/* transform.jl:3-5 -- 4x4 column-major SMatrix */
typedef struct { double m[16]; } Tf;

#define M(t, r, c) ((t).m[(c) * 4 + (r)])

static Tf tf_identity(void) { /* transform.jl:50-56 (zero(Transform) is the identity too) */
    Tf t; memset(&t, 0, sizeof t);
    M(t,0,0) = M(t,1,1) = M(t,2,2) = M(t,3,3) = 1.0;
    return t;
}

/* transform.jl:58-60 -- full 4x4 * 4x4 product (StaticArrays: row i, col j,
 * sum over k left to right) */
static Tf tf_mul(const Tf *a, const Tf *b) {
    Tf c;
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) {
            double s = M(*a,i,0) * M(*b,0,j);
            s = s + M(*a,i,1) * M(*b,1,j);
            s = s + M(*a,i,2) * M(*b,2,j);
            s = s + M(*a,i,3) * M(*b,3,j);
            M(c,i,j) = s;
        }
    return c;
}

/* Rotations.jl 1.0.2: UnitQuaternion(w,x,y,z) normalises, then
 * quaternion -> RotMatrix3 (restated; call sites mechanism.jl:96,
 * transform.jl:8,17,34). Output R column-major 3x3. */
static void quat_to_rot(double w, double x, double y, double z, double R[9]) {
    double inorm = 1.0 / sqrt(w*w + x*x + y*y + z*z);
    w *= inorm; x *= inorm; y *= inorm; z *= inorm;
    double xx = x*x, yy = y*y, zz = z*z;
    double xy = x*y, zw = w*z, xz = x*z, yw = y*w, yz = y*z, xw = w*x;
    R[0] = 1 - 2*(yy + zz); R[1] = 2*(xy + zw);     R[2] = 2*(xz - yw);
    R[3] = 2*(xy - zw);     R[4] = 1 - 2*(xx + zz); R[5] = 2*(yz + xw);
    R[6] = 2*(xz + yw);     R[7] = 2*(yz - xw);     R[8] = 1 - 2*(xx + yy);
}

/* transform.jl:7-14 */
static Tf tf_from_trans_rot(const double t[3], const double R[9]) {
    Tf o = tf_identity();
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) M(o,r,c) = R[c*3+r];
    M(o,0,3) = t[0]; M(o,1,3) = t[1]; M(o,2,3) = t[2];
    return o;
}
/* transform.jl:16-23 */
static Tf tf_from_rot(const double R[9]) { double z[3] = {0,0,0}; return tf_from_trans_rot(z, R); }
/* transform.jl:25-31 */
static Tf tf_from_trans(const double t[3]) {
    Tf o = tf_identity(); M(o,0,3) = t[0]; M(o,1,3) = t[1]; M(o,2,3) = t[2]; return o;
}
/* transform.jl:33-37 */
static Tf base_pose_to_transform(const double pose[3]) {
    double R[9]; quat_to_rot(cos(0.5 * pose[2]), 0.0, 0.0, sin(0.5 * pose[2]), R);
    double t[3] = {pose[0], pose[1], 0.0};
    return tf_from_trans_rot(t, R);
}
/* transform.jl:42-43 */
static void tf_translation(const Tf *t, double p[3]) { p[0] = M(*t,0,3); p[1] = M(*t,1,3); p[2] = M(*t,2,3); }
/* transform.jl:39-41 -- T*point = translation + rotation*point */
static void tf_apply(const Tf *t, const double p[3], double out[3]) {
    for (int r = 0; r < 3; ++r) {
        double s = M(*t,r,0)*p[0];
        s = s + M(*t,r,1)*p[1];
        s = s + M(*t,r,2)*p[2];
        out[r] = M(*t,r,3) + s;
    }
}
/* rotation(t) * v */
static void tf_rotate(const Tf *t, const double v[3], double out[3]) {
    for (int r = 0; r < 3; ++r) {
        double s = M(*t,r,0)*v[0];
        s = s + M(*t,r,1)*v[1];
        s = s + M(*t,r,2)*v[2];
        out[r] = s;
    }
}
/* transform.jl:62-65 -- rigid inverse: (-R' t, R') */
static Tf tf_inv(const Tf *t) {
    double Ri[9], p[3], q[3];
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) Ri[c*3+r] = M(*t,c,r);
    tf_translation(t, p);
    for (int r = 0; r < 3; ++r) {
        double s = (-Ri[0*3+r])*p[0];
        s = s + (-Ri[1*3+r])*p[1];
        s = s + (-Ri[2*3+r])*p[2];
        q[r] = s;
    }
    return tf_from_trans_rot(q, Ri);
}
/* transform.jl:45-48 -- rpy(t) = [theta3, theta2, theta1] of RotZYX(R)
 * (Rotations.jl 1.0.2 RotZYX(::RotMatrix), restated) -> [roll, pitch, yaw] */
static void tf_rpy(const Tf *t, double out[3]) {
    double t1 = atan2(M(*t,1,0), M(*t,0,0));
    double st1 = sin(t1), ct1 = cos(t1);
    double t2 = atan2(-M(*t,2,0), sqrt(M(*t,2,1)*M(*t,2,1) + M(*t,2,2)*M(*t,2,2)));
    double t3 = atan2(M(*t,0,2)*st1 - M(*t,1,2)*ct1, M(*t,1,1)*ct1 - M(*t,0,1)*st1);
    out[0] = t3; out[1] = t2; out[2] = t1;
}

/* ------------------------------------------------------------------------ */
/* mechanism.jl                                                             */
/* ------------------------------------------------------------------------ */
enum { JT_FIXED = 0, JT_REVOLUTE = 1, JT_PRISMATIC = 2 };

typedef struct {              /* mechanism.jl:74-88 */
    int id, plink_id, clink_id, type;
    Tf pose;
    double axis[3];
} Joint;

typedef struct {              /* mechanism.jl:35-49 (ids only) */
    int id, pjoint_id, plink_id;
    int n_child, cap_child;
    int *clink_ids;
} Link;

typedef struct {              /* mechanism.jl:147-164 */
    int n_links, n_joints, cap;
    Link *links; Joint *joints;
    Tf *tf_cache; unsigned char *tf_valid;          /* cache.jl:1-37 */
    double *axis_cache; unsigned char *axis_valid;  /* FloatingAxis: origin[3], axis[3] */
    double *angles;
    double base_pose[3];
    int with_base;
    unsigned char *rptable;                         /* [joint][link], mechanism.jl:117-139 */
    int *link_id_stack; Tf *tf_stack; int top;      /* stack.jl:1-25 */
    long n_tf_mul;                                  /* instrumentation only */
} Mech;

static void mech_reserve(Mech *m, int cap) {
    if (cap <= m->cap) return;
    m->links = realloc(m->links, sizeof(Link) * cap);
    m->joints = realloc(m->joints, sizeof(Joint) * cap);
    m->tf_cache = realloc(m->tf_cache, sizeof(Tf) * cap);
    m->tf_valid = realloc(m->tf_valid, cap);
    m->axis_cache = realloc(m->axis_cache, sizeof(double) * 6 * cap);
    m->axis_valid = realloc(m->axis_valid, cap);
    m->angles = realloc(m->angles, sizeof(double) * cap);
    m->link_id_stack = realloc(m->link_id_stack, sizeof(int) * cap);
    m->tf_stack = realloc(m->tf_stack, sizeof(Tf) * cap);
    m->cap = cap;
}

/* mechanism.jl:270 */
static void invalidate_cache(Mech *m) {
    memset(m->tf_valid, 0, m->n_links);
    memset(m->axis_valid, 0, m->n_joints);
}

/* mechanism.jl:117-139 */
static void rp_recurse(Mech *m, int joint_id, int link_id) {
    m->rptable[(size_t)(joint_id - 1) * m->n_links + (link_id - 1)] = 1;
    Link *l = &m->links[link_id - 1];
    for (int k = 0; k < l->n_child; ++k) rp_recurse(m, joint_id, l->clink_ids[k]);
}
static void create_rptable(Mech *m) {
    free(m->rptable);
    m->rptable = calloc((size_t)m->n_joints * m->n_links, 1);
    for (int j = 0; j < m->n_joints; ++j) rp_recurse(m, m->joints[j].id, m->joints[j].clink_id);
}

static void link_push_child(Link *l, int id) {
    if (l->n_child == l->cap_child) {
        l->cap_child = l->cap_child ? 2 * l->cap_child : 4;
        l->clink_ids = realloc(l->clink_ids, sizeof(int) * l->cap_child);
    }
    l->clink_ids[l->n_child++] = id;
}

/* load_urdf.jl:20-80 (the wiring half; the XML half lives in oracle/ref_model.py)
 * joint arrays are in id order (1..n_joints). pose is 4x4 column-major. */
OR_EXPORT Mech *or_mech_create(int n_links, int n_joints, const int *j_plink, const int *j_clink,
                               const int *j_type, const double *j_pose16, const double *j_axis3,
                               int with_base) {
    Mech *m = calloc(1, sizeof(Mech));
    mech_reserve(m, (n_links > n_joints ? n_links : n_joints) + 64);
    m->n_links = n_links; m->n_joints = n_joints; m->with_base = with_base;
    for (int i = 0; i < n_links; ++i) {
        Link *l = &m->links[i];
        memset(l, 0, sizeof *l);
        l->id = i + 1; l->pjoint_id = -1; l->plink_id = -1;
    }
    for (int j = 0; j < n_joints; ++j) {
        Joint *jt = &m->joints[j];
        jt->id = j + 1; jt->plink_id = j_plink[j]; jt->clink_id = j_clink[j]; jt->type = j_type[j];
        memcpy(jt->pose.m, j_pose16 + 16 * j, sizeof(double) * 16);
        memcpy(jt->axis, j_axis3 + 3 * j, sizeof(double) * 3);
        link_push_child(&m->links[jt->plink_id - 1], jt->clink_id);   /* load_urdf.jl:69-71 */
        m->links[jt->clink_id - 1].pjoint_id = jt->id;                /* load_urdf.jl:73-75 */
        m->links[jt->clink_id - 1].plink_id = jt->plink_id;
        m->angles[j] = 0.0;
    }
    create_rptable(m);
    invalidate_cache(m);
    return m;
}

OR_EXPORT void or_mech_destroy(Mech *m) {
    if (!m) return;
    for (int i = 0; i < m->n_links; ++i) free(m->links[i].clink_ids);
    free(m->links); free(m->joints); free(m->tf_cache); free(m->tf_valid);
    free(m->axis_cache); free(m->axis_valid); free(m->angles); free(m->rptable);
    free(m->link_id_stack); free(m->tf_stack); free(m);
}

static Mech *mech_clone(const Mech *s) {
    Mech *m = calloc(1, sizeof(Mech));
    mech_reserve(m, s->cap);
    m->n_links = s->n_links; m->n_joints = s->n_joints; m->with_base = s->with_base;
    memcpy(m->joints, s->joints, sizeof(Joint) * s->n_joints);
    memcpy(m->angles, s->angles, sizeof(double) * s->n_joints);
    memcpy(m->base_pose, s->base_pose, sizeof m->base_pose);
    for (int i = 0; i < s->n_links; ++i) {
        m->links[i] = s->links[i];
        m->links[i].clink_ids = malloc(sizeof(int) * (s->links[i].cap_child ? s->links[i].cap_child : 1));
        memcpy(m->links[i].clink_ids, s->links[i].clink_ids, sizeof(int) * s->links[i].n_child);
    }
    m->rptable = malloc((size_t)s->n_joints * s->n_links);
    memcpy(m->rptable, s->rptable, (size_t)s->n_joints * s->n_links);
    invalidate_cache(m);
    return m;
}

/* mechanism.jl:233-267 -- append a link under `parent` through a Fixed joint */
OR_EXPORT int or_add_new_link(Mech *m, int parent_link_id, const double *pose16) {
    int need = (m->n_links > m->n_joints ? m->n_links : m->n_joints) + 2;
    if (need > m->cap) mech_reserve(m, 2 * need);
    int hlink_id = m->n_links + 1;
    int joint_id = m->n_joints + 1;
    link_push_child(&m->links[parent_link_id - 1], hlink_id);
    Joint *jt = &m->joints[m->n_joints];
    jt->id = joint_id; jt->plink_id = parent_link_id; jt->clink_id = hlink_id; jt->type = JT_FIXED;
    memcpy(jt->pose.m, pose16, sizeof(double) * 16);
    jt->axis[0] = jt->axis[1] = jt->axis[2] = 0.0;
    Link *l = &m->links[m->n_links];
    memset(l, 0, sizeof *l);
    l->id = hlink_id; l->pjoint_id = joint_id; l->plink_id = parent_link_id;
    m->angles[m->n_joints] = 0.0;
    m->n_links++; m->n_joints++;
    create_rptable(m);
    invalidate_cache(m);
    return hlink_id;
}
/* mechanism.jl:233-236 -- position-only overload */
OR_EXPORT int or_add_new_link_pos(Mech *m, int parent_link_id, const double *pos3) {
    Tf t = tf_from_trans(pos3);
    return or_add_new_link(m, parent_link_id, t.m);
}

/* mechanism.jl:223-231 */
OR_EXPORT void or_set_joint_angles(Mech *m, const int *joint_ids, int n_joints, const double *angles) {
    for (int i = 0; i < n_joints; ++i) m->angles[joint_ids[i] - 1] = angles[i];
    if (m->with_base) { m->base_pose[0] = angles[n_joints]; m->base_pose[1] = angles[n_joints+1]; m->base_pose[2] = angles[n_joints+2]; }
    invalidate_cache(m);
}
OR_EXPORT int or_is_relevant(const Mech *m, int joint_id, int link_id) { /* mechanism.jl:277 */
    return m->rptable[(size_t)(joint_id - 1) * m->n_links + (link_id - 1)];
}
OR_EXPORT int or_n_links(const Mech *m) { return m->n_links; }
OR_EXPORT int or_n_joints(const Mech *m) { return m->n_joints; }
OR_EXPORT int or_link_parent(const Mech *m, int link_id) { return m->links[link_id - 1].plink_id; }
OR_EXPORT int or_link_n_children(const Mech *m, int link_id) { return m->links[link_id - 1].n_child; }
OR_EXPORT int or_link_child(const Mech *m, int link_id, int k) { return m->links[link_id - 1].clink_ids[k]; }

/* mechanism.jl:90-103 */
static Tf joint_transform(const Joint *j, double angle) {
    if (j->type == JT_FIXED) return j->pose;
    if (angle == 0.0) return j->pose;                     /* the a==0.0 short-circuit */
    if (j->type == JT_REVOLUTE) {
        double R[9], s = sin(0.5 * angle);
        quat_to_rot(cos(0.5 * angle), j->axis[0]*s, j->axis[1]*s, j->axis[2]*s, R);
        Tf r = tf_from_rot(R);
        return tf_mul(&j->pose, &r);
    } else {
        double t[3] = {j->axis[0]*angle, j->axis[1]*angle, j->axis[2]*angle};
        Tf r = tf_from_trans(t);
        return tf_mul(&j->pose, &r);
    }
}

/* ------------------------------------------------------------------------ */
/* algorithm.jl                                                             */
/* ------------------------------------------------------------------------ */

/* algorithm.jl:23-37 */
static Tf get_shallowest_cache(Mech *m, int hlink_id) {
    while (m->links[hlink_id - 1].plink_id != -1) {
        if (m->tf_valid[hlink_id - 1]) return m->tf_cache[hlink_id - 1];
        const Link *h = &m->links[hlink_id - 1];
        const Joint *hj = &m->joints[h->pjoint_id - 1];
        double angle = m->angles[hj->id - 1];
        m->tf_stack[m->top] = joint_transform(hj, angle);
        m->link_id_stack[m->top] = hlink_id;
        m->top++;
        hlink_id = h->plink_id;
    }
    return m->with_base ? base_pose_to_transform(m->base_pose) : tf_identity();
}

/* algorithm.jl:1-21 */
static Tf get_transform(Mech *m, int link_id) {
    if (m->tf_valid[link_id - 1]) return m->tf_cache[link_id - 1];
    Tf tf = get_shallowest_cache(m, link_id);
    while (m->top > 0) {
        m->top--;
        int hid = m->link_id_stack[m->top];
        tf = tf_mul(&tf, &m->tf_stack[m->top]);
        m->n_tf_mul++;
        m->tf_cache[hid - 1] = tf; m->tf_valid[hid - 1] = 1;
    }
    return tf;
}
OR_EXPORT void or_get_transform(Mech *m, int link_id, double *out16) {
    Tf t = get_transform(m, link_id); memcpy(out16, t.m, sizeof t.m);
}
OR_EXPORT void or_rpy(const double *tf16, double *out3) { Tf t; memcpy(t.m, tf16, sizeof t.m); tf_rpy(&t, out3); }

/* algorithm.jl:42-54 -- world joint origin and axis, memoised */
static const double *get_joint_axis(Mech *m, const Joint *hj) {
    double *fa = m->axis_cache + 6 * (hj->id - 1);
    if (m->axis_valid[hj->id - 1]) return fa;
    Tf wp = get_transform(m, hj->plink_id);
    Tf wj = tf_mul(&wp, &hj->pose);
    tf_translation(&wj, fa);
    tf_rotate(&wj, hj->axis, fa + 3);
    m->axis_valid[hj->id - 1] = 1;
    return fa;
}

/* algorithm.jl:56-63 */
static void rpy_derivative(const double rpy[3], const double axis[3], double out[3]) {
    double a2 = -rpy[1], a3 = -rpy[2];
    double x = axis[0], y = axis[1], z = axis[2];
    out[0] = cos(a3)/cos(a2)*x - sin(a3)/cos(a2)*y;
    out[1] = sin(a3)*x + cos(a3)*y;
    out[2] = -cos(a3)*sin(a2)/cos(a2)*x + sin(a3)*sin(a2)/cos(a2)*y + z;
}

/* algorithm.jl:65-81 -- one column; col points at `rows` doubles */
static void joint_jacobian(Mech *m, const Joint *j, const Tf *tf_link, int with_rot, int rpy_jac, double *col) {
    const double *fa = get_joint_axis(m, j);
    if (j->type == JT_REVOLUTE) {
        double p[3], d[3];
        tf_translation(tf_link, p);
        d[0] = p[0] - fa[0]; d[1] = p[1] - fa[1]; d[2] = p[2] - fa[2];
        const double *a = fa + 3;
        col[0] = a[1]*d[2] - a[2]*d[1];
        col[1] = a[2]*d[0] - a[0]*d[2];
        col[2] = a[0]*d[1] - a[1]*d[0];
        if (with_rot) {
            if (rpy_jac) { double r[3]; tf_rpy(tf_link, r); rpy_derivative(r, a, col + 3); }
            else { col[3] = a[0]; col[4] = a[1]; col[5] = a[2]; }
        }
    } else { /* prismatic: rows 4:6 are left untouched, algorithm.jl:78-81 */
        col[0] = fa[3]; col[1] = fa[4]; col[2] = fa[5];
    }
}

/* algorithm.jl:83-106 -- writes ONLY relevant columns; mat is rows x (n+3?) column-major */
static void get_jacobian_(Mech *m, int link_id, const int *joint_ids, int n_joint, int with_rot, int rpy_jac, double *mat) {
    int rows = with_rot ? 6 : 3;
    Tf tf = get_transform(m, link_id);
    for (int i = 0; i < n_joint; ++i) {
        const Joint *j = &m->joints[joint_ids[i] - 1];
        if (or_is_relevant(m, j->id, link_id))
            joint_jacobian(m, j, &tf, with_rot, rpy_jac, mat + (size_t)rows * i);
    }
    if (m->with_base) {
        double x = M(tf,0,3) - m->base_pose[0], y = M(tf,1,3) - m->base_pose[1];
        double *b = mat + (size_t)rows * n_joint;
        b[0] = 1.0; b[1] = 0.0; b[2] = 0.0;
        b[rows+0] = 0.0; b[rows+1] = 1.0; b[rows+2] = 0.0;
        b[2*rows+0] = -y; b[2*rows+1] = x; b[2*rows+2] = 0.0;
        if (with_rot) {
            for (int c = 0; c < 3; ++c) for (int r = 3; r < 6; ++r) b[c*rows + r] = 0.0;
            b[2*rows+5] = 1.0;
        }
    }
}
/* get_jacobian! -- caller-owned scratch, NOT cleared (algorithm.jl:83) */
OR_EXPORT void or_get_jacobian_inplace(Mech *m, int link_id, const int *joint_ids, int n_joint,
                                       int with_rot, int rpy_jac, double *mat) {
    get_jacobian_(m, link_id, joint_ids, n_joint, with_rot, rpy_jac, mat);
}
/* get_jacobian -- zero-initialised result (algorithm.jl:108-114) */
OR_EXPORT void or_get_jacobian(Mech *m, int link_id, const int *joint_ids, int n_joint,
                               int with_rot, int rpy_jac, double *mat) {
    int rows = with_rot ? 6 : 3, cols = n_joint + (m->with_base ? 3 : 0);
    memset(mat, 0, sizeof(double) * rows * cols);
    get_jacobian_(m, link_id, joint_ids, n_joint, with_rot, rpy_jac, mat);
}

/* ------------------------------------------------------------------------ */
/* sdf.jl                                                                   */
/* ------------------------------------------------------------------------ */
typedef struct {           /* sdf.jl:48-56 (stand-alone boxes; attached boxes are
                              resolved to a world pose by the caller, sdf.jl:14-32) */
    Tf pose, inv_pose;
    double width[3];
    double val_cache;
    int kind;              /* 0 = box (the only primitive of the reference, load_urdf.jl:10-15, sdf.jl:92-94);
                              EXTENSION (no reference counterpart, textbook formulas): 1 = sphere,
                              width[0] = radius; 2 = cylinder along the local z axis, width[0] = radius,
                              width[1] = length.  They plug into the same AbstractSDF contract: value at p,
                              generic forward-difference gradient! (sdf.jl:34-41), UnionSDF argmin. */
} Box;
typedef struct {           /* sdf.jl:76-80 */
    int n; Box *boxes; double *vals_cache; int min_idx_cache;   /* 1-based argmin */
} Sdf;

OR_EXPORT Sdf *or_sdf_create(int n_boxes, const double *pose16, const double *width3) {
    Sdf *s = calloc(1, sizeof(Sdf));
    s->n = n_boxes; s->boxes = calloc(n_boxes, sizeof(Box)); s->vals_cache = calloc(n_boxes, sizeof(double));
    for (int i = 0; i < n_boxes; ++i) {
        memcpy(s->boxes[i].pose.m, pose16 + 16*i, sizeof(double)*16);
        s->boxes[i].inv_pose = tf_inv(&s->boxes[i].pose);       /* sdf.jl:58-61 */
        memcpy(s->boxes[i].width, width3 + 3*i, sizeof(double)*3);
    }
    return s;
}
/* union of mixed primitives (extension, see Box.kind): size3 = widths | (radius, -, -) | (radius, length, -) */
OR_EXPORT Sdf *or_sdf_create_prims(int n, const int *kinds, const double *pose16, const double *size3) {
    Sdf *s = or_sdf_create(n, pose16, size3);
    for (int i = 0; i < n; ++i) s->boxes[i].kind = kinds ? kinds[i] : 0;
    return s;
}
OR_EXPORT void or_sdf_destroy(Sdf *s) { if (s) { free(s->boxes); free(s->vals_cache); free(s); } }
static Sdf *sdf_clone(const Sdf *s) {
    Sdf *c = calloc(1, sizeof(Sdf)); c->n = s->n;
    c->boxes = malloc(sizeof(Box) * s->n); memcpy(c->boxes, s->boxes, sizeof(Box) * s->n);
    c->vals_cache = calloc(s->n, sizeof(double));
    return c;
}

/* sdf.jl:67-74 */
static double box_eval(Box *b, const double p[3], int do_cache) {
    double q[3], pl[3];
    tf_apply(&b->inv_pose, p, pl);
    double d;
    if (b->kind == 1) {            /* sphere: |p - c| - r */
        d = sqrt(pl[0]*pl[0] + pl[1]*pl[1] + pl[2]*pl[2]) - b->width[0];
    } else if (b->kind == 2) {     /* capped cylinder: the box formula on (radial, axial) */
        double q0 = sqrt(pl[0]*pl[0] + pl[1]*pl[1]) - b->width[0], q1 = fabs(pl[2]) - 0.5 * b->width[1];
        double m0 = fmax(q0, 0.0), m1 = fmax(q1, 0.0);
        d = sqrt(m0*m0 + m1*m1) + fmin(fmax(q0, q1), 0.0);
    } else {
        for (int i = 0; i < 3; ++i) q[i] = fabs(pl[i]) - 0.5 * b->width[i];
        double m0 = fmax(q[0], 0.0), m1 = fmax(q[1], 0.0), m2 = fmax(q[2], 0.0);
        double nrm = sqrt(m0*m0 + m1*m1 + m2*m2);
        double mx = fmax(fmax(q[0], q[1]), q[2]);
        d = nrm + fmin(mx, 0.0);
    }
    if (do_cache) b->val_cache = d;
    return d;
}
/* sdf.jl:108-114 -- evaluates ALL children, argmin = first minimum */
static double sdf_eval(Sdf *s, const double p[3]) {
    for (int i = 0; i < s->n; ++i) s->vals_cache[i] = box_eval(&s->boxes[i], p, 1);
    int k = 0;
    for (int i = 1; i < s->n; ++i) if (s->vals_cache[i] < s->vals_cache[k]) k = i;
    s->min_idx_cache = k + 1;
    return s->vals_cache[k];
}
/* sdf.jl:116-119 + sdf.jl:34-41 -- forward difference, eps 1e-7, argmin child only,
 * differenced against that child's cached value */
static void sdf_gradient(Sdf *s, const double p[3], double g[3]) {
    const double eps = 1e-7;
    Box *b = &s->boxes[s->min_idx_cache - 1];
    for (int i = 0; i < 3; ++i) {
        double tmp[3] = {p[0], p[1], p[2]};
        tmp[i] += eps;
        g[i] = (box_eval(b, tmp, 0) - b->val_cache) / eps;
    }
}
OR_EXPORT double or_sdf_eval(Sdf *s, const double *p3) { return sdf_eval(s, p3); }
OR_EXPORT int or_sdf_argmin(const Sdf *s) { return s->min_idx_cache; }
OR_EXPORT void or_sdf_gradient(Sdf *s, const double *p3, double *g3) { sdf_gradient(s, p3, g3); }

/* Closed-form gradient of the argmin box (NOT in the reference; the oracle's
 * "analytic" mode used to pin the product's optional analytic mode). */
static void sdf_gradient_analytic(Sdf *s, const double p[3], double g[3]) {
    Box *b = &s->boxes[s->min_idx_cache - 1];
    double pl[3], q[3], gl[3] = {0,0,0};
    tf_apply(&b->inv_pose, p, pl);
    if (b->kind == 1) {
        double nrm = sqrt(pl[0]*pl[0] + pl[1]*pl[1] + pl[2]*pl[2]);
        if (nrm > 0.0) for (int i = 0; i < 3; ++i) gl[i] = pl[i] / nrm;
        else gl[0] = 1.0;      /* exactly at the centre the gradient is undefined: the +x axis of the primitive frame */
    } else if (b->kind == 2) {
        double rxy = sqrt(pl[0]*pl[0] + pl[1]*pl[1]);
        double q0 = rxy - b->width[0], q1 = fabs(pl[2]) - 0.5 * b->width[1];
        double m0 = fmax(q0, 0.0), m1 = fmax(q1, 0.0), nrm = sqrt(m0*m0 + m1*m1);
        double ux = rxy > 0.0 ? pl[0] / rxy : 0.0, uy = rxy > 0.0 ? pl[1] / rxy : 0.0, sz = pl[2] < 0 ? -1.0 : 1.0;
        if (nrm > 0.0) { gl[0] = m0 / nrm * ux; gl[1] = m0 / nrm * uy; gl[2] = m1 / nrm * sz; }
        else if (q0 > q1) { gl[0] = ux; gl[1] = uy; }
        else gl[2] = sz;
    } else {
    for (int i = 0; i < 3; ++i) q[i] = fabs(pl[i]) - 0.5 * b->width[i];
    double m[3] = {fmax(q[0],0.0), fmax(q[1],0.0), fmax(q[2],0.0)};
    double nrm = sqrt(m[0]*m[0] + m[1]*m[1] + m[2]*m[2]);
    if (nrm > 0.0) {
        for (int i = 0; i < 3; ++i) gl[i] = (m[i] / nrm) * (pl[i] < 0 ? -1.0 : 1.0);
    } else {
        int k = 0; if (q[1] > q[k]) k = 1; if (q[2] > q[k]) k = 2;
        gl[k] = pl[k] < 0 ? -1.0 : 1.0;
    }
    }
    /* world gradient = R * g_local  (inv_pose rotation is R') */
    for (int r = 0; r < 3; ++r) {
        double sum = M(b->inv_pose,0,r)*gl[0];
        sum = sum + M(b->inv_pose,1,r)*gl[1];
        sum = sum + M(b->inv_pose,2,r)*gl[2];
        g[r] = sum;
    }
}
OR_EXPORT void or_sdf_gradient_analytic(Sdf *s, const double *p3, double *g3) { sdf_gradient_analytic(s, p3, g3); }

/* ------------------------------------------------------------------------ */
/* collision.jl                                                             */
/* ------------------------------------------------------------------------ */

/* collision.jl:51-58 */
OR_EXPORT void or_compute_coll_dists(Mech *m, const int *sphere_link_ids, const double *radii, int n_col,
                                     Sdf *sdf, double *out_vals, int *out_argmin) {
    for (int i = 0; i < n_col; ++i) {
        Tf t = get_transform(m, sphere_link_ids[i]);
        double pt[3]; tf_translation(&t, pt);
        out_vals[i] = sdf_eval(sdf, pt) - radii[i];
        if (out_argmin) out_argmin[i] = sdf->min_idx_cache;
    }
}

/* collision.jl:67-94.
 * grad_mode   0 = forward difference (reference), 1 = analytic (oracle extension)
 * scratch_mode 0 = reference: ONE 3 x n_dof scratch, zeroed once per call and
 *                  reused across spheres so non-relevant columns keep whatever the
 *                  previous non-truncated sphere left there (collision.jl:76,90 +
 *                  algorithm.jl:91-96);
 *              1 = clean: scratch zeroed before every sphere.
 * out_grads is (n_dof, n_col) column-major (dof fastest). */
OR_EXPORT void or_compute_coll_dists_and_grads(Mech *m, const int *joint_ids, int n_joint,
                                               const int *sphere_link_ids, const double *radii, int n_col,
                                               Sdf *sdf, double truncation_dist, int grad_mode, int scratch_mode,
                                               double *out_vals, double *out_grads, int *out_argmin) {
    int n_dof = n_joint + (m->with_base ? 3 : 0);
    double grad[3];
    double *jac = calloc((size_t)3 * n_dof, sizeof(double));
    for (int i = 0; i < n_col; ++i) {
        int link = sphere_link_ids[i];
        Tf t = get_transform(m, link);
        double pt0[3]; tf_translation(&t, pt0);
        double dist0 = sdf_eval(sdf, pt0) - radii[i];
        if (out_argmin) out_argmin[i] = sdf->min_idx_cache;
        if (dist0 > truncation_dist) {
            out_vals[i] = truncation_dist;
            for (int d = 0; d < n_dof; ++d) out_grads[(size_t)i * n_dof + d] = 0.0;
        } else {
            out_vals[i] = dist0;
            if (grad_mode == 0) sdf_gradient(sdf, pt0, grad); else sdf_gradient_analytic(sdf, pt0, grad);
            if (scratch_mode == 1) memset(jac, 0, sizeof(double) * 3 * n_dof);
            get_jacobian_(m, link, joint_ids, n_joint, 0, 0, jac);
            for (int d = 0; d < n_dof; ++d) {      /* transpose(grad) * jac */
                double s = grad[0] * jac[3*d+0];
                s = s + grad[1] * jac[3*d+1];
                s = s + grad[2] * jac[3*d+2];
                out_grads[(size_t)i * n_dof + d] = s;
            }
        }
    }
    free(jac);
}

/* ------------------------------------------------------------------------ */
/* callers: inverse_kinematics.jl:38-50, planning.jl:55-68,114-138           */
/* ------------------------------------------------------------------------ */

/* inverse_kinematics.jl:38-50 -- f = sum(pose_diff^2), grad = -2 J' pose_diff */
OR_EXPORT double or_ik_objective(Mech *m, int link_id, const int *joint_ids, int n_joint, const double *angles,
                                 const double *target16, int with_rot, double *grad_out) {
    int n_dof = n_joint + (m->with_base ? 3 : 0), rows = with_rot ? 6 : 3;
    or_set_joint_angles(m, joint_ids, n_joint, angles);
    Tf now = get_transform(m, link_id), tgt; memcpy(tgt.m, target16, sizeof tgt.m);
    double e[6], r0[3], r1[3];
    e[0] = M(tgt,0,3) - M(now,0,3); e[1] = M(tgt,1,3) - M(now,1,3); e[2] = M(tgt,2,3) - M(now,2,3);
    tf_rpy(&tgt, r1); tf_rpy(&now, r0);
    e[3] = r1[0] - r0[0]; e[4] = r1[1] - r0[1]; e[5] = r1[2] - r0[2];
    double *jac = calloc((size_t)rows * n_dof, sizeof(double));   /* zero(SizedMatrix), :36 */
    get_jacobian_(m, link_id, joint_ids, n_joint, with_rot, 1, jac);
    double f = 0.0;
    for (int r = 0; r < rows; ++r) f += e[r] * e[r];
    if (grad_out) for (int d = 0; d < n_dof; ++d) {
        double s = 0.0;
        for (int r = 0; r < rows; ++r) s += jac[(size_t)rows * d + r] * e[r];
        grad_out[d] = -2.0 * s;
    }
    free(jac);
    return f;
}

/* planning.jl:114-138 for ONE (link, target) pair: val = [p - p_t; rpy - rpy_t],
 * jac_T is (n_dof, dim) column-major = transpose of the rpy-Jacobian */
OR_EXPORT void or_pose_constraint(Mech *m, int link_id, const int *joint_ids, int n_joint, const double *q,
                                  const double *target16, int with_rot, double *val, double *jac_T) {
    int n_dof = n_joint + (m->with_base ? 3 : 0), rows = with_rot ? 6 : 3;
    or_set_joint_angles(m, joint_ids, n_joint, q);
    Tf now = get_transform(m, link_id), tgt; memcpy(tgt.m, target16, sizeof tgt.m);
    double r0[3], r1[3];
    val[0] = M(now,0,3) - M(tgt,0,3); val[1] = M(now,1,3) - M(tgt,1,3); val[2] = M(now,2,3) - M(tgt,2,3);
    if (with_rot) { tf_rpy(&now, r0); tf_rpy(&tgt, r1); val[3] = r0[0]-r1[0]; val[4] = r0[1]-r1[1]; val[5] = r0[2]-r1[2]; }
    double *jac = calloc((size_t)rows * n_dof, sizeof(double));
    get_jacobian_(m, link_id, joint_ids, n_joint, with_rot, 1, jac);
    for (int d = 0; d < n_dof; ++d) for (int r = 0; r < rows; ++r) jac_T[(size_t)r * n_dof + d] = jac[(size_t)rows * d + r];
    free(jac);
}

/* planning.jl:55-68 -- per-waypoint collision stack.  xi is (n_dof, n_wp)
 * column-major.  val_vec has n_coll*n_wp entries (dists - margin);
 * jac_blocks holds ONLY the n_wp diagonal blocks, each (n_dof, n_coll)
 * column-major (the reference scatters them into a dense zero matrix). */
OR_EXPORT void or_ineq_const(Mech *m, const int *joint_ids, int n_joint, const int *sphere_link_ids,
                             const double *radii, int n_coll, Sdf *sdf, const double *xi, int n_wp,
                             double margin, int grad_mode, int scratch_mode, double *val_vec, double *jac_blocks) {
    int n_dof = n_joint + (m->with_base ? 3 : 0);
    double truncation_dist = margin + 0.05;
    for (int i = 0; i < n_wp; ++i) {
        or_set_joint_angles(m, joint_ids, n_joint, xi + (size_t)n_dof * i);
        or_compute_coll_dists_and_grads(m, joint_ids, n_joint, sphere_link_ids, radii, n_coll, sdf,
                                        truncation_dist, grad_mode, scratch_mode,
                                        val_vec + (size_t)n_coll * i, jac_blocks + (size_t)n_dof * n_coll * i, NULL);
        for (int k = 0; k < n_coll; ++k) val_vec[(size_t)n_coll * i + k] -= margin;
    }
}

/* ------------------------------------------------------------------------ */
/* batch drivers (tests + CPU baseline).  q is [N][n_dof] (one record per     */
/* configuration, i.e. Julia (n_dof, N)).  One Mech/Sdf clone per thread      */
/* because the reference's scratch is not shareable.  Threads    */
/* are plain pthreads over contiguous slices of the batch.                    */
/* ------------------------------------------------------------------------ */
#include <pthread.h>
#include <unistd.h>

typedef struct {
    int kind;                         /* 0 fk, 1 jacobian, 2 collision, 3 fused */
    const Mech *m0; const Sdf *sdf0;
    const int *joint_ids; int n_joint; const double *q; long n_begin, n_end;
    const int *link_ids; int n_req; int jac_link, with_rot, rpy_jac;
    const int *sphere_link_ids; const double *radii; int n_col;
    double truncation_dist; int grad_mode, scratch_mode;
    double *T_out, *J_out, *vals, *grads; int *argmin;
} Job;

static void *job_run(void *arg) {
    Job *b = arg;
    Mech *m = mech_clone(b->m0);
    Sdf *sdf = b->sdf0 ? sdf_clone(b->sdf0) : NULL;
    int n_dof = b->n_joint + (m->with_base ? 3 : 0), rows = b->with_rot ? 6 : 3;
    int nc = b->n_col > 0 ? b->n_col : 1;
    double *Jl = malloc(sizeof(double) * 6 * n_dof);
    double *vl = malloc(sizeof(double) * nc);
    double *gl = malloc(sizeof(double) * nc * n_dof);
    volatile double sink = 0.0;
    for (long n = b->n_begin; n < b->n_end; ++n) {
        or_set_joint_angles(m, b->joint_ids, b->n_joint, b->q + n * n_dof);
        if (b->kind == 0 || b->kind == 3) {
            /* exampel.jl:18-29 shape: get_transform for each requested link */
            for (int k = 0; k < b->n_req; ++k) {
                Tf t = get_transform(m, b->link_ids[k]);
                if (b->T_out) memcpy(b->T_out + ((size_t)n * b->n_req + k) * 16, t.m, sizeof t.m);
                else sink += t.m[12];
            }
        }
        if (b->kind == 1) {
            for (int k = 0; k < b->n_req; ++k)
                or_get_jacobian(m, b->link_ids[k], b->joint_ids, b->n_joint, b->with_rot, b->rpy_jac,
                                b->J_out + ((size_t)n * b->n_req + k) * rows * n_dof);
        }
        if (b->kind == 3 && b->jac_link > 0) {
            double *J = b->J_out ? b->J_out + (size_t)n * rows * n_dof : Jl;
            or_get_jacobian(m, b->jac_link, b->joint_ids, b->n_joint, b->with_rot, b->rpy_jac, J);
            sink += J[0];
        }
        if ((b->kind == 2 || b->kind == 3) && sdf && b->n_col > 0) {
            double *v = b->vals ? b->vals + (size_t)n * b->n_col : vl;
            int *am = b->argmin ? b->argmin + (size_t)n * b->n_col : NULL;
            if (b->kind == 2 && !b->grads) {
                or_compute_coll_dists(m, b->sphere_link_ids, b->radii, b->n_col, sdf, v, am);
            } else {
                double *g = b->grads ? b->grads + (size_t)n * b->n_col * n_dof : gl;
                or_compute_coll_dists_and_grads(m, b->joint_ids, b->n_joint, b->sphere_link_ids, b->radii,
                                                b->n_col, sdf, b->truncation_dist, b->grad_mode,
                                                b->scratch_mode, v, g, am);
            }
            sink += v[0];
        }
    }
    free(Jl); free(vl); free(gl);
    if (sdf) or_sdf_destroy(sdf);
    or_mech_destroy(m);
    return NULL;
}

static void run_jobs(Job *proto, long N, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if ((long)n_threads > N) n_threads = N > 0 ? (int)N : 1;
    Job *jobs = malloc(sizeof(Job) * n_threads);
    pthread_t *th = malloc(sizeof(pthread_t) * n_threads);
    for (int t = 0; t < n_threads; ++t) {
        jobs[t] = *proto;
        jobs[t].n_begin = N * t / n_threads;
        jobs[t].n_end = N * (t + 1) / n_threads;
    }
    for (int t = 1; t < n_threads; ++t) pthread_create(&th[t], NULL, job_run, &jobs[t]);
    job_run(&jobs[0]);
    for (int t = 1; t < n_threads; ++t) pthread_join(th[t], NULL);
    free(jobs); free(th);
}

/* T_out [N][n_req][16] */
OR_EXPORT void or_batch_fk(const Mech *m0, const int *joint_ids, int n_joint, const double *q, long N,
                           const int *link_ids, int n_req, double *T_out, int n_threads) {
    Job j; memset(&j, 0, sizeof j);
    j.kind = 0; j.m0 = m0; j.joint_ids = joint_ids; j.n_joint = n_joint; j.q = q;
    j.link_ids = link_ids; j.n_req = n_req; j.T_out = T_out;
    run_jobs(&j, N, n_threads);
}

/* per config: get_jacobian (zero-initialised) for each requested link.
 * J_out [N][n_req][rows*cols], column-major blocks */
OR_EXPORT void or_batch_jacobian(const Mech *m0, const int *joint_ids, int n_joint, const double *q, long N,
                                 const int *link_ids, int n_req, int with_rot, int rpy_jac, double *J_out,
                                 int n_threads) {
    Job j; memset(&j, 0, sizeof j);
    j.kind = 1; j.m0 = m0; j.joint_ids = joint_ids; j.n_joint = n_joint; j.q = q;
    j.link_ids = link_ids; j.n_req = n_req; j.with_rot = with_rot; j.rpy_jac = rpy_jac; j.J_out = J_out;
    run_jobs(&j, N, n_threads);
}

/* per config: compute_coll_dists_and_grads (or dists only when grads==NULL) */
OR_EXPORT void or_batch_collision(const Mech *m0, const int *joint_ids, int n_joint, const double *q, long N,
                                  const int *sphere_link_ids, const double *radii, int n_col, const Sdf *sdf0,
                                  double truncation_dist, int grad_mode, int scratch_mode,
                                  double *vals, double *grads, int *argmin, int n_threads) {
    Job j; memset(&j, 0, sizeof j);
    j.kind = 2; j.m0 = m0; j.sdf0 = sdf0; j.joint_ids = joint_ids; j.n_joint = n_joint; j.q = q;
    j.sphere_link_ids = sphere_link_ids; j.radii = radii; j.n_col = n_col;
    j.truncation_dist = truncation_dist; j.grad_mode = grad_mode; j.scratch_mode = scratch_mode;
    j.vals = vals; j.grads = grads; j.argmin = argmin;
    run_jobs(&j, N, n_threads);
}

/* The north-star unit of work per configuration, as the reference would do it:
 * set angles; get_transform for every requested link; rows x n_dof Jacobian of
 * jac_link; collision dists + grads.  Output pointers may be NULL (the value is
 * computed and dropped; the work is still done). */
OR_EXPORT void or_batch_fused(const Mech *m0, const int *joint_ids, int n_joint, const double *q, long N,
                              const int *link_ids, int n_req, int jac_link, int with_rot, int rpy_jac,
                              const int *sphere_link_ids, const double *radii, int n_col, const Sdf *sdf0,
                              double truncation_dist, int grad_mode, int scratch_mode,
                              double *T_out, double *J_out, double *vals, double *grads, int n_threads) {
    Job j; memset(&j, 0, sizeof j);
    j.kind = 3; j.m0 = m0; j.sdf0 = sdf0; j.joint_ids = joint_ids; j.n_joint = n_joint; j.q = q;
    j.link_ids = link_ids; j.n_req = n_req; j.jac_link = jac_link; j.with_rot = with_rot; j.rpy_jac = rpy_jac;
    j.sphere_link_ids = sphere_link_ids; j.radii = radii; j.n_col = n_col;
    j.truncation_dist = truncation_dist; j.grad_mode = grad_mode; j.scratch_mode = scratch_mode;
    j.T_out = T_out; j.J_out = J_out; j.vals = vals; j.grads = grads;
    run_jobs(&j, N, n_threads);
}

OR_EXPORT int or_max_threads(void) { long n = sysconf(_SC_NPROCESSORS_ONLN); return n > 0 ? (int)n : 1; }
OR_EXPORT long or_tf_mul_count(const Mech *m) { return m->n_tf_mul; }
