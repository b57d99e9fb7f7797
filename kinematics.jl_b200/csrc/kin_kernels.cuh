// kin_kernels.cuh -- the sm_100a kernel of libkin_b200: one pass per configuration over the
// compiled kinematic program (kin_program.h) doing
//   phase 1  forward kinematics of the dynamic nodes (algorithm.jl:1-37, mechanism.jl:90-103),
//            world joint frames (algorithm.jl:42-54), requested link transforms, requested
//            Jacobians (algorithm.jl:56-114) and collision-sphere centres (collision.jl:39-49);
//   phase 2  per sphere, in sscc.sphere_links order: union-of-boxes SDF with first-minimum argmin
//            (sdf.jl:67-74,108-114), truncation, forward-difference or analytic gradient
//            (sdf.jl:34-41,116-119) chain-ruled through the sphere's 3 x n_dof Jacobian
//            (collision.jl:67-94), with the reference's shared-scratch behaviour reproducible.
//
// Mapping: one thread = one configuration.  Everything indexed by a run-time table index lives in
// shared memory as [slot][thread] (bank-conflict free, no local memory); the running transform
// stays in registers.  Program tables are staged once per CTA into shared memory and read with
// warp-uniform (broadcast) loads.  CTAs are persistent over tiles of blockDim configurations.
// SoA outputs are written as full 256-B (f64) / 128-B (f32) warp rows.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "kin_program.h"

namespace kin {

struct KernelArgs {
    ProgHeader h;
    const int32_t *tab_i;       // device: int section
    const void *tab_r;          // device: real section (element type = real)
    const void *q;
    void *T_out, *J_out, *vals_out, *grads_out;
    int32_t *argmin_out;
    long long n, ld;
    int with_rot, rpy_jac, keep_irrelevant;
    int grad_mode, scratch_ref;  // scratch_ref = 1: reproduce the reference's shared Jacobian scratch
    double truncation_dist, vals_offset;
};

template <typename real> struct Tf { real r[9]; real p[3]; };   // rotation row-major

__device__ __forceinline__ void sincos_(double x, double *s, double *c) { sincos(x, s, c); }
__device__ __forceinline__ void sincos_(float x, float *s, float *c) { sincosf(x, s, c); }
__device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float fma_(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
__device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
__device__ __forceinline__ double atan2_(double y, double x) { return atan2(y, x); }
__device__ __forceinline__ float atan2_(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ double abs_(double x) { return fabs(x); }
__device__ __forceinline__ float abs_(float x) { return fabsf(x); }
__device__ __forceinline__ double max_(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float max_(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double min_(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ float min_(float a, float b) { return fminf(a, b); }

// out = a * b  (a: running transform, b: constant from the table at `c`, R row-major then t)
template <typename real>
__device__ __forceinline__ void tf_mul_const(const Tf<real> &a, const real *__restrict__ c, bool r_identity, Tf<real> &o) {
    #pragma unroll
    for (int i = 0; i < 3; ++i)
        o.p[i] = fma_(a.r[i * 3 + 0], c[9], fma_(a.r[i * 3 + 1], c[10], fma_(a.r[i * 3 + 2], c[11], a.p[i])));
    if (r_identity) {
        #pragma unroll
        for (int i = 0; i < 9; ++i) o.r[i] = a.r[i];
    } else {
        #pragma unroll
        for (int i = 0; i < 3; ++i)
            #pragma unroll
            for (int j = 0; j < 3; ++j)
                o.r[i * 3 + j] = fma_(a.r[i * 3 + 0], c[j], fma_(a.r[i * 3 + 1], c[3 + j], a.r[i * 3 + 2] * c[6 + j]));
    }
}

// Output addressing.  SoA: component-major with leading dimension ld; AoS: record-major.
template <bool AOS>
struct OutIdx {
    long long n, ld; int rec;
    __device__ __forceinline__ long long operator()(int comp) const {
        return AOS ? n * rec + comp : (long long)comp * ld + n;
    }
};

// BoxSDF call (sdf.jl:67-74) for the box whose table row starts at `b`.
template <typename real>
__device__ __forceinline__ real box_sdf(const real *__restrict__ b, real px, real py, real pz) {
    real lx = fma_(b[0], px, fma_(b[1], py, fma_(b[2], pz, b[9])));
    real ly = fma_(b[3], px, fma_(b[4], py, fma_(b[5], pz, b[10])));
    real lz = fma_(b[6], px, fma_(b[7], py, fma_(b[8], pz, b[11])));
    real qx = abs_(lx) - b[12], qy = abs_(ly) - b[13], qz = abs_(lz) - b[14];
    real mx = max_(qx, real(0)), my = max_(qy, real(0)), mz = max_(qz, real(0));
    real nrm = sqrt_(fma_(mx, mx, fma_(my, my, mz * mz)));
    return nrm + min_(max_(max_(qx, qy), qz), real(0));
}

// closed-form gradient of the same box, world frame (extension; KIN_GRAD_ANALYTIC)
template <typename real>
__device__ __forceinline__ void box_grad_analytic(const real *__restrict__ b, real px, real py, real pz, real g[3]) {
    real l[3], q[3], m[3], gl[3] = {0, 0, 0};
    l[0] = fma_(b[0], px, fma_(b[1], py, fma_(b[2], pz, b[9])));
    l[1] = fma_(b[3], px, fma_(b[4], py, fma_(b[5], pz, b[10])));
    l[2] = fma_(b[6], px, fma_(b[7], py, fma_(b[8], pz, b[11])));
    #pragma unroll
    for (int i = 0; i < 3; ++i) { q[i] = abs_(l[i]) - b[12 + i]; m[i] = max_(q[i], real(0)); }
    real nrm = sqrt_(fma_(m[0], m[0], fma_(m[1], m[1], m[2] * m[2])));
    if (nrm > real(0)) {
        #pragma unroll
        for (int i = 0; i < 3; ++i) gl[i] = (m[i] / nrm) * (l[i] < real(0) ? real(-1) : real(1));
    } else {
        int k = 0;
        if (q[1] > q[k]) k = 1;
        if (q[2] > q[k]) k = 2;
        #pragma unroll
        for (int i = 0; i < 3; ++i) if (i == k) gl[i] = l[i] < real(0) ? real(-1) : real(1);
    }
    // world = R * g_local, and the table holds inv_R = R' row-major => R[r][c] = b[c*3 + r]
    #pragma unroll
    for (int r = 0; r < 3; ++r) g[r] = fma_(b[0 + r], gl[0], fma_(b[3 + r], gl[1], b[6 + r] * gl[2]));
}

template <typename real, bool AOS>
__global__ void __launch_bounds__(128)
kin_eval_kernel(const __grid_constant__ KernelArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const ProgHeader &h = A.h;
    int32_t *ti = reinterpret_cast<int32_t *>(smem_raw);
    real *tr = reinterpret_cast<real *>(smem_raw + sizeof(int32_t) * (size_t)h.n_int);
    real *scr = tr + h.n_real;                         // [slot][thread]
    const int tid = threadIdx.x, bs = blockDim.x;

    // ---- stage the program tables once per CTA ----
    {
        const int4 *src = reinterpret_cast<const int4 *>(A.tab_i);
        int4 *dst = reinterpret_cast<int4 *>(ti);
        for (int i = tid; i < h.n_int / 4; i += bs) dst[i] = src[i];
        const real *rs = reinterpret_cast<const real *>(A.tab_r);
        for (int i = tid; i < h.n_real; i += bs) tr[i] = rs[i];
    }
    __syncthreads();

    const real *q = reinterpret_cast<const real *>(A.q);
    real *T_out = reinterpret_cast<real *>(A.T_out);
    real *J_out = reinterpret_cast<real *>(A.J_out);
    real *V_out = reinterpret_cast<real *>(A.vals_out);
    real *G_out = reinterpret_cast<real *>(A.grads_out);
    const int D = h.n_joints, ND = h.n_dof;
    const int rows = A.with_rot ? 6 : 3;
    #define SCR(slot) scr[(slot) * bs + tid]

    const long long n_tiles = (A.n + bs - 1) / bs;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long n = tile * bs + tid;
        if (n >= A.n) continue;      // no block-level sync below this point

        // ---- configuration -> scratch (all loads in flight together) ----
        for (int c = 0; c < ND; ++c)
            SCR(h.so_q + c) = AOS ? q[n * ND + c] : q[(long long)c * A.ld + n];

        real bx = 0, by = 0;          // planar base position (algorithm.jl:99)
        Tf<real> T;                   // running world transform of the current node

        // =========================== phase 1 ===========================
        for (int node = 0; node < h.n_nodes; ++node) {
            const int32_t *ni = ti + h.io_node + node * NODE_INTS;
            const real *nr = tr + h.ro_node + node * NODE_REALS;
            const int jtype = ni[1];
            if (jtype == NODE_ROOT) {
                #pragma unroll
                for (int i = 0; i < 9; ++i) T.r[i] = (i % 4 == 0) ? real(1) : real(0);
                T.p[0] = T.p[1] = T.p[2] = real(0);
                if (h.with_base) {   // base_pose_to_transform, transform.jl:33-37
                    bx = SCR(h.so_q + D); by = SCR(h.so_q + D + 1);
                    real s, c;
                    sincos_(SCR(h.so_q + D + 2), &s, &c);
                    T.r[0] = c; T.r[1] = -s; T.r[3] = s; T.r[4] = c;
                    T.p[0] = bx; T.p[1] = by;
                }
            } else {
                const int psrc = ni[0], flags = ni[2], qcol = ni[3];
                if (psrc >= 0) {
                    #pragma unroll
                    for (int i = 0; i < 9; ++i) T.r[i] = SCR(h.so_save + 12 * psrc + i);
                    #pragma unroll
                    for (int i = 0; i < 3; ++i) T.p[i] = SCR(h.so_save + 12 * psrc + 9 + i);
                }
                // A = T_parent * joint.pose : the joint frame (algorithm.jl:47-48)
                Tf<real> Aj;
                tf_mul_const(T, nr, flags & NF_OFF_R_IDENTITY, Aj);
                const int code = (flags >> NF_AXIS_SHIFT) & NF_AXIS_MASK;
                real ax, ay, az;          // world joint axis (algorithm.jl:50)
                real sgn = code >= 4 ? real(-1) : real(1);
                switch (code) {
                    case 1: case 4: ax = sgn * Aj.r[0]; ay = sgn * Aj.r[3]; az = sgn * Aj.r[6]; break;
                    case 2: case 5: ax = sgn * Aj.r[1]; ay = sgn * Aj.r[4]; az = sgn * Aj.r[7]; break;
                    case 3: case 6: ax = sgn * Aj.r[2]; ay = sgn * Aj.r[5]; az = sgn * Aj.r[8]; break;
                    default:
                        ax = fma_(Aj.r[0], nr[12], fma_(Aj.r[1], nr[13], Aj.r[2] * nr[14]));
                        ay = fma_(Aj.r[3], nr[12], fma_(Aj.r[4], nr[13], Aj.r[5] * nr[14]));
                        az = fma_(Aj.r[6], nr[12], fma_(Aj.r[7], nr[13], Aj.r[8] * nr[14]));
                }
                const int jf = h.so_jf + 6 * qcol;
                SCR(jf + 0) = Aj.p[0]; SCR(jf + 1) = Aj.p[1]; SCR(jf + 2) = Aj.p[2];
                SCR(jf + 3) = ax; SCR(jf + 4) = ay; SCR(jf + 5) = az;
                const real qa = SCR(h.so_q + qcol);
                T = Aj;
                if (jtype == 2) {          // prismatic: pose * Trans(axis * a), mechanism.jl:100-103
                    T.p[0] = fma_(ax, qa, Aj.p[0]); T.p[1] = fma_(ay, qa, Aj.p[1]); T.p[2] = fma_(az, qa, Aj.p[2]);
                } else {                   // revolute: pose * R(axis, a), mechanism.jl:94-98
                    real s, c;
                    sincos_(code >= 4 ? -qa : qa, &s, &c);
                    // rotation about a coordinate axis mixes the two other columns
                    #define GIVENS(u, v)                                                      \
                        _Pragma("unroll") for (int i = 0; i < 3; ++i) {                       \
                            real cu = Aj.r[i * 3 + u], cv = Aj.r[i * 3 + v];                  \
                            T.r[i * 3 + u] = fma_(c, cu, s * cv);                             \
                            T.r[i * 3 + v] = fma_(c, cv, -(s * cu));                          \
                        }
                    switch (code) {
                        case 1: case 4: GIVENS(1, 2) break;
                        case 2: case 5: GIVENS(2, 0) break;
                        case 3: case 6: GIVENS(0, 1) break;
                        default: {
                            // Rodrigues form of the reference's quaternion rotation:
                            // Rot = c I + s [a]x + (1-c) a a'
                            const real x = nr[12], y = nr[13], z = nr[14], t = real(1) - c;
                            real m[9];
                            m[0] = fma_(t * x, x, c);      m[1] = fma_(t * x, y, -(s * z)); m[2] = fma_(t * x, z, s * y);
                            m[3] = fma_(t * x, y, s * z);  m[4] = fma_(t * y, y, c);        m[5] = fma_(t * y, z, -(s * x));
                            m[6] = fma_(t * x, z, -(s * y)); m[7] = fma_(t * y, z, s * x);  m[8] = fma_(t * z, z, c);
                            #pragma unroll
                            for (int i = 0; i < 3; ++i)
                                #pragma unroll
                                for (int j = 0; j < 3; ++j)
                                    T.r[i * 3 + j] = fma_(Aj.r[i * 3 + 0], m[j], fma_(Aj.r[i * 3 + 1], m[3 + j], Aj.r[i * 3 + 2] * m[6 + j]));
                        }
                    }
                    #undef GIVENS
                }
            }
            if (ni[4] >= 0) {
                #pragma unroll
                for (int i = 0; i < 9; ++i) SCR(h.so_save + 12 * ni[4] + i) = T.r[i];
                #pragma unroll
                for (int i = 0; i < 3; ++i) SCR(h.so_save + 12 * ni[4] + 9 + i) = T.p[i];
            }

            // ---- requested links hanging from this node ----
            for (int a = ni[5]; a < ni[6]; ++a) {
                const int32_t *ai = ti + h.io_att + a * ATT_INTS;
                const real *ar = tr + h.ro_att + a * ATT_REALS;
                Tf<real> Tl;
                tf_mul_const(T, ar, ai[1] & AF_R_IDENTITY, Tl);
                if (ai[0] >= 0 && T_out) {       // get_transform, as 3x4 column-major
                    OutIdx<AOS> o{n, A.ld, 12 * h.n_fk};
                    const int base = 12 * ai[0];
                    #pragma unroll
                    for (int c = 0; c < 3; ++c)
                        #pragma unroll
                        for (int r = 0; r < 3; ++r) T_out[o(base + c * 3 + r)] = Tl.r[r * 3 + c];
                    #pragma unroll
                    for (int r = 0; r < 3; ++r) T_out[o(base + 9 + r)] = Tl.p[r];
                }
                if (ai[2] >= 0 && J_out) {       // get_jacobian, algorithm.jl:83-114
                    OutIdx<AOS> o{n, A.ld, rows * ND * h.n_jac};
                    const int jbase = ai[2] * rows * ND;
                    const unsigned mask = (unsigned)ai[3];
                    real k_rx = 0, k_ry = 0, k_px = 0, k_py = 0, k_yx = 0, k_yy = 0;
                    if (A.with_rot && A.rpy_jac) {
                        // rpy(T) (transform.jl:45-48, RotZYX) then the Euler-rate map (algorithm.jl:56-63)
                        const real yaw = atan2_(Tl.r[3], Tl.r[0]);
                        const real pitch = atan2_(-Tl.r[6], sqrt_(fma_(Tl.r[7], Tl.r[7], Tl.r[8] * Tl.r[8])));
                        real s2, c2, s3, c3;
                        sincos_(-pitch, &s2, &c2);
                        sincos_(-yaw, &s3, &c3);
                        k_rx = c3 / c2; k_ry = s3 / c2;
                        k_px = s3; k_py = c3;
                        k_yx = -c3 * s2 / c2; k_yy = s3 * s2 / c2;
                    }
                    for (int j = 0; j < D; ++j) {
                        const int cb = jbase + j * rows;
                        if ((mask >> j) & 1u) {
                            const int jf = h.so_jf + 6 * j;
                            const real ax = SCR(jf + 3), ay = SCR(jf + 4), az = SCR(jf + 5);
                            if (ti[h.io_col_type + j] == 1) {
                                const real dx = Tl.p[0] - SCR(jf + 0), dy = Tl.p[1] - SCR(jf + 1), dz = Tl.p[2] - SCR(jf + 2);
                                J_out[o(cb + 0)] = fma_(ay, dz, -(az * dy));
                                J_out[o(cb + 1)] = fma_(az, dx, -(ax * dz));
                                J_out[o(cb + 2)] = fma_(ax, dy, -(ay * dx));
                                if (A.with_rot) {
                                    if (A.rpy_jac) {
                                        J_out[o(cb + 3)] = k_rx * ax - k_ry * ay;
                                        J_out[o(cb + 4)] = fma_(k_px, ax, k_py * ay);
                                        J_out[o(cb + 5)] = fma_(k_yx, ax, k_yy * ay) + az;
                                    } else {
                                        J_out[o(cb + 3)] = ax; J_out[o(cb + 4)] = ay; J_out[o(cb + 5)] = az;
                                    }
                                }
                            } else {   // prismatic: rows 4:6 untouched by the reference (algorithm.jl:78-81)
                                J_out[o(cb + 0)] = ax; J_out[o(cb + 1)] = ay; J_out[o(cb + 2)] = az;
                                if (A.with_rot && !A.keep_irrelevant) {
                                    J_out[o(cb + 3)] = real(0); J_out[o(cb + 4)] = real(0); J_out[o(cb + 5)] = real(0);
                                }
                            }
                        } else if (!A.keep_irrelevant) {
                            for (int r = 0; r < rows; ++r) J_out[o(cb + r)] = real(0);
                        }
                    }
                    if (h.with_base) {   // algorithm.jl:98-105
                        const real x = Tl.p[0] - bx, y = Tl.p[1] - by;
                        const int cb = jbase + D * rows;
                        J_out[o(cb + 0)] = real(1); J_out[o(cb + 1)] = real(0); J_out[o(cb + 2)] = real(0);
                        J_out[o(cb + rows + 0)] = real(0); J_out[o(cb + rows + 1)] = real(1); J_out[o(cb + rows + 2)] = real(0);
                        J_out[o(cb + 2 * rows + 0)] = -y; J_out[o(cb + 2 * rows + 1)] = x; J_out[o(cb + 2 * rows + 2)] = real(0);
                        if (A.with_rot) {
                            for (int c = 0; c < 3; ++c)
                                for (int r = 3; r < 6; ++r) J_out[o(cb + c * rows + r)] = (c == 2 && r == 5) ? real(1) : real(0);
                        }
                    }
                }
            }

            // ---- collision-sphere centres on this node (collision.jl:54 / :80) ----
            for (int k = ni[7]; k < ni[8]; ++k) {
                const int s = ti[h.io_sph_order + k];
                const real *sr = tr + h.ro_sph + s * SPH_REALS;
                #pragma unroll
                for (int i = 0; i < 3; ++i)
                    SCR(h.so_cent + 3 * s + i) =
                        fma_(T.r[i * 3 + 0], sr[0], fma_(T.r[i * 3 + 1], sr[1], fma_(T.r[i * 3 + 2], sr[2], T.p[i])));
            }
        }

        // =========================== phase 2 ===========================
        if (h.n_sph > 0 && V_out) {
            const bool want_grads = G_out != nullptr;
            const bool stale = want_grads && A.scratch_ref;
            const real trunc = (real)A.truncation_dist;
            if (stale)
                for (int i = 0; i < 3 * D; ++i) SCR(h.so_stale + i) = real(0);   // jac = zeros(3, n_dof), collision.jl:76
            OutIdx<AOS> ov{n, A.ld, h.n_sph};
            OutIdx<AOS> og{n, A.ld, ND * h.n_sph};
            for (int s = 0; s < h.n_sph; ++s) {
                const real px = SCR(h.so_cent + 3 * s), py = SCR(h.so_cent + 3 * s + 1), pz = SCR(h.so_cent + 3 * s + 2);
                // UnionSDF: all boxes, first minimum wins (sdf.jl:108-114)
                real dmin = CUDART_INF;
                int kmin = 0;
                for (int b = 0; b < h.n_box; ++b) {
                    const real d = box_sdf(tr + h.ro_box + b * BOX_REALS, px, py, pz);
                    if (d < dmin) { dmin = d; kmin = b; }
                }
                const real dist0 = dmin - tr[h.ro_sph + s * SPH_REALS + 3];
                if (A.argmin_out) A.argmin_out[ov(s)] = kmin + 1;
                const bool truncated = dist0 > trunc;
                V_out[ov(s)] = (truncated ? trunc : dist0) - (real)A.vals_offset;
                if (!want_grads) continue;
                if (truncated) {            // collision.jl:84-86
                    for (int j = 0; j < ND; ++j) G_out[og(s * ND + j)] = real(0);
                    continue;
                }
                real g[3];
                const real *bk = tr + h.ro_box + kmin * BOX_REALS;
                if (A.grad_mode == 0) {     // forward difference on the argmin box, sdf.jl:34-41
                    const real eps = real(1e-7);
                    g[0] = (box_sdf(bk, px + eps, py, pz) - dmin) / eps;
                    g[1] = (box_sdf(bk, px, py + eps, pz) - dmin) / eps;
                    g[2] = (box_sdf(bk, px, py, pz + eps) - dmin) / eps;
                } else {
                    box_grad_analytic(bk, px, py, pz, g);
                }
                const unsigned mask = (unsigned)ti[h.io_sph_mask + s];
                for (int j = 0; j < D; ++j) {
                    real cx, cy, cz;
                    const bool rel = (mask >> j) & 1u;
                    if (rel) {              // joint_jacobian!, algorithm.jl:65-81
                        const int jf = h.so_jf + 6 * j;
                        const real ax = SCR(jf + 3), ay = SCR(jf + 4), az = SCR(jf + 5);
                        if (ti[h.io_col_type + j] == 1) {
                            const real dx = px - SCR(jf + 0), dy = py - SCR(jf + 1), dz = pz - SCR(jf + 2);
                            cx = fma_(ay, dz, -(az * dy)); cy = fma_(az, dx, -(ax * dz)); cz = fma_(ax, dy, -(ay * dx));
                        } else { cx = ax; cy = ay; cz = az; }
                        if (stale) { SCR(h.so_stale + 3 * j) = cx; SCR(h.so_stale + 3 * j + 1) = cy; SCR(h.so_stale + 3 * j + 2) = cz; }
                    } else if (stale) {     // column left over from an earlier sphere (collision.jl:76,90)
                        cx = SCR(h.so_stale + 3 * j); cy = SCR(h.so_stale + 3 * j + 1); cz = SCR(h.so_stale + 3 * j + 2);
                    } else { cx = cy = cz = real(0); }
                    G_out[og(s * ND + j)] = fma_(g[0], cx, fma_(g[1], cy, g[2] * cz));   // transpose(grad) * jac
                }
                if (h.with_base) {          // base columns are rewritten for every sphere (algorithm.jl:98-101)
                    const real x = px - bx, y = py - by;
                    G_out[og(s * ND + D)] = g[0];
                    G_out[og(s * ND + D + 1)] = g[1];
                    G_out[og(s * ND + D + 2)] = fma_(g[1], x, -(g[0] * y));
                }
            }
        }
    }
    #undef SCR
}

// sdf(p) and gradient!(sdf, p, out) for a batch of points (sdf.jl:34-41, 67-74, 108-119).
template <typename real, bool AOS>
__global__ void __launch_bounds__(256)
sdf_points_kernel(const real *__restrict__ boxes, int n_box, const real *__restrict__ pts, long long n_pts,
                  int grad_mode, real *__restrict__ vals, real *__restrict__ grads, int32_t *__restrict__ argmin) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    real *tb = reinterpret_cast<real *>(smem_raw);
    for (int i = threadIdx.x; i < n_box * BOX_REALS; i += blockDim.x) tb[i] = boxes[i];
    __syncthreads();
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < n_pts; n += (long long)gridDim.x * blockDim.x) {
        const real px = AOS ? pts[3 * n] : pts[n], py = AOS ? pts[3 * n + 1] : pts[n_pts + n],
                   pz = AOS ? pts[3 * n + 2] : pts[2 * n_pts + n];
        real dmin = CUDART_INF;
        int kmin = 0;
        for (int b = 0; b < n_box; ++b) {
            const real d = box_sdf(tb + b * BOX_REALS, px, py, pz);
            if (d < dmin) { dmin = d; kmin = b; }
        }
        vals[n] = dmin;
        if (argmin) argmin[n] = kmin + 1;
        if (grads) {
            real g[3];
            const real *bk = tb + kmin * BOX_REALS;
            if (grad_mode == 0) {
                const real eps = real(1e-7);
                g[0] = (box_sdf(bk, px + eps, py, pz) - dmin) / eps;
                g[1] = (box_sdf(bk, px, py + eps, pz) - dmin) / eps;
                g[2] = (box_sdf(bk, px, py, pz + eps) - dmin) / eps;
            } else {
                box_grad_analytic(bk, px, py, pz, g);
            }
            #pragma unroll
            for (int i = 0; i < 3; ++i) grads[AOS ? 3 * n + i : (long long)i * n_pts + n] = g[i];
        }
    }
}

}  // namespace kin
