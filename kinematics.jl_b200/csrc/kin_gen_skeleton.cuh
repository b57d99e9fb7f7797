// kin_gen_skeleton.cuh -- the fixed part of the model-specialised kernels (compiled by NVRTC at run time together with
// the text kin_codegen.cpp generates; this file is embedded in the library).  Compilation unit:
//     #include "kin_gen_config.h"      generated: KREAL, KND, KS, KREV, KSTALE, ..., KRADIUS[]
//     #include "kin_device_math.cuh"   the shared arithmetic (same functions as the interpreting kernels)
//     #include "kin_gen_skeleton.cuh"  this file; it includes
//         "kin_gen_phase1.inc"         generated straight-line phase 1 (uses KQ, KST_T, KST_J, KCEN_SET, KJF_OUT)
//         "kin_gen_phase2.inc"         generated: joint frames + one phase2_run<MASK> call per run of spheres
// One thread = one configuration, persistent CTAs, SoA or tiled layout (kin_b200.h).  No table is interpreted: the
// only run-time tables left are the boxes (kin_model_set_boxes moves them without recompiling) and the radii.
#pragma once

namespace kin {

typedef KREAL real;

// ---- phase 2 for the spheres [sb, se), which all have the relevance mask MASK: the code of kin_eval_kernel with
//      the mask, the column count / types and the scratch / gradient / argmin switches as compile-time constants ----
template <typename real_, int ND, unsigned MASK>
__device__ __forceinline__ void phase2_run(const int sb, const int se, const real_ *__restrict__ tb, const int n_box,
                                           const real_ *__restrict__ rad, const real_ *cent0, real_ *stale0, real_ *hand,
                                           const JFrame<real_> (&jfr)[ND > 0 ? ND : 1], const int grad_mode, const real_ trunc,
                                           const real_ voff, real_ *Vp0, real_ *Gp0, int32_t *Ap0, const size_t es) {
    typedef real_ real;
    constexpr int BS = KBS;
    #pragma unroll 1
    for (int s0 = sb; s0 < se; s0 += SPH_GROUP) {
        // ---- 2a: distances of SPH_GROUP spheres; one box-table row feeds all of them ----
        {
            real px[SPH_GROUP], py[SPH_GROUP], pz[SPH_GROUP], kmin[SPH_GROUP];
            int kidx[SPH_GROUP];
            #pragma unroll
            for (int g = 0; g < SPH_GROUP; ++g) {
                const real *cs = cent0 + 3 * min(s0 + g, se - 1) * BS;
                px[g] = cs[0]; py[g] = cs[BS]; pz[g] = cs[2 * BS];
                kmin[g] = CUDART_INF; kidx[g] = 0;
            }
            #pragma unroll 1
            for (int b = 0; b < n_box; ++b) {         // UnionSDF: all boxes, first minimum wins (sdf.jl:108-114)
                BoxRow<real> row;
                load_box(tb + b * BOX_REALS, row);
                real key[SPH_GROUP], qx[SPH_GROUP], qy[SPH_GROUP], qz[SPH_GROUP];
                bool any_inside = false;
                #pragma unroll
                for (int g = 0; g < SPH_GROUP; ++g) {
                    key[g] = box_key_outside(row, px[g], py[g], pz[g], qx[g], qy[g], qz[g]);
                    any_inside |= !(key[g] > real(0));
                }
                if (any_inside) {
                    #pragma unroll
                    for (int g = 0; g < SPH_GROUP; ++g)
                        if (!(key[g] > real(0))) key[g] = box_inside_key(qx[g], qy[g], qz[g]);
                }
                #pragma unroll
                for (int g = 0; g < SPH_GROUP; ++g)
                    if (key[g] < kmin[g]) { kmin[g] = key[g]; kidx[g] = b; }
            }
            #pragma unroll
            for (int g = 0; g < SPH_GROUP; ++g) {
                hand[g * BS] = key_to_dist(kmin[g]);
                reinterpret_cast<int *>(&hand[(SPH_GROUP + g) * BS])[0] = kidx[g];
            }
        }
        // ---- 2b: per sphere, IN sphere order (the shared scratch of collision.jl:76,90 makes the order observable) ----
        #pragma unroll 1
        for (int g = 0; g < SPH_GROUP && s0 + g < se; ++g) {
            const int s = s0 + g;
            const real dmin = hand[g * BS];
            const int kmin = reinterpret_cast<const int *>(&hand[(SPH_GROUP + g) * BS])[0];
            const real dist0 = dmin - rad[s];
            const bool truncated = dist0 > trunc;
            __stcs(Vp0 + (size_t)s * es, (truncated ? trunc : dist0) - voff);
            if (KARGMIN) __stcs(Ap0 + (size_t)s * es, kmin + 1);
            if (!KGRADS) continue;
            real *Gp = Gp0 + (size_t)s * ND * es;
            if (truncated) {            // collision.jl:84-86
                #pragma unroll
                for (int j = 0; j < ND; ++j) __stcs(&Gp[(size_t)j * es], real(0));
                continue;
            }
            const real *cs = cent0 + 3 * s * BS;
            const real px = cs[0], py = cs[BS], pz = cs[2 * BS];
            real grad[3];
            {
                BoxRow<real> row;
                load_box(tb + kmin * BOX_REALS, row);
                box_gradient(row, grad_mode, px, py, pz, dmin, grad);
            }
            #pragma unroll
            for (int j = 0; j < ND; ++j) {
                real *st = stale0 + 3 * j * BS;
                if ((MASK >> j) & 1u) {       // joint_jacobian!, algorithm.jl:65-81
                    real cx, cy, cz;
                    jac_col(jfr[j], ((KREV >> j) & 1u) != 0, px, py, pz, cx, cy, cz);
                    if (KSTALE) { st[0] = cx; st[BS] = cy; st[2 * BS] = cz; }
                    __stcs(&Gp[(size_t)j * es], fma_(grad[0], cx, fma_(grad[1], cy, grad[2] * cz)));   // transpose(grad) * jac
                } else if (KSTALE) {          // column left over from an earlier sphere (collision.jl:76,90)
                    const real cx = st[0], cy = st[BS], cz = st[2 * BS];
                    __stcs(&Gp[(size_t)j * es], fma_(grad[0], cx, fma_(grad[1], cy, grad[2] * cz)));
                } else {
                    __stcs(&Gp[(size_t)j * es], real(0));      // a zero column (clean scratch): transpose(grad) * 0
                }
            }
        }
    }
}

}  // namespace kin

// =====================================================================================================================
// Monolithic kernel: phase 1 and phase 2 in the same thread.
// Dynamic shared memory: [box table n_box * BOX_REALS][radii KS] then the per-thread scratch [slot][thread]:
//   3 KS sphere-centre coordinates, 3 KND stale-Jacobian columns (KSTALE), 2 SPH_GROUP hand-over slots.
// =====================================================================================================================
#if !KWS
extern "C" __global__ void __launch_bounds__(KBS, KMINB) kin_gen_kernel(const __grid_constant__ kin::GenArgs A) {
    using namespace kin;
    constexpr int BS = KBS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
#if KCOLL
    real *tb = reinterpret_cast<real *>(smem_raw);
    const int n_box = A.n_box;
    const int tab_reals = (n_box * BOX_REALS + KS + 1) & ~1;
    real *rad = tb + n_box * BOX_REALS;
    real *scr = tb + tab_reals + tid;                         // scr[slot * BS]
    real *cent0 = scr;
    real *stale0 = scr + 3 * KS * BS;
    real *hand = stale0 + (KSTALE ? 3 * KND : 0) * BS;
    {
        const real *src = reinterpret_cast<const real *>(A.boxes);
        for (int i = tid; i < n_box * BOX_REALS; i += BS) tb[i] = src[i];
        for (int i = tid; i < KS; i += BS) rad[i] = KRADIUS[i];
    }
    __syncthreads();
    const real trunc = (real)A.truncation_dist, voff = (real)A.vals_offset;
#endif
    const size_t es = KTILED ? size_t(32) : (size_t)A.ld;
    const long long n_tiles = (A.n + BS - 1) / BS;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // threads past the end of the batch redo the last configuration (identical values, benign duplicate stores)
        const long long n = min(tile * BS + tid, (long long)A.n - 1);
        #define KREC_BASE(rec) (KTILED ? (n >> 5) * ((long long)(rec) * 32) + (n & 31) : n)
        const real *qn = reinterpret_cast<const real *>(A.q) + KREC_BASE(KND);
        #define KQ(c) qn[(size_t)(c) * es]
#if KWANT_T
        real *Tn = reinterpret_cast<real *>(A.T_out) + KREC_BASE(12 * KNFK);
        #define KST_T(k, v) __stcs(Tn + (size_t)(k) * es, (v))
#else
        #define KST_T(k, v)
#endif
#if KWANT_J
        real *Jn = reinterpret_cast<real *>(A.J_out) + KREC_BASE(KROWS * KND * KNJAC);
        #define KST_J(k, v) __stcs(Jn + (size_t)(k) * es, (v))
#else
        #define KST_J(k, v)
#endif
#if KCOLL
        #define KCEN_SET(s, i, v) cent0[(3 * (s) + (i)) * BS] = (v)
#else
        #define KCEN_SET(s, i, v)
#endif
        #define KJF_OUT(j, i, v)
        {
#include "kin_gen_phase1.inc"
#if KCOLL
            if (KSTALE && KGRADS) {
                #pragma unroll
                for (int i = 0; i < 3 * KND; ++i) stale0[i * BS] = real(0);   // jac = zeros(3, n_dof), collision.jl:76
            }
            real *Vp0 = reinterpret_cast<real *>(A.vals_out) + KREC_BASE(KS);
            real *Gp0 = reinterpret_cast<real *>(A.grads_out) + KREC_BASE((long long)KND * KS);
            int32_t *Ap0 = KARGMIN ? A.argmin_out + KREC_BASE(KS) : nullptr;
            #define KP2ARGS tb, n_box, rad, cent0, stale0, hand, jfr, A.grad_mode, trunc, voff, Vp0, Gp0, Ap0, es
#include "kin_gen_phase2.inc"
#endif
        }
    }
}
#endif
