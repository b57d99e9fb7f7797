// kin_kernels_ws.cuh -- warp-specialised form of the fused kernel (FK + Jacobians + collision cost / gradient).
//
// Why: kin_eval_kernel keeps ALL of a configuration's state in one thread (96 shared-memory slots + ~110 live
// doubles), which caps residency at 8 warps per SM, and the kernel is bound by instruction-issue latency (IPC
// 0.41 per sub-partition with 2 warps each).  The two halves of the work need very different register budgets:
// the chain walk (phase 1: algorithm.jl:1-54,83-114) ~80 registers, the sphere / box / chain-rule work (phase 2:
// collision.jl:67-94, sdf.jl:34-41,108-119) ~195 because it holds the eight joint frames in registers.  Here one
// CTA of 384 threads per SM is split into three warpgroups with `setmaxnreg`:
//     warpgroup 0     PRODUCER   88 registers: walks the chain of one configuration per thread, writes link
//                     transforms and Jacobians to global memory and hands (joint frames, sphere centres) over
//                     through a ring of WS_STAGES tiles;
//     warpgroup 1, 2  CONSUMERS  208 registers each (88 + 2 x 208 = 3 x 168): copy the frames into registers and
//                     the centres into their private shared memory (cp.async), release the ring stage at once,
//                     and do phase 2 exactly as kin_eval_kernel does (same helpers, same order => bitwise-identical
//                     results: test_warp_specialised_kernel_is_bitwise_identical).
// 12 resident warps (3 per sub-partition: one producer + two consumers) instead of 8; the producer's ~5.7k
// instructions per configuration and the consumers' ~12k balance at one producer per two consumers (the producer
// waits for a free stage ~8 % of its time, the consumers for a full one < 1 %).
//
// The ring lives in GLOBAL memory: one region per CTA, WS_STAGES x 96 KB, reused every few microseconds and
// therefore L2-resident (measured per 2^24 configurations: 4 stages 17.1 ms, 3: 16.7, 2: 16.3 -- the smaller
// footprint stays in L2, DRAM traffic is back to the algorithmic 4 KB per configuration).  Shared memory cannot
// hold both the consumers' per-configuration state (80 slots x 256 threads = 160 KB) and a hand-over buffer deep
// enough to keep them busy: a first version with a two-stage ring in shared memory left each consumer idle while
// its only stage was being refilled and ended 3 % slower than kin_eval_kernel.  Producer and consumers run on the
// same SM; mbarrier full[stage] (128 producer arrivals) / empty[stage] (128 consumer arrivals) order the
// hand-over at CTA scope, and every thread only ever touches its own column of a stage.
//
// Eligible: FP64, SoA or tiled layout, collision requested, <= 8 control joints (+ planar base), a chain without
// save slots, and the
// consumers' state fits in shared memory (S <= ~16 spheres); kin_b200.cu falls back to kin_eval_kernel otherwise.
#pragma once
#include "kin_kernels.cuh"

namespace kin {

constexpr int WS_TILE = 128;                     // configurations per tile = threads per warpgroup
constexpr int WS_THREADS = 3 * WS_TILE;
#ifndef KIN_WS_STAGES
#define KIN_WS_STAGES 2
#endif
#ifndef KIN_WS_SLEEP_NS
#define KIN_WS_SLEEP_NS 200
#endif
constexpr int WS_STAGES = KIN_WS_STAGES;         // ring depth (tiles)
// Columns: up to JF_REGS control joints, whose frames the consumers hold in registers, plus the three columns of
// the planar base (prismatic x, prismatic y, revolute z: algorithm.jl:98-105), whose frames are the constants
// e_x, e_y and (e_z through (x, y, 0)) -- the consumers keep just (x, y).
constexpr int WS_MAX_COLS = JF_REGS + 3;
__host__ __device__ constexpr int ws_cols(bool base) { return base ? WS_MAX_COLS : JF_REGS; }   // column capacity of the layouts
__host__ __device__ inline bool ws_has_base(const ProgHeader &h) { return h.n_dof > h.n_joints; }
// ring slots [0, 6 cols): joint frames, then 3 S sphere-centre coordinates
// PRE (collision-only calls: no link transforms / Jacobians requested): the producer, which then has little to do,
// also runs the box search (phase 2a) of every sphere and hands (distance, argmin) over; the consumers keep phase 2b.
// ring slots [0, 6 cols): joint frames, then 3 S sphere-centre coordinates [, then S distances and S argmins]
__host__ __device__ inline int ws_ring_slots(int n_sph, bool base, bool pre) { return 6 * ws_cols(base) + 3 * n_sph + (pre ? 2 * n_sph : 0); }
// consumer-private shared slots (doubles): the shared Jacobian scratch of collision.jl:76 (3 x cols); (dmin, argmin)
// of the current sphere group [PRE: S distances + S int32 argmins from the producer]; the 3 S centre coordinates
__host__ __device__ inline int ws_hand_slots(int n_sph, bool pre) { return pre ? n_sph + (n_sph + 1) / 2 : 2 * SPH_GROUP; }
__host__ __device__ inline int ws_priv_slots(int n_sph, bool base, bool pre) { return 3 * n_sph + 3 * ws_cols(base) + ws_hand_slots(n_sph, pre); }
// bytes of global scratch one launch needs (n_cta regions)
__host__ __device__ inline size_t ws_ring_bytes(const ProgHeader &h, int n_cta, bool pre) {
    return sizeof(double) * (size_t)n_cta * WS_STAGES * ws_ring_slots(h.n_sph, ws_has_base(h), pre) * WS_TILE;
}
// bytes of dynamic shared memory: tables, 2 x WS_STAGES mbarriers, producer q double buffer, consumer state
__host__ __device__ inline size_t ws_smem_bytes(const ProgHeader &h, bool pre) {
    size_t b = sizeof(int32_t) * (size_t)h.n_int + sizeof(double) * (size_t)h.n_real;
    b = (b + 15) & ~size_t(15);
    b += 64;                                                            // mbarriers
    b += sizeof(double) * 2 * ws_cols(ws_has_base(h)) * WS_TILE;        // q double buffer
    b += sizeof(double) * 2 * (size_t)ws_priv_slots(h.n_sph, ws_has_base(h), pre) * WS_TILE; // two consumer warpgroups
    return b;
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}" : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// a waiting warp sleeps between polls so that it does not take issue slots from the warps it is waiting for.
// A hand-over that never completes would hang the GPU.  The debug build (-DKIN_DEBUG) traps after 2^24 polls
// (seconds, against microseconds per tile); the release build only after 2^28 (about a minute of sleeping -- far
// beyond any slowdown a debugger or sanitizer causes), because a trap is a sticky error for the whole context.
#ifdef KIN_DEBUG
#define KIN_WS_MAX_POLLS (1u << 24)
#else
#define KIN_WS_MAX_POLLS (1u << 28)
#endif
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    unsigned polls = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(KIN_WS_SLEEP_NS);
        if (++polls > KIN_WS_MAX_POLLS) __trap();
    }
}

// The ring is re-used every few microseconds while 4 KB of results per configuration stream through the same L2:
// ring loads carry an evict_last policy and bypass L1, and the result stores are streaming (st.global.cs); with
// plain accesses and a 4-stage ring every ring line was written back to HBM before it was read (+35 % DRAM
// traffic, measured).  Ring stores are plain C++ stores: `asm volatile` stores serialise the producer's schedule.
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void ring_st(double *p, double v, uint64_t) { *p = v; }
__device__ __forceinline__ double ring_ld(const double *p, uint64_t pol) {
    double v;
    asm volatile("ld.global.cg.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol) : "memory");
    return v;
}

// BASE: the model has the planar base (instantiated separately so that the base block is not in the hot loop of
// the fixed-base kernel: instruction fetch is one of its limiters)
template <int LAY, bool BASE, bool PRE>
__global__ void __launch_bounds__(WS_THREADS, 1) kin_eval_ws_kernel(const __grid_constant__ KernelArgs A) {
    typedef double real;
    constexpr bool TILED = LAY == 2;
    constexpr int BS = WS_TILE;
    static_assert(LAY == 0 || LAY == 2, "SoA or tiled");
    static_assert(2 * WS_STAGES * 8 <= 64, "mbarrier area");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const ProgHeader &h = A.h;
    int32_t *ti = reinterpret_cast<int32_t *>(smem_raw);
    real *tr = reinterpret_cast<real *>(smem_raw + sizeof(int32_t) * (size_t)h.n_int);
    size_t off = (sizeof(int32_t) * (size_t)h.n_int + sizeof(real) * (size_t)h.n_real + 15) & ~size_t(15);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + off);      // full[WS_STAGES], empty[WS_STAGES]
    off += 64;
    constexpr int COLS = ws_cols(BASE), FRAME_SLOTS = 6 * COLS;
    real *qs = reinterpret_cast<real *>(smem_raw + off);                // [2][COLS][BS]
    off += sizeof(real) * 2 * COLS * BS;
    real *priv = reinterpret_cast<real *>(smem_raw + off);              // [2 consumers][priv_slots][BS]
    const int ring_slots = ws_ring_slots(h.n_sph, BASE, PRE), priv_slots = ws_priv_slots(h.n_sph, BASE, PRE);
    const int ring_pre = FRAME_SLOTS + 3 * h.n_sph;                    // PRE: first distance slot; argmins follow
    // this CTA's region of the global ring: [WS_STAGES][ring_slots][BS]
    real *ring = reinterpret_cast<real *>(A.ws_ring) + (size_t)blockIdx.x * WS_STAGES * ring_slots * BS;

    const int tid = threadIdx.x, wg = tid / BS, t = tid % BS;
    {
        const int4 *src = reinterpret_cast<const int4 *>(A.tab_i);
        int4 *dst = reinterpret_cast<int4 *>(ti);
        for (int i = tid; i < h.n_int / 4; i += WS_THREADS) dst[i] = src[i];
        const real *rs = reinterpret_cast<const real *>(A.tab_r);
        for (int i = tid; i < h.n_real; i += WS_THREADS) tr[i] = rs[i];
        if (tid == 0) {
            for (int i = 0; i < 2 * WS_STAGES; ++i) mbar_init(&bars[i], BS);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();

    const int DC = h.n_joints, ND = h.n_dof;
    const int n_nodes = h.n_nodes, n_box = h.n_box, S = h.n_sph, n_fk = h.n_fk;
    const int io_node = h.io_node, io_att = h.io_att, io_sph_order = h.io_sph_order, io_sph_mask = h.io_sph_mask;
    const int ro_node = h.ro_node, ro_att = h.ro_att, ro_sph = h.ro_sph, ro_box = h.ro_box;
    unsigned rev_mask = 0;
    for (int j = 0; j < ND; ++j) rev_mask |= (ti[h.io_col_type + j] == 1 ? 1u : 0u) << j;
    const int rows = A.with_rot ? 6 : 3;
    const size_t es = TILED ? size_t(32) : (size_t)A.ld;
    auto rec_base = [&](long long n_, long long rec) -> long long {
        return TILED ? (n_ >> 5) * (rec * 32) + (n_ & 31) : n_;
    };
    const uint64_t pol = l2_evict_last_policy();
    const long long n_tiles = (A.n + BS - 1) / BS;
    const long long my_tiles = (long long)blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (wg == 0) {
        // =========================== PRODUCER: phase 1 ===========================
        if (PRE) asm volatile("setmaxnreg.dec.sync.aligned.u32 120;");      // 120 + 2 x 192 = 3 x 168
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");           //  88 + 2 x 208 = 3 x 168
        auto prefetch_q = [&](long long k_, int buf) {
            const long long n_ = min((blockIdx.x + k_ * gridDim.x) * BS + t, (long long)A.n - 1);
            const real *qn = reinterpret_cast<const real *>(A.q) + rec_base(n_, ND);
            real *dst = qs + (size_t)buf * COLS * BS + t;
            for (int c = 0; c < ND; ++c) cp_async_elem(dst + c * BS, qn + c * es);
            cp_async_commit();
        };
        if (my_tiles > 0) prefetch_q(0, 0);
        for (long long k = 0; k < my_tiles; ++k) {
            const int st = (int)(k % WS_STAGES), qb = (int)(k & 1);
            const long long n = min((blockIdx.x + k * gridDim.x) * BS + t, (long long)A.n - 1);
            if (k + 1 < my_tiles) { prefetch_q(k + 1, qb ^ 1); cp_async_wait<1>(); }
            else cp_async_wait<0>();
            const real *qv = qs + (size_t)qb * COLS * BS + t;           // qv[c * BS]
            real *rg = ring + (size_t)st * ring_slots * BS + t;         // rg[slot * BS], global
            mbar_wait(&bars[WS_STAGES + st], (unsigned)(((k / WS_STAGES) & 1) ^ 1));   // the stage has been copied out
            real *Tn = reinterpret_cast<real *>(A.T_out) + rec_base(n, 12 * n_fk);
            real *Jn = reinterpret_cast<real *>(A.J_out) + rec_base(n, rows * ND * h.n_jac);

            Tf<real> T;
            for (int node = 0; node < n_nodes; ++node) {
                const int32_t *ni = ti + io_node + node * NODE_INTS;
                const real *nr = tr + ro_node + node * NODE_REALS;
                const int jtype = ni[1];
                if (jtype == NODE_ROOT) {
                    #pragma unroll
                    for (int i = 0; i < 9; ++i) T.r[i] = (i % 4 == 0) ? real(1) : real(0);
                    T.p[0] = T.p[1] = T.p[2] = real(0);
                } else {
                    const int flags = ni[2], qcol = ni[3];
                    KIN_DASSERT(qcol >= 0 && qcol < ND && 6 * (qcol + 1) <= FRAME_SLOTS && ni[0] < 0);
                    // A = T_parent * joint.pose : the joint frame (algorithm.jl:47-48); the chain has no branching here
                    Tf<real> Aj;
                    tf_mul_const(T, nr, flags & NF_OFF_R_IDENTITY, Aj);
                    const int code = (flags >> NF_AXIS_SHIFT) & NF_AXIS_MASK;
                    JFrame<real> f;
                    f.o[0] = Aj.p[0]; f.o[1] = Aj.p[1]; f.o[2] = Aj.p[2];
                    const real sgn = code >= 4 ? real(-1) : real(1);
                    switch (code) {
                        case 1: case 4: f.a[0] = sgn * Aj.r[0]; f.a[1] = sgn * Aj.r[3]; f.a[2] = sgn * Aj.r[6]; break;
                        case 2: case 5: f.a[0] = sgn * Aj.r[1]; f.a[1] = sgn * Aj.r[4]; f.a[2] = sgn * Aj.r[7]; break;
                        case 3: case 6: f.a[0] = sgn * Aj.r[2]; f.a[1] = sgn * Aj.r[5]; f.a[2] = sgn * Aj.r[8]; break;
                        default:
                            f.a[0] = fma_(Aj.r[0], nr[12], fma_(Aj.r[1], nr[13], Aj.r[2] * nr[14]));
                            f.a[1] = fma_(Aj.r[3], nr[12], fma_(Aj.r[4], nr[13], Aj.r[5] * nr[14]));
                            f.a[2] = fma_(Aj.r[6], nr[12], fma_(Aj.r[7], nr[13], Aj.r[8] * nr[14]));
                    }
                    {
                        real *jf = rg + 6 * BS * qcol;
                        ring_st(jf, f.o[0], pol); ring_st(jf + BS, f.o[1], pol); ring_st(jf + 2 * BS, f.o[2], pol);
                        ring_st(jf + 3 * BS, f.a[0], pol); ring_st(jf + 4 * BS, f.a[1], pol); ring_st(jf + 5 * BS, f.a[2], pol);
                    }
                    const real qa = qv[qcol * BS];
                    T = Aj;
                    if (jtype == 2) {          // prismatic, mechanism.jl:100-103
                        T.p[0] = fma_(f.a[0], qa, Aj.p[0]); T.p[1] = fma_(f.a[1], qa, Aj.p[1]); T.p[2] = fma_(f.a[2], qa, Aj.p[2]);
                    } else {                   // revolute, mechanism.jl:94-98
                        real s, c;
                        sincos_(code >= 4 ? -qa : qa, &s, &c);
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
                                const real x = nr[12], y = nr[13], z = nr[14], tt = real(1) - c;
                                real m[9];
                                m[0] = fma_(tt * x, x, c);        m[1] = fma_(tt * x, y, -(s * z)); m[2] = fma_(tt * x, z, s * y);
                                m[3] = fma_(tt * x, y, s * z);    m[4] = fma_(tt * y, y, c);        m[5] = fma_(tt * y, z, -(s * x));
                                m[6] = fma_(tt * x, z, -(s * y)); m[7] = fma_(tt * y, z, s * x);    m[8] = fma_(tt * z, z, c);
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
                // ---- requested links hanging from this node (none in a collision-only call) ----
                if (!PRE) for (int a = ni[5]; a < ni[6]; ++a) {
                    const int32_t *ai = ti + io_att + a * ATT_INTS;
                    const real *ar = tr + ro_att + a * ATT_REALS;
                    Tf<real> Tl;
                    tf_mul_const(T, ar, ai[1] & AF_R_IDENTITY, Tl);
                    if (ai[0] >= 0 && A.T_out) {       // get_transform, as 3x4 column-major
                        real *o = Tn + (size_t)(12 * ai[0]) * es;
                        #pragma unroll
                        for (int c = 0; c < 3; ++c)
                            #pragma unroll
                            for (int r = 0; r < 3; ++r) __stcs(&o[(c * 3 + r) * es], Tl.r[r * 3 + c]);
                        #pragma unroll
                        for (int r = 0; r < 3; ++r) __stcs(&o[(9 + r) * es], Tl.p[r]);
                    }
                    if (ai[2] >= 0 && A.J_out) {       // get_jacobian, algorithm.jl:83-114
                        real *o = Jn + (size_t)(ai[2] * rows * ND) * es;
                        const unsigned mask = (unsigned)ai[3];
                        real k[6] = {0, 0, 0, 0, 0, 0};
                        if (A.with_rot && A.rpy_jac) rpy_rate_coeffs(Tl, k);
                        const real *jf = rg;
                        #pragma unroll 1
                        for (int j = 0; j < ND; ++j) {
                            if ((mask >> j) & 1u) {
                                JFrame<real> f;
                                // own stores of this tile, read back through L2 (once per requested Jacobian)
                                f.o[0] = ring_ld(jf, pol); f.o[1] = ring_ld(jf + BS, pol); f.o[2] = ring_ld(jf + 2 * BS, pol);
                                f.a[0] = ring_ld(jf + 3 * BS, pol); f.a[1] = ring_ld(jf + 4 * BS, pol); f.a[2] = ring_ld(jf + 5 * BS, pol);
                                real cx, cy, cz;
                                const bool rev = (rev_mask >> j) & 1u;
                                jac_col(f, rev, Tl.p[0], Tl.p[1], Tl.p[2], cx, cy, cz);
                                __stcs(&o[0], cx); __stcs(&o[es], cy); __stcs(&o[2 * es], cz);
                                if (A.with_rot) {
                                    if (rev) {
                                        if (A.rpy_jac) {
                                            real o3, o4, o5;
                                            rpy_rows(k, f.a[0], f.a[1], f.a[2], o3, o4, o5);
                                            __stcs(&o[3 * es], o3); __stcs(&o[4 * es], o4); __stcs(&o[5 * es], o5);
                                        } else {
                                            __stcs(&o[3 * es], f.a[0]); __stcs(&o[4 * es], f.a[1]); __stcs(&o[5 * es], f.a[2]);
                                        }
                                    } else if (!A.keep_irrelevant || j >= DC) {
                                        __stcs(&o[3 * es], real(0)); __stcs(&o[4 * es], real(0)); __stcs(&o[5 * es], real(0));
                                    }
                                }
                            } else if (!A.keep_irrelevant) {
                                for (int r = 0; r < rows; ++r) __stcs(&o[r * es], real(0));
                            }
                            o += rows * es;
                            jf += 6 * BS;
                        }
                    }
                }
                // ---- collision-sphere centres on this node (collision.jl:54 / :80) ----
                for (int k2 = ni[7]; k2 < ni[8]; ++k2) {
                    const int s = ti[io_sph_order + k2];
                    KIN_DASSERT(s >= 0 && s < S && FRAME_SLOTS + 3 * (s + 1) <= ring_slots);
                    const real *sr = tr + ro_sph + s * SPH_REALS;
                    const real c0 = sr[0], c1 = sr[1], c2 = sr[2];
                    real *cs = rg + (FRAME_SLOTS + 3 * s) * BS;
                    #pragma unroll
                    for (int i = 0; i < 3; ++i)
                        ring_st(cs + i * BS, fma_(T.r[i * 3 + 0], c0, fma_(T.r[i * 3 + 1], c1, fma_(T.r[i * 3 + 2], c2, T.p[i]))), pol);
                }
            }
            if (PRE) {
                // ---- phase 2a of every sphere (the consumers' code, so bitwise the same distances / argmins); the
                //      centres are this thread's own stores of a moment ago, read back through L2 ----
                for (int s0 = 0; s0 < S; s0 += SPH_GROUP) {
                    real px[SPH_GROUP], py[SPH_GROUP], pz[SPH_GROUP], kmin[SPH_GROUP];
                    int kidx[SPH_GROUP];
                    #pragma unroll
                    for (int g = 0; g < SPH_GROUP; ++g) {
                        const real *cs = rg + (FRAME_SLOTS + 3 * min(s0 + g, S - 1)) * BS;
                        px[g] = ring_ld(cs, pol); py[g] = ring_ld(cs + BS, pol); pz[g] = ring_ld(cs + 2 * BS, pol);
                        kmin[g] = CUDART_INF; kidx[g] = 0;
                    }
                    for (int b = 0; b < n_box; ++b) {
                        BoxRow<real> row;
                        load_box(tr + ro_box + b * BOX_REALS, row);
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
                    for (int g = 0; g < SPH_GROUP; ++g)
                        if (s0 + g < S) {
                            ring_st(rg + (ring_pre + s0 + g) * BS, key_to_dist(kmin[g]), pol);
                            ring_st(rg + (ring_pre + S + s0 + g) * BS, (real)kidx[g], pol);
                        }
                }
            }
            mbar_arrive(&bars[st]);                   // full: frames + centres of this tile are in the ring
        }
    } else {
        // =========================== CONSUMERS: phase 2 ===========================
        if (PRE) asm volatile("setmaxnreg.inc.sync.aligned.u32 192;");
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
        constexpr int JR = JF_REGS;
        // the output stride in a register the compiler cannot trace back to the kernel parameters (it otherwise
        // re-loads it from the constant bank in front of the stores of the column loop: LDC, 2.6 % of the stalls)
        size_t es_c = es;
        if (!TILED) asm volatile("" : "+l"(es_c));
        const int cw = wg - 1;                         // consumer 0 / 1 takes the even / odd tiles of this CTA
        real *pv = priv + (size_t)cw * priv_slots * BS + t;             // pv[slot * BS]: this thread's private state
        const bool want_grads = A.grads_out != nullptr;
        const bool stale = want_grads && A.scratch_ref;
        const real trunc = (real)A.truncation_dist, voff = (real)A.vals_offset;
        // fixed-size areas first: their slot offsets are compile-time constants (one base register + immediates;
        // with the S-dependent centres in front the compiler re-derived these addresses from the kernel parameters
        // in front of every access of the column loop)
        real *stale0 = pv;
        real *hand = pv + (3 * COLS) * BS;                                // !PRE: (dmin, argmin) of the current group
        real *pre_d = hand;                                               //  PRE: S distances, then S int32 argmins
        int *pre_k = reinterpret_cast<int *>(pv + (3 * COLS + S) * BS - t) + t;   // int column of this thread: pre_k[s * BS]
        real *cent0 = pv + (3 * COLS + ws_hand_slots(S, PRE)) * BS;
        constexpr bool with_base = BASE;
        // the control-joint columns (frames in registers, static indices); the base columns follow separately
        #define FOR_COLUMNS(j) _Pragma("unroll") for (int j = 0; j < JR; ++j) if (j >= DC) break; else
        for (long long k = cw; k < my_tiles; k += 2) {
            const int st = (int)(k % WS_STAGES);
            const long long n = min((blockIdx.x + k * gridDim.x) * BS + t, (long long)A.n - 1);
            const real *rg = ring + (size_t)st * ring_slots * BS + t;
            mbar_wait(&bars[st], (unsigned)((k / WS_STAGES) & 1));       // full
            // copy out: centres -> private shared memory (asynchronous global -> shared copies: no registers, all in
            // flight together with the frame loads), frames -> registers; then the stage is free again.
            // cp.async.ca may allocate in L1; producer and consumer share this SM's L1 and synchronise at CTA scope,
            // so a later refill of the stage is seen.
            {
                const real *src = rg + FRAME_SLOTS * BS;
                for (int i = 0; i < 3 * S; ++i) cp_async_elem(cent0 + i * BS, src + i * BS);
                if (PRE) for (int i = 0; i < S; ++i) cp_async_elem(pre_d + i * BS, rg + (ring_pre + i) * BS);
                cp_async_commit();
                if (PRE) for (int i = 0; i < S; ++i) pre_k[i * BS] = (int)ring_ld(rg + (ring_pre + S + i) * BS, pol);
            }
            JFrame<real> jfr[JR];
            FOR_COLUMNS(j) {
                const real *jf = rg + 6 * BS * j;
                jfr[j].o[0] = ring_ld(jf, pol); jfr[j].o[1] = ring_ld(jf + BS, pol); jfr[j].o[2] = ring_ld(jf + 2 * BS, pol);
                jfr[j].a[0] = ring_ld(jf + 3 * BS, pol); jfr[j].a[1] = ring_ld(jf + 4 * BS, pol); jfr[j].a[2] = ring_ld(jf + 5 * BS, pol);
            }
            real bx = real(0), by = real(0);           // origin of the base's revolute column: (x, y, 0)
            if (with_base) { bx = ring_ld(rg + 6 * BS * (DC + 2), pol); by = ring_ld(rg + 6 * BS * (DC + 2) + BS, pol); }
            cp_async_wait<0>();
            mbar_arrive(&bars[WS_STAGES + st]);        // empty: the producer may refill this stage
            if (stale)
                for (int i = 0; i < 3 * ND; ++i) stale0[i * BS] = real(0);   // jac = zeros(3, n_dof), collision.jl:76
            real *Vp = reinterpret_cast<real *>(A.vals_out) + rec_base(n, S);
            real *Gp = reinterpret_cast<real *>(A.grads_out) + rec_base(n, (long long)ND * S);
            int32_t *Ap = A.argmin_out ? A.argmin_out + rec_base(n, S) : nullptr;

            for (int s0 = 0; s0 < S; s0 += SPH_GROUP) {
                // ---- 2a: distances of SPH_GROUP spheres; one box-table row feeds all of them (PRE: done by the producer) ----
                if (!PRE) {
                    real px[SPH_GROUP], py[SPH_GROUP], pz[SPH_GROUP], kmin[SPH_GROUP];
                    int kidx[SPH_GROUP];
                    #pragma unroll
                    for (int g = 0; g < SPH_GROUP; ++g) {
                        const real *cs = cent0 + 3 * min(s0 + g, S - 1) * BS;
                        px[g] = cs[0]; py[g] = cs[BS]; pz[g] = cs[2 * BS];
                        kmin[g] = CUDART_INF; kidx[g] = 0;
                    }
                    for (int b = 0; b < n_box; ++b) {         // UnionSDF: all boxes, first minimum wins (sdf.jl:108-114)
                        BoxRow<real> row;
                        load_box(tr + ro_box + b * BOX_REALS, row);
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
                // ---- 2b: per sphere, IN sphere order: value, truncation, gradient, chain rule ----
                #pragma unroll 1
                for (int g = 0; g < SPH_GROUP && s0 + g < S; ++g) {
                    const int s = s0 + g;
                    const real dmin = PRE ? pre_d[s * BS] : hand[g * BS];
                    const int kmin = PRE ? pre_k[s * BS] : reinterpret_cast<const int *>(&hand[(SPH_GROUP + g) * BS])[0];
                    KIN_DASSERT(kmin >= 0 && kmin < n_box && st >= 0 && st < WS_STAGES);
                    const real dist0 = dmin - tr[ro_sph + s * SPH_REALS + 3];
                    const bool truncated = dist0 > trunc;
                    __stcs(Vp, (truncated ? trunc : dist0) - voff);
                    Vp += es_c;
                    if (Ap) { __stcs(Ap, kmin + 1); Ap += es_c; }
                    if (!want_grads) continue;
                    if (truncated) {            // collision.jl:84-86
                        for (int j = 0; j < ND; ++j) __stcs(&Gp[(size_t)j * es_c], real(0));
                        Gp += (size_t)ND * es_c;
                        continue;
                    }
                    const real *cs = cent0 + 3 * s * BS;
                    const real px = cs[0], py = cs[BS], pz = cs[2 * BS];
                    real grad[3];
                    {
                        BoxRow<real> row;
                        load_box(tr + ro_box + kmin * BOX_REALS, row);
                        box_gradient(row, A.grad_mode, px, py, pz, dmin, grad);
                    }
                    const unsigned mask = (unsigned)ti[io_sph_mask + s];
                    real *sp = stale0;
                    FOR_COLUMNS(j) {
                        real cx, cy, cz;
                        if ((mask >> j) & 1u) {   // joint_jacobian!, algorithm.jl:65-81
                            jac_col(jfr[j], (rev_mask >> j) & 1u, px, py, pz, cx, cy, cz);
                            if (stale) { sp[0] = cx; sp[BS] = cy; sp[2 * BS] = cz; }
                        } else if (stale) {     // column left over from an earlier sphere (collision.jl:76,90)
                            cx = sp[0]; cy = sp[BS]; cz = sp[2 * BS];
                        } else { cx = cy = cz = real(0); }
                        __stcs(&Gp[(size_t)j * es_c], fma_(grad[0], cx, fma_(grad[1], cy, grad[2] * cz)));   // transpose(grad) * jac
                        sp += 3 * BS;
                    }
                    if (with_base) {              // base block, algorithm.jl:98-105: same jac_col, constant frames
                        #pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const int j = DC + k;
                            JFrame<real> f;
                            f.o[0] = bx; f.o[1] = by; f.o[2] = real(0);
                            f.a[0] = k == 0 ? real(1) : real(0); f.a[1] = k == 1 ? real(1) : real(0); f.a[2] = k == 2 ? real(1) : real(0);
                            real cx, cy, cz;
                            if ((mask >> j) & 1u) {
                                jac_col(f, k == 2, px, py, pz, cx, cy, cz);
                                if (stale) { sp[0] = cx; sp[BS] = cy; sp[2 * BS] = cz; }
                            } else if (stale) {
                                cx = sp[0]; cy = sp[BS]; cz = sp[2 * BS];
                            } else { cx = cy = cz = real(0); }
                            __stcs(&Gp[(size_t)j * es_c], fma_(grad[0], cx, fma_(grad[1], cy, grad[2] * cz)));
                            sp += 3 * BS;
                        }
                    }
                    Gp += (size_t)ND * es_c;
                }
            }
        }
        #undef FOR_COLUMNS
    }
}

}  // namespace kin
