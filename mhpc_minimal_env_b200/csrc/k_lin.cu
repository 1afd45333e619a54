// Linearisation ("forward_sweep_partials_only", SinglePhase.cpp:147-180, and the derivative half of
// forward_sweep, :82-111): one thread per (knot, instance), lanes = 32 consecutive instances.
// Compiled once per knot class (-DLIN_CLS=...) so the large generated Jacobian bodies build in
// parallel and each gets its own register allocation:
//   CLS_BS / CLS_FS / CLS_FL   whole-body step knots: Dyn_{BS,FS,FL}_par + Euler A,B (PlanarQuadruped.cpp:30-53)
//                              + running-cost derivatives (CostBase.cpp:19-34) + barrier derivatives (SinglePhase.cpp:239-246)
//   CLS_IMPF / CLS_IMPB        whole-body terminal knots of flight phases: terminal cost derivatives (CostBase.cpp:50-60),
//                              touchdown constraint + AL terms (MHPCConstraints.cpp:91-107, SinglePhase.cpp:257-275),
//                              impact Jacobian Px (PlanarQuadruped.cpp:82-100, used by MultiPhaseDDP.cpp:320)
//   CLS_LIGHT                  stance-phase terminal knots (Px = I), rigid-body step and terminal knots
// Every thread writes one knot record of the `lin` array (common.cuh).
#include <stdlib.h>

#include "kernels.h"
#include "model.cuh"

#ifndef LIN_CLS
#error "compile with -DLIN_CLS=<knot class>"
#endif

#include "gen/gen_meta.h"

namespace mhpc {
namespace {

constexpr int LIN_THREADS = 128;
#ifndef LIN_MINBLOCKS
#define LIN_MINBLOCKS 1
#endif
#if LIN_CLS <= 4
constexpr size_t LIN_SMEM = (size_t)(MHPC_RA_SMEM_SLOTS + MHPC_RA_STAGE_SLOTS) * LIN_THREADS * sizeof(double);
#else
constexpr size_t LIN_SMEM = 0;
#endif

#if LIN_CLS <= 4 && MHPC_RA_STAGE_SLOTS > 0
// ---- asynchronous far spill level (codegen/regalloc.py, stage_slots > 0) ----
// Values that fit neither the register file nor the shared-memory slots go to a scratch array in global memory that is
// indexed by RESIDENT block slot, not by logical thread: (SM id, one of FAR_PER_SM slots, thread) - 148 x 3 x 128 threads x
// ~130 doubles = 59 MB however large the launch, so it lives in L2 (the knot records leave with an evict-first hint) and the
// spills cost no DRAM traffic.  A block takes a free slot of its SM when it starts and returns it when it ends; at most
// LIN_MINBLOCKS blocks of this kernel fit an SM (shared memory), so one of the LIN_MINBLOCKS + 1 slots is always free.
constexpr int FAR_SM_IDS = 256;                         // %smid values are not promised to be dense: sized generously, only touched rows cost
constexpr int FAR_PER_SM = LIN_MINBLOCKS + 1;
constexpr size_t FAR_STRIDE = (size_t)FAR_SM_IDS * FAR_PER_SM * LIN_THREADS;   // elements between consecutive far slots of a thread
constexpr size_t FAR_DOUBLES = FAR_STRIDE * MHPC_RA_FAR_SLOTS;
struct FarScratch { double* buf; unsigned* mask; };
__device__ __forceinline__ int far_acquire(unsigned* mask, int* s_slot) {
    if (threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        unsigned* m = mask + smid;
        int got = -1;
        while (got < 0) {
            for (int q = 0; q < FAR_PER_SM && got < 0; ++q) {
                const unsigned old = atomicOr(m, 1u << q);
                if (!(old & (1u << q))) got = q;
            }
        }
        *s_slot = (int)smid * FAR_PER_SM + got;
    }
    __syncthreads();
    return *s_slot;
}
__device__ __forceinline__ void far_release(unsigned* mask, int slot) {
    __syncthreads();      // every thread of the block is done with its scratch
    if (threadIdx.x == 0) atomicAnd(mask + slot / FAR_PER_SM, ~(1u << (slot % FAR_PER_SM)));
}
__device__ __forceinline__ void sx_far_st(double* p, double v) { asm volatile("st.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ double sx_far_ld(const double* p) { double v; asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void sx_far_issue(unsigned smem_addr, const double* p) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n\tcp.async.commit_group;" ::"r"(smem_addr), "l"(p) : "memory");
}
#define SX_FAR_DECL(n) static_assert((n) <= MHPC_RA_FAR_SLOTS, "far scratch: more slots than gen_meta.h promises")
#define SX_FAR_ST(s, r) sx_far_st(sx_far + (size_t)(s) * FAR_STRIDE, r)
#define SX_FAR_LD(s) sx_far_ld(sx_far + (size_t)(s) * FAR_STRIDE)
#define SX_FAR_ISSUE(st_, s) sx_far_issue(sx_stage_a + (unsigned)(st_) * (unsigned)(LIN_THREADS * sizeof(double)), sx_far + (size_t)(s) * FAR_STRIDE)
#define SX_FAR_WAIT(n) asm volatile("cp.async.wait_group %0;" ::"n"(n) : "memory")
#define SX_STAGE(st_) sx_sm[(MHPC_RA_SMEM_SLOTS + (st_)) * LIN_THREADS]
#define SX_FAR_PARAMS , double* sx_far, unsigned sx_stage_a
#define SX_FAR_ARGS , sx_far, sx_stage_a
#else
#define SX_FAR_PARAMS
#define SX_FAR_ARGS
#endif

#define REC_AT(e) rec[(e)]

template <int K, int ROW, int COL>
__device__ __forceinline__ void wb_par_out(double* __restrict__ rec, int j, double dt, double v) {
    if (K == 0) {
        if (ROW >= 7) REC_AT(R_AB(ROW - 7, COL - 2)) = (ROW == COL ? 1.0 : 0.0) + v * dt;   // A = I + Ac*dt
    } else if (K == 1) {
        REC_AT(R_BB(ROW - 7, COL)) = v * dt;                                                // B = Bc*dt
    } else if (K == 2) {
        REC_AT(R_CB(ROW & 1, COL - 2)) = v;
    } else {
        REC_AT(R_DB(ROW & 1, COL)) = v;
    }
}
// One aligned quad of record elements leaves as a single 32-byte store: the record is AoS (1408 bytes per knot and
// instance), so a scalar store per element touches a 32-byte sector four times with 8 bytes each - 4x the L2 write
// transactions and partial-sector merges - while a quad is exactly one full sector per thread.
#ifndef LIN_ST_HINT
#define LIN_ST_HINT 1     // the knot records leave as streaming stores (.cs): they are not read again before the sweep, and the spill slab of the
                          // resident threads stays in L2 (0: plain stores, 2: an explicit L2 evict-first policy; 0.993 / 0.972 / 0.988 ms per
                          // linearisation for 0 / 1 / 2 at the end of round 2)
#endif
#if LIN_ST_HINT == 2
// the policy word is made once per thread (k_lin: lin_pol) - createpolicy has the latency of a memory operation
__device__ __forceinline__ unsigned long long make_evict_first_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
#define ST_QUAD_POL , unsigned long long lin_pol
#define ST_QUAD_ARG , lin_pol
#else
#define ST_QUAD_POL
#define ST_QUAD_ARG
#endif
__device__ __forceinline__ void st_quad(double* p, double a, double b, double c, double d ST_QUAD_POL) {
#if LIN_ST_HINT == 1
    asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
#elif LIN_ST_HINT == 2
    asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d), "l"(lin_pol) : "memory");
#else
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
#endif
}
template <int K, int ROW, int COL>
__device__ __forceinline__ constexpr int wb_par_elem() {
    return K == 0 ? R_AB(ROW - 7, COL - 2) : K == 1 ? R_BB(ROW - 7, COL) : K == 2 ? R_CB(ROW & 1, COL - 2) : R_DB(ROW & 1, COL);
}
template <int K, int ROW, int COL>
__device__ __forceinline__ double wb_par_val(double dt, double v) {
    return K == 0 ? (ROW == COL ? 1.0 : 0.0) + v * dt : K == 1 ? v * dt : v;
}
template <int E0, int K0, int R0, int C0, int K1, int R1, int C1, int K2, int R2, int C2, int K3, int R3, int C3>
__device__ __forceinline__ void wb_par_out4(double* __restrict__ rec, double dt, double v0, double v1, double v2, double v3 ST_QUAD_POL) {
    static_assert((E0 & 3) == 0 && wb_par_elem<K0, R0, C0>() == E0 && wb_par_elem<K1, R1, C1>() == E0 + 1 &&
                  wb_par_elem<K2, R2, C2>() == E0 + 2 && wb_par_elem<K3, R3, C3>() == E0 + 3,
                  "codegen/gen_all.py: wb_record_element disagrees with the record layout");
    st_quad(rec + E0, wb_par_val<K0, R0, C0>(dt, v0), wb_par_val<K1, R1, C1>(dt, v1), wb_par_val<K2, R2, C2>(dt, v2),
            wb_par_val<K3, R3, C3>(dt, v3) ST_QUAD_ARG);
}
template <int ROW, int COL>
__device__ __forceinline__ void imp_par_out(double* __restrict__ rec, int j, double v) {
    if (ROW >= 7) REC_AT(R_AB(ROW - 7, COL - 2)) = v;
}
template <int E0, int R0, int C0, int R1, int C1, int R2, int C2, int R3, int C3>
__device__ __forceinline__ void imp_par_out4(double* __restrict__ rec, double v0, double v1, double v2, double v3 ST_QUAD_POL) {
    static_assert((E0 & 3) == 0 && R_AB(R0 - 7, C0 - 2) == E0 && R_AB(R1 - 7, C1 - 2) == E0 + 1 && R_AB(R2 - 7, C2 - 2) == E0 + 2 &&
                  R_AB(R3 - 7, C3 - 2) == E0 + 3, "codegen/gen_all.py: imp_record_element disagrees with the record layout");
    st_quad(rec + E0, v0, v1, v2, v3 ST_QUAD_ARG);
}

// The Jacobian bodies come in the generator-allocated form (codegen/regalloc.py: a fixed file of FP64 register
// variables + an explicit volatile local spill array with Belady eviction and software prefetch); inputs are read
// straight from the SoA trajectory (`st`, element stride `es`), outputs go straight to the knot record.
// (a volatile asm load: it stays where the generator put it - far ahead of its consumer - and is never re-executed)
__device__ __forceinline__ double sx_ld_in(const double* p) { double v; asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }
#define SX_IN_GLOBAL(i, j_) sx_ld_in(st + (size_t)((i) == 0 ? (j_) : 14 + (j_)) * es)
// whole-body step knots: the caller has the state and the control in registers already (it needed them for the cost partials)
#define SX_IN(i, j_) ((i) == 0 ? xin[(j_)] : uin[(j_)])
#define SX_SMEM(s) sx_sm[(s) * LIN_THREADS]     /* second spill level: per-thread shared-memory slots, conflict-free */
#define SX_OUT(k, nz, row, col, v) wb_par_out<k, row, col>(rec, j, dt, v)
#define SX_OUT4(e0, k0, n0, r0, c0, v0, k1, n1, r1, c1, v1, k2, n2, r2, c2, v2, k3, n3, r3, c3, v3) \
    wb_par_out4<e0, k0, r0, c0, k1, r1, c1, k2, r2, c2, k3, r3, c3>(rec, dt, v0, v1, v2, v3 ST_QUAD_ARG)
#if LIN_CLS == 0
__device__ __forceinline__ void dyn_par(const double* xin, const double* uin, double* __restrict__ rec, int j, double dt, volatile double* sx_sm SX_FAR_PARAMS ST_QUAD_POL) {
#include "gen/Dyn_BS_par_ra.inc"
}
#elif LIN_CLS == 1
__device__ __forceinline__ void dyn_par(const double* xin, const double* uin, double* __restrict__ rec, int j, double dt, volatile double* sx_sm SX_FAR_PARAMS ST_QUAD_POL) {
#include "gen/Dyn_FS_par_ra.inc"
}
#elif LIN_CLS == 2
__device__ __forceinline__ void dyn_par(const double* xin, const double* uin, double* __restrict__ rec, int j, double dt, volatile double* sx_sm SX_FAR_PARAMS ST_QUAD_POL) {
#include "gen/Dyn_FL_par_ra.inc"
}
#endif
#undef SX_OUT
#undef SX_OUT4
#undef SX_IN
#define SX_IN(i, j_) SX_IN_GLOBAL(i, j_)
#define SX_OUT(k, nz, row, col, v) imp_par_out<row, col>(rec, j, v)
#define SX_OUT4(e0, k0, n0, r0, c0, v0, k1, n1, r1, c1, v1, k2, n2, r2, c2, v2, k3, n3, r3, c3, v3) \
    imp_par_out4<e0, r0, c0, r1, c1, r2, c2, r3, c3>(rec, v0, v1, v2, v3 ST_QUAD_ARG)
#if LIN_CLS == 3
__device__ __forceinline__ void imp_par(const double* __restrict__ st, size_t es, double* __restrict__ rec, int j, volatile double* sx_sm SX_FAR_PARAMS ST_QUAD_POL) {
#include "gen/Imp_F_par_ra.inc"
}
#elif LIN_CLS == 4
__device__ __forceinline__ void imp_par(const double* __restrict__ st, size_t es, double* __restrict__ rec, int j, volatile double* sx_sm SX_FAR_PARAMS ST_QUAD_POL) {
#include "gen/Imp_B_par_ra.inc"
}
#endif
#undef SX_OUT
#undef SX_OUT4

// ---- standalone evaluation of the same Jacobian bodies (mhpc_eval_function "Dyn_*_par" / "Imp_*_par"): the CasADi
// function itself - raw Ac, Bc, C, D (or Px) scattered into dense column-major outputs, the structural zeros left to
// the caller's memset (CasadiGen.cpp:46-68) - with no Euler step and no knot record in between.
#if LIN_CLS <= 4
template <int K, int ROW, int COL>
__device__ __forceinline__ void ev_out(double* o0, double* o1, double* o2, double* o3, double v) {
    if (K == 0) o0[ROW + 14 * COL] = v;
    else if (K == 1) o1[ROW + 14 * COL] = v;
    else if (K == 2) o2[ROW + 4 * COL] = v;
    else o3[ROW + 4 * COL] = v;
}
#define SX_OUT(k, nz, row, col, v) ev_out<k, row, col>(o0, o1, o2, o3, v)
#define SX_OUT4(e0, k0, n0, r0, c0, v0, k1, n1, r1, c1, v1, k2, n2, r2, c2, v2, k3, n3, r3, c3, v3) \
    do { ev_out<k0, r0, c0>(o0, o1, o2, o3, v0); ev_out<k1, r1, c1>(o0, o1, o2, o3, v1);         \
         ev_out<k2, r2, c2>(o0, o1, o2, o3, v2); ev_out<k3, r3, c3>(o0, o1, o2, o3, v3); } while (0)
__device__ __forceinline__ void par_eval_body(const double* __restrict__ st, size_t es, double* o0, double* o1, double* o2, double* o3,
                                              volatile double* sx_sm SX_FAR_PARAMS) {
#if LIN_CLS == 0
#include "gen/Dyn_BS_par_ra.inc"
#elif LIN_CLS == 1
#include "gen/Dyn_FS_par_ra.inc"
#elif LIN_CLS == 2
#include "gen/Dyn_FL_par_ra.inc"
#elif LIN_CLS == 3
#include "gen/Imp_F_par_ra.inc"
#else
#include "gen/Imp_B_par_ra.inc"
#endif
}
#undef SX_OUT
#undef SX_OUT4
#endif
#undef SX_IN
#undef SX_IN_GLOBAL
#undef SX_SMEM

// running-cost derivatives of a WB step knot, with barrier terms when `reb` (SinglePhase.cpp:219-249)
__device__ inline void wb_cost_partials(const Consts& c, int mode, double dt, const double* x, const double* u, const double* y,
                                        double pos, double vel, double height, bool reb, const double* par,
                                        double* __restrict__ rec, int j ST_QUAD_POL) {
    const double two_dt = 2 * dt;
    double lx[14], lu[4], lxxq[4], luu[4];
#pragma unroll
    for (int i = 0; i < 14; ++i) lx[i] = (two_dt * c.wbQ[i]) * (x[i] - wb_xref(c, i, pos, vel, height));
#pragma unroll
    for (int i = 0; i < 4; ++i) { lu[i] = (two_dt * c.wbR[mode - 1][i]) * (u[i] - 0.0); luu[i] = two_dt * c.wbR[mode - 1][i]; lxxq[i] = two_dt * c.wbQ[3 + i]; }
    const int r0 = (mode == 1) ? 2 : 0;  // stance rows of y (mode 3: rows 0,1); flight: unused
    double ly2[2] = {0.0, 0.0}, lyy[3] = {0.0, 0.0, 0.0};
    if (mode == 1 || mode == 3) {
        ly2[0] = (two_dt * c.wbS[mode - 1][r0]) * (y[r0] - grf_ref(c, r0));
        ly2[1] = (two_dt * c.wbS[mode - 1][r0 + 1]) * (y[r0 + 1] - grf_ref(c, r0 + 1));
        lyy[0] = two_dt * c.wbS[mode - 1][r0];
        lyy[2] = two_dt * c.wbS[mode - 1][r0 + 1];
    }
    if (reb) {
        double bz, bzz;
        const double dT = par[0], dJ = par[1], dG = par[2], eT = par[3], eJ = par[4], eG = par[5];
#pragma unroll
        for (int i = 0; i < 8; ++i) {  // torque rows: g = -/+u + 33, gu = -/+1
            const double sgn = (i < 4) ? -1.0 : 1.0;
            const int a = i & 3;
            reb_eval_partials(sgn * u[a] + c.btorque, dT, bz, bzz);
            lu[a] += ((eT * bz) * sgn) * dt;
            luu[a] += (eT * ((sgn * bzz) * sgn + bz * 0.0)) * dt;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {  // joint rows: g = -/+q + b, gx = -/+1 on q[3:7]
            const double sgn = (i < 4) ? -1.0 : 1.0;
            const int a = i & 3;
            reb_eval_partials(sgn * x[3 + a] + c.bjoint[i], dJ, bz, bzz);
            lx[3 + a] += ((eJ * bz) * sgn) * dt;
            lxxq[a] += (eJ * ((sgn * bzz) * sgn + bz * 0.0)) * dt;
        }
        if (mode == 1 || mode == 3) {  // GRF rows on the stance foot: Fz >= 0, -/+Fx + mu Fz >= 0
            const double fx = y[r0], fz = y[r0 + 1];
            const double gy[3][2] = {{0.0, 1.0}, {-1.0, c.mu}, {1.0, c.mu}};
            const double g3[3] = {fz, -fx + c.mu * fz, fx + c.mu * fz};
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                reb_eval_partials(g3[i], dG, bz, bzz);
                ly2[0] += ((eG * bz) * gy[i][0]) * dt;
                ly2[1] += ((eG * bz) * gy[i][1]) * dt;
                lyy[0] += (eG * ((gy[i][0] * bzz) * gy[i][0] + bz * 0.0)) * dt;
                lyy[1] += (eG * ((gy[i][0] * bzz) * gy[i][1] + bz * 0.0)) * dt;
                lyy[2] += (eG * ((gy[i][1] * bzz) * gy[i][1] + bz * 0.0)) * dt;
            }
        }
    }
    static_assert(R_LX == 144 && R_LU == 158 && R_LY == 162 && R_LXXQ == 164 && R_LUU == 168 && R_LYY == 172 && REC == 176,
                  "the eight quads below spell out the cost part of the record");
    st_quad(rec + 144, lx[0], lx[1], lx[2], lx[3] ST_QUAD_ARG);
    st_quad(rec + 148, lx[4], lx[5], lx[6], lx[7] ST_QUAD_ARG);
    st_quad(rec + 152, lx[8], lx[9], lx[10], lx[11] ST_QUAD_ARG);
    st_quad(rec + 156, lx[12], lx[13], lu[0], lu[1] ST_QUAD_ARG);
    st_quad(rec + 160, lu[2], lu[3], ly2[0], ly2[1] ST_QUAD_ARG);
    st_quad(rec + 164, lxxq[0], lxxq[1], lxxq[2], lxxq[3] ST_QUAD_ARG);
    st_quad(rec + 168, luu[0], luu[1], luu[2], luu[3] ST_QUAD_ARG);
    st_quad(rec + 172, lyy[0], lyy[1], lyy[2], 0.0 ST_QUAD_ARG);
}

__device__ __forceinline__ void lin_body(const Dev& d, const Consts& c, const Opt& opt, int mask_mode, const int* __restrict__ knots, int n_pairs,
                                         volatile double* sx_sm SX_FAR_PARAMS) {
    (void)sx_sm;
#if LIN_ST_HINT == 2
    const unsigned long long lin_pol = make_evict_first_policy();
#endif
    int i, kk;
    if (n_pairs > 0) {   // non-uniform batch: explicit (knot, instance) work list of this knot class
        const int t = blockIdx.x * blockDim.x + threadIdx.x;
        if (t >= n_pairs) return;
        const unsigned pr = ((const unsigned*)knots)[t];
        kk = (int)(pr >> 20);
        i = (int)(pr & 0xFFFFFu);
    } else {
        i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= d.n) return;
        kk = knots[blockIdx.y];
    }
    // Everything the thread needs to know about its (knot, instance) is loaded before the first test, and the tests are one
    // predicate: behind a chain of early returns each group of loads waited for the previous one (four global-memory round
    // trips before the first useful instruction; ncu: 4 % of the kernel's stall samples).  All addresses are valid for any
    // i < n, kk < KT.
    const int p = kk / KSTRIDE, k = kk - p * KSTRIDE;
    const int nph = d.n_phases[i], nwb = d.n_wb[i];
    const int mode = d.ph_mode[ph_idx(d, p, i)], N = d.ph_N[ph_idx(d, p, i)];
    const bool enabled = inst_enabled(d, i, mask_mode);
    const bool wb = p < nwb;
    const bool term = (k == N - 1);
    // classify
    int cls;
    if (wb && !term) cls = (mode == 1) ? CLS_BS : (mode == 3) ? CLS_FS : CLS_FL;
    else if (wb && term && mode == 2) cls = CLS_IMPF;
    else if (wb && term && mode == 4) cls = CLS_IMPB;
    else cls = CLS_LIGHT;
    const int slot = d.si[(size_t)SI_ACC_SLOT * d.npad + i];
    const size_t es = d.npad;
    const double vel = d.vel[i], height = d.height[i];
    const double pos = d.pos_ref[(size_t)kk * es + i];
    const int lin_mode = d.si[(size_t)SI_AL_PARTIALS * d.npad + i];  // 2: derivatives of a full forward sweep (AL terms enter Phix/Phixx)
    // the scalars take part in the predicate (comparisons that hold for every valid batch), so their loads complete HERE, together:
    // ptxas otherwise sinks each load to its first use deep in the body, where it costs a round trip of its own
    bool ok = enabled & (p < nph) & (k < N) & (cls == LIN_CLS) & (slot >= 0) & (lin_mode >= 0) & (vel == vel) & (height == height) & (pos == pos);   // (p < nph && k < N: knot_to_phase)
#if LIN_CLS <= 2
    const double dt = d.dt_wb[i];
    const int reb_i = d.si[(size_t)SI_REB_ON * d.npad + i];
    double par[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) { par[q] = d.par[par_idx(d, p, PAR_DELTA_T + q, i)]; ok &= (par[q] == par[q]); }
    ok &= (dt == dt) & (reb_i >= 0);
#endif
    if (!ok) return;

    const int j = i & 31;
    double* __restrict__ rec = d.lin + lin_base(d, i, kk);
    const double* st = d.st_act + act_idx(d, slot, kk, 0, i);

#if LIN_CLS <= 2
    {
        // second (and last) round trip: state, control and outputs of the knot in one burst.  The cost partials go first - they need
        // all 22 values and the register file is still empty - and the Jacobian body then takes its inputs from the same registers.
        double x[14], u[4], y[4];
#pragma unroll
        for (int e = 0; e < 14; ++e) x[e] = sx_ld_in(st + (size_t)e * es);
#pragma unroll
        for (int e = 0; e < 4; ++e) { u[e] = sx_ld_in(st + (size_t)(14 + e) * es); y[e] = sx_ld_in(st + (size_t)(18 + e) * es); }
#if LIN_CLS == 2
#pragma unroll
        for (int e = R_CB(0, 0); e < R_LX; e += 4) st_quad(rec + e, 0.0, 0.0, 0.0, 0.0 ST_QUAD_ARG);   // C = D = 0 in flight
#endif
        wb_cost_partials(c, mode, dt, x, u, y, pos, vel, height, reb_i != 0, par, rec, j ST_QUAD_ARG);
        dyn_par(x, u, rec, j, dt, sx_sm SX_FAR_ARGS ST_QUAD_ARG);
    }
#elif LIN_CLS <= 4
    {
        imp_par(st, es, rec, j, sx_sm SX_FAR_ARGS ST_QUAD_ARG);
        double x[14];
#pragma unroll
        for (int e = 0; e < 14; ++e) x[e] = st[e * es];
        // terminal cost derivatives: Phix = Qf dx, Phixx = Qf  (CostBase.cpp:58-59)
#pragma unroll
        for (int e = 0; e < 14; ++e) {
            REC_AT(R_PHIX + e) = c.wbQf[mode - 1][e] * (x[e] - wb_xref_term(c, mode, e, pos, vel));
            REC_AT(R_PHIXXD + e) = c.wbQf[mode - 1][e];
        }
        double blk[36];
#pragma unroll
        for (int q = 0; q < 36; ++q) blk[q] = 0.0;
        if (lin_mode == 2 && opt.AL_active) {   // only the full forward sweep adds the AL terms
            double h, hx[14], hxx[196];
#pragma unroll
            for (int q = 0; q < 14; ++q) hx[q] = 0.0;
#pragma unroll
            for (int q = 0; q < 196; ++q) hxx[q] = 0.0;
#if LIN_CLS == 3
            tconstr_fl1(x, &h, hx, hxx);
#else
            tconstr_fl2(x, &h, hx, hxx);
#endif
            const double sigma = d.par[par_idx(d, p, PAR_SIGMA, i)], lam = d.par[par_idx(d, p, PAR_LAMBDA, i)];
            const double s2 = sigma * sigma / 2;
#pragma unroll
            for (int a = 1; a <= 6; ++a) {
                const double add = 50 * (s2 * hx[a] * h + lam * hx[a]);
                REC_AT(R_PHIX + a) = c.wbQf[mode - 1][a] * (x[a] - wb_xref_term(c, mode, a, pos, vel)) + add;
            }
#pragma unroll
            for (int b = 1; b <= 6; ++b)
#pragma unroll
                for (int a = 1; a <= 6; ++a)
                    blk[(a - 1) + 6 * (b - 1)] = 50 * (s2 * (hx[a] * hx[b] + h * hxx[a + 14 * b]) + lam * hxx[a + 14 * b]);
        }
#pragma unroll
        for (int q = 0; q < 36; ++q) REC_AT(R_PHIXXB + q) = blk[q];
    }
#else
    if (wb) {  // stance-phase terminal knot: Px = I, no terminal constraint
        double x[14];
#pragma unroll
        for (int e = 0; e < 14; ++e) x[e] = st[e * es];
#pragma unroll
        for (int cc = 0; cc < 12; ++cc)
#pragma unroll
            for (int s = 0; s < 7; ++s) REC_AT(R_AB(s, cc)) = (cc == s + 5) ? 1.0 : 0.0;
#pragma unroll
        for (int e = 0; e < 14; ++e) {
            REC_AT(R_PHIX + e) = c.wbQf[mode - 1][e] * (x[e] - wb_xref_term(c, mode, e, pos, vel));
            REC_AT(R_PHIXXD + e) = c.wbQf[mode - 1][e];
        }
#pragma unroll
        for (int q = 0; q < 36; ++q) REC_AT(R_PHIXXB + q) = 0.0;
    } else if (!term) {
        const double dt = d.dt_fb[i];
        const int off = d.ph_off[ph_idx(d, p, i)];
        double x[6], u[4], x0p[3], fh[4], ct[2], a2[2], b8[8];
#pragma unroll
        for (int e = 0; e < 6; ++e) x[e] = st[e * es];
#pragma unroll
        for (int e = 0; e < 4; ++e) u[e] = st[(14 + e) * es];
        const double* st0 = d.st_act + act_idx(d, slot, off, 0, i);   // phase-initial state (SinglePhase.cpp:150-153)
#pragma unroll
        for (int e = 0; e < 3; ++e) x0p[e] = st0[e * es];
        rb_foothold(c, mode, x0p, dt * (double)N, fh);
        rb_contact(mode, ct);
        fb_dyn_par(x, u, fh, ct, a2, b8);
        const double two_dt = 2 * dt;
        double rlx[6], rlu[4];
#pragma unroll
        for (int e = 0; e < 6; ++e) rlx[e] = (two_dt * c.rbQ[e]) * (x[e] - rb_xref(e, pos, vel, height));
#pragma unroll
        for (int e = 0; e < 4; ++e) rlu[e] = (two_dt * c.rbR[mode - 1][e]) * (u[e] - grf_ref(c, e));
        static_assert(R_RA == 0 && R_RB == 2 && R_RLX == 10 && R_RLU == 16 && RB_STEP_LEN == 20, "five quads = one rigid-body step record");
        st_quad(rec + 0, a2[0] * dt, a2[1] * dt, b8[0] * dt, b8[1] * dt ST_QUAD_ARG);
        st_quad(rec + 4, b8[2] * dt, b8[3] * dt, b8[4] * dt, b8[5] * dt ST_QUAD_ARG);
        st_quad(rec + 8, b8[6] * dt, b8[7] * dt, rlx[0], rlx[1] ST_QUAD_ARG);
        st_quad(rec + 12, rlx[2], rlx[3], rlx[4], rlx[5] ST_QUAD_ARG);
        st_quad(rec + 16, rlu[0], rlu[1], rlu[2], rlu[3] ST_QUAD_ARG);
    } else {
#pragma unroll
        for (int e = 0; e < 6; ++e) REC_AT(R_RPHIX + e) = c.rbQf[e] * (st[e * es] - rb_xref(e, pos, vel, height));
    }
#endif
}

#if LIN_CLS <= 4 && MHPC_RA_STAGE_SLOTS > 0
#define LIN_FAR_KPARAMS , FarScratch far
#define LIN_FAR_BEGIN                                                                                            \
    __shared__ int s_far_slot;                                                                                   \
    const int far_slot = far_acquire(far.mask, &s_far_slot);                                                     \
    double* sx_far = far.buf + (size_t)far_slot * LIN_THREADS + threadIdx.x;                                     \
    const unsigned sx_stage_a = (unsigned)__cvta_generic_to_shared(sx_smem_base + (size_t)MHPC_RA_SMEM_SLOTS * LIN_THREADS + threadIdx.x);
#define LIN_FAR_END far_release(far.mask, far_slot);
#else
#define LIN_FAR_KPARAMS
#define LIN_FAR_BEGIN
#define LIN_FAR_END
#endif
__global__ void __launch_bounds__(LIN_THREADS, LIN_MINBLOCKS) k_lin(Dev d, Consts c, Opt opt, int mask_mode, const int* __restrict__ knots, int n_pairs LIN_FAR_KPARAMS) {
    extern __shared__ __align__(16) double sx_smem_base[];
    volatile double* sx_sm = sx_smem_base + threadIdx.x;
    LIN_FAR_BEGIN
    lin_body(d, c, opt, mask_mode, knots, n_pairs, sx_sm SX_FAR_ARGS);
    LIN_FAR_END
}

#if LIN_CLS <= 4
// inputs in the trajectory layout (element-major, stride es: 14 x then 4 u), one evaluation per thread
__global__ void __launch_bounds__(LIN_THREADS, LIN_MINBLOCKS) k_par_eval(int n, size_t es, const double* __restrict__ st, double* o0, double* o1,
                                                                         double* o2, double* o3 LIN_FAR_KPARAMS) {
    extern __shared__ __align__(16) double sx_smem_base[];
    volatile double* sx_sm = sx_smem_base + threadIdx.x;
    LIN_FAR_BEGIN
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n)
        par_eval_body(st + t, es, o0 + (size_t)t * 196, o1 ? o1 + (size_t)t * 56 : nullptr, o2 ? o2 + (size_t)t * 56 : nullptr,
                      o3 ? o3 + (size_t)t * 16 : nullptr, sx_sm SX_FAR_ARGS);
    LIN_FAR_END
}
#endif

#if LIN_CLS <= 4 && MHPC_RA_STAGE_SLOTS > 0
// one scratch array + slot masks per device for this knot class, allocated at the first launch (a process drives one GPU;
// tests may touch several)
FarScratch far_scratch() {
    static FarScratch per_dev[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    FarScratch& f = per_dev[dev & 63];
    if (!f.buf) {
        if (cudaMalloc(&f.buf, FAR_DOUBLES * sizeof(double)) != cudaSuccess || cudaMalloc(&f.mask, FAR_SM_IDS * sizeof(unsigned)) != cudaSuccess) {
            f.buf = nullptr;
            return f;      // the launch below fails with an invalid pointer and the caller's cudaGetLastError reports it
        }
        cudaMemset(f.mask, 0, FAR_SM_IDS * sizeof(unsigned));
    }
    return f;
}
#define LIN_FAR_LAUNCH_ARG , far_scratch()
#else
#define LIN_FAR_LAUNCH_ARG
#endif

}  // anonymous namespace

#if LIN_CLS == 0
#define LIN_LAUNCH launch_lin_BS
#elif LIN_CLS == 1
#define LIN_LAUNCH launch_lin_FS
#elif LIN_CLS == 2
#define LIN_LAUNCH launch_lin_FL
#elif LIN_CLS == 3
#define LIN_LAUNCH launch_lin_IMPF
#elif LIN_CLS == 4
#define LIN_LAUNCH launch_lin_IMPB
#else
#define LIN_LAUNCH launch_lin_LIGHT
#endif
#if LIN_CLS == 0
#define PAR_EVAL_LAUNCH launch_par_eval_BS
#elif LIN_CLS == 1
#define PAR_EVAL_LAUNCH launch_par_eval_FS
#elif LIN_CLS == 2
#define PAR_EVAL_LAUNCH launch_par_eval_FL
#elif LIN_CLS == 3
#define PAR_EVAL_LAUNCH launch_par_eval_IMPF
#elif LIN_CLS == 4
#define PAR_EVAL_LAUNCH launch_par_eval_IMPB
#endif
#if LIN_CLS <= 4
void PAR_EVAL_LAUNCH(int n, size_t es, const double* st, double* o0, double* o1, double* o2, double* o3, cudaStream_t s) {
    if (LIN_SMEM > 48 * 1024) cudaFuncSetAttribute(k_par_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LIN_SMEM);
    k_par_eval<<<(n + LIN_THREADS - 1) / LIN_THREADS, LIN_THREADS, LIN_SMEM, s>>>(n, es, st, o0, o1, o2, o3 LIN_FAR_LAUNCH_ARG);
}
#endif

void LIN_LAUNCH(const Dev& d, const Consts& c, const Opt& opt, int mask_mode, const int* knots, int n_knots, cudaStream_t s) {
    if (n_knots == 0) return;
    // 128-thread blocks; the Jacobian classes carry MHPC_RA_SMEM_SLOTS shared-memory spill slots per thread
    // (codegen/regalloc.py), sized so that LIN_MINBLOCKS blocks are resident per SM.
    if (LIN_SMEM > 48 * 1024) cudaFuncSetAttribute(k_lin, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LIN_SMEM);
    if (n_knots < 0) {   // (knot, instance) pair list of -n_knots entries
        const int n_pairs = -n_knots;
        k_lin<<<(n_pairs + LIN_THREADS - 1) / LIN_THREADS, LIN_THREADS, LIN_SMEM, s>>>(d, c, opt, mask_mode, knots, n_pairs LIN_FAR_LAUNCH_ARG);
        return;
    }
    dim3 grid((d.n + LIN_THREADS - 1) / LIN_THREADS, n_knots);
    k_lin<<<grid, LIN_THREADS, LIN_SMEM, s>>>(d, c, opt, mask_mode, knots, 0 LIN_FAR_LAUNCH_ARG);
}

}  // namespace mhpc
