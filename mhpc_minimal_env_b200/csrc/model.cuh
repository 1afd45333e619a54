// Device-side models: the CasADi-derived dynamics / impact / constraint functions (bodies
// generated at build time into csrc/gen/*.inc by mhpc_minimal_env_b200/codegen from the same SX
// expressions the reference compiles on the CPU), the explicit-Euler wrappers of
// Dynamics/source/PlanarQuadruped.cpp and PlanarFloatingBase.cpp, the foothold planner
// (Dynamics/header/FootholdPlan.h), the quadratic costs (HSDDPSolver/source/CostBase.cpp with the
// weights of Controller/MHPCLocomotion/MHPCCost.cpp), the inequality rows of
// Controller/MHPCLocomotion/MHPCConstraints.cpp and the relaxed barrier of
// HSDDPSolver/source/SinglePhase.cpp:298-317.
#pragma once
#include "common.cuh"

namespace mhpc {

// Constants uploaded once by the host (weights are formed with the same IEEE operations the
// reference uses, e.g. 0.01 * q, 100 * qf, PI / 25) and passed to kernels by value.
struct Consts {
    double wbQ[14], wbR[4][4], wbS[4][4], wbQf[4][14];   // MHPCCost.cpp:24-53 (s[3] := 0)
    double rbQ[6], rbR[4][4], rbQf[6];                   // MHPCCost.cpp:56-75
    double xterm[4][14];                                 // ReferenceGen.cpp:42-49 ([0] and [7] filled per instance)
    double qbias[4];                                     // ReferenceGen.cpp:51
    double GRF;                                          // ReferenceGen.cpp:41
    double bjoint[8], btorque;                           // MHPCConstraints.cpp:49-54
    double mu;                                           // MHPCConstraints.h:33
    double fh_vel, fh_ground;                            // MHPCLocomotion.cpp:25
    double hip_x, hip_len, knee_len;                     // PlanarQuadruped.cpp:120-138
    double qjoint_nom[4], kp[4], legext_nom, kspring;    // boundingPDControl.cpp:6-13
};

#define SX_SINCOS(a, s, c) sincos((a), &(s), &(c))
// Reciprocal of a factorisation pivot (codegen emits SX_RCP for every 1/d node).  `1.0 / d` is a ~127-cycle dependent
// sequence on this GPU (range checks, slow-path branch; scripts/micro/fp64_lat.cu) and the dynamics have nine of them in
// a row; the pivots are ordinary well-scaled numbers, so the hardware seed (~20 bits) plus two Newton steps (error
// <= 2 ulp, ~60 cycles) is enough.  MHPC_EXACT_RCP restores the IEEE division.
#ifdef MHPC_EXACT_RCP
#define SX_RCP(x) (1.0 / (x))
#else
static __device__ __forceinline__ double sx_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
#define SX_RCP(x) sx_rcp(x)
#endif

// ---- whole-body continuous dynamics  xdot(14), y(4) --------------------------------------
#define SX_IN(i, j) ((i) == 0 ? x[j] : u[j])
#define SX_OUT(k, nz, row, col, v) do { if ((k) == 0) xdot[row] = (v); else y[row] = (v); } while (0)
static __device__ __forceinline__ void dyn_bs(const double* __restrict__ x, const double* __restrict__ u, double* __restrict__ xdot, double* __restrict__ y) {
#include "gen/Dyn_BS.inc"
}
static __device__ __forceinline__ void dyn_fs(const double* __restrict__ x, const double* __restrict__ u, double* __restrict__ xdot, double* __restrict__ y) {
#include "gen/Dyn_FS.inc"
}
static __device__ __forceinline__ void dyn_fl(const double* __restrict__ x, const double* __restrict__ u, double* __restrict__ xdot, double* __restrict__ y) {
#include "gen/Dyn_FL.inc"
}
#undef SX_IN
#undef SX_OUT
// ---- impact maps  x+(14), impulse y(4) ---------------------------------------------------
#define SX_IN(i, j) (x[j])
#define SX_OUT(k, nz, row, col, v) do { if ((k) == 0) xp[row] = (v); else y[row] = (v); } while (0)
static __device__ __forceinline__ void imp_f(const double* __restrict__ x, double* __restrict__ xp, double* __restrict__ y) {
#include "gen/Imp_F.inc"
}
static __device__ __forceinline__ void imp_b(const double* __restrict__ x, double* __restrict__ xp, double* __restrict__ y) {
#include "gen/Imp_B.inc"
}
#undef SX_OUT
// ---- touchdown constraint  h, hx(1x14, 4 nz), hxx(14x14, 9 nz) --------------------------
#define SX_OUT(k, nz, row, col, v) do { if ((k) == 0) *h = (v); else if ((k) == 1) hx[col] = (v); else hxx[(row) + 14 * (col)] = (v); } while (0)
static __device__ __forceinline__ void tconstr_fl1(const double* x, double* h, double* hx, double* hxx) {
#include "gen/WB_FL1_terminal_constr.inc"
}
static __device__ __forceinline__ void tconstr_fl2(const double* x, double* h, double* hx, double* hxx) {
#include "gen/WB_FL2_terminal_constr.inc"
}
#undef SX_OUT
// value only (line-search rollouts)
#define SX_OUT(k, nz, row, col, v) do { if ((k) == 0) *h = (v); } while (0)
static __device__ __forceinline__ void tconstr_fl1_h(const double* x, double* h) {
#include "gen/WB_FL1_terminal_constr.inc"
}
static __device__ __forceinline__ void tconstr_fl2_h(const double* x, double* h) {
#include "gen/WB_FL2_terminal_constr.inc"
}
#undef SX_OUT
// ---- foot Jacobians (warm start only): J (2x7) --------------------------------------------
#define SX_OUT(k, nz, row, col, v) do { if ((k) == 0) J[(row) + 2 * (col)] = (v); } while (0)
static __device__ __forceinline__ void jacob_f(const double* x, double* J) {
#include "gen/Jacob_F.inc"
}
static __device__ __forceinline__ void jacob_b(const double* x, double* J) {
#include "gen/Jacob_B.inc"
}
#undef SX_OUT
#undef SX_IN
// ---- link-point Jacobians J, Jdot (2x7) of a point p (2) on link 0..4: dormant in the reference (CasadiGen.h:22-26) ----
#define SX_IN(i, j) ((i) == 0 ? x[j] : pt[j])
#define SX_OUT(k, nz, row, col, v) do { if ((k) == 0) J[(row) + 2 * (col)] = (v); else Jd[(row) + 2 * (col)] = (v); } while (0)
static __device__ __forceinline__ void link_jacobian(int link, const double* x, const double* pt, double* J, double* Jd) {
    if (link == 0) {
#include "gen/Link0Jacobian.inc"
    } else if (link == 1) {
#include "gen/Link1Jacobian.inc"
    } else if (link == 2) {
#include "gen/Link2Jacobian.inc"
    } else if (link == 3) {
#include "gen/Link3Jacobian.inc"
    } else {
#include "gen/Link4Jacobian.inc"
    }
}
#undef SX_OUT
#undef SX_IN
// ---- rigid-body model: xdot(6) from x, u, foothold(4), contact(2) -------------------------
#define SX_IN(i, j) ((i) == 0 ? x[j] : (i) == 1 ? u[j] : (i) == 2 ? fh[j] : ct[j])
#define SX_OUT(k, nz, row, col, v) do { xdot[row] = (v); } while (0)
static __device__ __forceinline__ void fb_dyn(const double* x, const double* u, const double* fh, const double* ct, double* xdot) {
#include "gen/FBDynamics.inc"
}
#undef SX_OUT
// Ac (5 nz, 3 of them the constant 1) and Bc (8 nz, CCS order)
#define SX_OUT(k, nz, row, col, v) do { if ((k) == 0) { if ((nz) < 2) a2[nz] = (v); } else b8[nz] = (v); } while (0)
static __device__ __forceinline__ void fb_dyn_par(const double* x, const double* u, const double* fh, const double* ct, double* a2, double* b8) {
#include "gen/FBDynamics_par.inc"
}
#undef SX_OUT
#undef SX_IN

// PlanarQuadruped.cpp:8-27: x+ = x + xdot * dt, y = GRF
static __device__ __forceinline__ void wb_step(int mode, double dt, const double* x, const double* u, double* xn, double* y) {
    double xdot[14];
    if (mode == 1) dyn_bs(x, u, xdot, y);
    else if (mode == 3) dyn_fs(x, u, xdot, y);
    else dyn_fl(x, u, xdot, y);
#pragma unroll
    for (int i = 0; i < 14; ++i) xn[i] = x[i] + xdot[i] * dt;
}
// PlanarQuadruped.cpp:58-78
static __device__ __forceinline__ void wb_reset(int mode, const double* x, double* xp) {
    double y[4];
    if (mode == 2) imp_f(x, xp, y);
    else if (mode == 4) imp_b(x, xp, y);
    else {
#pragma unroll
        for (int i = 0; i < 14; ++i) xp[i] = x[i];
    }
}
// PlanarFloatingBase.cpp:9-23
static __device__ __forceinline__ void rb_contact(int mode, double* ct) {
    ct[0] = (mode == 3) ? 1.0 : 0.0;
    ct[1] = (mode == 1) ? 1.0 : 0.0;
}
// FootholdPlan.h:27-50 with PlanarQuadruped.cpp:142-186 at zero joint angles: the hip position is the
// translation column of T_body * T_hip, i.e. x -/+ hip_x * cos(pitch) (R_y(-pitch)(0,0) = cos(-pitch)).
static __device__ __forceinline__ void rb_foothold(const Consts& c, int mode, const double* x0, double stance_time, double* fh) {
    fh[0] = fh[1] = fh[2] = fh[3] = 0.0;
    if (mode == 1) {
        const double hx = cos(-x0[2]) * (-c.hip_x) + x0[0];
        fh[2] = hx + c.fh_vel * stance_time / 2;
        fh[3] = c.fh_ground;
    } else if (mode == 3) {
        const double hx = cos(-x0[2]) * c.hip_x + x0[0];
        fh[0] = hx + c.fh_vel * stance_time / 2;
        fh[1] = c.fh_ground;
    }
}
// PlanarFloatingBase.cpp:7-32
static __device__ __forceinline__ void rb_step(int mode, double dt, const double* x, const double* u, const double* fh, double* xn) {
    double ct[2], xdot[6];
    rb_contact(mode, ct);
    fb_dyn(x, u, fh, ct, xdot);
#pragma unroll
    for (int i = 0; i < 6; ++i) xn[i] = x[i] + xdot[i] * dt;
}

// ---- references (ReferenceGen.h:53-91) ----------------------------------------------------
static __device__ __forceinline__ double wb_xref(const Consts& c, int i, double pos, double vel, double height) {
    return i == 0 ? pos : i == 1 ? height : i == 2 ? 0.0 : i < 7 ? c.qbias[i - 3] : i == 7 ? vel : 0.0;
}
static __device__ __forceinline__ double wb_xref_term(const Consts& c, int mode, int i, double pos, double vel) {
    return i == 0 ? pos : i == 7 ? vel : c.xterm[mode - 1][i];
}
static __device__ __forceinline__ double rb_xref(int i, double pos, double vel, double height) {
    return i == 0 ? pos : i == 1 ? height : i == 3 ? vel : 0.0;
}
static __device__ __forceinline__ double grf_ref(const Consts& c, int j) { return (j & 1) ? c.GRF : 0.0; }

// ---- relaxed log barrier (SinglePhase.cpp:298-317, polynomial order k = 2) ----------------
static __device__ __forceinline__ void reb_eval(double g, double delta, double& b, double& bz, double& bzz) {
    if (g > delta) {
        b = -log(g);
        bz = -1.0 / g;
        bzz = 1.0 / (g * g);
    } else {
        const double r = (g - 2 * delta) / (1 * delta);
        b = 0.5 * (r * r - 1) - log(delta);
        bz = r / delta;
        bzz = 1.0;  // pow(r, k-2) with k = 2: NOT divided by delta^2
    }
}

// The derivative pair alone (SinglePhase.cpp:303-316) for the cost partials: bz = -1/g exactly as the reference computes it, and
// bzz = (1/g)^2 from the same quotient instead of a second division 1/(g*g) (<= 1.5 ulp apart; a DDIV is ~20 instructions and
// the 19 barrier rows of a knot were 10 % of the Jacobian kernels' instructions).
static __device__ __forceinline__ void reb_eval_partials(double g, double delta, double& bz, double& bzz) {
    if (g > delta) {
        const double rg = 1.0 / g;
        bz = -rg;
        bzz = rg * rg;
    } else {
        const double r = (g - 2 * delta) / (1 * delta);
        bz = r / delta;
        bzz = 1.0;  // pow(r, k-2) with k = 2: NOT divided by delta^2
    }
}

// Running cost of a WB knot (CostBase.cpp:4-16) plus, when `reb` is set, the barrier value
// terms of SinglePhase.cpp:236-237 for the 8 torque, 8 joint and (stance) 3 GRF rows.
static __device__ __forceinline__ double wb_running_cost(const Consts& c, int mode, double dt, const double* x, const double* u, const double* y,
                                         double pos, double vel, double height, bool reb, const double* par /*delta T,J,G eps T,J,G*/) {
    double l = 0.0;
#pragma unroll
    for (int i = 0; i < 14; ++i) { const double d = x[i] - wb_xref(c, i, pos, vel, height); const double t = (d * c.wbQ[i]) * d; l = (i == 0) ? t : l + t; }
    double lu = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const double d = u[i] - 0.0; const double t = (d * c.wbR[mode - 1][i]) * d; lu = (i == 0) ? t : lu + t; }
    l += lu;
    double ly = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const double d = y[i] - grf_ref(c, i); const double t = (d * c.wbS[mode - 1][i]) * d; ly = (i == 0) ? t : ly + t; }
    l += ly;
    l *= dt;
    if (reb) {
        double b, bz, bzz;
        const double dT = par[0], dJ = par[1], dG = par[2], eT = par[3], eJ = par[4], eG = par[5];
        if (eT != 0.0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { const double g = (i < 4 ? -u[i] : u[i - 4]) + c.btorque; reb_eval(g, dT, b, bz, bzz); l += eT * b * dt; }
        }
        if (eJ != 0.0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { const double g = (i < 4 ? -x[3 + i] : x[3 + i - 4]) + c.bjoint[i]; reb_eval(g, dJ, b, bz, bzz); l += eJ * b * dt; }
        }
        if ((mode == 1 || mode == 3) && eG != 0.0) {
            const double fx = (mode == 1) ? y[2] : y[0], fz = (mode == 1) ? y[3] : y[1];
            const double g3[3] = {fz, -fx + c.mu * fz, fx + c.mu * fz};
#pragma unroll
            for (int i = 0; i < 3; ++i) { reb_eval(g3[i], dG, b, bz, bzz); l += eG * b * dt; }
        }
    }
    return l;
}
static __device__ __forceinline__ double rb_running_cost(const Consts& c, int mode, double dt, const double* x, const double* u, double pos, double vel, double height) {
    double l = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) { const double d = x[i] - rb_xref(i, pos, vel, height); const double t = (d * c.rbQ[i]) * d; l = (i == 0) ? t : l + t; }
    double lu = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const double d = u[i] - grf_ref(c, i); const double t = (d * c.rbR[mode - 1][i]) * d; lu = (i == 0) ? t : lu + t; }
    l += lu;
    l += 0.0;  // S = 0 for the rigid-body cost
    l *= dt;
    return l;
}

// Terminal cost of a whole-body phase (SinglePhase.cpp:101-113 / 138-143): quadratic term, touchdown constraint h of the
// flight modes and, when the augmented Lagrangian is active, its penalty terms.  Shared by the rollout and the cost-only
// pass (k_recost_knots) so that both evaluate the same expression.
static __device__ __forceinline__ double wb_terminal_cost(const Consts& c, int mode, const double* x, double pos, double vel, bool al_active,
                                                          double sigma, double lam, double& h) {
    double Phi = 0.0;
#pragma unroll
    for (int e = 0; e < 14; ++e) {
        const double dd = x[e] - wb_xref_term(c, mode, e, pos, vel);
        const double t = (dd * c.wbQf[mode - 1][e]) * dd;
        Phi = (e == 0) ? t : Phi + t;
    }
    Phi *= 0.5;
    h = 0.0;
    if (mode == 2) tconstr_fl1_h(x, &h);
    else if (mode == 4) tconstr_fl2_h(x, &h);
    if ((mode == 2 || mode == 4) && al_active) {
        const double a = sigma * h / 2;
        Phi += 50 * (a * a + lam * h);
    }
    return Phi;
}
static __device__ __forceinline__ double rb_terminal_cost(const Consts& c, const double* x, double pos, double vel, double height) {
    double Phi = 0.0;
#pragma unroll
    for (int e = 0; e < 6; ++e) {
        const double dd = x[e] - rb_xref(e, pos, vel, height);
        const double t = (dd * c.rbQf[e]) * dd;
        Phi = (e == 0) ? t : Phi + t;
    }
    Phi *= 0.5;
    return Phi;
}

// phase p and in-phase index k of knot slot kk (phase p owns slots [p*KSTRIDE, p*KSTRIDE + N_p))
static __device__ __forceinline__ bool knot_to_phase(const Dev& d, int i, int kk, int& p, int& k) {
    p = kk / KSTRIDE;
    k = kk - p * KSTRIDE;
    return p < d.n_phases[i] && k < d.ph_N[ph_idx(d, p, i)];
}

}  // namespace mhpc
