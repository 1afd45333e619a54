// TEST INFRASTRUCTURE (oracle) - not part of the product; see oracle/README.md.
//
// CPU restatement of the reference's MHPC / HSDDP solver (NaCl-1374/MHPC_Minimal_Env),
// Eigen-free, linked against the reference's own *unmodified* CasADi-generated C
// (oracle/_ref/libcasadigen_ref_*.so, built by oracle/Makefile from the sources where
// they lie under /root/reference/CasadiGen/source).  Every function cites the reference
// file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
// legs may load this.
//
// PARITY STATUS: the reference ships no tests / golden vectors and
// cannot be built as shipped (Eigen3 is an absent, unpinned system dependency).  The
// restatement is pinned against the reference *sources compiled here* through the
// minimal Eigen stand-in in oracle/eigen_shim (oracle/README.md, tests/test_oracle_vs_ref.py:
// bit-identical on 8 solve cases and 4 receding-horizon update_problem cases).  Real Eigen's kernel rounding is not reproduced.
#pragma once
#include <vector>

#include "eigen_restated.h"
#include "smallmat.h"

namespace orc {

constexpr int N_TIMESTEPS_MAX = 110;  // Common/header/MHPC_CPPTypes.h:9
enum { CALC_DYN_AND_PAR = 0, CALC_PARTIALS_ONLY = 1, CALC_DYNAMICS_ONLY = 2 };  // MHPC_CPPTypes.h:6-8
constexpr double PI = 3.141592653589793238;  // MHPC_CPPTypes.h:18

// ---- Common/header/MHPC_CompoundTypes.h:7-194 ---------------------------------------
template <int n, int m, int p>
struct ModelState { Vec<n> x; Vec<m> u; Vec<p> y; };
template <int n, int m, int p>
struct DynDerivative { Mat<n, n> A; Mat<n, m> B; Mat<p, n> C; Mat<p, m> D; };
template <int n, int m, int p>
struct RCostStruct { double l; Vec<n> lx; Vec<m> lu; Vec<p> ly; Mat<n, n> lxx; Mat<m, n> lux; Mat<m, m> luu; Mat<p, p> lyy; };
template <int n>
struct TCostStruct { double Phi; Vec<n> Phix; Mat<n, n> Phixx; };
template <int n, int m>
struct CostToGoStruct { Vec<n> G; Mat<n, n> H; Vec<m> du; Mat<m, n> K; Vec<n> Qx; Vec<m> Qu; Mat<n, n> Qxx; Mat<m, n> Qux; Mat<m, m> Quu; };
template <int n>
struct TConstrStruct { double h; Vec<n> hx; Mat<n, n> hxx; };
template <int n, int m, int p>
struct IneqConstrStruct { double g; Vec<n> gx; Vec<m> gu; Vec<p> gy; };  // gxx,guu,gyy are identically zero in the reference

// MHPC_CompoundTypes.h:196-212 (same fields, same defaults)
struct Option {
    double alpha = 0.1, gamma = 0.01, update_penalty = 8, update_relax = 0.1, update_regularization = 2,
           update_ReB = 7, max_DDP_iter = 3, max_AL_iter = 2, DDP_thresh = 1e-3, AL_thresh = 1e-3;
    bool AL_active = 1, ReB_active = 1, smooth_active = 0;
};
// MHPC_CompoundTypes.h:214-235
struct ALReBParam {
    bool al_empty = true, reb_empty = true;
    double sigma = 0;
    std::vector<double> lambda, delta, delta_min, eps_ReB;
};

// One record per inner (DDP) iteration: the "bit-exact integers" + per-iteration scalars.
struct IterRecord {
    int iter_AL, iter_DDP, n_bws, n_fit;
    double reg_used, eps_accepted, cost, exp_cost_change, viol, cost_prev;
};

struct Problem {           // what test_main.cpp:16-32 + Gait.h + MHPCLocomotion.cpp:37-39 fix
    Option opt;
    int n_wbphase = 4, n_fbphase = 4;          // MHPCUserParameters
    float dt_wb = .001f, dt_fb = .001f;
    int cmode = 1;
    float vel = 1.5f, height = 0.f;            // USRCMD
    int gait_mode[4] = {1, 2, 3, 4};           // Gait.h:21-26
    float gait_timing[4] = {0.08f, 0.1f, 0.08f, 0.1f};
    double x0[14] = {0.0927, -0.1093, -0.1542, 1.0957, -2.2033, 0.9742, -1.7098,
                     0.9011, 0.2756, 0.7333, 0.0446, 0.0009, 1.3219, 2.7346};
};

class Solver;

}  // namespace orc
