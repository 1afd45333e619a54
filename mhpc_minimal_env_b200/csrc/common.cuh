// Device data model of the batched HSDDP path (see DESIGN.md "Data layout in HBM").
//
// Trajectory-type data is structure-of-arrays over instances so that the
// thread-per-instance kernels (rollout, linearise, line-search select) coalesce:
//     st[(knot * ST + e) * npad + inst]          e: 0..13 x | 14..17 u | 18..21 y
// Per-knot linearisation records are array-of-structures, one contiguous 1408 B record per
// (instance, knot), so that the backward sweep (a half-warp per instance) streams its instance's
// records with 16-byte cp.async into warp-private shared memory, one knot ahead of the math:
//     lin[(inst * KT + knot) * REC + e]
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mhpc {

constexpr int NXW = 14, NXR = 6, NU = 4, NY = 4;
constexpr int ST = 22;            // x(14) u(4) y(4) per knot
constexpr int KDIM = 56;          // K is 4x14 column-major (RB uses the first 24 = 4x6)
// Gain record of a knot: everything the backward sweep produces there, contiguous per (instance, knot) so that the warp that
// owns the instance writes whole sectors and the rollout streams K, du with 16-byte copies
constexpr int GREC = 74, GR_K = 0, GR_DU = 56, GR_G = 60;   // K 4x14 | du 4 | G 14  (592 B)
// The rest of a knot's CostToGoStruct (MHPC_CompoundTypes.h:100-144), written only by the replay sweep behind
// mhpc_batch_get_field (k_bws<1, 1>): column-major with leading dimension 14 (4 for Qux, Quu)
constexpr int CTGREC = 496, CG_H = 0, CG_G = 196, CG_QX = 210, CG_QU = 224, CG_QXX = 228, CG_QUX = 424, CG_QUU = 480;
constexpr int REC = 176;          // doubles per knot record
constexpr int MAXP = 16;          // phases per instance (reference default 8)
constexpr int KSTRIDE = 110;      // knot slots per phase = N_TIMESTEPS_MAX (MHPC_CPPTypes.h:9): phase p owns knots [p*110, p*110 + N_p),
                                  // exactly like the reference's per-phase arrays, so the receding-horizon buffer rotation
                                  // (MHPCLocomotion.cpp:107-158) moves whole blocks including their stale tails
constexpr int MAX_TRIALS = 48;    // line-search step sizes: eps = alpha^t > 1e-10 (reference default alpha = 0.1: 10; alpha = 0.6: 46)
constexpr int MAX_W = 16;         // trial rollouts evaluated side by side (slots of st_act)
constexpr int NPAR = 9;           // per-phase AL/ReB parameters

// ---- record element indices -------------------------------------------------------------
// WB step knot: A = I + dt*Ac, B = dt*Bc in the reference (PlanarQuadruped.cpp:51-52); only the
// dense blocks are stored: Ab = A[7:14, 2:14], Bb = B[7:14, :], Cb/Db = the two stance rows of C/D.
__host__ __device__ constexpr int R_AB(int s, int c) { return c * 7 + s; }            // 0..83
__host__ __device__ constexpr int R_BB(int s, int c) { return 84 + c * 7 + s; }       // 84..111
__host__ __device__ constexpr int R_CB(int t, int c) { return 112 + c * 2 + t; }      // 112..135
__host__ __device__ constexpr int R_DB(int t, int c) { return 136 + c * 2 + t; }      // 136..143
constexpr int R_LX = 144;    // 14
constexpr int R_LU = 158;    // 4
constexpr int R_LY = 162;    // 2 (stance rows of ly)
constexpr int R_LXXQ = 164;  // 4 (lxx diagonal at the joint positions 3..6)
constexpr int R_LUU = 168;   // 4 (luu diagonal)
constexpr int R_LYY = 172;   // 3 (stance 2x2 block of lyy: 00, 01(=10), 11)
constexpr int WB_STEP_LEN = 175;
// WB terminal knot: Pb = Px[7:14, 2:14] of the impact Jacobian (identity pattern for modes 1,3)
constexpr int R_PHIX = 84;     // 14
constexpr int R_PHIXXD = 98;   // 14 diagonal of Qf
constexpr int R_PHIXXB = 112;  // 6x6 AL block on state indices 1..6, column-major
constexpr int WB_TERM_LEN = 148;
// RB step knot
constexpr int R_RA = 0;    // A(5,0), A(5,1)
constexpr int R_RB = 2;    // 8 non-zeros of B in CCS order (3,0)(5,0)(4,1)(5,1)(3,2)(5,2)(4,3)(5,3)
constexpr int R_RLX = 10;  // 6
constexpr int R_RLU = 16;  // 4
constexpr int RB_STEP_LEN = 20;
constexpr int R_RPHIX = 0;  // 6
constexpr int RB_TERM_LEN = 6;

// per-phase parameter slots (AL_REB_PARAMETER, MHPC_CompoundTypes.h:214-235; the reference's
// delta / eps_ReB vectors are constant within each constraint group: torque, joint, GRF)
enum { PAR_SIGMA = 0, PAR_LAMBDA, PAR_DELTA_T, PAR_DELTA_J, PAR_DELTA_G, PAR_EPS_T, PAR_EPS_J, PAR_EPS_G, PAR_H };

// per-instance solver scalars
// SD_OPT_UPD / SI_OPT_REB are the per-instance copies of `_option.update_penalty` / `_option.ReB_active`, which
// MultiPhaseDDP::solve overwrites and never restores (MultiPhaseDDP.cpp:160-161,178-183,273-277): they survive
// initialization() and later solves exactly like the members of a re-used MHPCLocomotion object; *_BASE are solve()'s locals.
enum { SD_COST = 0, SD_COST_PREV, SD_DV, SD_VIOL, SD_REG, SD_REG_USED, SD_UPD_PENALTY, SD_EPS_ACC, SD_OPT_UPD, SD_UPD_BASE, SD_N };
enum { SI_STATUS = 0, SI_ITER_AL, SI_ITER_DDP, SI_DDP_ACTIVE, SI_REB_ON, SI_N_BWS, SI_N_FIT, SI_ACC_SLOT, SI_LS_DONE,
       SI_AL_PARTIALS, SI_N_REC, SI_BWS_OK, SI_OPT_REB, SI_REB_BASE, SI_SNAP, SI_N };
enum { STATUS_RUNNING = 0, STATUS_CONVERGED = 1, STATUS_MAXITER = 2, STATUS_REG_ABORT = 3 };
constexpr int REC_I = 4, REC_D = 6;  // per-iteration record: ints {iter_AL, iter_DDP, n_bws, n_fit}, doubles {reg, eps, cost, dV, viol, cost_prev}

struct Opt {  // HSDDP_OPTION<double>, MHPC_CompoundTypes.h:196-212
    double alpha, gamma, update_penalty, update_relax, update_regularization, update_ReB, max_DDP_iter, max_AL_iter,
        DDP_thresh, AL_thresh;
    int AL_active, ReB_active, smooth_active;
    double quu_shift;   // pow(0.1, 9)  (SinglePhase.cpp:202)
};

struct Dev {
    int n, npad, nblk;  // instances, padded to 32, blocks of 32
    int P;              // phase slots per instance
    int KT;             // knot slots per instance
    int W;              // line-search trial slots
    int max_rec;        // iteration records per instance
    // schedule
    int *n_phases, *n_wb, *n_knots;     // [npad]; n_knots = sum of N_p (knots actually used)
    int *ph_mode, *ph_N, *ph_off;       // [P][npad]; ph_off[p] = p * KSTRIDE
    double *dt_wb, *dt_fb, *vel, *height;  // [npad]
    double* x0;                         // [14][npad]
    double* pos_ref;                    // [KT][npad]
    // trajectories
    double* st_nom;  // [KT][ST][npad]
    double* st_act;  // [W][KT][ST][npad]
    double* gain;    // [npad][KT][GREC] gain records: K | du | G
    double* ctg_dump;   // [KT][CTGREC] of ONE instance, or null: filled by the replay sweep (capi.cu: replay_cost_to_go)
    double* lin;     // AoSoA records
    // parameters and solver state
    double* par;     // [P][NPAR][npad]
    double* tr_cost; // [W][npad]
    double* tr_viol2;// [W][npad]
    double* tr_h;    // [W][P][npad]
    double* knot_cost;  // [KT][npad] per-knot cost terms of the cost-only pass (k_recost_knots)
    double* sd;      // [SD_N][npad]
    int* si;         // [SI_N][npad]
    int* rec_i;      // [max_rec][REC_I][npad]
    double* rec_d;   // [max_rec][REC_D][npad]
    int* n_active;   // [1] device counter for optional host early exit
    // optional diagnostic (mhpc_batch_enable_bws_trace): per (iteration record, attempt) the backward step at which the
    // attempt of the backward sweep stopped: it + 1 if Quu was not positive definite at backward step `it`, -1 = passed
    int* bws_trace;  // [max_rec][BWS_TRACE_ATT][npad] or null
    // regularisation-ladder scheduler of the backward sweep (k_bws.cu: k_lad_plan / k_lad_resolve / k_lad_copy); null = off
    int lad_cap;          // lane groups / warps of the sweep kernel that are resident at once on this GPU
    int lad_tcap;         // length of the task arrays
    int lad_target;       // tasks a round should reach when the sweep kernel is throughput-bound (0: fill lad_cap only)
    int* lad_cnt;         // [LAD_CNT_N]
    int* lad_unres;       // [npad] the instance has no passing rung yet in this sweep
    int* lad_rung;        // [npad] next rung (0-based attempt index of this sweep) to test
    int* lad_t_inst;      // [cap] task -> instance slot (-1: idle)
    int* lad_t_rung;      // [cap] in: rung of the task; out: the rung that passed, -1 if none
    double* lad_t_reg;    // [cap] in: regularisation of the task's rung; out: of the rung that passed
    double* lad_t_dv;     // [cap] out: expected cost change of the passing sweep
    int* lad_u_inst;      // [cap] compact list of the unresolved instances of this round
    int* lad_u_copy;      // [cap] per unresolved instance: scratch column that holds its result (-1: none / already in place)
    double* lad_gain;     // [npad][KT][GREC] outputs of the speculative tasks: gain records, first index = scratch column
};
enum { LAD_N_TASKS = 0, LAD_N_UNRES, LAD_STRIDE, LAD_S, LAD_CNT_N };
// scratch column of speculative task t (t % S != 0) when every instance has S tasks: the first task of an instance writes in place
__host__ __device__ inline int lad_scratch_col(int t, int S) { return (t / S) * (S - 1) + (t % S) - 1; }
constexpr int BWS_TRACE_ATT = 24;   // the ladder 0, 1e-3, 2e-3, ... exceeds 1000 at the 22nd attempt

__host__ __device__ inline size_t st_idx(const Dev& d, int kk, int e, int i) { return ((size_t)kk * ST + e) * d.npad + i; }
// col: instance slot (gain) or scratch column (lad_gain).  GAIN_KNOT_MAJOR: [KT][npad][GREC] (the records of one knot side by side)
// instead of [npad][KT][GREC]
#ifndef GAIN_KNOT_MAJOR
#define GAIN_KNOT_MAJOR 0
#endif
__host__ __device__ inline size_t gain_idx(const Dev& d, int kk, size_t col) { return GAIN_KNOT_MAJOR ? ((size_t)kk * d.npad + col) * GREC : (col * d.KT + kk) * GREC; }
__host__ __device__ inline size_t gain_knot_stride(const Dev& d) { return GAIN_KNOT_MAJOR ? (size_t)d.npad * GREC : (size_t)GREC; }
__host__ __device__ inline size_t act_idx(const Dev& d, int w, int kk, int e, int i) {
    return (((size_t)w * d.KT + kk) * ST + e) * d.npad + i;
}
__host__ __device__ inline size_t lin_base(const Dev& d, int i, int kk) { return ((size_t)i * d.KT + kk) * REC; }
__host__ __device__ inline size_t par_idx(const Dev& d, int p, int q, int i) { return ((size_t)p * NPAR + q) * d.npad + i; }
__host__ __device__ inline size_t ph_idx(const Dev& d, int p, int i) { return (size_t)p * d.npad + i; }

// knot classes for the linearisation launches
enum { CLS_BS = 0, CLS_FS, CLS_FL, CLS_IMPF, CLS_IMPB, CLS_LIGHT, CLS_N };

// reference constants ---------------------------------------------------------------------
#define MHPC_PI 3.141592653589793238  /* MHPC_CPPTypes.h:18 */

}  // namespace mhpc
