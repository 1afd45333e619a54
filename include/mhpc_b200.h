/* mhpc_b200.h - C ABI of the batched, B200-native HSDDP hot path of Model-Hierarchy Predictive Control.
 *
 * The reference (NaCl-1374/MHPC_Minimal_Env) has no FFI layer: its surface is a set of C++ templates
 * instantiated for double and driven from one host thread (test_main.cpp:16-35).  This header is the
 * boundary a maintainer binds instead: plain C, caller-owned host buffers, library-owned device
 * buffers, int return codes (0 = ok), no exceptions, one host thread per handle.  Every entry point
 * names the reference interface it replaces (file:line under the reference tree); the semantics of one
 * instance through this ABI are those of one `MHPCLocomotion<double>` object.
 *
 * A batch holds N independent MHPC problems ("instances").  All arithmetic is FP64 on the device.
 */
#ifndef MHPC_B200_H
#define MHPC_B200_H
#ifdef __cplusplus
extern "C" {
#endif

/* HSDDP_OPTION<double>  -- Common/header/MHPC_CompoundTypes.h:196-212 (same fields, same order, same defaults;
 * the three bools are ints here). */
typedef struct mhpc_hsddp_option {
    double alpha;                 /* 0.1   line-search contraction            */
    double gamma;                 /* 0.01  Armijo slope factor                */
    double update_penalty;        /* 8     AL penalty growth                  */
    double update_relax;          /* 0.1   ReB relaxation shrink              */
    double update_regularization; /* 2     regularisation growth on failure   */
    double update_ReB;            /* 7     ReB weight growth                  */
    double max_DDP_iter;          /* 3     inner iterations (a double in the reference) */
    double max_AL_iter;           /* 2     outer iterations                   */
    double DDP_thresh;            /* 1e-3                                      */
    double AL_thresh;             /* 1e-3                                      */
    int AL_active;                /* 1 */
    int ReB_active;               /* 1 */
    int smooth_active;            /* 0 (dormant in the reference: SinglePhase.cpp:252-254) */
} mhpc_hsddp_option;

/* One problem instance = MHPCUserParameters + USRCMD (MHPC_CompoundTypes.h:237-251), the Gait tables
 * (Common/header/Gait.h:21-26) and the initial state the reference hard-codes in
 * MHPCLocomotion.cpp:37-39 (here per instance: the "batched entry point"). */
typedef struct mhpc_problem {
    int n_wbphase;        /* 4 */
    int n_fbphase;        /* 4 */
    float dt_wb;          /* .001 */
    float dt_fb;          /* .001 */
    int cmode;            /* 1 */
    float groundH;        /* -0.404; never read by the reference, kept for layout */
    float vel, height, roll, pitch, yaw;   /* USRCMD */
    int gait_mode[4];     /* 1,2,3,4 */
    float gait_timing[4]; /* .08,.1,.08,.1 */
    double x0[14];
} mhpc_problem;

/* per-instance result of a solve */
typedef struct mhpc_result {
    double cost;             /* MultiPhaseDDP::_actual_cost       */
    double exp_cost_change;  /* MultiPhaseDDP::_exp_cost_change   */
    double viol;             /* MultiPhaseDDP::_tconstr_violation */
    int status;              /* 1 converged (AL_thresh), 2 max_AL_iter reached, 3 "Regularization term exceeds maximum value" (MultiPhaseDDP.cpp:221-226) */
    int iter_AL;             /* outer iterations executed */
    int n_records;           /* inner iterations executed (records available) */
    int pad;
} mhpc_result;

/* one record per executed inner (DDP) iteration: the integers are the "bit-exact" quantities */
typedef struct mhpc_iter_record {
    int iter_AL, iter_DDP, n_bws, n_fit;   /* TIME_PER_ITERATION::{DDP_iter,n_bws,n_fit}, MHPC_CompoundTypes.h:253-262 */
    double reg, eps, cost, exp_cost_change, viol, cost_prev;
} mhpc_iter_record;

typedef struct mhpc_batch mhpc_batch;

/* trajectory / derivative fields readable per instance and phase (matrices column-major, like Eigen) */
enum mhpc_field {
    MHPC_F_X_ACT = 0, MHPC_F_U_ACT, MHPC_F_Y_ACT,   /* ModelState of the accepted / last rollout  */
    MHPC_F_X_NOM, MHPC_F_U_NOM, MHPC_F_Y_NOM,       /* nominal trajectory                         */
    MHPC_F_X_REF, MHPC_F_U_REF, MHPC_F_Y_REF,       /* tracking reference (ReferenceGen.h:53-91)  */
    MHPC_F_A, MHPC_F_B, MHPC_F_C, MHPC_F_D,         /* DynDerivative, dense                       */
    MHPC_F_LX, MHPC_F_LU, MHPC_F_LY, MHPC_F_LXX, MHPC_F_LUU, MHPC_F_LYY,   /* RCostStruct derivatives */
    MHPC_F_K, MHPC_F_DU, MHPC_F_G,                  /* CostToGoStruct K, du, G                    */
    MHPC_F_H,                                       /* CostToGoStruct H (recomputed on request, see below) */
    MHPC_F_PHIX, MHPC_F_PHIXX,                      /* TCostStruct derivatives                    */
    MHPC_F_PX,                                      /* impact Jacobian used by impact_aware_step  */
    MHPC_F_QX, MHPC_F_QU, MHPC_F_QXX, MHPC_F_QUX, MHPC_F_QUU   /* the Q-function of CostToGoStruct (MHPC_CompoundTypes.h:100-144), as it is
                                                       left by valuefunction_update: Qxx regularised and symmetrised, Quu regularised.
                                                       H and the Q-function are not kept by the solver (the sweep holds them on chip):
                                                       the first request for an instance re-runs its last accepted backward sweep on the
                                                       device (same knot records and regularisation; no solver state changes) and keeps
                                                       the result until any other entry point is called.  Valid after a solve or a
                                                       successful backward_sweep; an error if that sweep did not pass. */
};

/* ---- lifetime --------------------------------------------------------------------------------- */
void mhpc_default_option(mhpc_hsddp_option* o);          /* struct defaults of MHPC_CompoundTypes.h:196-212 */
void mhpc_default_problem(mhpc_problem* p);              /* MHPC_CompoundTypes.h:242-251, Gait.h:21-26, MHPCLocomotion.cpp:37-39, test_main.cpp:23-28 */
/* MHPCLocomotion<double>::MHPCLocomotion (MHPCLocomotion.cpp:8-43), for n instances on CUDA device `device`.
 * ls_width = number of line-search step sizes evaluated concurrently per launch (<= 16; 0 -> 8).  The outcome does not
 * depend on it (first acceptable step size in order, MultiPhaseDDP.cpp:132-148); 8 suits small batches, 4 is faster from
 * a few thousand instances on (the trial rollouts then compete for the FP64 pipe; bench.py uses 4). */
int mhpc_batch_create(int n_instances, const mhpc_hsddp_option* opt, int device, int ls_width, mhpc_batch** out);
void mhpc_batch_destroy(mhpc_batch* b);                  /* MHPCLocomotion::~MHPCLocomotion (:383-470) */
const char* mhpc_last_error(void);
/* sizes / field offsets (declaration order) of the structs above as this library was compiled, so that a binding in
 * another language can verify its mirror: which = 0 option, 1 problem, 2 result, 3 iter_record; returns the field count */
int mhpc_sizeof_option(void);
int mhpc_sizeof_problem(void);
int mhpc_sizeof_result(void);
int mhpc_sizeof_iter_record(void);
int mhpc_sizeof_timing(void);
int mhpc_struct_offsets(int which, int* out, int capacity);
int mhpc_batch_set_option(mhpc_batch* b, const mhpc_hsddp_option* opt);   /* MultiPhaseDDP::set_option, MultiPhaseDDP.h:22 */
/* Use the caller's CUDA stream (e.g. torch's current stream) for all work of this handle; NULL = own stream.
   (A solve forks two short pieces of work onto handle-private side streams and joins them back with events; ordering
   against this one stream is enough.) */
int mhpc_batch_set_stream(mhpc_batch* b, void* cuda_stream);

/* ---- problem definition ------------------------------------------------------------------------- */
/* n_instances problems (host array).  Replaces the constructor arguments + set_initial_condition
 * (MultiPhaseDDP.h:24-25) + build_problem's schedule (MHPCLocomotion.cpp:63-90). */
int mhpc_batch_set_problems(mhpc_batch* b, const mhpc_problem* problems);
/* MHPCLocomotion::initialization (MHPCLocomotion.cpp:47-53): memory_reset, build_problem (references,
 * AL/ReB parameters), warmstart (boundingPDControl.cpp) - all on the device. */
int mhpc_batch_initialization(mhpc_batch* b);
/* MultiPhaseDDP::set_initial_condition (MultiPhaseDDP.h:24-25) for every instance; x0 = [n][14] host doubles. */
int mhpc_batch_set_initial_condition(mhpc_batch* b, const double* x0);
/* MHPCLocomotion::update_problem (MHPCLocomotion.cpp:107-158): the receding-horizon shift - phase buffers of each
 * model group rotate by one, the gait advances one mode, every phase is re-initialised, the reference is regenerated
 * from the current x0.  (Never called by the reference's main, but part of its class; parity is checked against it.)
 * The derivative records (the reference also rotates dynpar / rcost / tcost) are NOT rotated: every solve and every
 * forward_sweep / forward_sweep_partials_only rewrites them before a backward sweep reads them; calling
 * mhpc_batch_backward_sweep directly after update_problem without one of those is not supported. */
int mhpc_batch_update_problem(mhpc_batch* b);

/* ---- MultiPhaseDDP<double> entry points, batched (all instances in lock step) -------------------- */
int mhpc_batch_forward_sweep(mhpc_batch* b, double eps);                 /* MultiPhaseDDP.cpp:28-48   */
int mhpc_batch_forward_sweep_dynamics_only(mhpc_batch* b, double eps);   /* MultiPhaseDDP.cpp:56-76   */
int mhpc_batch_forward_sweep_partials_only(mhpc_batch* b);               /* MultiPhaseDDP.cpp:84-91 ("linearization") */
/* MultiPhaseDDP.cpp:100-127; success[i] = 1/0 per instance (may be NULL). */
int mhpc_batch_backward_sweep(mhpc_batch* b, double regularization, int* success);
/* MultiPhaseDDP.cpp:130-151 with the step sizes evaluated as parallel rollouts; trials[i] = the
 * reference's return value (1-based index of the accepted step size, n_trials+1 if none). */
int mhpc_batch_forward_iteration(mhpc_batch* b, int* trials);
int mhpc_batch_update_nominal_trajectory(mhpc_batch* b);                 /* MultiPhaseDDP.cpp:389-399 */
int mhpc_batch_update_AL_ReB_param(mhpc_batch* b);                       /* MultiPhaseDDP.cpp:402-413 */
/* MultiPhaseDDP::solve (MultiPhaseDDP.cpp:154-289) for every instance; per-instance control flow
 * (regularisation retries, accepted step size, early convergence) follows the reference exactly.
 * sync_early_exit != 0 lets the host poll an "any instance still running" counter once per outer iteration. */
int mhpc_batch_solve(mhpc_batch* b, int sync_early_exit);
/* MHPCLocomotion::solve_mhpc end to end from HOST buffers: upload problems, initialization, solve,
 * download results (+ optional execution bag, MHPCLocomotion.cpp:178-194: nominal x,u,y and K,du of the knots of the
 * first min(exec_phases, n_wbphase) whole-body phases - the reference copies 2 - up to exec_knots knots per instance,
 * rows beyond an instance's execution horizon are zero; pass NULL / 0 to skip). */
int mhpc_batch_solve_host(mhpc_batch* b, const mhpc_problem* problems, mhpc_result* results,
                          int exec_phases, double* exec_state /* [n][exec_knots][22] */, double* exec_gain /* [n][exec_knots][60] */,
                          int exec_knots);
int mhpc_batch_synchronize(mhpc_batch* b);

/* ---- results -------------------------------------------------------------------------------------- */
int mhpc_batch_get_results(mhpc_batch* b, mhpc_result* results /* [n] */);
int mhpc_batch_get_records(mhpc_batch* b, int instance, mhpc_iter_record* out, int max_records);   /* returns count */
int mhpc_batch_get_scalars(mhpc_batch* b, double* cost, double* exp_cost_change, double* viol);    /* each [n] or NULL */
int mhpc_batch_n_phases(mhpc_batch* b, int instance);
int mhpc_batch_get_schedule(mhpc_batch* b, int instance, int* mode, int* n_knots, int* is_wb);     /* mode_seq / N_TIMESTEPS, MHPCLocomotion.cpp:65-85 */
/* per-knot field of one phase of one instance; returns doubles per knot (or total for PHIX/PHIXX/PX), <0 on error;
 * out may be NULL to query the size. */
int mhpc_batch_get_field(mhpc_batch* b, int instance, int phase, int field, double* out);
/* AL/ReB parameters of one phase: {sigma, lambda, delta_torque, delta_joint, delta_grf, eps_torque, eps_joint, eps_grf, h} */
int mhpc_batch_get_phase_params(mhpc_batch* b, int instance, int phase, double* out9);
/* device-resident result table [6][npad] doubles {cost, dV, viol, status, iter_AL, n_records} written into a
 * caller-provided DEVICE buffer (for an NCCL gather across ranks); returns npad. */
int mhpc_batch_results_to_device(mhpc_batch* b, double* dev_out, int capacity_doubles);

/* ---- standalone evaluation of the generated functions on the device (Jacobian parity) ------------
 * Replaces the reference's CasADi C ABI `int F(const double** arg, double** res, long long* iw, double* w, int mem)`
 * (CasadiGen/header/Dyn_BS.h:15-30 and siblings) together with the adaptor casadi_interface (CasadiGen.cpp:4-79), batched:
 *   "Dyn_BS" "Dyn_FS" "Dyn_FL"              (x[14], u[4])                    -> xdot[14], y[4]
 *   "Dyn_BS_par" "Dyn_FS_par" "Dyn_FL_par"  (x[14], u[4])                    -> Ac[14x14], Bc[14x4], C[4x14], D[4x4]
 *   "Imp_F" "Imp_B"                         (x[14])                          -> x+[14], impulse[4]
 *   "Imp_F_par" "Imp_B_par"                 (x[14])                          -> Px[14x14]
 *   "FBDynamics" / "FBDynamics_par"         (x[6], u[4], foothold[4], c[2])  -> xdot[6] / Ac[6x6], Bc[6x4]
 *   "WB_FL1_terminal_constr" / "WB_FL2_..." (x[14])                          -> h[1], hx[14], hxx[14x14]
 *   "Jacob_F" "Jacob_B"                     (x[14])                          -> J[2x7]
 *   "Link0Jacobian" .. "Link4Jacobian"      (x[14], p[2])                    -> J[2x7], Jdot[2x7]   (dormant in the reference:
 *                                            included by CasadiGen.h:22-26, never called)
 * inputs/outputs are host arrays of n evaluations, evaluation-major ([n][size]); outputs dense column-major with
 * structural zeros written as 0 (the adaptor's scatter into pre-zeroed Eigen storage). */
int mhpc_eval_function(const char* name, int n, int device, const double* in0, const double* in1, const double* in2,
                       const double* in3, double* out0, double* out1, double* out2, double* out3);

/* Control-smoothness penalty (HSDDP_OPTION::smooth_active, AL_REB_PARAMETER::eps_smooth = 1; MHPC_CompoundTypes.h:211,224): dormant in
 * the reference - update_running_cost_with_smooth is an empty body that nothing calls (SinglePhase.cpp:252-254).  Accepted, stored,
 * without effect on any result, exactly as there. */
int mhpc_batch_set_eps_smooth(mhpc_batch* b, double eps_smooth);
double mhpc_batch_get_eps_smooth(mhpc_batch* b);

/* ---- timing of the last solve (TIME_PER_ITERATION buckets, MHPC_CompoundTypes.h:253-262), ms, CUDA events */
typedef struct mhpc_timing { float total_ms, rollout_ms, linearize_ms, bws_ms, linesearch_ms, other_ms; int n_launches; } mhpc_timing;
int mhpc_batch_enable_timing(mhpc_batch* b, int on);
int mhpc_batch_get_timing(mhpc_batch* b, mhpc_timing* t);
/* the same per inner iteration: TIME_PER_ITERATION::{time_bws, time_fit, time_partial} (MHPC_CompoundTypes.h:253-262, filled by
 * MultiPhaseDDP.cpp:206-208,228-235,242-253,259-267 under TIME_BENCHMARK) for the whole batch, ms; the full forward sweep of an
 * outer iteration (MultiPhaseDDP.cpp:184-192) is reported with its first inner iteration.  Returns the number of records. */
typedef struct mhpc_iter_timing { int iter_AL, iter_DDP; float time_fwd_ms, time_bws_ms, time_fit_ms, time_partial_ms, total_ms; } mhpc_iter_timing;
int mhpc_batch_get_iteration_timing(mhpc_batch* b, mhpc_iter_timing* out, int max_records);
/* diagnostic: for every inner-iteration record r and every attempt a of its backward sweep (the regularisation ladder of
 * MultiPhaseDDP.cpp:209-227) the backward step at which the attempt stopped: it + 1 when Quu failed the positive-definite
 * test at backward step `it` (0 = the last knot's predecessor), -1 when the attempt passed, 0 when it was not made.
 * out = [n][rows] ints, rows = the return value (= record slots x 24 attempts). */
int mhpc_batch_enable_bws_trace(mhpc_batch* b, int on);
int mhpc_batch_get_bws_trace(mhpc_batch* b, int* out);
/* run `reps` linearisation launches back to back and return the average ms per launch (CUDA events) */
int mhpc_batch_time_kernel(mhpc_batch* b, int which /*0 linearize group, 1 backward sweep, 2 one rollout chunk, 10+c linearisation knot class c (0 BS,1 FS,2 FL,3 IMPF,4 IMPB,5 light)*/, int reps, float* ms_per_launch);
/* FP64 FMA peak probe on this device (TFLOP/s), used as the roofline denominator for the linearisation */
int mhpc_measure_fp64_peak(int device, double* tflops);
long long mhpc_batch_device_bytes(mhpc_batch* b);

#ifdef __cplusplus
}
#endif
#endif /* MHPC_B200_H */
