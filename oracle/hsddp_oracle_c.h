/* TEST INFRASTRUCTURE (oracle) - not part of the product.  C interface of the CPU
 * restatement (oracle/hsddp_oracle.cpp) for ctypes: tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs only. */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_problem {
    /* HSDDP_OPTION<double>  (Common/header/MHPC_CompoundTypes.h:196-212) */
    double alpha, gamma, update_penalty, update_relax, update_regularization, update_ReB,
           max_DDP_iter, max_AL_iter, DDP_thresh, AL_thresh;
    int AL_active, ReB_active, smooth_active;
    /* MHPCUserParameters + USRCMD (:237-251), Gait (Common/header/Gait.h:21-26), x0 (MHPCLocomotion.cpp:37-39) */
    int n_wbphase, n_fbphase;
    float dt_wb, dt_fb;
    int cmode;
    float vel, height;
    int gait_mode[4];
    float gait_timing[4];
    double x0[14];
} orc_problem;

enum {
    ORC_F_X_ACT = 0, ORC_F_U_ACT, ORC_F_Y_ACT, ORC_F_X_NOM, ORC_F_U_NOM, ORC_F_Y_NOM,
    ORC_F_X_REF, ORC_F_U_REF, ORC_F_Y_REF, ORC_F_A, ORC_F_B, ORC_F_C, ORC_F_D,
    ORC_F_LX, ORC_F_LU, ORC_F_LY, ORC_F_LXX, ORC_F_LUU, ORC_F_LYY,
    ORC_F_K, ORC_F_DU, ORC_F_G, ORC_F_H, ORC_F_PHIX, ORC_F_PHIXX,
    ORC_F_QX, ORC_F_QU, ORC_F_QXX, ORC_F_QUX, ORC_F_QUU   /* the rest of CostToGoStruct (MHPC_CompoundTypes.h:100-144) */
};

void* orc_create(const orc_problem* p);
void orc_destroy(void* h);
void orc_set_faithful_alloc(int on);
void orc_initialization(void* h);
void orc_forward_sweep(void* h, double eps);
void orc_forward_sweep_dynamics_only(void* h, double eps);
void orc_forward_sweep_partials_only(void* h);
int orc_backward_sweep(void* h, double reg);
int orc_forward_iteration(void* h);
void orc_update_nominal_trajectory(void* h);
void orc_update_AL_ReB_param(void* h);
void orc_update_problem(void* h);
void orc_set_x0(void* h, const double* x0_14);
void orc_set_ReB_active(void* h, int on);
void orc_set_update_penalty(void* h, double v);
int orc_solve(void* h);
int orc_n_phases(void* h);
void orc_get_schedule(void* h, int* mode, int* N, int* is_wb);
void orc_get_scalars(void* h, double* out4);
int orc_get_field(void* h, int phase, int field, double* out);
void orc_get_phase_params(void* h, int phase, double* out9);
void orc_get_Px(void* h, int phase, double* out196);
int orc_n_records(void* h);
void orc_get_record(void* h, int i, int* ints4, double* dbl6);
int orc_eval(const char* name, const double* a0, const double* a1, const double* a2, const double* a3,
             double* o0, double* o1, double* o2, double* o3);
#ifdef __cplusplus
}
#endif
