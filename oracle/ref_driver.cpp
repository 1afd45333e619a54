// TEST INFRASTRUCTURE (oracle) - not part of the product.
// Thin C driver around the reference's OWN, UNMODIFIED sources (compiled where they lie under /root/reference against
// oracle/eigen_shim) so that tests can pin the restatement in hsddp_oracle.cpp against the reference itself.
// `private` is opened for this translation unit only, to set a per-instance x0 (hard-coded in MHPCLocomotion.cpp:37-39)
// and to read the trajectory arrays that print_debugInfo dumps (MHPCLocomotion.cpp:293-380).
#include <algorithm>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include <unistd.h>
#define private public
#define protected public
#include "MHPCLocomotion.h"
#undef private
#undef protected

struct RefHandle {
    HSDDP_OPTION<double> option;
    USRCMD usrcmd;
    MHPCUserParameters params;
    Gait gait;
    MHPCLocomotion<double>* loco = nullptr;
};

extern "C" {
void* ref_create(const double* opt10, const int* flags3, int n_wb, int n_fb, float dt_wb, float dt_fb, int cmode, float vel, float height,
                 const int* gait_mode, const float* gait_timing, const double* x0) {
    RefHandle* h = new RefHandle();
    h->option.alpha = opt10[0]; h->option.gamma = opt10[1]; h->option.update_penalty = opt10[2]; h->option.update_relax = opt10[3];
    h->option.update_regularization = opt10[4]; h->option.update_ReB = opt10[5]; h->option.max_DDP_iter = opt10[6];
    h->option.max_AL_iter = opt10[7]; h->option.DDP_thresh = opt10[8]; h->option.AL_thresh = opt10[9];
    h->option.AL_active = flags3[0]; h->option.ReB_active = flags3[1]; h->option.smooth_active = flags3[2];
    h->usrcmd.vel = vel; h->usrcmd.height = height; h->usrcmd.roll = 0; h->usrcmd.pitch = 0; h->usrcmd.yaw = 0;
    h->params.n_wbphase = n_wb; h->params.n_fbphase = n_fb; h->params.dt_wb = dt_wb; h->params.dt_fb = dt_fb; h->params.cmode = cmode;
    h->params.usrcmd = &h->usrcmd;
    for (int i = 0; i < 4; ++i) { h->gait._gait_mode[i] = gait_mode[i]; h->gait._gait_timing[i] = gait_timing[i]; }
    h->loco = new MHPCLocomotion<double>(&h->params, &h->gait, h->option);
    for (int i = 0; i < 14; ++i) h->loco->x0[i] = x0[i];
    // The two places where the reference reads indeterminate values are given the values the oracle documents
    // (SURVEY.md A.3): s[3] of WBCost is an uninitialised stack vector (MHPCCost.cpp:29,42-44,50) -> S(mode 4) = 0;
    // _tconstr_violation is read before its first write (MultiPhaseDDP.cpp:179) -> 0.
    h->loco->wbcostGen->_S[3].setZero();
    h->loco->_tconstr_violation = 0;
    return h;
}
void ref_update_problem(void* p) { ((RefHandle*)p)->loco->update_problem(); }
void ref_set_x0(void* p, const double* x0) {
    MHPCLocomotion<double>* l = ((RefHandle*)p)->loco;
    for (int i = 0; i < 14; ++i) l->x0[i] = x0[i];
    l->set_initial_condition(l->x0);
}
void ref_initialization(void* p) { ((RefHandle*)p)->loco->initialization(); }
void ref_forward_sweep(void* p, double eps) { ((RefHandle*)p)->loco->forward_sweep(eps); }
void ref_update_nominal(void* p) { ((RefHandle*)p)->loco->update_nominal_trajectory(); }
int ref_backward_sweep(void* p, double reg) { return ((RefHandle*)p)->loco->backward_sweep(reg) ? 1 : 0; }
int ref_forward_iteration(void* p) { return ((RefHandle*)p)->loco->forward_iteration(); }
void ref_solve(void* p) { ((RefHandle*)p)->loco->solve_mhpc(); }
void ref_scalars(void* p, double* out3) {
    MHPCLocomotion<double>* l = ((RefHandle*)p)->loco;
    out3[0] = l->_actual_cost; out3[1] = l->_exp_cost_change; out3[2] = l->_tconstr_violation;
}
// MHPCLocomotion::print_debugInfo (MHPCLocomotion.cpp:293-380) writes state.txt, control.txt, gradient.txt, cost.txt into the
// current directory: run it inside `dir`
int ref_print_debugInfo(void* p, const char* dir) {
    char cwd[4096];
    if (!getcwd(cwd, sizeof(cwd)) || chdir(dir) != 0) return -1;
    ((RefHandle*)p)->loco->print_debugInfo();
    return chdir(cwd);
}
int ref_n_phases(void* p) { return ((RefHandle*)p)->loco->_n_phases; }
void ref_schedule(void* p, int* mode, int* N) {
    MHPCLocomotion<double>* l = ((RefHandle*)p)->loco;
    for (int i = 0; i < l->_n_phases; ++i) { mode[i] = l->mode_seq[i]; N[i] = l->N_TIMESTEPS[i]; }
}
// nominal x (which=0), u (1), y (2), K (3), du (4), G (5), H (6), Qx (7), Qu (8), Qxx (9), Qux (10), Quu (11) of one phase, knot-major;
// returns doubles per knot
int ref_get(void* p, int phase, int which, double* out) {
    MHPCLocomotion<double>* l = ((RefHandle*)p)->loco;
    const int N = l->N_TIMESTEPS[phase];
    if (phase < l->n_wbphase) {
        const int pidx = l->pidx_WB[phase];   // buffers rotate in update_problem (MHPCLocomotion.cpp:112-121)
        auto* ms = l->ms_nom_WB[pidx]; auto* ctg = l->CTG_WB[pidx];
        const int dims[12] = {14, 4, 4, 56, 4, 14, 196, 14, 4, 196, 56, 16};
        for (int k = 0; k < N && out; ++k) {
            const double* src = which == 0 ? ms[k].x.data() : which == 1 ? ms[k].u.data() : which == 2 ? ms[k].y.data()
                              : which == 3 ? ctg[k].K.data() : which == 4 ? ctg[k].du.data() : which == 5 ? ctg[k].G.data()
                              : which == 6 ? ctg[k].H.data() : which == 7 ? ctg[k].Qx.data() : which == 8 ? ctg[k].Qu.data()
                              : which == 9 ? ctg[k].Qxx.data() : which == 10 ? ctg[k].Qux.data() : ctg[k].Quu.data();
            std::memcpy(out + (size_t)k * dims[which], src, dims[which] * sizeof(double));
        }
        return dims[which];
    }
    const int q = l->pidx_FB[phase - l->n_wbphase];
    auto* ms = l->ms_nom_FB[q]; auto* ctg = l->CTG_FB[q];
    const int dims[12] = {6, 4, 4, 24, 4, 6, 36, 6, 4, 36, 24, 16};
    for (int k = 0; k < N && out; ++k) {
        const double* src = which == 0 ? ms[k].x.data() : which == 1 ? ms[k].u.data() : which == 2 ? ms[k].y.data()
                          : which == 3 ? ctg[k].K.data() : which == 4 ? ctg[k].du.data() : which == 5 ? ctg[k].G.data()
                          : which == 6 ? ctg[k].H.data() : which == 7 ? ctg[k].Qx.data() : which == 8 ? ctg[k].Qu.data()
                          : which == 9 ? ctg[k].Qxx.data() : which == 10 ? ctg[k].Qux.data() : ctg[k].Quu.data();
        std::memcpy(out + (size_t)k * dims[which], src, dims[which] * sizeof(double));
    }
    return dims[which];
}
}
