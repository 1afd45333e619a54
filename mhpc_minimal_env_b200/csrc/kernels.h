// Launch entry points of the CUDA kernels (one translation unit per kernel family so the large
// generated bodies compile in parallel).
#pragma once
#include "common.cuh"

namespace mhpc {

struct Consts;

struct EpsList {  // line-search step sizes eps_t = alpha^t formed by repeated multiplication (MultiPhaseDDP.cpp:132-148)
    double eps[MAX_TRIALS];
    int n;
};

// which instances a launch touches
enum { MASK_ALL = 0, MASK_RUNNING, MASK_DDP, MASK_LS, MASK_LIN, MASK_SNAP };

__device__ inline bool inst_enabled(const Dev& d, int i, int mask_mode) {
    switch (mask_mode) {
        case MASK_ALL: return true;
        case MASK_RUNNING: return d.si[(size_t)SI_STATUS * d.npad + i] == STATUS_RUNNING;
        case MASK_DDP:
        case MASK_LS: return d.si[(size_t)SI_STATUS * d.npad + i] == STATUS_RUNNING && d.si[(size_t)SI_DDP_ACTIVE * d.npad + i] != 0;
        case MASK_LIN: return d.si[(size_t)SI_STATUS * d.npad + i] == STATUS_RUNNING && d.si[(size_t)SI_AL_PARTIALS * d.npad + i] != 0;
        case MASK_SNAP: return d.si[(size_t)SI_SNAP * d.npad + i] != 0;   // a mask frozen by launch_snapshot_mask
    }
    return false;
}

// setup (k_misc.cu)
void launch_setup(const Dev& d, const Consts& c, cudaStream_t s);       // references, AL/ReB parameters, zeroed state
void launch_warmstart(const Dev& d, const Consts& c, cudaStream_t s);   // boundingPDControl.cpp on the device -> act slot 0
void launch_copy_nominal(const Dev& d, int mask_mode, cudaStream_t s);  // act[acc_slot] -> nom
void launch_snapshot_mask(const Dev& d, int mask_mode, cudaStream_t s); // SI_SNAP <- inst_enabled(mask_mode), for work that runs on a side stream
void launch_accept_slot0(const Dev& d, int mask_mode, int lin_mode, cudaStream_t s);
void launch_solve_begin(const Dev& d, cudaStream_t s);
void launch_al_begin(const Dev& d, const Opt& o, cudaStream_t s);
void launch_ddp_begin(const Dev& d, cudaStream_t s);
void launch_ls_begin(const Dev& d, int mask_mode, cudaStream_t s);
void launch_select(const Dev& d, const Opt& o, int mask_mode, int trial0, int n_slots, int last_chunk, const EpsList& e, cudaStream_t s);
void launch_ddp_end(const Dev& d, const Opt& o, int ddp_iter_host, cudaStream_t s);
void launch_al_end(const Dev& d, const Opt& o, const Consts& c, int mask_mode, int use_state_penalty, cudaStream_t s);
void launch_count_active(const Dev& d, cudaStream_t s);
void launch_reset_option_state(const Dev& d, const Opt& o, cudaStream_t s);
void launch_reb_from_option(const Dev& d, cudaStream_t s);
void launch_rotate(const Dev& d, const double* src, double* dst, int elems, int use_acc_slot, cudaStream_t s);
void launch_rotate_gain(const Dev& d, const double* src, double* dst, cudaStream_t s);   // the same for the gain records (dst <- rotated src)
void launch_spread_last_knot(const Dev& d, cudaStream_t s);
void launch_setup_refs_params(const Dev& d, const Consts& c, cudaStream_t s);
// rollout (k_rollout.cu)
void launch_rollout(const Dev& d, const Consts& c, const Opt& o, int mask_mode, int n_slots, int trial0, const EpsList& e, cudaStream_t s);
// forward_sweep(0) of an outer iteration after the first: costs of the unchanged trajectory under the new AL / ReB parameters (3 launches)
void launch_recost(const Dev& d, const Consts& c, const Opt& o, int mask_mode, int lin_mode, cudaStream_t s);
// linearisation (k_lin.cu compiled once per knot class)
void launch_linearize(const Dev& d, const Consts& c, const Opt& o, int mask_mode, const int* const* knot_lists, const int* n_knots, cudaStream_t s);
void launch_lin_BS(const Dev&, const Consts&, const Opt&, int, const int*, int, cudaStream_t);
void launch_lin_FS(const Dev&, const Consts&, const Opt&, int, const int*, int, cudaStream_t);
void launch_lin_FL(const Dev&, const Consts&, const Opt&, int, const int*, int, cudaStream_t);
void launch_lin_IMPF(const Dev&, const Consts&, const Opt&, int, const int*, int, cudaStream_t);
void launch_lin_IMPB(const Dev&, const Consts&, const Opt&, int, const int*, int, cudaStream_t);
void launch_lin_LIGHT(const Dev&, const Consts&, const Opt&, int, const int*, int, cudaStream_t);
// the same Jacobian bodies evaluated standalone (raw CasADi outputs, dense column-major); st = [18][es] element-major inputs
void launch_par_eval_BS(int n, size_t es, const double* st, double* o0, double* o1, double* o2, double* o3, cudaStream_t s);
void launch_par_eval_FS(int n, size_t es, const double* st, double* o0, double* o1, double* o2, double* o3, cudaStream_t s);
void launch_par_eval_FL(int n, size_t es, const double* st, double* o0, double* o1, double* o2, double* o3, cudaStream_t s);
void launch_par_eval_IMPF(int n, size_t es, const double* st, double* o0, double* o1, double* o2, double* o3, cudaStream_t s);
void launch_par_eval_IMPB(int n, size_t es, const double* st, double* o0, double* o1, double* o2, double* o3, cudaStream_t s);
// backward sweep (k_bws.cu)
// max_attempts > 0 (with the ladder scheduler allocated): stop walking the ladder in the kernel after that many rungs
void launch_bws(const Dev& d, const Consts& c, const Opt& o, int mask_mode, int single_attempt, int max_attempts, cudaStream_t s);
void launch_bws_tasks(const Dev& d, const Consts& c, const Opt& o, int mask_mode, int n_tasks, cudaStream_t s);   // n_tasks <= 0: unknown on the host
int bws_resident_lane_groups(int device);
// the same sweep with sixteen lanes per instance (k_bws16.cu)
void launch_bws16(const Dev& d, const Consts& c, const Opt& o, int mask_mode, int single_attempt, int max_attempts, cudaStream_t s);
void launch_bws16_tasks(const Dev& d, const Consts& c, const Opt& o, int mask_mode, int n_tasks, cudaStream_t s);
int bws16_resident_lane_groups(int device);
// ... and on the FP64 tensor cores, one warp per instance (k_bwsT.cu)
void launch_bwsT(const Dev& d, const Consts& c, const Opt& o, int mask_mode, int single_attempt, int max_attempts, cudaStream_t s);
void launch_bwsT_tasks(const Dev& d, const Consts& c, const Opt& o, int mask_mode, int n_tasks, cudaStream_t s);
int bwsT_resident_warps(int device);
void launch_lad_plan(const Dev& d, const Opt& o, int mask_mode, int last_round, int smax, cudaStream_t s);
void launch_bws_replay(const Dev& d, const Consts& c, const Opt& o, cudaStream_t s);   // task 1 of a two-entry task list, writing every knot's H, G, Q* to d.ctg_dump
void launch_lad_finish(const Dev& d, const Opt& o, int n_unres, cudaStream_t s);
// standalone evaluation of one generated function over a batch (Jacobian parity tests; k_eval.cu)
int launch_eval(int fn, int n, const double* in0, const double* in1, const double* in2, const double* in3,
                double* o0, double* o1, double* o2, double* o3, cudaStream_t s);

}  // namespace mhpc
