// Host side of the C ABI (include/mhpc_b200.h): problem setup, device memory, and the launch
// schedule that reproduces MultiPhaseDDP::solve (MultiPhaseDDP.cpp:154-289) for a whole batch with
// per-instance control flow kept on the device.  No CPU arithmetic of the hot path lives here.
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <string>
#include <vector>

#include "../../include/mhpc_b200.h"
#include "kernels.h"
#include "model.cuh"

namespace mhpc {
int eval_fn_id(const char* name);
double measure_fp64_peak_tflops(int device);
}
using namespace mhpc;

extern "C" int mhpc_batch_enable_bws_trace(mhpc_batch* b, int on);
static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return -1; }
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_));          \
    } while (0)

enum { TB_ROLLOUT = 0, TB_LIN, TB_BWS, TB_LS, TB_OTHER, TB_N };

struct mhpc_batch {
    int n = 0, device = 0, W = 8;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    Opt opt{};
    Consts c{};
    Dev d{};
    EpsList eps{};
    bool have_problems = false;
    std::vector<mhpc_problem> problems;      // as given by the caller (instance order)
    // Instances are independent, so their place in the batch is an internal matter: a batch with mixed schedules is
    // stored sorted by schedule (phase split, modes, knot counts), which keeps the instances that share a warp in the
    // sweep and the rollouts (4 per warp) on nearly the same knot classes.  slot_of[instance] / user_of[slot]; every
    // entry point speaks instance order.  Uniform batches are not permuted.
    bool permuted = false;
    std::vector<int> slot_of, user_of;
    int* d_user_of = nullptr;
    // side stream for the two impact-Jacobian launches of a linearisation (32 blocks each, one long dependent chain per
    // thread): they run underneath the stance/flight kernels instead of after them
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // second side stream: the nominal <- accepted-rollout copy (HBM-bound) runs underneath the linearisation and the
    // latency-bound backward sweep; it is joined before the next kernel that reads the nominal or rewrites the rollouts
    cudaStream_t side2 = nullptr;
    cudaEvent_t ev_cfork = nullptr, ev_cjoin = nullptr;
    bool copy_pending = false;
    std::vector<int> h_nph, h_nwb, h_nk, h_mode, h_N, h_off;   // host copy of the schedule ([P][npad] for the last three)
    std::vector<void*> allocs;
    long long bytes = 0;
    int* knot_list[CLS_N] = {nullptr};
    int knot_cnt[CLS_N] = {0};
    // non-uniform batches (instances with different schedules): per knot class an explicit list of (knot, instance)
    // pairs, so that every lane of a linearisation warp evaluates the SAME generated function (no class divergence)
    unsigned* pair_list[CLS_N] = {nullptr};
    int pair_cnt[CLS_N] = {0};
    bool uniform = true;
    size_t pair_cap[CLS_N] = {0};
    int alloc_KT = 0, alloc_P = 0;
    // timing
    bool timing = false;
    struct Ev { cudaEvent_t a, b; int bucket; int iter; };   // iter: inner iteration the stage belongs to (al * 1000 + it), -1: none
    int cur_iter = -1;
    std::vector<mhpc_iter_timing> iter_times;
    std::vector<Ev> evs;
    mhpc_timing last{};
    int launches = 0;
    void* stage = nullptr;
    size_t stage_bytes = 0;
    bool keep_dead_partials = false;
    // replay cache of mhpc_batch_get_field (H / Q fields): instance slot whose CostToGoStruct is in d.ctg_dump, and the call
    // generation it was computed in (every entry point other than get_field advances the generation)
    long long gen = 0, ctg_gen = -1;
    int ctg_inst = -1;
    bool recost = true;   // outer iterations after the first: cost-only pass instead of the step-size-0 rollout (MHPC_RECOST=0: rollout)
    void* rot_tmp = nullptr;
    size_t rot_bytes = 0;
    // regularisation-ladder scheduler of the backward sweep (k_bws.cu): rounds after the first launch, rungs the first
    // launch walks in the kernel, widest fan of rungs per instance.  MHPC_LADDER=0 puts the whole ladder back into one launch.
    bool ladder = true;
    int bws_kind = 2;                // backward sweep mapping: 2 = tensor cores, warp per instance (k_bwsT.cu, default); MHPC_BWS=8 / 16: eight / sixteen lanes per instance, DFMA (k_bws.cu, k_bws16.cu)
    int lad_rounds = 6, lad_first = 2, lad_smax = 16;
    bool lad_target_on = false;      // MHPC_LADDER_TARGET=0: rungs per round from the resident capacity only
    bool lad_sync = true;            // read the count of unresolved instances back after each round (one short host wait) instead of
                                     // launching a fixed number of rounds blind; MHPC_LADDER_SYNC=0 keeps the solve free of host waits
    int* h_unres = nullptr;          // pinned copy of lad_cnt
    double eps_smooth = 1.0;         // AL_REB_PARAMETER::eps_smooth (MHPC_CompoundTypes.h:224): stored, never read - as in the reference
    bool want_bws_trace = false;     // diagnostic: where each attempt of the backward sweep stopped
    int* bws_trace_buf = nullptr;
    size_t bws_trace_cap = 0;
};

static void fill_opt(Opt& o, const mhpc_hsddp_option& s) {
    o.alpha = s.alpha; o.gamma = s.gamma; o.update_penalty = s.update_penalty; o.update_relax = s.update_relax;
    o.update_regularization = s.update_regularization; o.update_ReB = s.update_ReB; o.max_DDP_iter = s.max_DDP_iter;
    o.max_AL_iter = s.max_AL_iter; o.DDP_thresh = s.DDP_thresh; o.AL_thresh = s.AL_thresh;
    o.AL_active = s.AL_active; o.ReB_active = s.ReB_active; o.smooth_active = s.smooth_active;
    o.quu_shift = std::pow(0.1, 9);   // SinglePhase.cpp:202
}
static void fill_eps(EpsList& e, double alpha) {   // MultiPhaseDDP.cpp:132-148
    double eps = 1;
    e.n = 0;
    while (eps > std::pow(0.1, 10) && e.n < MAX_TRIALS) { e.eps[e.n++] = eps; eps *= alpha; }
    for (int i = e.n; i < MAX_TRIALS; ++i) e.eps[i] = 0;
}
// Options the device path cannot follow the reference on are refused instead of silently truncated: the reference's
// line search runs `while (eps > 1e-10) ... eps *= alpha` (MultiPhaseDDP.cpp:132-148), i.e. forever for alpha >= 1 and
// more than MAX_TRIALS step sizes for alpha above ~0.62.
static int check_option(const mhpc_hsddp_option& o) {
    if (!(o.alpha > 0.0 && o.alpha < 1.0)) return fail("option: alpha must be in (0, 1) (the reference's line search does not terminate otherwise, MultiPhaseDDP.cpp:132-148)");
    int n = 0;
    for (double eps = 1; eps > std::pow(0.1, 10); eps *= o.alpha) if (++n > MAX_TRIALS) break;
    if (n > MAX_TRIALS) return fail("option: alpha gives more than " + std::to_string((int)MAX_TRIALS) + " line-search step sizes above 1e-10; not supported");
    if (!(o.max_AL_iter >= 0 && o.max_DDP_iter >= 0) || o.max_AL_iter * o.max_DDP_iter > 1e6) return fail("option: bad iteration limits");
    return 0;
}
static void fill_consts(Consts& c) {
    const double PI = MHPC_PI;
    const double q[14] = {0, 10, 5, 4, 4, 4, 4, 2, 1, .01, 6, 6, 6, 6};                       // MHPCCost.cpp:30
    const double qf[4][14] = {{0, 20, 8, 3, 3, 3, 3, 3, 2, 0.01, 5, 5, 0.01, 0.01},           // :32-35
                              {0, 20, 8, 3, 3, 3, 3, 3, 2, 0.01, 5, 5, 5, 5},
                              {0, 20, 8, 3, 3, 3, 3, 3, 2, 0.01, 0.01, 0.01, 5, 5},
                              {0, 20, 8, 3, 3, 3, 3, 3, 2, 0.01, 5, 5, 5, 5}};
    const double r[4][4] = {{5, 5, 1, 1}, {1, 1, 1, 1}, {1, 1, 5, 5}, {1, 1, 1, 1}};          // :37-40
    const double s[4][4] = {{0, 0, 0.3, 0.3}, {0, 0, 0, 0}, {0.15, 0.15, 0, 0}, {0, 0, 0, 0}};// :42-44 (s[3] undefined in the reference -> 0)
    for (int i = 0; i < 14; ++i) c.wbQ[i] = 0.01 * q[i];
    for (int m = 0; m < 4; ++m) {
        for (int i = 0; i < 4; ++i) { c.wbR[m][i] = 0.5 * r[m][i]; c.wbS[m][i] = s[m][i]; }
        for (int i = 0; i < 14; ++i) c.wbQf[m][i] = 100 * qf[m][i];
    }
    const double rq[6] = {0, 10, 5, 2, 1, 0.01}, rqf[6] = {1, 20, 8, 3, 1, 0.01};             // :63-64
    const double rr[4][4] = {{0, 0, 0.01, 0.01}, {0, 0, 0, 0}, {0.01, 0.01, 0, 0}, {0, 0, 0, 0}};
    for (int i = 0; i < 6; ++i) { c.rbQ[i] = 0.01 * rq[i]; c.rbQf[i] = 100 * rqf[i]; }
    for (int m = 0; m < 4; ++m) for (int i = 0; i < 4; ++i) c.rbR[m][i] = rr[m][i];
    const double xt[4][14] = {{0, -0.1432, -PI / 25, 0.35 * PI, -0.65 * PI, 0.35 * PI, -0.6 * PI, 0, 1, 0, 0, 0, 0, 0},   // ReferenceGen.cpp:42-49
                              {0, -0.1418, PI / 35, 0.2 * PI, -0.58 * PI, 0.25 * PI, -0.7 * PI, 0, -1, 0, 0, 0, 0, 0},
                              {0, -0.1325, -PI / 40, 0.33 * PI, -0.48 * PI, 0.33 * PI, -0.75 * PI, 0, 1, 0, 0, 0, 0, 0},
                              {0, -0.1490, -PI / 25, 0.35 * PI, -0.7 * PI, 0.25 * PI, -0.60 * PI, 0, -1, 0, 0, 0, 0, 0}};
    std::memcpy(c.xterm, xt, sizeof(xt));
    const double qb[4] = {0.3 * PI, -0.7 * PI, 0.3 * PI, -0.7 * PI};
    std::memcpy(c.qbias, qb, sizeof(qb));
    c.GRF = 8.252 * 9.81;
    const double bj[8] = {PI / 4, -0.1, 1.15 * PI, -0.1, PI, PI - 0.2, .1, PI - 0.2};         // MHPCConstraints.cpp:53-54
    std::memcpy(c.bjoint, bj, sizeof(bj));
    c.btorque = 33;
    c.mu = 0.5f;
    c.fh_vel = 1.5; c.fh_ground = -0.404;                                                    // MHPCLocomotion.cpp:25
    c.hip_x = (0.19 * 2) / 2; c.hip_len = 0.209; c.knee_len = 0.195;                         // PlanarQuadruped.cpp:120-138
    const double qn[4] = {PI / 4, -PI * 7 / 12, PI / 4, -PI * 7 / 12};                       // boundingPDControl.cpp:6
    std::memcpy(c.qjoint_nom, qn, sizeof(qn));
    const double kp[4] = {5 * 8.0, 5 * 1.0, 5 * 12.0, 5 * 10.0};
    std::memcpy(c.kp, kp, sizeof(kp));
    c.legext_nom = 0.2462; c.kspring = 2200;
}

template <class T>
static int dalloc(mhpc_batch* b, T** p, size_t count) {
    void* q = nullptr;
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    CK(cudaMalloc(&q, bytes));
    CK(cudaMemsetAsync(q, 0, bytes, b->stream));
    b->allocs.push_back(q);
    b->bytes += (long long)bytes;
    *p = (T*)q;
    return 0;
}
static void free_all(mhpc_batch* b) {
    for (int c = 0; c < CLS_N; ++c) { if (b->pair_list[c]) cudaFree(b->pair_list[c]); b->pair_cap[c] = 0; }
    for (void* p : b->allocs) cudaFree(p);
    b->allocs.clear();
    b->bytes = 0;
    for (int c = 0; c < CLS_N; ++c) { b->knot_list[c] = nullptr; b->knot_cnt[c] = 0; b->pair_list[c] = nullptr; b->pair_cnt[c] = 0; }
}

// stage timing helpers
struct Stage {
    mhpc_batch* b; int bucket; bool on; cudaEvent_t a{}, e{};
    Stage(mhpc_batch* b_, int bucket_) : b(b_), bucket(bucket_), on(b_->timing) {
        if (on) { cudaEventCreate(&a); cudaEventCreate(&e); cudaEventRecord(a, b->stream); }
    }
    ~Stage() { if (on) { cudaEventRecord(e, b->stream); b->evs.push_back({a, e, bucket, b->cur_iter}); } }
};

static int gait_next(const mhpc_problem& p, int cur) {   // Gait.h:46-53
    for (int i = 0; i < 4; ++i) if (cur == p.gait_mode[i]) return p.gait_mode[(i + 1) % 4];
    return cur;
}

static inline int SL(const mhpc_batch* b, int inst) { return b->permuted ? b->slot_of[inst] : inst; }
// n per-instance values that sit in slot order on the device -> the caller's instance order
template <class T>
static int pull_per_instance(mhpc_batch* b, const T* dev, T* out) {
    if (!b->permuted) {
        CK(cudaMemcpyAsync(out, dev, b->n * sizeof(T), cudaMemcpyDeviceToHost, b->stream));
        CK(cudaStreamSynchronize(b->stream));
        return 0;
    }
    std::vector<T> tmp(b->n);
    CK(cudaMemcpyAsync(tmp.data(), dev, b->n * sizeof(T), cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    for (int s = 0; s < b->n; ++s) out[b->user_of[s]] = tmp[s];
    return 0;
}

extern "C" {

const char* mhpc_last_error(void) { return g_err.c_str(); }

// layout of the ABI structs as this library was compiled (a binding checks its mirror against these)
int mhpc_sizeof_option(void) { return (int)sizeof(mhpc_hsddp_option); }
int mhpc_sizeof_problem(void) { return (int)sizeof(mhpc_problem); }
int mhpc_sizeof_result(void) { return (int)sizeof(mhpc_result); }
int mhpc_sizeof_iter_record(void) { return (int)sizeof(mhpc_iter_record); }
int mhpc_sizeof_timing(void) { return (int)sizeof(mhpc_timing); }
int mhpc_struct_offsets(int which, int* out, int cap) {
#define OFF(T, f) (int)offsetof(T, f)
    std::vector<int> v;
    switch (which) {
        case 0: v = {OFF(mhpc_hsddp_option, alpha), OFF(mhpc_hsddp_option, gamma), OFF(mhpc_hsddp_option, update_penalty), OFF(mhpc_hsddp_option, update_relax),
                     OFF(mhpc_hsddp_option, update_regularization), OFF(mhpc_hsddp_option, update_ReB), OFF(mhpc_hsddp_option, max_DDP_iter),
                     OFF(mhpc_hsddp_option, max_AL_iter), OFF(mhpc_hsddp_option, DDP_thresh), OFF(mhpc_hsddp_option, AL_thresh),
                     OFF(mhpc_hsddp_option, AL_active), OFF(mhpc_hsddp_option, ReB_active), OFF(mhpc_hsddp_option, smooth_active)}; break;
        case 1: v = {OFF(mhpc_problem, n_wbphase), OFF(mhpc_problem, n_fbphase), OFF(mhpc_problem, dt_wb), OFF(mhpc_problem, dt_fb), OFF(mhpc_problem, cmode),
                     OFF(mhpc_problem, groundH), OFF(mhpc_problem, vel), OFF(mhpc_problem, height), OFF(mhpc_problem, roll), OFF(mhpc_problem, pitch),
                     OFF(mhpc_problem, yaw), OFF(mhpc_problem, gait_mode), OFF(mhpc_problem, gait_timing), OFF(mhpc_problem, x0)}; break;
        case 2: v = {OFF(mhpc_result, cost), OFF(mhpc_result, exp_cost_change), OFF(mhpc_result, viol), OFF(mhpc_result, status), OFF(mhpc_result, iter_AL),
                     OFF(mhpc_result, n_records), OFF(mhpc_result, pad)}; break;
        case 3: v = {OFF(mhpc_iter_record, iter_AL), OFF(mhpc_iter_record, iter_DDP), OFF(mhpc_iter_record, n_bws), OFF(mhpc_iter_record, n_fit),
                     OFF(mhpc_iter_record, reg), OFF(mhpc_iter_record, eps), OFF(mhpc_iter_record, cost), OFF(mhpc_iter_record, exp_cost_change),
                     OFF(mhpc_iter_record, viol), OFF(mhpc_iter_record, cost_prev)}; break;
        default: return fail("mhpc_struct_offsets: which = 0 option, 1 problem, 2 result, 3 iter_record");
    }
#undef OFF
    for (size_t k = 0; k < v.size() && (int)k < cap; ++k) out[k] = v[k];
    return (int)v.size();
}

void mhpc_default_option(mhpc_hsddp_option* o) {
    o->alpha = 0.1; o->gamma = 0.01; o->update_penalty = 8; o->update_relax = 0.1; o->update_regularization = 2;
    o->update_ReB = 7; o->max_DDP_iter = 3; o->max_AL_iter = 2; o->DDP_thresh = 1e-03; o->AL_thresh = 1e-03;
    o->AL_active = 1; o->ReB_active = 1; o->smooth_active = 0;
}
void mhpc_default_problem(mhpc_problem* p) {
    p->n_wbphase = 4; p->n_fbphase = 4; p->dt_wb = .001f; p->dt_fb = .001f; p->cmode = 1; p->groundH = -0.404f;
    p->vel = 1.5f; p->height = 0; p->roll = 0; p->pitch = 0; p->yaw = 0;
    const int gm[4] = {1, 2, 3, 4};
    const float gt[4] = {0.08f, 0.1f, 0.08f, 0.1f};
    std::memcpy(p->gait_mode, gm, sizeof(gm));
    std::memcpy(p->gait_timing, gt, sizeof(gt));
    const double x0[14] = {0.0927, -0.1093, -0.1542, 1.0957, -2.2033, 0.9742, -1.7098,
                           0.9011, 0.2756, 0.7333, 0.0446, 0.0009, 1.3219, 2.7346};
    std::memcpy(p->x0, x0, sizeof(x0));
}

int mhpc_batch_create(int n, const mhpc_hsddp_option* opt, int device, int ls_width, mhpc_batch** out) {
    if (n <= 0 || !opt || !out) return fail("mhpc_batch_create: bad arguments");
    if (check_option(*opt)) return -1;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail("mhpc_batch_create: no CUDA device (the hot path has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail("mhpc_batch_create: bad device index");
    CK(cudaSetDevice(device));
    cudaStream_t st = nullptr;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    mhpc_batch* b = new mhpc_batch();
    b->n = n; b->device = device;
    b->W = ls_width <= 0 ? 8 : std::min(ls_width, (int)MAX_W);
    b->stream = st;
    fill_opt(b->opt, *opt);
    fill_eps(b->eps, b->opt.alpha);
    fill_consts(b->c);
    { const char* e = getenv("MHPC_KEEP_DEAD_PARTIALS"); b->keep_dead_partials = e && e[0] == '1'; }
    { const char* e = getenv("MHPC_RECOST"); b->recost = !(e && e[0] == '0'); }
    { const char* e = getenv("MHPC_LADDER"); if (e) b->ladder = e[0] != '0'; }
    { const char* e = getenv("MHPC_BWS"); if (e) b->bws_kind = (e[0] == 'T' || e[0] == 't') ? 2 : (!strcmp(e, "16") ? 1 : (!strcmp(e, "8") ? 0 : 2)); }
    { const char* e = getenv("MHPC_LADDER_ROUNDS"); if (e && atoi(e) > 0) b->lad_rounds = atoi(e); }
    { const char* e = getenv("MHPC_LADDER_SYNC"); if (e) b->lad_sync = e[0] != '0'; }
    { const char* e = getenv("MHPC_LADDER_TARGET"); if (e) b->lad_target_on = e[0] != '0'; }
    CK(cudaHostAlloc((void**)&b->h_unres, LAD_CNT_N * sizeof(int), cudaHostAllocDefault));
    { const char* e = getenv("MHPC_LADDER_FIRST"); if (e && atoi(e) > 0) b->lad_first = atoi(e); }
    { const char* e = getenv("MHPC_LADDER_SMAX"); if (e && atoi(e) > 0) b->lad_smax = atoi(e); }
    *out = b;
    return 0;
}
void mhpc_batch_destroy(mhpc_batch* b) {
    if (!b) return;
    cudaSetDevice(b->device);
    cudaStreamSynchronize(b->stream);
    free_all(b);
    if (b->stage) cudaFree(b->stage);
    if (b->rot_tmp) cudaFree(b->rot_tmp);
    if (b->bws_trace_buf) cudaFree(b->bws_trace_buf);
    if (b->h_unres) cudaFreeHost(b->h_unres);
    if (b->d_user_of) cudaFree(b->d_user_of);
    if (b->side) { cudaStreamDestroy(b->side); cudaEventDestroy(b->ev_fork); cudaEventDestroy(b->ev_join); }
    if (b->side2) { cudaStreamDestroy(b->side2); cudaEventDestroy(b->ev_cfork); cudaEventDestroy(b->ev_cjoin); }
    for (auto& e : b->evs) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    if (b->own_stream && b->stream) cudaStreamDestroy(b->stream);
    delete b;
}
int mhpc_batch_set_option(mhpc_batch* b, const mhpc_hsddp_option* opt) {
    if (!b || !opt) return fail("set_option: bad arguments");
    if (check_option(*opt)) return -1;
    fill_opt(b->opt, *opt);
    fill_eps(b->eps, b->opt.alpha);
    if (b->have_problems) {
        CK(cudaSetDevice(b->device));
        // one iteration record per inner iteration: grow the record arrays with the iteration budget (never dropped silently)
        const int need_rec = (int)std::floor(b->opt.max_AL_iter) * (int)std::floor(b->opt.max_DDP_iter);
        if (need_rec > b->d.max_rec) {
            CK(cudaStreamSynchronize(b->stream));
            Dev& d = b->d;
            d.max_rec = need_rec;
            if (dalloc(b, &d.rec_i, (size_t)d.max_rec * REC_I * d.npad) || dalloc(b, &d.rec_d, (size_t)d.max_rec * REC_D * d.npad)) return -1;
            CK(cudaMemsetAsync(d.si + (size_t)SI_N_REC * d.npad, 0, d.npad * sizeof(int), b->stream));
            if (b->want_bws_trace && mhpc_batch_enable_bws_trace(b, 1)) return -1;
        }
        launch_reset_option_state(b->d, b->opt, b->stream);
        CK(cudaGetLastError());
    }
    return 0;
}
int mhpc_batch_set_stream(mhpc_batch* b, void* s) {
    CK(cudaSetDevice(b->device));
    CK(cudaStreamSynchronize(b->stream));
    if (b->own_stream && b->stream) cudaStreamDestroy(b->stream);
    if (s) { b->stream = (cudaStream_t)s; b->own_stream = false; }
    else { CK(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking)); b->own_stream = true; }
    return 0;
}
int mhpc_batch_synchronize(mhpc_batch* b) {
    CK(cudaSetDevice(b->device));
    CK(cudaStreamSynchronize(b->stream));
    return 0;
}
long long mhpc_batch_device_bytes(mhpc_batch* b) { return b->bytes; }

static int set_problems_slots(mhpc_batch* b, const mhpc_problem* pr);


// schedule key of a problem: phase split, then (mode, knot count) per phase - what decides which code a knot runs
static std::vector<int> schedule_key(const mhpc_problem& q) {
    std::vector<int> key{q.n_wbphase, q.n_fbphase};
    const int np = q.n_wbphase + q.n_fbphase;
    int mode = q.cmode;
    for (int p = 0; p < np && p < MAXP; ++p) {
        if (p > 0) mode = gait_next(q, mode);
        const float timing = (mode >= 1 && mode <= 4) ? q.gait_timing[mode - 1] : 0.f;
        const double dt = p < q.n_wbphase ? (double)q.dt_wb : (double)q.dt_fb;
        key.push_back(mode);
        key.push_back(dt > 0 ? (int)std::round(timing / dt) : 0);
    }
    return key;
}

// `keep_order`: the device state already belongs to the current slots (receding-horizon shift): do not re-sort
static int set_problems_ordered(mhpc_batch* b, const mhpc_problem* pr, bool keep_order) {
    CK(cudaSetDevice(b->device));
    const int n = b->n;
    auto same_inputs = [](const mhpc_problem& a, const mhpc_problem& c) {   // the fields schedule_key reads
        return a.n_wbphase == c.n_wbphase && a.n_fbphase == c.n_fbphase && a.cmode == c.cmode && a.dt_wb == c.dt_wb && a.dt_fb == c.dt_fb &&
               std::memcmp(a.gait_mode, c.gait_mode, sizeof(a.gait_mode)) == 0 && std::memcmp(a.gait_timing, c.gait_timing, sizeof(a.gait_timing)) == 0;
    };
    bool uniform_inputs = true;
    for (int i = 1; i < n && uniform_inputs; ++i) uniform_inputs = same_inputs(pr[0], pr[i]);
    if (!keep_order && uniform_inputs) {   // the common case costs one pass and no sort
        b->permuted = false;
        b->slot_of.clear();
        b->user_of.clear();
    } else if (!keep_order || (b->permuted && (int)b->slot_of.size() != n)) {
        std::vector<std::vector<int>> keys(n);
        for (int i = 0; i < n; ++i) keys[i] = schedule_key(pr[i]);
        std::vector<int> order(n);
        for (int i = 0; i < n; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int c) { return keys[a] < keys[c]; });
        bool identity = true;
        for (int s = 0; s < n; ++s) if (order[s] != s) { identity = false; break; }
        b->user_of = order;
        b->slot_of.assign(n, 0);
        for (int s = 0; s < n; ++s) b->slot_of[order[s]] = s;
        b->permuted = !identity;
        if (b->permuted) {
            if (b->d_user_of) cudaFree(b->d_user_of);
            CK(cudaMalloc(&b->d_user_of, n * sizeof(int)));
            CK(cudaMemcpy(b->d_user_of, b->user_of.data(), n * sizeof(int), cudaMemcpyHostToDevice));
        }
    }
    b->problems.assign(pr, pr + n);
    if (!b->permuted) return set_problems_slots(b, pr);
    std::vector<mhpc_problem> ps(n);
    for (int s = 0; s < n; ++s) ps[s] = pr[b->user_of[s]];
    return set_problems_slots(b, ps.data());
}
int mhpc_batch_set_problems(mhpc_batch* b, const mhpc_problem* pr) { return set_problems_ordered(b, pr, false); }

// everything below works in slot order
static int set_problems_slots(mhpc_batch* b, const mhpc_problem* pr) {
    const int n = b->n, npad = (n + 31) / 32 * 32, nblk = npad / 32;
    int P = 1, KT = 1;
    for (int i = 0; i < n; ++i) {
        const int np = pr[i].n_wbphase + pr[i].n_fbphase;
        if (pr[i].n_wbphase < 1) return fail("set_problems: n_wbphase must be >= 1 (MHPCLocomotion.cpp:176)");
        if (np > MAXP) return fail("set_problems: too many phases");
        P = std::max(P, np);
    }
    const std::vector<int> old_nph = b->h_nph, old_nwb = b->h_nwb, old_mode = b->h_mode, old_N = b->h_N;
    b->h_nph.assign(npad, 0); b->h_nwb.assign(npad, 0); b->h_nk.assign(npad, 0);
    b->h_mode.assign((size_t)P * npad, 1); b->h_N.assign((size_t)P * npad, 0); b->h_off.assign((size_t)P * npad, 0);
    std::vector<double> dtw(npad, 0), dtf(npad, 0), vel(npad, 0), hgt(npad, 0), x0((size_t)14 * npad, 0);
    for (int i = 0; i < n; ++i) {
        const mhpc_problem& q = pr[i];
        const int np = q.n_wbphase + q.n_fbphase;
        b->h_nph[i] = np; b->h_nwb[i] = q.n_wbphase;
        const double dt_wb = q.dt_wb, dt_fb = q.dt_fb;   // float -> double (MHPCLocomotion.h:59)
        dtw[i] = dt_wb; dtf[i] = dt_fb; vel[i] = q.vel; hgt[i] = q.height;
        int mode = q.cmode, used = 0;
        for (int p = 0; p < np; ++p) {   // Gait.h:55-77, MHPCLocomotion.cpp:76,85
            if (p > 0) mode = gait_next(q, mode);
            if (mode < 1 || mode > 4) return fail("set_problems: modes must be 1..4");
            const float timing = q.gait_timing[mode - 1];
            const int N = (int)std::round(timing / (p < q.n_wbphase ? dt_wb : dt_fb));
            if (N < 2) return fail("set_problems: a phase needs at least 2 knots");
            if (N > KSTRIDE) return fail("set_problems: a phase may have at most N_TIMESTEPS_MAX = 110 knots (MHPC_CPPTypes.h:9)");
            b->h_mode[(size_t)p * npad + i] = mode; b->h_N[(size_t)p * npad + i] = N; b->h_off[(size_t)p * npad + i] = p * KSTRIDE;
            used += N;
        }
        b->h_nk[i] = used;
        KT = std::max(KT, np * KSTRIDE);
        for (int e = 0; e < 14; ++e) x0[(size_t)e * npad + i] = q.x0[e];
    }
    // same schedule as the previous batch (the common MPC case): only the per-instance scalars change
    const bool same_schedule = b->have_problems && b->d.npad == npad && b->d.P == P && old_nph == b->h_nph && old_nwb == b->h_nwb &&
                               old_mode == b->h_mode && old_N == b->h_N;
    auto upload_scalars = [&]() -> int {
        Dev& d = b->d;
        auto up = [&](void* dst, const void* src, size_t bytes) { return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, b->stream); };
        CK(up(d.dt_wb, dtw.data(), npad * sizeof(double))); CK(up(d.dt_fb, dtf.data(), npad * sizeof(double)));
        CK(up(d.vel, vel.data(), npad * sizeof(double))); CK(up(d.height, hgt.data(), npad * sizeof(double)));
        CK(up(d.x0, x0.data(), (size_t)14 * npad * sizeof(double)));
        CK(cudaStreamSynchronize(b->stream));   // host vectors go out of scope
        return 0;
    };
    if (same_schedule) return upload_scalars();
    // (re)allocate when the shape grows
    if (!b->have_problems || KT > b->alloc_KT || P > b->alloc_P || b->d.npad != npad) {
        CK(cudaStreamSynchronize(b->stream));
        free_all(b);
        Dev& d = b->d;
        d = Dev{};
        d.n = n; d.npad = npad; d.nblk = nblk; d.P = P; d.KT = KT; d.W = b->W;
        d.max_rec = std::max(1, (int)std::floor(b->opt.max_AL_iter) * (int)std::floor(b->opt.max_DDP_iter));
        d.max_rec = std::max(d.max_rec, 64);
        int rc = 0;
        rc |= dalloc(b, &d.n_phases, npad); rc |= dalloc(b, &d.n_wb, npad); rc |= dalloc(b, &d.n_knots, npad);
        rc |= dalloc(b, &d.ph_mode, (size_t)P * npad); rc |= dalloc(b, &d.ph_N, (size_t)P * npad); rc |= dalloc(b, &d.ph_off, (size_t)P * npad);
        rc |= dalloc(b, &d.dt_wb, npad); rc |= dalloc(b, &d.dt_fb, npad); rc |= dalloc(b, &d.vel, npad); rc |= dalloc(b, &d.height, npad);
        rc |= dalloc(b, &d.x0, (size_t)14 * npad);
        rc |= dalloc(b, &d.pos_ref, (size_t)KT * npad);
        rc |= dalloc(b, &d.st_nom, (size_t)KT * ST * npad);
        rc |= dalloc(b, &d.st_act, (size_t)d.W * KT * ST * npad);
        rc |= dalloc(b, &d.gain, (size_t)KT * GREC * npad);
        rc |= dalloc(b, &d.lin, (size_t)npad * KT * REC);
        rc |= dalloc(b, &d.par, (size_t)P * NPAR * npad);
        rc |= dalloc(b, &d.tr_cost, (size_t)d.W * npad); rc |= dalloc(b, &d.tr_viol2, (size_t)d.W * npad);
        rc |= dalloc(b, &d.tr_h, (size_t)d.W * P * npad);
        rc |= dalloc(b, &d.knot_cost, (size_t)KT * npad);
        rc |= dalloc(b, &d.sd, (size_t)SD_N * npad); rc |= dalloc(b, &d.si, (size_t)SI_N * npad);
        rc |= dalloc(b, &d.rec_i, (size_t)d.max_rec * REC_I * npad); rc |= dalloc(b, &d.rec_d, (size_t)d.max_rec * REC_D * npad);
        rc |= dalloc(b, &d.n_active, 1);
        // the tensor-core sweep addresses its outputs with 32-bit element offsets
        if (b->ladder) {
            d.lad_cap = b->bws_kind == 2 ? bwsT_resident_warps(b->device) : b->bws_kind == 1 ? bws16_resident_lane_groups(b->device) : bws_resident_lane_groups(b->device);
            d.lad_tcap = 2 * npad + d.lad_cap;
            d.lad_target = (b->bws_kind == 2 && b->lad_target_on) ? n : 0;   // the tensor-core sweep is throughput-bound from ~2400 tasks on
            const size_t tcap = (size_t)d.lad_tcap;
            rc |= dalloc(b, &d.lad_cnt, LAD_CNT_N); rc |= dalloc(b, &d.lad_unres, npad); rc |= dalloc(b, &d.lad_rung, npad);
            rc |= dalloc(b, &d.lad_t_inst, tcap); rc |= dalloc(b, &d.lad_t_rung, tcap); rc |= dalloc(b, &d.lad_t_reg, tcap); rc |= dalloc(b, &d.lad_t_dv, tcap);
            rc |= dalloc(b, &d.lad_u_inst, tcap); rc |= dalloc(b, &d.lad_u_copy, tcap);
            // scratch outputs of the speculative tasks: gain records, one column per task that does not write in place (< npad)
            rc |= dalloc(b, &d.lad_gain, (size_t)KT * GREC * npad);
        }
        for (int c = 0; c < CLS_N; ++c) rc |= dalloc(b, &b->knot_list[c], KT);
        if (rc) return -1;
        b->alloc_KT = KT; b->alloc_P = P;
        launch_reset_option_state(d, b->opt, b->stream);   // MultiPhaseDDP(n_phases, option): _option = option
    }
    Dev& d = b->d;
    // record lengths per (block, knot) and knot lists per class
    std::set<int> cls_knots[CLS_N];
    for (int i = 0; i < n; ++i) {
        const int np = b->h_nph[i], nwb = b->h_nwb[i];
        for (int p = 0; p < np; ++p) {
            const int mode = b->h_mode[(size_t)p * npad + i], N = b->h_N[(size_t)p * npad + i], off = b->h_off[(size_t)p * npad + i];
            const bool wb = p < nwb;
            for (int k = 0; k < N; ++k) {
                const bool term = (k == N - 1);
                int cls;
                if (wb && !term) cls = mode == 1 ? CLS_BS : mode == 3 ? CLS_FS : CLS_FL;
                else if (wb) cls = mode == 2 ? CLS_IMPF : mode == 4 ? CLS_IMPB : CLS_LIGHT;
                else cls = CLS_LIGHT;
                cls_knots[cls].insert(off + k);
            }
        }
    }
    // uniform batch <=> every instance has the schedule of instance 0
    b->uniform = true;
    for (int i = 1; i < n && b->uniform; ++i) {
        if (b->h_nph[i] != b->h_nph[0] || b->h_nwb[i] != b->h_nwb[0]) b->uniform = false;
        for (int p = 0; p < b->h_nph[0] && b->uniform; ++p)
            if (b->h_mode[(size_t)p * npad + i] != b->h_mode[(size_t)p * npad] || b->h_N[(size_t)p * npad + i] != b->h_N[(size_t)p * npad]) b->uniform = false;
    }
    if (!b->uniform) {
        if (n >= (1 << 20) || d.KT >= (1 << 12)) return fail("set_problems: non-uniform batches support < 2^20 instances and < 4096 knots");
        std::vector<std::vector<unsigned>> pairs(CLS_N);
        std::vector<std::vector<int>> cls_at((size_t)d.KT);   // per knot: (instance, class) in instance order
        for (int c = 0; c < CLS_N; ++c) pairs[c].reserve((size_t)n * 64);
        std::vector<unsigned char> cls_of((size_t)n * d.KT, 255);
        for (int i = 0; i < n; ++i) {
            const int np = b->h_nph[i], nwb = b->h_nwb[i];
            for (int p = 0; p < np; ++p) {
                const int mode = b->h_mode[(size_t)p * npad + i], N = b->h_N[(size_t)p * npad + i], off = b->h_off[(size_t)p * npad + i];
                const bool wb = p < nwb;
                for (int k = 0; k < N; ++k) {
                    const bool term = (k == N - 1);
                    int cls;
                    if (wb && !term) cls = mode == 1 ? CLS_BS : mode == 3 ? CLS_FS : CLS_FL;
                    else if (wb) cls = mode == 2 ? CLS_IMPF : mode == 4 ? CLS_IMPB : CLS_LIGHT;
                    else cls = CLS_LIGHT;
                    cls_of[(size_t)(off + k) * n + i] = (unsigned char)cls;
                }
            }
        }
        for (int kk = 0; kk < d.KT; ++kk)       // knot-major, instance-minor: neighbouring lanes read neighbouring instances
            for (int i = 0; i < n; ++i) {
                const unsigned char c = cls_of[(size_t)kk * n + i];
                if (c != 255) pairs[c].push_back(((unsigned)kk << 20) | (unsigned)i);
            }
        for (int c = 0; c < CLS_N; ++c) {
            b->pair_cnt[c] = (int)pairs[c].size();
            if (pairs[c].empty()) continue;
            if (!b->pair_list[c] || b->pair_cap[c] < pairs[c].size()) {
                if (b->pair_list[c]) cudaFree(b->pair_list[c]);
                CK(cudaMalloc(&b->pair_list[c], pairs[c].size() * sizeof(unsigned)));
                b->pair_cap[c] = pairs[c].size();
            }
            CK(cudaMemcpy(b->pair_list[c], pairs[c].data(), pairs[c].size() * sizeof(unsigned), cudaMemcpyHostToDevice));
        }
    }
    for (int c = 0; c < CLS_N; ++c) {
        std::vector<int> v(cls_knots[c].begin(), cls_knots[c].end());
        b->knot_cnt[c] = (int)v.size();
        if (!v.empty()) CK(cudaMemcpyAsync(b->knot_list[c], v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice, b->stream));
        CK(cudaStreamSynchronize(b->stream));
    }
    auto up = [&](void* dst, const void* src, size_t bytes) { return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, b->stream); };
    CK(up(d.n_phases, b->h_nph.data(), npad * sizeof(int))); CK(up(d.n_wb, b->h_nwb.data(), npad * sizeof(int)));
    CK(up(d.n_knots, b->h_nk.data(), npad * sizeof(int)));
    // schedule arrays may have been built with a smaller P than allocated: copy row by row
    for (int p = 0; p < P; ++p) {
        CK(up(d.ph_mode + (size_t)p * npad, b->h_mode.data() + (size_t)p * npad, npad * sizeof(int)));
        CK(up(d.ph_N + (size_t)p * npad, b->h_N.data() + (size_t)p * npad, npad * sizeof(int)));
        CK(up(d.ph_off + (size_t)p * npad, b->h_off.data() + (size_t)p * npad, npad * sizeof(int)));
    }
    if (upload_scalars()) return -1;
    b->have_problems = true;
    if (b->want_bws_trace) return mhpc_batch_enable_bws_trace(b, 1);
    return 0;
}

// nominal <- accepted rollout on the second side stream (the mask is frozen first: later kernels of the main stream change
// the flags it is made of); join_copy() before anything reads the nominal or overwrites the rollout slots
static int fork_copy_nominal(mhpc_batch* b, int mask_mode) {
    if (!b->side2) {
        CK(cudaStreamCreateWithFlags(&b->side2, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&b->ev_cfork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&b->ev_cjoin, cudaEventDisableTiming));
    }
    launch_snapshot_mask(b->d, mask_mode, b->stream);
    CK(cudaEventRecord(b->ev_cfork, b->stream));
    CK(cudaStreamWaitEvent(b->side2, b->ev_cfork, 0));
    launch_copy_nominal(b->d, MASK_SNAP, b->side2);
    CK(cudaEventRecord(b->ev_cjoin, b->side2));
    b->copy_pending = true;
    b->launches += 2;
    return 0;
}
static int join_copy(mhpc_batch* b) {
    if (b->copy_pending) { CK(cudaStreamWaitEvent(b->stream, b->ev_cjoin, 0)); b->copy_pending = false; }
    return 0;
}

static int linearize(mhpc_batch* b, int mask_mode) {
    Stage st(b, TB_LIN);
    const Dev& d = b->d;
    typedef void (*lin_fn)(const Dev&, const Consts&, const Opt&, int, const int*, int, cudaStream_t);
    const lin_fn fns[CLS_N] = {launch_lin_BS, launch_lin_FS, launch_lin_FL, launch_lin_IMPF, launch_lin_IMPB, launch_lin_LIGHT};
    if (!b->side) {
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&b->side, cudaStreamNonBlocking, hi));
        CK(cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&b->ev_join, cudaEventDisableTiming));
    }
    auto launch_cls = [&](int c, cudaStream_t s) {
        if (b->uniform) fns[c](d, b->c, b->opt, mask_mode, b->knot_list[c], b->knot_cnt[c], s);
        else fns[c](d, b->c, b->opt, mask_mode, (const int*)b->pair_list[c], -b->pair_cnt[c], s);   // negative count = pair list
    };
    // the knot classes write disjoint records: impacts fork onto the side stream, everything else stays in order
    CK(cudaEventRecord(b->ev_fork, b->stream));
    CK(cudaStreamWaitEvent(b->side, b->ev_fork, 0));
    launch_cls(CLS_IMPF, b->side);
    launch_cls(CLS_IMPB, b->side);
    CK(cudaEventRecord(b->ev_join, b->side));
    for (int c : {CLS_BS, CLS_FS, CLS_FL, CLS_LIGHT}) launch_cls(c, b->stream);
    CK(cudaStreamWaitEvent(b->stream, b->ev_join, 0));
    for (int c = 0; c < CLS_N; ++c) b->launches += b->knot_cnt[c] > 0;
    CK(cudaGetLastError());
    return 0;
}
static int need(mhpc_batch* b) {
    if (!b || !b->have_problems) return fail("no problems set (call mhpc_batch_set_problems first)");
    CK(cudaSetDevice(b->device));
    b->gen++;
    return 0;
}

// The per-knot H, Qx, Qu, Qxx, Qux, Quu of CostToGoStruct (MHPC_CompoundTypes.h:100-144) are not kept by the solver: the sweep holds
// them on chip.  On request they are recomputed for ONE instance by running its last accepted backward sweep again - same
// knot records, same regularisation - as a single task of the ladder scheduler's task kernel in dump mode (k_bws<1, 1>): the
// task writes its K, du, G into a scratch column and its outcome into the task arrays, so no solver state changes.
static int replay_cost_to_go(mhpc_batch* b, int slot) {
    Dev& d = b->d;
    if (!d.lad_cnt) {   // the ladder scheduler is off (MHPC_LADDER=0): a two-entry task list and one scratch column for the replay alone
        int rc = dalloc(b, &d.lad_cnt, LAD_CNT_N);
        rc |= dalloc(b, &d.lad_t_inst, 2); rc |= dalloc(b, &d.lad_t_rung, 2); rc |= dalloc(b, &d.lad_t_reg, 2); rc |= dalloc(b, &d.lad_t_dv, 2);
        rc |= dalloc(b, &d.lad_gain, (size_t)d.KT * GREC * (GAIN_KNOT_MAJOR ? d.npad : 1));
        if (rc) return -1;
    }
    if (!d.ctg_dump) { if (dalloc(b, &d.ctg_dump, (size_t)d.KT * CTGREC)) return -1; }
    double reg = 0.0; int nbws = 0, ok = 0;
    CK(cudaMemcpyAsync(&reg, d.sd + (size_t)SD_REG_USED * d.npad + slot, sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    CK(cudaMemcpyAsync(&nbws, d.si + (size_t)SI_N_BWS * d.npad + slot, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    CK(cudaMemcpyAsync(&ok, d.si + (size_t)SI_BWS_OK * d.npad + slot, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    if (!ok) return fail("get_field: the last backward sweep of this instance did not pass: no CostToGoStruct to report");
    const int cnt[LAD_CNT_N] = {2, 1, 0, 2};            // LAD_N_TASKS, LAD_N_UNRES, LAD_STRIDE, LAD_S: task 1 = speculative task of instance `slot`
    const int t_inst[2] = {-1, slot}, t_rung[2] = {0, nbws > 0 ? nbws - 1 : 0};
    const double t_reg[2] = {0.0, reg};
    static_assert(LAD_N_TASKS == 0 && LAD_N_UNRES == 1 && LAD_STRIDE == 2 && LAD_S == 3, "task-list header layout");
    CK(cudaMemcpyAsync(d.lad_cnt, cnt, sizeof(cnt), cudaMemcpyHostToDevice, b->stream));
    CK(cudaMemcpyAsync(d.lad_t_inst, t_inst, sizeof(t_inst), cudaMemcpyHostToDevice, b->stream));
    CK(cudaMemcpyAsync(d.lad_t_rung, t_rung, sizeof(t_rung), cudaMemcpyHostToDevice, b->stream));
    CK(cudaMemcpyAsync(d.lad_t_reg, t_reg, sizeof(t_reg), cudaMemcpyHostToDevice, b->stream));
    CK(cudaMemsetAsync(d.ctg_dump, 0, (size_t)d.KT * CTGREC * sizeof(double), b->stream));
    launch_bws_replay(d, b->c, b->opt, b->stream);
    int passed[2] = {0, -1};
    CK(cudaMemcpyAsync(passed, d.lad_t_rung, sizeof(passed), cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    CK(cudaGetLastError());
    if (passed[1] < 0) return fail("get_field: the replay of the accepted backward sweep did not pass (the knot records have changed since)");
    return 0;
}

int mhpc_batch_initialization(mhpc_batch* b) {
    if (need(b)) return -1;
    Dev& d = b->d;
    Stage st(b, TB_OTHER);
    // memory_reset (MHPCLocomotion.cpp:265-290)
    CK(cudaMemsetAsync(d.st_nom, 0, (size_t)d.KT * ST * d.npad * sizeof(double), b->stream));
    CK(cudaMemsetAsync(d.st_act, 0, (size_t)d.W * d.KT * ST * d.npad * sizeof(double), b->stream));
    CK(cudaMemsetAsync(d.gain, 0, (size_t)d.KT * GREC * d.npad * sizeof(double), b->stream));
    // (the knot records `lin` - more than half of the footprint - are not reset: every record of a used knot is written in
    //  full by the linearisation before the sweep reads it, and no kernel or getter reads the record of an unused knot)
    launch_setup(d, b->c, b->stream);                 // build_problem: references + AL/ReB parameters
    launch_warmstart(d, b->c, b->stream);             // warmstart()
    launch_copy_nominal(d, MASK_ALL, b->stream);      // update_nominal_trajectory()
    b->launches += 3;
    CK(cudaGetLastError());
    return 0;
}

// MultiPhaseDDP::set_initial_condition (MultiPhaseDDP.h:24-25) for every instance: x0 is [n][14] on the host
int mhpc_batch_set_initial_condition(mhpc_batch* b, const double* x0) {
    if (need(b)) return -1;
    std::vector<double> soa((size_t)14 * b->d.npad, 0.0);
    for (int i = 0; i < b->n; ++i)
        for (int e = 0; e < 14; ++e) { b->problems[i].x0[e] = x0[(size_t)i * 14 + e]; soa[(size_t)e * b->d.npad + SL(b, i)] = x0[(size_t)i * 14 + e]; }
    CK(cudaMemcpyAsync(b->d.x0, soa.data(), soa.size() * sizeof(double), cudaMemcpyHostToDevice, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    return 0;
}

// MHPCLocomotion::update_problem (MHPCLocomotion.cpp:107-158): receding-horizon shift by one phase.
int mhpc_batch_update_problem(mhpc_batch* b) {
    if (need(b)) return -1;
    Dev& d = b->d;
    // (1) rotate the phase buffers of each model group on the device (nominal, accepted rollout, K, du, G)
    const size_t tmp_need = (size_t)d.KT * GREC * d.npad * sizeof(double);
    if (b->rot_bytes < tmp_need) {
        if (b->rot_tmp) cudaFree(b->rot_tmp);
        CK(cudaMalloc(&b->rot_tmp, tmp_need));
        b->rot_bytes = tmp_need;
    }
    double* tmp = (double*)b->rot_tmp;
    struct Arr { double* p; int elems; int acc; };
    const Arr arrs[2] = {{d.st_nom, ST, 0}, {d.st_act, ST, 1}};
    {   // gain records (instance-major): same rotation, one thread per (knot slot, instance)
        const size_t bytes = (size_t)d.KT * GREC * d.npad * sizeof(double);
        CK(cudaMemcpyAsync(tmp, d.gain, bytes, cudaMemcpyDeviceToDevice, b->stream));
        launch_rotate_gain(d, d.gain, tmp, b->stream);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(d.gain, tmp, bytes, cudaMemcpyDeviceToDevice, b->stream));
    }
    for (const Arr& a : arrs) {
        const size_t bytes = (size_t)d.KT * a.elems * d.npad * sizeof(double);
        CK(cudaMemcpyAsync(tmp, a.p, bytes, cudaMemcpyDeviceToDevice, b->stream));      // rows of unused phases keep their content
        launch_rotate(d, a.p, tmp, a.elems, a.acc, b->stream);
        CK(cudaGetLastError());   // nothing of the handle's schedule changes unless every rotation was launched
        CK(cudaMemcpyAsync(a.p, tmp, bytes, cudaMemcpyDeviceToDevice, b->stream));      // (for st_act: the accepted rollout becomes slot 0)
    }
    CK(cudaStreamSynchronize(b->stream));
    // (2) the gait advances one mode; schedule, knot lists and x0 are re-uploaded through the normal path
    std::vector<mhpc_problem> pr = b->problems;
    for (mhpc_problem& q : pr) q.cmode = gait_next(q, q.cmode);
    if (set_problems_ordered(b, pr.data(), true)) return -1;
    // (3) every phase is re-initialised (AL/ReB parameters) and the reference regenerated from _x0
    launch_setup_refs_params(d, b->c, b->stream);
    launch_spread_last_knot(d, b->stream);
    CK(cudaGetLastError());
    return 0;
}

int mhpc_batch_forward_sweep(mhpc_batch* b, double eps) {
    if (need(b)) return -1;
    EpsList one{}; one.n = 1; one.eps[0] = eps;
    // the step-level API applies the option's ReB flag as is (the solve loop overrides it per AL iteration)
    CK(cudaMemsetAsync(b->d.si + (size_t)SI_STATUS * b->d.npad, 0, b->d.npad * sizeof(int), b->stream));
    launch_reb_from_option(b->d, b->stream);
    launch_rollout(b->d, b->c, b->opt, MASK_ALL, 1, 0, one, b->stream);
    launch_accept_slot0(b->d, MASK_ALL, 2, b->stream);
    if (linearize(b, MASK_ALL)) return -1;
    CK(cudaGetLastError());
    return 0;
}
int mhpc_batch_forward_sweep_dynamics_only(mhpc_batch* b, double eps) {
    if (need(b)) return -1;
    EpsList one{}; one.n = 1; one.eps[0] = eps;
    launch_reb_from_option(b->d, b->stream);
    launch_rollout(b->d, b->c, b->opt, MASK_ALL, 1, 0, one, b->stream);
    launch_accept_slot0(b->d, MASK_ALL, 1, b->stream);
    CK(cudaGetLastError());
    return 0;
}
int mhpc_batch_forward_sweep_partials_only(mhpc_batch* b) {
    if (need(b)) return -1;
    std::vector<int> v(b->d.npad, 1);
    CK(cudaMemcpyAsync(b->d.si + (size_t)SI_AL_PARTIALS * b->d.npad, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    launch_reb_from_option(b->d, b->stream);
    return linearize(b, MASK_ALL);
}
int mhpc_batch_backward_sweep(mhpc_batch* b, double reg, int* success) {
    if (need(b)) return -1;
    std::vector<double> r(b->d.npad, reg);
    CK(cudaMemcpyAsync(b->d.sd + (size_t)SD_REG * b->d.npad, r.data(), r.size() * sizeof(double), cudaMemcpyHostToDevice, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    {
        Stage st(b, TB_BWS);
        (b->bws_kind == 2 ? launch_bwsT : b->bws_kind == 1 ? launch_bws16 : launch_bws)(b->d, b->c, b->opt, MASK_ALL, 1, 0, b->stream);
    }
    CK(cudaGetLastError());
    if (success && pull_per_instance(b, b->d.si + (size_t)SI_BWS_OK * b->d.npad, success)) return -1;
    return 0;
}
static int line_search(mhpc_batch* b, int mask_ls, int mask_sel) {
    if (join_copy(b)) return -1;
    Stage st(b, TB_LS);
    const Dev& d = b->d;
    for (int t0 = 0; t0 < b->eps.n; t0 += d.W) {
        const int ns = std::min(d.W, b->eps.n - t0);
        const int last = (t0 + ns >= b->eps.n);
        launch_rollout(d, b->c, b->opt, mask_ls, ns, t0, b->eps, b->stream);
        launch_select(d, b->opt, mask_sel, t0, ns, last, b->eps, b->stream);
        b->launches += 2;
    }
    CK(cudaGetLastError());
    return 0;
}
int mhpc_batch_forward_iteration(mhpc_batch* b, int* trials) {
    if (need(b)) return -1;
    std::vector<int> act(b->d.npad, 1);
    CK(cudaMemsetAsync(b->d.si + (size_t)SI_STATUS * b->d.npad, 0, b->d.npad * sizeof(int), b->stream));
    CK(cudaMemcpyAsync(b->d.si + (size_t)SI_DDP_ACTIVE * b->d.npad, act.data(), act.size() * sizeof(int), cudaMemcpyHostToDevice, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    launch_reb_from_option(b->d, b->stream);
    launch_ls_begin(b->d, MASK_ALL, b->stream);
    if (line_search(b, MASK_LS, MASK_ALL)) return -1;
    if (trials && pull_per_instance(b, b->d.si + (size_t)SI_N_FIT * b->d.npad, trials)) return -1;
    return 0;
}
int mhpc_batch_update_nominal_trajectory(mhpc_batch* b) {
    if (need(b)) return -1;
    launch_copy_nominal(b->d, MASK_ALL, b->stream);
    CK(cudaGetLastError());
    return 0;
}
int mhpc_batch_update_AL_ReB_param(mhpc_batch* b) {
    if (need(b)) return -1;
    launch_al_end(b->d, b->opt, b->c, MASK_ALL, 0, b->stream);
    CK(cudaGetLastError());
    return 0;
}

int mhpc_batch_solve(mhpc_batch* b, int sync_early_exit) {
    if (need(b)) return -1;
    const Dev& d = b->d;
    cudaStream_t s = b->stream;
    const int maxAL = (int)std::floor(b->opt.max_AL_iter), maxDDP = (int)std::floor(b->opt.max_DDP_iter);
    // forward_sweep(0) runs as `nfs` identical trials of step size 0: a lane group of four shares the shared-memory staging
    // of the nominal trajectory and gains, which is faster (1.25 ms) than 32 instances per warp each reading in place
    // (1.8 ms); trial slot 0 is the one that is used, slots 1.. are scratch that the next line search overwrites
    const int nfs = std::min(d.W, 4);
    EpsList zero{}; zero.n = nfs;
    for (int t = 0; t < nfs; ++t) zero.eps[t] = 0.0;
    b->launches = 0;
    for (auto& e : b->evs) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    b->evs.clear();
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (b->timing) { cudaEventCreate(&t0); cudaEventCreate(&t1); cudaEventRecord(t0, s); }
    launch_solve_begin(d, s); b->launches++;
    for (int al = 1; al <= maxAL; ++al) {
        b->cur_iter = al * 1000 + 1;   // the full forward sweep of an outer iteration is booked on its first inner iteration
        { Stage st(b, TB_OTHER); launch_al_begin(d, b->opt, s); b->launches++; }
        if (join_copy(b)) return -1;
        if (al == 1 || !b->recost) {
            { Stage st(b, TB_ROLLOUT);                                   // forward_sweep(0): rollout ...
              launch_rollout(d, b->c, b->opt, MASK_RUNNING, nfs, 0, zero, s);
              launch_accept_slot0(d, MASK_RUNNING, 2, s); b->launches += 2; }
            if (linearize(b, MASK_RUNNING)) return -1;                   // ... and its derivatives
            { Stage st(b, TB_OTHER); if (fork_copy_nominal(b, MASK_RUNNING)) return -1; }   // update_nominal_trajectory
        } else {
            // After the first outer iteration the step-size-0 rollout reproduces the nominal trajectory bit for bit (k_rollout.cu,
            // "cost-only pass"): only its costs are new.  The accepted trial slot of the last line search still holds that
            // trajectory and the nominal copy of it, so neither is rewritten.
            { Stage st(b, TB_ROLLOUT); launch_recost(d, b->c, b->opt, MASK_RUNNING, 2, s); b->launches += 3; }
            if (linearize(b, MASK_RUNNING)) return -1;
        }
        for (int it = 1; it <= maxDDP; ++it) {
            b->cur_iter = al * 1000 + it;
            { Stage st(b, TB_OTHER); launch_ls_begin(d, MASK_DDP, s); b->launches++; }
            { Stage st(b, TB_BWS);   // backward_sweep + regularisation retries (MultiPhaseDDP.cpp:209-227)
              auto sweep = b->bws_kind == 2 ? launch_bwsT : b->bws_kind == 1 ? launch_bws16 : launch_bws;
              if (!b->ladder) { sweep(d, b->c, b->opt, MASK_DDP, 0, 0, s); b->launches++; }
              else {
                  sweep(d, b->c, b->opt, MASK_DDP, 0, b->lad_first, s); b->launches++;
                  auto round = [&](int n_tasks, int n_unres) {
                      if (b->bws_kind == 2) launch_bwsT_tasks(d, b->c, b->opt, MASK_DDP, n_tasks, s); else if (b->bws_kind == 1) launch_bws16_tasks(d, b->c, b->opt, MASK_DDP, n_tasks, s);
                      else launch_bws_tasks(d, b->c, b->opt, MASK_DDP, n_tasks, s);
                      launch_lad_finish(d, b->opt, n_unres, s);
                      b->launches += 3;
                  };
                  if (b->lad_sync) {
                      // plan -> how many instances still have no passing rung? -> tasks, resolve, copy; until none is left.  One
                      // 4-byte read-back per round: a stage whose first launch settles everything (5 of the 6 stages of the bench
                      // workload) costs one plan kernel and one short wait instead of rounds of empty launches.
                      for (int r = 1; r <= 64; ++r) {
                          launch_lad_plan(d, b->opt, MASK_DDP, 0, b->lad_smax, s); b->launches++;
                          CK(cudaMemcpyAsync(b->h_unres, d.lad_cnt, LAD_CNT_N * sizeof(int), cudaMemcpyDeviceToHost, s));
                          CK(cudaStreamSynchronize(s));
                          if (b->h_unres[LAD_N_UNRES] == 0) break;
                          round(b->h_unres[LAD_N_TASKS], b->h_unres[LAD_N_UNRES]);
                      }
                  } else {
                      for (int r = 1; r <= b->lad_rounds; ++r) {
                          const int last = r == b->lad_rounds;
                          launch_lad_plan(d, b->opt, MASK_DDP, last, b->lad_smax, s); b->launches++;
                          round(0, 0);
                      }
                  }
              } }
            if (line_search(b, MASK_LS, MASK_DDP)) return -1;
            { Stage st(b, TB_OTHER);
              if (fork_copy_nominal(b, MASK_DDP)) return -1;
              launch_ddp_end(d, b->opt, it, s); b->launches++; }
            // forward_sweep_partials_only (MultiPhaseDDP.cpp:262).  After the LAST inner iteration the reference still
            // linearises, but every derivative it writes is overwritten by the next outer iteration's forward_sweep(0)
            // (same trajectory: K (x - x_nom) = 0 exactly) or never read (end of solve): the launch is dead work and is
            // skipped; no result, record or trajectory depends on it.  MHPC_KEEP_DEAD_PARTIALS=1 restores it.
            if (it < maxDDP || b->keep_dead_partials) { if (linearize(b, MASK_LIN)) return -1; }
        }
        { Stage st(b, TB_OTHER); launch_al_end(d, b->opt, b->c, MASK_RUNNING, 1, s); b->launches++; }
        if (sync_early_exit && al < maxAL) {
            launch_count_active(d, s);
            int na = 0;
            CK(cudaMemcpyAsync(&na, d.n_active, sizeof(int), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            if (na == 0) break;
        }
    }
    if (join_copy(b)) return -1;
    CK(cudaGetLastError());
    if (b->timing) {
        cudaEventRecord(t1, s);
        CK(cudaStreamSynchronize(s));
        mhpc_timing t{};
        cudaEventElapsedTime(&t.total_ms, t0, t1);
        float acc[TB_N] = {0};
        for (auto& e : b->evs) { float ms = 0; cudaEventElapsedTime(&ms, e.a, e.b); acc[e.bucket] += ms; }
        t.rollout_ms = acc[TB_ROLLOUT]; t.linearize_ms = acc[TB_LIN]; t.bws_ms = acc[TB_BWS]; t.linesearch_ms = acc[TB_LS]; t.other_ms = acc[TB_OTHER];
        t.n_launches = b->launches;
        b->last = t;
        // TIME_PER_ITERATION buckets (MHPC_CompoundTypes.h:253-262; MultiPhaseDDP.cpp:163-170,184-192,206-208,228-235,242-253,259-267):
        // per inner iteration the backward sweep with its retries, the line search and the partials; the full forward sweep of
        // the outer iteration goes to its first inner iteration
        b->iter_times.clear();
        for (auto& e : b->evs) {
            if (e.iter < 0) continue;
            mhpc_iter_timing* it = nullptr;
            for (auto& x : b->iter_times) if (x.iter_AL == e.iter / 1000 && x.iter_DDP == e.iter % 1000) it = &x;
            if (!it) { b->iter_times.push_back(mhpc_iter_timing{e.iter / 1000, e.iter % 1000, 0, 0, 0, 0, 0}); it = &b->iter_times.back(); }
            float ms = 0;
            cudaEventElapsedTime(&ms, e.a, e.b);
            if (e.bucket == TB_ROLLOUT) it->time_fwd_ms += ms;
            else if (e.bucket == TB_BWS) it->time_bws_ms += ms;
            else if (e.bucket == TB_LS) it->time_fit_ms += ms;
            else if (e.bucket == TB_LIN) { if (e.iter % 1000 == 1 && it->time_fwd_ms > 0 && it->time_bws_ms == 0) it->time_fwd_ms += ms; else it->time_partial_ms += ms; }
            it->total_ms += ms;
        }
        b->cur_iter = -1;
        cudaEventDestroy(t0); cudaEventDestroy(t1);
    }
    return 0;
}

int mhpc_batch_get_results(mhpc_batch* b, mhpc_result* r) {
    if (need(b)) return -1;
    const Dev& d = b->d;
    std::vector<double> sd((size_t)SD_N * d.npad);
    std::vector<int> si((size_t)SI_N * d.npad);
    CK(cudaMemcpyAsync(sd.data(), d.sd, sd.size() * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    CK(cudaMemcpyAsync(si.data(), d.si, si.size() * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    for (int u = 0; u < b->n; ++u) {
        const int i = SL(b, u);
        r[u].cost = sd[(size_t)SD_COST * d.npad + i]; r[u].exp_cost_change = sd[(size_t)SD_DV * d.npad + i];
        r[u].viol = sd[(size_t)SD_VIOL * d.npad + i]; r[u].status = si[(size_t)SI_STATUS * d.npad + i];
        r[u].iter_AL = si[(size_t)SI_ITER_AL * d.npad + i]; r[u].n_records = si[(size_t)SI_N_REC * d.npad + i]; r[u].pad = 0;
    }
    return 0;
}
int mhpc_batch_get_scalars(mhpc_batch* b, double* cost, double* dv, double* viol) {
    if (need(b)) return -1;
    const Dev& d = b->d;
    if (cost && pull_per_instance(b, d.sd + (size_t)SD_COST * d.npad, cost)) return -1;
    if (dv && pull_per_instance(b, d.sd + (size_t)SD_DV * d.npad, dv)) return -1;
    if (viol && pull_per_instance(b, d.sd + (size_t)SD_VIOL * d.npad, viol)) return -1;
    return 0;
}
int mhpc_batch_get_records(mhpc_batch* b, int inst, mhpc_iter_record* out, int max_records) {
    if (need(b)) return -1;
    const Dev& d = b->d;
    if (inst < 0 || inst >= b->n) return fail("get_records: bad instance");
    inst = SL(b, inst);
    int nrec = 0;
    CK(cudaMemcpyAsync(&nrec, d.si + (size_t)SI_N_REC * d.npad + inst, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    nrec = std::min(nrec, std::min(max_records, d.max_rec));
    if (nrec <= 0) return 0;
    std::vector<int> ri((size_t)nrec * REC_I);
    std::vector<double> rd((size_t)nrec * REC_D);
    CK(cudaMemcpy2DAsync(ri.data(), sizeof(int), d.rec_i + inst, (size_t)d.npad * sizeof(int), sizeof(int), (size_t)nrec * REC_I, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaMemcpy2DAsync(rd.data(), sizeof(double), d.rec_d + inst, (size_t)d.npad * sizeof(double), sizeof(double), (size_t)nrec * REC_D, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    for (int r = 0; r < nrec; ++r) {
        out[r].iter_AL = ri[r * REC_I + 0]; out[r].iter_DDP = ri[r * REC_I + 1]; out[r].n_bws = ri[r * REC_I + 2]; out[r].n_fit = ri[r * REC_I + 3];
        out[r].reg = rd[r * REC_D + 0]; out[r].eps = rd[r * REC_D + 1]; out[r].cost = rd[r * REC_D + 2];
        out[r].exp_cost_change = rd[r * REC_D + 3]; out[r].viol = rd[r * REC_D + 4]; out[r].cost_prev = rd[r * REC_D + 5];
    }
    return nrec;
}
int mhpc_batch_n_phases(mhpc_batch* b, int inst) {
    if (!b || !b->have_problems || inst < 0 || inst >= b->n) return -1;
    return b->h_nph[SL(b, inst)];
}
int mhpc_batch_get_schedule(mhpc_batch* b, int inst, int* mode, int* nk, int* is_wb) {
    if (!b || !b->have_problems || inst < 0 || inst >= b->n) return fail("get_schedule: bad arguments");
    inst = SL(b, inst);
    const int npad = b->d.npad;
    for (int p = 0; p < b->h_nph[inst]; ++p) {
        if (mode) mode[p] = b->h_mode[(size_t)p * npad + inst];
        if (nk) nk[p] = b->h_N[(size_t)p * npad + inst];
        if (is_wb) is_wb[p] = p < b->h_nwb[inst];
    }
    return 0;
}
int mhpc_batch_get_phase_params(mhpc_batch* b, int inst, int phase, double* out9) {
    if (need(b)) return -1;
    if (inst < 0 || inst >= b->n) return fail("get_phase_params: bad instance");
    inst = SL(b, inst);
    if (phase < 0 || phase >= b->h_nph[inst]) return fail("get_phase_params: bad phase");
    const Dev& d = b->d;
    CK(cudaMemcpy2DAsync(out9, sizeof(double), d.par + par_idx(d, phase, 0, inst), (size_t)d.npad * sizeof(double), sizeof(double), NPAR,
                         cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    return 0;
}

// strided gather of `rows` doubles of one instance out of an [rows][npad] array
static int pull_rows(mhpc_batch* b, const double* base, int rows, int inst, double* out) {
    CK(cudaMemcpy2DAsync(out, sizeof(double), base + inst, (size_t)b->d.npad * sizeof(double), sizeof(double), rows, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    return 0;
}

int mhpc_batch_get_field(mhpc_batch* b, int inst, int phase, int field, double* out) {
    if (need(b)) return -1;
    b->gen--;   // a read-only call: the replay cache below stays valid across consecutive get_field calls
    const Dev& d = b->d;
    if (inst < 0 || inst >= b->n || phase < 0 || phase >= b->h_nph[SL(b, inst)]) return fail("get_field: bad instance/phase");
    const mhpc_problem& prob = b->problems[inst];   // caller's order
    inst = SL(b, inst);                             // slot from here on
    const int npad = d.npad;
    const bool wb = phase < b->h_nwb[inst];
    const int nx = wb ? 14 : 6;
    const int N = b->h_N[(size_t)phase * npad + inst], off = b->h_off[(size_t)phase * npad + inst], mode = b->h_mode[(size_t)phase * npad + inst];
    const double dt = wb ? (double)prob.dt_wb : (double)prob.dt_fb;
    auto state_field = [&](const double* base, int e0, int dim) -> int {
        if (!out) return dim;
        std::vector<double> tmp((size_t)N * ST);
        if (pull_rows(b, base + (size_t)off * ST * npad, N * ST, inst, tmp.data())) return -1;
        for (int k = 0; k < N; ++k) for (int e = 0; e < dim; ++e) out[(size_t)k * dim + e] = tmp[(size_t)k * ST + e0 + e];
        return dim;
    };
    int slot = 0;
    if (field <= MHPC_F_Y_ACT) {
        CK(cudaMemcpyAsync(&slot, d.si + (size_t)SI_ACC_SLOT * npad + inst, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
        CK(cudaStreamSynchronize(b->stream));
    }
    const double* act = d.st_act + (size_t)slot * d.KT * ST * npad;
    switch (field) {
        case MHPC_F_X_ACT: return state_field(act, 0, nx);
        case MHPC_F_U_ACT: return state_field(act, 14, 4);
        case MHPC_F_Y_ACT: return state_field(act, 18, 4);
        case MHPC_F_X_NOM: return state_field(d.st_nom, 0, nx);
        case MHPC_F_U_NOM: return state_field(d.st_nom, 14, 4);
        case MHPC_F_Y_NOM: return state_field(d.st_nom, 18, 4);
        case MHPC_F_K: case MHPC_F_DU: case MHPC_F_G: {
            const int e0 = field == MHPC_F_K ? GR_K : field == MHPC_F_DU ? GR_DU : GR_G;
            const int dim = field == MHPC_F_K ? 4 * nx : field == MHPC_F_DU ? 4 : nx;
            if (!out) return dim;
            std::vector<double> tmp((size_t)N * GREC);
            CK(cudaMemcpy2DAsync(tmp.data(), GREC * sizeof(double), d.gain + gain_idx(d, off, (size_t)inst), gain_knot_stride(d) * sizeof(double), GREC * sizeof(double), N,
                                 cudaMemcpyDeviceToHost, b->stream));
            CK(cudaStreamSynchronize(b->stream));
            for (int k = 0; k < N; ++k) for (int e = 0; e < dim; ++e) out[(size_t)k * dim + e] = tmp[(size_t)k * GREC + e0 + e];
            return dim;
        }
        case MHPC_F_X_REF: case MHPC_F_U_REF: case MHPC_F_Y_REF: {
            const int dim = field == MHPC_F_X_REF ? nx : 4;
            if (!out) return dim;
            std::vector<double> pos(N);
            if (pull_rows(b, d.pos_ref + (size_t)off * npad, N, inst, pos.data())) return -1;
            const double vel = prob.vel, hgt = prob.height;
            for (int k = 0; k < N; ++k) {   // ReferenceGen.h:53-91 (the device evaluates these on the fly from pos_ref)
                double* o = out + (size_t)k * dim;
                for (int e = 0; e < dim; ++e) o[e] = 0;
                const bool term = (k == N - 1);
                if (field == MHPC_F_X_REF) {
                    if (wb && !term) { o[0] = pos[k]; o[1] = hgt; for (int q = 0; q < 4; ++q) o[3 + q] = b->c.qbias[q]; o[7] = vel; }
                    else if (wb) { for (int e = 0; e < 14; ++e) o[e] = b->c.xterm[mode - 1][e]; o[0] = pos[k]; o[7] = vel; }
                    else { o[0] = pos[k]; o[1] = hgt; o[3] = vel; }
                } else if (field == MHPC_F_Y_REF) {
                    if (wb && !term) { o[1] = b->c.GRF; o[3] = b->c.GRF; }
                } else {
                    if (!wb && !term) { o[1] = b->c.GRF; o[3] = b->c.GRF; }
                }
            }
            return dim;
        }
        case MHPC_F_H: case MHPC_F_QX: case MHPC_F_QU: case MHPC_F_QXX: case MHPC_F_QUX: case MHPC_F_QUU: {
            // H, G exist on every knot; the Q-function on the step knots 0..N-2 (zero rows for the terminal knot, like the reference's zeroed struct)
            const int rows = (field == MHPC_F_QU || field == MHPC_F_QUX || field == MHPC_F_QUU) ? 4 : nx;
            const int cols = (field == MHPC_F_H || field == MHPC_F_QXX || field == MHPC_F_QUX) ? nx : field == MHPC_F_QUU ? 4 : 1;
            const int ld = rows == 4 ? 4 : 14;
            const int e0 = field == MHPC_F_H ? CG_H : field == MHPC_F_QX ? CG_QX : field == MHPC_F_QU ? CG_QU : field == MHPC_F_QXX ? CG_QXX : field == MHPC_F_QUX ? CG_QUX : CG_QUU;
            const int dim = rows * cols;
            if (!out) return dim;
            if (b->ctg_inst != inst || b->ctg_gen != b->gen) {
                if (replay_cost_to_go(b, inst)) return -1;
                b->ctg_inst = inst; b->ctg_gen = b->gen;
            }
            std::vector<double> tmp((size_t)N * CTGREC);
            CK(cudaMemcpyAsync(tmp.data(), d.ctg_dump + (size_t)off * CTGREC, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
            CK(cudaStreamSynchronize(b->stream));
            for (int k = 0; k < N; ++k) for (int cc = 0; cc < cols; ++cc) for (int r = 0; r < rows; ++r)
                out[(size_t)k * dim + r + rows * cc] = tmp[(size_t)k * CTGREC + e0 + r + ld * cc];
            return dim;
        }
        default: break;
    }
    // fields decoded from the AoSoA knot records
    std::vector<double> recs((size_t)N * REC);
    {
        CK(cudaMemcpyAsync(recs.data(), d.lin + lin_base(d, inst, off), recs.size() * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
        CK(cudaStreamSynchronize(b->stream));
    }
    auto R = [&](int k, int e) { return recs[(size_t)k * REC + e]; };
    const int r0 = (mode == 1) ? 2 : 0;
    const bool stance = (mode == 1 || mode == 3);
    int dim = 0;
    switch (field) {
        case MHPC_F_A: dim = nx * nx; break; case MHPC_F_B: dim = nx * 4; break; case MHPC_F_C: dim = 4 * nx; break; case MHPC_F_D: dim = 16; break;
        case MHPC_F_LX: dim = nx; break; case MHPC_F_LU: dim = 4; break; case MHPC_F_LY: dim = 4; break;
        case MHPC_F_LXX: dim = nx * nx; break; case MHPC_F_LUU: dim = 16; break; case MHPC_F_LYY: dim = 16; break;
        case MHPC_F_PHIX: dim = nx; break; case MHPC_F_PHIXX: dim = nx * nx; break; case MHPC_F_PX: dim = nx * nx; break;
        default: return fail("get_field: unknown field");
    }
    if (!out) return dim;
    if (field == MHPC_F_PHIX || field == MHPC_F_PHIXX || field == MHPC_F_PX) {
        const int k = N - 1;
        for (int e = 0; e < dim; ++e) out[e] = 0;
        if (field == MHPC_F_PHIX) { for (int e = 0; e < nx; ++e) out[e] = R(k, wb ? R_PHIX + e : R_RPHIX + e); }
        else if (field == MHPC_F_PHIXX) {
            for (int e = 0; e < nx; ++e) out[e + nx * e] = wb ? R(k, R_PHIXXD + e) : b->c.rbQf[e];
            if (wb) for (int cb = 1; cb <= 6; ++cb) for (int a = 1; a <= 6; ++a) out[a + 14 * cb] += R(k, R_PHIXXB + (a - 1) + 6 * (cb - 1));
        } else {
            for (int e = 0; e < nx; ++e) out[e + nx * e] = 1.0;
            if (wb) {
                for (int e = 7; e < 14; ++e) out[e + 14 * e] = 0.0;
                for (int cc = 0; cc < 12; ++cc) for (int s = 0; s < 7; ++s) out[(7 + s) + 14 * (2 + cc)] = R(k, R_AB(s, cc));
            }
        }
        return dim;
    }
    for (int k = 0; k < N; ++k) {
        double* o = out + (size_t)k * dim;
        for (int e = 0; e < dim; ++e) o[e] = 0;
        if (k == N - 1) continue;   // derivatives exist for knots 0..N-2 only (SinglePhase.cpp:73)
        if (wb) {
            switch (field) {
                case MHPC_F_A:
                    for (int e = 0; e < 7; ++e) { o[e + 14 * e] = 1.0; o[e + 14 * (7 + e)] = 1.0 * dt; }
                    for (int cc = 0; cc < 12; ++cc) for (int s = 0; s < 7; ++s) o[(7 + s) + 14 * (2 + cc)] = R(k, R_AB(s, cc));
                    break;
                case MHPC_F_B: for (int cc = 0; cc < 4; ++cc) for (int s = 0; s < 7; ++s) o[(7 + s) + 14 * cc] = R(k, R_BB(s, cc)); break;
                case MHPC_F_C: if (stance) for (int cc = 0; cc < 12; ++cc) for (int t = 0; t < 2; ++t) o[(r0 + t) + 4 * (2 + cc)] = R(k, R_CB(t, cc)); break;
                case MHPC_F_D: if (stance) for (int cc = 0; cc < 4; ++cc) for (int t = 0; t < 2; ++t) o[(r0 + t) + 4 * cc] = R(k, R_DB(t, cc)); break;
                case MHPC_F_LX: for (int e = 0; e < 14; ++e) o[e] = R(k, R_LX + e); break;
                case MHPC_F_LU: for (int e = 0; e < 4; ++e) o[e] = R(k, R_LU + e); break;
                case MHPC_F_LY: if (stance) { o[r0] = R(k, R_LY); o[r0 + 1] = R(k, R_LY + 1); } break;
                case MHPC_F_LXX: for (int e = 0; e < 14; ++e) o[e + 14 * e] = (e >= 3 && e < 7) ? R(k, R_LXXQ + e - 3) : (2 * dt) * b->c.wbQ[e]; break;
                case MHPC_F_LUU: for (int e = 0; e < 4; ++e) o[e + 4 * e] = R(k, R_LUU + e); break;
                case MHPC_F_LYY:
                    if (stance) { o[r0 + 4 * r0] = R(k, R_LYY); o[r0 + 4 * (r0 + 1)] = R(k, R_LYY + 1); o[(r0 + 1) + 4 * r0] = R(k, R_LYY + 1); o[(r0 + 1) + 4 * (r0 + 1)] = R(k, R_LYY + 2); }
                    break;
            }
        } else {
            const int rows[8] = {3, 5, 4, 5, 3, 5, 4, 5}, cols[8] = {0, 0, 1, 1, 2, 2, 3, 3};
            switch (field) {
                case MHPC_F_A:
                    for (int e = 0; e < 6; ++e) o[e + 6 * e] = 1.0;
                    for (int e = 0; e < 3; ++e) o[e + 6 * (3 + e)] = 1.0 * dt;
                    o[5 + 6 * 0] = R(k, R_RA); o[5 + 6 * 1] = R(k, R_RA + 1);
                    break;
                case MHPC_F_B: for (int q = 0; q < 8; ++q) o[rows[q] + 6 * cols[q]] = R(k, R_RB + q); break;
                case MHPC_F_LX: for (int e = 0; e < 6; ++e) o[e] = R(k, R_RLX + e); break;
                case MHPC_F_LU: for (int e = 0; e < 4; ++e) o[e] = R(k, R_RLU + e); break;
                case MHPC_F_LXX: for (int e = 0; e < 6; ++e) o[e + 6 * e] = (2 * dt) * b->c.rbQ[e]; break;
                case MHPC_F_LUU: for (int e = 0; e < 4; ++e) o[e + 4 * e] = (2 * dt) * b->c.rbR[mode - 1][e]; break;
                default: break;
            }
        }
    }
    return dim;
}

// execution bag (MHPCLocomotion.cpp:178-194): nominal x,u,y and K,du of the leading knots, instance-major
__global__ void k_gather_exec(Dev d, const int* __restrict__ user_of, int kn, int exec_phases, double* st_out, double* gain_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    const int u = user_of ? user_of[i] : i;   // output row: the caller's instance order
    const int k = blockIdx.y;   // k-th knot of the execution horizon -> knot slot (phase blocks of KSTRIDE)
    int kk = -1, c = k;
    const int np_exec = min(exec_phases, d.n_wb[i]);   // whole-body phases only (ms_exec / CTG_exec are WB-typed, MHPCLocomotion.cpp:178-194)
    for (int p = 0; p < np_exec; ++p) {
        const int N = d.ph_N[ph_idx(d, p, i)];
        if (c < N) { kk = p * KSTRIDE + c; break; }
        c -= N;
    }
    if (st_out) for (int e = 0; e < ST; ++e) st_out[((size_t)u * kn + k) * ST + e] = kk >= 0 ? d.st_nom[((size_t)kk * ST + e) * d.npad + i] : 0.0;
    if (gain_out) {
        const double* gr = d.gain + gain_idx(d, kk >= 0 ? kk : 0, (size_t)i);     // K | du are the first 60 entries of the record
        for (int e = 0; e < 60; ++e) gain_out[((size_t)u * kn + k) * 60 + e] = kk >= 0 ? gr[e] : 0.0;
    }
}
__global__ void k_results_pack(Dev d, const int* __restrict__ user_of, mhpc_result* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    mhpc_result r;
    r.cost = d.sd[(size_t)SD_COST * d.npad + i]; r.exp_cost_change = d.sd[(size_t)SD_DV * d.npad + i]; r.viol = d.sd[(size_t)SD_VIOL * d.npad + i];
    r.status = d.si[(size_t)SI_STATUS * d.npad + i]; r.iter_AL = d.si[(size_t)SI_ITER_AL * d.npad + i]; r.n_records = d.si[(size_t)SI_N_REC * d.npad + i]; r.pad = 0;
    out[user_of ? user_of[i] : i] = r;
}
__global__ void k_results_table(Dev d, const int* __restrict__ user_of, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.npad) return;
    const bool v = i < d.n;
    const int u = (v && user_of) ? user_of[i] : i;   // padding columns stay in place
    out[0 * (size_t)d.npad + u] = v ? d.sd[(size_t)SD_COST * d.npad + i] : 0.0;
    out[1 * (size_t)d.npad + u] = v ? d.sd[(size_t)SD_DV * d.npad + i] : 0.0;
    out[2 * (size_t)d.npad + u] = v ? d.sd[(size_t)SD_VIOL * d.npad + i] : 0.0;
    out[3 * (size_t)d.npad + u] = v ? (double)d.si[(size_t)SI_STATUS * d.npad + i] : -1.0;
    out[4 * (size_t)d.npad + u] = v ? (double)d.si[(size_t)SI_ITER_AL * d.npad + i] : 0.0;
    out[5 * (size_t)d.npad + u] = v ? (double)d.si[(size_t)SI_N_REC * d.npad + i] : 0.0;
}
int mhpc_batch_results_to_device(mhpc_batch* b, double* dev_out, int capacity) {
    if (need(b)) return -1;
    if (capacity < 6 * b->d.npad) return fail("results_to_device: buffer too small (need 6*npad doubles)");
    k_results_table<<<(b->d.npad + 127) / 128, 128, 0, b->stream>>>(b->d, b->permuted ? b->d_user_of : nullptr, dev_out);
    CK(cudaGetLastError());
    return b->d.npad;
}

int mhpc_batch_solve_host(mhpc_batch* b, const mhpc_problem* problems, mhpc_result* results, int exec_phases,
                          double* exec_state, double* exec_gain, int exec_knots) {
    if (!b) return fail("null handle");
    if (mhpc_batch_set_problems(b, problems)) return -1;
    launch_reset_option_state(b->d, b->opt, b->stream);   // behaves like a freshly constructed MHPCLocomotion per call
    if (mhpc_batch_initialization(b)) return -1;
    if (mhpc_batch_solve(b, 0)) return -1;
    const Dev& d = b->d;
    const int kn = (exec_phases > 0 && exec_knots > 0) ? std::min(exec_knots, d.KT) : 0;
    // staging buffers (device): packed results + instance-major execution bag, one contiguous D2H each
    const size_t need_res = (size_t)b->n * sizeof(mhpc_result);
    const size_t need_st = exec_state ? (size_t)b->n * kn * ST * sizeof(double) : 0, need_gn = exec_gain ? (size_t)b->n * kn * 60 * sizeof(double) : 0;
    if (b->stage_bytes < need_res + need_st + need_gn) {
        if (b->stage) cudaFree(b->stage);
        CK(cudaMalloc(&b->stage, need_res + need_st + need_gn));
        b->stage_bytes = need_res + need_st + need_gn;
    }
    char* base = (char*)b->stage;
    mhpc_result* dres = (mhpc_result*)base;
    double* dst = exec_state && kn ? (double*)(base + need_res) : nullptr;
    double* dgn = exec_gain && kn ? (double*)(base + need_res + need_st) : nullptr;
    const int* uo = b->permuted ? b->d_user_of : nullptr;
    if (results) k_results_pack<<<(b->n + 127) / 128, 128, 0, b->stream>>>(d, uo, dres);
    if (kn && (dst || dgn)) { dim3 g((b->n + 127) / 128, kn); k_gather_exec<<<g, 128, 0, b->stream>>>(d, uo, kn, exec_phases, dst, dgn); }
    if (results) CK(cudaMemcpyAsync(results, dres, need_res, cudaMemcpyDeviceToHost, b->stream));
    if (dst) CK(cudaMemcpyAsync(exec_state, dst, need_st, cudaMemcpyDeviceToHost, b->stream));
    if (dgn) CK(cudaMemcpyAsync(exec_gain, dgn, need_gn, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    CK(cudaGetLastError());
    return 0;
}

int mhpc_eval_function(const char* name, int n, int device, const double* in0, const double* in1, const double* in2,
                       const double* in3, double* o0, double* o1, double* o2, double* o3) {
    if (!name || n <= 0 || !in0 || !o0) return fail("mhpc_eval_function: bad arguments");
    {   // the five Jacobian functions (CasadiGen/header/Dyn_BS_par.h ..., adaptor CasadiGen.cpp:4-79): the bodies the
        // linearisation kernels run, evaluated standalone: (x[14], u[4]) -> Ac 14x14, Bc 14x4, C 4x14, D 4x4; x -> Px 14x14
        typedef void (*pe_fn)(int, size_t, const double*, double*, double*, double*, double*, cudaStream_t);
        static const struct { const char* nm; pe_fn fn; bool dyn; } par[5] = {
            {"Dyn_BS_par", launch_par_eval_BS, true}, {"Dyn_FS_par", launch_par_eval_FS, true}, {"Dyn_FL_par", launch_par_eval_FL, true},
            {"Imp_F_par", launch_par_eval_IMPF, false}, {"Imp_B_par", launch_par_eval_IMPB, false}};
        for (const auto& pf : par) {
            if (std::strcmp(name, pf.nm)) continue;
            if (pf.dyn && (!in1 || !o1 || !o2 || !o3)) return fail("mhpc_eval_function: Dyn_*_par takes (x, u) and returns (Ac, Bc, C, D)");
            CK(cudaSetDevice(device));
            const size_t es = ((size_t)n + 31) / 32 * 32;
            std::vector<double> soa(18 * es, 0.0);
            for (int t = 0; t < n; ++t) {
                for (int e = 0; e < 14; ++e) soa[(size_t)e * es + t] = in0[(size_t)t * 14 + e];
                if (pf.dyn) for (int e = 0; e < 4; ++e) soa[(size_t)(14 + e) * es + t] = in1[(size_t)t * 4 + e];
            }
            const size_t osz[4] = {196, pf.dyn ? 56u : 0u, pf.dyn ? 56u : 0u, pf.dyn ? 16u : 0u};
            double* hout[4] = {o0, o1, o2, o3};
            double *din = nullptr, *dout[4] = {nullptr, nullptr, nullptr, nullptr};
            CK(cudaMalloc(&din, soa.size() * sizeof(double)));
            CK(cudaMemcpy(din, soa.data(), soa.size() * sizeof(double), cudaMemcpyHostToDevice));
            for (int q = 0; q < 4; ++q) {
                if (!osz[q]) continue;
                CK(cudaMalloc(&dout[q], (size_t)n * osz[q] * sizeof(double)));
                CK(cudaMemset(dout[q], 0, (size_t)n * osz[q] * sizeof(double)));   // structural zeros (CasadiGen.cpp:58-66 writes non-zeros only)
            }
            pf.fn(n, es, din, dout[0], dout[1], dout[2], dout[3], 0);
            CK(cudaGetLastError());
            CK(cudaDeviceSynchronize());
            for (int q = 0; q < 4; ++q) {
                if (!osz[q]) continue;
                CK(cudaMemcpy(hout[q], dout[q], (size_t)n * osz[q] * sizeof(double), cudaMemcpyDeviceToHost));
                cudaFree(dout[q]);
            }
            cudaFree(din);
            return 0;
        }
    }
    const int fn = eval_fn_id(name);
    if (fn < 0) return fail(std::string("mhpc_eval_function: unknown function '") + name + "'");
    CK(cudaSetDevice(device));
    const int isz[16][4] = {{14, 4, 0, 0}, {14, 4, 0, 0}, {14, 4, 0, 0}, {14, 0, 0, 0}, {14, 0, 0, 0}, {6, 4, 4, 2}, {6, 4, 4, 2}, {14, 0, 0, 0}, {14, 0, 0, 0}, {14, 0, 0, 0}, {14, 0, 0, 0},
                            {14, 2, 0, 0}, {14, 2, 0, 0}, {14, 2, 0, 0}, {14, 2, 0, 0}, {14, 2, 0, 0}};
    const int osz[16][4] = {{14, 4, 0, 0}, {14, 4, 0, 0}, {14, 4, 0, 0}, {14, 4, 0, 0}, {14, 4, 0, 0}, {6, 0, 0, 0}, {36, 24, 0, 0}, {1, 14, 196, 0}, {1, 14, 196, 0}, {14, 0, 0, 0}, {14, 0, 0, 0},
                            {14, 14, 0, 0}, {14, 14, 0, 0}, {14, 14, 0, 0}, {14, 14, 0, 0}, {14, 14, 0, 0}};
    const double* hin[4] = {in0, in1, in2, in3};
    double* hout[4] = {o0, o1, o2, o3};
    double* din[4] = {nullptr, nullptr, nullptr, nullptr};
    double* dout[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int q = 0; q < 4; ++q) {
        if (isz[fn][q]) { CK(cudaMalloc(&din[q], (size_t)n * isz[fn][q] * sizeof(double))); CK(cudaMemcpy(din[q], hin[q], (size_t)n * isz[fn][q] * sizeof(double), cudaMemcpyHostToDevice)); }
        if (osz[fn][q]) CK(cudaMalloc(&dout[q], (size_t)n * osz[fn][q] * sizeof(double)));
    }
    launch_eval(fn, n, din[0], din[1], din[2], din[3], dout[0], dout[1], dout[2], dout[3], 0);
    CK(cudaDeviceSynchronize());
    for (int q = 0; q < 4; ++q) {
        if (osz[fn][q]) CK(cudaMemcpy(hout[q], dout[q], (size_t)n * osz[fn][q] * sizeof(double), cudaMemcpyDeviceToHost));
        if (din[q]) cudaFree(din[q]);
        if (dout[q]) cudaFree(dout[q]);
    }
    return 0;
}

int mhpc_batch_enable_bws_trace(mhpc_batch* b, int on) {
    b->want_bws_trace = on != 0;
    if (!b->have_problems) return 0;
    CK(cudaSetDevice(b->device));
    Dev& d = b->d;
    if (!on) { d.bws_trace = nullptr; return 0; }
    const size_t cnt = (size_t)d.max_rec * BWS_TRACE_ATT * d.npad;
    if (b->bws_trace_cap < cnt) {
        CK(cudaStreamSynchronize(b->stream));
        if (b->bws_trace_buf) cudaFree(b->bws_trace_buf);
    if (b->h_unres) cudaFreeHost(b->h_unres);
        CK(cudaMalloc(&b->bws_trace_buf, cnt * sizeof(int)));
        b->bws_trace_cap = cnt;
    }
    CK(cudaMemsetAsync(b->bws_trace_buf, 0, cnt * sizeof(int), b->stream));
    d.bws_trace = b->bws_trace_buf;
    return 0;
}
int mhpc_batch_get_bws_trace(mhpc_batch* b, int* out) {
    if (need(b)) return -1;
    const Dev& d = b->d;
    if (!d.bws_trace) return fail("get_bws_trace: call mhpc_batch_enable_bws_trace first");
    const size_t rows = (size_t)d.max_rec * BWS_TRACE_ATT;
    std::vector<int> tmp(rows * d.npad);
    CK(cudaMemcpyAsync(tmp.data(), d.bws_trace, tmp.size() * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    for (int u = 0; u < b->n; ++u) {
        const int i = SL(b, u);
        for (size_t r = 0; r < rows; ++r) out[(size_t)u * rows + r] = tmp[r * d.npad + i];
    }
    return (int)rows;
}
// Control-smoothness penalty: the reference carries the switch (HSDDP_OPTION::smooth_active, MHPC_CompoundTypes.h:211) and the
// weight (AL_REB_PARAMETER::eps_smooth, :224), but SinglePhase::update_running_cost_with_smooth is an empty body that nothing
// calls (SinglePhase.cpp:252-254).  Both values are accepted and kept; like there, they change no result.
int mhpc_batch_set_eps_smooth(mhpc_batch* b, double eps_smooth) { if (!b) return fail("null handle"); b->eps_smooth = eps_smooth; return 0; }
double mhpc_batch_get_eps_smooth(mhpc_batch* b) { return b ? b->eps_smooth : 0.0; }
int mhpc_batch_enable_timing(mhpc_batch* b, int on) { b->timing = on != 0; return 0; }
int mhpc_batch_get_timing(mhpc_batch* b, mhpc_timing* t) { *t = b->last; return 0; }
int mhpc_batch_get_iteration_timing(mhpc_batch* b, mhpc_iter_timing* out, int max_records) {
    const int n = std::min((int)b->iter_times.size(), max_records);
    for (int k = 0; k < n; ++k) out[k] = b->iter_times[k];
    return n;
}

int mhpc_batch_time_kernel(mhpc_batch* b, int which, int reps, float* ms_per_launch) {
    if (need(b)) return -1;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const bool saved = b->timing;
    b->timing = false;
    auto run = [&]() -> int {
        if (which == 0) {
            return linearize(b, MASK_ALL);
        } else if (which >= 10 && which < 10 + CLS_N) {   // one knot class of the linearisation only
            const int cls = which - 10;
            typedef void (*lin_fn)(const Dev&, const Consts&, const Opt&, int, const int*, int, cudaStream_t);
            const lin_fn fns[CLS_N] = {launch_lin_BS, launch_lin_FS, launch_lin_FL, launch_lin_IMPF, launch_lin_IMPB, launch_lin_LIGHT};
            fns[cls](b->d, b->c, b->opt, MASK_ALL, b->knot_list[cls], b->knot_cnt[cls], b->stream);
        } else if (which == 1) {
            (b->bws_kind == 2 ? launch_bwsT : b->bws_kind == 1 ? launch_bws16 : launch_bws)(b->d, b->c, b->opt, MASK_ALL, 1, 0, b->stream);
        } else {
            launch_rollout(b->d, b->c, b->opt, MASK_ALL, std::min(b->d.W, b->eps.n), 0, b->eps, b->stream);
        }
        return 0;
    };
    if (run()) return -1;   // warm-up
    CK(cudaStreamSynchronize(b->stream));
    CK(cudaEventRecord(e0, b->stream));
    for (int r = 0; r < reps; ++r) if (run()) return -1;
    CK(cudaEventRecord(e1, b->stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    *ms_per_launch = ms / reps;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    b->timing = saved;
    CK(cudaGetLastError());
    return 0;
}

int mhpc_measure_fp64_peak(int device, double* tflops) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev) return fail("no CUDA device");
    *tflops = measure_fp64_peak_tflops(device);
    return 0;
}

}  // extern "C"
