// mhpc_b200.hpp - header-only C++ host mirror of the reference's surface for the batched B200 path.
//
// Same type names, field names, defaults and method names as the reference so that host code written against
// NaCl-1374/MHPC_Minimal_Env keeps compiling with a batch dimension added:
//   HSDDP_OPTION<T>        Common/header/MHPC_CompoundTypes.h:196-212
//   USRCMD, MHPCUserParameters                         :237-251
//   Gait                   Common/header/Gait.h:13-78
//   BatchedMHPCLocomotion<T>  =  MHPCLocomotion<T> (Controller/MHPCLocomotion/MHPCLocomotion.h:13-83) over N instances,
//                             exposing the MultiPhaseDDP<T> entry points (HSDDPSolver/header/MultiPhaseDDP.h:17-63)
// Everything forwards to the C ABI in mhpc_b200.h (libmhpc_b200.so); there is no arithmetic in this header.
#ifndef MHPC_B200_HPP
#define MHPC_B200_HPP
#include <cassert>
#include <algorithm>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "mhpc_b200.h"

template <typename T>
struct HSDDP_OPTION {
    T alpha = 0.1;               // line search udpate paramer
    T gamma = 0.01;              // scale the expected cost reduction
    T update_penalty = 8;        // penalty update parameter
    T update_relax = 0.1;        // relaxation parameter udpate parameter
    T update_regularization = 2; // regularization parameter update parameter
    T update_ReB = 7;            // update barrier function weighting
    T max_DDP_iter = 3;          // maximum inner loop iteration
    T max_AL_iter = 2;           // maximum outer loop iteration
    T DDP_thresh = 1e-03;        // inner loop convergence threshhold
    T AL_thresh = 1e-03;         // outer loop convergence threshhold
    bool AL_active = 1;          // activate terminal constraint
    bool ReB_active = 1;         // activate path constraint
    bool smooth_active = 0;      // activate control smoothness penalization
};

struct USRCMD { float vel, height, roll, pitch, yaw; };

struct MHPCUserParameters {
    int n_wbphase = 4;
    int n_fbphase = 4;
    float dt_wb = .001f;
    float dt_fb = .001f;
    int cmode = 1;
    float groundH = -0.404f;
    USRCMD* usrcmd = nullptr;
};

struct TIME_PER_ITERATION {   // MHPC_CompoundTypes.h:253-262 (times in ms from CUDA events, whole batch)
    int DDP_iter = 0;
    int n_bws = 0;
    double time_bws = 0;
    int n_fit = 0;
    double time_fit = 0;
    double time_partial = 0;
};

enum class GaitType2D { STAND, BOUND, PRONK };

class Gait {   // Gait.h:13-78
public:
    std::vector<int> _gait_mode{1, 2, 3, 4};
    std::vector<float> _gait_timing{0.08f, 0.1f, 0.08f, 0.1f};
    std::string _gait_name = "BOUND";
    Gait() {}
    explicit Gait(GaitType2D gait) { if (gait != GaitType2D::BOUND) _gait_timing = {0.08f, 0.08f, 0.08f, 0.08f}; }
    int get_next_mode(int current_mode) const {
        for (size_t i = 0; i < _gait_mode.size(); ++i)
            if (current_mode == _gait_mode[i]) return _gait_mode[(i + 1) % _gait_mode.size()];
        return current_mode;
    }
    std::vector<int> get_mode_seq(int current_mode, int num_phases) const {
        std::vector<int> s(num_phases, current_mode);
        for (int p = 0; p + 1 < num_phases; ++p) s[p + 1] = get_next_mode(s[p]);
        return s;
    }
    std::vector<float> get_timings(const std::vector<int>& mode_seq) const {
        std::vector<float> t(mode_seq.size());
        for (size_t i = 0; i < mode_seq.size(); ++i) t[i] = _gait_timing[mode_seq[i] - 1];
        return t;
    }
};

template <typename T>
class BatchedMHPCLocomotion {
    static_assert(sizeof(T) == sizeof(double), "the device path is FP64 (the reference instantiates <double> only)");
public:
    // MHPCLocomotion(MHPCUserParameters*, Gait*, HSDDP_OPTION<T>)  (MHPCLocomotion.cpp:8-43), for n_instances on `device`.
    BatchedMHPCLocomotion(int n_instances, MHPCUserParameters* params, Gait* gait, HSDDP_OPTION<T> option_in, int device = 0, int ls_width = 8)
        : _n_instances(n_instances), _n_phases(params->n_wbphase + params->n_fbphase), _option(option_in) {
        mhpc_hsddp_option o = to_c(option_in);
        ck(mhpc_batch_create(n_instances, &o, device, ls_width, &_h));
        mhpc_problem p;
        mhpc_default_problem(&p);   // x0 of MHPCLocomotion.cpp:37-39
        p.n_wbphase = params->n_wbphase; p.n_fbphase = params->n_fbphase; p.dt_wb = params->dt_wb; p.dt_fb = params->dt_fb;
        p.cmode = params->cmode; p.groundH = params->groundH;
        if (params->usrcmd) { p.vel = params->usrcmd->vel; p.height = params->usrcmd->height; p.roll = params->usrcmd->roll; p.pitch = params->usrcmd->pitch; p.yaw = params->usrcmd->yaw; }
        for (int i = 0; i < 4; ++i) { p.gait_mode[i] = gait->_gait_mode[i]; p.gait_timing[i] = gait->_gait_timing[i]; }
        _problems.assign(n_instances, p);
    }
    ~BatchedMHPCLocomotion() { mhpc_batch_destroy(_h); }
    BatchedMHPCLocomotion(const BatchedMHPCLocomotion&) = delete;
    BatchedMHPCLocomotion& operator=(const BatchedMHPCLocomotion&) = delete;

    // per-instance problem data (the "batched entry point" the controller gains)
    mhpc_problem& problem(int i) { _dirty = true; return _problems[i]; }
    void set_initial_condition(int i, const T* x0) { for (int e = 0; e < 14; ++e) _problems[i].x0[e] = x0[e]; _dirty = true; }   // MultiPhaseDDP.h:24-25
    void set_option(HSDDP_OPTION<T>& option) { _option = option; mhpc_hsddp_option o = to_c(option); ck(mhpc_batch_set_option(_h, &o)); }   // MultiPhaseDDP.h:22

    void initialization() { upload(); ck(mhpc_batch_initialization(_h)); }                 // MHPCLocomotion.cpp:47-53
    // MHPCLocomotion::update_problem (MHPCLocomotion.cpp:107-158): receding-horizon shift by one phase; initial states set
    // with set_initial_condition() since the last upload are sent along (update_problem itself leaves _x0 untouched)
    void update_problem() {
        ck(mhpc_batch_update_problem(_h));
        for (auto& p : _problems) p.cmode = next_mode(p);
        if (_dirty) {
            std::vector<double> x0((size_t)_n_instances * 14);
            for (int i = 0; i < _n_instances; ++i) for (int e = 0; e < 14; ++e) x0[(size_t)i * 14 + e] = _problems[i].x0[e];
            ck(mhpc_batch_set_initial_condition(_h, x0.data()));
            _dirty = false;
        }
    }
    void warmstart() { initialization(); }                                                  // :200-215 (part of initialization on the device)
    void solve_mhpc() { ck(mhpc_batch_solve(_h, 0)); refresh(); }                           // :167-195
    void solve() { solve_mhpc(); }                                                          // MultiPhaseDDP.cpp:154-289
    void forward_sweep(T eps) { ck(mhpc_batch_forward_sweep(_h, eps)); refresh(); }         // :28-48
    void forward_sweep_dynamics_only(T eps) { ck(mhpc_batch_forward_sweep_dynamics_only(_h, eps)); refresh(); }   // :56-76
    void forward_sweep_partials_only() { ck(mhpc_batch_forward_sweep_partials_only(_h)); }  // :84-91
    std::vector<int> backward_sweep(T regularization) {                                     // :100-127 (bool per instance)
        std::vector<int> ok(_n_instances);
        ck(mhpc_batch_backward_sweep(_h, regularization, ok.data()));
        refresh();
        return ok;
    }
    std::vector<int> forward_iteration() {                                                  // :130-151
        std::vector<int> t(_n_instances);
        ck(mhpc_batch_forward_iteration(_h, t.data()));
        refresh();
        return t;
    }
    void update_nominal_trajectory() { ck(mhpc_batch_update_nominal_trajectory(_h)); }      // :389-399
    void update_AL_ReB_param() { ck(mhpc_batch_update_AL_ReB_param(_h)); }                  // :402-413

    // result fields of MultiPhaseDDP (MultiPhaseDDP.h:49-58), one entry per instance
    std::vector<T> _actual_cost, _exp_cost_change, _tconstr_violation;
    int _n_instances, _n_phases;
    HSDDP_OPTION<T> _option;

    std::vector<mhpc_result> results() { std::vector<mhpc_result> r(_n_instances); ck(mhpc_batch_get_results(_h, r.data())); return r; }
    std::vector<mhpc_iter_record> records(int i) {
        std::vector<mhpc_iter_record> r(256);
        int n = mhpc_batch_get_records(_h, i, r.data(), 256);
        ck(n);
        r.resize(n);
        return r;
    }
    // MultiPhaseDDP.cpp:163-170,206-253 (TIME_BENCHMARK): one record per executed inner iteration of instance i.  n_bws / n_fit
    // are the instance's own counts; the times are those of the batch (CUDA events, ms; call enable_timing() before the
    // solve), booked as the reference books them: time_partial of an iteration is the time of the sweep that produced the
    // derivatives it used - the full forward sweep for the first inner iteration of an outer iteration, the partials-only
    // sweep of the previous inner iteration otherwise.
    void enable_timing(bool on = true) { ck(mhpc_batch_enable_timing(_h, on ? 1 : 0)); }
    std::vector<TIME_PER_ITERATION> time_ddp(int i) {
        std::vector<TIME_PER_ITERATION> out;
        std::vector<mhpc_iter_timing> tm(256);
        const int nt = mhpc_batch_get_iteration_timing(_h, tm.data(), 256);
        int abs_it = 1;
        for (const mhpc_iter_record& r : records(i)) {
            TIME_PER_ITERATION t;
            t.DDP_iter = abs_it++; t.n_bws = r.n_bws; t.n_fit = r.n_fit;
            for (int k = 0; k < nt; ++k) {
                if (tm[k].iter_AL != r.iter_AL) continue;
                if (tm[k].iter_DDP == r.iter_DDP) { t.time_bws = tm[k].time_bws_ms; t.time_fit = tm[k].time_fit_ms; if (r.iter_DDP == 1) t.time_partial = tm[k].time_fwd_ms; }
                if (r.iter_DDP > 1 && tm[k].iter_DDP == r.iter_DDP - 1) t.time_partial = tm[k].time_partial_ms;
            }
            out.push_back(t);
        }
        return out;
    }
    // [N_knots x dim] row-major copy of a per-knot field of one phase (matrices column-major inside a row)
    std::vector<double> field(int i, int phase, mhpc_field f, int* dim_out = nullptr, int* knots_out = nullptr) {
        int np = mhpc_batch_n_phases(_h, i);
        ck(np);
        std::vector<int> mode(np), N(np), wb(np);
        ck(mhpc_batch_get_schedule(_h, i, mode.data(), N.data(), wb.data()));
        int dim = mhpc_batch_get_field(_h, i, phase, f, nullptr);
        ck(dim);
        const bool per_phase = (f == MHPC_F_PHIX || f == MHPC_F_PHIXX || f == MHPC_F_PX);
        std::vector<double> out((size_t)dim * (per_phase ? 1 : N[phase]));
        ck(mhpc_batch_get_field(_h, i, phase, f, out.data()));
        if (dim_out) *dim_out = dim;
        if (knots_out) *knots_out = per_phase ? 1 : N[phase];
        return out;
    }
    // print_debugInfo (MHPCLocomotion.cpp:293-380) of instance i, file for file and row for row:
    //   state.txt     nominal x of every knot, whole-body phases then rigid-body phases
    //   control.txt   nominal u;  gradient.txt  value gradient G
    //   cost.txt      per phase the running-cost gradients lx of knots 0..N-2, then the terminal gradient Phix
    // in Eigen's default stream format (6 significant digits, coefficients right-aligned to the longest one of the row).
    // The reference sizes the rigid-body loops of control / gradient / cost with N_TIMESTEPS[i + 2] instead of
    // N_TIMESTEPS[i + n_wbphase] (:330,350,371); the row counts follow it (rows past a phase's own knots are the zeros of
    // memory_reset).  lx and Phix are those of forward_sweep_partials_only() at the final nominal trajectory, which is what
    // the reference's arrays hold when the inner loop ran to max_DDP_iter; it is re-run here because solve() skips that
    // last, otherwise dead, sweep.  (When the reference's inner loop stops on DDP_thresh instead, its arrays keep the
    // derivatives of the previous linearisation point plus what the line-search trials added to Phix - SURVEY A.2-1; that
    // left-over is not reproduced.)
    void print_debugInfo(int i, const std::string& prefix = "", int precision = 6) {
        ck(mhpc_batch_forward_sweep_partials_only(_h));
        std::ofstream state(prefix + "state.txt"), control(prefix + "control.txt"), gradient(prefix + "gradient.txt"), cost(prefix + "cost.txt");
        for (std::ofstream* f : {&state, &control, &gradient, &cost}) f->precision(precision);
        const int np = mhpc_batch_n_phases(_h, i);
        ck(np);
        std::vector<int> mode(np), N(np), wb(np);
        ck(mhpc_batch_get_schedule(_h, i, mode.data(), N.data(), wb.data()));
        int nwb = 0;
        for (int p = 0; p < np; ++p) nwb += wb[p];
        for (int p = 0; p < np; ++p) {
            const int rows_ref = wb[p] ? N[p] : N[std::min(p - nwb + 2, np - 1)];   // the reference's loop bound
            dump(state, i, p, MHPC_F_X_NOM, N[p]);
            dump(control, i, p, MHPC_F_U_NOM, rows_ref);
            dump(gradient, i, p, MHPC_F_G, rows_ref);
            dump(cost, i, p, MHPC_F_LX, rows_ref - 1);
            dump(cost, i, p, MHPC_F_PHIX, 1);
        }
    }
    mhpc_batch* handle() { return _h; }

private:
    static mhpc_hsddp_option to_c(const HSDDP_OPTION<T>& s) {
        mhpc_hsddp_option o;
        o.alpha = s.alpha; o.gamma = s.gamma; o.update_penalty = s.update_penalty; o.update_relax = s.update_relax;
        o.update_regularization = s.update_regularization; o.update_ReB = s.update_ReB; o.max_DDP_iter = s.max_DDP_iter;
        o.max_AL_iter = s.max_AL_iter; o.DDP_thresh = s.DDP_thresh; o.AL_thresh = s.AL_thresh;
        o.AL_active = s.AL_active; o.ReB_active = s.ReB_active; o.smooth_active = s.smooth_active;
        return o;
    }
    static void ck(int rc) { if (rc < 0) throw std::runtime_error(mhpc_last_error()); }
    static int next_mode(const mhpc_problem& p) {   // Gait::get_next_mode
        for (int i = 0; i < 4; ++i) if (p.cmode == p.gait_mode[i]) return p.gait_mode[(i + 1) % 4];
        return p.cmode;
    }
    void upload() { ck(mhpc_batch_set_problems(_h, _problems.data())); _dirty = false; }
    void refresh() {
        _actual_cost.resize(_n_instances); _exp_cost_change.resize(_n_instances); _tconstr_violation.resize(_n_instances);
        ck(mhpc_batch_get_scalars(_h, _actual_cost.data(), _exp_cost_change.data(), _tconstr_violation.data()));
    }
    // `rows` rows of a field in Eigen's default IOFormat; rows the phase does not have are zeros
    void dump(std::ofstream& os, int i, int p, mhpc_field f, int rows) {
        int dim = 0, kn = 0;
        std::vector<double> v = field(i, p, f, &dim, &kn);
        for (int k = 0; k < rows; ++k) {
            std::vector<std::string> txt(dim);
            size_t width = 0;
            for (int e = 0; e < dim; ++e) {
                std::ostringstream ss;
                ss.copyfmt(os);
                ss << (k < kn ? v[(size_t)k * dim + e] : 0.0);
                txt[e] = ss.str();
                width = std::max(width, txt[e].size());
            }
            for (int e = 0; e < dim; ++e) os << (e ? " " : "") << std::string(width - txt[e].size(), ' ') << txt[e];
            os << "\n";
        }
    }
    mhpc_batch* _h = nullptr;
    std::vector<mhpc_problem> _problems;
    bool _dirty = true;
};

#endif  // MHPC_B200_HPP
