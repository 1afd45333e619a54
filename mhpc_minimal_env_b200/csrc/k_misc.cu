// Per-instance bookkeeping kernels: problem setup (reference generator, AL/ReB parameters), the
// warm-start heuristic, nominal update, Armijo selection, and the scalar control flow of
// MultiPhaseDDP::solve (MultiPhaseDDP.cpp:154-289) as a per-instance state machine.
#include "kernels.h"
#include "model.cuh"

namespace mhpc {

#define SI(f) d.si[(size_t)(f) * d.npad + i]
#define SD(f) d.sd[(size_t)(f) * d.npad + i]

// ReferenceGen.h:94-109 (forward position reference) + MHPCConstraints.cpp:59-88 (initial AL/ReB parameters)
__global__ void k_setup(Dev d, Consts c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    const int np = d.n_phases[i], nwb = d.n_wb[i];
    const double vel = d.vel[i];
    double pos = d.x0[i];  // x0[0]
    for (int p = 0; p < np; ++p) {
        const double dt = (p < nwb) ? d.dt_wb[i] : d.dt_fb[i];
        const int off = d.ph_off[ph_idx(d, p, i)], N = d.ph_N[ph_idx(d, p, i)];
        // pos[p][0] = pos[p-1][N-1]; pos[p][k] = pos[p][k-1] + vel*dt
        d.pos_ref[(size_t)off * d.npad + i] = pos;
        for (int k = 1; k < N; ++k) {
            pos = pos + vel * dt;
            d.pos_ref[(size_t)(off + k) * d.npad + i] = pos;
        }
        const int mode = d.ph_mode[ph_idx(d, p, i)];
        const bool wb = p < nwb;
        const bool flight = (mode == 2 || mode == 4);
        d.par[par_idx(d, p, PAR_SIGMA, i)] = (wb && flight) ? 5.0 : 0.0;
        d.par[par_idx(d, p, PAR_LAMBDA, i)] = 0.0;
        d.par[par_idx(d, p, PAR_DELTA_T, i)] = wb ? 0.1 : 0.0;
        d.par[par_idx(d, p, PAR_DELTA_J, i)] = wb ? 0.1 : 0.0;
        d.par[par_idx(d, p, PAR_DELTA_G, i)] = (wb && !flight) ? 0.1 : 0.0;
        d.par[par_idx(d, p, PAR_EPS_T, i)] = wb ? 0.01 * 1.0 : 0.0;
        d.par[par_idx(d, p, PAR_EPS_J, i)] = 0.0;
        d.par[par_idx(d, p, PAR_EPS_G, i)] = (wb && !flight) ? 0.01 * 1.0 : 0.0;
        d.par[par_idx(d, p, PAR_H, i)] = 0.0;
    }
    for (int f = 0; f < SI_N; ++f) if (f != SI_OPT_REB) SI(f) = 0;            // _option is not touched by initialization()
    for (int f = 0; f < SD_N; ++f) if (f != SD_OPT_UPD) SD(f) = 0.0;
}

// planar homogeneous transform restricted to the (x,z) plane: R = [[r00 r02],[r20 r22]], t = (tx,tz)
struct T2 { double r00, r02, r20, r22, tx, tz; };
__device__ inline T2 t2_make(double theta, double tx, double tz) {   // orientation_tools.h:58-76 (axis Y), :383-392
    double s, cs;
    sincos(theta, &s, &cs);
    return T2{cs, -s, s, cs, tx, tz};
}
__device__ inline T2 t2_mul(const T2& a, const T2& b) {   // 4x4 product with the structural zeros dropped
    T2 o;
    o.r00 = a.r00 * b.r00 + a.r02 * b.r20;
    o.r02 = a.r00 * b.r02 + a.r02 * b.r22;
    o.r20 = a.r20 * b.r00 + a.r22 * b.r20;
    o.r22 = a.r20 * b.r02 + a.r22 * b.r22;
    o.tx = (a.r00 * b.tx + a.r02 * b.tz) + a.tx;
    o.tz = (a.r20 * b.tx + a.r22 * b.tz) + a.tz;
    return o;
}
// PlanarQuadruped.cpp:195-205 via :142-186 (leg 0 front, 1 hind): foot position - hip position
__device__ inline void leg_ext_vec(const Consts& c, const double* q, int leg, double* out) {
    const T2 body = t2_make(-q[2], q[0], q[1]);
    const T2 hip = t2_mul(body, t2_make(leg == 0 ? -q[3] : -q[5], leg == 0 ? c.hip_x : -c.hip_x, 0.0));
    const T2 knee = t2_mul(hip, t2_make(leg == 0 ? -q[4] : -q[6], 0.0, -c.hip_len));
    const T2 ident{1.0, 0.0, 0.0, 1.0, 0.0, -c.knee_len};
    const T2 foot = t2_mul(knee, ident);
    out[0] = foot.tx - hip.tx;
    out[1] = foot.tz - hip.tz;
}

// Controller/Heuristics/boundingPDControl.cpp:3-46 + MHPCLocomotion.cpp:200-215 on the device:
// spring-leg / joint-PD rollout of the whole-body phases into act slot 0.
__global__ void __launch_bounds__(128) k_warmstart(Dev d, Consts c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    const int nwb = d.n_wb[i];
    const double dt = d.dt_wb[i];
    const size_t es = d.npad;
    double* act = d.st_act;
    double x[14];
#pragma unroll
    for (int e = 0; e < 14; ++e) x[e] = d.x0[(size_t)e * es + i];
    for (int p = 0; p < nwb; ++p) {
        const int mode = d.ph_mode[ph_idx(d, p, i)], N = d.ph_N[ph_idx(d, p, i)], off = d.ph_off[ph_idx(d, p, i)];
        for (int k = 0; k < N - 1; ++k) {
            const size_t b = st_idx(d, off + k, 0, i);
            double u[4], y[4], xn[14];
            if (mode == 1 || mode == 3) {
                const int leg = (mode == 1) ? 1 : 0;
                const double gain = (mode == 1) ? 3 : 2.2;
                double J[14];
#pragma unroll
                for (int q = 0; q < 14; ++q) J[q] = 0.0;
                if (leg == 0) jacob_f(x, J); else jacob_b(x, J);
                double lv[2];
                leg_ext_vec(c, x, leg, lv);
                const double sq = lv[0] * lv[0] + lv[1] * lv[1];
                const double nrm = sqrt(sq);
                double nv0 = lv[0], nv1 = lv[1];
                if (sq > 0) { nv0 = lv[0] / sqrt(sq); nv1 = lv[1] / sqrt(sq); }
                const double F0 = ((-nv0) * c.kspring) * (nrm - c.legext_nom), F1 = ((-nv1) * c.kspring) * (nrm - c.legext_nom);
#pragma unroll
                for (int q = 0; q < 4; ++q) u[q] = (J[0 + 2 * (3 + q)] * F0 + J[1 + 2 * (3 + q)] * F1) * gain;
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) u[q] = c.kp[q] * (c.qjoint_nom[q] - x[3 + q]) - 1.0 * x[10 + q];
            }
            wb_step(mode, dt, x, u, xn, y);
#pragma unroll
            for (int e = 0; e < 14; ++e) act[b + e * es] = x[e];
#pragma unroll
            for (int e = 0; e < 4; ++e) { act[b + (14 + e) * es] = u[e]; act[b + (18 + e) * es] = y[e]; }
#pragma unroll
            for (int e = 0; e < 14; ++e) x[e] = xn[e];
        }
        const size_t b = st_idx(d, off + N - 1, 0, i);
#pragma unroll
        for (int e = 0; e < 14; ++e) act[b + e * es] = x[e];
        if (p + 1 < nwb) {
            double xp[14];
            wb_reset(mode, x, xp);
#pragma unroll
            for (int e = 0; e < 14; ++e) x[e] = xp[e];
        }
    }
}

// SinglePhase.cpp:358-361: nominal <- actual (of the accepted line-search slot)
__global__ void k_copy_nominal(Dev d, int mask_mode) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    if (!inst_enabled(d, i, mask_mode)) return;
    const int kk = blockIdx.y;    // one knot (22 rows of the SoA trajectory) per thread: the index work is done once
    int p, k;
    if (!knot_to_phase(d, i, kk, p, k)) return;   // memcpy of N_p knots per phase (SinglePhase.cpp:358-361)
    const int slot = SI(SI_ACC_SLOT);
    const double* __restrict__ src = d.st_act + ((size_t)slot * d.KT + kk) * ST * d.npad + i;
    double* __restrict__ dst = d.st_nom + (size_t)kk * ST * d.npad + i;
    double v[ST];
#pragma unroll
    for (int e = 0; e < ST; ++e) v[e] = src[(size_t)e * d.npad];
#pragma unroll
    for (int e = 0; e < ST; ++e) dst[(size_t)e * d.npad] = v[e];
}

__global__ void k_snapshot_mask(Dev d, int mask_mode) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    SI(SI_SNAP) = inst_enabled(d, i, mask_mode) ? 1 : 0;
}

__global__ void k_accept_slot0(Dev d, int mask_mode, int lin_mode) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n || !inst_enabled(d, i, mask_mode)) return;
    SD(SD_COST) = d.tr_cost[i];
    SD(SD_VIOL) = sqrt(d.tr_viol2[i]);
    const int np = d.n_phases[i];
    for (int p = 0; p < np; ++p) d.par[par_idx(d, p, PAR_H, i)] = d.tr_h[(size_t)p * d.npad + i];
    SI(SI_ACC_SLOT) = 0;
    SI(SI_AL_PARTIALS) = lin_mode;
}

__global__ void k_solve_begin(Dev d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    SI(SI_STATUS) = STATUS_RUNNING;
    SI(SI_ITER_AL) = 1;
    SI(SI_N_REC) = 0;
    SI(SI_DDP_ACTIVE) = 0;
    SI(SI_REB_BASE) = SI(SI_OPT_REB);      // bool ReB_active = _option.ReB_active;   (MultiPhaseDDP.cpp:161)
    SD(SD_UPD_BASE) = SD(SD_OPT_UPD);      // T update_penalty = _option.update_penalty; (:160)
}

// MultiPhaseDDP.cpp:178-195
__global__ void k_al_begin(Dev d, Opt opt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n || SI(SI_STATUS) != STATUS_RUNNING) return;
    int reb = SI(SI_REB_BASE);
    if ((SD(SD_VIOL) > 0.05) || 1 == SI(SI_ITER_AL)) reb = 0;
    SI(SI_OPT_REB) = reb;
    SI(SI_REB_ON) = reb;
    SI(SI_ITER_DDP) = 1;
    SD(SD_REG) = 0.0;
    SI(SI_DDP_ACTIVE) = 1;
}

// entry of forward_iteration (MultiPhaseDDP.cpp:132-134) / top of the inner loop (:203)
__global__ void k_ls_begin(Dev d, int mask_mode) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n || !inst_enabled(d, i, mask_mode)) return;
    SD(SD_COST_PREV) = SD(SD_COST);
    SI(SI_LS_DONE) = 0;
}

// Armijo test over one chunk of parallel trials: the FIRST step size of the reference's sequence that
// satisfies MultiPhaseDDP.cpp:142 wins; if none does, the last trial is kept.
__global__ void k_select(Dev d, Opt opt, int mask_mode, int trial0, int n_slots, int last_chunk, EpsList el) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n || !inst_enabled(d, i, mask_mode)) return;
    if (SI(SI_LS_DONE)) return;
    const double cost_prev = SD(SD_COST_PREV), dV = SD(SD_DV);
    int acc = -1, nfit = 0;
    for (int w = 0; w < n_slots; ++w) {
        const double eps = el.eps[trial0 + w];
        const double cst = d.tr_cost[(size_t)w * d.npad + i];
        if (cst <= cost_prev + opt.gamma * eps * (1 - eps / 2) * dV) { acc = w; nfit = trial0 + w + 1; break; }
    }
    if (acc < 0 && last_chunk) { acc = n_slots - 1; nfit = el.n + 1; }
    if (acc < 0) return;
    SI(SI_ACC_SLOT) = acc;
    SI(SI_N_FIT) = nfit;
    SI(SI_LS_DONE) = 1;
    SD(SD_EPS_ACC) = el.eps[trial0 + acc];
    SD(SD_COST) = d.tr_cost[(size_t)acc * d.npad + i];
    SD(SD_VIOL) = sqrt(d.tr_viol2[(size_t)acc * d.npad + i]);
    const int np = d.n_phases[i];
    for (int p = 0; p < np; ++p) d.par[par_idx(d, p, PAR_H, i)] = d.tr_h[((size_t)acc * d.P + p) * d.npad + i];
}

// MultiPhaseDDP.cpp:255-270: record, convergence test, iteration counter
__global__ void k_ddp_end(Dev d, Opt opt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n || SI(SI_STATUS) != STATUS_RUNNING) return;
    if (!SI(SI_DDP_ACTIVE)) { SI(SI_AL_PARTIALS) = 0; return; }
    const int r = SI(SI_N_REC);
    if (r < d.max_rec) {
        int* ri = d.rec_i + (size_t)r * REC_I * d.npad + i;
        double* rd = d.rec_d + (size_t)r * REC_D * d.npad + i;
        ri[0 * d.npad] = SI(SI_ITER_AL); ri[1 * d.npad] = SI(SI_ITER_DDP); ri[2 * d.npad] = SI(SI_N_BWS); ri[3 * d.npad] = SI(SI_N_FIT);
        rd[0 * d.npad] = SD(SD_REG_USED); rd[1 * d.npad] = SD(SD_EPS_ACC); rd[2 * d.npad] = SD(SD_COST);
        rd[3 * d.npad] = SD(SD_DV); rd[4 * d.npad] = SD(SD_VIOL); rd[5 * d.npad] = SD(SD_COST_PREV);
        SI(SI_N_REC) = r + 1;
    }
    if (SD(SD_COST_PREV) - SD(SD_COST) < opt.DDP_thresh) {
        SI(SI_DDP_ACTIVE) = 0;
        SI(SI_AL_PARTIALS) = 0;
    } else {
        SI(SI_AL_PARTIALS) = 1;
        const int it = SI(SI_ITER_DDP) + 1;
        SI(SI_ITER_DDP) = it;
        if (!((double)it <= opt.max_DDP_iter)) SI(SI_DDP_ACTIVE) = 0;
    }
}

// MultiPhaseDDP.cpp:273-287 + SinglePhase.cpp:334-354
__global__ void k_al_end(Dev d, Opt opt, Consts c, int mask_mode, int solve_mode) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n || !inst_enabled(d, i, mask_mode)) return;
    if (solve_mode) {   // _option.update_penalty = update_penalty; if (viol < 0.03) _option.update_penalty = 0;
        double u = SD(SD_UPD_BASE);
        if (SD(SD_VIOL) < 0.03) u = 0;
        SD(SD_OPT_UPD) = u;
    }
    const double upd = SD(SD_OPT_UPD);
    const int reb = SI(SI_OPT_REB);
    SD(SD_UPD_PENALTY) = upd;
    const int np = d.n_phases[i], nwb = d.n_wb[i];
    for (int p = 0; p < nwb && p < np; ++p) {
        const int mode = d.ph_mode[ph_idx(d, p, i)];
        double sigma = d.par[par_idx(d, p, PAR_SIGMA, i)];
        if (mode == 2 || mode == 4) d.par[par_idx(d, p, PAR_LAMBDA, i)] += sigma * d.par[par_idx(d, p, PAR_H, i)];
        sigma *= upd;
        d.par[par_idx(d, p, PAR_SIGMA, i)] = sigma;
        if (reb) {
            for (int g = 0; g < 3; ++g) {
                if (g == 2 && (mode == 2 || mode == 4)) continue;
                double dl = d.par[par_idx(d, p, PAR_DELTA_T + g, i)] * opt.update_relax;
                if (dl < 0.01) dl = 0.01;   // delta_min = 0.01 (MHPCConstraints.cpp:62)
                d.par[par_idx(d, p, PAR_DELTA_T + g, i)] = dl;
                d.par[par_idx(d, p, PAR_EPS_T + g, i)] *= opt.update_ReB;
            }
        }
    }
    if (solve_mode) {
        if (SD(SD_VIOL) < opt.AL_thresh) SI(SI_STATUS) = STATUS_CONVERGED;
        else {
            const int it = SI(SI_ITER_AL) + 1;
            if (!((double)it <= opt.max_AL_iter)) SI(SI_STATUS) = STATUS_MAXITER;
            else SI(SI_ITER_AL) = it;
        }
    }
}

// MultiPhaseDDP(n_phases, option) / set_option: the per-instance copy of the option members solve() overwrites
__global__ void k_reset_option_state(Dev d, Opt opt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    SI(SI_OPT_REB) = opt.ReB_active;
    SD(SD_OPT_UPD) = opt.update_penalty;
}
// the bare sweep entry points read `_option.ReB_active` as it currently is
__global__ void k_reb_from_option(Dev d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    SI(SI_REB_ON) = SI(SI_OPT_REB);
}

// update_problem (MHPCLocomotion.cpp:107-121,134-151): the whole-body and the rigid-body phase buffers rotate by one
// (new phase p uses the buffer of old phase p+1, the last one of each group gets the first).  One thread per
// (knot slot, element, instance); `src`/`dst` are different allocations.  The accepted rollout becomes slot 0.
__global__ void k_rotate(Dev d, const double* __restrict__ src, double* __restrict__ dst, int elems, int use_acc_slot) {
    const int i = blockIdx.y * blockDim.x + threadIdx.x;   // rows on grid.x: KT * elems exceeds the 65535 limit of grid.y from 11 phases on
    if (i >= d.n) return;
    const int row = blockIdx.x;            // dst knot slot * elems + e
    const int kk = row / elems, e = row - kk * elems;
    const int p = kk / KSTRIDE, k = kk - p * KSTRIDE;
    const int np = d.n_phases[i], nwb = d.n_wb[i];
    if (p >= np) return;
    int q;
    if (p < nwb) q = (p + 1 < nwb) ? p + 1 : 0;
    else q = (p + 1 < np) ? p + 1 : nwb;
    const size_t slot_off = use_acc_slot ? (size_t)SI(SI_ACC_SLOT) * d.KT * elems * d.npad : 0;
    dst[(size_t)row * d.npad + i] = src[slot_off + ((size_t)(q * KSTRIDE + k) * elems + e) * d.npad + i];
}
// the same rotation for the gain records [npad][KT][GREC]: one thread per (knot slot, instance)
__global__ void k_rotate_gain(Dev d, const double* __restrict__ src, double* __restrict__ dst) {
    const int i = blockIdx.y * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    const int kk = blockIdx.x;
    const int p = kk / KSTRIDE, k = kk - p * KSTRIDE;
    const int np = d.n_phases[i], nwb = d.n_wb[i];
    if (p >= np) return;
    int q;
    if (p < nwb) q = (p + 1 < nwb) ? p + 1 : 0;
    else q = (p + 1 < np) ? p + 1 : nwb;
    const double* s = src + gain_idx(d, q * KSTRIDE + k, (size_t)i);
    double* t = dst + gain_idx(d, kk, (size_t)i);
    for (int e = 0; e < GREC; ++e) t[e] = s[e];
}
// after the rotation every line-search slot must carry the same never-recomputed u,y of the last knot of each phase
//  as the single `act` buffer of the reference
__global__ void k_spread_last_knot(Dev d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    const int p = blockIdx.y;
    if (p >= d.n_phases[i]) return;
    const int kk = p * KSTRIDE + d.ph_N[ph_idx(d, p, i)] - 1;
    for (int w = 1; w < d.W; ++w)
        for (int e = 14; e < ST; ++e) d.st_act[act_idx(d, w, kk, e, i)] = d.st_act[act_idx(d, 0, kk, e, i)];
}
// build_problem part of update_problem: references + re-initialised AL/ReB parameters (no other state is touched)
__global__ void k_setup_refs_params(Dev d, Consts c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    const int np = d.n_phases[i], nwb = d.n_wb[i];
    const double vel = d.vel[i];
    double pos = d.x0[i];
    for (int p = 0; p < np; ++p) {
        const double dt = (p < nwb) ? d.dt_wb[i] : d.dt_fb[i];
        const int off = d.ph_off[ph_idx(d, p, i)], N = d.ph_N[ph_idx(d, p, i)];
        d.pos_ref[(size_t)off * d.npad + i] = pos;
        for (int k = 1; k < N; ++k) {
            pos = pos + vel * dt;
            d.pos_ref[(size_t)(off + k) * d.npad + i] = pos;
        }
        const int mode = d.ph_mode[ph_idx(d, p, i)];
        const bool wb = p < nwb;
        const bool flight = (mode == 2 || mode == 4);
        d.par[par_idx(d, p, PAR_SIGMA, i)] = (wb && flight) ? 5.0 : 0.0;
        d.par[par_idx(d, p, PAR_LAMBDA, i)] = 0.0;
        d.par[par_idx(d, p, PAR_DELTA_T, i)] = wb ? 0.1 : 0.0;
        d.par[par_idx(d, p, PAR_DELTA_J, i)] = wb ? 0.1 : 0.0;
        d.par[par_idx(d, p, PAR_DELTA_G, i)] = (wb && !flight) ? 0.1 : 0.0;
        d.par[par_idx(d, p, PAR_EPS_T, i)] = wb ? 0.01 * 1.0 : 0.0;
        d.par[par_idx(d, p, PAR_EPS_J, i)] = 0.0;
        d.par[par_idx(d, p, PAR_EPS_G, i)] = (wb && !flight) ? 0.01 * 1.0 : 0.0;
        d.par[par_idx(d, p, PAR_H, i)] = 0.0;     // TConstrStruct() re-created zeroed (SinglePhase.cpp:35-36)
    }
    SI(SI_ACC_SLOT) = 0;
}

__global__ void k_count_active(Dev d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    if (SI(SI_STATUS) == STATUS_RUNNING) atomicAdd(d.n_active, 1);
}

static inline unsigned nb(int n, int t) { return (unsigned)((n + t - 1) / t); }

void launch_setup(const Dev& d, const Consts& c, cudaStream_t s) { k_setup<<<nb(d.n, 128), 128, 0, s>>>(d, c); }
void launch_warmstart(const Dev& d, const Consts& c, cudaStream_t s) { k_warmstart<<<nb(d.n, 128), 128, 0, s>>>(d, c); }
void launch_copy_nominal(const Dev& d, int mask_mode, cudaStream_t s) {
    dim3 grid(nb(d.n, 128), d.KT);
    k_copy_nominal<<<grid, 128, 0, s>>>(d, mask_mode);
}
void launch_snapshot_mask(const Dev& d, int mask_mode, cudaStream_t s) { k_snapshot_mask<<<nb(d.n, 128), 128, 0, s>>>(d, mask_mode); }
void launch_accept_slot0(const Dev& d, int mask_mode, int lin_mode, cudaStream_t s) { k_accept_slot0<<<nb(d.n, 128), 128, 0, s>>>(d, mask_mode, lin_mode); }
void launch_solve_begin(const Dev& d, cudaStream_t s) { k_solve_begin<<<nb(d.n, 128), 128, 0, s>>>(d); }
void launch_al_begin(const Dev& d, const Opt& o, cudaStream_t s) { k_al_begin<<<nb(d.n, 128), 128, 0, s>>>(d, o); }
void launch_ls_begin(const Dev& d, int mask_mode, cudaStream_t s) { k_ls_begin<<<nb(d.n, 128), 128, 0, s>>>(d, mask_mode); }
void launch_select(const Dev& d, const Opt& o, int mask_mode, int trial0, int n_slots, int last_chunk, const EpsList& e, cudaStream_t s) {
    k_select<<<nb(d.n, 128), 128, 0, s>>>(d, o, mask_mode, trial0, n_slots, last_chunk, e);
}
void launch_ddp_end(const Dev& d, const Opt& o, int, cudaStream_t s) { k_ddp_end<<<nb(d.n, 128), 128, 0, s>>>(d, o); }
void launch_al_end(const Dev& d, const Opt& o, const Consts& c, int mask_mode, int solve_mode, cudaStream_t s) {
    k_al_end<<<nb(d.n, 128), 128, 0, s>>>(d, o, c, mask_mode, solve_mode);
}
void launch_reset_option_state(const Dev& d, const Opt& o, cudaStream_t s) { k_reset_option_state<<<nb(d.n, 128), 128, 0, s>>>(d, o); }
void launch_reb_from_option(const Dev& d, cudaStream_t s) { k_reb_from_option<<<nb(d.n, 128), 128, 0, s>>>(d); }
void launch_rotate(const Dev& d, const double* src, double* dst, int elems, int use_acc_slot, cudaStream_t s) {
    dim3 grid(d.KT * elems, nb(d.n, 128));
    k_rotate<<<grid, 128, 0, s>>>(d, src, dst, elems, use_acc_slot);
}
void launch_rotate_gain(const Dev& d, const double* src, double* dst, cudaStream_t s) {
    dim3 grid(d.KT, nb(d.n, 128));
    k_rotate_gain<<<grid, 128, 0, s>>>(d, src, dst);
}
void launch_spread_last_knot(const Dev& d, cudaStream_t s) {
    dim3 grid(nb(d.n, 128), d.P);
    k_spread_last_knot<<<grid, 128, 0, s>>>(d);
}
void launch_setup_refs_params(const Dev& d, const Consts& c, cudaStream_t s) { k_setup_refs_params<<<nb(d.n, 128), 128, 0, s>>>(d, c); }
void launch_count_active(const Dev& d, cudaStream_t s) {
    cudaMemsetAsync(d.n_active, 0, sizeof(int), s);
    k_count_active<<<nb(d.n, 128), 128, 0, s>>>(d);
}

}  // namespace mhpc
