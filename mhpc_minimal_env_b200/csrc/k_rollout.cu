// Forward rollout (SinglePhase.cpp:62-144 + MultiPhaseDDP.cpp:28-76,351-379): one thread per
// (line-search slot, instance); all step sizes of the Armijo search are evaluated as parallel
// rollouts.  Time is sequential inside the thread.  Lane mapping: with S = 2^k slots per launch a warp
// holds the S trials of 32/S consecutive instances (lane = instance_in_warp * S + slot), so the trials
// of one instance - which share its schedule, nominal trajectory and gains - never diverge and read K,
// du, x_nom, u_nom as broadcasts, while each store still fills whole 32 B sectors (32/S consecutive
// instances per row).  A single-slot launch (forward_sweep) maps lanes to 32 consecutive instances.
#include "kernels.h"
#include "model.cuh"

namespace mhpc {

__global__ void __launch_bounds__(128) k_rollout(Dev d, Consts c, Opt opt, int mask_mode, int n_slots, int trial0, EpsList eps_list) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    int w, i;
    if (n_slots > 1 && n_slots <= 32 && (n_slots & (n_slots - 1)) == 0) {
        const int lane = tid & 31, per_warp = 32 / n_slots;
        w = lane % n_slots;
        i = (tid >> 5) * per_warp + lane / n_slots;
    } else {
        w = tid / d.npad;
        i = tid - w * d.npad;
    }
    if (w >= n_slots || i >= d.n) return;
    if (!inst_enabled(d, i, mask_mode)) return;
    if (mask_mode == MASK_LS && d.si[(size_t)SI_LS_DONE * d.npad + i]) return;
    const double eps = eps_list.eps[trial0 + w];
    const bool reb = d.si[(size_t)SI_REB_ON * d.npad + i] != 0;
    const double vel = d.vel[i], height = d.height[i];
    const int np = d.n_phases[i], nwb = d.n_wb[i];
    double* act = d.st_act + act_idx(d, w, 0, 0, 0);
    const size_t es = d.npad;  // element stride

    double x[14];
#pragma unroll
    for (int e = 0; e < 14; ++e) x[e] = d.x0[(size_t)e * d.npad + i];
    double total = 0.0, viol2 = 0.0;
    for (int p = 0; p < np; ++p) {
        const int mode = d.ph_mode[ph_idx(d, p, i)], N = d.ph_N[ph_idx(d, p, i)], off = d.ph_off[ph_idx(d, p, i)];
        const bool wb = p < nwb;
        double V = 0.0;
        if (wb) {
            const double dt = d.dt_wb[i];
            double par[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) par[q] = d.par[par_idx(d, p, PAR_DELTA_T + q, i)];
            for (int k = 0; k < N - 1; ++k) {
                const int kk = off + k;
                const size_t b = st_idx(d, kk, 0, i);
                double dx[14], u[4], y[4], xn[14];
#pragma unroll
                for (int e = 0; e < 14; ++e) { act[b + e * es] = x[e]; dx[e] = x[e] - d.st_nom[b + e * es]; }
                // u = u_nom + eps*du + K (x - x_nom)   (SinglePhase.cpp:76)
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    double s = d.Kfb[((size_t)kk * KDIM + r) * es + i] * dx[0];
#pragma unroll
                    for (int e = 1; e < 14; ++e) s += d.Kfb[((size_t)kk * KDIM + r + 4 * e) * es + i] * dx[e];
                    u[r] = (d.st_nom[b + (14 + r) * es] + eps * d.dufb[((size_t)kk * 4 + r) * es + i]) + s;
                }
                wb_step(mode, dt, x, u, xn, y);
#pragma unroll
                for (int r = 0; r < 4; ++r) { act[b + (14 + r) * es] = u[r]; act[b + (18 + r) * es] = y[r]; }
                V += wb_running_cost(c, mode, dt, x, u, y, d.pos_ref[(size_t)kk * es + i], vel, height, reb, par);
#pragma unroll
                for (int e = 0; e < 14; ++e) x[e] = xn[e];
            }
            // terminal knot (SinglePhase.cpp:101-113 / 138-143)
            const int kk = off + N - 1;
            const size_t b = st_idx(d, kk, 0, i);
            const double pos = d.pos_ref[(size_t)kk * es + i];
            double Phi = 0.0;
#pragma unroll
            for (int e = 0; e < 14; ++e) {
                act[b + e * es] = x[e];
                const double dd = x[e] - wb_xref_term(c, mode, e, pos, vel);
                const double t = (dd * c.wbQf[mode - 1][e]) * dd;
                Phi = (e == 0) ? t : Phi + t;
            }
            Phi *= 0.5;
            double h = 0.0;
            if (mode == 2) tconstr_fl1_h(x, &h);
            else if (mode == 4) tconstr_fl2_h(x, &h);
            if (mode == 2 || mode == 4) {
                viol2 += h * h;
                if (opt.AL_active) {
                    const double sigma = d.par[par_idx(d, p, PAR_SIGMA, i)], lam = d.par[par_idx(d, p, PAR_LAMBDA, i)];
                    const double a = sigma * h / 2;
                    Phi += 50 * (a * a + lam * h);
                }
            }
            d.tr_h[((size_t)w * d.P + p) * es + i] = h;
            V += Phi;
            // phase transition (MultiPhaseDDP.cpp:351-379)
            if (p + 1 < np) {
                double xp[14];
                wb_reset(mode, x, xp);
                if (p + 1 < nwb) {
#pragma unroll
                    for (int e = 0; e < 14; ++e) x[e] = xp[e];
                } else {  // x_RB = P x_WB, P = [I3 0 0 0; 0 0 I3 0]  (MHPCLocomotion.cpp:32-34)
                    x[0] = xp[0]; x[1] = xp[1]; x[2] = xp[2]; x[3] = xp[7]; x[4] = xp[8]; x[5] = xp[9];
                }
            }
        } else {
            const double dt = d.dt_fb[i];
            double fh[4];
            rb_foothold(c, mode, x, dt * (double)N, fh);
            for (int k = 0; k < N - 1; ++k) {
                const int kk = off + k;
                const size_t b = st_idx(d, kk, 0, i);
                double dx[6], u[4], xn[6];
#pragma unroll
                for (int e = 0; e < 6; ++e) { act[b + e * es] = x[e]; dx[e] = x[e] - d.st_nom[b + e * es]; }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    double s = d.Kfb[((size_t)kk * KDIM + r) * es + i] * dx[0];
#pragma unroll
                    for (int e = 1; e < 6; ++e) s += d.Kfb[((size_t)kk * KDIM + r + 4 * e) * es + i] * dx[e];
                    u[r] = (d.st_nom[b + (14 + r) * es] + eps * d.dufb[((size_t)kk * 4 + r) * es + i]) + s;
                }
                rb_step(mode, dt, x, u, fh, xn);
#pragma unroll
                for (int r = 0; r < 4; ++r) { act[b + (14 + r) * es] = u[r]; act[b + (18 + r) * es] = 0.0; }
                V += rb_running_cost(c, mode, dt, x, u, d.pos_ref[(size_t)kk * es + i], vel, height);
#pragma unroll
                for (int e = 0; e < 6; ++e) x[e] = xn[e];
            }
            const int kk = off + N - 1;
            const size_t b = st_idx(d, kk, 0, i);
            const double pos = d.pos_ref[(size_t)kk * es + i];
            double Phi = 0.0;
#pragma unroll
            for (int e = 0; e < 6; ++e) {
                act[b + e * es] = x[e];
                const double dd = x[e] - rb_xref(e, pos, vel, height);
                const double t = (dd * c.rbQf[e]) * dd;
                Phi = (e == 0) ? t : Phi + t;
            }
            Phi *= 0.5;
            V += Phi;
            d.tr_h[((size_t)w * d.P + p) * es + i] = 0.0;
            // RB -> RB transition is the identity (MultiPhaseDDP.cpp:374-378)
        }
        total += V;
    }
    d.tr_cost[(size_t)w * es + i] = total;
    d.tr_viol2[(size_t)w * es + i] = viol2;
}

void launch_rollout(const Dev& d, const Consts& c, const Opt& opt, int mask_mode, int n_slots, int trial0, const EpsList& eps, cudaStream_t s) {
    const int threads = 128;
    const long total = (long)n_slots * d.npad;
    k_rollout<<<(unsigned)((total + threads - 1) / threads), threads, 0, s>>>(d, c, opt, mask_mode, n_slots, trial0, eps);
}

}  // namespace mhpc
