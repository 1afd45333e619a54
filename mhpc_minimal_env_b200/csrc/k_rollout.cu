// Forward rollout (SinglePhase.cpp:62-144 + MultiPhaseDDP.cpp:28-76,351-379): one thread per
// (line-search slot, instance); all step sizes of the Armijo search are evaluated as parallel
// rollouts.  Time is sequential inside the thread.  Lane mapping: with S = 2^k slots per launch a warp
// holds the S trials of 32/S consecutive instances (lane = instance_in_warp * S + slot), so the trials
// of one instance - which share its schedule, nominal trajectory and gains - never diverge and read K,
// du, x_nom, u_nom as shared-memory broadcasts, while each store still fills whole 32 B sectors (32/S
// consecutive instances per row).  A single-slot launch (forward_sweep) maps lanes to 32 consecutive instances.
#include <cuda_pipeline.h>

#include <algorithm>

#include "kernels.h"
#include "model.cuh"

namespace mhpc {

// Everything a step reads from HBM besides its own state (nominal x/u, the gain K, du, the reference position) sits at
// addresses known in advance; the trajectories and gains of a batch (GBs) do not stay in L2 between kernels, so read in
// place each step exposes several dependent DRAM/L2 round trips (ncu: 54 % of the warp time on the long scoreboard).
// The lanes that hold the trials of one instance stream that instance's next knot into shared memory with cp.async, one
// two steps ahead of the math (three-deep ring, synchronised inside the lane group only, so instances with different
// phase lengths may share a warp).
constexpr int RO_X = 0, RO_U = 14, RO_K = 18, RO_DU = 74, RO_POS = 78, RO_STAGE = 80;

constexpr int RO_DEPTH = 3;   // ring of staged knots: the copy of knot k+2 is issued while knot k is computed

// NL > 0: the lane-group size is a compile-time constant (the 8-trial line search), so the copy loops unroll into copies at
// constant multiples of the element stride; NL == 0 takes it from n_lanes.  Always commits one group (possibly empty):
// the consumer counts groups.
template <int NXP, int NL>   // NXP: 14 whole-body, 6 rigid-body
__device__ __forceinline__ void stage_knot(const Dev& d, bool copy, int kk, int i, int w, int n_lanes, double* dst) {
    if (copy) {
        const int nl = NL > 0 ? NL : n_lanes;
        const size_t es = d.npad;
        const double* sn = d.st_nom + st_idx(d, kk, 0, i) + (size_t)w * es;
        const double* kf = d.Kfb + ((size_t)kk * KDIM + w) * es + i;
        double* dw = dst + w;
#pragma unroll
        for (int t = 0; t < 4 * NXP; t += (NL > 0 ? NL : 1)) {
            if (NL > 0) { if (t + w < 4 * NXP) __pipeline_memcpy_async(dw + RO_K + t, kf + (size_t)t * es, 8); }
            else if (t % nl == 0 && t + w < 4 * NXP) __pipeline_memcpy_async(dw + RO_K + t, kf + (size_t)t * es, 8);
        }
#pragma unroll
        for (int t = 0; t < NXP; t += (NL > 0 ? NL : 1)) {
            if (NL > 0) { if (t + w < NXP) __pipeline_memcpy_async(dw + RO_X + t, sn + (size_t)t * es, 8); }
            else if (t % nl == 0 && t + w < NXP) __pipeline_memcpy_async(dw + RO_X + t, sn + (size_t)t * es, 8);
        }
        for (int t = w; t < 4; t += nl) {
            __pipeline_memcpy_async(dst + RO_U + t, d.st_nom + st_idx(d, kk, 14 + t, i), 8);
            __pipeline_memcpy_async(dst + RO_DU + t, d.dufb + ((size_t)kk * 4 + t) * es + i, 8);
        }
        if (w == nl - 1) __pipeline_memcpy_async(dst + RO_POS, d.pos_ref + (size_t)kk * es + i, 8);
    }
    __pipeline_commit();
}

template <bool STAGED, int NL>   // STAGED: the trials of an instance share a lane group (NL lanes, 0 = n_used) that streams the knot data through shared memory
__global__ void __launch_bounds__(128) k_rollout(Dev d, Consts c, Opt opt, int mask_mode, int n_slots, int n_used, int trial0, EpsList eps_list) {
    extern __shared__ __align__(16) double ro_smem[];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    // n_slots is a power of two <= 32 (launch_rollout): lane = instance_in_warp * n_slots + slot; slots >= n_used are padding
    const int lane = tid & 31, per_warp = 32 / n_slots;
    const int w = lane % n_slots;
    const int i = (tid >> 5) * per_warp + lane / n_slots;
    const unsigned grp = (n_used == 32 ? 0xffffffffu : ((1u << n_used) - 1u)) << (lane - w);   // lanes holding the trials of my instance
    double* stage = ro_smem + (size_t)(threadIdx.x / n_slots) * RO_DEPTH * RO_STAGE;         // [RO_DEPTH][RO_STAGE] of my instance
    if (w >= n_used || i >= d.n) return;
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
            if (STAGED) {
                __syncwarp(grp);
                stage_knot<14, NL>(d, 0 < N - 1, off, i, w, n_used, stage);
                stage_knot<14, NL>(d, 1 < N - 1, off + 1, i, w, n_used, stage + RO_STAGE);
            }
            for (int k = 0; k < N - 1; ++k) {
                const int kk = off + k;
                const size_t b = st_idx(d, kk, 0, i);
                double dx[14], u[4], y[4], xn[14];
                const double* sg = stage + (k % RO_DEPTH) * RO_STAGE;
                if (STAGED) {
                    __pipeline_wait_prior(1);   // everything but the newest group (knot k + 1) has landed
                    __syncwarp(grp);
                    stage_knot<14, NL>(d, k + 2 < N - 1, kk + 2, i, w, n_used, stage + ((k + 2) % RO_DEPTH) * RO_STAGE);
                }
#define RO_NOM(e) (STAGED ? sg[RO_X + (e)] : d.st_nom[b + (size_t)(e) * es])
#define RO_UNOM(r) (STAGED ? sg[RO_U + (r)] : d.st_nom[b + (size_t)(14 + (r)) * es])
#define RO_KF(q) (STAGED ? sg[RO_K + (q)] : d.Kfb[((size_t)kk * KDIM + (q)) * es + i])
#define RO_DUF(r) (STAGED ? sg[RO_DU + (r)] : d.dufb[((size_t)kk * 4 + (r)) * es + i])
#define RO_PREF() (STAGED ? sg[RO_POS] : d.pos_ref[(size_t)kk * es + i])
#pragma unroll
                for (int e = 0; e < 14; ++e) { act[b + e * es] = x[e]; dx[e] = x[e] - RO_NOM(e); }
                // u = u_nom + eps*du + K (x - x_nom)   (SinglePhase.cpp:76)
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    double s = RO_KF(r) * dx[0];
#pragma unroll
                    for (int e = 1; e < 14; ++e) s += RO_KF(r + 4 * e) * dx[e];
                    u[r] = (RO_UNOM(r) + eps * RO_DUF(r)) + s;
                }
                wb_step(mode, dt, x, u, xn, y);
#pragma unroll
                for (int r = 0; r < 4; ++r) { act[b + (14 + r) * es] = u[r]; act[b + (18 + r) * es] = y[r]; }
                V += wb_running_cost(c, mode, dt, x, u, y, RO_PREF(), vel, height, reb, par);
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
            if (STAGED) {
                __syncwarp(grp);
                stage_knot<6, NL>(d, 0 < N - 1, off, i, w, n_used, stage);
                stage_knot<6, NL>(d, 1 < N - 1, off + 1, i, w, n_used, stage + RO_STAGE);
            }
            for (int k = 0; k < N - 1; ++k) {
                const int kk = off + k;
                const size_t b = st_idx(d, kk, 0, i);
                double dx[6], u[4], xn[6];
                const double* sg = stage + (k % RO_DEPTH) * RO_STAGE;
                if (STAGED) {
                    __pipeline_wait_prior(1);   // everything but the newest group (knot k + 1) has landed
                    __syncwarp(grp);
                    stage_knot<6, NL>(d, k + 2 < N - 1, kk + 2, i, w, n_used, stage + ((k + 2) % RO_DEPTH) * RO_STAGE);
                }
#pragma unroll
                for (int e = 0; e < 6; ++e) { act[b + e * es] = x[e]; dx[e] = x[e] - RO_NOM(e); }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    double s = RO_KF(r) * dx[0];
#pragma unroll
                    for (int e = 1; e < 6; ++e) s += RO_KF(r + 4 * e) * dx[e];
                    u[r] = (RO_UNOM(r) + eps * RO_DUF(r)) + s;
                }
                rb_step(mode, dt, x, u, fh, xn);
#pragma unroll
                for (int r = 0; r < 4; ++r) { act[b + (14 + r) * es] = u[r]; act[b + (18 + r) * es] = 0.0; }
                V += rb_running_cost(c, mode, dt, x, u, RO_PREF(), vel, height);
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

#undef RO_NOM
#undef RO_UNOM
#undef RO_KF
#undef RO_DUF
#undef RO_PREF

void launch_rollout(const Dev& d, const Consts& c, const Opt& opt, int mask_mode, int n_slots, int trial0, const EpsList& eps, cudaStream_t s) {
    // slots per launch rounded up to a power of two (unused slots return at once); at most 32 instances (40 KB of staging)
    // per block: 128-thread blocks for the line search, one warp per block for single-slot sweeps so that their 4096
    // threads spread over all SMs
    int ns = 1;
    while (ns < n_slots) ns *= 2;
    if (ns > 32) ns = 32;
    const int threads = ns == 1 ? 32 : std::min(128, 16 * ns);      // at most 16 staged instances (30 KB) per block
    const int inst_per_block = threads / ns;
    const unsigned blocks = (unsigned)((d.npad + inst_per_block - 1) / inst_per_block);
    const size_t smem = (size_t)inst_per_block * RO_DEPTH * RO_STAGE * sizeof(double);
    if (n_slots == 8) {          // the lane-group sizes of the usual line-search widths are compile-time constants
        k_rollout<true, 8><<<blocks, threads, smem, s>>>(d, c, opt, mask_mode, ns, n_slots, trial0, eps);
    } else if (n_slots == 4) {
        k_rollout<true, 4><<<blocks, threads, smem, s>>>(d, c, opt, mask_mode, ns, n_slots, trial0, eps);
    } else if (n_slots == 2) {
        k_rollout<true, 2><<<blocks, threads, smem, s>>>(d, c, opt, mask_mode, ns, n_slots, trial0, eps);
    } else if (n_slots > 1) {
        k_rollout<true, 0><<<blocks, threads, smem, s>>>(d, c, opt, mask_mode, ns, n_slots, trial0, eps);
    } else {   // one trial per instance: nothing to share, every lane reads its own instance in place
        k_rollout<false, 0><<<blocks, threads, 0, s>>>(d, c, opt, mask_mode, ns, n_slots, trial0, eps);
    }
}

}  // namespace mhpc
