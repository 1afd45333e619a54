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
// A whole-body knot is 79 doubles (x 14 | u 4 | K 56 | du 4 | pos), a rigid-body knot 39 (6 | 4 | 24 | 4 | 1): the same
// 240 doubles per instance hold a ring of three whole-body knots or of six rigid-body knots.  The rigid-body steps are
// short (~0.4 us), so two steps of lead do not cover a DRAM round trip there (ncu: a quarter of the kernel's samples sat
// on the copy wait and the first reads of the rigid-body loop); five do.
template <int NXP> struct RoLay;
#ifndef RO_WB_DEPTH
#define RO_WB_DEPTH 3
#endif
template <> struct RoLay<14> { enum { X = 0, U = 14, K = 18, DU = 74, POS = 78, STAGE = 80, DEPTH = RO_WB_DEPTH }; };
template <> struct RoLay<6> { enum { X = 0, U = 6, K = 10, DU = 34, POS = 38, STAGE = 40, DEPTH = 6 }; };
constexpr int RO_RING = RO_WB_DEPTH * 80 > 240 ? RO_WB_DEPTH * 80 : 240;   // doubles per instance

// The lanes of a warp address the stages of 32 / n_slots different instances at once; with the stages a multiple of 16
// doubles apart all of them hit the same bank pair (ncu: 79 % of the shared-memory wavefronts of the 4-trial launch were
// bank conflicts).  n_slots doubles of padding put instance j of a warp at bank offset j * n_slots, so the copies
// (n_slots consecutive doubles per instance) and the broadcast reads are conflict-free.
__host__ __device__ constexpr int ro_stride(int n_slots) { return RO_RING + (n_slots < 16 ? n_slots : 0); }

// NL > 0: the lane-group size is a compile-time constant (the 8-trial line search), so the copy loops unroll into copies at
// constant multiples of the element stride; NL == 0 takes it from n_lanes.  Always commits one group (possibly empty):
// the consumer counts groups.
template <int NXP, int NL>   // NXP: 14 whole-body, 6 rigid-body
__device__ __forceinline__ void stage_knot(const Dev& d, bool copy, int kk, int i, int w, int n_lanes, double* dst) {
    if (copy) {
        typedef RoLay<NXP> L;
        const int nl = NL > 0 ? NL : n_lanes;
        const size_t es = d.npad;
        const double* sn = d.st_nom + st_idx(d, kk, 0, i) + (size_t)w * es;
        // K | du: the head of the knot's gain record, contiguous -> 16-byte copies (whole-body: one run of 30; rigid body: K 4x6, then du)
        const double* gr = d.gain + gain_idx(d, kk, (size_t)i);
        constexpr int NK16 = 2 * NXP + (NXP == 14 ? 2 : 0);
#pragma unroll
        for (int t = 0; t < NK16; t += (NL > 0 ? NL : 1)) {
            if (NL > 0) { if (t + w < NK16) __pipeline_memcpy_async(dst + L::K + 2 * (t + w), gr + 2 * (t + w), 16); }
            else if (t % nl == 0 && t + w < NK16) __pipeline_memcpy_async(dst + L::K + 2 * (t + w), gr + 2 * (t + w), 16);
        }
        if (NXP != 14 && w < 2) __pipeline_memcpy_async(dst + L::DU + 2 * w, gr + GR_DU + 2 * w, 16);
        double* dw = dst + w;
#pragma unroll
        for (int t = 0; t < NXP; t += (NL > 0 ? NL : 1)) {
            if (NL > 0) { if (t + w < NXP) __pipeline_memcpy_async(dw + L::X + t, sn + (size_t)t * es, 8); }
            else if (t % nl == 0 && t + w < NXP) __pipeline_memcpy_async(dw + L::X + t, sn + (size_t)t * es, 8);
        }
        for (int t = w; t < 4; t += nl) __pipeline_memcpy_async(dst + L::U + t, d.st_nom + st_idx(d, kk, 14 + t, i), 8);
        if (w == nl - 1) __pipeline_memcpy_async(dst + L::POS, d.pos_ref + (size_t)kk * es + i, 8);
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
    double* stage = ro_smem + (size_t)(threadIdx.x / n_slots) * ro_stride(n_slots);          // [RO_DEPTH][RO_STAGE] of my instance
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
#pragma unroll
                for (int q = 0; q < RoLay<14>::DEPTH - 1; ++q) stage_knot<14, NL>(d, q < N - 1, off + q, i, w, n_used, stage + q * RoLay<14>::STAGE);
            }
            for (int k = 0; k < N - 1; ++k) {
                const int kk = off + k;
                const size_t b = st_idx(d, kk, 0, i);
                double dx[14], u[4], y[4], xn[14];
                typedef RoLay<14> L;
                const double* sg = stage + (k % L::DEPTH) * L::STAGE;
                if (STAGED) {
                    __pipeline_wait_prior(L::DEPTH - 2);   // everything but the newest groups (knots k + 1 ..) has landed
                    __syncwarp(grp);
                    stage_knot<14, NL>(d, k + L::DEPTH - 1 < N - 1, kk + L::DEPTH - 1, i, w, n_used, stage + ((k + L::DEPTH - 1) % L::DEPTH) * L::STAGE);
                }
#define RO_NOM(e) (STAGED ? sg[L::X + (e)] : d.st_nom[b + (size_t)(e) * es])
#define RO_UNOM(r) (STAGED ? sg[L::U + (r)] : d.st_nom[b + (size_t)(14 + (r)) * es])
#define RO_KF(q) (STAGED ? sg[L::K + (q)] : d.gain[gain_idx(d, kk, (size_t)i) + GR_K + (q)])
#define RO_DUF(r) (STAGED ? sg[L::DU + (r)] : d.gain[gain_idx(d, kk, (size_t)i) + GR_DU + (r)])
#define RO_PREF() (STAGED ? sg[L::POS] : d.pos_ref[(size_t)kk * es + i])
#pragma unroll
                for (int e = 0; e < 14; ++e) { act[b + e * es] = x[e]; dx[e] = x[e] - RO_NOM(e); }
                // u = u_nom + eps*du + K (x - x_nom)   (SinglePhase.cpp:76)
                if (STAGED) {   // the four gains of a state are 32 contiguous bytes of the staged record: two 16-byte loads (-7 % on the launch)
                    const double2* k2 = reinterpret_cast<const double2*>(sg + L::K);
                    double s[4];
#pragma unroll
                    for (int e = 0; e < 14; ++e) {
                        const double2 ka = k2[2 * e], kb = k2[2 * e + 1];
                        if (e == 0) { s[0] = ka.x * dx[0]; s[1] = ka.y * dx[0]; s[2] = kb.x * dx[0]; s[3] = kb.y * dx[0]; }
                        else { s[0] += ka.x * dx[e]; s[1] += ka.y * dx[e]; s[2] += kb.x * dx[e]; s[3] += kb.y * dx[e]; }
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r) u[r] = (RO_UNOM(r) + eps * RO_DUF(r)) + s[r];
                } else {
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        double s = RO_KF(r) * dx[0];
#pragma unroll
                        for (int e = 1; e < 14; ++e) s += RO_KF(r + 4 * e) * dx[e];
                        u[r] = (RO_UNOM(r) + eps * RO_DUF(r)) + s;
                    }
                }
                wb_step(mode, dt, x, u, xn, y);
#pragma unroll
                for (int r = 0; r < 4; ++r) { act[b + (14 + r) * es] = u[r]; act[b + (18 + r) * es] = y[r]; }
                V = __dadd_rn(V, wb_running_cost(c, mode, dt, x, u, y, RO_PREF(), vel, height, reb, par));
#pragma unroll
                for (int e = 0; e < 14; ++e) x[e] = xn[e];
            }
            // terminal knot (SinglePhase.cpp:101-113 / 138-143)
            const int kk = off + N - 1;
            const size_t b = st_idx(d, kk, 0, i);
            const double pos = d.pos_ref[(size_t)kk * es + i];
#pragma unroll
            for (int e = 0; e < 14; ++e) act[b + e * es] = x[e];
            const bool al_term = (mode == 2 || mode == 4) && opt.AL_active;
            const double sigma = al_term ? d.par[par_idx(d, p, PAR_SIGMA, i)] : 0.0, lam = al_term ? d.par[par_idx(d, p, PAR_LAMBDA, i)] : 0.0;
            double h;
            const double Phi = wb_terminal_cost(c, mode, x, pos, vel, opt.AL_active != 0, sigma, lam, h);
            if (mode == 2 || mode == 4) viol2 = fma(h, h, viol2);
            d.tr_h[((size_t)w * d.P + p) * es + i] = h;
            V = __dadd_rn(V, Phi);
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
#pragma unroll
                for (int q = 0; q < RoLay<6>::DEPTH - 1; ++q) stage_knot<6, NL>(d, q < N - 1, off + q, i, w, n_used, stage + q * RoLay<6>::STAGE);
            }
            for (int k = 0; k < N - 1; ++k) {
                const int kk = off + k;
                const size_t b = st_idx(d, kk, 0, i);
                double dx[6], u[4], xn[6];
                typedef RoLay<6> L;
                const double* sg = stage + (k % L::DEPTH) * L::STAGE;
                if (STAGED) {
                    __pipeline_wait_prior(L::DEPTH - 2);   // everything but the newest four groups (knots k + 1 .. k + 4) has landed
                    __syncwarp(grp);
                    stage_knot<6, NL>(d, k + L::DEPTH - 1 < N - 1, kk + L::DEPTH - 1, i, w, n_used, stage + ((k + L::DEPTH - 1) % L::DEPTH) * L::STAGE);
                }
#pragma unroll
                for (int e = 0; e < 6; ++e) { act[b + e * es] = x[e]; dx[e] = x[e] - RO_NOM(e); }
                if (STAGED) {
                    const double2* k2 = reinterpret_cast<const double2*>(sg + L::K);
                    double s[4];
#pragma unroll
                    for (int e = 0; e < 6; ++e) {
                        const double2 ka = k2[2 * e], kb = k2[2 * e + 1];
                        if (e == 0) { s[0] = ka.x * dx[0]; s[1] = ka.y * dx[0]; s[2] = kb.x * dx[0]; s[3] = kb.y * dx[0]; }
                        else { s[0] += ka.x * dx[e]; s[1] += ka.y * dx[e]; s[2] += kb.x * dx[e]; s[3] += kb.y * dx[e]; }
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r) u[r] = (RO_UNOM(r) + eps * RO_DUF(r)) + s[r];
                } else {
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        double s = RO_KF(r) * dx[0];
#pragma unroll
                        for (int e = 1; e < 6; ++e) s += RO_KF(r + 4 * e) * dx[e];
                        u[r] = (RO_UNOM(r) + eps * RO_DUF(r)) + s;
                    }
                }
                rb_step(mode, dt, x, u, fh, xn);
#pragma unroll
                for (int r = 0; r < 4; ++r) { act[b + (14 + r) * es] = u[r]; act[b + (18 + r) * es] = 0.0; }
                V = __dadd_rn(V, rb_running_cost(c, mode, dt, x, u, RO_PREF(), vel, height));
#pragma unroll
                for (int e = 0; e < 6; ++e) x[e] = xn[e];
            }
            const int kk = off + N - 1;
            const size_t b = st_idx(d, kk, 0, i);
            const double pos = d.pos_ref[(size_t)kk * es + i];
#pragma unroll
            for (int e = 0; e < 6; ++e) act[b + e * es] = x[e];
            V = __dadd_rn(V, rb_terminal_cost(c, x, pos, vel, height));
            d.tr_h[((size_t)w * d.P + p) * es + i] = 0.0;
            // RB -> RB transition is the identity (MultiPhaseDDP.cpp:374-378)
        }
        total = __dadd_rn(total, V);
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
    const int threads = ns == 1 ? 32 : std::min(128, 16 * ns);      // at most 16 staged instances (30 KB) per block.  (Two-warp blocks for the 8-trial
    // launch balance the SMs better when it is timed alone - 1.70 ms against 1.89 - but inside a solve the line search of 4096 ... 32768
    // instances was 4-10 % slower with them.)
    const int inst_per_block = threads / ns;
    const unsigned blocks = (unsigned)((d.npad + inst_per_block - 1) / inst_per_block);
    const size_t smem = (size_t)inst_per_block * ro_stride(ns) * sizeof(double);
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

// ---- cost-only pass --------------------------------------------------------------------------------------------
// forward_sweep(0) at the top of an outer (AL) iteration other than the first re-integrates the nominal trajectory with
// step size 0: u = u_nom + 0 du + K (x - x_nom) reproduces u_nom and, step by step, x_nom exactly, so the
// sweep changes nothing but the costs, which are re-evaluated with the new multipliers, penalties and barrier
// parameters (MultiPhaseDDP.cpp:178-200).  Those are independent per knot: one thread per (knot, instance) evaluates the
// knot's term from the trajectory the last line search accepted (still in its trial slot), one thread per (phase,
// instance) and then one per instance add them in the rollout's order.  The rollout sums with explicit adds for the same reason: knot term and running sum are
// then separate roundings in both kernels, and the result is bit-identical (tests: MHPC_RECOST=0 against 1).
__global__ void __launch_bounds__(128) k_recost_knots(Dev d, Consts c, Opt opt, int mask_mode) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n || !inst_enabled(d, i, mask_mode)) return;
    const int kk = blockIdx.y;
    int p, k;
    if (!knot_to_phase(d, i, kk, p, k)) return;
    const size_t es = d.npad;
    const int slot = d.si[(size_t)SI_ACC_SLOT * es + i];
    const double* st = d.st_act + act_idx(d, slot, kk, 0, i);
    const int mode = d.ph_mode[ph_idx(d, p, i)], N = d.ph_N[ph_idx(d, p, i)];
    const bool wb = p < d.n_wb[i];
    const double vel = d.vel[i], height = d.height[i], pos = d.pos_ref[(size_t)kk * es + i];
    double out;
    if (wb) {
        double x[14];
#pragma unroll
        for (int e = 0; e < 14; ++e) x[e] = st[(size_t)e * es];
        if (k < N - 1) {
            double u[4], y[4], par[6];
#pragma unroll
            for (int r = 0; r < 4; ++r) { u[r] = st[(size_t)(14 + r) * es]; y[r] = st[(size_t)(18 + r) * es]; }
#pragma unroll
            for (int q = 0; q < 6; ++q) par[q] = d.par[par_idx(d, p, PAR_DELTA_T + q, i)];
            const bool reb = d.si[(size_t)SI_REB_ON * es + i] != 0;
            out = wb_running_cost(c, mode, d.dt_wb[i], x, u, y, pos, vel, height, reb, par);
        } else {
            const bool al_term = (mode == 2 || mode == 4) && opt.AL_active;
            const double sigma = al_term ? d.par[par_idx(d, p, PAR_SIGMA, i)] : 0.0, lam = al_term ? d.par[par_idx(d, p, PAR_LAMBDA, i)] : 0.0;
            double h;
            out = wb_terminal_cost(c, mode, x, pos, vel, opt.AL_active != 0, sigma, lam, h);
            d.par[par_idx(d, p, PAR_H, i)] = h;
        }
    } else {
        double x[6];
#pragma unroll
        for (int e = 0; e < 6; ++e) x[e] = st[(size_t)e * es];
        if (k < N - 1) {
            double u[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) u[r] = st[(size_t)(14 + r) * es];
            out = rb_running_cost(c, mode, d.dt_fb[i], x, u, pos, vel, height);
        } else {
            out = rb_terminal_cost(c, x, pos, vel, height);
            d.par[par_idx(d, p, PAR_H, i)] = 0.0;
        }
    }
    d.knot_cost[(size_t)kk * es + i] = out;
}

// one thread per (phase, instance): the phase's knot terms in knot order
__global__ void __launch_bounds__(128) k_recost_phase(Dev d, int mask_mode) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n || !inst_enabled(d, i, mask_mode)) return;
    const int p = blockIdx.y;
    if (p >= d.n_phases[i]) return;
    const size_t es = d.npad;
    const int N = d.ph_N[ph_idx(d, p, i)], off = d.ph_off[ph_idx(d, p, i)];
    const double* kc = d.knot_cost + (size_t)off * es + i;
    double V = 0.0;
    int k = 0;
    for (; k + 8 <= N; k += 8) {     // eight loads in flight, added in order
        double t[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) t[q] = kc[(size_t)(k + q) * es];
#pragma unroll
        for (int q = 0; q < 8; ++q) V = __dadd_rn(V, t[q]);
    }
    for (; k < N; ++k) V = __dadd_rn(V, kc[(size_t)k * es]);
    d.tr_h[(size_t)p * es + i] = V;     // trial slot 0 of tr_h is free here (k_recost_knots wrote h straight into the parameters)
}

__global__ void __launch_bounds__(128) k_recost_sum(Dev d, int mask_mode, int lin_mode) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n || !inst_enabled(d, i, mask_mode)) return;
    const size_t es = d.npad;
    const int np = d.n_phases[i];
    double total = 0.0, viol2 = 0.0;
    for (int p = 0; p < np; ++p) {
        const int mode = d.ph_mode[ph_idx(d, p, i)];
        if (p < d.n_wb[i] && (mode == 2 || mode == 4)) { const double h = d.par[par_idx(d, p, PAR_H, i)]; viol2 = fma(h, h, viol2); }
        total = __dadd_rn(total, d.tr_h[(size_t)p * es + i]);
    }
    // what k_accept_slot0 does after a rollout; the accepted trial slot stays the one the last line search chose
    d.sd[(size_t)SD_COST * es + i] = total;
    d.sd[(size_t)SD_VIOL * es + i] = sqrt(viol2);
    d.si[(size_t)SI_AL_PARTIALS * es + i] = lin_mode;
}

void launch_recost(const Dev& d, const Consts& c, const Opt& opt, int mask_mode, int lin_mode, cudaStream_t s) {
    const dim3 grid((unsigned)((d.n + 127) / 128), (unsigned)d.KT);
    k_recost_knots<<<grid, 128, 0, s>>>(d, c, opt, mask_mode);
    k_recost_phase<<<dim3((unsigned)((d.n + 127) / 128), (unsigned)d.P), 128, 0, s>>>(d, mask_mode);
    k_recost_sum<<<(unsigned)((d.n + 127) / 128), 128, 0, s>>>(d, mask_mode, lin_mode);
}

}  // namespace mhpc
