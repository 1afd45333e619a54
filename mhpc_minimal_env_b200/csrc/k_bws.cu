// Impact-aware Riccati backward sweep (MultiPhaseDDP.cpp:100-127, 300-341; SinglePhase.cpp:183-216;
// MHPC_CompoundTypes.h:116-144) with the regularisation retry loop of MultiPhaseDDP::solve
// (MultiPhaseDDP.cpp:209-227) folded in.
//
// EIGHT lanes own one instance (four instances per warp, one warp per CTA).  In the whole-body phases sub-lane m
// (m < 7) owns rows m and m+7 of the value Hessian H in registers; in the rigid-body phases sub-lane m (m < 6) owns
// row m.  Two rows per lane means every shared operand of the 14x14 products (a column of A, a row of M = H A, a row of
// Qux) is read from shared memory once and used twice, and a warp instruction serves four instances: the kernel is
// issue- and latency-bound (712 dependent knots, ~1 warp per scheduler at 4096 instances), so instructions per instance
// are what counts.  Time is sequential; every warp streams the knot records of its four instances into warp-private
// shared memory with 16-byte cp.async, one knot ahead of the math (double buffered, synchronised with __syncwarp only:
// warps never wait for each other, there is no CTA barrier in the sweep).
// The structure of the reference's matrices is used instead of dense 14x14 products:
//   A = [I7  dt*I7 ; 0(7x2) Ab(7x12)],  B = [0 ; Bb],  C/D = two stance rows,  lxx/luu diagonal,
// which removes only exact-zero / exact-one terms of the dense sums; every stored entry is produced by the same
// operation sequence (ascending inner index) as in the reference, whichever lane computes it.
#include <cuda_pipeline.h>

#include <algorithm>

#include "kernels.h"
#include "model.cuh"

#include "bws_shared.cuh"

namespace mhpc {

constexpr int BWS_LPI = 8;                   // lanes per instance
constexpr int BWS_THREADS = 32;              // one warp per CTA
constexpr int BWS_INST = BWS_THREADS / BWS_LPI;
constexpr size_t BWS_SMEM = (size_t)BWS_REGION * BWS_INST * sizeof(double);
#ifndef BWS_MINBLOCKS
#define BWS_MINBLOCKS 7
#endif

// Work assignment.  task_mode == 0: lane group q4 of block b owns instance 4 b + q4, enters with the instance's current
// regularisation and walks the ladder reg <- max(reg * update_regularization, 1e-3) inside the kernel (at most
// `max_attempts` rungs when the ladder scheduler is on: what is left goes to the scheduler's rounds below).
// task_mode == 1: lane group = one task (instance, rung, regularisation) of the list built by k_lad_plan; tasks that are
// not the first of their instance write K, du, G into the task-indexed scratch arrays instead of the instance's own.
template <int task_mode, int dump = 0>   // dump: also write H, G, Qx, Qu, Qxx, Qux, Quu of every knot to d.ctg_dump (replay of one instance)
__global__ void __launch_bounds__(BWS_THREADS, BWS_MINBLOCKS) k_bws(Dev d, Consts c, Opt opt, int mask_mode, int single_attempt, int max_attempts) {
    extern __shared__ __align__(16) double smem[];
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x;
    const int q4 = tid >> 3, m = tid & 7;           // instance slot in the warp, sub-lane
    double* bufs = smem + (size_t)q4 * BWS_REGION;                     // [2 stages][REC] of my instance
    double* scr = bufs + 2 * REC;
    const size_t es = d.npad;
    int i = blockIdx.x * BWS_INST + q4, rung = 0, stride = 0;          // rung: 0-based attempt index of this sweep (n_bws = rung + 1)
    bool valid, enabled, to_scratch = false;
    double reg;
    if (!task_mode) {
        valid = i < d.n;
        if (!valid) i = 0;
        enabled = valid && inst_enabled(d, i, mask_mode);
        reg = enabled ? d.sd[(size_t)SD_REG * es + i] : 0.0;
    } else {
        const int nt = d.lad_cnt[LAD_N_TASKS];
        if (blockIdx.x * BWS_INST >= nt) return;                       // the whole warp is beyond the task list
        const int ti = i < nt ? d.lad_t_inst[i] : -1;
        valid = enabled = ti >= 0;
        reg = valid ? d.lad_t_reg[i] : 0.0;
        rung = valid ? d.lad_t_rung[i] : 0;
        stride = d.lad_cnt[LAD_STRIDE];
        to_scratch = (i % d.lad_cnt[LAD_S]) != 0;
    }
    const int task = blockIdx.x * BWS_INST + q4;
    if (task_mode) i = valid ? d.lad_t_inst[task] : 0;
    const int np = valid ? d.n_phases[i] : 0, nwb = valid ? d.n_wb[i] : 0, nk = enabled ? d.n_knots[i] : 0;
    const double dtw = valid ? d.dt_wb[i] : 0.0, dtf = valid ? d.dt_fb[i] : 0.0;
    bool done = !enabled;
    int ktb = max(nk, __shfl_xor_sync(FULL, nk, 8));                   // knots this warp walks
    ktb = max(ktb, __shfl_xor_sync(FULL, ktb, 16));
    const double* lin_i = d.lin + (size_t)i * d.KT * REC;
    const bool rowl = m < 7;                                           // this sub-lane owns WB rows m and m + 7
    // lane-constant parts of the output addresses (SoA over instances: the four instances of a warp fill one 32-byte
    // sector per element): only `kk * stride` is added per knot
    // (the scratch arrays have the element stride of the instance arrays, npad, so that only the base and the column differ)
    const size_t ocol = to_scratch ? (size_t)lad_scratch_col(task, d.lad_cnt[LAD_S]) : (size_t)i;
    double* const grec = (to_scratch ? d.lad_gain : d.gain) + gain_idx(d, 0, ocol);     // [KT][GREC] of this instance / scratch column
    double* const gv_lane = grec + GR_G + m;
    double* const kf_lane = grec + GR_K + 4 * m;
    double* const du_lane = grec + GR_DU + m;
    const size_t gv_knot = gain_knot_stride(d), kf_knot = gv_knot, du_knot = gv_knot;

    // per-instance schedule walker (phase p, in-phase knot k; knot slot = p*KSTRIDE + k), used both by the prefetcher
    // (one knot ahead) and by the math; it visits the nk used knots from the last one backwards
    struct Walk { int p, N, mode, k; };
    auto walk_load = [&](Walk& w) {
        w.N = d.ph_N[ph_idx(d, w.p, i)];
        w.mode = d.ph_mode[ph_idx(d, w.p, i)];
    };
    auto walk_init = [&](Walk& w) {
        w.p = np - 1; w.N = 0; w.mode = 1; w.k = 0;
        if (enabled && np > 0) { walk_load(w); w.k = w.N - 1; }
    };
    auto walk_step = [&](Walk& w) {   // to the previous knot
        if (--w.k < 0 && w.p > 0) { --w.p; walk_load(w); w.k = w.N - 1; }
    };
    auto prefetch = [&](Walk& w, int it_next, double* dst) {
        if (enabled && it_next < nk) {
            const int len = knot_rec_len(w.p < nwb, w.k == w.N - 1);
            const double* src = lin_i + (size_t)(w.p * KSTRIDE + w.k) * REC;
            for (int q = m; 2 * q < len; q += BWS_LPI) __pipeline_memcpy_async(dst + 2 * q, src + 2 * q, 16);
            walk_step(w);
        }
        __pipeline_commit();
    };

    while (true) {
        // ---- one attempt -----------------------------------------------------------------
        double Ha[14], Hb[14];
#pragma unroll
        for (int q = 0; q < 14; ++q) { Ha[q] = 0.0; Hb[q] = 0.0; }
        scr[S_G + m] = 0.0;
        scr[S_G + m + 8] = 0.0;
        double dV = 0.0;
        bool failed = false;
        int fail_it = 0;
        Walk wk, wp;
        walk_init(wk); walk_init(wp);
        int& p = wk.p; int& Nn = wk.N; int& mode = wk.mode;
        __syncwarp();
        if (ktb > 0) prefetch(wp, 0, bufs);
        for (int it = 0; it < ktb; ++it) {
            const double* buf = bufs + (size_t)(it & 1) * REC;
            double* nbuf = bufs + (size_t)((it & 1) ^ 1) * REC;
            __pipeline_wait_prior(0);
            __syncwarp();
            const int alive = (!done && !failed) ? 1 : 0;
            if (!__any_sync(FULL, alive)) break;
            prefetch(wp, it + 1, nbuf);

            // ---- which kind of knot is this for my instance? ----
            const int kk = wk.p * KSTRIDE + wk.k;   // knot slot of this step
            int ty = 0;  // 0 idle, 1 WB step, 2 RB step, 3 WB terminal, 4 RB terminal
            if (alive && it < nk) {
                const bool wb = p < nwb, term = (wk.k == Nn - 1);
                ty = wb ? (term ? 3 : 1) : (term ? 4 : 2);
            }
            const unsigned b1 = __ballot_sync(FULL, ty == 1), b2 = __ballot_sync(FULL, ty == 2), bt34 = __ballot_sync(FULL, ty >= 3);
#define RBUF(e) buf[(e)]

            // ============ terminal knots: value reset across the phase boundary ============
            if (bt34) {
                const bool on = (ty == 3 || ty == 4);
                const bool last = on && (p == np - 1);
                const bool wbp = (ty == 3);
                const bool next_rb = on && !last && (p + 1 >= nwb);
                // (a) WB <- RB: embed the 6-dim value into 14 dims through P^T (.) P  (MultiPhaseDDP.cpp:330-331)
                const bool embed = wbp && next_rb;
                double goa = 0.0, gob = 0.0;
                if (embed) {
                    if (m < 6) {
#pragma unroll
                        for (int q = 0; q < 6; ++q) scr[S_MQ + m * MQ_LD + q] = Ha[q];
                    }
                    if (rowl) {
                        const int pa = wb_to_rb(m), pb = wb_to_rb(m + 7);
                        goa = (pa >= 0) ? scr[S_G + pa] : 0.0;
                        gob = (pb >= 0) ? scr[S_G + pb] : 0.0;
                    }
                }
                __syncwarp();
                if (embed && rowl) { embed_row(m, scr, Ha); embed_row(m + 7, scr, Hb); }
                __syncwarp();
                if (embed && rowl) { scr[S_G + m] = goa; scr[S_G + m + 7] = gob; }
                __syncwarp();
                // (b) WB phase end: G = Px^T G', H = Px^T H' Px  (MultiPhaseDDP.cpp:320-326), Px = [I7 0; 0 Pb]
                const bool px = wbp && !last;
                if (px && rowl) { px_right_row(m, buf, scr, Ha); px_right_row(m + 7, buf, scr, Hb); }
                __syncwarp();
                double gna = 0.0, gnb = 0.0;
                if (px && rowl) { gna = px_left_row(m, buf, scr, Ha); gnb = px_left_row(m + 7, buf, scr, Hb); }
                __syncwarp();
                if (px && rowl) { scr[S_G + m] = gna; scr[S_G + m + 7] = gnb; }
                __syncwarp();
                // (c) terminal cost: G = Phix + Gnext, H = Phixx + Hnext  (SinglePhase.cpp:190-191)
                if (on) {
                    if (last) {
#pragma unroll
                        for (int q = 0; q < 14; ++q) { Ha[q] = 0.0; Hb[q] = 0.0; }
                    }
                    if (wbp) {
                        if (rowl) {
                            const double ga = wb_terminal_row(m, buf, last ? 0.0 : scr[S_G + m], Ha);
                            const double gb = wb_terminal_row(m + 7, buf, last ? 0.0 : scr[S_G + m + 7], Hb);
                            scr[S_G + m] = ga; scr[S_G + m + 7] = gb;
                            gv_lane[(size_t)kk * gv_knot] = ga;
                            gv_lane[(size_t)kk * gv_knot + 7] = gb;
                            if (dump) {
                                double* cg = d.ctg_dump + (size_t)kk * CTGREC;
                                cg[CG_G + m] = ga; cg[CG_G + m + 7] = gb;
#pragma unroll
                                for (int jc = 0; jc < 14; ++jc) { cg[CG_H + m + 14 * jc] = Ha[jc]; cg[CG_H + m + 7 + 14 * jc] = Hb[jc]; }
                            }
                        }
                    } else if (m < 6) {
                        double gv = last ? 0.0 : scr[S_G + m];
#pragma unroll
                        for (int jc = 0; jc < 6; ++jc) Ha[jc] = ((jc == m) ? c.rbQf[m] : 0.0) + Ha[jc];
                        gv = RBUF(R_RPHIX + m) + gv;
                        scr[S_G + m] = gv;
                        gv_lane[(size_t)kk * gv_knot] = gv;
                        if (dump) {
                            double* cg = d.ctg_dump + (size_t)kk * CTGREC;
                            cg[CG_G + m] = gv;
#pragma unroll
                            for (int jc = 0; jc < 6; ++jc) cg[CG_H + m + 14 * jc] = Ha[jc];
                        }
                    }
                }
                __syncwarp();
            }

            // ============ whole-body step ============
            if (b1) {
                const bool on = (ty == 1);
                const bool rw = on && rowl;
                const double dt = dtw;
                const int la = m, lb = m + 7;
                double Ma[14];
                if (rw) {
                    // M = H A, rows la and lb; rows 7..13 are shared through scratch.  Columns go in blocks of four so that
                    // eight accumulation chains (4 columns x 2 rows) are in flight; each chain keeps the ascending-s order.
                    Ma[0] = Ha[0]; Ma[1] = Ha[1];
                    scr[S_MQ + m * MQ_LD + 0] = Hb[0]; scr[S_MQ + m * MQ_LD + 1] = Hb[1];
#pragma unroll
                    for (int j0 = 2; j0 < 14; j0 += 4) {
                        double xa[4], xb[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int jc = j0 + k;
                            xa[k] = (jc < 7) ? Ha[jc] : Ha[jc - 7] * dt;
                            xb[k] = (jc < 7) ? Hb[jc] : Hb[jc - 7] * dt;
                        }
#pragma unroll
                        for (int s = 0; s < 7; ++s) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const double e = RBUF(R_AB(s, j0 + k - 2));
                                xa[k] += Ha[7 + s] * e;
                                xb[k] += Hb[7 + s] * e;
                            }
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) { Ma[j0 + k] = xa[k]; scr[S_MQ + m * MQ_LD + j0 + k] = xb[k]; }
                    }
                    {   // N = H B (rows 7..13 are all B^T N needs)
                        double nn[4];
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) nn[cc] = Hb[7] * RBUF(R_BB(0, cc));
#pragma unroll
                        for (int s = 1; s < 7; ++s)
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) nn[cc] += Hb[7 + s] * RBUF(R_BB(s, cc));
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) scr[S_N + m * 4 + cc] = nn[cc];
                    }
                }
                __syncwarp();
                double quxa[4] = {0, 0, 0, 0}, quxb[4] = {0, 0, 0, 0}, qxa = 0.0, qxb = 0.0;
                if (on) {
                    const double L00 = RBUF(R_LYY), L01 = RBUF(R_LYY + 1), L11 = RBUF(R_LYY + 2);
                    const double ly0 = RBUF(R_LY), ly1 = RBUF(R_LY + 1);
                    if (rowl) {
                        double aba[7], abb[7];
#pragma unroll
                        for (int s = 0; s < 7; ++s) { aba[s] = (la >= 2) ? RBUF(R_AB(s, la - 2)) : 0.0; abb[s] = RBUF(R_AB(s, lb - 2)); }
                        const double c0a = (la >= 2) ? RBUF(R_CB(0, la - 2)) : 0.0, c1a = (la >= 2) ? RBUF(R_CB(1, la - 2)) : 0.0;
                        const double c0b = RBUF(R_CB(0, lb - 2)), c1b = RBUF(R_CB(1, lb - 2));
                        const double w0a = c0a * L00 + c1a * L01, w1a = c0a * L01 + c1a * L11;   // row of C^T lyy
                        const double w0b = c0b * L00 + c1b * L01, w1b = c0b * L01 + c1b * L11;
                        const double lxxda = (la >= 3) ? RBUF(R_LXXQ + (la - 3)) : (2 * dt) * c.wbQ[la];
                        const double lxxdb = (2 * dt) * c.wbQ[lb];
                        // Qxx rows = (lxx + C^T lyy C) + A^T M, four columns (eight chains) at a time
                        double Qa[14], Qb[14];
#pragma unroll
                        for (int j0 = 0; j0 < 14; j0 += 4) {
                            double qa[4], qb[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (j0 + k < 14) { qa[k] = Ma[j0 + k]; qb[k] = Ma[j0 + k] * dt; }
                            }
#pragma unroll
                            for (int s = 0; s < 7; ++s) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    if (j0 + k < 14) {
                                        const double mr = scr[S_MQ + s * MQ_LD + j0 + k];
                                        if (la >= 2) qa[k] += aba[s] * mr;
                                        qb[k] += abb[s] * mr;
                                    }
                                }
                            }
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int jc = j0 + k;
                                if (jc < 14) {
                                    double lca = (jc == la) ? lxxda : 0.0, lcb = (jc == lb) ? lxxdb : 0.0;
                                    if (jc >= 2) {
                                        const double cj0 = RBUF(R_CB(0, jc - 2)), cj1 = RBUF(R_CB(1, jc - 2));
                                        if (la >= 2) lca = lca + (w0a * cj0 + w1a * cj1);
                                        lcb = lcb + (w0b * cj0 + w1b * cj1);
                                    }
                                    Qa[jc] = lca + qa[k];
                                    if (jc == la) Qa[jc] += 1.0 * reg;
                                    Qb[jc] = lcb + qb[k];
                                    if (jc == lb) Qb[jc] += 1.0 * reg;
                                }
                            }
                        }
#pragma unroll
                        for (int jc = 0; jc < 14; ++jc) { Ha[jc] = Qa[jc]; Hb[jc] = Qb[jc]; }
                        // Qx = (lx + A^T G) + C^T ly
                        double aga, agb;
                        if (la < 2) aga = scr[S_G + la];
                        else {
                            aga = scr[S_G + la];
#pragma unroll
                            for (int s = 0; s < 7; ++s) aga += aba[s] * scr[S_G + 7 + s];
                        }
                        agb = scr[S_G + la] * dt;
#pragma unroll
                        for (int s = 0; s < 7; ++s) agb += abb[s] * scr[S_G + 7 + s];
                        qxa = (RBUF(R_LX + la) + aga) + (c0a * ly0 + c1a * ly1);
                        qxb = (RBUF(R_LX + lb) + agb) + (c0b * ly0 + c1b * ly1);
                        // Qux columns la, lb = (0 + D^T lyy C) + B^T M
                        {
                            double xa[4], xb[4];
                            const double m0a = scr[S_MQ + 0 * MQ_LD + la], m0b = scr[S_MQ + 0 * MQ_LD + lb];
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) { const double b0 = RBUF(R_BB(0, cc)); xa[cc] = b0 * m0a; xb[cc] = b0 * m0b; }
#pragma unroll
                            for (int s = 1; s < 7; ++s) {
                                const double msa = scr[S_MQ + s * MQ_LD + la], msb = scr[S_MQ + s * MQ_LD + lb];
#pragma unroll
                                for (int cc = 0; cc < 4; ++cc) {
                                    const double bs = RBUF(R_BB(s, cc));
                                    xa[cc] += bs * msa;
                                    xb[cc] += bs * msb;
                                }
                            }
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) {
                                const double d0 = RBUF(R_DB(0, cc)), d1 = RBUF(R_DB(1, cc));
                                const double v0 = d0 * L00 + d1 * L01, v1 = d0 * L01 + d1 * L11;
                                quxa[cc] = (0.0 + (v0 * c0a + v1 * c1a)) + xa[cc];
                                quxb[cc] = (0.0 + (v0 * c0b + v1 * c1b)) + xb[cc];
                            }
                        }
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {   // Quu entries (row l&3, col l>>2), l = m and m + 8: (luu + D^T lyy D) + B^T N
                        const int l = m + 8 * h;
                        const int rr = l & 3, cc = l >> 2;
                        double acc = RBUF(R_BB(0, rr)) * scr[S_N + 0 * 4 + cc];
#pragma unroll
                        for (int s = 1; s < 7; ++s) acc += RBUF(R_BB(s, rr)) * scr[S_N + s * 4 + cc];
                        const double d0 = RBUF(R_DB(0, rr)), d1 = RBUF(R_DB(1, rr));
                        const double v0 = d0 * L00 + d1 * L01, v1 = d0 * L01 + d1 * L11;
                        double q = ((rr == cc ? RBUF(R_LUU + rr) : 0.0) + (v0 * RBUF(R_DB(0, cc)) + v1 * RBUF(R_DB(1, cc)))) + acc;
                        if (rr == cc) q += 1.0 * reg;
                        scr[S_QUU + rr + 4 * cc] = q;
                    }
                    if (m < 4) {   // Qu = (lu + B^T G) + D^T ly
                        double acc = RBUF(R_BB(0, m)) * scr[S_G + 7];
#pragma unroll
                        for (int s = 1; s < 7; ++s) acc += RBUF(R_BB(s, m)) * scr[S_G + 7 + s];
                        scr[S_QU + m] = (RBUF(R_LU + m) + acc) + (RBUF(R_DB(0, m)) * ly0 + RBUF(R_DB(1, m)) * ly1);
                    }
                }
                __syncwarp();   // everyone is done reading M
                if (rw) {
#pragma unroll
                    for (int jc = 0; jc < 14; ++jc) { scr[S_MQ + la * MQ_LD + jc] = Ha[jc]; scr[S_MQ + lb * MQ_LD + jc] = Hb[jc]; }
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) { scr[S_QUX + cc * 14 + la] = quxa[cc]; scr[S_QUX + cc * 14 + lb] = quxb[cc]; }
                }
                __syncwarp();
                if (rw) {
#pragma unroll
                    for (int jc = 0; jc < 14; ++jc) {   // Qxx = (Qxx + Qxx^T)/2
                        Ha[jc] = (Ha[jc] + scr[S_MQ + jc * MQ_LD + la]) / 2;
                        Hb[jc] = (Hb[jc] + scr[S_MQ + jc * MQ_LD + lb]) / 2;
                    }
                }
                if (dump && on) {   // the Q-function as CostToGoStruct holds it after valuefunction_update: Qxx regularised and symmetrised, Quu regularised
                    double* cg = d.ctg_dump + (size_t)kk * CTGREC;
                    if (rowl) {
#pragma unroll
                        for (int jc = 0; jc < 14; ++jc) { cg[CG_QXX + la + 14 * jc] = Ha[jc]; cg[CG_QXX + lb + 14 * jc] = Hb[jc]; }
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) { cg[CG_QUX + cc + 4 * la] = quxa[cc]; cg[CG_QUX + cc + 4 * lb] = quxb[cc]; }
                        cg[CG_QX + la] = qxa; cg[CG_QX + lb] = qxb;
                    }
                    cg[CG_QUU + m] = scr[S_QUU + m]; cg[CG_QUU + m + 8] = scr[S_QUU + m + 8];
                    if (m < 4) cg[CG_QU + m] = scr[S_QU + m];
                }
                double ka[4], kb[4], du[4], goa = 0.0, gob = 0.0, dvs = 0.0;
                __syncwarp();
                const bool ok = value_update<14, true, BWS_LPI>(on, rw, m, scr, quxa, quxb, qxa, qxb, Ha, Hb, opt.quu_shift, ka, kb, du, goa, gob, dvs);
                if (on) {   // (no hand-over needed here or after the stores: S_G is not read by value_update, and the next reader
                            //  of anything written below sits behind the __syncwarp at the top of the knot loop)
                    if (!ok) { failed = true; fail_it = it; }
                    else {
                        dV += dvs;
                        if (dump && rowl) {
                            double* cg = d.ctg_dump + (size_t)kk * CTGREC;
                            cg[CG_G + la] = goa; cg[CG_G + lb] = gob;
#pragma unroll
                            for (int jc = 0; jc < 14; ++jc) { cg[CG_H + la + 14 * jc] = Ha[jc]; cg[CG_H + lb + 14 * jc] = Hb[jc]; }
                        }
                        if (rowl) {
                            scr[S_G + la] = goa; scr[S_G + lb] = gob;
                            gv_lane[(size_t)kk * gv_knot] = goa;
                            gv_lane[(size_t)kk * gv_knot + 7] = gob;
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) {
                                kf_lane[(size_t)kk * kf_knot + cc] = ka[cc];
                                kf_lane[(size_t)kk * kf_knot + 28 + cc] = kb[cc];
                            }
                        }
                        if (m < 4) du_lane[(size_t)kk * du_knot] = du[m];
                    }
                }
            }

            // ============ rigid-body step ============
            // A = I + dt*[0 I3; 0 0] + e5 (a0, a1, 0, 0, 0, 0), B has two non-zeros per column (rows 3 + (c&1) and 5):
            // only exact-zero terms of the reference's dense 6x6 products are dropped, the remaining terms keep the
            // ascending inner-index order.  Sub-lane m < 6 owns row m (in Ha).
            if (b2) {
                const bool on = (ty == 2);
                const bool rw = on && m < 6;
                const int l = m;
                const double dt = dtf;
                const double a0 = on ? RBUF(R_RA) : 0.0, a1 = on ? RBUF(R_RA + 1) : 0.0;
                double bt[4], b5[4];   // B(3 + (c&1), c), B(5, c)
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) { bt[cc] = on ? RBUF(R_RB + 2 * cc) : 0.0; b5[cc] = on ? RBUF(R_RB + 2 * cc + 1) : 0.0; }
                if (rw) {
                    // M = H A (row l)
                    scr[S_MQ + l * MQ_LD + 0] = Ha[0] + Ha[5] * a0;
                    scr[S_MQ + l * MQ_LD + 1] = Ha[1] + Ha[5] * a1;
                    scr[S_MQ + l * MQ_LD + 2] = Ha[2];
                    scr[S_MQ + l * MQ_LD + 3] = Ha[0] * dt + Ha[3];
                    scr[S_MQ + l * MQ_LD + 4] = Ha[1] * dt + Ha[4];
                    scr[S_MQ + l * MQ_LD + 5] = Ha[2] * dt + Ha[5];
                    // N = H B (row l)
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) scr[S_N + l * 4 + cc] = Ha[3 + (cc & 1)] * bt[cc] + Ha[5] * b5[cc];
                }
                __syncwarp();
                double qux[4] = {0, 0, 0, 0}, qx = 0.0;
                if (on) {
                    if (l < 6) {
                        const double acol = (l == 0) ? a0 : a1;
#pragma unroll
                        for (int jc = 0; jc < 6; ++jc) {   // Qxx row l = lxx + A^T M
                            double q;
                            if (l < 2) q = scr[S_MQ + l * MQ_LD + jc] + acol * scr[S_MQ + 5 * MQ_LD + jc];
                            else if (l == 2) q = scr[S_MQ + 2 * MQ_LD + jc];
                            else q = dt * scr[S_MQ + (l - 3) * MQ_LD + jc] + scr[S_MQ + l * MQ_LD + jc];
                            Ha[jc] = ((jc == l) ? (2 * dt) * c.rbQ[l] : 0.0) + q;
                            if (jc == l) Ha[jc] += 1.0 * reg;
                        }
                        double ag;
                        if (l < 2) ag = scr[S_G + l] + acol * scr[S_G + 5];
                        else if (l == 2) ag = scr[S_G + 2];
                        else ag = dt * scr[S_G + l - 3] + scr[S_G + l];
                        qx = RBUF(R_RLX + l) + ag;
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc)   // Qux column l = B^T M
                            qux[cc] = bt[cc] * scr[S_MQ + (3 + (cc & 1)) * MQ_LD + l] + b5[cc] * scr[S_MQ + 5 * MQ_LD + l];
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int e = m + 8 * h;
                        const int rr = e & 3, cc = e >> 2;
                        const double btr = (rr == 0) ? bt[0] : (rr == 1) ? bt[1] : (rr == 2) ? bt[2] : bt[3];
                        const double b5r = (rr == 0) ? b5[0] : (rr == 1) ? b5[1] : (rr == 2) ? b5[2] : b5[3];
                        const double acc = btr * scr[S_N + (3 + (rr & 1)) * 4 + cc] + b5r * scr[S_N + 5 * 4 + cc];
                        double q = ((rr == cc) ? (2 * dt) * c.rbR[mode - 1][rr] : 0.0) + acc;
                        if (rr == cc) q += 1.0 * reg;
                        scr[S_QUU + rr + 4 * cc] = q;
                    }
                    if (m < 4) {
                        const double btm = (m == 0) ? bt[0] : (m == 1) ? bt[1] : (m == 2) ? bt[2] : bt[3];
                        const double b5m = (m == 0) ? b5[0] : (m == 1) ? b5[1] : (m == 2) ? b5[2] : b5[3];
                        scr[S_QU + m] = RBUF(R_RLU + m) + (btm * scr[S_G + 3 + (m & 1)] + b5m * scr[S_G + 5]);
                    }
                }
                __syncwarp();
                if (rw) {
#pragma unroll
                    for (int jc = 0; jc < 6; ++jc) scr[S_MQ + l * MQ_LD + jc] = Ha[jc];
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) scr[S_QUX + cc * 14 + l] = qux[cc];
                }
                __syncwarp();
                if (rw) {
#pragma unroll
                    for (int jc = 0; jc < 6; ++jc) Ha[jc] = (Ha[jc] + scr[S_MQ + jc * MQ_LD + l]) / 2;
                }
                if (dump && on) {
                    double* cg = d.ctg_dump + (size_t)kk * CTGREC;
                    if (l < 6) {
#pragma unroll
                        for (int jc = 0; jc < 6; ++jc) cg[CG_QXX + l + 14 * jc] = Ha[jc];
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) cg[CG_QUX + cc + 4 * l] = qux[cc];
                        cg[CG_QX + l] = qx;
                    }
                    cg[CG_QUU + m] = scr[S_QUU + m]; cg[CG_QUU + m + 8] = scr[S_QUU + m + 8];
                    if (m < 4) cg[CG_QU + m] = scr[S_QU + m];
                }
                double ka[4], kb[4], du[4], gout = 0.0, gdummy = 0.0, dvs = 0.0;
                __syncwarp();
                const bool ok = value_update<6, false, BWS_LPI>(on, rw, m, scr, qux, qux, qx, qx, Ha, Hb, opt.quu_shift, ka, kb, du, gout, gdummy, dvs);
                if (on) {
                    if (!ok) { failed = true; fail_it = it; }
                    else {
                        dV += dvs;
                        if (dump && l < 6) {
                            double* cg = d.ctg_dump + (size_t)kk * CTGREC;
                            cg[CG_G + l] = gout;
#pragma unroll
                            for (int jc = 0; jc < 6; ++jc) cg[CG_H + l + 14 * jc] = Ha[jc];
                        }
                        if (l < 6) {
                            scr[S_G + l] = gout;
                            gv_lane[(size_t)kk * gv_knot] = gout;
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) kf_lane[(size_t)kk * kf_knot + cc] = ka[cc];
                        }
                        if (m < 4) du_lane[(size_t)kk * du_knot] = du[m];
                    }
                }
            }
#undef RBUF
            if (enabled && it < nk) walk_step(wk);
        }
        __pipeline_wait_prior(0);
        __syncwarp();

        // ---- outcome of the attempt (MultiPhaseDDP.cpp:209-241) ----
        int retry = 0;
        if (!dump && !done && d.bws_trace && m == 0 && rung < BWS_TRACE_ATT) {
            const int r = d.si[(size_t)SI_N_REC * es + i];
            if (r < d.max_rec) d.bws_trace[((size_t)r * BWS_TRACE_ATT + rung) * es + i] = failed ? fail_it + 1 : -1;
        }
        if (!done) {
            if (task_mode) {
                if (!failed) {
                    if (m == 0) { d.lad_t_rung[task] = rung; d.lad_t_reg[task] = reg; d.lad_t_dv[task] = dV; }
                    done = true;
                } else {
                    // a chain of the final round goes on with every `stride`-th rung of the ladder until one passes or the
                    // ladder runs past 1000 (a rung the reference never attempts); single-rung tasks just report
                    bool more = stride > 0;
                    double nr = reg;
                    for (int k = 0; k < stride && more; ++k) { nr = fmax(nr * opt.update_regularization, 1e-03); if (nr > 1000) more = false; }
                    if (more) { reg = nr; rung += stride; retry = 1; }
                    else { if (m == 0) d.lad_t_rung[task] = -1; done = true; }
                }
            } else if (single_attempt) {
                if (m == 0) {
                    d.si[(size_t)SI_BWS_OK * es + i] = failed ? 0 : 1;
                    if (!failed) d.sd[(size_t)SD_DV * es + i] = dV;
                    if (d.lad_unres) d.lad_unres[i] = 0;
                }
                done = true;
            } else if (failed) {
                reg = fmax(reg * opt.update_regularization, 1e-03);
                ++rung;
                if (reg > 1000) {
                    if (m == 0) {
                        d.si[(size_t)SI_STATUS * es + i] = STATUS_REG_ABORT; d.si[(size_t)SI_DDP_ACTIVE * es + i] = 0; d.si[(size_t)SI_N_BWS * es + i] = rung + 1;
                        if (d.lad_unres) d.lad_unres[i] = 0;
                    }
                    done = true;
                } else if (max_attempts > 0 && rung >= max_attempts) {   // the rest of the ladder goes to the scheduler's rounds
                    if (m == 0) { d.lad_unres[i] = 1; d.lad_rung[i] = rung; d.si[(size_t)SI_BWS_OK * es + i] = 0; }
                    done = true;
                } else retry = 1;
            } else {
                if (m == 0) {
                    d.sd[(size_t)SD_DV * es + i] = dV;
                    d.sd[(size_t)SD_REG_USED * es + i] = reg;
                    double r2 = reg / 20;
                    if (r2 < 1e-06) r2 = 0;
                    d.sd[(size_t)SD_REG * es + i] = r2;
                    d.si[(size_t)SI_N_BWS * es + i] = rung + 1;
                    d.si[(size_t)SI_BWS_OK * es + i] = 1;
                    if (d.lad_unres) d.lad_unres[i] = 0;
                }
                done = true;
            }
        }
        if (!__any_sync(FULL, retry)) break;
    }
}

// ---- regularisation-ladder scheduler ---------------------------------------------------------------------------------
// MultiPhaseDDP::solve retries a failed backward sweep with reg <- max(2 reg, 1e-3) until it passes (MultiPhaseDDP.cpp:
// 209-227).  The ladder is deterministic, so its rungs can be tested side by side: the first rung IN LADDER ORDER that
// passes is the reference's result whichever order the rungs were evaluated in (no monotonicity is assumed: every rung
// below the winner is actually run and seen to fail).  A sweep stage is: one launch that walks the first rungs in the
// kernel (all instances), then rounds of {plan, task launch, resolve, copy} over the instances that still have no passing
// rung: the plan packs them densely and gives each S = floor(resident lane groups / unresolved instances) consecutive
// rungs, so the SMs the finished instances left idle test the next rungs in advance.  The last round (and any round with
// S >= 4) lets each of the S tasks of an instance walk every S-th rung to the end of the ladder, which terminates the stage
// without a host round trip.  n_bws, the accepted reg and dV are those of the sequential ladder, bit for bit.
__device__ __forceinline__ double ladder_reg(double reg0, int rung, double upd) {   // regularisation of attempt `rung` of a sweep entered with reg0
    double r = reg0;
    for (int k = 0; k < rung; ++k) r = fmax(r * upd, 1e-03);
    return r;
}

__global__ void __launch_bounds__(1024) k_lad_plan(Dev d, Opt opt, int mask_mode, int last_round, int smax) {
    __shared__ int s_scan[1024];
    const int tid = threadIdx.x;
    const int per = (d.n + 1023) / 1024;
    const int lo = min(d.n, tid * per), hi = min(d.n, lo + per);
    int cnt = 0;
    for (int i = lo; i < hi; ++i) cnt += (d.lad_unres[i] != 0 && inst_enabled(d, i, mask_mode)) ? 1 : 0;
    s_scan[tid] = cnt;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {   // inclusive scan: the unresolved instances keep their slot order
        const int v = tid >= off ? s_scan[tid - off] : 0;
        __syncthreads();
        s_scan[tid] += v;
        __syncthreads();
    }
    int total = s_scan[1023];
    int cidx = s_scan[tid] - cnt;
    // A little more than one resident wave of single-rung tasks costs a second wave (2621 tasks: 1.68 ms, 2298: 1.20 ms): the
    // excess instances (highest slots) sit this round out - they stay unresolved at their rung and join the next plan.
    const int cap = min(d.lad_cap, d.npad);
    const bool defer = !last_round && total > cap && 4 * total <= 5 * cap;
    if (defer) total = cap;
    // rungs per instance this round.  Latency regime (few unresolved instances): as many as fill the resident capacity - they
    // cost no time.  Throughput regime (lad_target > 0): about one batch worth of tasks per round, because a round cannot
    // be shorter than one sweep whatever its size; fewer, fuller rounds win although some speculative rungs are wasted.
    int S = 1;
    if (total > 0) {
        const int s_cap = min(d.lad_cap, d.npad) / total;
        const int s_tgt = d.lad_target > 0 ? (d.lad_target + total / 2) / total : 0;
        S = max(1, min(smax, max(s_cap, s_tgt)));
        while (S > 1 && ((long long)total * (S - 1) > d.npad || (long long)total * S > d.lad_tcap)) --S;   // scratch columns, task slots
    }
    if (tid == 0) {
        d.lad_cnt[LAD_N_TASKS] = total * S;
        d.lad_cnt[LAD_N_UNRES] = total;
        d.lad_cnt[LAD_STRIDE] = last_round ? S : 0;
        d.lad_cnt[LAD_S] = S;
    }
    for (int i = lo; i < hi; ++i) {
        if (!(d.lad_unres[i] != 0 && inst_enabled(d, i, mask_mode))) continue;
        if (cidx >= total) break;           // deferred to the next round
        d.lad_u_inst[cidx] = i;
        d.lad_u_copy[cidx] = -1;
        const int r0 = d.lad_rung[i];
        double reg = ladder_reg(d.sd[(size_t)SD_REG * d.npad + i], r0, opt.update_regularization);
        for (int sidx = 0; sidx < S; ++sidx) {
            const int t = cidx * S + sidx;
            const bool ok = !(reg > 1000);      // rungs past the abort threshold are never attempted (MultiPhaseDDP.cpp:221-226)
            d.lad_t_inst[t] = ok ? i : -1;
            d.lad_t_rung[t] = ok ? r0 + sidx : -1;
            d.lad_t_reg[t] = reg;
            reg = fmax(reg * opt.update_regularization, 1e-03);
        }
        ++cidx;
    }
}

__global__ void k_lad_resolve(Dev d, Opt opt) {
    const int cidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (cidx >= d.lad_cnt[LAD_N_UNRES]) return;
    const int S = d.lad_cnt[LAD_S], stride = d.lad_cnt[LAD_STRIDE];
    const int i = d.lad_u_inst[cidx];
    const size_t es = d.npad;
    int best = -1, best_t = -1, tested = 0;
    for (int sidx = 0; sidx < S; ++sidx) {
        const int t = cidx * S + sidx;
        if (d.lad_t_inst[t] < 0) continue;
        ++tested;
        const int r = d.lad_t_rung[t];
        if (r >= 0 && (best < 0 || r < best)) { best = r; best_t = t; }
    }
    if (best >= 0) {   // the lowest passing rung: every rung below it was run and failed
        const double reg = d.lad_t_reg[best_t];
        d.sd[(size_t)SD_DV * es + i] = d.lad_t_dv[best_t];
        d.sd[(size_t)SD_REG_USED * es + i] = reg;
        double r2 = reg / 20;
        if (r2 < 1e-06) r2 = 0;
        d.sd[(size_t)SD_REG * es + i] = r2;
        d.si[(size_t)SI_N_BWS * es + i] = best + 1;
        d.si[(size_t)SI_BWS_OK * es + i] = 1;
        d.lad_unres[i] = 0;
        d.lad_u_copy[cidx] = (best_t % S != 0) ? lad_scratch_col(best_t, S) : -1;
        return;
    }
    const double reg0 = d.sd[(size_t)SD_REG * es + i];
    int L = d.lad_rung[i] + tested;                    // rungs 0 .. L-1 have failed
    if (stride > 0) {                                  // the chains ran to the end of the ladder: L = the first rung past 1000
        L = d.lad_rung[i];
        double r = ladder_reg(reg0, L, opt.update_regularization);
        while (!(r > 1000) && L < 64) { r = fmax(r * opt.update_regularization, 1e-03); ++L; }
    }
    if (ladder_reg(reg0, L, opt.update_regularization) > 1000 || stride > 0) {   // "Regularization term exceeds maximum value!"
        d.si[(size_t)SI_STATUS * es + i] = STATUS_REG_ABORT;
        d.si[(size_t)SI_DDP_ACTIVE * es + i] = 0;
        d.si[(size_t)SI_N_BWS * es + i] = L + 1;
        d.lad_unres[i] = 0;
    } else {
        d.lad_rung[i] = L;
    }
}

// K, du, G of a winner that ran as a speculative task: scratch column -> the instance's own records, exactly the entries a
// passing sweep writes (step knots: K, du, G; terminal knots: G).  One block row per unresolved instance; consecutive threads
// copy consecutive elements of the instance's records.
__global__ void k_lad_copy(Dev d) {
  for (int cidx = blockIdx.y; cidx < d.lad_cnt[LAD_N_UNRES]; cidx += gridDim.y) {
    const int t = d.lad_u_copy[cidx];
    if (t < 0) continue;
    const int i = d.lad_u_inst[cidx];
    const double* __restrict__ src = d.lad_gain + gain_idx(d, 0, (size_t)t);
    double* __restrict__ dst = d.gain + gain_idx(d, 0, (size_t)i);
    const size_t ks = gain_knot_stride(d);
    const int nwb = d.n_wb[i], np = d.n_phases[i];
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < np * KSTRIDE * GREC; idx += gridDim.x * blockDim.x) {
        const int kk = idx / GREC, e = idx - kk * GREC;
        const int p = kk / KSTRIDE, k = kk - p * KSTRIDE;
        const int N = d.ph_N[ph_idx(d, p, i)];
        if (k >= N) continue;
        const int nx = (p < nwb) ? 14 : 6;
        const bool term = (k == N - 1);
        const bool mine = e >= GR_G ? (e - GR_G < nx) : (!term && (e >= GR_DU || e < 4 * nx));
        if (mine) dst[(size_t)kk * ks + e] = src[(size_t)kk * ks + e];
    }
  }
}

void launch_bws(const Dev& d, const Consts& c, const Opt& opt, int mask_mode, int single_attempt, int max_attempts, cudaStream_t s) {
    cudaFuncSetAttribute(k_bws<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWS_SMEM);
    k_bws<0><<<(d.n + BWS_INST - 1) / BWS_INST, BWS_THREADS, BWS_SMEM, s>>>(d, c, opt, mask_mode, single_attempt, d.lad_unres ? max_attempts : 0);
}
// one round of the ladder scheduler: plan -> tasks -> resolve -> copy (4 launches, no host synchronisation)
void launch_lad_plan(const Dev& d, const Opt& opt, int mask_mode, int last_round, int smax, cudaStream_t s) {
    k_lad_plan<<<1, 1024, 0, s>>>(d, opt, mask_mode, last_round, smax);
}
void launch_lad_finish(const Dev& d, const Opt& opt, int n_unres, cudaStream_t s) {
    k_lad_resolve<<<(d.n + 127) / 128, 128, 0, s>>>(d, opt);
    // one block row per unresolved instance of this round (n_unres <= 0: not known on the host - rows loop over the device count)
    dim3 g(8, (unsigned)(n_unres > 0 ? std::min(n_unres, 32768) : std::min(d.npad, 2048)));
    k_lad_copy<<<g, 256, 0, s>>>(d);
}
void launch_bws_tasks(const Dev& d, const Consts& c, const Opt& opt, int mask_mode, int n_tasks, cudaStream_t s) {
    cudaFuncSetAttribute(k_bws<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWS_SMEM);
    const int tcap = n_tasks > 0 ? n_tasks : d.lad_tcap;
    k_bws<1><<<(tcap + BWS_INST - 1) / BWS_INST, BWS_THREADS, BWS_SMEM, s>>>(d, c, opt, mask_mode, 0, 0);
}
// one task in dump mode (capi.cu: replay_cost_to_go): a single block
void launch_bws_replay(const Dev& d, const Consts& c, const Opt& opt, cudaStream_t s) {
    cudaFuncSetAttribute(k_bws<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWS_SMEM);
    k_bws<1, 1><<<1, BWS_THREADS, BWS_SMEM, s>>>(d, c, opt, MASK_ALL, 0, 0);
}
// lane groups of k_bws that are resident at once on this device
int bws_resident_lane_groups(int device) {
    int nsm = 0, per_sm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
    cudaFuncSetAttribute(k_bws<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWS_SMEM);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bws<1>, BWS_THREADS, BWS_SMEM);
    return nsm * (per_sm > 0 ? per_sm : 1) * BWS_INST;
}

}  // namespace mhpc
