// Impact-aware Riccati backward sweep, SIXTEEN lanes per instance (two instances per warp, one warp per CTA).
//
// Same algorithm, same structure exploitation and - entry by entry - the same operation sequence as k_bws.cu (see the
// header there for the reference lines); what changes is the mapping.  The 8-lane kernel keeps two rows of the value
// Hessian per lane in 255 registers, which leaves ~1.7 warps per SM sub-partition at 4096 instances: every dependent
// DFMA chain, shared-memory hand-over and reciprocal is exposed.  Here lane m < 14 owns ONE row of H (lane m < 6 in the
// rigid-body phases), the kernel needs half the registers, twice as many warps are resident (14 per SM) and a warp
// instruction still serves two instances.  All 14 rows of M = H A go through shared memory (the row 7 + m of A^T M needs
// row m of M, which now sits in another lane).
#include <cuda_pipeline.h>

#include <algorithm>

#include "bws_shared.cuh"

namespace mhpc {

constexpr int B16_LPI = 16;                  // lanes per instance
constexpr int B16_THREADS = 32;              // one warp per CTA
constexpr int B16_INST = B16_THREADS / B16_LPI;
constexpr size_t B16_SMEM = (size_t)BWS_REGION * B16_INST * sizeof(double);
#ifndef B16_MINBLOCKS
#define B16_MINBLOCKS 14
#endif

template <int task_mode>
__global__ void __launch_bounds__(B16_THREADS, B16_MINBLOCKS) k_bws16(Dev d, Consts c, Opt opt, int mask_mode, int single_attempt, int max_attempts) {
    extern __shared__ __align__(16) double smem[];
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x;
    const int q2 = tid >> 4, m = tid & 15;          // instance slot in the warp, sub-lane
    double* bufs = smem + (size_t)q2 * BWS_REGION;                     // [2 stages][REC] of my instance
    double* scr = bufs + 2 * REC;
    const size_t es = d.npad;
    int i = blockIdx.x * B16_INST + q2, rung = 0, stride = 0;          // rung: 0-based attempt index of this sweep (n_bws = rung + 1)
    bool valid, enabled, to_scratch = false;
    double reg;
    if (!task_mode) {
        valid = i < d.n;
        if (!valid) i = 0;
        enabled = valid && inst_enabled(d, i, mask_mode);
        reg = enabled ? d.sd[(size_t)SD_REG * es + i] : 0.0;
    } else {
        const int nt = d.lad_cnt[LAD_N_TASKS];
        if (blockIdx.x * B16_INST >= nt) return;                       // the whole warp is beyond the task list
        const int ti = i < nt ? d.lad_t_inst[i] : -1;
        valid = enabled = ti >= 0;
        reg = valid ? d.lad_t_reg[i] : 0.0;
        rung = valid ? d.lad_t_rung[i] : 0;
        stride = d.lad_cnt[LAD_STRIDE];
        to_scratch = (i % d.lad_cnt[LAD_S]) != 0;
    }
    const int task = blockIdx.x * B16_INST + q2;
    if (task_mode) i = valid ? d.lad_t_inst[task] : 0;
    const int np = valid ? d.n_phases[i] : 0, nwb = valid ? d.n_wb[i] : 0, nk = enabled ? d.n_knots[i] : 0;
    const double dtw = valid ? d.dt_wb[i] : 0.0, dtf = valid ? d.dt_fb[i] : 0.0;
    bool done = !enabled;
    const int ktb = max(nk, __shfl_xor_sync(FULL, nk, 16));            // knots this warp walks
    const double* lin_i = d.lin + (size_t)i * d.KT * REC;
    const bool rowl = m < 14;                                          // this sub-lane owns WB row m
    const size_t ocol = to_scratch ? (size_t)lad_scratch_col(task, d.lad_cnt[LAD_S]) : (size_t)i;
    double* const grec = (to_scratch ? d.lad_gain : d.gain) + gain_idx(d, 0, ocol);     // [KT][GREC] of this instance / scratch column
    double* const gv_lane = grec + GR_G + m;
    double* const kf_lane = grec + GR_K + 4 * m;
    double* const du_lane = grec + GR_DU + m;
    const size_t gv_knot = gain_knot_stride(d), kf_knot = gv_knot, du_knot = gv_knot;

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
            for (int q = m; 2 * q < len; q += B16_LPI) __pipeline_memcpy_async(dst + 2 * q, src + 2 * q, 16);
            walk_step(w);
        }
        __pipeline_commit();
    };

    while (true) {
        // ---- one attempt -----------------------------------------------------------------
        double Hr[14];
#pragma unroll
        for (int q = 0; q < 14; ++q) Hr[q] = 0.0;
        scr[S_G + m] = 0.0;
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
                double go = 0.0;
                if (embed) {
                    if (m < 6) {
#pragma unroll
                        for (int q = 0; q < 6; ++q) scr[S_MQ + m * MQ_LD + q] = Hr[q];
                    }
                    if (rowl) {
                        const int pa = wb_to_rb(m);
                        go = (pa >= 0) ? scr[S_G + pa] : 0.0;
                    }
                }
                __syncwarp();
                if (embed && rowl) embed_row(m, scr, Hr);
                __syncwarp();
                if (embed && rowl) scr[S_G + m] = go;
                __syncwarp();
                // (b) WB phase end: G = Px^T G', H = Px^T H' Px  (MultiPhaseDDP.cpp:320-326), Px = [I7 0; 0 Pb]
                const bool px = wbp && !last;
                if (px && rowl) px_right_row(m, buf, scr, Hr);
                __syncwarp();
                double gn = 0.0;
                if (px && rowl) gn = px_left_row(m, buf, scr, Hr);
                __syncwarp();
                if (px && rowl) scr[S_G + m] = gn;
                __syncwarp();
                // (c) terminal cost: G = Phix + Gnext, H = Phixx + Hnext  (SinglePhase.cpp:190-191)
                if (on) {
                    if (last) {
#pragma unroll
                        for (int q = 0; q < 14; ++q) Hr[q] = 0.0;
                    }
                    if (wbp) {
                        if (rowl) {
                            const double ga = wb_terminal_row(m, buf, last ? 0.0 : scr[S_G + m], Hr);
                            scr[S_G + m] = ga;
                            gv_lane[(size_t)kk * gv_knot] = ga;
                        }
                    } else if (m < 6) {
                        double gv = last ? 0.0 : scr[S_G + m];
#pragma unroll
                        for (int jc = 0; jc < 6; ++jc) Hr[jc] = ((jc == m) ? c.rbQf[m] : 0.0) + Hr[jc];
                        gv = RBUF(R_RPHIX + m) + gv;
                        scr[S_G + m] = gv;
                        gv_lane[(size_t)kk * gv_knot] = gv;
                    }
                }
                __syncwarp();
            }

            // ============ whole-body step ============
            if (b1) {
                const bool on = (ty == 1);
                const bool rw = on && rowl;
                const double dt = dtw;
                const int l = m;                    // my row
                const bool top = l < 7;             // rows 0..6: positions, rows 7..13: velocities
                const int lt = top ? l : l - 7;     // the row of M that enters my row of A^T M through the [I; dt I] block
                double Mr[14];
                if (rw) {
                    // M = H A, row l.  Columns go in blocks of four (four accumulation chains), each chain in ascending s.
                    Mr[0] = Hr[0]; Mr[1] = Hr[1];
#pragma unroll
                    for (int j0 = 2; j0 < 14; j0 += 4) {
                        double xa[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int jc = j0 + k;
                            xa[k] = (jc < 7) ? Hr[jc] : Hr[jc - 7] * dt;
                        }
#pragma unroll
                        for (int s = 0; s < 7; ++s) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) xa[k] += Hr[7 + s] * RBUF(R_AB(s, j0 + k - 2));
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) Mr[j0 + k] = xa[k];
                    }
#pragma unroll
                    for (int jc = 0; jc < 14; ++jc) scr[S_MQ + l * MQ_LD + jc] = Mr[jc];
                    if (!top) {   // N = H B (rows 7..13 are all B^T N needs)
                        double nn[4];
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) nn[cc] = Hr[7] * RBUF(R_BB(0, cc));
#pragma unroll
                        for (int s = 1; s < 7; ++s)
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) nn[cc] += Hr[7 + s] * RBUF(R_BB(s, cc));
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) scr[S_N + (l - 7) * 4 + cc] = nn[cc];
                    }
                }
                __syncwarp();
                double qux[4] = {0, 0, 0, 0}, qx = 0.0;
                if (on) {
                    const double L00 = RBUF(R_LYY), L01 = RBUF(R_LYY + 1), L11 = RBUF(R_LYY + 2);
                    const double ly0 = RBUF(R_LY), ly1 = RBUF(R_LY + 1);
                    if (rowl) {
                        const bool has_ab = l >= 2;     // column l of A has a dense part (rows 7..13) from column 2 on
                        double ab[7];
#pragma unroll
                        for (int s = 0; s < 7; ++s) ab[s] = has_ab ? RBUF(R_AB(s, l - 2)) : 0.0;
                        const double c0 = has_ab ? RBUF(R_CB(0, l - 2)) : 0.0, c1 = has_ab ? RBUF(R_CB(1, l - 2)) : 0.0;
                        const double w0 = c0 * L00 + c1 * L01, w1 = c0 * L01 + c1 * L11;   // row of C^T lyy
                        const double lxxd = (l >= 3 && l < 7) ? RBUF(R_LXXQ + (l - 3)) : (2 * dt) * c.wbQ[l];
                        // Qxx row = (lxx + C^T lyy C) + A^T M, four columns at a time
                        double Qr[14];
#pragma unroll
                        for (int j0 = 0; j0 < 14; j0 += 4) {
                            double qa[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (j0 + k < 14) {
                                    const double mt = top ? Mr[j0 + k] : scr[S_MQ + lt * MQ_LD + j0 + k];
                                    qa[k] = top ? mt : mt * dt;
                                }
                            }
#pragma unroll
                            for (int s = 0; s < 7; ++s) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    if (j0 + k < 14) {
                                        const double mr = scr[S_MQ + (7 + s) * MQ_LD + j0 + k];
                                        if (has_ab) qa[k] += ab[s] * mr;
                                    }
                                }
                            }
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int jc = j0 + k;
                                if (jc < 14) {
                                    double lc = (jc == l) ? lxxd : 0.0;
                                    if (jc >= 2) {
                                        const double cj0 = RBUF(R_CB(0, jc - 2)), cj1 = RBUF(R_CB(1, jc - 2));
                                        if (has_ab) lc = lc + (w0 * cj0 + w1 * cj1);
                                    }
                                    Qr[jc] = lc + qa[k];
                                    if (jc == l) Qr[jc] += 1.0 * reg;
                                }
                            }
                        }
#pragma unroll
                        for (int jc = 0; jc < 14; ++jc) Hr[jc] = Qr[jc];
                        // Qx = (lx + A^T G) + C^T ly
                        double ag = top ? scr[S_G + l] : scr[S_G + lt] * dt;
                        if (has_ab) {
#pragma unroll
                            for (int s = 0; s < 7; ++s) ag += ab[s] * scr[S_G + 7 + s];
                        }
                        qx = (RBUF(R_LX + l) + ag) + (c0 * ly0 + c1 * ly1);
                        // Qux column l = (0 + D^T lyy C) + B^T M
                        {
                            double xa[4];
                            const double m0 = scr[S_MQ + 7 * MQ_LD + l];
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) xa[cc] = RBUF(R_BB(0, cc)) * m0;
#pragma unroll
                            for (int s = 1; s < 7; ++s) {
                                const double ms = scr[S_MQ + (7 + s) * MQ_LD + l];
#pragma unroll
                                for (int cc = 0; cc < 4; ++cc) xa[cc] += RBUF(R_BB(s, cc)) * ms;
                            }
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) {
                                const double d0 = RBUF(R_DB(0, cc)), d1 = RBUF(R_DB(1, cc));
                                const double v0 = d0 * L00 + d1 * L01, v1 = d0 * L01 + d1 * L11;
                                qux[cc] = (0.0 + (v0 * c0 + v1 * c1)) + xa[cc];
                            }
                        }
                    }
                    {   // Quu entry (row m&3, col m>>2): (luu + D^T lyy D) + B^T N
                        const int rr = m & 3, cc = m >> 2;
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
                    for (int jc = 0; jc < 14; ++jc) scr[S_MQ + l * MQ_LD + jc] = Hr[jc];
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) scr[S_QUX + cc * 14 + l] = qux[cc];
                }
                __syncwarp();
                if (rw) {
#pragma unroll
                    for (int jc = 0; jc < 14; ++jc) Hr[jc] = (Hr[jc] + scr[S_MQ + jc * MQ_LD + l]) / 2;   // Qxx = (Qxx + Qxx^T)/2
                }
                double kc[4], du[4], gout = 0.0, gdummy = 0.0, dvs = 0.0;
                __syncwarp();
                const bool ok = value_update<14, false, B16_LPI>(on, rw, m, scr, qux, qux, qx, qx, Hr, Hr, opt.quu_shift, kc, kc, du, gout, gdummy, dvs);
                if (on) {
                    if (!ok) { failed = true; fail_it = it; }
                    else {
                        dV += dvs;
                        if (rowl) {
                            scr[S_G + l] = gout;
                            gv_lane[(size_t)kk * gv_knot] = gout;
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) kf_lane[(size_t)kk * kf_knot + cc] = kc[cc];
                        }
                        if (m < 4) du_lane[(size_t)kk * du_knot] = du[m];
                    }
                }
            }

            // ============ rigid-body step ============  (see k_bws.cu; sub-lane m < 6 owns row m)
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
                    scr[S_MQ + l * MQ_LD + 0] = Hr[0] + Hr[5] * a0;
                    scr[S_MQ + l * MQ_LD + 1] = Hr[1] + Hr[5] * a1;
                    scr[S_MQ + l * MQ_LD + 2] = Hr[2];
                    scr[S_MQ + l * MQ_LD + 3] = Hr[0] * dt + Hr[3];
                    scr[S_MQ + l * MQ_LD + 4] = Hr[1] * dt + Hr[4];
                    scr[S_MQ + l * MQ_LD + 5] = Hr[2] * dt + Hr[5];
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) scr[S_N + l * 4 + cc] = Hr[3 + (cc & 1)] * bt[cc] + Hr[5] * b5[cc];
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
                            Hr[jc] = ((jc == l) ? (2 * dt) * c.rbQ[l] : 0.0) + q;
                            if (jc == l) Hr[jc] += 1.0 * reg;
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
                    {
                        const int rr = m & 3, cc = m >> 2;
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
                    for (int jc = 0; jc < 6; ++jc) scr[S_MQ + l * MQ_LD + jc] = Hr[jc];
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) scr[S_QUX + cc * 14 + l] = qux[cc];
                }
                __syncwarp();
                if (rw) {
#pragma unroll
                    for (int jc = 0; jc < 6; ++jc) Hr[jc] = (Hr[jc] + scr[S_MQ + jc * MQ_LD + l]) / 2;
                }
                double kc[4], du[4], gout = 0.0, gdummy = 0.0, dvs = 0.0;
                __syncwarp();
                const bool ok = value_update<6, false, B16_LPI>(on, rw, m, scr, qux, qux, qx, qx, Hr, Hr, opt.quu_shift, kc, kc, du, gout, gdummy, dvs);
                if (on) {
                    if (!ok) { failed = true; fail_it = it; }
                    else {
                        dV += dvs;
                        if (l < 6) {
                            scr[S_G + l] = gout;
                            gv_lane[(size_t)kk * gv_knot] = gout;
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) kf_lane[(size_t)kk * kf_knot + cc] = kc[cc];
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

        // ---- outcome of the attempt (MultiPhaseDDP.cpp:209-241); identical to k_bws.cu ----
        int retry = 0;
        if (!done && d.bws_trace && m == 0 && rung < BWS_TRACE_ATT) {
            const int r = d.si[(size_t)SI_N_REC * es + i];
            if (r < d.max_rec) d.bws_trace[((size_t)r * BWS_TRACE_ATT + rung) * es + i] = failed ? fail_it + 1 : -1;
        }
        if (!done) {
            if (task_mode) {
                if (!failed) {
                    if (m == 0) { d.lad_t_rung[task] = rung; d.lad_t_reg[task] = reg; d.lad_t_dv[task] = dV; }
                    done = true;
                } else {
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
                } else if (max_attempts > 0 && rung >= max_attempts) {
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

void launch_bws16(const Dev& d, const Consts& c, const Opt& opt, int mask_mode, int single_attempt, int max_attempts, cudaStream_t s) {
    cudaFuncSetAttribute(k_bws16<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B16_SMEM);
    k_bws16<0><<<(d.n + B16_INST - 1) / B16_INST, B16_THREADS, B16_SMEM, s>>>(d, c, opt, mask_mode, single_attempt, d.lad_unres ? max_attempts : 0);
}
void launch_bws16_tasks(const Dev& d, const Consts& c, const Opt& opt, int mask_mode, int n_tasks, cudaStream_t s) {
    cudaFuncSetAttribute(k_bws16<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B16_SMEM);
    const int tcap = n_tasks > 0 ? n_tasks : d.lad_tcap;
    k_bws16<1><<<(tcap + B16_INST - 1) / B16_INST, B16_THREADS, B16_SMEM, s>>>(d, c, opt, mask_mode, 0, 0);
}
int bws16_resident_lane_groups(int device) {
    int nsm = 0, per_sm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
    cudaFuncSetAttribute(k_bws16<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B16_SMEM);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bws16<1>, B16_THREADS, B16_SMEM);
    return nsm * (per_sm > 0 ? per_sm : 1) * B16_INST;
}

}  // namespace mhpc
