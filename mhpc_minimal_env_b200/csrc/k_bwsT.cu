// Impact-aware Riccati backward sweep on the FP64 tensor cores: ONE warp per instance, every matrix product of a knot
// as m8n8k4 DMMA instructions on 8x8 register tiles (MultiPhaseDDP.cpp:100-127, 300-341; SinglePhase.cpp:183-216;
// MHPC_CompoundTypes.h:116-144; regularisation ladder MultiPhaseDDP.cpp:209-227).
//
// Why: on B200 a DMMA has the same FP64 peak as DFMA (scripts/micro/dmma.cu: 63.8 vs 58 FMA/clk/SM), so this is not about
// flops.  The lane-per-row kernels (k_bws.cu, k_bws16.cu) are bound by the shared-memory data pipe: every DFMA of the
// 14x14 products needs an operand another lane produced (ncu: LSU wavefronts 56 % of peak, FP64 pipe 27 %).  A DMMA does
// 256 FMAs from two operand registers per lane - the tensor core does the cross-lane broadcast - and with the layout trick
// below NO operand ever changes layout between products, so a knot needs ~40 shared-memory loads instead of ~600.
//
// Homogeneous form.  The value function V(x) = 1/2 x'Hx + G'x is carried as the symmetric matrix [H G; G' *] and the
// dynamics as [A 0; 0 1], so that Qx, Qu, G and du fall out of the same products as Qxx, Qux, H and K.  The whole-body state
// splits into positions (0..6) and velocities (7..13): tile index 0..6 + one spare, 8 = the DMMA tile edge:
//     top tile index 0..6 = x[0..6], 7 = the homogeneous 1;   bottom tile index 0..6 = x[7..13], 7 = unused (zero)
//     T11 = [H11 G1; G1' *]   T12 = [H12 0; G2' 0]   T22 = [H22 0; 0 0]        (T21 = T12' is never needed)
//     A~ = [I8 dtI~; E~ F~],  E~ = A[7:14, 0:7] (columns 0,1 empty), F~ = A[7:14, 7:14],  dtI~ = diag(dt x 7, 0)
// The rigid-body model (6 states) lives in one tile: index 0..5 = x, 6 unused, 7 = homogeneous.
//
// Layout trick.  A DMMA accumulator tile X sits in "C layout": lane (g, t) = (lane / 4, lane % 4) holds X[g][2t], X[g][2t+1].
// Feeding c0 / c1 of two C-layout tiles X, Y as the A / B operands of two DMMAs computes D += X * Y' with the
// contraction index permuted identically on both sides (hardware k = t of step s <-> column 2t + s) - i.e. ANY accumulator
// is directly usable as a left factor, or (transposed) as a right factor.  With H symmetric the recursion can be arranged
// so that this is all that is ever needed:
//     M' = A~' T        (X = E~', F~' from the knot record, Y = T12, T22)         8 DMMA
//     Q  = M' A~        (X = M', Y = E~', F~' - the same registers)               6 DMMA   (Q21 = Q12' is not formed)
//     U  = M' B~        (Qxu, columns spread to 0,2,4,6; row "hom" = Qu)          4 DMMA
//     Quu = (B~' T22) B~                                                          4 DMMA
//     K' = U (-Quu^-1),  T+ = Q + U K                                             2 + 3 DMMA (single k-step: the spread
//                                                                                 columns are hardware k)
// plus six single-step DMMAs for the output-cost terms C'lyyC, C'lyyD, D'lyyD, C'ly, D'ly on stance knots.
// Every per-lane record element is read through a lane-constant offset that points at a zero slot of the record buffer
// when the lane has no such element: no predicates, no selects in the knot loop.
// The 4x4 part: Quu is symmetrised with one shuffle and written to shared memory as a doubly periodic 7x7 table, so that
// each even-row lane forms the one cofactor-inverse entry it feeds to the K product from ONE lane-constant base address
// (the determinant is a 4-lane shuffle sum); the positive-definite test is decided by the pivots of an unpivoted LDL' when
// they are far from zero and by the restated Eigen LDLT (bws_shared.cuh) otherwise.
// The step knots of a phase run in one of three specialised loops (stance / flight / rigid body) so that only the lane
// constants of that kind are live; the knot records arrive by one cp.async.bulk per knot (mbarrier completion), issued
// one knot ahead by a single lane.
// Phase-boundary knots (8 per sweep) convert the tiles to rows and run the row code shared with k_bws.cu.
// Products are accumulated by the tensor core in its own order: results differ from the lane-per-row kernels in the last
// bits (1e-13 relative over a sweep); every integer trace of the test batches is unchanged.
#include <algorithm>
#include <type_traits>

#include "bws_shared.cuh"

namespace mhpc {

constexpr int BT_THREADS = 32;
constexpr int ZSLOT = REC - 1;        // element 175 of a record stage: 0.0 (k_lin writes it as the pad of the last quad of a WB step
                                      // record; no other record kind reaches it; both stages are cleared at kernel start)
static_assert(WB_STEP_LEN == REC - 1 && WB_TERM_LEN < REC - 1 && RB_STEP_LEN < REC - 1, "the zero slot must be outside every record but the WB step pad");
#ifndef BT_AHEAD
#define BT_AHEAD 2        // knots in flight ahead of the one being computed (1..3); measured 2.31 / 2.25 / 2.27 ms per sweep
#endif
#ifndef BT_SYM_MASK
#define BT_SYM_MASK 0     // Qxx is symmetrised on the knots k with (k & mask) == 0 (every knot, like the reference; mask 3 is not faster)
#endif
#ifndef BT_CT_SELECT
#define BT_CT_SELECT 1      // right factor of the output-cost products by selects instead of three more loads
#endif
constexpr int BT_NST = BT_AHEAD == 1 ? 2 : 4;       // record stages (a power of two > BT_AHEAD)
constexpr int BT_NSH = BT_AHEAD == 1 ? 1 : 2;       // log2(BT_NST)
constexpr int BT_SCR = BT_NST * REC;  // scratch behind the record stages
constexpr int BT_BAR = BT_NST * REC + SCR + 14;     // one 8-byte mbarrier per stage
constexpr int BT_NPH = BT_BAR + BT_NST;             // knot counts per phase (ints)
constexpr size_t BT_SMEM = (size_t)(BT_NPH + MAXP / 2 + 2) * sizeof(double);
#ifndef BT_MINBLOCKS
#define BT_MINBLOCKS 13   // a hint to ptxas only: 13 ... 16 all compile to 128 registers (16 warps resident per SM) but schedule differently -
                          // 1.99 / 2.00 / 2.06 / 2.06 ms per sweep; 12 takes 164 registers (2.69 ms), 18 and 20 take 96 (21 warps: 2.09 ms)
#endif
constexpr int S_QP = S_QUX;           // Quu as a doubly periodic 7 x 7 table QP[a + 7 b] = Quu(a & 3, b & 3) (the Qux area is unused here)
constexpr int S_SYM = S_MQ;           // two 8 x 10 transposition areas (the row workspace is idle during step knots)
constexpr int SYM_LD = 10;            // even: [g][2t] is 16-byte aligned; the transposed reads of a half-warp hit 16 distinct banks
static_assert(S_SYM % 2 == 0 && 2 * 8 * SYM_LD <= 14 * MQ_LD, "transposition areas");

struct Tile { double c0, c1; };

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// D += X * Y^T, X and Y in C layout (contraction over the eight columns)
__device__ __forceinline__ void mm_xyT(Tile& D, const Tile& X, const Tile& Y) {
    dmma(D.c0, D.c1, X.c0, Y.c0);
    dmma(D.c0, D.c1, X.c1, Y.c1);
}
__device__ __forceinline__ Tile tzero() { Tile z; z.c0 = 0.0; z.c1 = 0.0; return z; }
// ---- one-dimensional bulk copy global -> shared, completion on an mbarrier (one lane issues) ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(void* bar) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar))); }
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, void* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bar_wait(void* bar, unsigned parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// positive-definite test of S = Quu - shift I (SinglePhase.cpp:202-205) by an LDL^T elimination without pivoting (lower triangle,
// as Eigen reads it).  By Sylvester's law of inertia the signs of the pivots are those Eigen's pivoted LDL^T finds, whatever the
// order - unless a pivot is lost in rounding.  Every pivot is therefore compared with its own diagonal entry: 1 = all pivots are
// positive and none is smaller than 1e-9 of its diagonal entry (its rounding error is ~1e-16 of it), 0 = the first pivot that is
// not clearly positive is clearly negative, -1 = too close to call (the caller then runs the restated Eigen LDLT, whose pivot
// order decides ties).
__device__ __forceinline__ int pd_quick(const double* q, double shift) {
#define QL(r, c) q[(r) + 7 * (c)]
    const double tol = 1e-9;
    const double s00 = QL(0, 0) - shift, s11 = QL(1, 1) - shift, s22 = QL(2, 2) - shift, s33 = QL(3, 3) - shift;
    const double s10 = QL(1, 0), s20 = QL(2, 0), s30 = QL(3, 0), s21 = QL(2, 1), s31 = QL(3, 1), s32 = QL(3, 2);
#undef QL
    const double d0 = s00;
    const double r0 = sx_rcp(d0);
    const double l10 = s10 * r0, l20 = s20 * r0, l30 = s30 * r0;
    const double d1 = s11 - l10 * s10;
    const double r1 = sx_rcp(d1);
    const double t21 = s21 - l20 * s10, t31 = s31 - l30 * s10;
    const double l21 = t21 * r1, l31 = t31 * r1;
    const double d2 = (s22 - l20 * s20) - l21 * t21;
    const double r2 = sx_rcp(d2);
    const double t32 = (s32 - l30 * s20) - l31 * t21;
    const double l32 = t32 * r2;
    const double d3 = ((s33 - l30 * s30) - l31 * t31) - l32 * t32;
    const double a0 = tol * fabs(s00), a1 = tol * fabs(s11), a2 = tol * fabs(s22), a3 = tol * fabs(s33);
    if (d0 > a0 && d1 > a1 && d2 > a2 && d3 > a3) return 1;
    if (!(d0 > a0)) return (d0 < -a0) ? 0 : -1;
    if (!(d1 > a1)) return (d1 < -a1) ? 0 : -1;
    if (!(d2 > a2)) return (d2 < -a2) ? 0 : -1;
    return (d3 < -a3) ? 0 : -1;
}

template <int task_mode>
__global__ void __launch_bounds__(BT_THREADS, BT_MINBLOCKS) k_bwsT(Dev d, Consts c, Opt opt, int mask_mode, int single_attempt, int max_attempts) {
    extern __shared__ __align__(16) double smem[];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x;
    const int g = lane >> 2, t = lane & 3;          // C-layout coordinates: row g, columns 2t and 2t + 1
    const int n0 = 2 * t, n1 = 2 * t + 1;
    const int m = lane;                             // row owner index of the phase-boundary code (lanes 0..13)
    double* bufs = smem;                            // [2 stages][REC]
    double* scr = smem + BT_SCR;
    int* sN = reinterpret_cast<int*>(smem + BT_NPH);
    const size_t es = d.npad;
    int i = blockIdx.x, rung = 0, stride = 0;       // rung: 0-based attempt index of this sweep (n_bws = rung + 1)
    bool valid, enabled, to_scratch = false;
    double reg;
    if (!task_mode) {
        valid = i < d.n;
        if (!valid) i = 0;
        enabled = valid && inst_enabled(d, i, mask_mode);
        reg = enabled ? d.sd[(size_t)SD_REG * es + i] : 0.0;
    } else {
        const int nt = d.lad_cnt[LAD_N_TASKS];
        if (i >= nt) return;
        const int ti = d.lad_t_inst[i];
        valid = enabled = ti >= 0;
        reg = valid ? d.lad_t_reg[i] : 0.0;
        rung = valid ? d.lad_t_rung[i] : 0;
        stride = d.lad_cnt[LAD_STRIDE];
        to_scratch = (i % d.lad_cnt[LAD_S]) != 0;
    }
    const int task = blockIdx.x;
    if (task_mode) i = valid ? d.lad_t_inst[task] : 0;
    if (!enabled) return;                           // one instance per warp: nothing to wait for
    const int np = d.n_phases[i], nwb = d.n_wb[i];
    const double dtw = d.dt_wb[i], dtf = d.dt_fb[i];
    const double* lin_i = d.lin + (size_t)i * d.KT * REC;
    const size_t ocol = to_scratch ? (size_t)lad_scratch_col(task, d.lad_cnt[LAD_S]) : (size_t)i;
    double* const grec = (to_scratch ? d.lad_gain : d.gain) + gain_idx(d, 0, ocol);   // gain records [KT][GREC] of this instance / scratch column
    const unsigned bar_a = smem_u32(smem + BT_BAR), buf_a = smem_u32(bufs);   // shared-window addresses of barrier 0 / stage 0
    if (lane < np) sN[lane] = d.ph_N[ph_idx(d, lane, i)];
    if (lane < BT_NST) {
        bufs[lane * REC + ZSLOT] = 0.0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a + 8u * lane));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();

    // ---- record stream: one bulk copy per knot into stage (cnt % BT_NST), issued BT_AHEAD knots ahead by lane 0 from a cursor
    // over the knot sequence (phases last to first; in a phase the boundary knot, then the step knots downwards: one
    // descending run of records).  The barrier of a stage completes once per use, so its parity is the bit of the running
    // knot counter above the stage index. ----
    unsigned cnt = 0;                               // knots consumed since the kernel started (all attempts)
    unsigned fcnt = 0;                              // knots fetched
    const double* c_src = lin_i;                    // cursor of the next fetch: record, knots left in its phase, phase
    int c_left = 0, c_p = -1;
    unsigned c_bytes = 0;
    auto cursor_phase = [&](int p) {                // put the cursor on the boundary knot of phase p (p < 0: end of the stream)
        c_p = p;
        if (p >= 0) {
            c_left = sN[p];
            c_src = lin_i + (size_t)(p * KSTRIDE + c_left - 1) * REC;
            c_bytes = (p < nwb) ? WB_TERM_LEN * 8u : RB_TERM_LEN * 8u;
        } else c_left = 0;
    };
    auto fetch_next = [&]() {                       // issue the copy of the cursor knot, if any, and advance
        if (c_left > 0) {
            if (lane == 0) {
                const unsigned st = fcnt & (BT_NST - 1), bar = bar_a + 8u * st, dst = buf_a + (unsigned)(REC * 8) * st;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(c_bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(c_src), "r"(c_bytes), "r"(bar) : "memory");
            }
            ++fcnt;
            c_src -= REC;
            c_bytes = (c_p < nwb) ? (unsigned)REC * 8u : RB_STEP_LEN * 8u;
            if (--c_left == 0) cursor_phase(c_p - 1);
        }
    };
    auto landed = [&](unsigned knot_cnt) {
        const unsigned bar = bar_a + 8u * (knot_cnt & (BT_NST - 1)), parity = (knot_cnt >> BT_NSH) & 1u;
        asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(bar), "r"(parity) : "memory");
    };

    // ---- lane constants shared by all kinds of step knots ----
    const bool gf = g < 7, gev = !(g & 1), hom = g == 7, k1ok = n1 < 7;
    const bool dgu = gev && t == (g >> 1);                                // my quu.c0 is a diagonal entry of Quu
    const int sOwn = S_SYM + g * SYM_LD + n0, sT0 = S_SYM + n0 * SYM_LD + g, sT1 = S_SYM + n1 * SYM_LD + g;
    // 4x4 part: even-row lane (g, t) holds Quu(r, cq) = Quu(g/2, t) and produces entry (t, g/2) of the inverse
    const int qr = g >> 1, qc = t;
    const int src_T = 8 * t + (g >> 1);                                   // the lane that holds Quu(cq, r)
    const int oQP = S_QP + qr + 7 * qc;                                   // my entry in the periodic table (+4, +28, +32: its replicas)
    const bool rep_r = qr < 3, rep_c = qc < 3;
    const int oCof = S_QP + qr + 7 * qc;                                  // cofactor of (i, jj) = (g/2, t): elements QP[oCof + a + 7 b], a, b = 1..3
    const double sgn = ((qr + qc) & 1) ? -1.0 : 1.0;
    // my elements of a knot's gain record (K | du | G, contiguous: a store instruction of the warp fills whole sectors); the knot
    // offset added per store is the same for all lanes
    double* const pK1 = grec + GR_K + 4 * (gf ? g : 0) + t;
    double* const pG = grec + GR_G + (hom ? 7 + n0 : (gf ? g : 0));
    double* const pdU = grec + GR_DU + t;

    while (true) {
        // ---- one attempt -----------------------------------------------------------------
        Tile T11 = tzero(), T12 = tzero(), T22 = tzero();   // whole-body value (homogeneous form); T11 doubles as the rigid-body tile
        double dVp = 0.0;                                   // my share of dV (lanes of row "hom")
        bool failed = false;
        int fail_it = 0, it = 0;
        __syncwarp();
        fcnt = cnt;
        cursor_phase(np - 1);
        for (int q = 0; q < BT_AHEAD; ++q) fetch_next();

        // the step knots k = N-2 .. 0 of phase p; KIND 0 = whole-body stance, 1 = whole-body flight, 2 = rigid body
        auto run_steps = [&](auto KIND, int p, int N, int mode) {
            constexpr int kind = decltype(KIND)::value;
            constexpr bool WB = kind != 2;
            const double dt = WB ? dtw : dtf;
            // ---- lane constants of this kind: record offsets (ZSLOT = "this lane has no such element"), cost constants ----
            const bool ge = (g >= 2 && g < 7), g6 = g < 6;
            const bool dg0 = (WB ? gf : g6) && n0 == g, dg1 = (WB ? gf : g6) && n1 == g;   // my c0 / c1 sits on a state diagonal
            int oE0 = ZSLOT, oE1 = ZSLOT, oF0 = ZSLOT, oF1 = ZSLOT, oB0, oB1, oQa, oQb, oQ12a = ZSLOT, oQ12b = ZSLOT, oU, oQuu = ZSLOT;
            double kQa, kQb, kQ22a = 0.0, kQ22b = 0.0, kQuu, aR0 = 0.0, aR1 = 0.0, sdt = 0.0, cdt1 = 0.0;
            if (WB) {
                oE0 = ge ? R_AB(n0, g - 2) : ZSLOT; oE1 = (ge && k1ok) ? R_AB(n1, g - 2) : ZSLOT;      // C layout of E~'
                oF0 = gf ? R_AB(n0, 5 + g) : ZSLOT; oF1 = (gf && k1ok) ? R_AB(n1, 5 + g) : ZSLOT;      //             F~'
                oB0 = gev ? R_BB(n0, g >> 1) : ZSLOT; oB1 = (gev && k1ok) ? R_BB(n1, g >> 1) : ZSLOT;  //             spread B~'
                oQa = dg0 ? ((g >= 3) ? R_LXXQ + g - 3 : ZSLOT) : (hom ? R_LX + n0 : ZSLOT);
                oQb = dg1 ? ((g >= 3) ? R_LXXQ + g - 3 : ZSLOT) : ((gf && t == 3) ? R_LX + g : ((hom && k1ok) ? R_LX + n1 : ZSLOT));
                oQ12a = hom ? R_LX + 7 + n0 : ZSLOT; oQ12b = (hom && k1ok) ? R_LX + 7 + n1 : ZSLOT;
                oU = hom ? R_LU + t : ZSLOT;
                oQuu = dgu ? R_LUU + t : ZSLOT;
                const double cQ11 = ((dg0 || dg1) && g < 3) ? (2 * dt) * c.wbQ[g & 3] : 0.0;            // lxx where it is a constant
                const double cQ22 = (dg0 || dg1) ? (2 * dt) * c.wbQ[7 + (gf ? g : 0)] : 0.0;
                kQa = dg0 ? cQ11 + 1.0 * reg : 0.0; kQb = dg1 ? cQ11 + 1.0 * reg : 0.0;
                kQ22a = dg0 ? cQ22 + 1.0 * reg : 0.0; kQ22b = dg1 ? cQ22 + 1.0 * reg : 0.0;
                kQuu = dgu ? 1.0 * reg : 0.0;
                sdt = gf ? dt : 0.0;                 // row scale of dtI~ for my row
                cdt1 = (t < 3) ? dt : 0.0;           // column scale of my c1 (column 7 when t == 3); c0 columns are always < 7
            } else {
                // A~' : lane holds A~[k' = 2t + s][n = g];  A = I + dt [0 I3; 0 0] + e5 (a0, a1, 0, ...); spread B~': B(3 + (c&1), c), B(5, c)
                auto aconst = [&](int kq, int nn) -> double { return (kq == nn && nn != 6) ? 1.0 : ((kq < 3 && nn == kq + 3) ? dt : 0.0); };
                aR0 = aconst(n0, g); aR1 = aconst(n1, g);
                oE1 = (t == 2 && g < 2) ? R_RA + g : ZSLOT;               // A(5, 0), A(5, 1): k' = 5 = n1 of t = 2
                const int cc = g >> 1;
                oB0 = (gev && n0 == 3 + (cc & 1)) ? R_RB + 2 * cc : ZSLOT;
                oB1 = (gev && n1 == 3 + (cc & 1)) ? R_RB + 2 * cc : ((gev && n1 == 5) ? R_RB + 2 * cc + 1 : ZSLOT);
                oQa = (hom && n0 < 6) ? R_RLX + n0 : ZSLOT;
                oQb = (g6 && t == 3) ? R_RLX + g : ((hom && n1 < 6) ? R_RLX + n1 : ZSLOT);
                oU = hom ? R_RLU + t : ZSLOT;
                const double cRQ = (dg0 || dg1) ? (2 * dt) * c.rbQ[g6 ? g : 0] : 0.0;
                kQa = dg0 ? cRQ + 1.0 * reg : 0.0; kQb = dg1 ? cRQ + 1.0 * reg : 0.0;
                kQuu = dgu ? (2 * dt) * c.rbR[mode - 1][t] + 1.0 * reg : 0.0;
            }
            // output-cost terms (stance knots): Z^' L^ Z^ with Z^ = [C D 0; 0 0 1] (rows y0, y1, hom), L^ = [lyy ly; ly' 0]
            int oZt0 = ZSLOT, oZt1 = ZSLOT, oZb0 = ZSLOT, oZb1 = ZSLOT, oZu0 = ZSLOT, oZu1 = ZSLOT, oL0 = ZSLOT, oL1 = ZSLOT, oL2 = ZSLOT, oYt = ZSLOT, oYb = ZSLOT, oYu = ZSLOT;
            double zt2 = 0.0, cYt = 0.0;
            if (kind == 0) {
                oZt0 = ge ? R_CB(0, g - 2) : ZSLOT; oZt1 = ge ? R_CB(1, g - 2) : ZSLOT;
                oZb0 = gf ? R_CB(0, 5 + g) : ZSLOT; oZb1 = gf ? R_CB(1, 5 + g) : ZSLOT;
                oZu0 = gev ? R_DB(0, g >> 1) : ZSLOT; oZu1 = gev ? R_DB(1, g >> 1) : ZSLOT;
                oL0 = (t == 0) ? R_LYY : (t == 1) ? R_LYY + 1 : (t == 2) ? R_LY : ZSLOT;       // column t of L^
                oL1 = (t == 0) ? R_LYY + 1 : (t == 1) ? R_LYY + 2 : (t == 2) ? R_LY + 1 : ZSLOT;
                oL2 = (t == 0) ? R_LY : (t == 1) ? R_LY + 1 : ZSLOT;
                oYt = (t == 0) ? oZt0 : (t == 1) ? oZt1 : ZSLOT; oYb = (t == 0) ? oZb0 : (t == 1) ? oZb1 : ZSLOT; oYu = (t == 0) ? oZu0 : (t == 1) ? oZu1 : ZSLOT;
                zt2 = hom ? 1.0 : 0.0; cYt = (hom && t == 2) ? 1.0 : 0.0;
            }
            const size_t gks = gain_knot_stride(d);
            size_t koff = (size_t)(p * KSTRIDE + N - 2) * gks;       // element offset of the current knot's gain record
            for (int k = N - 2; k >= 0; --k, ++it, ++cnt, koff -= gks) {
                const double* buf = bufs + (cnt & (BT_NST - 1)) * REC;
                landed(cnt);
                __syncwarp();                        // everyone is done with the stage of the previous knot (and with its scratch)
                fetch_next();                        // ... which the copy of knot cnt + BT_AHEAD may now overwrite
                Tile u1 = tzero(), u2 = tzero(), quu = tzero(), q11, q12 = tzero(), q22 = tzero();
                if (WB) {
                    Tile eT, fT, bT;
                    eT.c0 = buf[oE0]; eT.c1 = buf[oE1]; fT.c0 = buf[oF0]; fT.c1 = buf[oF1]; bT.c0 = buf[oB0]; bT.c1 = buf[oB1];
                    // M' = A~' T
                    Tile m11 = T11, m12 = T12, m21, m22;
                    m21.c0 = T11.c0 * sdt; m21.c1 = T11.c1 * sdt; m22.c0 = T12.c0 * sdt; m22.c1 = T12.c1 * sdt;
                    mm_xyT(m11, eT, T12);
                    mm_xyT(m12, eT, T22);
                    mm_xyT(m21, fT, T12);
                    mm_xyT(m22, fT, T22);
                    // Q = M' A~ ; U = M' B~ ; Quu = (B~' T22) B~
                    Tile w = tzero();
                    q11 = m11;
                    q12.c0 = m11.c0 * dt; q12.c1 = m11.c1 * cdt1; q22.c0 = m21.c0 * dt; q22.c1 = m21.c1 * cdt1;
                    mm_xyT(w, bT, T22);
                    mm_xyT(q11, m12, eT);
                    mm_xyT(q12, m12, fT);
                    mm_xyT(q22, m22, fT);
                    mm_xyT(u1, m12, bT);
                    mm_xyT(u2, m22, bT);
                    mm_xyT(quu, w, bT);
                    if (kind == 0) {                 // one k-step each
                        const double lc0 = buf[oL0], lc1 = buf[oL1], lc2 = buf[oL2];
                        const double zt0 = buf[oZt0], zt1 = buf[oZt1], zb0 = buf[oZb0], zb1 = buf[oZb1], zu0 = buf[oZu0], zu1 = buf[oZu1];
                        const double xt = (zt0 * lc0 + zt1 * lc1) + zt2 * lc2;               // (Z^' L^)[g][t]: top / bottom / input tiles
                        const double xb = zb0 * lc0 + zb1 * lc1;
                        const double xu = zu0 * lc0 + zu1 * lc1;
#if BT_CT_SELECT
                        const double yt = (t == 0) ? zt0 : (t == 1) ? zt1 : cYt;             // Z^[t][g]
                        const double yb = (t == 0) ? zb0 : (t == 1) ? zb1 : 0.0, yu = (t == 0) ? zu0 : (t == 1) ? zu1 : 0.0;
#else
                        const double yt = buf[oYt] + cYt, yb = buf[oYb], yu = buf[oYu];      // Z^[t][g]
#endif
                        dmma(q11.c0, q11.c1, xt, yt);
                        dmma(q12.c0, q12.c1, xt, yb);
                        dmma(q22.c0, q22.c1, xb, yb);
                        dmma(u1.c0, u1.c1, xt, yu);
                        dmma(u2.c0, u2.c1, xb, yu);
                        dmma(quu.c0, quu.c1, xu, yu);
                    }
                    // running-cost terms: lxx (+ reg) on the diagonal, lx in the homogeneous row / column, lu, luu (+ reg)
                    q12.c0 += buf[oQ12a]; q12.c1 += buf[oQ12b];
                    q22.c0 += kQ22a; q22.c1 += kQ22b;
                    quu.c0 += buf[oQuu];
                } else {
                    // rigid body: one tile (index 0..5 state, 6 unused, 7 homogeneous)
                    Tile aT, bT;
                    aT.c0 = aR0; aT.c1 = aR1 + buf[oE1]; bT.c0 = buf[oB0]; bT.c1 = buf[oB1];
                    Tile mt = tzero(), w = tzero();
                    q11 = tzero();
                    mm_xyT(mt, aT, T11);        // M' = A~' T
                    mm_xyT(w, bT, T11);
                    mm_xyT(q11, mt, aT);        // Q = M' A~
                    mm_xyT(u1, mt, bT);         // U = M' B~
                    mm_xyT(quu, w, bT);
                }
                q11.c0 += buf[oQa] + kQa; q11.c1 += buf[oQb] + kQb;
                u1.c0 += buf[oU];
                quu.c0 += kQuu;
                if (lane == 31) q11.c1 = 0.0;                            // [hom][hom]: the constant term is not carried
                // Qxx = (Qxx + Qxx')/2 on the diagonal tiles (the off-diagonal block is one tile used for both halves);
                // Quu = (Quu + Quu')/2 -> doubly periodic table
                const bool do_sym = BT_SYM_MASK == 0 || ((k & BT_SYM_MASK) == 0);   // rounding-level asymmetry only
                if (do_sym) {
                    *reinterpret_cast<double2*>(scr + sOwn) = make_double2(q11.c0, q11.c1);
                    if (WB) *reinterpret_cast<double2*>(scr + sOwn + 80) = make_double2(q22.c0, q22.c1);
                }
                const double qv = (quu.c0 + __shfl_sync(FULL, quu.c0, src_T)) / 2;   // Quu(g/2, t), symmetrised
                if (gev) {
                    scr[oQP] = qv;
                    if (rep_r) scr[oQP + 4] = qv;
                    if (rep_c) scr[oQP + 28] = qv;
                    if (rep_r && rep_c) scr[oQP + 32] = qv;
                }
                __syncwarp();
                if (do_sym) {
                    q11.c0 = (q11.c0 + scr[sT0]) / 2; q11.c1 = (q11.c1 + scr[sT1]) / 2;
                    if (WB) { q22.c0 = (q22.c0 + scr[sT0 + 80]) / 2; q22.c1 = (q22.c1 + scr[sT1 + 80]) / 2; }
                }
                // ---- 4x4 part: positive-definite test, my entry of the cofactor inverse ----
                int pd = pd_quick(scr + S_QP, opt.quu_shift);
                if (pd < 0) {
                    double tst[4][4];
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int cq = 0; cq < 4; ++cq) tst[r][cq] = scr[S_QP + r + 7 * cq] - (r == cq ? 1.0 * opt.quu_shift : 0.0);
                    pd = ldlt4_is_positive(tst) ? 1 : 0;
                }
                if (!pd) { failed = true; fail_it = it; return; }
                double nq;
                {
#define EQ(a, b) scr[oCof + (a) + 7 * (b)]
                    const double cf = EQ(1, 1) * (EQ(2, 2) * EQ(3, 3) - EQ(2, 3) * EQ(3, 2)) + EQ(2, 1) * (EQ(3, 2) * EQ(1, 3) - EQ(3, 3) * EQ(1, 2)) +
                                      EQ(3, 1) * (EQ(1, 2) * EQ(2, 3) - EQ(1, 3) * EQ(2, 2));
#undef EQ
                    const double adj = sgn * cf;                         // adj(t, g/2) = adj(g/2, t): Quu is exactly symmetric here
                    double det = qv * adj;                               // Quu(g/2, t) adj(t, g/2), summed over t within the quad
                    det += __shfl_xor_sync(FULL, det, 1);
                    det += __shfl_xor_sync(FULL, det, 2);
                    nq = gev ? -(adj * sx_rcp(det)) : 0.0;               // -Quu^-1 entry (t, g/2) on the even rows
                }
                // K' = U (-Quu^-1); T+ = Q + U K
                Tile kt1 = tzero(), kt2 = tzero();
                dmma(kt1.c0, kt1.c1, u1.c0, nq);
                dmma(q11.c0, q11.c1, u1.c0, kt1.c0);
                if (WB) {
                    dmma(kt2.c0, kt2.c1, u2.c0, nq);
                    dmma(q12.c0, q12.c1, u1.c0, kt2.c0);
                    dmma(q22.c0, q22.c1, u2.c0, kt2.c0);
                }
                dVp += u1.c0 * kt1.c0;                                   // on row "hom": Qu[t] du[t]
                if (lane == 31) q11.c1 = 0.0;
                T11 = q11; T12 = q12; T22 = q22;
                // K (4 x nx column-major), du, G of this knot
                if (WB) {
                    if (gf) {
                        pK1[koff] = kt1.c0;
                        pK1[koff + 28] = kt2.c0;
                        if (t == 3) pG[koff] = T11.c1;
                    } else {
                        pdU[koff] = kt1.c0;
                        pG[koff] = T12.c0;
                        if (k1ok) pG[koff + 1] = T12.c1;
                    }
                } else {
                    if (g6) {
                        pK1[koff] = kt1.c0;
                        if (t == 3) pG[koff] = T11.c1;
                    } else if (hom) {
                        pdU[koff] = kt1.c0;
                    }
                }
            }
        };

        for (int p = np - 1; p >= 0 && !failed; --p) {
            const int N = sN[p];
            const bool wb = p < nwb;
            const int mode = d.ph_mode[ph_idx(d, p, i)];
            // ============ phase boundary (last knot of the phase): tiles -> rows, the row code of k_bws.cu, rows -> tiles ============
            {
                const double* buf = bufs + (cnt & (BT_NST - 1)) * REC;
                landed(cnt);
                __syncwarp();
                fetch_next();
                const int kk = p * KSTRIDE + N - 1;
                const bool last = (p == np - 1);
                const bool next_rb = !last && (p + 1 >= nwb);
                double Hr[14];
#pragma unroll
                for (int q = 0; q < 14; ++q) Hr[q] = 0.0;
                if (m < 16) scr[S_G + m] = 0.0;
                __syncwarp();
                if (!last) {
                    if (next_rb) {   // rigid-body tile: index 0..5 state, 7 homogeneous
                        if (g < 6) {
                            if (n0 < 6) scr[S_MQ + g * MQ_LD + n0] = T11.c0;
                            if (n1 < 6) scr[S_MQ + g * MQ_LD + n1] = T11.c1;
                            if (t == 3) scr[S_G + g] = T11.c1;
                        }
                    } else {
                        if (g < 7) {
                            scr[S_MQ + g * MQ_LD + n0] = T11.c0;
                            if (n1 < 7) scr[S_MQ + g * MQ_LD + n1] = T11.c1; else scr[S_G + g] = T11.c1;
                            scr[S_MQ + g * MQ_LD + 7 + n0] = T12.c0;
                            scr[S_MQ + (7 + n0) * MQ_LD + g] = T12.c0;
                            if (n1 < 7) { scr[S_MQ + g * MQ_LD + 7 + n1] = T12.c1; scr[S_MQ + (7 + n1) * MQ_LD + g] = T12.c1; }
                            scr[S_MQ + (7 + g) * MQ_LD + 7 + n0] = T22.c0;
                            if (n1 < 7) scr[S_MQ + (7 + g) * MQ_LD + 7 + n1] = T22.c1;
                        } else {
                            scr[S_G + 7 + n0] = T12.c0;
                            if (n1 < 7) scr[S_G + 7 + n1] = T12.c1;
                        }
                    }
                    __syncwarp();
                    const int nxn = next_rb ? 6 : 14;
                    if (m < nxn) {
#pragma unroll
                        for (int q = 0; q < 14; ++q) if (q < nxn) Hr[q] = scr[S_MQ + m * MQ_LD + q];
                    }
                    __syncwarp();
                }
                const bool rowl = m < 14;
                // (a) WB <- RB: embed the 6-dim value into 14 dims through P^T (.) P  (MultiPhaseDDP.cpp:330-331)
                const bool embed = wb && next_rb;
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
                const bool px = wb && !last;
                if (px && rowl) px_right_row(m, buf, scr, Hr);
                __syncwarp();
                double gn = 0.0;
                if (px && rowl) gn = px_left_row(m, buf, scr, Hr);
                __syncwarp();
                if (px && rowl) scr[S_G + m] = gn;
                __syncwarp();
                // (c) terminal cost: G = Phix + Gnext, H = Phixx + Hnext  (SinglePhase.cpp:190-191)
                if (last) {
#pragma unroll
                    for (int q = 0; q < 14; ++q) Hr[q] = 0.0;
                }
                if (wb) {
                    if (rowl) {
                        const double ga = wb_terminal_row(m, buf, last ? 0.0 : scr[S_G + m], Hr);
                        scr[S_G + m] = ga;
                        grec[(size_t)kk * gain_knot_stride(d) + GR_G + m] = ga;
                    }
                } else if (m < 6) {
                    double gv = last ? 0.0 : scr[S_G + m];
#pragma unroll
                    for (int jc = 0; jc < 6; ++jc) Hr[jc] = ((jc == m) ? c.rbQf[m] : 0.0) + Hr[jc];
                    gv = buf[R_RPHIX + m] + gv;
                    scr[S_G + m] = gv;
                    grec[(size_t)kk * gain_knot_stride(d) + GR_G + m] = gv;
                }
                __syncwarp();
                // rows -> tiles of this phase's model
                const int nxc = wb ? 14 : 6;
                if (m < nxc) {
#pragma unroll
                    for (int q = 0; q < 14; ++q) if (q < nxc) scr[S_MQ + m * MQ_LD + q] = Hr[q];
                }
                __syncwarp();
                if (wb) {
                    if (g < 7) {
                        T11.c0 = scr[S_MQ + g * MQ_LD + n0];
                        T11.c1 = (n1 < 7) ? scr[S_MQ + g * MQ_LD + n1] : scr[S_G + g];
                        T12.c0 = scr[S_MQ + g * MQ_LD + 7 + n0];
                        T12.c1 = (n1 < 7) ? scr[S_MQ + g * MQ_LD + 7 + n1] : 0.0;
                        T22.c0 = scr[S_MQ + (7 + g) * MQ_LD + 7 + n0];
                        T22.c1 = (n1 < 7) ? scr[S_MQ + (7 + g) * MQ_LD + 7 + n1] : 0.0;
                    } else {
                        T11.c0 = scr[S_G + n0];
                        T11.c1 = (n1 < 7) ? scr[S_G + n1] : 0.0;
                        T12.c0 = scr[S_G + 7 + n0];
                        T12.c1 = (n1 < 7) ? scr[S_G + 7 + n1] : 0.0;
                        T22 = tzero();
                    }
                } else {
                    T12 = tzero(); T22 = tzero();
                    if (g < 6) {
                        T11.c0 = (n0 < 6) ? scr[S_MQ + g * MQ_LD + n0] : 0.0;
                        T11.c1 = (n1 < 6) ? scr[S_MQ + g * MQ_LD + n1] : (n1 == 7 ? scr[S_G + g] : 0.0);
                    } else if (g == 7) {
                        T11.c0 = (n0 < 6) ? scr[S_G + n0] : 0.0;
                        T11.c1 = (n1 < 6) ? scr[S_G + n1] : 0.0;
                    } else {
                        T11 = tzero();
                    }
                }
                __syncwarp();
                ++it; ++cnt;
            }
            // ============ step knots of the phase, last to first ============
            if (!wb) run_steps(std::integral_constant<int, 2>{}, p, N, mode);
            else if (mode == 1 || mode == 3) run_steps(std::integral_constant<int, 0>{}, p, N, mode);
            else run_steps(std::integral_constant<int, 1>{}, p, N, mode);
        }
        // after a failure the copy issued for the knot behind the failing one is still in flight: wait for it, so that the
        // next attempt starts with idle stages and barriers
        __syncwarp();
        if (failed) {   // the failing knot used counter value cnt; the knots behind it that are in flight: wait for them, skip them
            for (unsigned q = cnt + 1; q < fcnt; ++q) landed(q);
            cnt = fcnt;
        }
        __syncwarp();
        double dV = dVp;                                                 // sum of the four lanes of row "hom"
        dV += __shfl_xor_sync(FULL, dV, 1);
        dV += __shfl_xor_sync(FULL, dV, 2);
        dV = __shfl_sync(FULL, dV, 28);

        // ---- outcome of the attempt (MultiPhaseDDP.cpp:209-241); identical to k_bws.cu ----
        int retry = 0;
        if (d.bws_trace && lane == 0 && rung < BWS_TRACE_ATT) {
            const int r = d.si[(size_t)SI_N_REC * es + i];
            if (r < d.max_rec) d.bws_trace[((size_t)r * BWS_TRACE_ATT + rung) * es + i] = failed ? fail_it + 1 : -1;
        }
        if (task_mode) {
            if (!failed) {
                if (lane == 0) { d.lad_t_rung[task] = rung; d.lad_t_reg[task] = reg; d.lad_t_dv[task] = dV; }
            } else {
                bool more = stride > 0;
                double nr = reg;
                for (int k = 0; k < stride && more; ++k) { nr = fmax(nr * opt.update_regularization, 1e-03); if (nr > 1000) more = false; }
                if (more) { reg = nr; rung += stride; retry = 1; }
                else if (lane == 0) d.lad_t_rung[task] = -1;
            }
        } else if (single_attempt) {
            if (lane == 0) {
                d.si[(size_t)SI_BWS_OK * es + i] = failed ? 0 : 1;
                if (!failed) d.sd[(size_t)SD_DV * es + i] = dV;
                if (d.lad_unres) d.lad_unres[i] = 0;
            }
        } else if (failed) {
            reg = fmax(reg * opt.update_regularization, 1e-03);
            ++rung;
            if (reg > 1000) {
                if (lane == 0) {
                    d.si[(size_t)SI_STATUS * es + i] = STATUS_REG_ABORT; d.si[(size_t)SI_DDP_ACTIVE * es + i] = 0; d.si[(size_t)SI_N_BWS * es + i] = rung + 1;
                    if (d.lad_unres) d.lad_unres[i] = 0;
                }
            } else if (max_attempts > 0 && rung >= max_attempts) {
                if (lane == 0) { d.lad_unres[i] = 1; d.lad_rung[i] = rung; d.si[(size_t)SI_BWS_OK * es + i] = 0; }
            } else retry = 1;
        } else {
            if (lane == 0) {
                d.sd[(size_t)SD_DV * es + i] = dV;
                d.sd[(size_t)SD_REG_USED * es + i] = reg;
                double r2 = reg / 20;
                if (r2 < 1e-06) r2 = 0;
                d.sd[(size_t)SD_REG * es + i] = r2;
                d.si[(size_t)SI_N_BWS * es + i] = rung + 1;
                d.si[(size_t)SI_BWS_OK * es + i] = 1;
                if (d.lad_unres) d.lad_unres[i] = 0;
            }
        }
        if (!retry) break;
    }
}

void launch_bwsT(const Dev& d, const Consts& c, const Opt& opt, int mask_mode, int single_attempt, int max_attempts, cudaStream_t s) {
    k_bwsT<0><<<d.n, BT_THREADS, BT_SMEM, s>>>(d, c, opt, mask_mode, single_attempt, d.lad_unres ? max_attempts : 0);
}
void launch_bwsT_tasks(const Dev& d, const Consts& c, const Opt& opt, int mask_mode, int n_tasks, cudaStream_t s) {
    const int tcap = n_tasks > 0 ? n_tasks : d.lad_tcap;
    k_bwsT<1><<<tcap, BT_THREADS, BT_SMEM, s>>>(d, c, opt, mask_mode, 0, 0);
}
int bwsT_resident_warps(int device) {
    int nsm = 0, per_sm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bwsT<1>, BT_THREADS, BT_SMEM);
    return nsm * (per_sm > 0 ? per_sm : 1);
}

}  // namespace mhpc
