// TEST INFRASTRUCTURE (oracle) - not part of the product; see oracle/README.md and the
// header of hsddp_oracle.h.  CPU restatement of the reference solver around the
// reference's own generated C.  Citations are to files under /root/reference.
#include "hsddp_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>

#include "hsddp_oracle_c.h"

// ---- the reference's generated C (CasadiGen/header/*.h:15-30), linked from oracle/_ref --
extern "C" {
typedef long long int casadi_int_t;
#define ORC_DECL(F)                                                                        \
    int F(const double** arg, double** res, casadi_int_t* iw, double* w, int mem);          \
    const casadi_int_t* F##_sparsity_out(casadi_int_t i);                                   \
    int F##_work(casadi_int_t* sz_arg, casadi_int_t* sz_res, casadi_int_t* sz_iw, casadi_int_t* sz_w);
ORC_DECL(Dyn_BS) ORC_DECL(Dyn_FS) ORC_DECL(Dyn_FL)
ORC_DECL(Dyn_BS_par) ORC_DECL(Dyn_FS_par) ORC_DECL(Dyn_FL_par)
ORC_DECL(Imp_F) ORC_DECL(Imp_B) ORC_DECL(Imp_F_par) ORC_DECL(Imp_B_par)
ORC_DECL(FBDynamics) ORC_DECL(FBDynamics_par)
ORC_DECL(WB_FL1_terminal_constr) ORC_DECL(WB_FL2_terminal_constr)
ORC_DECL(Jacob_F) ORC_DECL(Jacob_B)
ORC_DECL(Link0Jacobian) ORC_DECL(Link1Jacobian) ORC_DECL(Link2Jacobian) ORC_DECL(Link3Jacobian) ORC_DECL(Link4Jacobian)
#undef ORC_DECL
}

namespace orc {

typedef int (*casadi_fn)(const double**, double**, casadi_int_t*, double*, int);
typedef const casadi_int_t* (*casadi_sp)(casadi_int_t);
typedef int (*casadi_wk)(casadi_int_t*, casadi_int_t*, casadi_int_t*, casadi_int_t*);

static bool g_faithful_alloc = false;  // mimic the adaptor's per-call heap traffic (timing only)

// CasadiGen/source/CasadiGen.cpp:4-79: evaluate into per-output scratch, then scatter the
// CCS non-zeros column-major into the caller's dense storage; zeros are never written.
static void casadi_interface(std::initializer_list<const double*> ARG, std::initializer_list<double*> RES,
                             int max_sz_res, casadi_fn f, casadi_sp f_sparse_out, casadi_wk f_work) {
    const double* arg[8];
    double* res[8];
    double stackbuf[8][200];
    casadi_int_t sz_arg, sz_res, sz_iw, sz_w;
    f_work(&sz_arg, &sz_res, &sz_iw, &sz_w);
    int na = 0, nr = 0;
    for (const double* a : ARG) arg[na++] = a;
    std::vector<double*> heap;
    for (double* r : RES) {
        (void)r;
        if (g_faithful_alloc) { heap.push_back(new double[max_sz_res]); res[nr] = heap.back(); }
        else res[nr] = stackbuf[nr];
        ++nr;
    }
    casadi_int_t iw0[1];
    double w0[1];
    f(arg, res, iw0, w0, 1);
    int idx = 0;
    for (double* R : RES) {
        if (idx >= sz_res) break;
        const casadi_int_t* sp = f_sparse_out(idx);
        const int nrow = (int)sp[0], ncol = (int)sp[1];
        const casadi_int_t* colinfo = sp + 2;
        const casadi_int_t* rowinfo = colinfo + ncol + 1;
        int nz = 0;
        for (int c = 0; c < ncol; ++c)
            while (nz < colinfo[c + 1]) { R[rowinfo[nz] + nrow * c] = res[idx][nz]; ++nz; }
        ++idx;
    }
    for (double* h : heap) delete[] h;
}
#define CI(F) F, F##_sparsity_out, F##_work

// =====================================================================================
// Dynamics/source/PlanarQuadruped.cpp  (whole-body "FULL" model, 14/4/4)
// =====================================================================================
struct WBModel {
    double dt;
    Mat<14, 14> Ac;
    Mat<14, 4> Bc;
    Vec<14> xdot;
    // PlanarQuadruped.cpp:120-138
    double bodyLength = 0.19 * 2, hipLinkLength = 0.209, kneeLinkLength = 0.195;
    double hipLoc[2], kneeLoc[2];
    explicit WBModel(double dt_) : dt(dt_) {
        Ac.setZero(); Bc.setZero(); xdot.setZero();  // PlanarQuadruped.h:45-47
        hipLoc[0] = bodyLength / 2; hipLoc[1] = 0;
        kneeLoc[0] = 0; kneeLoc[1] = -hipLinkLength;
    }
    // PlanarQuadruped.cpp:8-27
    void dynamics(const Vec<14>& x, const Vec<4>& u, Vec<14>& x_next, Vec<4>& y, int mode) {
        switch (mode) {
            case 2: case 4: casadi_interface({x.d, u.d}, {xdot.d, y.d}, 14, CI(Dyn_FL)); break;
            case 1: casadi_interface({x.d, u.d}, {xdot.d, y.d}, 14, CI(Dyn_BS)); break;
            case 3: casadi_interface({x.d, u.d}, {xdot.d, y.d}, 14, CI(Dyn_FS)); break;
        }
        Vec<14> xn;
        for (int i = 0; i < 14; ++i) xn[i] = x[i] + xdot[i] * dt;
        x_next = xn;
    }
    // PlanarQuadruped.cpp:30-53
    void dynamics_par(const Vec<14>& x, const Vec<4>& u, Mat<14, 14>& A, Mat<14, 4>& B, Mat<4, 14>& C, Mat<4, 4>& D, int mode) {
        switch (mode) {
            case 2: case 4: casadi_interface({x.d, u.d}, {Ac.d, Bc.d, C.d, D.d}, 196, CI(Dyn_FL_par)); break;
            case 1: casadi_interface({x.d, u.d}, {Ac.d, Bc.d, C.d, D.d}, 196, CI(Dyn_BS_par)); break;
            case 3: casadi_interface({x.d, u.d}, {Ac.d, Bc.d, C.d, D.d}, 196, CI(Dyn_FS_par)); break;
        }
        for (int c = 0; c < 14; ++c)
            for (int r = 0; r < 14; ++r) A(r, c) = (r == c ? 1.0 : 0.0) + Ac(r, c) * dt;
        for (int i = 0; i < 56; ++i) B[i] = Bc[i] * dt;
    }
    // PlanarQuadruped.cpp:58-78 (modes 1,3 leave y untouched)
    void resetmap(const Vec<14>& x, Vec<14>& x_next, Vec<4>& y, int mode) {
        switch (mode) {
            case 1: case 3: x_next = x; break;
            case 2: casadi_interface({x.d}, {x_next.d, y.d}, 14, CI(Imp_F)); break;
            case 4: casadi_interface({x.d}, {x_next.d, y.d}, 14, CI(Imp_B)); break;
        }
    }
    // PlanarQuadruped.cpp:82-100
    void resetmap_par(const Vec<14>& x, Mat<14, 14>& Px, int mode) {
        switch (mode) {
            case 1: case 3: Px.setIdentity(); break;
            case 2: casadi_interface({x.d}, {Px.d}, 196, CI(Imp_F_par)); break;
            case 4: casadi_interface({x.d}, {Px.d}, 196, CI(Imp_B_par)); break;
        }
    }
    // PlanarQuadruped.cpp:103-117 (foot: 0 front, 1 hind)
    void FootJacobian(const Vec<14>& x, Mat<2, 7>& J, Mat<2, 7>& Jd, int foot) {
        if (foot == 0) casadi_interface({x.d}, {J.d, Jd.d}, 14, CI(Jacob_F));
        else if (foot == 1) casadi_interface({x.d}, {J.d, Jd.d}, 14, CI(Jacob_B));
    }
    // Common/header/orientation_tools.h:58-76 (Y axis) and :383-392
    static Mat<4, 4> homoY(double theta, double vx, double vy, double vz, bool identityR = false) {
        Mat<4, 4> H; H.setZero();
        if (identityR) { H(0, 0) = 1; H(1, 1) = 1; H(2, 2) = 1; }
        else {
            const double s = std::sin(theta), c = std::cos(theta);
            H(0, 0) = c; H(0, 1) = 0; H(0, 2) = -s;
            H(1, 0) = 0; H(1, 1) = 1; H(1, 2) = 0;
            H(2, 0) = s; H(2, 1) = 0; H(2, 2) = c;
        }
        H(0, 3) = vx; H(1, 3) = vy; H(2, 3) = vz; H(3, 3) = 1;
        return H;
    }
    // PlanarQuadruped.cpp:142-175; link ids PlanarQuadruped.h:15-24
    Mat<4, 4> get_homoTransformation(const Vec<7>& q, int linkidx) const {
        Mat<4, 4> T_body = homoY(-q[2], q[0], 0, q[1]);
        if (linkidx == 0) return T_body;
        Mat<4, 4> T_Fhip = T_body * homoY(-q[3], hipLoc[0], 0, hipLoc[1]);
        if (linkidx == 1) return T_Fhip;
        Mat<4, 4> T_Hhip = T_body * homoY(-q[5], -hipLoc[0], 0, hipLoc[1]);
        if (linkidx == 3) return T_Hhip;
        Mat<4, 4> T_Fknee = T_Fhip * homoY(-q[4], kneeLoc[0], 0, kneeLoc[1]);
        if (linkidx == 2) return T_Fknee;
        Mat<4, 4> T_Hknee = T_Hhip * homoY(-q[6], kneeLoc[0], 0, kneeLoc[1]);
        if (linkidx == 4) return T_Hknee;
        Mat<4, 4> T_Ffoot = T_Fknee * homoY(0, 0, 0, -kneeLinkLength, true);
        if (linkidx == 5) return T_Ffoot;
        Mat<4, 4> T_Hfoot = T_Hknee * homoY(0, 0, 0, -kneeLinkLength, true);
        return T_Hfoot;
    }
    // PlanarQuadruped.cpp:178-186
    void get_contact_position(const Vec<7>& q, int linkidx, const double loc[2], double out[2]) const {
        Mat<4, 4> H = get_homoTransformation(q, linkidx);
        Vec<4> v; v[0] = loc[0]; v[1] = 0; v[2] = loc[1]; v[3] = 1;
        Vec<4> p = H * v;
        out[0] = p[0]; out[1] = p[2];
    }
    // PlanarQuadruped.cpp:195-205 (legidx 0 front, 1 hind)
    void get_leg_ext_vec(const Vec<7>& q, int legidx, double out[2]) const {
        const double loc[2] = {0, 0};
        double a[2], b[2];
        if (legidx == 0) { get_contact_position(q, 5, loc, a); get_contact_position(q, 1, loc, b); }
        else { get_contact_position(q, 6, loc, a); get_contact_position(q, 3, loc, b); }
        out[0] = a[0] - b[0]; out[1] = a[1] - b[1];
    }
};

// Dynamics/header/FootholdPlan.h:6-51
struct FootholdPlanner {
    double velcmd, ground;
    Vec<7> config;
    WBModel* robot;
    FootholdPlanner(WBModel* r, double v, double g) : velcmd(v), ground(g), robot(r) { config.setZero(); }
    Vec<4> get_foothold_location(const double pos[3], double stance_time, int mode) {
        double pos_hip[2] = {0, 0};
        Vec<4> foothold; foothold.setZero();
        config[0] = pos[0]; config[1] = pos[1]; config[2] = pos[2];
        const double loc[2] = {0, 0};
        switch (mode) {
            case 1:
                robot->get_contact_position(config, 3, loc, pos_hip);
                foothold[2] = pos_hip[0] + velcmd * stance_time / 2;
                foothold[3] = ground;
                break;
            case 3:
                robot->get_contact_position(config, 1, loc, pos_hip);
                foothold[0] = pos_hip[0] + velcmd * stance_time / 2;
                foothold[1] = ground;
                break;
        }
        return foothold;
    }
};

// Dynamics/source/PlanarFloatingBase.cpp (rigid-body "RBD" model, 6/4/4)
struct FBModel {
    double dt;
    Vec<4> foothold;
    Vec<2> contact;
    Vec<6> xdot;
    Mat<6, 6> Ac;
    Mat<6, 4> Bc;
    FootholdPlanner* planner;
    FBModel(FootholdPlanner* p, double dt_) : dt(dt_), planner(p) {
        foothold.setZero(); contact.setZero(); xdot.setZero(); Ac.setZero(); Bc.setZero();
    }
    void set_contact(int mode) {  // PlanarFloatingBase.cpp:9-23
        switch (mode) {
            case 1: contact[0] = 0; contact[1] = 1; break;
            case 2: contact[0] = 0; contact[1] = 0; break;
            case 3: contact[0] = 1; contact[1] = 0; break;
            case 4: contact[0] = 0; contact[1] = 0; break;
        }
    }
    void dynamics(const Vec<6>& x, const Vec<4>& u, Vec<6>& x_next, Vec<4>& y, int mode) {  // :7-32
        set_contact(mode);
        xdot.setZero();
        casadi_interface({x.d, u.d, foothold.d, contact.d}, {xdot.d}, 6, CI(FBDynamics));
        y.setZero();
        Vec<6> xn;
        for (int i = 0; i < 6; ++i) xn[i] = x[i] + xdot[i] * dt;
        x_next = xn;
    }
    void dynamics_par(const Vec<6>& x, const Vec<4>& u, Mat<6, 6>& A, Mat<6, 4>& B, Mat<4, 6>& C, Mat<4, 4>& D, int mode) {  // :44-72
        set_contact(mode);
        Ac.setZero(); Bc.setZero();
        casadi_interface({x.d, u.d, foothold.d, contact.d}, {Ac.d, Bc.d}, 36, CI(FBDynamics_par));
        for (int c = 0; c < 6; ++c)
            for (int r = 0; r < 6; ++r) A(r, c) = (r == c ? 1.0 : 0.0) + Ac(r, c) * dt;
        for (int i = 0; i < 24; ++i) B[i] = Bc[i] * dt;
        C.setZero(); D.setZero();
    }
    void plan_foothold(const double* x, double stance_time, int mode) {  // :81-84
        foothold = planner->get_foothold_location(x, stance_time, mode);
    }
};

// =====================================================================================
// HSDDPSolver/source/CostBase.cpp + Controller/MHPCLocomotion/MHPCCost.cpp (diagonal weights)
// =====================================================================================
template <int n, int m, int p>
struct CostW {
    double dt;
    Mat<n, n> Q[4], Qf[4];
    Mat<m, m> R[4];
    Mat<p, p> S[4];
    void running_cost(const ModelState<n, m, p>& s, const ModelState<n, m, p>& r, RCostStruct<n, m, p>& rc, int mode) const {  // CostBase.cpp:4-16
        Vec<n> dx = s.x - r.x; Vec<m> du = s.u - r.u; Vec<p> dy = s.y - r.y;
        double l = (tmul(dx, Q[mode - 1]) * dx)[0];
        l += (tmul(du, R[mode - 1]) * du)[0];
        l += (tmul(dy, S[mode - 1]) * dy)[0];
        l *= dt;
        rc.l = l;
    }
    void running_cost_par(const ModelState<n, m, p>& s, const ModelState<n, m, p>& r, RCostStruct<n, m, p>& rc, int mode) const {  // :19-34
        rc.lx = ((2 * dt) * Q[mode - 1]) * (s.x - r.x);
        rc.lu = ((2 * dt) * R[mode - 1]) * (s.u - r.u);
        rc.ly = ((2 * dt) * S[mode - 1]) * (s.y - r.y);
        rc.lxx = (2 * dt) * Q[mode - 1];
        rc.luu = (2 * dt) * R[mode - 1];
        rc.lux.setZero();
        rc.lyy = (2 * dt) * S[mode - 1];
    }
    void terminal_cost(const ModelState<n, m, p>& s, const ModelState<n, m, p>& r, TCostStruct<n>& tc, int mode) const {  // :37-47
        Vec<n> dx = s.x - r.x;
        tc.Phi = (tmul(dx, Qf[mode - 1]) * dx)[0];
        tc.Phi *= 0.5;
    }
    void terminal_cost_par(const ModelState<n, m, p>& s, const ModelState<n, m, p>& r, TCostStruct<n>& tc, int mode) const {  // :50-60
        tc.Phix = Qf[mode - 1] * (s.x - r.x);
        tc.Phixx = Qf[mode - 1];
    }
};

template <int N>
static Mat<N, N> diag(const double* v, double scale) {
    Mat<N, N> M; M.setZero();
    for (int i = 0; i < N; ++i) M(i, i) = scale * v[i];
    return M;
}

static CostW<14, 4, 4> make_wb_cost(double dt) {  // MHPCCost.cpp:24-53
    CostW<14, 4, 4> c; c.dt = dt;
    const double q[14] = {0, 10, 5, 4, 4, 4, 4, 2, 1, .01, 6, 6, 6, 6};
    const double qf[4][14] = {{0, 20, 8, 3, 3, 3, 3, 3, 2, 0.01, 5, 5, 0.01, 0.01},
                              {0, 20, 8, 3, 3, 3, 3, 3, 2, 0.01, 5, 5, 5, 5},
                              {0, 20, 8, 3, 3, 3, 3, 3, 2, 0.01, 0.01, 0.01, 5, 5},
                              {0, 20, 8, 3, 3, 3, 3, 3, 2, 0.01, 5, 5, 5, 5}};
    const double r[4][4] = {{5, 5, 1, 1}, {1, 1, 1, 1}, {1, 1, 5, 5}, {1, 1, 1, 1}};
    // s[3] is uninitialised in the reference (MHPCCost.cpp:29,42-44): ORACLE CHOICE s4 = 0
    const double s[4][4] = {{0, 0, 0.3, 0.3}, {0, 0, 0, 0}, {0.15, 0.15, 0, 0}, {0, 0, 0, 0}};
    for (int i = 0; i < 4; ++i) {
        c.Q[i] = diag<14>(q, 0.01);
        c.R[i] = diag<4>(r[i], 0.5);
        c.S[i] = diag<4>(s[i], 1.0);
        c.Qf[i] = diag<14>(qf[i], 100);
    }
    return c;
}
static CostW<6, 4, 4> make_fb_cost(double dt) {  // MHPCCost.cpp:56-75
    CostW<6, 4, 4> c; c.dt = dt;
    const double q[6] = {0, 10, 5, 2, 1, 0.01}, qf[6] = {1, 20, 8, 3, 1, 0.01};
    const double r[4][4] = {{0, 0, 0.01, 0.01}, {0, 0, 0, 0}, {0.01, 0.01, 0, 0}, {0, 0, 0, 0}};
    for (int i = 0; i < 4; ++i) {
        c.Q[i] = diag<6>(q, 0.01);
        c.R[i] = diag<4>(r[i], 1.0);
        c.Qf[i] = diag<6>(qf, 100);
        c.S[i].setZero();
    }
    return c;
}

// =====================================================================================
// Controller/MHPCLocomotion/MHPCConstraints.cpp
// =====================================================================================
struct WBConstraint {
    int num_torque = 8, num_joint = 8, num_grf = 3;
    double C_torque[8][4], b_torque[8], C_joint[8][4], b_joint[8], C_grf[3][4], b_grf[3];
    float friccoeff = 0.5f;  // MHPCConstraints.h:33
    ALReBParam params[4];
    int num_pconstr[4] = {19, 16, 19, 16}, num_tconstr[4] = {0, 1, 0, 1};  // :20-30
    WBConstraint() {
        for (int md = 0; md < 4; ++md) {  // MHPC_CompoundTypes.h:228-234
            params[md].lambda.assign(num_tconstr[md], 0.0);
            params[md].delta.assign(num_pconstr[md], 0.1);
            params[md].delta_min.assign(num_pconstr[md], 0.01);
            params[md].eps_ReB.assign(num_pconstr[md], 0.0);
        }
        for (int md = 0; md < 4; ++md) {  // MHPCConstraints.cpp:59-88
            for (int i = 0; i < num_pconstr[md]; ++i) {
                params[md].delta[i] = 0.1; params[md].delta_min[i] = 0.01;
                params[md].eps_ReB[i] = (i < 8) ? 0.01 * 1.0 : (i < 16 ? 0.0 : 0.01 * 1.0);
            }
            params[md].reb_empty = false;
            if (md == 1 || md == 3) { params[md].sigma = 5; params[md].al_empty = false; }
        }
        std::memset(C_torque, 0, sizeof(C_torque)); std::memset(C_joint, 0, sizeof(C_joint));
        std::memset(C_grf, 0, sizeof(C_grf)); std::memset(b_grf, 0, sizeof(b_grf));
        for (int i = 0; i < 4; ++i) { C_torque[i][i] = -1; C_torque[4 + i][i] = 1; C_joint[i][i] = -1; C_joint[4 + i][i] = 1; }
        for (int i = 0; i < 8; ++i) b_torque[i] = 33;
        const double bj[8] = {PI / 4, -0.1, 1.15 * PI, -0.1, PI, PI - 0.2, .1, PI - 0.2};
        for (int i = 0; i < 8; ++i) b_joint[i] = bj[i];
    }
    static double rowdot4(const double* row, const double* v) {
        double s = row[0] * v[0];
        for (int i = 1; i < 4; ++i) s += row[i] * v[i];
        return s;
    }
    // :110-176
    void path_constraint(const ModelState<14, 4, 4>& ms, IneqConstrStruct<14, 4, 4>* pc, int mode) {
        if (!pc) return;
        for (int i = 0; i < 8; ++i) {
            pc[i].g = rowdot4(C_torque[i], ms.u.d) + b_torque[i];
            for (int j = 0; j < 4; ++j) pc[i].gu[j] = C_torque[i][j];
        }
        for (int i = 0; i < 8; ++i) {
            pc[8 + i].g = rowdot4(C_joint[i], ms.x.d + 3) + b_joint[i];
            for (int j = 0; j < 4; ++j) pc[8 + i].gx[3 + j] = C_joint[i][j];
        }
        if (mode == 1 || mode == 3) {
            const double mu = friccoeff;
            if (mode == 1) { const double c[3][4] = {{0, 0, 0, 1}, {0, 0, -1, mu}, {0, 0, 1, mu}}; std::memcpy(C_grf, c, sizeof(c)); }
            else { const double c[3][4] = {{0, 1, 0, 0}, {-1, mu, 0, 0}, {1, mu, 0, 0}}; std::memcpy(C_grf, c, sizeof(c)); }
            for (int i = 0; i < 3; ++i) {
                pc[16 + i].g = rowdot4(C_grf[i], ms.y.d) + b_grf[i];
                for (int j = 0; j < 4; ++j) pc[16 + i].gy[j] = C_grf[i][j];
            }
        }
    }
    // :91-107
    void terminal_constraint(const ModelState<14, 4, 4>& ms, TConstrStruct<14>* tc, int mode) {
        if (!tc) return;
        if (mode == 2) casadi_interface({ms.x.d}, {&tc[0].h, tc[0].hx.d, tc[0].hxx.d}, 196, CI(WB_FL1_terminal_constr));
        else if (mode == 4) casadi_interface({ms.x.d}, {&tc[0].h, tc[0].hx.d, tc[0].hxx.d}, 196, CI(WB_FL2_terminal_constr));
    }
};

// =====================================================================================
// HSDDPSolver/source/SinglePhase.cpp
// =====================================================================================
struct PhaseBase {
    int mode = 1, phaseidx = 0, N = 0, xdim = 0;
    bool is_wb = true;
    double dt = 0, V = 0, dV = 0, dVnext = 0;
    const Option* option = nullptr;
    ALReBParam param;
    virtual ~PhaseBase() {}
    virtual void set_x0(const double* x) = 0;
    virtual void get_terminal_state(double* x) const = 0;
    virtual void forward_sweep(double eps) = 0;
    virtual void forward_sweep_dynamics_only(double eps) = 0;
    virtual void forward_sweep_partials_only() = 0;
    virtual bool backward_sweep(double reg) = 0;
    virtual void update_nominal_trajectory() = 0;
    virtual void udpate_AL_ReB_Param() = 0;
    virtual double get_tconstr_violation() const = 0;
    virtual void get_CTG0(double& dV_, double* G, double* H) const = 0;
    virtual void set_CTG_next(double dVn, const double* G, const double* H) = 0;
    virtual int get_field(int field, double* out) const = 0;
    virtual double tconstr_h() const = 0;
    virtual void zero_data() = 0;
    virtual void swap_data(PhaseBase& other) = 0;
    virtual void initialization() = 0;
};

template <int n, int m, int p>
struct Phase : PhaseBase {
    std::vector<ModelState<n, m, p>> act, nom, ref;
    std::vector<DynDerivative<n, m, p>> dynpar;
    std::vector<RCostStruct<n, m, p>> rcost;
    std::vector<CostToGoStruct<n, m>> CTG;
    TCostStruct<n> tcost;
    std::vector<TConstrStruct<n>> tconstr;
    std::vector<IneqConstrStruct<n, m, p>> pconstr;
    std::vector<double> B, Bz, Bzz;
    int num_pconstr = 0, num_tconstr = 0;
    Vec<n> x0, Gnext;
    Mat<n, n> Hnext;
    WBModel* wb = nullptr;
    FBModel* fb = nullptr;
    const CostW<n, m, p>* cost = nullptr;
    WBConstraint* wbcon = nullptr;

    Phase() {
        this->xdim = n;
        auto z = [](auto& v) { std::memset((void*)v.data(), 0, v.size() * sizeof(v[0])); };
        act.resize(N_TIMESTEPS_MAX); nom.resize(N_TIMESTEPS_MAX); ref.resize(N_TIMESTEPS_MAX);
        dynpar.resize(N_TIMESTEPS_MAX); rcost.resize(N_TIMESTEPS_MAX); CTG.resize(N_TIMESTEPS_MAX);
        z(act); z(nom); z(ref); z(dynpar); z(rcost); z(CTG);  // MHPCLocomotion.cpp:265-290
        std::memset(&tcost, 0, sizeof(tcost));
        x0.setZero(); Gnext.setZero(); Hnext.setZero();
    }
    void zero_data() override {   // MHPCLocomotion.cpp:265-290 (memset of every per-phase array)
        auto z = [](auto& v) { std::memset((void*)v.data(), 0, v.size() * sizeof(v[0])); };
        z(act); z(nom); z(ref); z(dynpar); z(rcost); z(CTG);
        std::memset(&tcost, 0, sizeof(tcost));
    }
    void swap_data(PhaseBase& other) override {   // set_data with rotated buffers (MHPCLocomotion.cpp:134-151)
        auto& o = static_cast<Phase<n, m, p>&>(other);
        act.swap(o.act); nom.swap(o.nom); ref.swap(o.ref); dynpar.swap(o.dynpar); rcost.swap(o.rcost); CTG.swap(o.CTG);
        std::swap(tcost, o.tcost);
    }
    void initialization() override {  // SinglePhase.cpp:26-42
        if (is_wb) { num_pconstr = wbcon->num_pconstr[mode - 1]; num_tconstr = wbcon->num_tconstr[mode - 1]; param = wbcon->params[mode - 1]; }
        else { num_pconstr = 0; num_tconstr = 0; param = ALReBParam(); }  // MHPCConstraints.cpp:5-11
        pconstr.assign(num_pconstr, IneqConstrStruct<n, m, p>());
        for (auto& c : pconstr) std::memset(&c, 0, sizeof(c));
        tconstr.assign(num_tconstr, TConstrStruct<n>());
        for (auto& c : tconstr) std::memset(&c, 0, sizeof(c));
        B.assign(num_pconstr, 0); Bz.assign(num_pconstr, 0); Bzz.assign(num_pconstr, 0);
    }
    void set_x0(const double* x) override { for (int i = 0; i < n; ++i) x0[i] = x[i]; }
    void get_terminal_state(double* x) const override { for (int i = 0; i < n; ++i) x[i] = act[N - 1].x[i]; }
    double tconstr_h() const override { return num_tconstr ? tconstr[0].h : 0.0; }

    void model_dynamics(const Vec<n>& x, const Vec<m>& u, Vec<n>& xn, Vec<p>& y) {
        if constexpr (n == 14) wb->dynamics(x, u, xn, y, mode); else fb->dynamics(x, u, xn, y, mode);
    }
    void model_dynamics_par(const Vec<n>& x, const Vec<m>& u, DynDerivative<n, m, p>& d) {
        if constexpr (n == 14) wb->dynamics_par(x, u, d.A, d.B, d.C, d.D, mode); else fb->dynamics_par(x, u, d.A, d.B, d.C, d.D, mode);
    }
    void plan_foothold() {
        if constexpr (n == 6) fb->plan_foothold(x0.d, dt * (double)N, mode);
    }
    void path_constraint(const ModelState<n, m, p>& s) {
        if constexpr (n == 14) wbcon->path_constraint(s, pconstr.empty() ? nullptr : pconstr.data(), mode);
    }
    void terminal_constraint(const ModelState<n, m, p>& s) {
        if constexpr (n == 14) wbcon->terminal_constraint(s, tconstr.empty() ? nullptr : tconstr.data(), mode);
    }
    void control_update(int k, double eps) {  // SinglePhase.cpp:76
        Vec<n> dx = act[k].x - nom[k].x;
        act[k].u = (nom[k].u + eps * CTG[k].du) + CTG[k].K * dx;
    }
    // SinglePhase.cpp:62-114
    void forward_sweep(double eps) override {
        V = 0;
        act[0].x = x0;
        if (!is_wb) plan_foothold();
        for (int k = 0; k < N - 1; ++k) {
            control_update(k, eps);
            model_dynamics(act[k].x, act[k].u, act[k + 1].x, act[k].y);
            model_dynamics_par(act[k].x, act[k].u, dynpar[k]);
            cost->running_cost(act[k], ref[k], rcost[k], mode);
            cost->running_cost_par(act[k], ref[k], rcost[k], mode);
            path_constraint(act[k]);
            if (option->ReB_active && !param.reb_empty) update_running_cost_with_pconstr(rcost[k], CALC_DYN_AND_PAR);
            V += rcost[k].l;
        }
        cost->terminal_cost(act[N - 1], ref[N - 1], tcost, mode);
        cost->terminal_cost_par(act[N - 1], ref[N - 1], tcost, mode);
        terminal_constraint(act[N - 1]);
        if (option->AL_active & (!param.al_empty)) update_terminal_cost_with_tconstr(CALC_DYN_AND_PAR);
        V += tcost.Phi;
    }
    // SinglePhase.cpp:117-144
    void forward_sweep_dynamics_only(double eps) override {
        V = 0;
        act[0].x = x0;
        if (!is_wb) plan_foothold();
        for (int k = 0; k < N - 1; ++k) {
            control_update(k, eps);
            model_dynamics(act[k].x, act[k].u, act[k + 1].x, act[k].y);
            cost->running_cost(act[k], ref[k], rcost[k], mode);
            path_constraint(act[k]);
            if (option->ReB_active && !param.reb_empty) update_running_cost_with_pconstr(rcost[k], CALC_DYNAMICS_ONLY);
            V += rcost[k].l;
        }
        cost->terminal_cost(act[N - 1], ref[N - 1], tcost, mode);
        terminal_constraint(act[N - 1]);
        if (option->AL_active & (!param.al_empty)) update_terminal_cost_with_tconstr(CALC_DYNAMICS_ONLY);
        V += tcost.Phi;
    }
    // SinglePhase.cpp:147-180
    void forward_sweep_partials_only() override {
        if (!is_wb) plan_foothold();
        for (int k = 0; k < N - 1; ++k) {
            model_dynamics_par(act[k].x, act[k].u, dynpar[k]);
            cost->running_cost_par(act[k], ref[k], rcost[k], mode);
            path_constraint(act[k]);
            if (option->ReB_active && !param.reb_empty) update_running_cost_with_pconstr(rcost[k], CALC_PARTIALS_ONLY);
        }
        cost->terminal_cost_par(act[N - 1], ref[N - 1], tcost, mode);
        terminal_constraint(act[N - 1]);
        if (option->AL_active & (!param.al_empty)) update_terminal_cost_with_tconstr(CALC_PARTIALS_ONLY);
    }
    // MHPC_CompoundTypes.h:116-126
    static void compute_Qfunction(CostToGoStruct<n, m>& c, const RCostStruct<n, m, p>& rc, const DynDerivative<n, m, p>& dp,
                                  const Vec<n>& Gn, const Mat<n, n>& Hn) {
        c.Qx = (rc.lx + tmul(dp.A, Gn)) + tmul(dp.C, rc.ly);
        c.Qu = (rc.lu + tmul(dp.B, Gn)) + tmul(dp.D, rc.ly);
        c.Qxx = (rc.lxx + tmul(dp.C, rc.lyy) * dp.C) + tmul(dp.A, Hn) * dp.A;
        c.Quu = (rc.luu + tmul(dp.D, rc.lyy) * dp.D) + tmul(dp.B, Hn) * dp.B;
        c.Qux = (rc.lux + tmul(dp.D, rc.lyy) * dp.C) + tmul(dp.B, Hn) * dp.A;
    }
    // MHPC_CompoundTypes.h:128-144
    static double valuefunction_update(CostToGoStruct<n, m>& c) {
        static_assert(m == 4, "Quu.inverse() restated for 4x4 only");
        Mat<m, m> inv = inverse4(c.Quu);
        Mat<m, m> Quu_inv = (inv + T(inv)) / 2;
        c.Qxx = (c.Qxx + T(c.Qxx)) / 2;
        c.du = (-Quu_inv) * c.Qu;
        c.K = (-Quu_inv) * c.Qux;
        Mat<n, m> QuxT_Qinv = tmul(c.Qux, Quu_inv);
        c.G = c.Qx - QuxT_Qinv * c.Qu;
        c.H = c.Qxx - QuxT_Qinv * c.Qux;
        Mat<1, m> t = tmul(-c.Qu, inv);
        return (t * c.Qu)[0];
    }
    // SinglePhase.cpp:183-216
    bool backward_sweep(double reg) override {
        bool success = true;
        dV = dVnext;
        CTG[N - 1].G = tcost.Phix + Gnext;
        CTG[N - 1].H = tcost.Phixx + Hnext;
        const double tiny = std::pow(0.1, 9);
        for (int k = N - 2; k >= 0; --k) {
            compute_Qfunction(CTG[k], rcost[k], dynpar[k], CTG[k + 1].G, CTG[k + 1].H);
            for (int i = 0; i < n; ++i) CTG[k].Qxx(i, i) += 1.0 * reg;
            for (int i = 0; i < m; ++i) CTG[k].Quu(i, i) += 1.0 * reg;
            Mat<m, m> Qt = CTG[k].Quu;
            for (int i = 0; i < m; ++i) Qt(i, i) -= 1.0 * tiny;
            if (!ldlt_is_positive<m>(Qt)) { success = false; break; }
            dV += valuefunction_update(CTG[k]);
        }
        return success;
    }
    // SinglePhase.cpp:298-317
    void reduced_barrier() {
        const int k = 2;
        for (int i = 0; i < num_pconstr; ++i) {
            const double g = pconstr[i].g, d = param.delta[i];
            if (g > d) {
                B[i] = -std::log(g);
                Bz[i] = -1.0 / g;
                Bzz[i] = std::pow(g, -2);
            } else {
                const double r = (g - k * d) / ((k - 1) * d);
                B[i] = (double)(k - 1) / k * (std::pow(r, k) - 1) - std::log(d);
                Bz[i] = std::pow(r, (k - 1)) / d;
                Bzz[i] = std::pow(r, k - 2);
            }
        }
    }
    // SinglePhase.cpp:219-249 (gxx = guu = gyy = 0 for every constraint of the reference)
    void update_running_cost_with_pconstr(RCostStruct<n, m, p>& rc, int calcflag) {
        if (pconstr.empty()) return;
        std::fill(B.begin(), B.end(), 0.0); std::fill(Bz.begin(), Bz.end(), 0.0); std::fill(Bzz.begin(), Bzz.end(), 0.0);
        reduced_barrier();
        for (int i = 0; i < num_pconstr; ++i) {
            const double e = param.eps_ReB[i];
            const IneqConstrStruct<n, m, p>& q = pconstr[i];
            if (calcflag == CALC_DYNAMICS_ONLY || calcflag == CALC_DYN_AND_PAR) rc.l += e * B[i] * dt;
            if (calcflag == CALC_PARTIALS_ONLY || calcflag == CALC_DYN_AND_PAR) {
                for (int a = 0; a < n; ++a) rc.lx[a] += e * Bz[i] * q.gx[a] * dt;
                for (int a = 0; a < m; ++a) rc.lu[a] += e * Bz[i] * q.gu[a] * dt;
                for (int a = 0; a < p; ++a) rc.ly[a] += e * Bz[i] * q.gy[a] * dt;
                for (int b = 0; b < n; ++b) for (int a = 0; a < n; ++a) rc.lxx(a, b) += e * (q.gx[a] * Bzz[i] * q.gx[b] + Bz[i] * 0.0) * dt;
                for (int b = 0; b < m; ++b) for (int a = 0; a < m; ++a) rc.luu(a, b) += e * (q.gu[a] * Bzz[i] * q.gu[b] + Bz[i] * 0.0) * dt;
                for (int b = 0; b < p; ++b) for (int a = 0; a < p; ++a) rc.lyy(a, b) += e * (q.gy[a] * Bzz[i] * q.gy[b] + Bz[i] * 0.0) * dt;
            }
        }
    }
    // SinglePhase.cpp:257-275 -- NOTE both tests use the same flags
    void update_terminal_cost_with_tconstr(int calcflag) {
        if (tconstr.empty()) return;
        const double sigma = param.sigma;
        for (int i = 0; i < num_tconstr; ++i) {
            const TConstrStruct<n>& t = tconstr[i];
            const double lam = param.lambda[i];
            if (calcflag == CALC_DYNAMICS_ONLY || calcflag == CALC_DYN_AND_PAR) {
                const double a = sigma * t.h / 2;
                tcost.Phi += 50 * (a * a + lam * t.h);
            }
            if (calcflag == CALC_DYNAMICS_ONLY || calcflag == CALC_DYN_AND_PAR) {
                const double s2 = sigma * sigma / 2;
                for (int a = 0; a < n; ++a) tcost.Phix[a] += 50 * (s2 * t.hx[a] * t.h + lam * t.hx[a]);
                for (int b = 0; b < n; ++b)
                    for (int a = 0; a < n; ++a)
                        tcost.Phixx(a, b) += 50 * (s2 * (t.hx[a] * t.hx[b] + t.h * t.hxx(a, b)) + lam * t.hxx(a, b));
            }
        }
    }
    // SinglePhase.cpp:334-354
    void udpate_AL_ReB_Param() override {
        for (size_t e = 0; e < param.lambda.size(); ++e) param.lambda[e] += param.sigma * tconstr[e].h;
        param.sigma *= option->update_penalty;
        if (option->ReB_active) {
            for (size_t i = 0; i < param.delta.size(); ++i) {
                param.delta[i] *= option->update_relax;
                if (param.delta[i] < param.delta_min[i]) param.delta[i] = param.delta_min[i];
                param.eps_ReB[i] *= option->update_ReB;
            }
        }
    }
    void update_nominal_trajectory() override { std::memcpy((void*)nom.data(), (const void*)act.data(), N * sizeof(ModelState<n, m, p>)); }  // :358-361
    double get_tconstr_violation() const override {
        double s = 0;
        for (int i = 0; i < num_tconstr; ++i) s += tconstr[i].h * tconstr[i].h;
        return s;
    }
    void get_CTG0(double& dV_, double* G, double* H) const override {  // :320-325
        dV_ = dV;
        std::memcpy(G, CTG[0].G.d, n * sizeof(double));
        std::memcpy(H, CTG[0].H.d, n * n * sizeof(double));
    }
    void set_CTG_next(double dVn, const double* G, const double* H) override {
        dVnext = dVn;
        std::memcpy(Gnext.d, G, n * sizeof(double));
        std::memcpy(Hnext.d, H, n * n * sizeof(double));
    }
    // flat export for the tests: returns doubles per knot, fills out[k*stride + e] for k < N
    int get_field(int field, double* out) const override {
        auto put = [&](int stride, auto getter) {
            if (out) for (int k = 0; k < N; ++k) { const double* s = getter(k); std::memcpy(out + (size_t)k * stride, s, stride * sizeof(double)); }
            return stride;
        };
        switch (field) {
            case ORC_F_X_ACT: return put(n, [&](int k) { return act[k].x.d; });
            case ORC_F_U_ACT: return put(m, [&](int k) { return act[k].u.d; });
            case ORC_F_Y_ACT: return put(p, [&](int k) { return act[k].y.d; });
            case ORC_F_X_NOM: return put(n, [&](int k) { return nom[k].x.d; });
            case ORC_F_U_NOM: return put(m, [&](int k) { return nom[k].u.d; });
            case ORC_F_Y_NOM: return put(p, [&](int k) { return nom[k].y.d; });
            case ORC_F_X_REF: return put(n, [&](int k) { return ref[k].x.d; });
            case ORC_F_U_REF: return put(m, [&](int k) { return ref[k].u.d; });
            case ORC_F_Y_REF: return put(p, [&](int k) { return ref[k].y.d; });
            case ORC_F_A: return put(n * n, [&](int k) { return dynpar[k].A.d; });
            case ORC_F_B: return put(n * m, [&](int k) { return dynpar[k].B.d; });
            case ORC_F_C: return put(p * n, [&](int k) { return dynpar[k].C.d; });
            case ORC_F_D: return put(p * m, [&](int k) { return dynpar[k].D.d; });
            case ORC_F_LX: return put(n, [&](int k) { return rcost[k].lx.d; });
            case ORC_F_LU: return put(m, [&](int k) { return rcost[k].lu.d; });
            case ORC_F_LY: return put(p, [&](int k) { return rcost[k].ly.d; });
            case ORC_F_LXX: return put(n * n, [&](int k) { return rcost[k].lxx.d; });
            case ORC_F_LUU: return put(m * m, [&](int k) { return rcost[k].luu.d; });
            case ORC_F_LYY: return put(p * p, [&](int k) { return rcost[k].lyy.d; });
            case ORC_F_K: return put(m * n, [&](int k) { return CTG[k].K.d; });
            case ORC_F_DU: return put(m, [&](int k) { return CTG[k].du.d; });
            case ORC_F_G: return put(n, [&](int k) { return CTG[k].G.d; });
            case ORC_F_H: return put(n * n, [&](int k) { return CTG[k].H.d; });
            case ORC_F_PHIX: if (out) std::memcpy(out, tcost.Phix.d, n * sizeof(double)); return n;
            case ORC_F_PHIXX: if (out) std::memcpy(out, tcost.Phixx.d, n * n * sizeof(double)); return n * n;
            case ORC_F_QX: return put(n, [&](int k) { return CTG[k].Qx.d; });
            case ORC_F_QU: return put(m, [&](int k) { return CTG[k].Qu.d; });
            case ORC_F_QXX: return put(n * n, [&](int k) { return CTG[k].Qxx.d; });
            case ORC_F_QUX: return put(m * n, [&](int k) { return CTG[k].Qux.d; });
            case ORC_F_QUU: return put(m * m, [&](int k) { return CTG[k].Quu.d; });
        }
        return -1;
    }
};

// =====================================================================================
// HSDDPSolver/source/MultiPhaseDDP.cpp + Controller/MHPCLocomotion/MHPCLocomotion.cpp
// =====================================================================================
class Solver {
public:
    Problem prob;
    Option option;  // the "shared" option every phase points at (MHPCLocomotion.cpp:77,86)
    int n_phases, n_wb, n_fb;
    double dt_wb, dt_fb;
    std::vector<int> mode_seq, N_TIMESTEPS;
    std::vector<float> timings;
    std::unique_ptr<WBModel> wbmodel;
    std::unique_ptr<FootholdPlanner> planner;
    std::unique_ptr<FBModel> fbmodel;
    CostW<14, 4, 4> wbcost;
    CostW<6, 4, 4> fbcost;
    WBConstraint wbcon;
    std::vector<std::unique_ptr<PhaseBase>> phases;
    Mat<6, 14> stateProj;
    Vec<14> x0;
    double actual_cost = 0, exp_cost_change = 0, tconstr_violation = 0;
    std::vector<double> Gnext, Hnext;
    double dVnext = 0;
    std::vector<IterRecord> records;
    int status = 0;  // 0 ok, 1 regularisation abort (MultiPhaseDDP.cpp:221-226)
    int last_iter_AL = 0;
    std::vector<double> Px_store;  // Px used at each phase boundary in the last backward sweep (14x14 each)

    explicit Solver(const Problem& p) : prob(p), option(p.opt) {
        // MHPCLocomotion.cpp:8-43
        n_wb = p.n_wbphase; n_fb = p.n_fbphase; n_phases = n_wb + n_fb;
        dt_wb = p.dt_wb; dt_fb = p.dt_fb;  // float -> double (MHPCLocomotion.h:59)
        wbmodel.reset(new WBModel(dt_wb));
        planner.reset(new FootholdPlanner(wbmodel.get(), 1.5, -0.404));
        fbmodel.reset(new FBModel(planner.get(), dt_fb));
        wbcost = make_wb_cost(dt_wb);
        fbcost = make_fb_cost(dt_fb);
        stateProj.setZero();
        for (int i = 0; i < 3; ++i) { stateProj(i, i) = 1; stateProj(3 + i, 7 + i) = 1; }
        for (int i = 0; i < 14; ++i) x0[i] = p.x0[i];
        for (int i = 0; i < n_phases; ++i) {
            if (i < n_wb) { auto* ph = new Phase<14, 4, 4>(); ph->is_wb = true; phases.emplace_back(ph); }
            else { auto* ph = new Phase<6, 4, 4>(); ph->is_wb = false; phases.emplace_back(ph); }
        }
        Px_store.assign((size_t)std::max(n_phases, 1) * 196, 0.0);
    }
    int gait_next_mode(int cur) const {  // Gait.h:46-53
        for (int i = 0; i < 4; ++i) if (cur == prob.gait_mode[i]) return prob.gait_mode[(i + 1) % 4];
        return cur;
    }
    // MHPCLocomotion.cpp:47-53
    void initialization() {
        memory_reset();
        build_problem();
        warmstart();
    }
    // MHPCLocomotion.cpp:265-290
    void memory_reset() {
        for (auto& ph : phases) ph->zero_data();
    }
    // MHPCLocomotion.cpp:107-158 (dead code in the reference's main, but it compiles and defines the receding-horizon
    // shift): the whole-body / rigid-body phase buffers rotate by one, the gait advances one mode, every phase is
    // re-initialised (AL/ReB parameters back to their initial values) and the reference is regenerated from _x0.
    void update_problem() {
        for (int i = 0; i + 1 < n_wb; ++i) phases[i]->swap_data(*phases[i + 1]);
        if (n_fb > 0)
            for (int i = n_wb; i + 1 < n_phases; ++i) phases[i]->swap_data(*phases[i + 1]);
        prob.cmode = gait_next_mode(prob.cmode);
        mode_seq[0] = prob.cmode;
        for (int i = 0; i + 1 < n_phases; ++i) mode_seq[i + 1] = gait_next_mode(mode_seq[i]);
        for (int i = 0; i < n_phases; ++i) timings[i] = prob.gait_timing[mode_seq[i] - 1];
        for (int i = 0; i < n_phases; ++i) {
            PhaseBase* ph = phases[i].get();
            N_TIMESTEPS[i] = (int)std::round(timings[i] / ph->dt);
            ph->mode = mode_seq[i]; ph->phaseidx = i; ph->N = N_TIMESTEPS[i];
        }
        for (int i = 0; i < n_phases; ++i) phases[i]->initialization();
        generate_ref();
    }
    void set_x0(const double* x) { for (int i = 0; i < 14; ++i) x0[i] = x[i]; }
    // MHPCLocomotion.cpp:63-104
    void build_problem() {
        mode_seq.assign(n_phases, 1); timings.assign(n_phases, 0.f); N_TIMESTEPS.assign(n_phases, 0);
        mode_seq[0] = prob.cmode;  // Gait.h:55-66
        for (int i = 0; i + 1 < n_phases; ++i) mode_seq[i + 1] = gait_next_mode(mode_seq[i]);
        for (int i = 0; i < n_phases; ++i) timings[i] = prob.gait_timing[mode_seq[i] - 1];  // Gait.h:68-77
        for (int i = 0; i < n_phases; ++i) {
            PhaseBase* ph = phases[i].get();
            ph->option = &option; ph->mode = mode_seq[i]; ph->phaseidx = i;
            if (i < n_wb) {
                N_TIMESTEPS[i] = (int)std::round(timings[i] / dt_wb);
                auto* w = static_cast<Phase<14, 4, 4>*>(ph);
                w->wb = wbmodel.get(); w->cost = &wbcost; w->wbcon = &wbcon; w->dt = dt_wb; w->N = N_TIMESTEPS[i];
                w->initialization();
            } else {
                N_TIMESTEPS[i] = (int)std::round(timings[i] / dt_fb);
                auto* f = static_cast<Phase<6, 4, 4>*>(ph);
                f->fb = fbmodel.get(); f->cost = &fbcost; f->dt = dt_fb; f->N = N_TIMESTEPS[i];
                f->initialization();
            }
        }
        generate_ref();
    }
    // HSDDPSolver/header/ReferenceGen.h:53-109 + source/ReferenceGen.cpp:23-66
    void generate_ref() {
        const double vel = prob.vel, height = prob.height;  // float -> double
        const double GRF = 8.252 * 9.81;
        double xterm[4][14] = {{0, -0.1432, -PI / 25, 0.35 * PI, -0.65 * PI, 0.35 * PI, -0.6 * PI, vel, 1, 0, 0, 0, 0, 0},
                               {0, -0.1418, PI / 35, 0.2 * PI, -0.58 * PI, 0.25 * PI, -0.7 * PI, vel, -1, 0, 0, 0, 0, 0},
                               {0, -0.1325, -PI / 40, 0.33 * PI, -0.48 * PI, 0.33 * PI, -0.75 * PI, vel, 1, 0, 0, 0, 0, 0},
                               {0, -0.1490, -PI / 25, 0.35 * PI, -0.7 * PI, 0.25 * PI, -0.60 * PI, vel, -1, 0, 0, 0, 0, 0}};
        const double qb[4] = {0.3 * PI, -0.7 * PI, 0.3 * PI, -0.7 * PI};
        std::vector<std::vector<double>> pos(n_phases, std::vector<double>(N_TIMESTEPS_MAX, 0.0));
        pos[0][0] = x0[0];
        for (int pi = 0; pi < n_phases; ++pi) {
            const double dt = (pi < n_wb) ? dt_wb : dt_fb;
            if (pi > 0) pos[pi][0] = pos[pi - 1][N_TIMESTEPS[pi - 1] - 1];
            for (int k = 1; k < N_TIMESTEPS[pi]; ++k) pos[pi][k] = pos[pi][k - 1] + vel * dt;
        }
        for (int pi = 0; pi < n_phases; ++pi) {
            const int md = mode_seq[pi], N = N_TIMESTEPS[pi];
            if (pi < n_wb) {
                auto* w = static_cast<Phase<14, 4, 4>*>(phases[pi].get());
                for (int k = 0; k < N - 1; ++k) {
                    const double xr[14] = {pos[pi][k], height, 0, qb[0], qb[1], qb[2], qb[3], vel, 0, 0, 0, 0, 0, 0};
                    std::memcpy(w->ref[k].x.d, xr, sizeof(xr));
                    w->ref[k].y[0] = 0; w->ref[k].y[1] = GRF; w->ref[k].y[2] = 0; w->ref[k].y[3] = GRF;
                }
                std::memcpy(w->ref[N - 1].x.d, xterm[md - 1], 14 * sizeof(double));
                w->ref[N - 1].x[0] = pos[pi][N - 1];
            } else {
                auto* f = static_cast<Phase<6, 4, 4>*>(phases[pi].get());
                for (int k = 0; k < N - 1; ++k) {
                    const double xr[6] = {pos[pi][k], height, 0, vel, 0, 0};
                    std::memcpy(f->ref[k].x.d, xr, sizeof(xr));
                    f->ref[k].u[0] = 0; f->ref[k].u[1] = GRF; f->ref[k].u[2] = 0; f->ref[k].u[3] = GRF;
                }
                const double tail[5] = {height, 0, vel, 0, 0};
                for (int i = 0; i < 5; ++i) f->ref[N - 1].x[1 + i] = tail[i];
                f->ref[N - 1].x[0] = pos[pi][N - 1];
            }
        }
    }
    // Controller/Heuristics/boundingPDControl.cpp:3-46
    void bounding_PDcontrol(Phase<14, 4, 4>* ph) {
        const double qjoint_nom[4] = {PI / 4, -PI * 7 / 12, PI / 4, -PI * 7 / 12};
        const double legext_nom = 0.2462, Kspring = 2200;
        const double kp[4] = {5 * 8.0, 5 * 1.0, 5 * 12.0, 5 * 10.0};
        auto& ms = ph->act;
        for (int k = 0; k < ph->N - 1; ++k) {
            Mat<2, 7> J, Jd; J.setZero(); Jd.setZero();
            Vec<7> q; for (int i = 0; i < 7; ++i) q[i] = ms[k].x[i];
            const int md = ph->mode;
            if (md == 1 || md == 3) {
                const int leg = (md == 1) ? 1 : 0;
                const double gain = (md == 1) ? 3 : 2.2;
                wbmodel->FootJacobian(ms[k].x, J, Jd, leg);
                double lv[2]; wbmodel->get_leg_ext_vec(q, leg, lv);
                const double sq = lv[0] * lv[0] + lv[1] * lv[1];
                const double nrm = std::sqrt(sq);
                double nv[2] = {lv[0], lv[1]};
                if (sq > 0) { nv[0] = lv[0] / std::sqrt(sq); nv[1] = lv[1] / std::sqrt(sq); }
                double F[2];
                for (int i = 0; i < 2; ++i) F[i] = ((-nv[i]) * Kspring) * (nrm - legext_nom);
                for (int j = 0; j < 4; ++j) ms[k].u[j] = (J(0, 3 + j) * F[0] + J(1, 3 + j) * F[1]) * gain;
            } else {
                for (int j = 0; j < 4; ++j) ms[k].u[j] = kp[j] * (qjoint_nom[j] - ms[k].x[3 + j]) - 1.0 * ms[k].x[10 + j];
            }
            wbmodel->dynamics(ms[k].x, ms[k].u, ms[k + 1].x, ms[k].y, md);
        }
    }
    // MHPCLocomotion.cpp:200-215
    void warmstart() {
        Vec<14> x0_phase = x0;
        for (int i = 0; i < n_wb; ++i) {
            auto* w = static_cast<Phase<14, 4, 4>*>(phases[i].get());
            w->act[0].x = x0_phase;
            bounding_PDcontrol(w);
            if (i + 1 < n_wb) {
                std::vector<double> xn = phase_transition(i);
                for (int j = 0; j < 14; ++j) x0_phase[j] = xn[j];
            }
        }
        update_nominal_trajectory();
    }
    // MultiPhaseDDP.cpp:351-379
    std::vector<double> phase_transition(int ci) {
        PhaseBase* c = phases[ci].get();
        PhaseBase* nx = phases[ci + 1].get();
        if (c->is_wb) {
            Vec<14> xend, xin; Vec<4> y;
            c->get_terminal_state(xend.d);
            xin.setZero(); y.setZero();
            wbmodel->resetmap(xend, xin, y, c->mode);
            if (nx->is_wb) return std::vector<double>(xin.d, xin.d + 14);
            Vec<6> xr = stateProj * xin;
            return std::vector<double>(xr.d, xr.d + 6);
        }
        std::vector<double> x(6);
        c->get_terminal_state(x.data());
        return x;
    }
    template <class F>
    void sweep(F call) {  // MultiPhaseDDP.cpp:28-76
        actual_cost = 0;
        std::vector<double> xs(x0.d, x0.d + 14);
        tconstr_violation = 0;
        for (int i = 0; i < n_phases; ++i) {
            phases[i]->set_x0(xs.data());
            call(phases[i].get());
            if (i + 1 < n_phases) xs = phase_transition(i);
            actual_cost += phases[i]->V;
            tconstr_violation += phases[i]->get_tconstr_violation();
        }
        tconstr_violation = std::sqrt(tconstr_violation);
    }
    void forward_sweep(double eps) { sweep([&](PhaseBase* p) { p->forward_sweep(eps); }); }
    void forward_sweep_dynamics_only(double eps) { sweep([&](PhaseBase* p) { p->forward_sweep_dynamics_only(eps); }); }
    void forward_sweep_partials_only() { for (auto& p : phases) p->forward_sweep_partials_only(); }  // :84-91
    void update_nominal_trajectory() { for (auto& p : phases) p->update_nominal_trajectory(); }
    void update_AL_ReB_param() { for (auto& p : phases) p->udpate_AL_ReB_Param(); }

    // MultiPhaseDDP.cpp:300-341
    void impact_aware_step(int ci, double& dV, std::vector<double>& G, std::vector<double>& H) {
        PhaseBase* c = phases[ci].get();
        PhaseBase* nx = phases[ci + 1].get();
        const int nn = nx->xdim;
        std::vector<double> Gp(nn, 0.0), Hp((size_t)nn * nn, 0.0);
        nx->get_CTG0(dV, Gp.data(), Hp.data());
        if (c->is_wb) {
            Mat<14, 14> Px; Vec<14> xend;
            Px.setZero();
            c->get_terminal_state(xend.d);
            wbmodel->resetmap_par(xend, Px, c->mode);
            std::memcpy(&Px_store[(size_t)ci * 196], Px.d, 196 * sizeof(double));
            Vec<14> Gv; Mat<14, 14> Hv;
            if (nx->is_wb) {
                Vec<14> g; Mat<14, 14> h;
                std::memcpy(g.d, Gp.data(), sizeof(g.d)); std::memcpy(h.d, Hp.data(), sizeof(h.d));
                Gv = tmul(Px, g);
                Hv = tmul(Px, h) * Px;
            } else {
                Vec<6> g; Mat<6, 6> h;
                std::memcpy(g.d, Gp.data(), sizeof(g.d)); std::memcpy(h.d, Hp.data(), sizeof(h.d));
                Mat<14, 6> PxT_PT = tmul(Px, T(stateProj));
                Gv = PxT_PT * g;
                Hv = ((PxT_PT * h) * stateProj) * Px;
            }
            G.assign(Gv.d, Gv.d + 14); H.assign(Hv.d, Hv.d + 196);
        } else {
            G = Gp; H = Hp;
        }
    }
    // MultiPhaseDDP.cpp:100-127
    bool backward_sweep(double reg) {
        bool success = true;
        dVnext = 0;
        const int nl = phases[n_phases - 1]->xdim;
        Gnext.assign(nl, 0.0); Hnext.assign((size_t)nl * nl, 0.0);
        for (int i = n_phases - 1; i >= 0; --i) {
            if (i + 1 < n_phases) impact_aware_step(i, dVnext, Gnext, Hnext);
            phases[i]->set_CTG_next(dVnext, Gnext.data(), Hnext.data());
            if (!phases[i]->backward_sweep(reg)) { success = false; break; }
        }
        if (success) exp_cost_change = phases[0]->dV;
        return success;
    }
    double last_eps = 0;
    // MultiPhaseDDP.cpp:130-151
    int forward_iteration() {
        double eps = 1;
        const double cost_prev = actual_cost;
        int iter = 1;
        while (eps > std::pow(0.1, 10)) {
            forward_sweep_dynamics_only(eps);
            last_eps = eps;
            if (actual_cost <= cost_prev + option.gamma * eps * (1 - eps / 2) * exp_cost_change) break;
            eps *= option.alpha;
            iter++;
        }
        return iter;
    }
    // MultiPhaseDDP.cpp:154-289
    void solve() {
        int iter_AL = 1;
        double cost_prev;
        bool bws_success;
        const double update_penalty = option.update_penalty;
        const bool ReB_active = option.ReB_active;
        records.clear();
        status = 0;
        while (iter_AL <= option.max_AL_iter) {
            last_iter_AL = iter_AL;
            option.ReB_active = ReB_active;
            if ((tconstr_violation > 0.05) || 1 == iter_AL) option.ReB_active = 0;
            forward_sweep(0);
            update_nominal_trajectory();
            int iter_DDP = 1;
            double regularization = 0;
            while (iter_DDP <= option.max_DDP_iter) {
                cost_prev = actual_cost;
                bws_success = false;
                int bws_iter = 1;
                while (!bws_success) {
                    bws_success = backward_sweep(regularization);
                    if (bws_success) break;
                    regularization = std::max(regularization * option.update_regularization, 1e-03);
                    bws_iter++;
                    if (regularization > 1000) { status = 1; return; }
                }
                const double reg_used = regularization;
                regularization = regularization / 20;
                if (regularization < 1e-06) regularization = 0;
                const int fit_iter = forward_iteration();
                update_nominal_trajectory();
                records.push_back(IterRecord{iter_AL, iter_DDP, bws_iter, fit_iter, reg_used, last_eps, actual_cost,
                                             exp_cost_change, tconstr_violation, cost_prev});
                if (cost_prev - actual_cost < option.DDP_thresh) break;
                forward_sweep_partials_only();
                iter_DDP++;
            }
            option.update_penalty = update_penalty;
            if (tconstr_violation < 0.03) option.update_penalty = 0;
            update_AL_ReB_param();
            if (tconstr_violation < option.AL_thresh) break;
            iter_AL++;
        }
    }
};

}  // namespace orc

// =====================================================================================
// C interface for ctypes (tests / bench CPU baseline only)
// =====================================================================================
using orc::Solver;
extern "C" {

void* orc_create(const orc_problem* p) {
    orc::Problem q;
    q.opt.alpha = p->alpha; q.opt.gamma = p->gamma; q.opt.update_penalty = p->update_penalty;
    q.opt.update_relax = p->update_relax; q.opt.update_regularization = p->update_regularization;
    q.opt.update_ReB = p->update_ReB; q.opt.max_DDP_iter = p->max_DDP_iter; q.opt.max_AL_iter = p->max_AL_iter;
    q.opt.DDP_thresh = p->DDP_thresh; q.opt.AL_thresh = p->AL_thresh;
    q.opt.AL_active = p->AL_active != 0; q.opt.ReB_active = p->ReB_active != 0; q.opt.smooth_active = p->smooth_active != 0;
    q.n_wbphase = p->n_wbphase; q.n_fbphase = p->n_fbphase; q.dt_wb = p->dt_wb; q.dt_fb = p->dt_fb; q.cmode = p->cmode;
    q.vel = p->vel; q.height = p->height;
    for (int i = 0; i < 4; ++i) { q.gait_mode[i] = p->gait_mode[i]; q.gait_timing[i] = p->gait_timing[i]; }
    for (int i = 0; i < 14; ++i) q.x0[i] = p->x0[i];
    return new Solver(q);
}
void orc_destroy(void* h) { delete (Solver*)h; }
void orc_set_faithful_alloc(int on) { orc::g_faithful_alloc = on != 0; }
void orc_initialization(void* h) { ((Solver*)h)->initialization(); }
void orc_forward_sweep(void* h, double eps) { ((Solver*)h)->forward_sweep(eps); }
void orc_forward_sweep_dynamics_only(void* h, double eps) { ((Solver*)h)->forward_sweep_dynamics_only(eps); }
void orc_forward_sweep_partials_only(void* h) { ((Solver*)h)->forward_sweep_partials_only(); }
int orc_backward_sweep(void* h, double reg) { return ((Solver*)h)->backward_sweep(reg) ? 1 : 0; }
int orc_forward_iteration(void* h) { return ((Solver*)h)->forward_iteration(); }
void orc_update_nominal_trajectory(void* h) { ((Solver*)h)->update_nominal_trajectory(); }
void orc_update_AL_ReB_param(void* h) { ((Solver*)h)->update_AL_ReB_param(); }
void orc_update_problem(void* h) { ((Solver*)h)->update_problem(); }
void orc_set_x0(void* h, const double* x0) { ((Solver*)h)->set_x0(x0); }
void orc_set_ReB_active(void* h, int on) { ((Solver*)h)->option.ReB_active = on != 0; }
void orc_set_update_penalty(void* h, double v) { ((Solver*)h)->option.update_penalty = v; }
int orc_solve(void* h) { ((Solver*)h)->solve(); return ((Solver*)h)->status; }
int orc_n_phases(void* h) { return ((Solver*)h)->n_phases; }
void orc_get_schedule(void* h, int* mode, int* N, int* is_wb) {
    Solver* s = (Solver*)h;
    for (int i = 0; i < s->n_phases; ++i) { mode[i] = s->mode_seq[i]; N[i] = s->N_TIMESTEPS[i]; is_wb[i] = i < s->n_wb; }
}
void orc_get_scalars(void* h, double* out) {
    Solver* s = (Solver*)h;
    out[0] = s->actual_cost; out[1] = s->exp_cost_change; out[2] = s->tconstr_violation; out[3] = s->last_eps;
}
int orc_get_field(void* h, int phase, int field, double* out) { return ((Solver*)h)->phases[phase]->get_field(field, out); }
void orc_get_phase_params(void* h, int phase, double* out /* sigma, lambda, delta[0], delta[8], delta[16], eps[0], eps[8], eps[16], h */) {
    Solver* s = (Solver*)h;
    const orc::ALReBParam& a = s->phases[phase]->param;
    auto at = [](const std::vector<double>& v, size_t i) { return i < v.size() ? v[i] : 0.0; };
    out[0] = a.sigma; out[1] = at(a.lambda, 0);
    out[2] = at(a.delta, 0); out[3] = at(a.delta, 8); out[4] = at(a.delta, 16);
    out[5] = at(a.eps_ReB, 0); out[6] = at(a.eps_ReB, 8); out[7] = at(a.eps_ReB, 16);
    out[8] = s->phases[phase]->tconstr_h();
}
void orc_get_Px(void* h, int phase, double* out) { std::memcpy(out, &((Solver*)h)->Px_store[(size_t)phase * 196], 196 * sizeof(double)); }
int orc_n_records(void* h) { return (int)((Solver*)h)->records.size(); }
void orc_get_record(void* h, int i, int* ints /*4*/, double* dbl /*6*/) {
    const orc::IterRecord& r = ((Solver*)h)->records[i];
    ints[0] = r.iter_AL; ints[1] = r.iter_DDP; ints[2] = r.n_bws; ints[3] = r.n_fit;
    dbl[0] = r.reg_used; dbl[1] = r.eps_accepted; dbl[2] = r.cost; dbl[3] = r.exp_cost_change; dbl[4] = r.viol; dbl[5] = r.cost_prev;
}

// Direct evaluation of one generated function into dense column-major outputs (zero-filled first),
// i.e. what casadi_interface leaves in a freshly zeroed destination.  name: e.g. "Dyn_BS_par".
int orc_eval(const char* name, const double* a0, const double* a1, const double* a2, const double* a3,
             double* o0, double* o1, double* o2, double* o3) {
    using namespace orc;
#define TRY(F, NIN, NOUT)                                                                                   \
    if (!std::strcmp(name, #F)) {                                                                           \
        const double* in[4] = {a0, a1, a2, a3};                                                             \
        double* out[4] = {o0, o1, o2, o3};                                                                  \
        const double* arg[8]; double* res[8]; double buf[4][200];                                           \
        for (int i = 0; i < NIN; ++i) arg[i] = in[i];                                                       \
        for (int i = 0; i < NOUT; ++i) res[i] = buf[i];                                                     \
        casadi_int_t iw0[1]; double w0[1];                                                                  \
        F(arg, res, iw0, w0, 1);                                                                            \
        for (int i = 0; i < NOUT; ++i) {                                                                    \
            const casadi_int_t* sp = F##_sparsity_out(i);                                                   \
            const int nrow = (int)sp[0], ncol = (int)sp[1];                                                 \
            const casadi_int_t* ci = sp + 2; const casadi_int_t* ri = ci + ncol + 1;                        \
            std::memset(out[i], 0, sizeof(double) * nrow * ncol);                                           \
            int nz = 0;                                                                                     \
            for (int c = 0; c < ncol; ++c) while (nz < ci[c + 1]) { out[i][ri[nz] + nrow * c] = buf[i][nz]; ++nz; } \
        }                                                                                                   \
        return 0;                                                                                           \
    }
    TRY(Dyn_BS, 2, 2) TRY(Dyn_FS, 2, 2) TRY(Dyn_FL, 2, 2)
    TRY(Dyn_BS_par, 2, 4) TRY(Dyn_FS_par, 2, 4) TRY(Dyn_FL_par, 2, 4)
    TRY(Imp_F, 1, 2) TRY(Imp_B, 1, 2) TRY(Imp_F_par, 1, 1) TRY(Imp_B_par, 1, 1)
    TRY(FBDynamics, 4, 1) TRY(FBDynamics_par, 4, 2)
    TRY(WB_FL1_terminal_constr, 1, 3) TRY(WB_FL2_terminal_constr, 1, 3)
    TRY(Jacob_F, 1, 2) TRY(Jacob_B, 1, 2)
    TRY(Link0Jacobian, 2, 2) TRY(Link1Jacobian, 2, 2) TRY(Link2Jacobian, 2, 2) TRY(Link3Jacobian, 2, 2) TRY(Link4Jacobian, 2, 2)
#undef TRY
    return -1;
}

}  // extern "C"
