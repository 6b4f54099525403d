// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path
// (machline_b200/, include/): only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use it, and only as the checker.
//
// CPU restatement of MachLine's AIC assembly (lower-order linear doublet / constant source panels for the Dirichlet and
// the least-squares Neumann formulations; higher-order quadratic doublet / linear source panels for the Dirichlet
// formulations), following the reference statement by statement:
//   flow_point_in_dod                         src/flow.f90:282-310
//   panel_check_dod                           src/panel.f90:1732-1901
//   panel_calc_basic_geom                     src/panel.f90:1904-1938
//   panel_calc_subsonic_geom                  src/panel.f90:1941-1997
//   panel_calc_supersonic_subinc_geom         src/panel.f90:2000-2091
//   panel_calc_basic_F_integrals_subsonic     src/panel.f90:2232-2283
//   panel_calc_basic_F_integrals_supersonic_subinc  src/panel.f90:2286-2407
//   panel_calc_hH113_subsonic                 src/panel.f90:2472-2509
//   panel_calc_hH113_supersonic_subinc        src/panel.f90:2512-2573   (binary128 F1,F2,b)
//   panel_calc_remaining_integrals            src/panel.f90:2631-2683   (order 1 and order 2 potential integrals)
//   panel_assemble_phi_s_S_space / phi_d_M_space   src/panel.f90:2815-2914
//   panel_calc_potential_influences           src/panel.f90:2917-2971
//   panel_solver_update_system_row            src/panel_solver.f90:1203-1287
//   panel_solver_calc_body_influences         src/panel_solver.f90:1290-1501 (Dirichlet + strength matching)
//   panel_solver_calc_wake_influences         src/panel_solver.f90:1504-1706 (Dirichlet)
// Compiled with -O2 -ffp-contract=off (gfortran -O2 on x86-64 does not contract), real(16) ->
// __float128.  Superinclined panels are rejected by the reference at init (panel.f90:439-443), so
// the *_supinc branches are not restated.
//
// Parity pinning: through the golden tuples of test/test_machline.py (tests/test_oracle_golden.py), through the
// reference's stored off-body potentials at 400 field points, sub- and supersonic (tests/test_oracle_offbody.py), and
// through known-answer integrals generated from dev/unit_tests/panel.py (tests/test_oracle_integrals.py); fixtures and
// the script that made them are under tests/golden/.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#include <quadmath.h>
#include <omp.h>

#include "../include/machline_gpu.h"
#include "oracle.h"

namespace {

typedef __float128 quad;
const double pi = 3.14159265358979323846264338327950288419716939937510;

inline double inner3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline void cross3(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
inline double fsign(double a, double b) { return std::signbit(b) ? -std::fabs(a) : std::fabs(a); }

// Calibration switch (tests only): evaluate log / atan2 in binary128 and round once, i.e. a libm that is
// correctly rounded.  The difference between the two modes is the noise floor that ANY two conforming
// libms (glibc vs CUDA, or two glibc versions under the reference itself) put on an AIC entry.
// mode 2: the correctly rounded value moved by one ulp in a pseudo-random ~1/3 of the calls, i.e. a libm whose
// results are "faithfully" but not correctly rounded (what CUDA's 1-2 ulp log/atan2 look like from outside).
int g_exact_libm = 0;
inline double o_jitter(double v) {
    if (g_exact_libm != 2) return v;
    unsigned long long u;
    std::memcpy(&u, &v, 8);
    unsigned h = (unsigned)((u * 0x9E3779B97F4A7C15ull) >> 61);   // 0..7 from the value's own bits
    if (h == 0 || h == 1) return std::nextafter(v, 1e300);
    if (h == 2) return std::nextafter(v, -1e300);
    return v;
}
inline double o_log(double x) { return g_exact_libm ? o_jitter((double)logq((quad)x)) : std::log(x); }
inline double o_atan2(double y, double x) { return g_exact_libm ? o_jitter((double)atan2q((quad)y, (quad)x)) : std::atan2(y, x); }

struct Rec {  // one panel image
    const double *centr, *A, *vls, *nh, *b, *sb, *vg, *T;
    double J;
    int r;
    // higher-order tables
    int order = 1, M_dim = 3, S_dim = 1;
    const double *T6 = nullptr, *Ts = nullptr;
};
inline Rec get_rec(const ml_panel_soa* t, int j, int img) {
    size_t rec = (size_t)j + (size_t)img * t->n_panels;
    Rec R;
    R.centr = t->centr + 3 * rec;
    R.A = t->A_g_to_ls + 9 * rec;
    R.vls = t->vertices_ls + 6 * rec;
    R.nh = t->n_hat_ls + 6 * rec;
    R.b = t->b + 3 * rec;
    R.sb = t->sqrt_b + 3 * rec;
    R.vg = t->vert_g + 9 * rec;
    R.T = t->T_mu + 9 * rec;
    R.J = t->J[rec];
    R.r = t->r[rec];
    if (t->order2) {
        R.order = t->order[j];
        R.M_dim = t->M_dim[j];
        R.S_dim = t->S_dim[j];
        R.T6 = t->T_mu6 + 36 * rec;
        R.Ts = t->T_sigma + 12 * rec;
    }
    return R;
}

// flow.f90:282-310
inline bool point_in_dod(const ml_flow* fs, const double* Q, const double* P) {
    double d[3] = {P[0] - Q[0], P[1] - Q[1], P[2] - Q[2]};
    if (inner3(d, fs->c_hat_g) >= 0.) {
        double Cd[3];
        for (int i = 0; i < 3; ++i) Cd[i] = fs->C_mat_g[3 * i] * d[0] + fs->C_mat_g[3 * i + 1] * d[1] + fs->C_mat_g[3 * i + 2] * d[2];
        if (inner3(d, Cd) >= 0.) return true;
    }
    return false;
}

struct Dod {
    bool in_dod = true;
    bool e[3] = {true, true, true};
};

// panel.f90:1732-1901
Dod check_dod(const Rec& p, const double* P, const ml_flow* fs) {
    Dod dod;
    if (!fs->supersonic) return dod;
    bool vin[3];
    for (int i = 0; i < 3; ++i) vin[i] = point_in_dod(fs, p.vg + 3 * i, P);
    if (vin[0] && vin[1] && vin[2]) return dod;
    double dfv[3][3];
    for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 3; ++k) dfv[i][k] = P[k] - p.vg[3 * i + k];
    bool downstream = false;
    for (int i = 0; i < 3; ++i) {
        double x = inner3(dfv[i], fs->c_hat_g);
        downstream = x > 0. || downstream;
    }
    if (downstream) {
        for (int i = 0; i < 3; ++i) {
            int i_next = (i + 1) % 3;
            if (vin[i] || vin[i_next]) {
                dod.e[i] = true;
            } else if (p.b[i] <= 0.) {
                dod.e[i] = false;
            } else {
                const double* Q_end = p.vg + 3 * i_next;
                double d[3] = {Q_end[0] - p.vg[3 * i], Q_end[1] - p.vg[3 * i + 1], Q_end[2] - p.vg[3 * i + 2]};
                double a[3], b[3], nd[3] = {-dfv[i_next][0], -dfv[i_next][1], -dfv[i_next][2]};
                cross3(fs->c_hat_g, d, a);
                cross3(fs->c_hat_g, nd, b);
                double s_star = inner3(a, b) / std::fabs(inner3(a, a));
                double R_star[3] = {Q_end[0] - s_star * d[0], Q_end[1] - s_star * d[1], Q_end[2] - s_star * d[2]};
                if (s_star > 0. && s_star < 1.) dod.e[i] = point_in_dod(fs, R_star, P);
                else dod.e[i] = false;
            }
        }
        if (vin[0] || vin[1] || vin[2] || dod.e[0] || dod.e[1] || dod.e[2]) dod.in_dod = true;
        else dod.in_dod = false;  // (superinclined fallback, panel.f90:1859-1877, is dead: r > 0)
    } else {
        dod.e[0] = dod.e[1] = dod.e[2] = false;
        dod.in_dod = false;
    }
    return dod;
}

struct Geom {  // eval_point_geom, base_geom.f90:97-112
    double P_ls[2], h, h2, d_ls[3][2], a[3], l1[3], l2[3], R1[3], R2[3], dR[3], g2[3], v_xi[3], v_eta[3];
};

// panel.f90:1904-1938 + base_geom.f90:501-521
void basic_geom(const Rec& p, const double* P, Geom& g) {
    double d[3] = {P[0] - p.centr[0], P[1] - p.centr[1], P[2] - p.centr[2]};
    double Pls[3];
    for (int i = 0; i < 3; ++i) Pls[i] = p.A[3 * i] * d[0] + p.A[3 * i + 1] * d[1] + p.A[3 * i + 2] * d[2];
    g.P_ls[0] = Pls[0];
    g.P_ls[1] = Pls[1];
    g.h = Pls[2];
    g.h2 = g.h * g.h;
    for (int i = 0; i < 3; ++i) {
        g.R1[i] = g.R2[i] = g.a[i] = 0.;
        g.l1[i] = g.l2[i] = g.g2[i] = g.dR[i] = 0.;  // garbage in the reference; multiplied by zeros
        g.v_xi[i] = p.nh[2 * i];
        g.v_eta[i] = p.nh[2 * i + 1];
        g.d_ls[i][0] = p.vls[2 * i] - g.P_ls[0];
        g.d_ls[i][1] = p.vls[2 * i + 1] - g.P_ls[1];
    }
}

// panel.f90:1941-1997
void subsonic_geom(const Rec& p, const double* P, bool mirror, Geom& g) {
    basic_geom(p, P, g);
    for (int i = 0; i < 3; ++i) {
        int n = (i + 1) % 3;
        g.l1[i] = -g.d_ls[i][0] * g.v_eta[i] + g.d_ls[i][1] * g.v_xi[i];
        g.l2[i] = -g.d_ls[n][0] * g.v_eta[i] + g.d_ls[n][1] * g.v_xi[i];
    }
    for (int i = 0; i < 3; ++i) {
        g.a[i] = g.d_ls[i][0] * g.v_xi[i] + g.d_ls[i][1] * g.v_eta[i];
        g.g2[i] = g.a[i] * g.a[i] + g.h2;
        g.R1[i] = std::sqrt(g.d_ls[i][0] * g.d_ls[i][0] + g.d_ls[i][1] * g.d_ls[i][1] + g.h2);
    }
    for (int i = 0; i < 3; ++i) g.R2[i] = g.R1[(i + 1) % 3];  // cshift(R1, 1)
    if (mirror) {
        for (int i = 0; i < 3; ++i) {
            double t = g.l1[i];
            g.l1[i] = g.l2[i];
            g.l2[i] = t;
        }
        double R1o[3] = {g.R1[0], g.R1[1], g.R1[2]};
        for (int i = 0; i < 3; ++i) {
            g.R1[i] = g.R2[i];
            g.R2[i] = R1o[i];
        }
    }
    for (int i = 0; i < 3; ++i) g.dR[i] = g.R2[i] - g.R1[i];
}

// panel.f90:2000-2091
void supersonic_subinc_geom(const Rec& p, const double* P, bool mirror, const Dod& dod, Geom& g) {
    basic_geom(p, P, g);
    for (int i = 0; i < 3; ++i) {
        if (!dod.e[i]) continue;
        int n = (i + 1) % 3;
        g.l1[i] = g.v_eta[i] * g.d_ls[i][0] + g.v_xi[i] * g.d_ls[i][1];
        g.l2[i] = g.v_eta[i] * g.d_ls[n][0] + g.v_xi[i] * g.d_ls[n][1];
        g.a[i] = g.v_xi[i] * g.d_ls[i][0] + g.v_eta[i] * g.d_ls[i][1];
        g.g2[i] = g.a[i] * g.a[i] - p.b[i] * g.h2;
        double x = g.d_ls[i][0] * g.d_ls[i][0] - g.d_ls[i][1] * g.d_ls[i][1] - g.h2;
        if (x > 0. && g.d_ls[i][0] < 0.) {
            g.R1[i] = std::sqrt(x);
        } else {
            g.l1[i] = -std::sqrt(std::fabs(g.g2[i]));
            g.R1[i] = 0.;
        }
        x = g.d_ls[n][0] * g.d_ls[n][0] - g.d_ls[n][1] * g.d_ls[n][1] - g.h2;
        if (x > 0. && g.d_ls[n][0] < 0.) {
            g.R2[i] = std::sqrt(x);
        } else {
            g.l2[i] = std::sqrt(std::fabs(g.g2[i]));
            g.R2[i] = 0.;
        }
        if (mirror) {
            double dummy = g.l1[i];
            if (g.R2[i] == 0.) g.l1[i] = -g.l2[i];
            else g.l1[i] = g.l2[i];
            if (g.R1[i] == 0.) g.l2[i] = -dummy;
            else g.l2[i] = dummy;
            dummy = g.R1[i];
            g.R1[i] = g.R2[i];
            g.R2[i] = dummy;
        }
    }
    for (int i = 0; i < 3; ++i) g.dR[i] = g.R2[i] - g.R1[i];
}

struct Integrals {
    int r, s, rs;
    double H111, hH113, H213, H123;
    double F111[3];
    double F121[3] = {0., 0., 0.}, F211[3] = {0., 0., 0.};   // order 2 (panel.f90:2275-2279, 2325-2392)
    double H211 = 0., H121 = 0., H313 = 0., H223 = 0., H133 = 0.;
    // velocity integrals (panel.f90:2686-2763)
    double F113[3] = {0., 0., 0.}, F123[3] = {0., 0., 0.}, F133[3] = {0., 0., 0.};
    double h3H115 = 0., H125 = 0., hH135 = 0., H145 = 0., H215 = 0., H225 = 0., H235 = 0., hH315 = 0., H325 = 0., H415 = 0.,
           H113_3rsh2H115 = 0.;
    double hH113_abs = 0.;  // running-error scale of hH113: sum over edges of |term| + |cancelled products behind it|
                            // (not a reference quantity; tests only)
};

// panel.f90:2232-2283
void F_subsonic(const Geom& g, Integrals& I) {
    for (int i = 0; i < 3; ++i) {
        if (fsign(1., g.l1[i]) != fsign(1., g.l2[i])) {
            I.F111[i] = o_log(((g.R1[i] - g.l1[i]) * (g.R2[i] + g.l2[i])) / g.g2[i]);
        } else {
            I.F111[i] = fsign(1., g.l1[i]) * o_log((g.R2[i] + std::fabs(g.l2[i])) / (g.R1[i] + std::fabs(g.l1[i])));
        }
    }
    for (int i = 0; i < 3; ++i) {   // :2278-2279
        I.F121[i] = g.a[i] * g.v_eta[i] * I.F111[i] + g.v_xi[i] * g.dR[i];
        I.F211[i] = g.a[i] * g.v_xi[i] * I.F111[i] - g.v_eta[i] * g.dR[i];
    }
}

// panel.f90:2286-2407
void F_supersonic_subinc(const Rec& p, const Geom& g, const Dod& dod, bool mirror, Integrals& I) {
    for (int i = 0; i < 3; ++i) {
        if (!dod.e[i]) continue;
        const int i_next = (i + 1) % 3;
        double b = p.b[i], s_b = p.sb[i];
        if (g.R1[i] == 0. && g.R2[i] == 0.) {
            I.F111[i] = pi / s_b;
            I.F121[i] = -g.a[i] * g.v_eta[i] * I.F111[i] / b;
            I.F211[i] = g.a[i] * g.v_xi[i] * I.F111[i] / b;
        } else {
            double F1, F2;
            if (b > 0.) {
                F1 = (g.l1[i] * g.R2[i] - g.l2[i] * g.R1[i]) / g.g2[i];
                F2 = (b * g.R1[i] * g.R2[i] + g.l1[i] * g.l2[i]) / g.g2[i];
            } else {
                F1 = (g.R2[i] - g.R1[i]) * (g.R2[i] + g.R1[i]) / (g.l1[i] * g.R2[i] + g.l2[i] * g.R1[i]);
                F2 = (g.g2[i] - g.l1[i] * g.l1[i] - g.l2[i] * g.l2[i]) / (b * g.R1[i] * g.R2[i] - g.l1[i] * g.l2[i]);
            }
            if (std::fabs(F2) > 125.0 * std::fabs(s_b * F1)) {
                double eps = F1 / F2;
                double eps2 = eps * eps;
                double series = eps * eps2 * (1. / 3. - b * eps2 / 5. + (b * eps2) * (b * eps2) / 7.);
                I.F111[i] = -eps + b * series;
                // :2339-2351 (p.vls is the mirrored panel's own vertex array for a mirror image)
                const double eta_a = mirror ? p.vls[2 * i_next + 1] : p.vls[2 * i + 1];
                const double eta_b = mirror ? p.vls[2 * i + 1] : p.vls[2 * i_next + 1];
                I.F121[i] = (-g.v_xi[i] * g.dR[i] * g.R1[i] * g.R2[i] + g.l2[i] * g.R1[i] * (eta_a - g.P_ls[1]) -
                             g.l1[i] * g.R2[i] * (eta_b - g.P_ls[1])) /
                                (g.g2[i] * F2) -
                            g.a[i] * g.v_eta[i] * series;
                I.F211[i] = -g.v_eta[i] * g.dR[i] + g.a[i] * g.v_xi[i] * I.F111[i] - 2. * g.v_xi[i] * g.v_eta[i] * I.F121[i];
            } else if (b > 0.) {
                I.F111[i] = -o_atan2(s_b * F1, F2) / s_b;
                I.F121[i] = -(g.v_xi[i] * g.dR[i] + g.a[i] * g.v_eta[i] * I.F111[i]) / b;
                I.F211[i] = -g.v_eta[i] * g.dR[i] + g.a[i] * g.v_xi[i] * I.F111[i] - 2. * g.v_xi[i] * g.v_eta[i] * I.F121[i];
            } else {
                F1 = s_b * g.R1[i] + std::fabs(g.l1[i]);
                F2 = s_b * g.R2[i] + std::fabs(g.l2[i]);
                if (F1 != 0. && F2 != 0.) I.F111[i] = -fsign(1., g.v_eta[i]) * o_log(F1 / F2) / s_b;
                I.F121[i] = -(g.v_xi[i] * g.dR[i] + g.a[i] * g.v_eta[i] * I.F111[i]) / b;
                I.F211[i] = -g.v_eta[i] * g.dR[i] + g.a[i] * g.v_xi[i] * I.F111[i] - 2. * g.v_xi[i] * g.v_eta[i] * I.F121[i];
            }
        }
    }
}

// panel.f90:2472-2509
void hH113_subsonic(const Geom& g, Integrals& I) {
    I.hH113 = 0.;
    for (int i = 0; i < 3; ++i) {
        double c1 = g.g2[i] + std::fabs(g.h) * g.R1[i];
        double c2 = g.g2[i] + std::fabs(g.h) * g.R2[i];
        double S = g.a[i] * (g.l2[i] * c1 - g.l1[i] * c2);
        double C = c1 * c2 + g.a[i] * g.a[i] * g.l1[i] * g.l2[i];
        double x = o_atan2(S, C);
        I.hH113 = I.hH113 + x;
        // scale of the term's rounding noise: |x| plus what the cancellations inside S and C feed into the angle,
        // d(theta) = (C dS - S dC) / (S^2 + C^2) with dS, dC ~ the summed |products| of S and C
        const double S_abs = std::fabs(g.a[i]) * (std::fabs(g.l2[i] * c1) + std::fabs(g.l1[i] * c2));
        const double C_abs = std::fabs(c1 * c2) + std::fabs(g.a[i] * g.a[i] * g.l1[i] * g.l2[i]);
        const double den = S * S + C * C;
        I.hH113_abs += std::fabs(x) + (den > 0. ? (std::fabs(C) * S_abs + std::fabs(S) * C_abs) / den : 0.);
    }
    I.hH113 = fsign(I.hH113, g.h);
}

// panel.f90:2512-2573 -- F1, F2, b are real(16); operands of pure-binary64 subexpressions are
// evaluated in binary64 first and then widened, as Fortran's mixed-mode rules prescribe.
void hH113_supersonic_subinc(const Rec& p, const Geom& g, const Dod& dod, Integrals& I) {
    I.hH113 = 0.;
    for (int i = 0; i < 3; ++i) {
        if (!dod.e[i]) continue;
        quad b = (quad)p.b[i];
        if (std::fabs(g.h) > 1.e-12) {
            if (g.R1[i] == 0. && g.R2[i] == 0.) {
                I.hH113 = I.hH113 + pi * fsign(1., g.h * g.v_xi[i]);
                I.hH113_abs += pi;
            } else {
                quad F1, F2;
                if (b > 0) {
                    F1 = (quad)((g.l1[i] * g.R2[i] - g.l2[i] * g.R1[i]) / g.g2[i]);
                    F2 = (b * (quad)g.R1[i] * (quad)g.R2[i] + (quad)(g.l1[i] * g.l2[i])) / (quad)g.g2[i];
                } else {
                    F1 = (quad)(g.dR[i] * (g.R2[i] + g.R1[i]) / (g.l1[i] * g.R2[i] + g.l2[i] * g.R1[i]));
                    F2 = (quad)(g.g2[i] - g.l1[i] * g.l1[i] - g.l2[i] * g.l2[i]) /
                         (b * (quad)g.R1[i] * (quad)g.R2[i] - (quad)(g.l1[i] * g.l2[i]));
                }
                quad y = (quad)(g.h * g.a[i]) * F1;
                quad x = (quad)(g.R1[i] * g.R2[i]) + (quad)g.h2 * F2;
                quad t = atan2q(y, x);
                I.hH113 = (double)((quad)I.hH113 + t);
                I.hH113_abs += std::fabs((double)t);
            }
        }
    }
}

// panel.f90:2181-2229: E(M,N,K) of every edge; integer powers as Fortran evaluates them (x**0 = 1, x**1 = x, x**2 = x*x)
inline double ipow(double x, int n) { return n == 0 ? 1. : (n == 1 ? x : x * x); }
void EMNK(const Geom& g, int M, int N, int K, bool mirror, double E[3]) {
    for (int i = 0; i < 3; ++i) {
        const int n = (i + 1) % 3;
        double E1, E2;
        if (g.R1[i] == 0.) E1 = 0.0;
        else if (mirror) E1 = ipow(g.d_ls[n][0], M - 1) * ipow(g.d_ls[n][1], N - 1) / ipow(g.R1[i], K);
        else E1 = ipow(g.d_ls[i][0], M - 1) * ipow(g.d_ls[i][1], N - 1) / ipow(g.R1[i], K);
        if (g.R2[i] == 0.) E2 = 0.0;
        else if (mirror) E2 = ipow(g.d_ls[i][0], M - 1) * ipow(g.d_ls[i][1], N - 1) / ipow(g.R2[i], K);
        else E2 = ipow(g.d_ls[n][0], M - 1) * ipow(g.d_ls[n][1], N - 1) / ipow(g.R2[i], K);
        E[i] = E2 - E1;
    }
}

// panel.f90:2686-2763: the F and H integrals of the velocity influences (the reference notes that its H recursions are
// "ONLY SUBSONIC RIGHT NOW", :2740; restated as they are)
void velocity_recursions(const Geom& g, const Dod& dod, bool mirror, Integrals& I) {
    double E211[3], E121[3], E131[3], E221[3];
    EMNK(g, 2, 1, 1, mirror, E211);
    EMNK(g, 1, 2, 1, mirror, E121);
    EMNK(g, 1, 3, 1, mirror, E131);
    EMNK(g, 2, 2, 1, mirror, E221);
    const double s = I.s, r = I.r, rs = I.rs;
    for (int i = 0; i < 3; ++i) {
        if (!dod.e[i]) {
            I.F113[i] = I.F123[i] = I.F133[i] = 0.;
            continue;
        }
        const double vx = g.v_xi[i], ve = g.v_eta[i];
        I.F113[i] = (-s * ve * E211[i] + r * vx * E121[i]) / g.g2[i];
        I.F123[i] = (-r * (vx * vx) * I.F121[i] + r * vx * E131[i] - s * ve * E221[i] + s * ve * vx * I.F211[i]) / g.g2[i];
        I.F133[i] = (r * ve * g.a[i] * I.F123[i] + (vx * vx) * I.F111[i] - vx * E121[i]) / (s * (vx * vx) + r * (ve * ve));
    }
    auto sum3 = [](const double* u, const double* v) { return (u[0] * v[0] + u[1] * v[1]) + u[2] * v[2]; };
    I.h3H115 = rs * (I.hH113 + g.h * sum3(g.a, I.F113)) / 3;
    I.H125 = -s * sum3(g.v_eta, I.F113) / 3.;
    I.hH135 = s * (I.hH113 - g.h * sum3(g.v_eta, I.F123)) / 3.;
    I.H145 = s * (2. * I.H123 - sum3(g.v_eta, I.F133)) / 3.;
    I.H215 = -r * sum3(g.v_xi, I.F113) / 3.;
    I.H225 = -r * sum3(g.v_xi, I.F123) / 3.;
    I.H235 = -r * sum3(g.v_xi, I.F133) / 3.;
    I.hH315 = -rs * I.hH135 - s * I.h3H115 + r * I.hH113;
    I.H325 = -rs * I.H145 - s * g.h2 * I.H125 + r * I.H123;
    I.H415 = -rs * I.H235 - s * g.h2 * I.H215 + r * I.H223;   // sic (:2760): R^2 = xi^2 + eta^2 + h^2 gives H213 here; see tests/test_oracle_identities.py
    I.H113_3rsh2H115 = -sum3(g.a, I.F113);
}

// panel.f90:2766-2812 + 2631-2647
void calc_integrals(const Rec& p, const Geom& g, const ml_flow* fs, const Dod& dod, bool mirror, bool has_sources, Integrals& I) {
    I.F111[0] = I.F111[1] = I.F111[2] = 0.;
    I.r = p.r;
    I.s = (int)fs->s;
    I.rs = I.r * I.s;
    if (fs->supersonic) {
        F_supersonic_subinc(p, g, dod, mirror, I);
        hH113_supersonic_subinc(p, g, dod, I);
    } else {
        F_subsonic(g, I);
        hH113_subsonic(g, I);
    }
    double s1 = 0., s2 = 0., s3 = 0.;
    for (int i = 0; i < 3; ++i) s1 = s1 + g.a[i] * I.F111[i];
    for (int i = 0; i < 3; ++i) s2 = s2 + g.v_xi[i] * I.F111[i];
    for (int i = 0; i < 3; ++i) s3 = s3 + g.v_eta[i] * I.F111[i];
    I.H111 = s1 - I.rs * g.h * I.hH113;
    I.H213 = -I.r * s2;
    I.H123 = -I.s * s3;
    if (p.order == 2) {   // panel.f90:2649-2662
        if (has_sources) {
            double sa211 = 0., sa121 = 0.;
            for (int i = 0; i < 3; ++i) sa211 = sa211 + g.a[i] * I.F211[i];
            for (int i = 0; i < 3; ++i) sa121 = sa121 + g.a[i] * I.F121[i];
            I.H211 = 0.5 * (-I.rs * g.h2 * I.H213 + sa211);
            I.H121 = 0.5 * (-I.rs * g.h2 * I.H123 + sa121);
        }
        double sx211 = 0., sx121 = 0., se121 = 0.;
        for (int i = 0; i < 3; ++i) sx211 = sx211 + g.v_xi[i] * I.F211[i];
        for (int i = 0; i < 3; ++i) sx121 = sx121 + g.v_xi[i] * I.F121[i];
        for (int i = 0; i < 3; ++i) se121 = se121 + g.v_eta[i] * I.F121[i];
        I.H313 = I.r * (I.H111 - sx211);
        I.H223 = -I.r * sx121;
        I.H133 = I.s * (I.H111 - se121);
    }
}

}  // namespace

// panel.f90:2917-2971.  phi_d has 3 entries (the wake's negated copy, :2909-2912, is applied by the caller).
extern "C" void orc_set_exact_libm(int on) { g_exact_libm = on; }

extern "C" void orc_pair_influence(const ml_flow* fs, const ml_panel_soa* t, int j, int img, const double* P,
                                   orc_pair_out* out) {
    std::memset(out, 0, sizeof *out);
    Rec p = get_rec(t, j, img);
    bool mirror = img == 1;
    Dod dod = check_dod(p, P, fs);
    out->in_dod = dod.in_dod;
    for (int i = 0; i < 3; ++i) out->edges_in_dod[i] = dod.e[i];
    if (!(dod.in_dod && t->area[j] > 0.)) return;
    Geom g;
    if (fs->supersonic) supersonic_subinc_geom(p, P, mirror, dod, g);
    else subsonic_geom(p, P, mirror, g);
    bool has_src = !t->in_wake && t->has_sources && t->has_sources[j];
    Integrals I;
    calc_integrals(p, g, fs, dod, mirror, has_src, I);
    for (int i = 0; i < 3; ++i) out->F111[i] = I.F111[i];
    out->hH113 = I.hH113;
    out->H111 = I.H111;
    out->H213 = I.H213;
    out->H123 = I.H123;
    out->h = g.h;
    // assemble_phi_s_S_space (order 1), panel.f90:2852-2859
    out->phi_s = has_src ? -p.J * fs->K_inv * I.H111 : 0.;
    // assemble_phi_d_M_space, panel.f90:2888-2907
    double m[3];
    m[0] = I.hH113;
    m[1] = I.hH113 * g.P_ls[0] + g.h * I.H213;
    m[2] = I.hH113 * g.P_ls[1] + g.h * I.H123;
    for (int c = 0; c < 3; ++c) {
        double acc = 0.;
        for (int k = 0; k < 3; ++k) acc = acc + m[k] * p.T[3 * k + c];
        out->phi_d[c] = I.s * fs->K_inv * acc;
    }
    for (int i = 0; i < 3; ++i) {
        out->F121[i] = I.F121[i];
        out->F211[i] = I.F211[i];
    }
    out->H211 = I.H211;
    out->H121 = I.H121;
    out->H313 = I.H313;
    out->H223 = I.H223;
    out->H133 = I.H133;
    if (t->order2) {
        // assemble_phi_s_S_space, panel.f90:2815-2863
        if (has_src) {
            if (p.order == 2) {
                const double sg[3] = {I.H111, I.H111 * g.P_ls[0] + I.H211, I.H111 * g.P_ls[1] + I.H121};
                for (int c = 0; c < p.S_dim; ++c) {
                    double acc = 0.;
                    for (int k = 0; k < 3; ++k) acc = acc + sg[k] * p.Ts[4 * k + c];
                    out->phi_s_S[c] = -p.J * fs->K_inv * acc;
                }
            } else {
                out->phi_s_S[0] = out->phi_s;
            }
        }
        // assemble_phi_d_M_space, panel.f90:2866-2914
        double mu6[6] = {m[0], m[1], m[2], 0., 0., 0.};
        double mu6a[6] = {0., 0., 0., 0., 0., 0.};
        const int md = p.order == 2 ? 6 : 3;
        if (p.order == 2) {
            mu6[3] = 0.5 * I.hH113 * (g.P_ls[0] * g.P_ls[0]) + g.h * (g.P_ls[0] * I.H213 + 0.5 * I.H313);
            mu6[4] = I.hH113 * g.P_ls[0] * g.P_ls[1] + g.h * (g.P_ls[1] * I.H213 + g.P_ls[0] * I.H123 + I.H223);
            mu6[5] = 0.5 * I.hH113 * (g.P_ls[1] * g.P_ls[1]) + g.h * (g.P_ls[1] * I.H123 + 0.5 * I.H133);
        }
        {   // forward-error scale of the six parameters-space influences (tests only)
            double s2a = 0., s3a = 0.;
            for (int i = 0; i < 3; ++i) s2a += std::fabs(g.v_xi[i] * I.F111[i]);
            for (int i = 0; i < 3; ++i) s3a += std::fabs(g.v_eta[i] * I.F111[i]);
            const double x = std::fabs(g.P_ls[0]), y = std::fabs(g.P_ls[1]), ah = std::fabs(g.h);
            double sx211 = 0., sx121 = 0., se121 = 0., h111a = 0.;
            for (int i = 0; i < 3; ++i) sx211 += std::fabs(g.v_xi[i] * I.F211[i]);
            for (int i = 0; i < 3; ++i) sx121 += std::fabs(g.v_xi[i] * I.F121[i]);
            for (int i = 0; i < 3; ++i) se121 += std::fabs(g.v_eta[i] * I.F121[i]);
            for (int i = 0; i < 3; ++i) h111a += std::fabs(g.a[i] * I.F111[i]);
            h111a += ah * I.hH113_abs;
            mu6a[0] = I.hH113_abs;
            mu6a[1] = I.hH113_abs * x + ah * s2a;
            mu6a[2] = I.hH113_abs * y + ah * s3a;
            mu6a[3] = 0.5 * I.hH113_abs * x * x + ah * (x * s2a + 0.5 * (h111a + sx211));
            mu6a[4] = I.hH113_abs * x * y + ah * (y * s2a + x * s3a + sx121);
            mu6a[5] = 0.5 * I.hH113_abs * y * y + ah * (y * s3a + 0.5 * (h111a + se121));
        }
        for (int c = 0; c < p.M_dim; ++c) {
            double acc = 0., acc_a = 0.;
            for (int k = 0; k < md; ++k) acc = acc + mu6[k] * p.T6[6 * k + c];
            for (int k = 0; k < md; ++k) acc_a += mu6a[k] * std::fabs(p.T6[6 * k + c]);
            out->phi_d_M[c] = I.s * fs->K_inv * acc;
            out->phi_d_M_abs[c] = fs->K_inv * acc_a;
        }
    }
    // assemble_v_s_S_space (order 1), panel.f90:3029-3072: local (r H213, s H123, -rs hH113), times -K_inv J, to global
    {
        double vl[3] = {I.r * I.H213, I.s * I.H123, -I.rs * I.hH113};
        for (int k = 0; k < 3; ++k) vl[k] = has_src ? -vl[k] * fs->K_inv * p.J : 0.;
        for (int i = 0; i < 3; ++i) {   // matmul(transpose(A_g_to_ls), v)
            double acc = 0.;
            for (int k = 0; k < 3; ++k) acc = acc + p.A[3 * k + i] * vl[k];
            out->v_s[i] = acc;
        }
    }
    // assemble_v_d_M_space (order 1), panel.f90:3118-3166
    {
        double vmu[9] = {0., I.hH113, 0., 0., 0., I.hH113, 0., I.H213, I.H123};   // row-major 3 x mu_dim
        double vM[9];
        for (int i = 0; i < 3; ++i)
            for (int c = 0; c < 3; ++c) {
                double acc = 0.;
                for (int k = 0; k < 3; ++k) acc = acc + vmu[3 * i + k] * p.T[3 * k + c];
                vM[3 * i + c] = I.s * fs->K_inv * acc;
            }
        for (int i = 0; i < 3; ++i)
            for (int c = 0; c < 3; ++c) {
                double acc = 0.;
                for (int k = 0; k < 3; ++k) acc = acc + p.A[3 * k + i] * vM[3 * k + c];
                out->v_d[3 * i + c] = acc;
            }
    }
    // the same for a higher-order table (panel.f90:3011-3170): S_dim / M_dim columns
    if (t->order2) {
        for (int i = 0; i < 3; ++i) {
            out->F113[i] = 0.;
            out->F123[i] = 0.;
            out->F133[i] = 0.;
        }
        if (p.order == 2) {
            velocity_recursions(g, dod, mirror, I);
            for (int i = 0; i < 3; ++i) {
                out->F113[i] = I.F113[i];
                out->F123[i] = I.F123[i];
                out->F133[i] = I.F133[i];
            }
            out->h3H115 = I.h3H115; out->H125 = I.H125; out->hH135 = I.hH135; out->H145 = I.H145; out->H215 = I.H215;
            out->H225 = I.H225; out->H235 = I.H235; out->hH315 = I.hH315; out->H325 = I.H325; out->H415 = I.H415;
            out->H113_3rsh2H115 = I.H113_3rsh2H115;
            const double r = I.r, sg = I.s, rs = I.rs, x = g.P_ls[0], y = g.P_ls[1], h = g.h, h2 = g.h2;
            // assemble_v_s_S_space, panel.f90:3029-3072
            if (has_src) {
                double vs[3][3];   // [component][sigma parameter]
                vs[0][0] = r * I.H213;
                vs[1][0] = sg * I.H123;
                vs[2][0] = -rs * I.hH113;
                vs[0][1] = r * (I.H213 * x + I.H313);
                vs[0][2] = r * (I.H213 * y + I.H223);
                vs[1][1] = sg * (I.H123 * x + I.H223);
                vs[1][2] = sg * (I.H123 * y + I.H133);
                vs[2][1] = -rs * (I.hH113 * x + h * I.H213);
                vs[2][2] = -rs * (I.hH113 * y + h * I.H123);
                double vS[3][4] = {};
                for (int i = 0; i < 3; ++i)
                    for (int c = 0; c < p.S_dim; ++c) {
                        double acc = 0.;
                        for (int k = 0; k < 3; ++k) acc = acc + vs[i][k] * p.Ts[4 * k + c];
                        vS[i][c] = -acc * fs->K_inv * p.J;
                    }
                for (int i = 0; i < 3; ++i)
                    for (int c = 0; c < p.S_dim; ++c) {
                        double acc = 0.;
                        for (int k = 0; k < 3; ++k) acc = acc + p.A[3 * k + i] * vS[k][c];
                        out->v_s_S[4 * i + c] = acc;
                    }
            }
            // assemble_v_d_M_space, panel.f90:3118-3166
            double vd[3][6] = {};
            vd[0][1] = I.hH113;
            vd[1][2] = I.hH113;
            vd[2][1] = I.H213;
            vd[2][2] = I.H123;
            vd[0][3] = 3. * r * (0.5 * I.H215 * (x * x) * h + I.hH315 * x + 0.5 * I.H415 * h);
            vd[0][4] = 3. * r * (I.H215 * h * x * y + I.hH315 * y + I.H225 * h * x + I.H325 * h);
            vd[0][5] = 3. * r * (0.5 * I.H215 * (y * y) * h + I.H225 * h * y + 0.5 * I.H235 * h);
            vd[1][3] = 3. * sg * (0.5 * I.H125 * (x * x) * h + I.H225 * h * x + 0.5 * I.H325 * h);
            vd[1][4] = 3. * sg * (I.H125 * h * x * y + I.hH135 * x + I.H225 * h * y + I.H235 * h);
            vd[1][5] = 3. * sg * (0.5 * I.H125 * (y * y) * h + I.hH135 * y + 0.5 * I.H145 * h);
            vd[2][3] = 0.5 * (x * x) * I.H113_3rsh2H115 + x * (I.H213 - 3. * rs * h2 * I.H215) + 0.5 * (I.H313 - 3. * rs * h * I.hH315);
            vd[2][4] = x * y * (I.H113_3rsh2H115) + y * (I.H213 - 3. * rs * h2 * I.H215) + x * (I.H123 - 3. * rs * h2 * I.H125) + I.H223 -
                       3. * rs * h2 * I.H225;
            vd[2][5] = 0.5 * (y * y) * I.H113_3rsh2H115 + y * (I.H123 - 3. * rs * h2 * I.H125) + 0.5 * (I.H133 - 3. * rs * h * I.hH135);
            double vM[3][6] = {};
            for (int i = 0; i < 3; ++i)
                for (int c = 0; c < p.M_dim; ++c) {
                    double acc = 0.;
                    for (int k = 0; k < 6; ++k) acc = acc + vd[i][k] * p.T6[6 * k + c];
                    vM[i][c] = I.s * fs->K_inv * acc;
                }
            for (int i = 0; i < 3; ++i)
                for (int c = 0; c < p.M_dim; ++c) {
                    double acc = 0.;
                    for (int k = 0; k < 3; ++k) acc = acc + p.A[3 * k + i] * vM[k][c];
                    out->v_d_M[6 * i + c] = acc;
                }
        } else {   // an order-1 panel of the table
            for (int i = 0; i < 3; ++i) {
                out->v_s_S[4 * i] = out->v_s[i];
                for (int c = 0; c < 3; ++c) out->v_d_M[6 * i + c] = 0.;
            }
            // its T_mu lives in the upper left corner of T6 (p.T is filled as well for order 1)
            for (int i = 0; i < 3; ++i)
                for (int c = 0; c < 3; ++c) out->v_d_M[6 * i + c] = out->v_d[3 * i + c];
        }
    }
    // magnitude of the terms that were summed into phi_d[c] (forward-error scale, tests only)
    double ma[3];
    double s2a = 0., s3a = 0.;
    for (int i = 0; i < 3; ++i) s2a += std::fabs(g.v_xi[i] * I.F111[i]);
    for (int i = 0; i < 3; ++i) s3a += std::fabs(g.v_eta[i] * I.F111[i]);
    ma[0] = I.hH113_abs;
    ma[1] = I.hH113_abs * std::fabs(g.P_ls[0]) + std::fabs(g.h) * s2a;
    ma[2] = I.hH113_abs * std::fabs(g.P_ls[1]) + std::fabs(g.h) * s3a;
    for (int c = 0; c < 3; ++c) {
        double acc = 0.;
        for (int k = 0; k < 3; ++k) acc += ma[k] * std::fabs(p.T[3 * k + c]);
        out->phi_d_abs[c] = fs->K_inv * acc;
    }
}

// Batch of pair evaluations (tests only): records are (panel j, image img) with index r = j + img * n_panels.
// phi_d[n_pts][n_rec][3], phi_d_abs likewise, phi_s[n_pts][n_rec], in_dod[n_pts][n_rec].
extern "C" void orc_pair_batch(const ml_flow* fs, const ml_panel_soa* t, int n_pts, const double* pts, double* phi_d,
                               double* phi_d_abs, double* phi_s, unsigned char* in_dod) {
    const int n_rec = t->n_panels * t->n_images;
#pragma omp parallel for schedule(static)
    for (int p = 0; p < n_pts; ++p) {
        for (int r = 0; r < n_rec; ++r) {
            orc_pair_out o;
            orc_pair_influence(fs, t, r % t->n_panels, r / t->n_panels, pts + 3 * (size_t)p, &o);
            const size_t k = (size_t)p * n_rec + r;
            const bool on = o.in_dod && t->area[r % t->n_panels] > 0.;
            in_dod[k] = on;
            phi_s[k] = on ? o.phi_s : 0.;
            for (int c = 0; c < 3; ++c) {
                phi_d[3 * k + c] = on ? o.phi_d[c] : 0.;
                phi_d_abs[3 * k + c] = on ? o.phi_d_abs[c] : 0.;
            }
        }
    }
}

// The same for a higher-order table: phi_d6[n_pts][n_rec][6] (the panel's M_dim strength-space influences, zero padded),
// phi_d6_abs likewise, phi_s[n_pts][n_rec] = the sum of the panel's S_dim source influences.
extern "C" void orc_pair_batch_ho(const ml_flow* fs, const ml_panel_soa* t, int n_pts, const double* pts, double* phi_d6,
                                  double* phi_d6_abs, double* phi_s, unsigned char* in_dod) {
    const int n_rec = t->n_panels * t->n_images;
#pragma omp parallel for schedule(static)
    for (int p = 0; p < n_pts; ++p) {
        for (int r = 0; r < n_rec; ++r) {
            orc_pair_out o;
            const int j = r % t->n_panels;
            orc_pair_influence(fs, t, j, r / t->n_panels, pts + 3 * (size_t)p, &o);
            const size_t k = (size_t)p * n_rec + r;
            const bool on = o.in_dod && t->area[j] > 0.;
            in_dod[k] = on;
            double s = 0.;
            for (int c = 0; c < t->S_dim[j]; ++c) s += o.phi_s_S[c];
            phi_s[k] = on ? s : 0.;
            for (int c = 0; c < 6; ++c) {
                phi_d6[6 * k + c] = (on && c < t->M_dim[j]) ? o.phi_d_M[c] : 0.;
                phi_d6_abs[6 * k + c] = (on && c < t->M_dim[j]) ? o.phi_d_M_abs[c] : 0.;
            }
        }
    }
}

// panel_solver.f90:1290-1501 + :1504-1706, Dirichlet / strength-matching rows.
// Rows row0..row0+nrows-1 of the permuted system are computed; A is column-major nrows x n_unknown with
// leading dimension ld, output row index = row - row0 (likewise I_known and A_abs).
extern "C" int orc_assemble(const ml_flow* fs, const ml_panel_soa* body, const ml_panel_soa* wake,
                            const ml_system_map* map, int n_cp, const double* cp_loc, const int* cp_bc,
                            const int* row_perm, int row0, int nrows, double* A, int ld, double* I_known,
                            int n_threads, double* A_abs) {
    return orc_assemble_n(fs, body, wake, map, n_cp, cp_loc, cp_bc, nullptr, row_perm, row0, nrows, A, ld, I_known, n_threads, A_abs);
}

// source_inf = matmul(n_g, matmul(B_mat_g, v_s)) / matmul(n_g, v_s)  (panel_solver.f90:1335-1336, 1409-1410)
static double project_velocity(const ml_flow* fs, const double* n_g, const double* v, int stride, bool mass_flux) {
    double w[3];
    for (int i = 0; i < 3; ++i) {
        if (mass_flux) {
            double acc = 0.;
            for (int k = 0; k < 3; ++k) acc = acc + fs->B_mat_g[3 * i + k] * v[k * stride];
            w[i] = acc;
        } else {
            w[i] = v[i * stride];
        }
    }
    double acc = 0.;
    for (int i = 0; i < 3; ++i) acc = acc + n_g[i] * w[i];
    return acc;
}

extern "C" int orc_assemble_n(const ml_flow* fs, const ml_panel_soa* body, const ml_panel_soa* wake,
                              const ml_system_map* map, int n_cp, const double* cp_loc, const int* cp_bc, const double* cp_n_g,
                              const int* row_perm, int row0, int nrows, double* A, int ld, double* I_known,
                              int n_threads, double* A_abs) {
    const int N_unknown = map->n_unknown, N_panels = map->n_body_panels, N_verts = map->n_verts;
    const int* P = map->P;
    if (n_threads <= 0) n_threads = omp_get_max_threads();
    int status = 0;
#pragma omp parallel for schedule(dynamic) num_threads(n_threads)
    for (int i = 0; i < n_cp; ++i) {
        int row = row_perm[i];
        if (row < row0 || row >= row0 + nrows) continue;
        std::vector<double> A_i(N_unknown, 0.);
        // sum of |contributions| per entry: the scale rounding differences must be judged against
        std::vector<double> S_i(A_abs ? N_unknown : 0, 0.);
        double I_known_i = 0.;
        const double* Pt = cp_loc + 3 * (size_t)i;
        if (cp_bc[i] == ML_BC_STRENGTH_MATCHING) {
            A_i[P[i]] = 1.;
            A_i[P[i - n_cp / 2]] = -1.;
        } else if (cp_bc[i] == ML_BC_ZERO_POTENTIAL || cp_bc[i] == ML_BC_SF_POTENTIAL) {
            for (int j = 0; j < N_panels; ++j) {
                for (int img = 0; img < body->n_images; ++img) {
                    orc_pair_out o;
                    orc_pair_influence(fs, body, j, img, Pt, &o);
                    if (!o.in_dod) continue;  // panel_solver.f90:1448 / 1462
                    bool mirrored_panel = (img == 1) && map->asym_flow;  // :1470-1471
                    if (body->order2) {
                        // update_system_row, panel_solver.f90:1203-1287, with the panel's own S_dim / M_dim
                        if (body->has_sources[j]) {
                            for (int k = 0; k < body->S_dim[j]; ++k) {
                                int ips = body->i_panel_s4[(size_t)j * 4 + k];
                                int index;
                                if (mirrored_panel) index = (ips >= N_panels) ? ips - N_panels : ips + N_panels;
                                else index = (ips >= N_panels) ? ips - N_panels : ips;
                                if (map->sigma_known[index]) I_known_i = I_known_i + o.phi_s_S[k] * map->sigma[index];
                                else A_i[P[map->i_sigma_in_sys[index]]] += o.phi_s_S[k];
                            }
                        }
                        for (int k = 0; k < body->M_dim[j]; ++k) {
                            int iv = body->i_vert_d[(size_t)j * body->n_cols + k];
                            int index;
                            if (mirrored_panel) index = (iv >= N_verts) ? iv - N_verts : iv + N_verts;
                            else index = (iv >= N_verts) ? iv - N_verts : iv;
                            A_i[P[index]] = A_i[P[index]] + o.phi_d_M[k];
                            if (A_abs) S_i[P[index]] += o.phi_d_M_abs[k];
                        }
                        continue;
                    }
                    // update_system_row, panel_solver.f90:1203-1287 (S_dim = 1, M_dim = 3)
                    if (body->has_sources[j]) {
                        int ips = body->i_panel_s[j];
                        int index;
                        if (mirrored_panel) index = (ips >= N_panels) ? ips - N_panels : ips + N_panels;
                        else index = (ips >= N_panels) ? ips - N_panels : ips;
                        if (map->sigma_known[index]) I_known_i = I_known_i + o.phi_s * map->sigma[index];
                        else A_i[P[map->i_sigma_in_sys[index]]] += o.phi_s;
                    }
                    for (int k = 0; k < 3; ++k) {
                        int iv = body->i_vert_d[(size_t)j * body->n_cols + k];
                        int index;
                        if (mirrored_panel) index = (iv >= N_verts) ? iv - N_verts : iv + N_verts;
                        else index = (iv >= N_verts) ? iv - N_verts : iv;
                        A_i[P[index]] = A_i[P[index]] + o.phi_d[k];
                        if (A_abs) S_i[P[index]] += o.phi_d_abs[k];
                    }
                }
            }
        } else if ((cp_bc[i] == ML_BC_ZERO_NORMAL_MF || cp_bc[i] == ML_BC_ZERO_NORMAL_VEL) && cp_n_g) {
            // Neumann rows, panel_solver.f90:1322-1440: normal mass flux n . B v or normal velocity n . v
            const bool mf = cp_bc[i] == ML_BC_ZERO_NORMAL_MF;
            const double* n_g = cp_n_g + 3 * (size_t)i;
            for (int j = 0; j < N_panels; ++j) {
                for (int img = 0; img < body->n_images; ++img) {
                    orc_pair_out o;
                    orc_pair_influence(fs, body, j, img, Pt, &o);
                    if (!o.in_dod) continue;
                    bool mirrored_panel = (img == 1) && map->asym_flow;
                    if (body->order2) {   // S_dim / M_dim columns of a higher-order table
                        if (body->has_sources[j]) {
                            for (int k = 0; k < body->S_dim[j]; ++k) {
                                const double source_inf = project_velocity(fs, n_g, o.v_s_S + k, 4, mf);
                                int ips = body->i_panel_s4[(size_t)j * 4 + k];
                                int index;
                                if (mirrored_panel) index = (ips >= N_panels) ? ips - N_panels : ips + N_panels;
                                else index = (ips >= N_panels) ? ips - N_panels : ips;
                                if (map->sigma_known[index]) I_known_i = I_known_i + source_inf * map->sigma[index];
                                else A_i[P[map->i_sigma_in_sys[index]]] += source_inf;
                            }
                        }
                        for (int k = 0; k < body->M_dim[j]; ++k) {
                            const double doublet_inf = project_velocity(fs, n_g, o.v_d_M + k, 6, mf);
                            int iv = body->i_vert_d[(size_t)j * body->n_cols + k];
                            int index;
                            if (mirrored_panel) index = (iv >= N_verts) ? iv - N_verts : iv + N_verts;
                            else index = (iv >= N_verts) ? iv - N_verts : iv;
                            A_i[P[index]] = A_i[P[index]] + doublet_inf;
                            if (A_abs) S_i[P[index]] += std::fabs(doublet_inf);
                        }
                        continue;
                    }
                    if (body->has_sources[j]) {
                        const double source_inf = project_velocity(fs, n_g, o.v_s, 1, mf);
                        int ips = body->i_panel_s[j];
                        int index;
                        if (mirrored_panel) index = (ips >= N_panels) ? ips - N_panels : ips + N_panels;
                        else index = (ips >= N_panels) ? ips - N_panels : ips;
                        if (map->sigma_known[index]) I_known_i = I_known_i + source_inf * map->sigma[index];
                        else A_i[P[map->i_sigma_in_sys[index]]] += source_inf;
                    }
                    for (int k = 0; k < 3; ++k) {
                        const double doublet_inf = project_velocity(fs, n_g, o.v_d + k, 3, mf);
                        int iv = body->i_vert_d[(size_t)j * body->n_cols + k];
                        int index;
                        if (mirrored_panel) index = (iv >= N_verts) ? iv - N_verts : iv + N_verts;
                        else index = (iv >= N_verts) ? iv - N_verts : iv;
                        A_i[P[index]] = A_i[P[index]] + doublet_inf;
                        if (A_abs) S_i[P[index]] += std::fabs(doublet_inf);
                    }
                }
            }
        } else {
#pragma omp critical
            status = ML_UNSUPPORTED;
        }
        // wake pass (panel_solver.f90:1650-1697): a separate row accumulated from zero, then added
        if (wake && wake->n_panels > 0 && cp_bc[i] != ML_BC_STRENGTH_MATCHING) {
            const bool neumann = (cp_bc[i] == ML_BC_ZERO_NORMAL_MF || cp_bc[i] == ML_BC_ZERO_NORMAL_VEL) && cp_n_g;
            std::vector<double> W_i(N_unknown, 0.);
            for (int l = 0; l < wake->n_panels; ++l) {
                for (int img = 0; img < wake->n_images; ++img) {
                    if (img == 1 && !(wake->image_present && wake->image_present[l])) continue;
                    orc_pair_out o;
                    orc_pair_influence(fs, wake, l, img, Pt, &o);
                    if (neumann) {   // panel_solver.f90:1530-1566, 1612-1650: the doublet velocity influences, projected
                        for (int c = 0; c < 3; ++c)
                            o.phi_d[c] = o.in_dod ? project_velocity(fs, cp_n_g + 3 * (size_t)i, o.v_d + c, 3, cp_bc[i] == ML_BC_ZERO_NORMAL_MF) : 0.;
                        for (int c = 0; c < 3; ++c) o.phi_d_abs[c] = std::fabs(o.phi_d[c]);
                    }
                    // in_dod = false leaves zeros (panel.f90:2959-2967), which are still "added"
                    for (int k = 0; k < 6; ++k) {
                        int iv = wake->i_vert_d[(size_t)l * wake->n_cols + k];
                        double v = (k < 3) ? o.phi_d[k] : -o.phi_d[k - 3];
                        W_i[P[iv]] = W_i[P[iv]] + v;
                        if (A_abs) S_i[P[iv]] += o.phi_d_abs[k % 3];
                    }
                }
            }
            for (int c = 0; c < N_unknown; ++c) A_i[c] = A_i[c] + W_i[c];
        }
        const size_t orow = (size_t)(row - row0);
        for (int c = 0; c < N_unknown; ++c) A[orow + (size_t)c * ld] = A_i[c];
        if (A_abs)
            for (int c = 0; c < N_unknown; ++c) A_abs[orow + (size_t)c * ld] = std::max(S_i[c], std::fabs(A_i[c]));
        if (I_known) I_known[orow] = I_known_i;
    }
    return status;
}
