// Freestream constants and pressure rules: host restatement of src/flow.f90 plus matinv from
// common/linalg.f90:9-115.  O(1) work per case; feeds ml_flow (include/machline_gpu.h).
#include <cmath>
#include <limits>
#include <stdexcept>
#include <quadmath.h>

#include "model.hpp"

namespace mlh {

quad sqrt_quad(quad x) { return sqrtq(x); }

// linalg.f90:9-115
void matinv(int n, const double* a, double* ai) {
    // n <= 6 everywhere in the setup (3 x 3 panel frames, 4 x 4 and 6 x 6 fits of the higher-order distributions): work arrays on
    // the stack -- this runs several times per panel on the host threads, where heap allocations contend
    double d_small[72];
    int io_small[6];
    std::vector<double> d_big;
    std::vector<int> io_big;
    double* d = d_small;
    int* io = io_small;
    if (n > 6) {
        d_big.assign((size_t)n * 2 * n, 0.0);
        io_big.assign(n, 0);
        d = d_big.data();
        io = io_big.data();
    }
    auto D = [&](int i, int k) -> double& { return d[(size_t)i * 2 * n + k]; };
    for (int i = 0; i < n; ++i) io[i] = i;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            D(i, j) = a[i * n + j];
            D(i, n + j) = (i == j) ? 1.0 : 0.0;
        }
    // Scaling (linalg.f90:51-60)
    for (int i = 0; i < n; ++i) {
        int m = 0;
        for (int k = 1; k < n; ++k)
            if (std::fabs(D(i, k)) > std::fabs(D(i, m))) m = k;
        double tmp = D(i, m);
        for (int k = 0; k < 2 * n; ++k) D(i, k) = D(i, k) / tmp;
    }
    // Lower elimination (linalg.f90:64-87)
    for (int i = 0; i < n - 1; ++i) {
        int m = i;
        for (int j = i + 1; j < n; ++j)
            if (std::fabs(D(io[j], i)) > std::fabs(D(io[m], i))) m = j;
        int itmp = io[m];
        io[m] = io[i];
        io[i] = itmp;
        double r = D(io[i], i);
        for (int k = 0; k < 2 * n; ++k) D(io[i], k) = D(io[i], k) / r;
        for (int j = i + 1; j < n; ++j) {
            r = D(io[j], i);
            for (int k = 0; k < 2 * n; ++k) D(io[j], k) = D(io[j], k) - r * D(io[i], k);
        }
    }
    // Upper elimination (linalg.f90:91-102)
    double r = D(io[n - 1], n - 1);
    for (int k = 0; k < 2 * n; ++k) D(io[n - 1], k) = D(io[n - 1], k) / r;
    for (int i = n - 2; i >= 0; --i)
        for (int j = i + 1; j < n; ++j) {
            r = D(io[i], j);
            for (int k = 0; k < 2 * n; ++k) D(io[i], k) = D(io[i], k) - r * D(io[j], k);
        }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) ai[i * n + j] = D(io[i], n + j);
}

// flow.f90:58-147
void Flow::init(const Json& settings, const std::string& spanwise_axis) {
    const Json* v = settings.find("freestream_velocity");
    if (!v || v->type != Json::Array || v->arr.size() != 3)
        throw std::runtime_error("Freestream velocity was not specified.");
    for (int i = 0; i < 3; ++i) v_inf[i] = v->arr[i].num;
    M_inf = settings.get("freestream_mach_number", 0.0);
    gamma = settings.get("gamma", 1.4);
    if (M_inf < 0.) throw std::runtime_error("Invalid freestream Mach number selected.");
    for (int i = 0; i < 3; ++i) sym_about[i] = (v_inf[i] == 0.);
    U = norm2(v_inf);
    U_inv = 1. / U;
    c_hat_g = v_inf * U_inv;
    if (M_inf == 1.) throw std::runtime_error("A freestream Mach number of 1.0 is not allowed in MachLine.");
    supersonic = M_inf > 1.0;
    incompressible = M_inf == 0.;
    const double pi = 3.14159265358979323846264338327950288419716939937510;
    if (supersonic) {
        B = std::sqrt(M_inf * M_inf - 1.);
        s = -1.;
        K = 2. * pi;
    } else {
        B = std::sqrt(1. - M_inf * M_inf);
        s = 1.;
        K = 4. * pi;
    }
    K_inv = 1. / K;
    c = M_inf * U;
    if (supersonic) {
        mu = std::asin(1.0 / M_inf);
        C_mu = std::cos(mu);
    }

    // calc_metric_matrices, flow.f90:150-187 (off-diagonals start from zero storage)
    double M2 = M_inf * M_inf;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double outer_ij = c_hat_g[i] * c_hat_g[j];
            B_mat_g[i][j] = (i == j ? 1. : 0.) - M2 * outer_ij;
            C_mat_g[i][j] = (i == j ? 1. - M2 : 0.) + M2 * outer_ij;
        }
    B_mat_g_inv = matinv3(B_mat_g);
    for (auto& row : B_mat_c) row = {0., 0., 0.};
    B_mat_c[0][0] = s * (B * B);
    B_mat_c[1][1] = 1.;
    B_mat_c[2][2] = 1.;
    for (auto& row : C_mat_c) row = {0., 0., 0.};
    C_mat_c[0][0] = 1.;
    C_mat_c[1][1] = s * (B * B);
    C_mat_c[2][2] = s * (B * B);

    // calc_transforms, flow.f90:190-251
    V3 j_g{0., 0., 0.};
    if (spanwise_axis == "+x") j_g[0] = 1.;
    else if (spanwise_axis == "-x") j_g[0] = -1.;
    else if (spanwise_axis == "+y") j_g[1] = 1.;
    else if (spanwise_axis == "-y") j_g[1] = -1.;
    else if (spanwise_axis == "+z") j_g[2] = 1.;
    else if (spanwise_axis == "-z") j_g[2] = -1.;
    else j_g[1] = 1.;
    for (auto& row : A_g_to_c) row = {0., 0., 0.};
    A_g_to_c[0] = c_hat_g;
    V3 r3 = cross(c_hat_g, j_g);
    r3 = r3 / norm2(r3);
    A_g_to_c[2] = r3;
    A_g_to_c[1] = cross(r3, c_hat_g);
    for (auto& row : A_c_to_s) row = {0., 0., 0.};
    A_c_to_s[0][0] = 1.;
    A_c_to_s[1][1] = B;
    A_c_to_s[2][2] = B;
    A_g_to_s = matmul(A_c_to_s, A_g_to_c);

    if (!incompressible) {
        a_ise = 2. / (gamma * (M_inf * M_inf));
        b_ise = 0.5 * (gamma - 1.) * (M_inf * M_inf);
        c_ise = gamma / (gamma - 1.);
        C_P_vac = -a_ise;
        C_P_stag = a_ise * (std::pow(1. + b_ise, c_ise) - 1.);
    } else {
        C_P_vac = -std::numeric_limits<double>::max();
        C_P_stag = 1.;
    }
}

// flow.f90:282-310
bool Flow::point_in_dod(const V3& Q, const V3& P) const {
    V3 d = P - Q;
    if (inner(d, c_hat_g) >= 0.) {
        if (C_g_inner(d, d) >= 0.) return true;
    }
    return false;
}

// The rules themselves live in pressure_rules.hpp (shared with the device post-processing kernel)
mlpr::FlowConst Flow::pressure_const() const {
    mlpr::FlowConst f;
    f.U_inv = U_inv;
    f.M_inf = M_inf;
    f.gamma = gamma;
    f.a_ise = a_ise;
    f.b_ise = b_ise;
    f.c_ise = c_ise;
    f.C_P_vac = C_P_vac;
    f.C_P_stag = C_P_stag;
    for (int i = 0; i < 3; ++i) {
        f.v_inf[i] = v_inf[i];
        for (int k = 0; k < 3; ++k) f.A_g_to_c[3 * i + k] = A_g_to_c[i][k];
    }
    return f;
}
static inline void v3_arr(const V3& v, double a[3]) {
    a[0] = v[0];
    a[1] = v[1];
    a[2] = v[2];
}
double Flow::get_C_P_inc(const V3& v) const {
    double a[3];
    v3_arr(v, a);
    return mlpr::C_P_inc(pressure_const(), a);
}
double Flow::get_C_P_ise(const V3& v) const {
    double a[3];
    v3_arr(v, a);
    return mlpr::C_P_ise(pressure_const(), a);
}
V3 Flow::get_v_pert_c(const V3& v) const { return matvec(A_g_to_c, v - v_inf); }
void Flow::restrict_pressure(double& C_P) const { C_P = mlpr::restrict_pressure(pressure_const(), C_P); }
double Flow::get_C_P_lin(const V3& v) const {
    double a[3];
    v3_arr(v, a);
    return mlpr::C_P_lin(pressure_const(), a);
}
double Flow::get_C_P_sln(const V3& v) const {
    double a[3];
    v3_arr(v, a);
    return mlpr::C_P_sln(pressure_const(), a);
}
double Flow::get_C_P_2nd(const V3& v) const {
    double a[3];
    v3_arr(v, a);
    return mlpr::C_P_2nd(pressure_const(), a);
}

double Flow::get_C_P_crit(double M) const {
    double M2 = M * M;
    double x = 0.5 * (gamma - 1.);
    double n = 1. + x * M2;
    double d = 1. + x;
    return 2. / (gamma * M2) * (std::pow(n / d, gamma / (gamma - 1.)) - 1.);
}

int pressure_rule_id(const std::string& rule) {
    static const char* names[mlpr::RULE_COUNT] = {"incompressible", "isentropic", "second-order", "slender-body", "linear",
                                                  "prandtl-glauert", "karman-tsien", "laitone"};
    for (int i = 0; i < mlpr::RULE_COUNT; ++i)
        if (rule == names[i]) return i;
    return -1;
}

// flow.f90:528-585
double Flow::get_C_P(const V3& v, const std::string& rule, double M_corr) const {
    const int id = pressure_rule_id(rule);
    if (id < 0) throw std::runtime_error("unknown pressure rule " + rule);
    double a[3];
    v3_arr(v, a);
    return mlpr::C_P(pressure_const(), a, id, M_corr);
}

}  // namespace mlh
