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
    std::vector<double> d((size_t)n * 2 * n, 0.0);
    std::vector<int> io(n);
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

double Flow::get_C_P_inc(const V3& v) const { return 1. - inner(v, v) * U_inv * U_inv; }

double Flow::get_C_P_ise(const V3& v) const {
    double C = get_C_P_inc(v);
    C = a_ise * (std::pow(1. + b_ise * C, c_ise) - 1.);
    if (std::isnan(C)) C = C_P_vac;
    return C;
}

V3 Flow::get_v_pert_c(const V3& v) const { return matvec(A_g_to_c, v - v_inf); }

void Flow::restrict_pressure(double& C_P) const {
    if (C_P > C_P_stag) C_P = C_P_stag;
    else if (C_P < C_P_vac) C_P = C_P_vac;
}

double Flow::get_C_P_lin(const V3& v) const {
    V3 vp = get_v_pert_c(v);
    double C = -2. * vp[0] * U_inv;
    restrict_pressure(C);
    return C;
}

double Flow::get_C_P_sln(const V3& v) const {
    double C_lin = get_C_P_lin(v);
    V3 vp = get_v_pert_c(v);
    double C = C_lin - (vp[1] * vp[1] + vp[2] * vp[2]) * (U_inv * U_inv);
    restrict_pressure(C);
    return C;
}

double Flow::get_C_P_2nd(const V3& v) const {
    double C_sln = get_C_P_sln(v);
    V3 vp = get_v_pert_c(v);
    double C = C_sln - (1. - M_inf * M_inf) * (vp[0] * vp[0]) * (U_inv * U_inv);
    restrict_pressure(C);
    return C;
}

double Flow::get_C_P_crit(double M) const {
    double M2 = M * M;
    double x = 0.5 * (gamma - 1.);
    double n = 1. + x * M2;
    double d = 1. + x;
    return 2. / (gamma * M2) * (std::pow(n / d, gamma / (gamma - 1.)) - 1.);
}

// flow.f90:528-585
double Flow::get_C_P(const V3& v, const std::string& rule, double M_corr) const {
    if (rule == "incompressible") return get_C_P_inc(v);
    if (rule == "isentropic") return get_C_P_ise(v);
    if (rule == "second-order") return get_C_P_2nd(v);
    if (rule == "slender-body") return get_C_P_sln(v);
    if (rule == "linear") return get_C_P_lin(v);
    if (rule == "prandtl-glauert") {  // flow.f90:453-466
        double C = get_C_P_inc(v);
        return C / std::sqrt(1. - M_corr * M_corr);
    }
    if (rule == "karman-tsien") {  // flow.f90:469-487
        double C = get_C_P_inc(v);
        double M2 = M_corr * M_corr, sM2 = std::sqrt(1. - M2);
        double x = M2 / (1. + sM2);
        return C / (sM2 + 0.5 * x * C);
    }
    if (rule == "laitone") {  // flow.f90:490-508
        double C = get_C_P_inc(v);
        double M2 = M_corr * M_corr, sM2 = std::sqrt(1. - M2);
        double x = M2 * (1. + (0.5 * (gamma - 1.) * M2)) / (2 * sM2);
        return C / (sM2 + (x * C));
    }
    throw std::runtime_error("unknown pressure rule " + rule);
}

}  // namespace mlh
