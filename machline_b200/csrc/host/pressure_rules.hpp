// Pressure rules of the reference (src/flow.f90:313-585: get_C_P_inc / _ise / _2nd / _sln / _lin, restrict_pressure and the
// subsonic corrections Prandtl-Glauert :453-466, Karman-Tsien :469-487, Laitone :490-508) on plain arrays, compiled BOTH into the
// host library (Flow::get_C_P, flow.cpp) and into the device post-processing kernel (csrc/gpu/post.cu, built with -fmad=false as
// the host is built with -ffp-contract=off): one statement of the arithmetic, the same IEEE operations on both sides (pow differs
// by the libraries' last bit).
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define ML_PR_HD __host__ __device__ inline
#else
#define ML_PR_HD inline
#endif

namespace mlpr {

enum Rule {   // bit positions of ml_post_flow::rules (include/machline_gpu.h)
    RULE_INCOMPRESSIBLE = 0,
    RULE_ISENTROPIC = 1,
    RULE_SECOND_ORDER = 2,
    RULE_SLENDER_BODY = 3,
    RULE_LINEAR = 4,
    RULE_PRANDTL_GLAUERT = 5,
    RULE_KARMAN_TSIEN = 6,
    RULE_LAITONE = 7,
    RULE_COUNT = 8
};

struct FlowConst {
    double U_inv, M_inf, gamma;
    double a_ise, b_ise, c_ise, C_P_vac, C_P_stag;
    double v_inf[3];
    double A_g_to_c[9];   // row-major
};

ML_PR_HD double C_P_inc(const FlowConst& f, const double v[3]) {
    return 1. - ((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]) * f.U_inv * f.U_inv;
}
ML_PR_HD double C_P_ise(const FlowConst& f, const double v[3]) {
    double C = C_P_inc(f, v);
    C = f.a_ise * (pow(1. + f.b_ise * C, f.c_ise) - 1.);
    if (C != C) C = f.C_P_vac;   // NaN: beyond vacuum
    return C;
}
ML_PR_HD void v_pert_c(const FlowConst& f, const double v[3], double vp[3]) {
    const double d[3] = {v[0] - f.v_inf[0], v[1] - f.v_inf[1], v[2] - f.v_inf[2]};
    for (int i = 0; i < 3; ++i) vp[i] = (f.A_g_to_c[3 * i] * d[0] + f.A_g_to_c[3 * i + 1] * d[1]) + f.A_g_to_c[3 * i + 2] * d[2];
}
ML_PR_HD double restrict_pressure(const FlowConst& f, double C) {
    if (C > f.C_P_stag) return f.C_P_stag;
    if (C < f.C_P_vac) return f.C_P_vac;
    return C;
}
ML_PR_HD double C_P_lin(const FlowConst& f, const double v[3]) {
    double vp[3];
    v_pert_c(f, v, vp);
    return restrict_pressure(f, -2. * vp[0] * f.U_inv);
}
ML_PR_HD double C_P_sln(const FlowConst& f, const double v[3]) {
    const double C_lin = C_P_lin(f, v);
    double vp[3];
    v_pert_c(f, v, vp);
    return restrict_pressure(f, C_lin - (vp[1] * vp[1] + vp[2] * vp[2]) * (f.U_inv * f.U_inv));
}
ML_PR_HD double C_P_2nd(const FlowConst& f, const double v[3]) {
    const double C_sln = C_P_sln(f, v);
    double vp[3];
    v_pert_c(f, v, vp);
    return restrict_pressure(f, C_sln - (1. - f.M_inf * f.M_inf) * (vp[0] * vp[0]) * (f.U_inv * f.U_inv));
}

// rule: one of Rule; M_corr: the Mach number of the subsonic corrections (solver.M_inf_corr)
ML_PR_HD double C_P(const FlowConst& f, const double v[3], int rule, double M_corr) {
    switch (rule) {
        case RULE_INCOMPRESSIBLE: return C_P_inc(f, v);
        case RULE_ISENTROPIC: return C_P_ise(f, v);
        case RULE_SECOND_ORDER: return C_P_2nd(f, v);
        case RULE_SLENDER_BODY: return C_P_sln(f, v);
        case RULE_LINEAR: return C_P_lin(f, v);
        case RULE_PRANDTL_GLAUERT: {
            const double C = C_P_inc(f, v);
            return C / sqrt(1. - M_corr * M_corr);
        }
        case RULE_KARMAN_TSIEN: {
            const double C = C_P_inc(f, v);
            const double M2 = M_corr * M_corr, sM2 = sqrt(1. - M2);
            const double x = M2 / (1. + sM2);
            return C / (sM2 + 0.5 * x * C);
        }
        case RULE_LAITONE: {
            const double C = C_P_inc(f, v);
            const double M2 = M_corr * M_corr, sM2 = sqrt(1. - M2);
            const double x = M2 * (1. + (0.5 * (f.gamma - 1.) * M2)) / (2 * sM2);
            return C / (sM2 + (x * C));
        }
        default: return 0.;
    }
}

}  // namespace mlpr
