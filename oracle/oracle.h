/* ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_aic.cpp header).  C ABI of the CPU restatement. */
#ifndef MACHLINE_ORACLE_H
#define MACHLINE_ORACLE_H
#include "../include/machline_gpu.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_pair_out {
    int in_dod;
    int edges_in_dod[3];
    double F111[3];
    double hH113, H111, H213, H123, h;
    double phi_s;      /* source influence (S space, S_dim = 1)      */
    double phi_d[3];   /* doublet influences (M space, M_dim = 3)    */
    double phi_d_abs[3]; /* sum of |terms| behind phi_d[c]: the scale of its rounding noise (tests only) */
    double v_s[3];     /* source-induced velocity influence, global coordinates (S_dim = 1; panel.f90:3011-3075)   */
    double v_d[9];     /* doublet-induced velocity influences, global coordinates, v_d[3 * i + c] = component i of */
                       /* column c (M_dim = 3; panel.f90:3078-3170)                                                 */
    /* higher-order panels (table->order2): strength-space influences of panel.f90:2815-2914 with S_dim <= 4 and M_dim <= 6;
       for an order-1 panel of such a table phi_s_S[0] = phi_s and phi_d_M[0..2] = phi_d */
    double phi_s_S[4];
    double phi_d_M[6];
    double phi_d_M_abs[6];
    double F121[3], F211[3];
    double H211, H121, H313, H223, H133;
    /* velocity influences of a higher-order table (panel.f90:3011-3170 with S_dim <= 4, M_dim <= 6), global coordinates:
       v_s_S[4 * i + c] / v_d_M[6 * i + c] = component i of column c; for an order-1 panel the first column / three columns */
    double v_s_S[12];
    double v_d_M[18];
    double F113[3], F123[3], F133[3];
    double h3H115, H125, hH135, H145, H215, H225, H235, hH315, H325, H415, H113_3rsh2H115;
} orc_pair_out;

/* tests only: log/atan2 through binary128, rounded once (noise-floor calibration) */
void orc_set_exact_libm(int on);

void orc_pair_influence(const ml_flow *fs, const ml_panel_soa *t, int j, int img, const double *P, orc_pair_out *out);

void orc_pair_batch(const ml_flow *fs, const ml_panel_soa *t, int n_pts, const double *pts, double *phi_d,
                    double *phi_d_abs, double *phi_s, unsigned char *in_dod);

void orc_pair_batch_ho(const ml_flow *fs, const ml_panel_soa *t, int n_pts, const double *pts, double *phi_d6,
                       double *phi_d6_abs, double *phi_s, unsigned char *in_dod);

int orc_assemble(const ml_flow *fs, const ml_panel_soa *body, const ml_panel_soa *wake, const ml_system_map *map,
                 int n_cp, const double *cp_loc, const int *cp_bc, const int *row_perm, int row0, int nrows,
                 double *A_colmajor, int ld, double *I_known, int n_threads,
                 double *A_abs /* NULL or same layout: per entry, the sum of |panel contributions| */);

/* The same with the Neumann rows of panel_solver.f90:1322-1440 / 1570-1650: cp_n_g[n_cp][3] are the control points' normals
   (boundary conditions ML_BC_ZERO_NORMAL_MF: n . B v, ML_BC_ZERO_NORMAL_VEL: n . v). */
int orc_assemble_n(const ml_flow *fs, const ml_panel_soa *body, const ml_panel_soa *wake, const ml_system_map *map,
                   int n_cp, const double *cp_loc, const int *cp_bc, const double *cp_n_g, const int *row_perm, int row0,
                   int nrows, double *A_colmajor, int ld, double *I_known, int n_threads, double *A_abs);

/* common/linalg.f90 solvers on a host system; A is column-major N x N and is overwritten the way
   the reference overwrites A_p.  Returns ml_status. */
int orc_lu_decomp(int N, double *A, int *indx);
int orc_lu_solve(int N, double *A, const double *b, double *x);
int orc_gmres(int N, const double *A, const double *b, double tol, int max_iter, int *total_iter, double *x,
              double *err_history /* NULL or [max_iter] */);
int orc_restarted_gmres(int N, const double *A, const double *b, double tol, int max_iter, int restart_iter,
                        int *total_iter, double *x);
int orc_block_jacobi(int N, double *A, const double *b, int block_size, double tol, double rel, int max_iter,
                     int *total_iter, double *x);
void orc_set_block_history(double *dx, double *err, int capacity);
int orc_block_ssor(int N, double *A, const double *b, int block_size, double tol, double rel, int max_iter,
                   int *total_iter, double *x);
int orc_qr_givens_up(int N, double *A, double *b, double *x);
int orc_qr_fast_givens_up(int N, double *A, double *b, double *x);
int orc_purcell(int N, const double *A, const double *b, double *x);
int orc_lower_bandwidth(int N, const double *A);
/* number of OpenMP threads used by orc_assemble (n_threads <= 0) and the matvecs of the solvers */
void orc_set_threads(int n);

/* panel_solver.f90:1802-2027: b = BC - I_known, preconditioner, dispatch, residual.  A (n x n,
   column-major) is not modified. */
int orc_solve_system(int N, const double *A, const double *I_known, const double *BC, const ml_solver_opts *opts,
                     double *x, ml_solve_info *info);

/* The overdetermined least-squares branch (panel_solver.f90:1842-1895: neumann-mass-flux / neumann-velocity): A is n_cp x N
   column-major (leading dimension n_cp); the normal equations A^T A x = A^T b go through the same preconditioner and solvers;
   the residual is that of the original system. */
int orc_solve_system_ls(int n_cp, int N, const double *A, const double *I_known, const double *BC, const ml_solver_opts *opts,
                        double *x, ml_solve_info *info);

#ifdef __cplusplus
}
#endif
#endif
