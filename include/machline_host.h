/*
 * machline_host.h -- C ABI of the host-side setup library (libmachline_host.so).
 *
 * It restates what MachLine's `main` does around the two hot paths (src/main.f90:102-160):
 * read the JSON input, load and analyse the mesh, build the wake, place control points, compute
 * the per-panel tables, and -- after the solve -- turn x into mu/sigma, cell velocities, pressure
 * coefficients, forces and moments.  It contains NO influence-coefficient or linear-solver code:
 * those are only reachable through include/machline_gpu.h (CUDA).  The tables it hands out are
 * exactly the arguments of ml_set_flow / ml_set_panels / ml_set_control_points / ml_set_system_map.
 */
#ifndef MACHLINE_HOST_H
#define MACHLINE_HOST_H

#include "machline_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mlh_case mlh_case;

typedef struct mlh_cp_table {
    int n_cp;
    const double *loc;      /* [n_cp][3]  control_point%loc                                         */
    const int *bc;          /* [n_cp]     control_point%bc (ML_BC_*)                                */
    const double *n_g;      /* [n_cp][3]  normal for Neumann conditions (zeros for Dirichlet)       */
    const int *row_perm;    /* [n_cp]     row of A owned by control point i                         */
    const double *BC;       /* [n_cp]     boundary-condition vector in row order (:1104-1159)       */
} mlh_cp_table;

typedef struct mlh_solver_settings {   /* solver.* keys after defaults (panel_solver.f90:173-209) */
    ml_solver_opts opts;
    char matrix_solver_name[16];
    char formulation[48];
    int sort_system;
    int write_A_and_b;
    int run_checks;         /* solver.run_checks (main.f90:66 / panel_solver.f90:1828): check_system before the solve */
} mlh_solver_settings;

typedef struct mlh_results {
    double C_p_max, C_p_min;  /* of the rule test/test_machline.py:62-66 reads                      */
    double C_F[3], C_M[3];
    int n_cells, n_mu;
    const double *mu;         /* [n_mu]                                                             */
    const double *C_p;        /* [n_cells] same rule as C_p_max/min                                 */
    const double *V_cells;    /* [n_cells][3]                                                       */
} mlh_results;

typedef struct mlh_mesh_info {
    int n_body_panels, n_body_verts, n_wake_panels, n_wake_strips, n_edges, n_cp, n_unknown;
    int mirrored, asym_flow, mirror_plane, supersonic;
    double sort_seconds;
} mlh_mesh_info;

/* Returns 0 on success; on failure returns nonzero and mlh_last_error() describes it (the
   reference would have printed "!!! ..." and stopped). */
int mlh_case_create(const char *json_text, const char *base_dir, mlh_case **out);
void mlh_case_destroy(mlh_case *c);
const char *mlh_last_error(void);

int mlh_case_info(const mlh_case *c, mlh_mesh_info *out);
/* Pointers stay valid until mlh_case_destroy. wake->n_panels == 0 when no wake is appended. */
int mlh_case_tables(mlh_case *c, ml_flow *flow, ml_panel_soa *body, ml_panel_soa *wake,
                    ml_system_map *map, mlh_cp_table *cps);
int mlh_case_solver_settings(const mlh_case *c, mlh_solver_settings *out);
/* x[n_unknown] in permuted order, as returned by ml_solve. Result pointers valid until the next
   mlh_case_post or destroy. */
int mlh_case_post(mlh_case *c, const double *x, mlh_results *out);
/* The same for the formulations without a prescribed inner flow (neumann-*): v_inner[n_points][3] = the induced velocity
   (doublet + source, per unit freestream speed) at the points of mlh_case_inner_points -- just inside every panel, where
   panel_solver_calc_cell_velocities evaluates it (src/panel_solver.f90:2063-2066, 2080-2083).  pts may be NULL to query the count. */
int mlh_case_inner_points(mlh_case *c, double *pts, int *n_points);
int mlh_case_post2(mlh_case *c, const double *x, const double *v_inner, mlh_results *out);
/* Per-rule arrays of the last mlh_case_post / mlh_case_post2: rule = ML_RULE_* (pressure coefficients, [n_cells]) or -1 (the
   cells' force contributions dC_f, [n_cells][3]).  dst may be NULL to query the length; *n = 0 when the rule was not computed. */
int mlh_case_result_array(mlh_case *c, int rule, double *dst, int *n);
/* Tables and constants of ml_post_process (include/machline_gpu.h: the lower-order post-processing on the device,
   panel_solver.f90:2030-2615), prepared with the operations of mlh_case_post.  v_inner as in mlh_case_post2 (NULL for the
   Dirichlet formulations).  Pointers stay valid until the next call or mlh_case_destroy. */
int mlh_case_post_tables(mlh_case *c, const double *v_inner, ml_post_tables *tables, ml_post_flow *flow);
/* Write report.json in the reference's layout (panel_solver.f90:2618-2746) */
int mlh_case_write_report(mlh_case *c, const char *path, const ml_solve_info *info, int solver_stat,
                          double total_runtime);

/* Result files of the last mlh_case_post / mlh_case_post2 in the reference's legacy-VTK layout (src/vtk.f90:42-431):
   output.body_file / output.mirrored_body_file (surface_mesh_write_body / write_body_mirror, src/surface_mesh.f90:2547-2735;
   mirrored != 0 needs an asymmetric mirrored flow), output.wake_file (wake_mesh_write_strips, src/wake_mesh.f90:148-222;
   *exported = 0 and no file when there is no wake) and output.control_point_file (surface_mesh_write_control_points,
   src/surface_mesh.f90:2738-2777; residual[n_cp] = A x - b in row order as ml_residual returns it, or NULL). */
int mlh_case_write_body(mlh_case *c, const char *path, int mirrored);
int mlh_case_write_wake(mlh_case *c, const char *path, int *exported);
int mlh_case_write_control_points(mlh_case *c, const char *path, const double *residual);

#ifdef __cplusplus
}
#endif
#endif
