/*
 * machline_gpu.h -- C ABI of the B200 (sm_100a) replacement for MachLine's two hot paths:
 *   (1) AIC assembly  (reference: src/panel_solver.f90:651-775 DoD pre-pass,
 *                                  src/panel_solver.f90:1290-1501 calc_body_influences,
 *                                  src/panel_solver.f90:1504-1706 calc_wake_influences,
 *                                  src/panel_solver.f90:1203-1287 update_system_row,
 *                                  src/panel.f90:1732-2971 check_dod .. calc_potential_influences)
 *   (2) dense solve   (reference: src/panel_solver.f90:1802-2027 solve_system,
 *                                  common/linalg.f90:118-342 lu_solve, :1235-1334 GMRES,
 *                                  :1337-1453 restarted_GMRES, :601-728 block_jacobi_solve,
 *                                  :1798-1831 diagonal_preconditioner)
 *
 * The reference has no FFI; the two seams are the Fortran call sites
 *   call this%calc_body_influences(body) / call this%calc_wake_influences(body)   (panel_solver.f90:1072,1075)
 *   select case(this%matrix_solver) ... end select                                (panel_solver.f90:1915-1975)
 * A Fortran bind(C) shim flattens type(panel)/type(control_point) into the
 * plain arrays below and calls these entry points.  Everything is 0-based on this side of the ABI.
 *
 * All pointers are HOST pointers owned by the caller for the duration of the call only.
 * Every entry point returns ml_status; none of them aborts the process (the reference `stop`s).
 * A context is bound to one CUDA device and is not thread-safe.
 */
#ifndef MACHLINE_GPU_H
#define MACHLINE_GPU_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ml_ctx ml_ctx;

/* Status codes: 0..4 follow the reference's solver_stat (panel_solver.f90:1722-1761, 2006-2010). */
typedef enum ml_status {
    ML_OK = 0,
    ML_NAN_IN_SYSTEM = 1,      /* NaN in A or b          (check_system, panel_solver.f90:1723-1730) */
    ML_UNINFLUENCED = 2,       /* zero row / column      (panel_solver.f90:1735-1761)               */
    ML_SINGULAR = 3,           /* lu_decomp code 1       (linalg.f90:138-140, 205-208)              */
    ML_NAN_RESIDUAL = 4,       /* NaN residual           (panel_solver.f90:2006-2010)               */
    ML_BAD_ARGUMENT = 10,
    ML_NOT_READY = 11,         /* e.g. ml_solve before ml_assemble                                  */
    ML_UNSUPPORTED = 12,       /* feature outside the hot-path scope                    */
    ML_CUDA_ERROR = 100,
    ML_NCCL_ERROR = 101
} ml_status;

/* Boundary-condition codes, identical to base_geom.f90:14-20. */
enum { ML_BC_ZERO_POTENTIAL = 1, ML_BC_SF_POTENTIAL = 2, ML_BC_ZERO_NORMAL_MF = 3,
       ML_BC_STRENGTH_MATCHING = 4, ML_BC_ZERO_NORMAL_VEL = 5, ML_BC_ZERO_X_VEL = 6,
       ML_BC_MF_INNER_FLOW = 7 };

/* Freestream constants (type flow, src/flow.f90:11-29).  3x3 matrices are row-major. */
typedef struct ml_flow {
    double M_inf;
    double B;            /* sqrt|1-M^2|                       flow.f90:100-108 */
    double s;            /* sign(1-M^2): +1 elliptic, -1 hyperbolic            */
    double K_inv;        /* 1/(4 pi) subsonic, 1/(2 pi) supersonic             */
    double c_hat_g[3];   /* compressibility axis = freestream direction        */
    double B_mat_g[9];   /* dual metric matrix   flow.f90:158-163              */
    double C_mat_g[9];   /* metric matrix        flow.f90:174-179              */
    int supersonic;
    int mirror_plane;    /* 0 = none, else 1..3 = index of the normal to the mirror plane (mesh.f90:16) */
} ml_flow;

/*
 * Panel table (type panel members used at evaluation time, src/panel.f90:38-70).
 * n_rec = n_panels * n_images records; record r < n_panels is the panel itself, record
 * r + n_panels its mirrored twin (the *_mir members).  Arrays are record-major (AoS inside a
 * record, e.g. centr[3*r + c]); 3x3 matrices are row-major.
 */
typedef struct ml_panel_soa {
    int n_panels;
    int n_images;              /* 1, or 2 when mirrored twins are present                              */
    int n_cols;                /* entries of i_vert_d per panel: M_dim (3) for body, 2*M_dim (6) wake; 6
                                  for a higher-order body table (order2 below)                         */
    int in_wake;               /* 1: wake table (doublet only, bottom = -top, panel.f90:2909-2912)     */
    const double *centr;       /* [n_rec][3]      centr / centr_mir                                    */
    const double *A_g_to_ls;   /* [n_rec][3][3]   A_g_to_ls / A_g_to_ls_mir                            */
    const double *vertices_ls; /* [n_rec][3][2]   vertex k -> (xi, eta) = Fortran vertices_ls(:,k)     */
    const double *n_hat_ls;    /* [n_rec][3][2]   edge k   -> (n_xi, n_eta) = n_hat_ls(:,k)            */
    const double *b;           /* [n_rec][3]      edge parameter (panel.f90:525-539)                   */
    const double *sqrt_b;      /* [n_rec][3]                                                           */
    const double *J;           /* [n_rec]         area Jacobian  (panel.f90:469)                       */
    const int    *r;           /* [n_rec]         inclination indicator; must be +1 (panel.f90:439)    */
    const double *area;        /* [n_panels]      A > 0 test (panel.f90:2933)                          */
    const double *vert_g;      /* [n_rec][3][3]   global vertex locations (mirrored for the twin)      */
    const double *T_mu;        /* [n_rec][3][3]   T_mu / T_mu_mir, row-major (mu_dim x M_dim)          */
    const int    *i_vert_d;    /* [n_panels][n_cols] 0-based doublet unknown ids (panel.f90:618-638)   */
    const int    *i_panel_s;   /* [n_panels]      0-based source panel id (panel.f90:680); body only   */
    const unsigned char *has_sources;   /* [n_panels]; body only                                       */
    const unsigned char *image_present; /* [n_panels] or NULL (= all): whether record r+n_panels is
                                           evaluated (wake_strip%mirrored, wake_strip.f90:49)          */
    /* Higher-order distributions (geometry.singularity_order = "higher": quadratic doublets, linear sources;
       panel.f90:544-969).  order2 = 0 (and the pointers below NULL) for a lower-order table.  With order2 = 1 the body
       table has n_cols = 6: row j of i_vert_d holds the M_dim(j) doublet ids of panel j (its 3 vertices, then the vertex
       opposite every continuous edge, panel.f90:605-665), -1 padded.  Wake tables are always lower order. */
    int order2;
    const unsigned char *order; /* [n_panels]      1 or 2 (panel.f90:556-565: 3 discontinuous edges -> 1)      */
    const int    *M_dim;        /* [n_panels]      3..6                                                        */
    const double *T_mu6;        /* [n_rec][6][6]   T_mu / T_mu_mir (mu_dim x M_dim), zero padded; order-1 panels
                                                   carry their 3 x 3 in the upper left corner                  */
    const int    *S_dim;        /* [n_panels]      1..4                                                        */
    const int    *i_panel_s4;   /* [n_panels][4]   source panel ids (panel.f90:668-693), -1 padded             */
    const double *T_sigma;      /* [n_rec][3][4]   T_sigma / T_sigma_mir (sigma_dim x S_dim), zero padded;
                                                   order-1 panels: T_sigma(0,0) = 1                            */
} ml_panel_soa;

/* Unknown / index bookkeeping used by update_system_row (panel_solver.f90:1203-1287). */
typedef struct ml_system_map {
    int n_cp;                  /* rows of A                                                            */
    int n_unknown;             /* columns of A                                                         */
    int n_verts;               /* body%N_verts (after cloning)                                         */
    int n_body_panels;         /* body%N_panels                                                        */
    int n_sigma;               /* N_panels or 2*N_panels                                               */
    int mirrored;              /* body%mirrored                                                        */
    int asym_flow;             /* body%asym_flow                                                       */
    const int *P;              /* [n_unknown] permutation, 0-based (panel_solver.f90:778-1030)         */
    const unsigned char *sigma_known; /* [n_sigma]                                                     */
    const int *i_sigma_in_sys; /* [n_sigma] 0-based unknown id (before P) or -1                        */
    const double *sigma;       /* [n_sigma] known source strengths (panel_solver.f90:1162-1200)        */
} ml_system_map;

typedef enum ml_matrix_solver {     /* solver.matrix_solver values, panel_solver.f90:1915-1975 */
    ML_SOLVER_LU = 0, ML_SOLVER_QRUP = 1, ML_SOLVER_FQRUP = 2, ML_SOLVER_GMRES = 3,
    ML_SOLVER_RGMRES = 4, ML_SOLVER_PURC = 5, ML_SOLVER_BSSOR = 6, ML_SOLVER_BJAC = 7
} ml_matrix_solver;

typedef enum ml_preconditioner { ML_PREC_NONE = 0, ML_PREC_DIAG = 1 } ml_preconditioner;

typedef struct ml_solver_opts {     /* defaults: panel_solver.f90:173-209 */
    int matrix_solver;         /* ml_matrix_solver                                                      */
    int preconditioner;        /* ml_preconditioner; DIAG reproduces linalg.f90:1813-1816 (1/A(N,N))    */
    double tol;                /* 1e-12                                                                 */
    double rel;                /* 0.8                                                                   */
    int max_iterations;        /* 1000                                                                  */
    int restart_iterations;    /* 20                                                                    */
    int block_size;            /* <= 0 -> N/5 (panel_solver.f90:1910-1912)                              */
    const char *iteration_file;/* NULL or "none": no iteration history; else the file the reference's solver
                                  writes (solver.iterative_solver_output): GMRES linalg.f90:1273-1280,1316;
                                  RGMRES :1376-1383,1438; BJAC :659-666,717; BSSOR :514-520,587            */
} ml_solver_opts;

typedef struct ml_solve_info {
    int iterations;            /* -1 for direct solvers (panel_solver.f90:189)                          */
    double res_max;            /* max |A x - b|         (panel_solver.f90:1997)                         */
    double res_norm;           /* ||A x - b||_2         (panel_solver.f90:1998)                         */
    double assemble_ms;        /* device time of the last ml_assemble                                   */
    double solve_ms;           /* device time of the last ml_solve                                      */
} ml_solve_info;

/* ---- lifecycle ------------------------------------------------------------------------------ */
int  ml_abi_version(void);
ml_status ml_ctx_create(ml_ctx **out, int device_id);
/* Single-process multi-GPU context (the reference's main is one process, src/main.f90:133): ONE host thread
   drives the two hot paths on n_dev devices through the same entry points as a single-device context -- ml_set_* take the
   same tables, ml_assemble returns the whole I_known, ml_get_A / ml_set_A address any rows, ml_solve returns the whole x,
   ml_check_system covers all shards.  Rows of the permuted system are dealt to the devices in contiguous blocks, or
   block-cyclically (blocks of block_rows) after ml_multi_set_dealing (load balance of the sharded LU); on several devices
   ml_solve supports the solvers of a row-sharded system (GMRES, RGMRES, LU).  Inside, one ordinary context per device,
   NCCL + peer-memory windows between them (peer access inside the process instead of CUDA IPC), one worker thread per
   device for the duration of a call.  ml_set_row_shard*, ml_set_communicator, ml_device_system and ml_device_stream do
   not apply to such a handle. */
ml_status ml_ctx_create_multi(ml_ctx **out, const int *device_ids, int n_dev);
ml_status ml_multi_set_dealing(ml_ctx *ctx, int block_rows);   /* 0: contiguous blocks; > 0: block-cyclic; -1 (default): contiguous
                                                                  for subsonic flows, block-cyclic 128 for supersonic ones (the rows
                                                                  of a sorted supersonic system cost more the further downstream) */
int ml_device_count(const ml_ctx *ctx);                        /* devices behind the handle (1 for ml_ctx_create) */
void ml_ctx_destroy(ml_ctx *ctx);
const char *ml_last_error(const ml_ctx *ctx);

/* ---- inputs (host -> device) ---------------------------------------------------------------- */
ml_status ml_set_flow(ml_ctx *ctx, const ml_flow *flow);
/* body is required; wake may be NULL or have n_panels == 0 (panel_solver.f90:1075). */
ml_status ml_set_panels(ml_ctx *ctx, const ml_panel_soa *body, const ml_panel_soa *wake);
/* loc[n_cp][3], bc[n_cp] (ML_BC_*), n_g[n_cp][3] or NULL, row_perm[n_cp] = row of A that control
   point i owns (P(i) when use_sort_for_cp, else i; panel_solver.f90:1484-1490). */
ml_status ml_set_control_points(ml_ctx *ctx, int n_cp, const double *loc, const int *bc,
                                const double *n_g, const int *row_perm);
ml_status ml_set_system_map(ml_ctx *ctx, const ml_system_map *map);
/* Multi-GPU: this context builds and owns rows [row0, row0+nrows) of the permuted system.
   Default: all rows. */
ml_status ml_set_row_shard(ml_ctx *ctx, int row0, int nrows);
/* Multi-GPU, block-cyclic alternative (load balance of the sharded LU): this context owns the blocks
   b = rank (mod world) of block_rows consecutive rows.  ml_get_A / ml_set_A then address runs of rows that are
   consecutive inside the context; ml_local_rows lists the global row of every local row (after ml_assemble; pass
   rows_out = NULL to query the count). */
ml_status ml_set_row_shard_cyclic(ml_ctx *ctx, int block_rows, int rank, int world_size);
ml_status ml_local_rows(ml_ctx *ctx, int *rows_out, int *n_out);
/* Multi-GPU: join an NCCL communicator (id = the 128-byte ncclUniqueId made by rank 0). */
ml_status ml_set_communicator(ml_ctx *ctx, const void *nccl_unique_id, int rank, int world_size);

/* ---- hot path 1: AIC assembly ---------------------------------------------------------------- */
/* Builds A (device resident, column-major, ld = local rows) and I_known.  I_known_out (host,
   length = local rows, indexed by local row) may be NULL. */
ml_status ml_assemble(ml_ctx *ctx, double *I_known_out);
/* Copy rows [row0,row0+nrows) (global row ids inside this context's shard) of A to a host
   column-major array with leading dimension ld (parity channel = write_A_and_b). */
ml_status ml_get_A(ml_ctx *ctx, int row0, int nrows, double *dst_colmajor, int ld);
/* The inverse channel: overwrite rows [row0,row0+nrows) of the resident system from a host column-major array (a system
   assembled elsewhere, or a test matrix, solved through ml_solve on the row shards). */
ml_status ml_set_A(ml_ctx *ctx, int row0, int nrows, const double *src_colmajor, int ld);
/* Pair count of the last ml_assemble on this context: local rows x (body records + wake records). */
long long ml_pair_count(const ml_ctx *ctx);

/* panel_solver_check_system (panel_solver.f90:1709-1764; run when solver.run_checks is set, :1828-1831) on the resident
   system: ML_NAN_IN_SYSTEM when A or b = BC - I_known holds a NaN, else ML_UNINFLUENCED when a row of A (a control point
   that nothing influences) or a column (a vertex that exerts no influence) is entirely zero, else ML_OK.  n_zero_rows /
   n_zero_cols (either may be NULL) receive the counts.  With a communicator the verdict covers all shards. */
ml_status ml_check_system(ml_ctx *ctx, const double *BC, int *n_zero_rows, int *n_zero_cols);

/* R_cp = A x - b on this context's rows, b = BC - I_known (panel_solver.f90:1992; the point data "residual" of
   output.control_point_file): r_out is indexed like I_known_out of ml_assemble. */
ml_status ml_residual(ml_ctx *ctx, const double *BC, const double *x, double *r_out);

/* ---- hot path 2: dense solve ------------------------------------------------------------------ */
/* BC[n_cp] is the boundary-condition vector in row order (panel_solver.f90:1104-1159); the library
   forms b = BC - I_known (:1818), applies the reference's "preconditioner", dispatches on
   matrix_solver, and returns x[n_unknown] (in permuted order; mu(i) = x(P(i)), :2018-2020).
   With a communicator every rank passes the full BC and receives the full x. */
ml_status ml_solve(ml_ctx *ctx, const ml_solver_opts *opts, const double *BC, double *x_out,
                   ml_solve_info *info);

/* Stand-alone dense solve of a host system (the lu_solve / GMRES / block_jacobi_solve signatures,
   linalg.f90:118, 1235, 601): A is column-major N x N and is not modified. */
ml_status ml_solve_dense(ml_ctx *ctx, int N, const double *A_colmajor, const double *b,
                         const ml_solver_opts *opts, double *x_out, ml_solve_info *info);

/* ---- post-processing on the device ---------------------------------------------------------
 * panel_solver_calc_cell_velocities, calc_pressures, calc_forces and calc_moments (src/panel_solver.f90:2030-2095,
 * 2218-2321, 2440-2528, 2551-2615) with panel_get_velocity_jump (src/panel.f90:3415-3512) and the pressure rules
 * (src/flow.f90:313-585) for LOWER-ORDER panels: one thread per cell (= panel image), reading the solution that the
 * last ml_solve left on the device, so that x only has to leave the GPU as results.  Higher-order panels (quadratic
 * pressure distributions, src/panel.f90:3541-3740) are post-processed by the host library (mlh_case_post). */
typedef struct ml_post_tables {
    int n_cells;                /* N_panels, or 2 N_panels in an asymmetric mirrored flow (panel_solver.f90:2040-2047)   */
    const int *mu_index;        /* [n_cells][3] position in x of each vertex's doublet strength (mu(i) = x(P(i)),
                                   panel_solver.f90:2018-2020, with the mirror shift of panel.f90:3380-3400); -1: zero  */
    const double *T_mu;         /* [n_cells][9] row-major T_mu of the cell's image (rows 2 and 3 give the gradient)      */
    const double *A_g_to_ls;    /* [n_cells][9] row-major                                                               */
    const double *s_dir;        /* [n_cells][3] n_g / (nu_g . n_g) (panel.f90:3489-3493); zeros: panel without sources   */
    const int *sigma_index;     /* [n_cells] position in x of an unknown source strength, -1: use sigma_known            */
    const double *sigma_known;  /* [n_cells]                                                                             */
    const double *v_inner;      /* [n_cells][3] velocity just inside the cell per unit U: the prescribed inner flow of the
                                   Dirichlet formulations, v_inf / U + induced velocity for the Neumann ones (:2063-2066) */
    const double *n_g;          /* [n_cells][3] normal of the cell's image                                               */
    const double *area;         /* [n_cells]                                                                             */
    const double *centr;        /* [n_cells][3] centroid of the cell's image                                             */
    const int *force_cell;      /* [n_cells] the cell whose force enters this cell's moment: the reference uses the
                                   un-mirrored panel's force for the mirrored cell (panel_solver.f90:2583)               */
} ml_post_tables;

enum { ML_RULE_INCOMPRESSIBLE = 0, ML_RULE_ISENTROPIC, ML_RULE_SECOND_ORDER, ML_RULE_SLENDER_BODY, ML_RULE_LINEAR,
       ML_RULE_PRANDTL_GLAUERT, ML_RULE_KARMAN_TSIEN, ML_RULE_LAITONE, ML_RULE_COUNT };

typedef struct ml_post_flow {
    double U, U_inv, M_inf, gamma;                          /* flow.f90:58-147                                          */
    double a_ise, b_ise, c_ise, C_P_vac, C_P_stag;          /* constants of the isentropic rule and the limits          */
    double M_inf_corr;                                      /* Mach number of the subsonic corrections                  */
    double v_inf[3], A_g_to_c[9];
    double CG[3], S_ref, l_ref;                             /* references of the force / moment coefficients            */
    int rules;                                              /* bit ML_RULE_*: which pressure coefficients to compute    */
    int force_rule;                                         /* ML_RULE_* of solver.pressure_for_forces                  */
    int mirrored_symmetric;                                 /* mirrored mesh in a symmetric flow: C_F, C_M doubled /
                                                               zeroed as panel_solver.f90:2513-2521, 2600-2612          */
    int mirror_plane;                                       /* 1..3                                                     */
} ml_post_flow;

typedef struct ml_post_out {        /* host buffers; any array may be NULL                                               */
    double *V_cells;                /* [n_cells][3]                                                                      */
    double *C_p[ML_RULE_COUNT];     /* [n_cells] each, for the rules selected in ml_post_flow::rules                     */
    double *dC_f;                   /* [n_cells][3]                                                                      */
    double C_F[3], C_M[3];
    double C_p_max, C_p_min;        /* of the rule test/test_machline.py:62-66 reads: incompressible if computed, else isentropic */
} ml_post_out;

/* Needs a successful ml_solve on this context (ML_BAD_ARGUMENT otherwise).  x_override (host, n_unknown) replaces the device
   solution when non-NULL (tests). */
ml_status ml_post_process(ml_ctx *ctx, const ml_post_tables *tables, const ml_post_flow *flow, const double *x_override,
                          ml_post_out *out);

/* ---- introspection for tests / benches --------------------------------------------------------- */
/* Number of kernel launches issued by this context since creation (bench.py "gpu_launches"). */
long long ml_launch_count(const ml_ctx *ctx);
/* Device pointers of the resident system (NULL before ml_assemble): for benches that time kernels
   with inputs already resident. */
ml_status ml_device_system(ml_ctx *ctx, double **A_dev, int *ld, int *nrows_local, int *ncols);
/* The CUDA stream (cudaStream_t) every kernel and copy of this context is enqueued on: lets a caller bracket a
   sequence of entry-point calls with its own CUDA events (bench.py times steps on the device this way). */
ml_status ml_device_stream(ml_ctx *ctx, void **stream_out);
/* Census of the pair classes of the assembled case (the algorithmic flop count of a supersonic assembly, SURVEY 8(d):
   30 flop per culled pair, 76 + 44 e per evaluated pair with e edges in the domain of dependence): counts4[0] = pairs
   outside the domain of dependence (panel_check_dod, src/panel.f90:1732-1901), counts4[e] = evaluated pairs with e = 1..3
   edges inside.  Subsonic: everything in counts4[3]. */
ml_status ml_dod_census(ml_ctx *ctx, long long *counts4);
/* Re-run the assembly kernels only (inputs resident, no host transfers); returns device ms. */
ml_status ml_assemble_resident(ml_ctx *ctx, double *device_ms);
/* Roofline denominators measured on this device: FP64 vector pipe (register-resident DFMA loop,
   TFLOP/s) and a streaming copy (read+write GB/s).  Either pointer may be NULL. */
ml_status ml_measure_peaks(ml_ctx *ctx, double *fp64_tflops, double *hbm_gbs);
/* FP64 tensor pipe (DMMA, mma.sync.m8n8k4.f64) peak measured on this device, TFLOP/s: the roofline of the
   blocked LU's trailing update. */
ml_status ml_measure_dmma_peak(ml_ctx *ctx, double *tflops);
/* Accounting since context creation (or the last ml_reset_profile): host<->device bytes moved by the
   entry points, and -- when profiling is on -- CUDA-event time of the HBM-bound gemv kernel of the
   Krylov solvers (one event pair per launch, on the launching stream). */
typedef struct ml_profile {
    long long h2d_bytes, d2h_bytes;
    long long gemv_launches;   /* profiled launches of gemv_n_partial_kernel                       */
    long long gemv_bytes;      /* algorithmic bytes of those launches: 8*rows*cols each             */
    double gemv_ms;            /* summed device time of those launches                               */
    double assemble_ms;        /* device time of the last assembly                                   */
    double comm_ms;            /* multi-GPU: summed device time of the exchange step after those launches
                                  (ncclAllGather of the Krylov vector + compaction)                     */
} ml_profile;
ml_status ml_set_profiling(ml_ctx *ctx, int on);
ml_status ml_get_profile(ml_ctx *ctx, ml_profile *out);
ml_status ml_reset_profile(ml_ctx *ctx);
/* Rank 0 makes the 128-byte NCCL unique id that every rank passes to ml_set_communicator. */
ml_status ml_nccl_unique_id(void *out128);

#ifdef __cplusplus
}
#endif
#endif /* MACHLINE_GPU_H */
