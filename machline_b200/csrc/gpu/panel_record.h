// Packed per-image panel record, the unit the assembly kernel stages in shared memory with a bulk
// (TMA) copy.  One record per (panel, image); doubles first, then an int tail, 16-byte multiple.
// Source of every field: type(panel) members used at evaluation time (src/panel.f90:38-70) after
// the index resolution of panel_solver_update_system_row (src/panel_solver.f90:1203-1287).
#pragma once

namespace mlgpu {

// offsets in units of doubles
enum : int {
    R_CENTR = 0,    // [3]   centr / centr_mir
    R_A = 3,        // [9]   A_g_to_ls (row-major)
    R_VLS = 12,     // [6]   vertices_ls (vertex k: xi, eta)
    R_NH = 18,      // [6]   n_hat_ls   (edge k: xi, eta)
    R_T = 24,       // [9]   T_mu (row-major, mu_dim x M_dim)
    R_J = 33,       // [1]   J
    R_SIGMA = 34,   // [1]   known source strength of the panel this record's source feeds (0 if none)
    R_COLS = 35,    // 6 ints (3 doubles): permuted column of doublet slot k (0..2 add, 3..5 subtract), -1 unused
    R_FLAGS = 38,   // 2 ints: [0] flags (bit0 evaluate, bit1 mirror image, bit2 has known source), [1] spare
    R_SUB_DOUBLES = 40,  // subsonic record length (320 B)
    R_B = 40,       // [3]   edge parameter b            (supersonic only from here on)
    R_SB = 43,      // [3]   sqrt|b|
    R_VG = 46,      // [9]   global vertex locations of this image (DoD tests)
    R_SUP_DOUBLES = 56   // supersonic record length (448 B)
};

enum : int { RF_EVAL = 1, RF_MIRROR = 2, RF_SOURCE = 4 };

}  // namespace mlgpu
