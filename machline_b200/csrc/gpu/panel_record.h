// Packed per-image panel record, the unit the assembly kernel stages in shared memory with a bulk
// (TMA) copy.  One record per (panel, image); doubles first, then an int tail.
// Source of every field: type(panel) members used at evaluation time (src/panel.f90:38-70).
// The record stride is an EVEN number of doubles (38 / 52), so every record is 16-byte aligned and its fields are read
// with 128-bit shared-memory loads (two doubles per LDS); the strides are 304 B and 416 B = 12 and 8 banks (mod 32), so
// the 16-byte accesses of the 2, 4 or 8 different records a warp reads at once (row tiles narrower than a warp) fall
// in distinct banks.
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
    R_FLAGS = 35,   // 2 ints: [0] flags (bit0 evaluate, bit1 mirror image, bit2 has known source),
                    //         [1] subsonic: float bits of the near-edge threshold (0.05 * longest edge)^2
    R_AREA2 = 36,   // [1]   |(v2-v1) x (v3-v1)| in local scaled coordinates = twice the panel area there (subsonic only)
    R_SUB_DOUBLES = 37,  // subsonic payload
    R_SUB_STRIDE = 38,   // subsonic record stride (304 B)
    R_B = 36,       // [3]   edge parameter b            (supersonic only from here on)
    R_SB = 39,      // [3]   sqrt|b|
    R_VG = 42,      // [9]   global vertex locations of this image (DoD tests)
    R_SUP_DOUBLES = 51,  // supersonic payload
    R_SUP_STRIDE = 52    // supersonic record stride (416 B)
};

// Higher-order tables (quadratic doublets, linear sources; panel.f90:544-969): the record is the lower-order record followed
// by an extension of R_HO_EXTRA doubles (so the strides stay even: 78 / 92 doubles).  Offsets relative to the extension:
enum : int {
    R_HO_T = 0,       // [36]  T_mu (6 x 6 row-major, mu_dim x M_dim zero padded; an order-1 panel carries its 3 x 3 upper left)
    R_HO_W = 36,      // [3]   T_sigma (3 x S_dim) times the known strengths of the panel's S_dim source panels: the source
                      //       influence of the pair on I_known is -J K_inv (phi_s_sigma_space . w)  (panel.f90:2838-2849 with
                      //       panel_solver.f90:1245-1246 folded in); an order-1 panel has w = (sigma, 0, 0)
    R_HO_EXTRA = 40
};
constexpr int record_stride(bool sup, bool ho) { return (sup ? R_SUP_STRIDE : R_SUB_STRIDE) + (ho ? R_HO_EXTRA : 0); }
constexpr int record_ho_offset(bool sup) { return sup ? R_SUP_STRIDE : R_SUB_STRIDE; }

enum : int { RF_EVAL = 1, RF_MIRROR = 2, RF_SOURCE = 4 };

// ---- per-chunk scatter list (built on the host, staged next to the records) -------------------------------
// A chunk is C consecutive records of the stream.  Its list names every column of A that the chunk's
// records feed and, per column, the (record, slot) items in the reference's order of addition
// (panel_solver.f90:1445-1476 / 1656-1686: record order, slot order inside a record).
//   int  head[4]      : n_cols, n_items, flags (bit0: wake pass), spare
//   int  col[6C]      : target column; bit 31 set = first chunk of the pass that touches it (start from 0)
//   u16  beg[6C + 2]  : item range of column i = [beg[i], beg[i+1])
//   u32  item[6C]     : byte offset of the staged value, (position_in_chunk * S + slot) * R * 8 (R = tile rows, S = staged
//                       values per record and row: 3 subsonic, 4 supersonic; 6 / 7 for higher-order tables), with
//                       bit 31 set when the item is subtracted (the bottom side of a wake panel: its items 3..5 read the
//                       slots 0..2 again)
constexpr int list_max_items(int C) { return 6 * C; }
constexpr int list_bytes(int C) { return ((16 + 4 * list_max_items(C) + 2 * (list_max_items(C) + 2) + 4 * list_max_items(C)) + 15) / 16 * 16; }
constexpr unsigned ITEM_NEG = 0x80000000u;
enum : int { LF_WAKE = 1, LF_SOURCES = 2 };   // LF_SOURCES: some record of the chunk feeds a known source strength
constexpr unsigned COL_FIRST = 0x80000000u;

}  // namespace mlgpu
