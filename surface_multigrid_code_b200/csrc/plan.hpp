// plan.hpp -- host-side index/topology planning for libsmg (pure C++17, no CUDA).
//
// Everything here is integer work on sparsity patterns: the pieces of
// min_quad_with_fixed_mg_precompute whose outputs must be bit-identical to the
// reference (unknown = setdiff, igl::slice patterns, column pruning, the patterns
// of the conservative sparse products) plus the data-layout decisions that are
// ours (smoother phases, row permutation, SELL-32 slices).  Floating-point work on
// the hot path lives in kernels.cu; the only values touched here are the entries of
// the prolongation matrices, which are copied / sliced, never combined.
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

namespace smg {

struct Csc {
  int rows = 0, cols = 0;
  std::vector<int> colptr;  // cols + 1
  std::vector<int> rowidx;  // nnz
  std::vector<double> val;  // nnz or empty (pattern only)
  int nnz() const { return colptr.empty() ? 0 : colptr.back(); }
};

// igl::setdiff(0..n-1, known): ascending complement (setdiff.cpp:19-75)
std::vector<int> setdiff_range(int n, const int* known, int nknown);

// igl::slice(X,R,C,Y) for sparse X (slice.cpp:13-77). R / C == nullptr: all rows /
// columns in order.  src[p] = index of the X entry that produced Y entry p.
Csc slice(const Csc& X, const int* R, int nr, const int* C, int nc, std::vector<int>* src);

// Y = X^T with explicit zeros kept; map[p] = position in Y of X entry p.
Csc transpose(const Csc& X, std::vector<int>* map);

// pattern of the conservative product L*R (Eigen ConservativeSparseSparseProduct):
// structural, sorted rows per column.
Csc spgemm_pattern(const Csc& L, const Csc& R);

// per-entry column index of a CSC pattern
std::vector<int> entry_columns(const Csc& X);

// ---- smoother schedule -----------------------------------------------------
// Pattern must be structurally symmetric; column i is read as row i
// (mg_VCycle.cpp:152).
// wavefront: level[i] = 1 + max(level[j] : j < i, A(j,i) stored), 0 if none.
std::vector<int> wavefront_levels(const Csc& A, int* n_levels);
// greedy distance-1 colouring, visiting rows in `order` (rank -> row), followed by
// a balancing pass; returns colour per row.
std::vector<int> greedy_colours(const Csc& A, const std::vector<int>& order, int* n_colours);
// breadth-first (Cuthill-McKee style) visiting order: order[rank] = row
std::vector<int> bfs_order(const Csc& A);

// ---- SELL-32 layout ----------------------------------------------------------
constexpr int kSliceRows = 32;
constexpr int kBlockRows = 256;  // rows per CTA of the hot kernels
struct Sell {
  int nrows = 0;                // rows (permuted numbering)
  int nslices = 0;
  std::vector<int> slice_ptr;   // nslices + 1, offsets in entries (multiples of 32)
  std::vector<int> col;         // padded entries: column (permuted numbering)
  std::vector<int> src;         // padded entries: source CSC entry, -1 = padding
  int64_t padded() const { return slice_ptr.empty() ? 0 : slice_ptr.back(); }
};
// Rows of the SELL matrix are the COLUMNS of `X` (column c of the CSC is stored as
// row c: this is how the reference's smoother reads A, and it turns the CSC of P /
// PT into the row-wise storage of PT / P).  Row r of the result is X column
// row_perm[r]; entry order inside a row = CSC order (ascending original index, the
// reference's accumulation order); stored column index = col_iperm[X.rowidx].
// sort_cols: order the entries of a row by stored (permuted) column instead, which
// makes the gathers of neighbouring lanes fall into the same sectors (fast mode only:
// it changes the floating-point summation order).
Sell build_sell(const Csc& X, const std::vector<int>& row_perm, const std::vector<int>& col_iperm,
                bool sort_cols = false);

struct RowOrder {
  std::vector<int> perm;       // new -> old
  std::vector<int> iperm;      // old -> new
  std::vector<int> phase_ptr;  // n_phases + 1, row offsets in the new numbering
};
// rows sorted by (phase, rank); inside sigma-aligned windows of a phase, by
// descending row length (stable).
RowOrder make_row_order(const Csc& A, const std::vector<int>& phase, int n_phases,
                        const std::vector<int>& rank, int sigma);

// ---- row partition across ranks (multi-GPU) --------------------------------------
// One halo-exchange pattern of a level: idx[src * world + dst] = rows (permuted numbering of
// the level, ascending) whose values rank `src` owns and rank `dst` reads.  The same list is
// the sender's gather list and the receiver's scatter list.
struct Exchange {
  std::vector<std::vector<int>> idx;
  bool empty() const {
    for (const auto& v : idx)
      if (!v.empty()) return false;
    return true;
  }
  size_t max_count() const {
    size_t m = 0;
    for (const auto& v : idx) m = std::max(m, v.size());
    return m;
  }
};
enum LevelLayout {
  LAYOUT_PLAIN = 0,       // rows ordered (phase, locality rank); every rank holds and computes all rows
  LAYOUT_PARTITIONED = 1, // rows ordered (part, phase, rank); rank r computes part r only
  LAYOUT_SPLIT = 2        // replicated level directly below a partitioned one: rows ordered
                          // (phase, part, rank); part r's rows are what rank r restricts into
};

// ---- whole-hierarchy plan ----------------------------------------------------
struct LevelPlan {
  int n = 0;
  Csc A;                      // pattern of mg[lv].A (values live on the device)
  std::vector<int> a_col;     // column of every entry of A
  std::vector<int> tmap;      // entry (r,c) -> position of entry (c,r)
  std::vector<int> diag_pos;  // position of A(i,i) per column, -1 if absent
  // Compute pattern: the entries of A that can be non-zero.  The reference keeps the
  // explicit zeros of P (get_prolong.cpp:45-56) and Eigen's conservative products
  // propagate them into structural entries of A_l whose value is exactly 0; they are
  // part of the parity surface (A above) but contribute nothing to A*x or to a
  // Gauss-Seidel sum, so the SELL matrix the kernels stream omits them.
  Csc Alive;                  // pattern only, subset of A
  std::vector<int> live_src;  // Alive entry -> entry of A
  int n_phases = 0;
  std::vector<int> phase;     // per row (reference numbering)
  RowOrder order;
  Sell sellA;
  // lv >= 1 (operators between level lv-1 (fine) and lv (coarse)):
  Csc P, PT;                  // with values, as the reference leaves mg[lv].P / .PT
  bool pruned = false;
  std::vector<int> keep;
  Sell sellP;                 // rows: fine level lv-1; y = P x
  Sell sellPT;                // rows: coarse level lv; y = PT x
  Csc T1;                     // pattern of PT * A_{lv-1}
  std::vector<int> t1_col;
  // multi-GPU (PlanOptions::world > 1).  With nparts > 1 order.phase_ptr has
  // nparts * n_phases + 1 entries: group (part, phase) for LAYOUT_PARTITIONED, (phase, part)
  // for LAYOUT_SPLIT.
  int layout = LAYOUT_PLAIN;
  int nparts = 1;
  std::vector<int> part;      // owner of every row (reference numbering)
  Exchange halo_u;            // u read by the A rows of another part (PARTITIONED)
  std::vector<Exchange> halo_u_phase;  // the same, split by the phase of the row sent
  Exchange halo_r;            // r read by the PT rows (level lv+1) of another part (PARTITIONED)
  Exchange halo_pu;           // this level's u read by prolongation rows (level lv-1) of another
                              // part (PARTITIONED below PARTITIONED)
  Exchange gather_all;        // every row of part src -> every other rank (PARTITIONED, SPLIT)
  // rows of this rank's phases / of a whole part, in the permuted numbering
  int group(int part_id, int phase_id) const {
    return layout == LAYOUT_PARTITIONED ? part_id * n_phases + phase_id
                                        : (layout == LAYOUT_SPLIT ? phase_id * nparts + part_id : phase_id);
  }
};

struct PlanOptions {
  int smoother = 1;  // 0 wavefront, 1 multicolour
  int locality_reorder = 1;
  int sigma = 256;
  int sort_cols = -1;  // -1: on for multicolour, off for wavefront (bit-parity order)
  // multi-GPU: number of ranks the fine levels are partitioned over, and how many levels
  // (from level 0) are partitioned; < 0: automatic (always level 0, never the coarsest; below it
  // levels of >= 500 000 rows with >= 50 000 rows per rank, or, with dist_min_rows > 0, levels
  // with at least dist_min_rows rows per rank)
  int world = 1;
  int dist_levels = -1;
  int dist_min_rows = 0;
};

struct Plan {
  int n = 0;  // size of the caller's A
  bool has_fixed = false;
  std::vector<int> known, unknown;
  Csc LHS;                    // pattern
  std::vector<int> lhs_src;   // LHS entry -> entry of the caller's A
  Csc Auk;                    // pattern, columns in caller order of `known`
  std::vector<int> auk_src;
  std::vector<LevelPlan> lv;
  int world = 1;
  int dist_levels = 0;  // levels 0 .. dist_levels-1 are PARTITIONED, level dist_levels is SPLIT
  std::string error;
};

// Full planning pass. P_full[l-1] (l = 1..nlev-1) are the prolongations handed to
// smg_set_hierarchy. Returns 0 or an smg_status code (error text in plan.error).
int build_plan(const Csc& A, const int* known, int nknown /* <0: free variant */,
               const std::vector<Csc>& P_full, const PlanOptions& opt, Plan* plan);

}  // namespace smg
