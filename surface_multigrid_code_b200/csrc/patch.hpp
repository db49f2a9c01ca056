// patch.hpp -- communication-avoiding ("patch") schedule of the multicolour smoother on the
// small levels of the hierarchy (host-side planning, pure C++17; kernels: patch_kernels.cu).
//
// Why: on a level with fewer rows than the GPU has threads, one colour phase of Gauss-Seidel is
// a ~2 us kernel boundary that moves a few hundred KB; a V(2,2) cycle has 19 such dependent
// steps per level.  Here the rows of a level are cut into connected patches, one CTA per
// patch, and a CTA runs ALL colour phases of a relax call (mg_VCycle.cpp:36 / :56) on its
// patch in shared memory.  Rows outside the patch whose intermediate values the patch needs
// are recomputed redundantly (a halo whose depth shrinks phase by phase), so no CTA ever
// waits for another one: the results are bit-identical to the phase-by-phase kernels (same
// per-row arithmetic in the same entry order), with one launch instead of 2 x colours.
// The residual and the restriction that follow the pre-smoothing (mg_VCycle.cpp:41-47) and
// the prolongation + correction that precede the post-smoothing (:52-53) run inside the same
// launches, so a level costs two launches per V-cycle.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "plan.hpp"

namespace smg {

constexpr int kPatchMaxColours = 12;
constexpr int kPatchMaxPhases = 32;
enum PatchKind { PATCH_DOWN = 0, PATCH_UP = 1 };

// First bytes of every patch blob; the same POD on host and device.  Local rows are numbered
// [0, n_b): rows whose right-hand side is needed (rows that are updated in some phase, and
// the residual rows), grouped by colour and, inside a colour, sorted by the last phase they
// are updated in (descending), so the rows active in a phase are a prefix of their group;
// [n_b, n_loc): rows that are only read.  Offsets are in bytes from the start of the blob.
struct PatchHeader {
  int n_loc, n_b, n_own, n_R, n_C;
  int T, C;              // phases of the relax call (sweeps x colours), colours
  int W_r, W_pt, W_p;    // ELL widths of the residual rows / restriction rows / prolongation rows
  int kind, blob_bytes;
  int o_gid;             // int32[n_loc]: row in the level's (permuted) numbering
  int o_w;               // uint8[n_b]: stored entries of the Gauss-Seidel row
  int o_col, o_val;      // uint16 / double, per colour group column-major: [grp_ent[g] + j * rows(g) + r]
  int o_diag;            // double[n_b]
  int o_own;             // uint16[n_own]: local index of the rows this patch owns (writes back)
  // PATCH_DOWN: residual of the rows the patch's coarse rows restrict from, then the restriction
  int o_ridx, o_rw, o_rcol, o_rval;      // uint16[n_R] local row, uint8[n_R], uint16 / double [j * n_R + r]
  int o_cgid, o_ptw, o_ptcol, o_ptval;   // int32[n_C] coarse row, uint8[n_C], uint16 (index into the residual rows) / double
  // PATCH_UP: u = u + P uc for every local row before the sweeps
  int o_pw, o_pcol, o_pval;              // uint8[n_loc], int32 (coarse row) / double [j * n_loc + i]
  int grp_row[kPatchMaxColours + 1];
  int grp_ent[kPatchMaxColours];
  int grp_w[kPatchMaxColours];
  short n_active[kPatchMaxPhases];       // rows of phase t's colour group updated in phase t (t = 0 .. T-1)
  int pad[2];
};
static_assert(sizeof(PatchHeader) % 16 == 0, "patch blobs are copied with 16-byte granularity");

struct PatchSet {
  int level = -1, kind = PATCH_DOWN, iters = 0, n_patches = 0;
  std::vector<long long> off;       // n_patches + 1 byte offsets into blob (multiples of 16)
  std::vector<unsigned char> blob;  // headers, index arrays, transfer-operator values; matrix values zero
  // numeric fill: double slot fill_dst[i] of the blob (index in doubles) <- a_val[fill_src[i]] of the level
  std::vector<int> fill_dst, fill_src;
  int max_blob_bytes = 0;
  int max_vec_doubles = 0;          // most (n_loc + n_b + n_R) of any patch: shared-memory vectors per column
  int max_active = 0;               // most rows updated in one phase by one patch
  int max_passes = 0;               // most sum_t ceil(rows updated in phase t / 256) of any patch
  int max_width = 0;                // widest stored row
  int64_t sum_own = 0, sum_loc = 0, sum_b = 0, sum_updates = 0, sum_entries = 0;  // statistics
  bool empty() const { return n_patches == 0; }
};

// Cut level `l` of the plan into about n / target_rows patches and lay out one relax call of
// `iters` sweeps (+ residual / restriction for PATCH_DOWN, prolongation for PATCH_UP).
// smem_limit: bytes of shared memory one patch may need with k_cols right-hand-side columns;
// the patch count is raised until every patch fits.  Returns false (and leaves `out` empty) when
// the level cannot be laid out (too many colours / phases, a patch that never fits).
bool build_patches(const Plan& pl, int l, int kind, int iters, int target_rows, int smem_limit, int k_cols,
                   PatchSet* out, std::string* why);

// Symbolic check of a patch set against the level's matrices: every update reads neighbour
// values of exactly the version the phase-by-phase schedule would (no floating point
// involved), rows / coarse rows are owned exactly once, entry lists equal the SELL rows.
// Returns "" or a description of the first violation.
std::string verify_patches(const Plan& pl, const PatchSet& ps);

}  // namespace smg
