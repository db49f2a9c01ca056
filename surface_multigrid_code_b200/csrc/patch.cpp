// patch.cpp -- host-side planning of the communication-avoiding patch smoother (patch.hpp).
// Integer work on sparsity patterns only; the values of the level matrices are filled on the
// device (fill lists), the static values of the transfer operators are copied.
#include "patch.hpp"

#include <algorithm>
#include <cstring>
#include <numeric>

namespace smg {

namespace {

// stored (non-padding) entries of every SELL row, in stored order
struct RowList {
  std::vector<int> ptr, col, src;
  int rows() const { return static_cast<int>(ptr.size()) - 1; }
};

RowList sell_rows(const Sell& S) {
  RowList R;
  R.ptr.assign(static_cast<size_t>(S.nrows) + 1, 0);
  for (int r = 0; r < S.nrows; r++) {
    const int s = r / kSliceRows, lane = r % kSliceRows;
    const int base = S.slice_ptr[s], w = (S.slice_ptr[s + 1] - base) / kSliceRows;
    int c = 0;
    for (int j = 0; j < w; j++)
      if (S.src[base + j * kSliceRows + lane] >= 0) c++;
    R.ptr[r + 1] = R.ptr[r] + c;
  }
  R.col.resize(R.ptr.back());
  R.src.resize(R.ptr.back());
  for (int r = 0; r < S.nrows; r++) {
    const int s = r / kSliceRows, lane = r % kSliceRows;
    const int base = S.slice_ptr[s], w = (S.slice_ptr[s + 1] - base) / kSliceRows;
    int o = R.ptr[r];
    for (int j = 0; j < w; j++) {
      const int e = base + j * kSliceRows + lane;
      if (S.src[e] >= 0) {
        R.col[o] = S.col[e];
        R.src[o] = S.src[e];
        o++;
      }
    }
  }
  return R;
}

// ---- recursive bisection along breadth-first orders ------------------------------------
// A piece is ordered breadth-first from a pseudo-peripheral row of its own and cut across that
// order, so consecutive cuts run across the longest extent of what is left: compact pieces
// of (almost) equal size.
struct Bisector {
  const RowList& G;
  std::vector<int> in_set, seen;  // stamps
  int stamp = 0;
  std::vector<int> part;
  int next_id = 0;
  explicit Bisector(const RowList& g) : G(g), in_set(g.rows(), 0), seen(g.rows(), 0), part(g.rows(), -1) {}

  // breadth-first order of `rows` (all marked in_set == sid) starting at `start`
  void bfs(const std::vector<int>& rows, int start, int sid, std::vector<int>* order) {
    order->clear();
    const int vid = ++stamp;
    size_t scan = 0;
    int s = start;
    while (order->size() < rows.size()) {
      if (s < 0) {  // another component of the piece
        while (seen[rows[scan]] == vid) scan++;
        s = rows[scan];
      }
      size_t head = order->size();
      order->push_back(s);
      seen[s] = vid;
      for (; head < order->size(); head++) {
        const int v = (*order)[head];
        for (int p = G.ptr[v]; p < G.ptr[v + 1]; p++) {
          const int w = G.col[p];
          if (in_set[w] == sid && seen[w] != vid) {
            seen[w] = vid;
            order->push_back(w);
          }
        }
      }
      s = -1;
    }
  }

  void split(std::vector<int>& rows, int m) {
    if (m <= 1 || rows.size() <= 1) {
      for (int v : rows) part[v] = next_id;
      next_id++;
      return;
    }
    const int sid = ++stamp;
    for (int v : rows) in_set[v] = sid;
    std::vector<int> order;
    bfs(rows, rows[0], sid, &order);
    const int far = order.back();
    bfs(rows, far, sid, &order);
    const int m1 = m / 2;
    const size_t n1 = std::max<size_t>(1, std::min(rows.size() - 1, rows.size() * static_cast<size_t>(m1) / m));
    std::vector<int> left(order.begin(), order.begin() + n1), right(order.begin() + n1, order.end());
    rows.clear();
    rows.shrink_to_fit();
    split(left, m1);
    split(right, m - m1);
  }
};

inline int align16(int x) { return (x + 15) & ~15; }

struct Section {
  int off = 0;
  template <class T>
  int take(size_t count) {
    const int o = off;
    off = align16(off + static_cast<int>(count * sizeof(T)));
    return o;
  }
};

struct Ctx {
  const Plan& pl;
  const LevelPlan& L;
  int l, kind, iters, C, T;
  RowList A;                 // level matrix rows (permuted numbering)
  RowList X;                 // PATCH_DOWN: rows of PT (coarse rows); PATCH_UP: rows of P (fine rows)
  std::vector<int> colour;   // per permuted row
  int t_last(int c) const { return iters > 0 ? (iters - 1) * C + c + 1 : 0; }
  int prev(int t, int c) const {  // last phase before t in which colour c is updated (0: none)
    for (int s = t - 1; s >= 1; s--)
      if ((s - 1) % C == c) return s;
    return 0;
  }
};

}  // namespace

bool build_patches(const Plan& pl, int l, int kind, int iters, int target_rows, int smem_limit, int k_cols,
                   PatchSet* out, std::string* why) {
  *out = PatchSet();
  auto fail = [&](const std::string& msg) {
    if (why) *why = msg;
    *out = PatchSet();
    return false;
  };
  const int nlev = static_cast<int>(pl.lv.size());
  if (l < 0 || l + 1 >= nlev) return fail("no coarser level");
  const LevelPlan& L = pl.lv[l];
  const LevelPlan& Cl = pl.lv[l + 1];
  const int n = L.n;
  if (n <= 0 || iters < 0) return fail("empty level");
  const int C = L.n_phases;
  if (C < 1 || C > kPatchMaxColours || iters * C > kPatchMaxPhases) return fail("too many colours / phases");
  if (static_cast<int>(L.order.phase_ptr.size()) < 1) return fail("no phases");
  Ctx cx{pl, L, l, kind, iters, C, iters * C, sell_rows(L.sellA),
         sell_rows(kind == PATCH_DOWN ? Cl.sellPT : Cl.sellP), {}};
  cx.colour.resize(n);
  for (int r = 0; r < n; r++) cx.colour[r] = L.phase[L.order.perm[r]];
  const int T = cx.T;
  const int nc = Cl.n;

  int npatch = std::max(1, (n + std::max(1, target_rows) - 1) / std::max(1, target_rows));
  for (int attempt = 0; attempt < 12; attempt++, npatch = npatch + npatch / 2 + 1) {
    npatch = std::min(npatch, n);
    // ---- partition ------------------------------------------------------------------
    Bisector bis(cx.A);
    {
      std::vector<int> all(n);
      std::iota(all.begin(), all.end(), 0);
      bis.split(all, npatch);
    }
    const int np = bis.next_id;
    const std::vector<int>& part = bis.part;
    std::vector<std::vector<int>> own(np), cown(np);
    for (int r = 0; r < n; r++) own[part[r]].push_back(r);
    if (kind == PATCH_DOWN) {
      // a coarse row belongs to the patch of the fine row that carries its largest weight
      const std::vector<double>& pv = Cl.P.val;
      for (int I = 0; I < nc; I++) {
        int best = -1;
        double bw = -1.0;
        for (int p = cx.X.ptr[I]; p < cx.X.ptr[I + 1]; p++) {
          const double w = std::abs(pv[cx.X.src[p]]);
          if (w > bw) {
            bw = w;
            best = cx.X.col[p];
          }
        }
        cown[best >= 0 ? part[best] : 0].push_back(I);  // (an empty restriction row still has to be written)
      }
    }
    // ---- per patch: needs, local numbering, blob ---------------------------------------
    PatchSet ps;
    ps.level = l;
    ps.kind = kind;
    ps.iters = iters;
    ps.n_patches = np;
    ps.off.assign(static_cast<size_t>(np) + 1, 0);
    std::vector<int> need(n, -1), local(n, -1), rpos(n, -1), touched;
    std::vector<char> in_R(n, 0);
    std::vector<std::vector<int>> bucket(static_cast<size_t>(T) + 1);
    bool fits = true;
    for (int pid = 0; pid < np && fits; pid++) {
      touched.clear();
      for (auto& b : bucket) b.clear();
      auto set_need = [&](int v, int t) {
        if (need[v] < t) {
          if (need[v] < 0) touched.push_back(v);
          need[v] = t;
          bucket[t].push_back(v);
        }
      };
      std::vector<int> R;
      for (int v : own[pid]) set_need(v, cx.t_last(cx.colour[v]));
      if (kind == PATCH_DOWN) {
        for (int I : cown[pid])
          for (int p = cx.X.ptr[I]; p < cx.X.ptr[I + 1]; p++) {
            const int f = cx.X.col[p];
            if (!in_R[f]) {
              in_R[f] = 1;
              R.push_back(f);
            }
          }
        std::sort(R.begin(), R.end());
        for (int f : R)  // r(f) reads the final u of f and of every neighbour
          for (int p = cx.A.ptr[f]; p < cx.A.ptr[f + 1]; p++) set_need(cx.A.col[p], cx.t_last(cx.colour[cx.A.col[p]]));
        for (int f : R) set_need(f, cx.t_last(cx.colour[f]));
      }
      for (int t = T; t >= 1; t--)
        for (size_t q = 0; q < bucket[t].size(); q++) {
          const int v = bucket[t][q];
          if (need[v] != t) continue;
          for (int p = cx.A.ptr[v]; p < cx.A.ptr[v + 1]; p++) {
            const int w = cx.A.col[p];
            if (w != v) set_need(w, cx.prev(t, cx.colour[w]));
          }
        }
      // local numbering
      std::vector<int> rows_b, rows_ro;
      for (int v : touched) (need[v] >= 1 || in_R[v] ? rows_b : rows_ro).push_back(v);
      std::sort(rows_b.begin(), rows_b.end(), [&](int a, int b) {
        if (cx.colour[a] != cx.colour[b]) return cx.colour[a] < cx.colour[b];
        if (need[a] != need[b]) return need[a] > need[b];
        return a < b;
      });
      std::sort(rows_ro.begin(), rows_ro.end());
      const int n_b = static_cast<int>(rows_b.size()), n_loc = n_b + static_cast<int>(rows_ro.size());
      const int n_R = static_cast<int>(R.size()), n_C = static_cast<int>(cown[pid].size());
      const int n_own = static_cast<int>(own[pid].size());
      for (int i = 0; i < n_b; i++) local[rows_b[i]] = i;
      for (int i = n_b; i < n_loc; i++) local[rows_ro[i - n_b]] = i;
      for (int i = 0; i < n_R; i++) rpos[R[i]] = i;

      PatchHeader H;
      std::memset(&H, 0, sizeof(H));
      H.n_loc = n_loc;
      H.n_b = n_b;
      H.n_own = n_own;
      H.n_R = n_R;
      H.n_C = n_C;
      H.T = T;
      H.C = C;
      H.kind = kind;
      // colour groups and their ELL widths (the diagonal entry is not stored: the sweep skips it)
      std::vector<int> wrow(n_b, 0);
      {
        int i = 0;
        int ent = 0;
        for (int g = 0; g < C; g++) {
          H.grp_row[g] = i;
          H.grp_ent[g] = ent;
          int w = 0;
          const int i0 = i;
          while (i < n_b && cx.colour[rows_b[i]] == g) {
            const int v = rows_b[i];
            if (need[v] >= 1) {
              int c = 0;
              for (int p = cx.A.ptr[v]; p < cx.A.ptr[v + 1]; p++) c += cx.A.col[p] != v;
              wrow[i] = c;
              ps.sum_entries += c;
              w = std::max(w, c);
            }
            i++;
          }
          H.grp_w[g] = w;
          ent += w * (i - i0);
        }
        for (int g = C; g <= kPatchMaxColours; g++) H.grp_row[g] = n_b;
        int passes = 0;
        for (int t = 1; t <= T; t++) {
          const int g = (t - 1) % C;
          int c = 0;
          for (int i2 = H.grp_row[g]; i2 < H.grp_row[g + 1] && need[rows_b[i2]] >= t; i2++) c++;
          H.n_active[t - 1] = static_cast<short>(c);
          ps.max_active = std::max(ps.max_active, c);
          ps.sum_updates += c;
          passes += (c + 255) / 256;
        }
        ps.max_passes = std::max(ps.max_passes, passes);
        int W_r = 0, W_pt = 0, W_p = 0;
        for (int f : R) W_r = std::max(W_r, cx.A.ptr[f + 1] - cx.A.ptr[f]);
        if (kind == PATCH_DOWN)
          for (int I : cown[pid]) W_pt = std::max(W_pt, cx.X.ptr[I + 1] - cx.X.ptr[I]);
        if (kind == PATCH_UP)
          for (int v : touched) W_p = std::max(W_p, cx.X.ptr[v + 1] - cx.X.ptr[v]);
        H.W_r = W_r;
        H.W_pt = W_pt;
        H.W_p = W_p;
        for (int g = 0; g < C; g++) W_r = std::max(W_r, H.grp_w[g]);
        if (std::max(W_r, std::max(W_pt, W_p)) > 255) return fail("a row has more than 255 entries");
        ps.max_width = std::max(ps.max_width, W_r);
        W_r = H.W_r;
        // sections: doubles first, then 4-, 2-, 1-byte arrays
        Section sec;
        sec.off = static_cast<int>(sizeof(PatchHeader));
        H.o_val = sec.take<double>(ent);
        H.o_diag = sec.take<double>(n_b);
        H.o_rval = sec.take<double>(static_cast<size_t>(W_r) * n_R);
        H.o_ptval = sec.take<double>(static_cast<size_t>(W_pt) * n_C);
        H.o_pval = sec.take<double>(static_cast<size_t>(W_p) * n_loc);
        H.o_gid = sec.take<int>(n_loc);
        H.o_cgid = sec.take<int>(n_C);
        H.o_pcol = sec.take<int>(static_cast<size_t>(W_p) * n_loc);
        H.o_col = sec.take<unsigned short>(ent);
        H.o_own = sec.take<unsigned short>(n_own);
        H.o_ridx = sec.take<unsigned short>(n_R);
        H.o_rcol = sec.take<unsigned short>(static_cast<size_t>(W_r) * n_R);
        H.o_ptcol = sec.take<unsigned short>(static_cast<size_t>(W_pt) * n_C);
        H.o_w = sec.take<unsigned char>(n_b);
        H.o_rw = sec.take<unsigned char>(n_R);
        H.o_ptw = sec.take<unsigned char>(n_C);
        H.o_pw = sec.take<unsigned char>(n_loc);
        H.blob_bytes = sec.off;
      }
      const int64_t smem1 = static_cast<int64_t>(H.blob_bytes) + 8ll * std::max(1, k_cols) * (n_loc + n_b + n_R);
      if (n_loc > 65535 || smem1 > smem_limit) {
        fits = false;
      } else {
        const long long base = ps.off[pid];
        ps.off[pid + 1] = base + H.blob_bytes;
        ps.blob.resize(static_cast<size_t>(base + H.blob_bytes), 0);
        unsigned char* B = ps.blob.data() + base;
        std::memcpy(B, &H, sizeof(H));
        auto I32 = [&](int o) { return reinterpret_cast<int*>(B + o); };
        auto U16 = [&](int o) { return reinterpret_cast<unsigned short*>(B + o); };
        auto F64 = [&](int o) { return reinterpret_cast<double*>(B + o); };
        auto slot = [&](int o, size_t i) { return static_cast<int>((base + o) / 8 + static_cast<long long>(i)); };
        for (int i = 0; i < n_b; i++) I32(H.o_gid)[i] = rows_b[i];
        for (int i = n_b; i < n_loc; i++) I32(H.o_gid)[i] = rows_ro[i - n_b];
        for (int i = 0; i < n_own; i++) U16(H.o_own)[i] = static_cast<unsigned short>(local[own[pid][i]]);
        for (int g = 0; g < C; g++) {
          const int g0 = H.grp_row[g], gn = H.grp_row[g + 1] - g0;
          for (int r = 0; r < gn; r++) {
            const int i = g0 + r, v = rows_b[i];
            B[H.o_w + i] = static_cast<unsigned char>(wrow[i]);
            ps.fill_dst.push_back(slot(H.o_diag, i));
            ps.fill_src.push_back(L.diag_pos[L.order.perm[v]]);
            if (need[v] < 1) continue;
            int j = 0;
            for (int p = cx.A.ptr[v]; p < cx.A.ptr[v + 1]; p++) {
              if (cx.A.col[p] == v) continue;
              const size_t e = static_cast<size_t>(H.grp_ent[g]) + static_cast<size_t>(j) * gn + r;
              U16(H.o_col)[e] = static_cast<unsigned short>(local[cx.A.col[p]]);
              ps.fill_dst.push_back(slot(H.o_val, e));
              ps.fill_src.push_back(cx.A.src[p]);  // column v of the CSC read as row v (mg_VCycle.cpp:152)
              j++;
            }
          }
        }
        for (int r = 0; r < n_R; r++) {
          const int f = R[r];
          U16(H.o_ridx)[r] = static_cast<unsigned short>(local[f]);
          B[H.o_rw + r] = static_cast<unsigned char>(cx.A.ptr[f + 1] - cx.A.ptr[f]);
          int j = 0;
          for (int p = cx.A.ptr[f]; p < cx.A.ptr[f + 1]; p++, j++) {
            const size_t e = static_cast<size_t>(j) * n_R + r;
            U16(H.o_rcol)[e] = static_cast<unsigned short>(local[cx.A.col[p]]);
            ps.fill_dst.push_back(slot(H.o_rval, e));
            ps.fill_src.push_back(L.tmap[cx.A.src[p]]);  // the true row f (what A * u reads)
          }
        }
        if (kind == PATCH_DOWN)
          for (int r = 0; r < n_C; r++) {
            const int I = cown[pid][r];
            I32(H.o_cgid)[r] = I;
            B[H.o_ptw + r] = static_cast<unsigned char>(cx.X.ptr[I + 1] - cx.X.ptr[I]);
            int j = 0;
            for (int p = cx.X.ptr[I]; p < cx.X.ptr[I + 1]; p++, j++) {
              const size_t e = static_cast<size_t>(j) * n_C + r;
              U16(H.o_ptcol)[e] = static_cast<unsigned short>(rpos[cx.X.col[p]]);
              F64(H.o_ptval)[e] = Cl.P.val[cx.X.src[p]];
            }
          }
        if (kind == PATCH_UP)
          for (int i = 0; i < n_loc; i++) {
            const int v = I32(H.o_gid)[i];
            B[H.o_pw + i] = static_cast<unsigned char>(cx.X.ptr[v + 1] - cx.X.ptr[v]);
            int j = 0;
            for (int p = cx.X.ptr[v]; p < cx.X.ptr[v + 1]; p++, j++) {
              const size_t e = static_cast<size_t>(j) * n_loc + i;
              I32(H.o_pcol)[e] = cx.X.col[p];
              F64(H.o_pval)[e] = Cl.PT.val[cx.X.src[p]];
            }
          }
        ps.max_blob_bytes = std::max(ps.max_blob_bytes, H.blob_bytes);
        ps.max_vec_doubles = std::max(ps.max_vec_doubles, n_loc + n_b + n_R);
        ps.sum_own += n_own;
        ps.sum_loc += n_loc;
        ps.sum_b += n_b;
      }
      for (int v : touched) {
        need[v] = -1;
        local[v] = -1;
      }
      for (int f : R) {
        in_R[f] = 0;
        rpos[f] = -1;
      }
    }
    if (fits) {
      *out = std::move(ps);
      return true;
    }
    if (npatch >= n) break;
  }
  return fail("a patch does not fit into shared memory");
}

std::string verify_patches(const Plan& pl, const PatchSet& ps) {
  if (ps.empty()) return "empty patch set";
  const int l = ps.level;
  const LevelPlan& L = pl.lv[l];
  const LevelPlan& Cl = pl.lv[l + 1];
  const int n = L.n, C = L.n_phases, T = ps.iters * C;
  const RowList A = sell_rows(L.sellA);
  const RowList X = sell_rows(ps.kind == PATCH_DOWN ? Cl.sellPT : Cl.sellP);
  std::vector<int> colour(n);
  for (int r = 0; r < n; r++) colour[r] = L.phase[L.order.perm[r]];
  auto t_last = [&](int c) { return ps.iters > 0 ? (ps.iters - 1) * C + c + 1 : 0; };
  auto prev = [&](int t, int c) {
    for (int s = t - 1; s >= 1; s--)
      if ((s - 1) % C == c) return s;
    return 0;
  };
  std::vector<int> owned(n, 0), cowned(Cl.n, 0);
  // fill lists by destination slot
  std::vector<int> fill_of(ps.blob.size() / 8, -1);
  for (size_t i = 0; i < ps.fill_dst.size(); i++) {
    if (ps.fill_dst[i] < 0 || static_cast<size_t>(ps.fill_dst[i]) >= fill_of.size()) return "fill slot out of range";
    if (fill_of[ps.fill_dst[i]] >= 0) return "fill slot written twice";
    fill_of[ps.fill_dst[i]] = ps.fill_src[i];
  }
  for (int pid = 0; pid < ps.n_patches; pid++) {
    const std::string at = " (patch " + std::to_string(pid) + ")";
    const long long base = ps.off[pid];
    const unsigned char* B = ps.blob.data() + base;
    PatchHeader H;
    std::memcpy(&H, B, sizeof(H));
    if (base + H.blob_bytes != ps.off[pid + 1] || base % 16 != 0) return "blob offsets" + at;
    if (H.T != T || H.C != C || H.kind != ps.kind) return "header" + at;
    auto I32 = [&](int o) { return reinterpret_cast<const int*>(B + o); };
    auto U16 = [&](int o) { return reinterpret_cast<const unsigned short*>(B + o); };
    auto F64 = [&](int o) { return reinterpret_cast<const double*>(B + o); };
    auto src_of = [&](int o, size_t i) { return fill_of[static_cast<size_t>((base + o) / 8) + i]; };
    const int* gid = I32(H.o_gid);
    for (int i = 0; i < H.n_loc; i++)
      if (gid[i] < 0 || gid[i] >= n) return "row id" + at;
    std::vector<int> ver(H.n_loc, 0);
    for (int t = 1; t <= T; t++) {
      const int g = (t - 1) % C;
      const int g0 = H.grp_row[g], gn = H.grp_row[g + 1] - g0;
      if (H.n_active[t - 1] > gn) return "active rows" + at;
      for (int r = 0; r < H.n_active[t - 1]; r++) {
        const int i = g0 + r, v = gid[i];
        if (colour[v] != g) return "colour group" + at;
        // the stored row equals the SELL row without its diagonal, in order
        int j = 0;
        for (int p = A.ptr[v]; p < A.ptr[v + 1]; p++) {
          if (A.col[p] == v) continue;
          if (j >= B[H.o_w + i] || j >= H.grp_w[g]) return "row width" + at;
          const size_t e = static_cast<size_t>(H.grp_ent[g]) + static_cast<size_t>(j) * gn + r;
          const int c = U16(H.o_col)[e];
          if (c >= H.n_loc || gid[c] != A.col[p]) return "column" + at;
          if (src_of(H.o_val, e) != A.src[p]) return "value source" + at;
          if (ver[c] != prev(t, colour[A.col[p]])) return "a neighbour has the wrong version in phase " + std::to_string(t) + at;
          j++;
        }
        if (j != B[H.o_w + i]) return "row width" + at;
        if (src_of(H.o_diag, i) != L.diag_pos[L.order.perm[v]]) return "diagonal source" + at;
      }
      for (int r = 0; r < H.n_active[t - 1]; r++) ver[g0 + r] = t;
    }
    for (int i = 0; i < H.n_own; i++) {
      const int li = U16(H.o_own)[i];
      if (li >= H.n_loc) return "owned row index" + at;
      const int v = gid[li];
      if (ver[li] != t_last(colour[v])) return "an owned row is not final" + at;
      owned[v]++;
    }
    if (ps.kind == PATCH_DOWN) {
      for (int r = 0; r < H.n_R; r++) {
        const int li = U16(H.o_ridx)[r];
        if (li >= H.n_b) return "residual row index" + at;
        const int f = gid[li];
        if (B[H.o_rw + r] != A.ptr[f + 1] - A.ptr[f]) return "residual row width" + at;
        int j = 0;
        for (int p = A.ptr[f]; p < A.ptr[f + 1]; p++, j++) {
          const size_t e = static_cast<size_t>(j) * H.n_R + r;
          const int c = U16(H.o_rcol)[e];
          if (c >= H.n_loc || gid[c] != A.col[p]) return "residual column" + at;
          if (ver[c] != t_last(colour[A.col[p]])) return "the residual reads a row that is not final" + at;
          if (src_of(H.o_rval, e) != L.tmap[A.src[p]]) return "residual value source" + at;
        }
      }
      for (int r = 0; r < H.n_C; r++) {
        const int I = I32(H.o_cgid)[r];
        if (I < 0 || I >= Cl.n) return "coarse row id" + at;
        cowned[I]++;
        if (B[H.o_ptw + r] != X.ptr[I + 1] - X.ptr[I]) return "restriction row width" + at;
        int j = 0;
        for (int p = X.ptr[I]; p < X.ptr[I + 1]; p++, j++) {
          const size_t e = static_cast<size_t>(j) * H.n_C + r;
          const int rr = U16(H.o_ptcol)[e];
          if (rr >= H.n_R || gid[U16(H.o_ridx)[rr]] != X.col[p]) return "restriction column" + at;
          if (F64(H.o_ptval)[e] != Cl.P.val[X.src[p]]) return "restriction value" + at;
        }
      }
    } else {
      for (int i = 0; i < H.n_loc; i++) {
        const int v = gid[i];
        if (B[H.o_pw + i] != X.ptr[v + 1] - X.ptr[v]) return "prolongation row width" + at;
        int j = 0;
        for (int p = X.ptr[v]; p < X.ptr[v + 1]; p++, j++) {
          const size_t e = static_cast<size_t>(j) * H.n_loc + i;
          if (I32(H.o_pcol)[e] != X.col[p]) return "prolongation column" + at;
          if (F64(H.o_pval)[e] != Cl.PT.val[X.src[p]]) return "prolongation value" + at;
        }
      }
    }
  }
  for (int v = 0; v < n; v++)
    if (owned[v] != 1) return "row " + std::to_string(v) + " is owned " + std::to_string(owned[v]) + " times";
  if (ps.kind == PATCH_DOWN)
    for (int I = 0; I < Cl.n; I++)
      if (cowned[I] != 1) return "coarse row " + std::to_string(I) + " is owned " + std::to_string(cowned[I]) + " times";
  return "";
}

}  // namespace smg
