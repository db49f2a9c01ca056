// plan.cpp -- host-side index/topology planning (see plan.hpp).
#include "plan.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <numeric>
#include <thread>

#include "../../include/smg.h"

namespace smg {

namespace {
// SMG_PLAN_TIMING=1: stage times of build_plan on stderr
struct StageTimer {
  bool on;
  std::chrono::steady_clock::time_point t;
  StageTimer() : on(std::getenv("SMG_PLAN_TIMING") != nullptr), t(std::chrono::steady_clock::now()) {}
  void lap(const char* what) {
    if (!on) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "plan: %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
    t = now;
  }
};

// fn(begin, end) over [0, n) in contiguous chunks on up to 8 host threads (SMG_PLAN_THREADS);
// every use below writes disjoint output ranges, so results do not depend on the thread count
template <class Fn>
void parallel_chunks(int64_t n, int64_t min_per_thread, Fn fn) {
  static const int max_threads = [] {
    int t = static_cast<int>(std::thread::hardware_concurrency());
    if (const char* e = std::getenv("SMG_PLAN_THREADS")) t = std::atoi(e);
    return std::max(1, std::min(t, 8));
  }();
  const int nt = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(max_threads, n / std::max<int64_t>(1, min_per_thread))));
  if (nt <= 1) {
    fn(static_cast<int64_t>(0), n);
    return;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < nt; t++) th.emplace_back([=] { fn(n * t / nt, n * (t + 1) / nt); });
  for (auto& x : th) x.join();
}
}  // namespace

std::vector<int> setdiff_range(int n, const int* known, int nknown) {
  std::vector<char> mark(static_cast<size_t>(n) + 1, 0);
  for (int i = 0; i < nknown; i++)
    if (known[i] >= 0 && known[i] < n) mark[known[i]] = 1;
  std::vector<int> out;
  out.reserve(n);
  for (int i = 0; i < n; i++)
    if (!mark[i]) out.push_back(i);
  return out;
}

Csc slice(const Csc& X, const int* R, int nr, const int* C, int nc, std::vector<int>* src) {
  const int ym = R ? nr : X.rows, yn = C ? nc : X.cols;
  const bool with_val = !X.val.empty();
  // bucket the output rows that read every input row (stable: ascending output index)
  std::vector<int> rptr(static_cast<size_t>(X.rows) + 1, 0), ridx(ym);
  if (R) {
    for (int i = 0; i < ym; i++) rptr[R[i] + 1]++;
    for (int i = 0; i < X.rows; i++) rptr[i + 1] += rptr[i];
    std::vector<int> nx(rptr.begin(), rptr.end() - 1);
    for (int i = 0; i < ym; i++) ridx[nx[R[i]]++] = i;
  } else {
    std::iota(rptr.begin(), rptr.end(), 0);
    std::iota(ridx.begin(), ridx.end(), 0);
  }
  Csc Y;
  Y.rows = ym;
  Y.cols = yn;
  Y.colptr.assign(static_cast<size_t>(yn) + 1, 0);
  for (int j = 0; j < yn; j++) {
    const int cj = C ? C[j] : j;
    int c = 0;
    for (int p = X.colptr[cj]; p < X.colptr[cj + 1]; p++) {
      const int r = X.rowidx[p];
      c += rptr[r + 1] - rptr[r];
    }
    Y.colptr[j + 1] = Y.colptr[j] + c;
  }
  const int nnz = Y.colptr[yn];
  Y.rowidx.resize(nnz);
  if (with_val) Y.val.resize(nnz);
  std::vector<int> s(nnz);
  std::vector<std::pair<int, int>> tmp;
  for (int j = 0; j < yn; j++) {
    const int cj = C ? C[j] : j;
    tmp.clear();
    for (int p = X.colptr[cj]; p < X.colptr[cj + 1]; p++) {
      const int r = X.rowidx[p];
      for (int t = rptr[r]; t < rptr[r + 1]; t++) tmp.emplace_back(ridx[t], p);
    }
    // setFromTriplets orders rows inside a column; (row, source) pairs are unique
    if (!std::is_sorted(tmp.begin(), tmp.end())) std::stable_sort(tmp.begin(), tmp.end());
    int q = Y.colptr[j];
    for (auto& e : tmp) {
      Y.rowidx[q] = e.first;
      s[q] = e.second;
      if (with_val) Y.val[q] = X.val[e.second];
      q++;
    }
  }
  if (src) *src = std::move(s);
  return Y;
}

Csc transpose(const Csc& X, std::vector<int>* map) {
  Csc T;
  T.rows = X.cols;
  T.cols = X.rows;
  const int nnz = X.nnz();
  T.colptr.assign(static_cast<size_t>(X.rows) + 1, 0);
  for (int p = 0; p < nnz; p++) T.colptr[X.rowidx[p] + 1]++;
  for (int i = 0; i < X.rows; i++) T.colptr[i + 1] += T.colptr[i];
  T.rowidx.resize(nnz);
  const bool with_val = !X.val.empty();
  if (with_val) T.val.resize(nnz);
  std::vector<int> next(T.colptr.begin(), T.colptr.end() - 1);
  if (map) map->resize(nnz);
  for (int j = 0; j < X.cols; j++)
    for (int p = X.colptr[j]; p < X.colptr[j + 1]; p++) {
      const int q = next[X.rowidx[p]]++;
      T.rowidx[q] = j;
      if (with_val) T.val[q] = X.val[p];
      if (map) (*map)[p] = q;
    }
  return T;
}

Csc spgemm_pattern(const Csc& L, const Csc& R) {
  Csc Y;
  Y.rows = L.rows;
  Y.cols = R.cols;
  Y.colptr.assign(static_cast<size_t>(R.cols) + 1, 0);
  // columns of the product are independent: every thread computes a contiguous range of
  // them into its own buffer, the buffers are concatenated in column order
  struct Part {
    int64_t c0 = 0, c1 = 0;
    std::vector<int> rows, cnt;
  };
  std::vector<Part> parts(8);
  std::vector<char> used(8, 0);
  std::mutex mu;
  int next_part = 0;
  parallel_chunks(R.cols, 4096, [&](int64_t c0, int64_t c1) {
    int id;
    {
      std::lock_guard<std::mutex> lk(mu);
      id = next_part++;
    }
    Part& P = parts[static_cast<size_t>(id)];
    used[static_cast<size_t>(id)] = 1;
    P.c0 = c0;
    P.c1 = c1;
    P.cnt.assign(static_cast<size_t>(c1 - c0), 0);
    std::vector<int> mask(static_cast<size_t>(L.rows), -1);
    std::vector<int> touched;
    for (int64_t j = c0; j < c1; j++) {
      touched.clear();
      for (int p = R.colptr[j]; p < R.colptr[j + 1]; p++) {
        const int k = R.rowidx[p];
        for (int q = L.colptr[k]; q < L.colptr[k + 1]; q++) {
          const int i = L.rowidx[q];
          if (mask[i] != static_cast<int>(j)) {
            mask[i] = static_cast<int>(j);
            touched.push_back(i);
          }
        }
      }
      std::sort(touched.begin(), touched.end());
      P.rows.insert(P.rows.end(), touched.begin(), touched.end());
      P.cnt[static_cast<size_t>(j - c0)] = static_cast<int>(touched.size());
    }
  });
  std::vector<const Part*> ord;
  for (size_t i = 0; i < parts.size(); i++)
    if (used[i]) ord.push_back(&parts[i]);
  std::sort(ord.begin(), ord.end(), [](const Part* a, const Part* b) { return a->c0 < b->c0; });
  size_t total = 0;
  for (const Part* P : ord) total += P->rows.size();
  Y.rowidx.reserve(total);
  for (const Part* P : ord) {
    for (int64_t j = P->c0; j < P->c1; j++)
      Y.colptr[static_cast<size_t>(j) + 1] = Y.colptr[static_cast<size_t>(j)] + P->cnt[static_cast<size_t>(j - P->c0)];
    Y.rowidx.insert(Y.rowidx.end(), P->rows.begin(), P->rows.end());
  }
  return Y;
}

std::vector<int> entry_columns(const Csc& X) {
  std::vector<int> c(X.nnz());
  for (int j = 0; j < X.cols; j++)
    for (int p = X.colptr[j]; p < X.colptr[j + 1]; p++) c[p] = j;
  return c;
}

std::vector<int> wavefront_levels(const Csc& A, int* n_levels) {
  std::vector<int> lvl(A.cols, 0);
  int mx = -1;
  for (int i = 0; i < A.cols; i++) {
    int l = 0;
    for (int p = A.colptr[i]; p < A.colptr[i + 1]; p++) {
      const int j = A.rowidx[p];
      if (j < i && lvl[j] + 1 > l) l = lvl[j] + 1;
    }
    lvl[i] = l;
    mx = std::max(mx, l);
  }
  *n_levels = mx + 1;
  return lvl;
}

std::vector<int> greedy_colours(const Csc& A, const std::vector<int>& order, int* n_colours) {
  const int n = A.cols;
  std::vector<int> colour(n, -1);
  std::vector<int> stamp;  // stamp[c] == v  <=> colour c used by a neighbour of v
  int nc = 0;
  for (int t = 0; t < n; t++) {
    const int v = order[t];
    for (int p = A.colptr[v]; p < A.colptr[v + 1]; p++) {
      const int w = A.rowidx[p];
      if (w != v && colour[w] >= 0) stamp[colour[w]] = v;
    }
    int c = 0;
    while (c < nc && stamp[c] == v) c++;
    if (c == nc) {
      stamp.push_back(-1);
      nc++;
    }
    colour[v] = c;
  }
  // Balancing pass: the last colours of a greedy colouring are tiny, which costs a
  // whole kernel phase each.  Move rows from over-full colours into an admissible
  // under-full colour (never creates conflicts: admissibility is re-checked against
  // the live colouring).
  if (nc > 1) {
    std::vector<int64_t> cnt(nc, 0);
    for (int v = 0; v < n; v++) cnt[colour[v]]++;
    const int64_t target = (n + nc - 1) / nc;
    std::vector<char> used(nc);
    for (int t = n - 1; t >= 0; t--) {
      const int v = order[t];
      const int cv = colour[v];
      if (cnt[cv] <= target) continue;
      std::fill(used.begin(), used.end(), 0);
      for (int p = A.colptr[v]; p < A.colptr[v + 1]; p++) {
        const int w = A.rowidx[p];
        if (w != v) used[colour[w]] = 1;
      }
      int best = -1;
      for (int c = 0; c < nc; c++)
        if (!used[c] && c != cv && cnt[c] < target && (best < 0 || cnt[c] < cnt[best])) best = c;
      if (best >= 0) {
        colour[v] = best;
        cnt[cv]--;
        cnt[best]++;
      }
    }
  }
  *n_colours = nc;
  return colour;
}

std::vector<int> bfs_order(const Csc& A) {
  const int n = A.cols;
  std::vector<int> order;
  order.reserve(n);
  std::vector<char> seen(n, 0);
  // start every component at a pseudo-peripheral vertex: two BFS sweeps from the
  // lowest-numbered unseen vertex
  std::vector<int> comp;
  for (int s0 = 0; s0 < n; s0++) {
    if (seen[s0]) continue;
    int start = s0;
    for (int sweep = 0; sweep < 2; sweep++) {
      comp.clear();
      comp.push_back(start);
      seen[start] = 2;
      for (size_t h = 0; h < comp.size(); h++) {
        const int v = comp[h];
        for (int p = A.colptr[v]; p < A.colptr[v + 1]; p++) {
          const int w = A.rowidx[p];
          if (seen[w] != 2 && seen[w] != 1) {
            seen[w] = 2;
            comp.push_back(w);
          }
        }
      }
      start = comp.back();
      for (int v : comp) seen[v] = 0;
    }
    const size_t base = order.size();
    order.push_back(start);
    seen[start] = 1;
    for (size_t h = base; h < order.size(); h++) {
      const int v = order[h];
      for (int p = A.colptr[v]; p < A.colptr[v + 1]; p++) {
        const int w = A.rowidx[p];
        if (!seen[w]) {
          seen[w] = 1;
          order.push_back(w);
        }
      }
    }
  }
  return order;
}

RowOrder make_row_order(const Csc& A, const std::vector<int>& phase, int n_phases,
                        const std::vector<int>& rank, int sigma) {
  const int n = A.cols;
  RowOrder ro;
  ro.perm.resize(n);
  ro.iperm.resize(n);
  ro.phase_ptr.assign(static_cast<size_t>(n_phases) + 1, 0);
  for (int i = 0; i < n; i++) ro.phase_ptr[phase[i] + 1]++;
  for (int p = 0; p < n_phases; p++) ro.phase_ptr[p + 1] += ro.phase_ptr[p];
  // counting sort by phase of the rank-ordered rows (stable)
  std::vector<int> by_rank(n);
  for (int i = 0; i < n; i++) by_rank[rank[i]] = i;
  std::vector<int> next(ro.phase_ptr.begin(), ro.phase_ptr.end() - 1);
  for (int t = 0; t < n; t++) {
    const int v = by_rank[t];
    ro.perm[next[phase[v]]++] = v;
  }
  if (sigma > 1) {
    sigma = std::max(kSliceRows, (sigma / kSliceRows) * kSliceRows);
    auto len = [&](int v) { return A.colptr[v + 1] - A.colptr[v]; };
    for (int p = 0; p < n_phases; p++) {
      const int ps = ro.phase_ptr[p], pe = ro.phase_ptr[p + 1];
      int ws = ps;
      while (ws < pe) {
        const int we = std::min(pe, (ws / sigma + 1) * sigma);
        std::stable_sort(ro.perm.begin() + ws, ro.perm.begin() + we,
                         [&](int a, int b) { return len(a) > len(b); });
        ws = we;
      }
    }
  }
  for (int i = 0; i < n; i++) ro.iperm[ro.perm[i]] = i;
  return ro;
}

Sell build_sell(const Csc& X, const std::vector<int>& row_perm,
                const std::vector<int>& col_iperm, bool sort_cols) {
  Sell S;
  const int n = X.cols;
  S.nrows = n;
  S.nslices = (n + kSliceRows - 1) / kSliceRows;
  S.slice_ptr.assign(static_cast<size_t>(S.nslices) + 1, 0);
  std::vector<int> width(static_cast<size_t>(S.nslices), 0);
  parallel_chunks(S.nslices, 2048, [&](int64_t s0, int64_t s1) {
    for (int64_t s = s0; s < s1; s++) {
      int w = 0;
      const int r1 = static_cast<int>(std::min<int64_t>(n, (s + 1) * kSliceRows));
      for (int r = static_cast<int>(s) * kSliceRows; r < r1; r++) {
        const int c = row_perm[r];
        w = std::max(w, X.colptr[c + 1] - X.colptr[c]);
      }
      width[static_cast<size_t>(s)] = w;
    }
  });
  for (int s = 0; s < S.nslices; s++) {
    const int64_t nxt = static_cast<int64_t>(S.slice_ptr[s]) + static_cast<int64_t>(width[s]) * kSliceRows;
    if (nxt > INT32_MAX) {
      S.nrows = -1;  // signals overflow to the caller
      return S;
    }
    S.slice_ptr[s + 1] = static_cast<int>(nxt);
  }
  const int64_t tot = S.slice_ptr[S.nslices];
  S.col.assign(tot, 0);
  S.src.assign(tot, -1);
  parallel_chunks(S.nslices, 2048, [&](int64_t s0, int64_t s1) {
    std::vector<std::pair<int, int>> ent;
    for (int64_t s = s0; s < s1; s++) {
      const int base = S.slice_ptr[s];
      const int w = (S.slice_ptr[s + 1] - base) / kSliceRows;
      for (int lane = 0; lane < kSliceRows; lane++) {
        const int r = static_cast<int>(s) * kSliceRows + lane;
        int j = 0;
        int last_col = 0;
        if (r < n) {
          const int c = row_perm[r];
          ent.clear();
          for (int p = X.colptr[c]; p < X.colptr[c + 1]; p++) ent.emplace_back(col_iperm[X.rowidx[p]], p);
          if (sort_cols) std::sort(ent.begin(), ent.end());
          for (const auto& e : ent) {
            last_col = e.first;
            S.col[base + j * kSliceRows + lane] = e.first;
            S.src[base + j * kSliceRows + lane] = e.second;
            j++;
          }
        }
        // padding: zero value, column = last real column of the row (same sector)
        for (; j < w; j++) S.col[base + j * kSliceRows + lane] = last_col;
      }
    }
  });
  return S;
}

// ---------------------------------------------------------------------------
// entries of X with a non-zero value (pattern + values); src: entry -> entry of X
static Csc drop_zeros(const Csc& X, std::vector<int>* src) {
  Csc Y;
  Y.rows = X.rows;
  Y.cols = X.cols;
  Y.colptr.assign(static_cast<size_t>(X.cols) + 1, 0);
  src->clear();
  for (int j = 0; j < X.cols; j++) {
    for (int p = X.colptr[j]; p < X.colptr[j + 1]; p++)
      if (X.val[p] != 0.0) {
        Y.rowidx.push_back(X.rowidx[p]);
        Y.val.push_back(X.val[p]);
        src->push_back(p);
      }
    Y.colptr[j + 1] = static_cast<int>(Y.rowidx.size());
  }
  return Y;
}

// Sub (a sorted sub-pattern of A, same shape): position in A of every entry of Sub
static std::vector<int> locate_entries(const Csc& A, const Csc& Sub) {
  std::vector<int> pos(Sub.nnz());
  for (int j = 0; j < A.cols; j++) {
    int p = A.colptr[j];
    for (int q = Sub.colptr[j]; q < Sub.colptr[j + 1]; q++) {
      while (p < A.colptr[j + 1] && A.rowidx[p] < Sub.rowidx[q]) p++;
      pos[q] = (p < A.colptr[j + 1] && A.rowidx[p] == Sub.rowidx[q]) ? p : -1;
    }
  }
  return pos;
}

static void remap_src(Sell& S, const std::vector<int>& map) {
  for (auto& s : S.src)
    if (s >= 0) s = map[s];
}

static void plan_level_layout(LevelPlan& L, const PlanOptions& opt) {
  StageTimer st;
  const int n = L.A.cols;
  L.n = n;
  L.a_col = entry_columns(L.A);
  L.diag_pos.assign(n, -1);
  for (int j = 0; j < n; j++)
    for (int p = L.A.colptr[j]; p < L.A.colptr[j + 1]; p++)
      if (L.A.rowidx[p] == j) L.diag_pos[j] = p;
  st.lap("  columns, diag positions");
  const Csc& A = L.Alive;  // layout decisions look at the compute pattern only
  std::vector<int> order;
  if (opt.locality_reorder) {
    order = bfs_order(A);
  } else {
    order.resize(n);
    std::iota(order.begin(), order.end(), 0);
  }
  st.lap("  bfs order");
  std::vector<int> rank(n);
  for (int t = 0; t < n; t++) rank[order[t]] = t;
  if (opt.smoother == SMG_SMOOTHER_WAVEFRONT) {
    L.phase = wavefront_levels(A, &L.n_phases);
  } else {
    L.phase = greedy_colours(A, order, &L.n_phases);
  }
  if (n == 0) L.n_phases = 0;
  st.lap("  phases (colouring)");
  if (L.nparts > 1 && L.layout != LAYOUT_PLAIN) {
    // multi-GPU: group rows by (part, phase) or (phase, part); see LevelPlan::group
    if (L.layout == LAYOUT_PARTITIONED && L.part.empty()) {
      // finest partitioned level: equal chunks of the breadth-first order (strips of BFS fronts)
      L.part.resize(n);
      for (int t = 0; t < n; t++)
        L.part[order[t]] = static_cast<int>(static_cast<int64_t>(t) * L.nparts / std::max(n, 1));
    }
    std::vector<int> grp(n);
    for (int i = 0; i < n; i++) grp[i] = L.group(L.part[i], L.phase[i]);
    L.order = make_row_order(A, grp, L.nparts * L.n_phases, rank, opt.sigma);
    L.sellA = build_sell(A, L.order.perm, L.order.iperm, opt.sort_cols > 0);
    remap_src(L.sellA, L.live_src);
    return;
  }
  L.layout = LAYOUT_PLAIN;
  L.nparts = 1;
  L.order = make_row_order(A, L.phase, L.n_phases, rank, opt.sigma);
  st.lap("  row order");
  L.sellA = build_sell(A, L.order.perm, L.order.iperm, opt.sort_cols > 0);
  remap_src(L.sellA, L.live_src);
  st.lap("  build SELL A");
}

// ---- multi-GPU exchange lists ---------------------------------------------------------
namespace {
struct Need {
  int dst, src, idx;
  bool operator<(const Need& o) const {
    return dst != o.dst ? dst < o.dst : (src != o.src ? src < o.src : idx < o.idx);
  }
  bool operator==(const Need& o) const { return dst == o.dst && src == o.src && idx == o.idx; }
};
Exchange make_exchange(std::vector<Need>& needs, int world) {
  std::sort(needs.begin(), needs.end());
  needs.erase(std::unique(needs.begin(), needs.end()), needs.end());
  Exchange x;
  x.idx.assign(static_cast<size_t>(world) * world, {});
  for (const Need& nd : needs) x.idx[static_cast<size_t>(nd.src) * world + nd.dst].push_back(nd.idx);
  return x;
}
}  // namespace

// owner of a coarse row = owner of the fine row that carries the largest weight of its
// prolongation column (for subdivision hierarchies: the fine copy of the coarse vertex)
static std::vector<int> coarse_parts(const Csc& P, const std::vector<int>& fine_part) {
  std::vector<int> part(P.cols, 0);
  for (int c = 0; c < P.cols; c++) {
    double best = -1.0;
    for (int p = P.colptr[c]; p < P.colptr[c + 1]; p++) {
      const double w = P.val[p] < 0 ? -P.val[p] : P.val[p];
      if (w > best) {
        best = w;
        part[c] = fine_part[P.rowidx[p]];
      }
    }
  }
  return part;
}

static void plan_exchanges(Plan& pl, const std::vector<Csc>& Pz) {
  const int W = pl.world;
  const int nlev = static_cast<int>(pl.lv.size());
  for (int l = 0; l < nlev; l++) {
    LevelPlan& L = pl.lv[l];
    if (L.nparts <= 1) continue;
    const std::vector<int>& ip = L.order.iperm;
    {  // every row of part src -> all other ranks
      L.gather_all.idx.assign(static_cast<size_t>(W) * W, {});
      std::vector<std::vector<int>> own(W);
      for (int i = 0; i < L.n; i++) own[L.part[i]].push_back(ip[i]);
      for (int s = 0; s < W; s++) {
        std::sort(own[s].begin(), own[s].end());
        for (int d = 0; d < W; d++)
          if (d != s) L.gather_all.idx[static_cast<size_t>(s) * W + d] = own[s];
      }
    }
    if (L.layout != LAYOUT_PARTITIONED) continue;
    std::vector<Need> nu;
    std::vector<std::vector<Need>> nup(L.n_phases);
    for (int j = 0; j < L.n; j++)
      for (int p = L.Alive.colptr[j]; p < L.Alive.colptr[j + 1]; p++) {
        const int i = L.Alive.rowidx[p];  // symmetric pattern: row i reads column j
        if (L.part[i] == L.part[j]) continue;
        const Need nd{L.part[i], L.part[j], ip[j]};
        nu.push_back(nd);
        nup[L.phase[j]].push_back(nd);
      }
    L.halo_u = make_exchange(nu, W);
    L.halo_u_phase.clear();
    for (auto& v : nup) L.halo_u_phase.push_back(make_exchange(v, W));
    if (l + 1 < nlev) {
      LevelPlan& C = pl.lv[l + 1];  // PARTITIONED or SPLIT: C.part is set
      const Csc& P = Pz[l + 1];     // rows: level l, columns: level l+1, zero-free
      std::vector<Need> nr, npu;
      for (int c = 0; c < P.cols; c++)
        for (int p = P.colptr[c]; p < P.colptr[c + 1]; p++) {
          const int f = P.rowidx[p];
          if (L.part[f] == C.part[c]) continue;
          nr.push_back(Need{C.part[c], L.part[f], ip[f]});                    // restriction row c reads r[f]
          npu.push_back(Need{L.part[f], C.part[c], C.order.iperm[c]});        // prolongation row f reads u_c[c]
        }
      L.halo_r = make_exchange(nr, W);
      if (C.layout == LAYOUT_PARTITIONED) C.halo_pu = make_exchange(npu, W);
    }
  }
}

static bool pattern_symmetric(const Csc& A, std::vector<int>* tmap) {
  if (A.rows != A.cols) return false;
  Csc T = transpose(A, tmap);
  return T.colptr == A.colptr && T.rowidx == A.rowidx;
}

int build_plan(const Csc& A, const int* known, int nknown, const std::vector<Csc>& P_full,
               const PlanOptions& opt_in, Plan* plan) {
  PlanOptions opt = opt_in;
  if (opt.sort_cols < 0) opt.sort_cols = opt.smoother == SMG_SMOOTHER_MULTICOLOUR ? 1 : 0;
  Plan& pl = *plan;
  pl = Plan();
  const int nlev = static_cast<int>(P_full.size()) + 1;
  if (nlev < 2) {
    pl.error = "at least 2 multigrid levels are required";
    return SMG_E_NLEVELS;
  }
  if (A.rows != A.cols) {
    pl.error = "A must be square";
    return SMG_E_INVALID;
  }
  if (P_full[0].rows != A.rows) {
    pl.error = "P[1] must have as many rows as A";
    return SMG_E_INVALID;
  }
  for (int l = 1; l + 1 < nlev; l++)
    if (P_full[l].rows != P_full[l - 1].cols) {
      pl.error = "prolongation sizes do not chain";
      return SMG_E_INVALID;
    }
  StageTimer stage;
  pl.n = A.rows;
  pl.lv.resize(nlev);
  Csc Apat = A;
  Apat.val.clear();
  if (nknown < 0) {
    // variant without fixed values: src/min_quad_with_fixed_mg.cpp:3-51
    pl.has_fixed = false;
    pl.unknown.resize(A.rows);
    std::iota(pl.unknown.begin(), pl.unknown.end(), 0);
    pl.LHS = Apat;
    pl.lhs_src.resize(A.nnz());
    std::iota(pl.lhs_src.begin(), pl.lhs_src.end(), 0);
    for (int l = 1; l < nlev; l++) {
      pl.lv[l].P = P_full[l - 1];
      pl.lv[l].PT = transpose(P_full[l - 1], nullptr);
    }
  } else {
    // variant with fixed values: src/min_quad_with_fixed_mg.cpp:137-257
    for (int i = 0; i < nknown; i++)
      if (known[i] < 0 || known[i] >= A.rows) {
        pl.error = "known index out of range";
        return SMG_E_INVALID;
      }
    pl.has_fixed = true;
    pl.known.assign(known, known + nknown);
    pl.unknown = setdiff_range(A.rows, known, nknown);                              // :156-158
    const int nu = static_cast<int>(pl.unknown.size());
    static const int kEmpty = 0;  // slice(): a null index list means "all", so never pass one
    const int* kn = nknown > 0 ? pl.known.data() : &kEmpty;
    const int* un = nu > 0 ? pl.unknown.data() : &kEmpty;
    pl.LHS = slice(Apat, un, nu, un, nu, &pl.lhs_src);      // :167
    pl.Auk = slice(Apat, un, nu, kn, nknown, &pl.auk_src);  // :170
    for (int l = 1; l < nlev; l++) pl.lv[l].P = P_full[l - 1];
    pl.lv[1].P = slice(P_full[0], un, nu, nullptr, 0, nullptr);  // :185
    for (int l = 1; l < nlev; l++) {                                            // :186-220
      Csc& P = pl.lv[l].P;
      std::vector<int> keep;
      keep.reserve(P.cols);
      for (int c = 0; c < P.cols; c++)
        for (int p = P.colptr[c]; p < P.colptr[c + 1]; p++)
          if (P.val[p] > 1e-15) {
            keep.push_back(c);
            break;
          }
      if (static_cast<int>(keep.size()) < P.cols) {
        const int nk = static_cast<int>(keep.size());
        Csc Pn = slice(P, nullptr, 0, keep.data(), nk, nullptr);
        P = std::move(Pn);
        pl.lv[l].pruned = true;
        if (l < nlev - 1) pl.lv[l + 1].P = slice(P_full[l], keep.data(), nk, nullptr, 0, nullptr);
        pl.lv[l].keep = std::move(keep);
      } else {
        break;
      }
    }
    for (int l = 1; l < nlev; l++) pl.lv[l].PT = transpose(pl.lv[l].P, nullptr);  // :226
  }
  stage.lap("slices, pruning, transposes");
  // patterns of the Galerkin products, :223-228 (left to right: (PT*A)*P)
  pl.lv[0].A = pl.LHS;
  for (int l = 1; l < nlev; l++) {
    if (pl.lv[l].P.rows != pl.lv[l - 1].A.rows) {
      pl.error = "prolongation rows do not match the level matrix";
      return SMG_E_INVALID;
    }
    pl.lv[l].T1 = spgemm_pattern(pl.lv[l].PT, pl.lv[l - 1].A);
    pl.lv[l].t1_col = entry_columns(pl.lv[l].T1);
    pl.lv[l].A = spgemm_pattern(pl.lv[l].T1, pl.lv[l].P);
  }
  stage.lap("Galerkin patterns");
  // compute patterns (LevelPlan::Alive): Galerkin patterns of the zero-free part of P
  std::vector<Csc> Pz(nlev), PTz(nlev);
  std::vector<std::vector<int>> pz_src(nlev), ptz_src(nlev);
  pl.lv[0].Alive = pl.lv[0].A;
  pl.lv[0].live_src.resize(pl.lv[0].A.nnz());
  std::iota(pl.lv[0].live_src.begin(), pl.lv[0].live_src.end(), 0);
  for (int l = 1; l < nlev; l++) {
    Pz[l] = drop_zeros(pl.lv[l].P, &pz_src[l]);
    PTz[l] = drop_zeros(pl.lv[l].PT, &ptz_src[l]);
    Csc Tz = spgemm_pattern(PTz[l], pl.lv[l - 1].Alive);
    pl.lv[l].Alive = spgemm_pattern(Tz, Pz[l]);
    pl.lv[l].live_src = locate_entries(pl.lv[l].A, pl.lv[l].Alive);
  }
  stage.lap("compute patterns");
  // multi-GPU: which levels are partitioned by rows
  pl.world = std::max(1, opt.world);
  pl.dist_levels = 0;
  if (pl.world > 1) {
    int nd = opt.dist_levels;
    if (nd < 0) {
      // A level below level 0 is partitioned when that pays for its halo exchanges (~45 us per
      // iteration, measured on B200 / NVSwitch): a colour-phase kernel costs ~2.7 us however few
      // rows it has, so only levels of 500 000 rows or more get faster when split, whatever the
      // number of ranks, as long as a rank keeps at least 50 000 rows.  dist_min_rows > 0
      // replaces this rule by "at least dist_min_rows rows per rank".
      nd = 1;
      while (nd < nlev - 1 &&
             (opt.dist_min_rows > 0
                  ? pl.lv[nd].A.cols >= static_cast<int64_t>(opt.dist_min_rows) * pl.world
                  : (pl.lv[nd].A.cols >= 500000 && pl.lv[nd].A.cols >= static_cast<int64_t>(50000) * pl.world)))
        nd++;
    }
    pl.dist_levels = std::max(1, std::min(nd, nlev - 1));
  }
  for (int l = 0; l < nlev; l++) {
    LevelPlan& L = pl.lv[l];
    if (pl.world > 1 && l <= pl.dist_levels) {
      L.nparts = pl.world;
      L.layout = l < pl.dist_levels ? LAYOUT_PARTITIONED : LAYOUT_SPLIT;
      if (l > 0) L.part = coarse_parts(Pz[l], pl.lv[l - 1].part);
    }
    if (!pattern_symmetric(L.A, &L.tmap) || !pattern_symmetric(L.Alive, nullptr)) {
      pl.error = "sparsity pattern of the level matrix is not symmetric (level " +
                 std::to_string(l) + ")";
      return SMG_E_NOT_SYMMETRIC;
    }
    plan_level_layout(L, opt);
    if (L.sellA.nrows < 0) {
      pl.error = "level too large for 32-bit SELL offsets";
      return SMG_E_UNSUPPORTED;
    }
    for (int j = 0; j < L.n; j++)
      if (L.diag_pos[j] < 0) {
        pl.error = "level matrix has a structurally missing diagonal entry";
        return SMG_E_INVALID;
      }
  }
  stage.lap("level layouts (order, SELL A)");
  for (int l = 1; l < nlev; l++) {
    LevelPlan& L = pl.lv[l];
    // y_fine = P x_coarse : rows of P = columns of PT's CSC (explicit zeros skipped)
    L.sellP = build_sell(PTz[l], pl.lv[l - 1].order.perm, L.order.iperm, opt.sort_cols > 0);
    remap_src(L.sellP, ptz_src[l]);
    // y_coarse = PT x_fine : rows of PT = columns of P's CSC
    L.sellPT = build_sell(Pz[l], L.order.perm, pl.lv[l - 1].order.iperm, opt.sort_cols > 0);
    remap_src(L.sellPT, pz_src[l]);
    if (L.sellP.nrows < 0 || L.sellPT.nrows < 0) {
      pl.error = "transfer operator too large for 32-bit SELL offsets";
      return SMG_E_UNSUPPORTED;
    }
  }
  stage.lap("SELL P / PT");
  if (pl.world > 1) plan_exchanges(pl, Pz);
  stage.lap("exchange lists");
  return SMG_OK;
}

}  // namespace smg
