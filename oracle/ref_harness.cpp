// ref_harness.cpp -- C entry points (the same orc_* names as oracle/smg_oracle.c) in front of
// the REFERENCE'S OWN hot-path sources, compiled unmodified from where they lie:
//     /root/reference/src/mg_VCycle.cpp, /root/reference/src/min_quad_with_fixed_mg.cpp
// against the functional Eigen / igl stand-in of oracle/ref_shim (Eigen itself is not vendored
// by the reference and is absent from this image; see ref_shim/Eigen/Core for what exactly is
// restated).  TEST INFRASTRUCTURE: built by `make -C oracle ref` into oracle/_ref/ (git-ignored),
// loaded only by tests/ and by bench.py's reference arm.  Nothing here is product code.
//
// Every entry point only marshals raw CSC / column-major arrays into the reference's types
// (mg_data, min_quad_with_fixed_mg_data) and calls the reference function named in its comment.
// The two reference files are pulled into THIS translation unit (found through -I$(REF)/src):
// their function templates are only instantiated for the types their own callers use, and an
// optimising build inlines most of those instantiations away, so a separately compiled harness
// could not link against them.  Including them makes every template visible here; the files
// themselves are byte-for-byte the reference's.
//
// -DSMG_HARNESS_USE_ADAPTER builds the SAME entry points over adapter/smg_eigen_adapter.cpp
// instead (the drop-in replacement of those two files, which forwards to libsmg.so on the
// GPU): tests/test_gpu_adapter_dropin.py then drives one identical call sequence through the
// reference's code and through the drop-in and compares what comes back.
#ifdef SMG_HARNESS_USE_ADAPTER
#include <smg_eigen_adapter.cpp>

#include <cstdlib>
#else
#include <mg_VCycle.cpp>
#include <min_quad_with_fixed_mg.cpp>
#endif

#include <cstdio>
#include <cstring>
#include <exception>
#include <iostream>
#include <sstream>
#include <vector>

namespace {
typedef Eigen::SparseMatrix<double> SpMat;
typedef Eigen::SimplicialLDLT<SpMat> Ldlt;

struct Ref {
  int nlev = 0;
  std::vector<mg_data> mg;
  min_quad_with_fixed_mg_data data;
  Ldlt solver;
  bool has_fixed = false;
  int nknown = 0;
};

SpMat from_csc(int rows, int cols, const int* colptr, const int* rowidx, const double* val) {
  SpMat M(rows, cols);
  const int nnz = colptr[cols];
  M.resizeNonZeros(nnz);
  std::memcpy(M.outerIndexPtr(), colptr, sizeof(int) * (static_cast<size_t>(cols) + 1));
  if (nnz > 0) {
    std::memcpy(M.innerIndexPtr(), rowidx, sizeof(int) * static_cast<size_t>(nnz));
    std::memcpy(M.valuePtr(), val, sizeof(double) * static_cast<size_t>(nnz));
  }
  return M;
}

Eigen::MatrixXd from_colmajor(const double* p, Eigen::Index n, int k) {
  Eigen::MatrixXd M(n, k);
  if (n * k > 0) std::memcpy(M.data(), p, sizeof(double) * static_cast<size_t>(n * k));
  return M;
}
Eigen::VectorXd vec_from(const double* p, Eigen::Index n) {
  Eigen::VectorXd v(n);
  if (n > 0) std::memcpy(v.data(), p, sizeof(double) * static_cast<size_t>(n));
  return v;
}
template <class M>
void to_colmajor(const M& m, double* out) {
  if (m.size() > 0) std::memcpy(out, m.data(), sizeof(double) * static_cast<size_t>(m.size()));
}

// the drop-in reports failures (no CUDA device, ...) as exceptions: never let one cross the C ABI
#define ORC_TRY try {
#define ORC_CATCH(ret)                                        \
  }                                                           \
  catch (const std::exception& e) {                           \
    std::fprintf(stderr, "ref_harness: %s\n", e.what());      \
    ret;                                                      \
  }

// the reference prints one line per iteration (min_quad_with_fixed_mg.cpp:111,334,349)
struct Quiet {
  std::ostringstream sink;  // constructed before `old` takes its buffer
  std::streambuf* old;
  Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~Quiet() { std::cout.rdbuf(old); }
};
}  // namespace

extern "C" {

void* orc_create(int nlev) {
#ifdef SMG_HARNESS_USE_ADAPTER
  setenv("SMG_MIRROR_TO_HOST", "1", 1);  // mg[lv].A / A_diag / P / PT, data.LHS / Auk come back
  setenv("SMG_QUIET", "1", 1);
#endif
  Ref* s = new Ref();
  s->nlev = nlev;
  s->mg.resize(static_cast<size_t>(nlev));
  return s;
}

void orc_destroy(void* h) {
  Ref* s = static_cast<Ref*>(h);
#ifdef SMG_HARNESS_USE_ADAPTER
  if (s) smg_adapter_release(&s->data);  // what a caller does before its solver objects die
#endif
  delete s;
}
int orc_live_handles(void) {
#ifdef SMG_HARNESS_USE_ADAPTER
  return smg_adapter_live_handles();
#else
  return 0;
#endif
}

// what mg_precompute leaves for level lv >= 1 (src/mg_precompute.cpp:71-77)
int orc_set_prolongation(void* h, int lv, int rows, int cols, const int* colptr, const int* rowidx,
                         const double* val) {
  Ref* s = static_cast<Ref*>(h);
  if (lv < 1 || lv >= s->nlev) return -1;
  mg_data& d = s->mg[static_cast<size_t>(lv)];
  d.P_full = from_csc(rows, cols, colptr, rowidx, val);
  d.P = d.P_full;
  d.PT = d.P.transpose();
  return 0;
}

// min_quad_with_fixed_mg_precompute, both variants (src/min_quad_with_fixed_mg.cpp:3-51, :137-257)
int orc_precompute(void* h, int n, const int* colptr, const int* rowidx, const double* val,
                   const int* known, int nknown) {
  ORC_TRY
  Ref* s = static_cast<Ref*>(h);
  if (s->nlev < 2) return -2;
  for (size_t lv = 1; lv < s->mg.size(); lv++) {  // a fresh mg_precompute state
    s->mg[lv].P = s->mg[lv].P_full;
    s->mg[lv].PT = s->mg[lv].P.transpose();
  }
  const SpMat A = from_csc(n, n, colptr, rowidx, val);
  s->data = min_quad_with_fixed_mg_data();
  Quiet q;
  if (nknown < 0) {
    s->has_fixed = false;
    s->nknown = 0;
    min_quad_with_fixed_mg_precompute(A, s->data, s->mg, s->solver);
    s->data.unknown = Eigen::VectorXi::LinSpaced(n, 0, n - 1);  // the free variant leaves it empty
  } else {
    s->has_fixed = true;
    s->nknown = nknown;
    Eigen::VectorXi kn(nknown);
    for (int i = 0; i < nknown; i++) kn(i) = known[i];
    min_quad_with_fixed_mg_precompute(A, kn, s->data, s->mg, s->solver);
  }
#ifdef SMG_HARNESS_USE_ADAPTER
  return 0;  // the drop-in ignores the caller's SimplicialLDLT (coarse factorisation on the device)
#else
  return s->solver.ok() ? 0 : -3;
#endif
  ORC_CATCH(return -4)
}

// relax (src/mg_VCycle.cpp:113-178)
void orc_relax(void* h, int lv, int iters, const double* B, double* u, int k) {
  ORC_TRY
  Ref* s = static_cast<Ref*>(h);
  const Eigen::Index n = s->mg[static_cast<size_t>(lv)].A.rows();
  Eigen::MatrixXd b = from_colmajor(B, n, k), x = from_colmajor(u, n, k);
  relax(b, lv, iters, x, s->mg);
  to_colmajor(x, u);
  ORC_CATCH(return)
}

// A (src/mg_VCycle.cpp:62-70)
void orc_apply_A(void* h, int lv, const double* u, double* Au, int k) {
  ORC_TRY
  Ref* s = static_cast<Ref*>(h);
  Eigen::MatrixXd x = from_colmajor(u, s->mg[static_cast<size_t>(lv)].A.cols(), k), y;
  A(x, s->mg, lv, y);
  to_colmajor(y, Au);
  ORC_CATCH(return)
}

// restrict (src/mg_VCycle.cpp:72-81)
void orc_restrict(void* h, int lv, const double* xin, double* Rx, int k) {
  ORC_TRY
  Ref* s = static_cast<Ref*>(h);
  Eigen::MatrixXd x = from_colmajor(xin, s->mg[static_cast<size_t>(lv) + 1].PT.cols(), k), y;
  restrict(x, s->mg, lv, y);
  to_colmajor(y, Rx);
  ORC_CATCH(return)
}

// prolong (src/mg_VCycle.cpp:83-92)
void orc_prolong(void* h, int lv, const double* xin, double* Px, int k) {
  ORC_TRY
  Ref* s = static_cast<Ref*>(h);
  Eigen::MatrixXd x = from_colmajor(xin, s->mg[static_cast<size_t>(lv) + 1].P.cols(), k), y;
  prolong(x, s->mg, lv, y);
  to_colmajor(y, Px);
  ORC_CATCH(return)
}

// coarseSolve (src/mg_VCycle.cpp:181-201)
void orc_coarse_solve(void* h, const double* B, double* u, int k) {
  ORC_TRY
  Ref* s = static_cast<Ref*>(h);
  const int lv = s->nlev - 1;
  const Eigen::Index n = s->mg[static_cast<size_t>(lv)].A.rows();
  Eigen::MatrixXd b = from_colmajor(B, n, k), x = from_colmajor(u, n, k);
  coarseSolve(s->solver, b, lv, x, s->mg);
  to_colmajor(x, u);
  ORC_CATCH(return)
}

// mg_VCycle (src/mg_VCycle.cpp:3-59)
void orc_vcycle(void* h, const double* B, int pre, int post, int lv, double* u, int k) {
  ORC_TRY
  Ref* s = static_cast<Ref*>(h);
  const Eigen::Index n = s->mg[static_cast<size_t>(lv)].A.rows();
  Eigen::MatrixXd b = from_colmajor(B, n, k), x = from_colmajor(u, n, k);
  mg_VCycle(s->solver, b, pre, post, lv, x, s->mg);
  to_colmajor(x, u);
  ORC_CATCH(return)
}

// min_quad_with_fixed_mg_solve with explicit tolerance and maxIter
// (src/min_quad_with_fixed_mg.cpp:80-135 free, :288-361 fixed); k == 1 goes through the
// VectorXd instantiation like 03/04, k > 1 through the MatrixXd one like 05
int orc_solve(void* h, const double* RHS, const double* known_val, const double* z0, int k, double tol,
              int max_iter, double* z, double* r_his, int* n_his) {
  ORC_TRY
  Ref* s = static_cast<Ref*>(h);
  const Eigen::Index n = s->data.n;
  std::vector<double> hist;
  bool ok;
  Quiet q;
  if (k == 1) {
    Eigen::VectorXd b = vec_from(RHS, n), x0 = vec_from(z0, n), x;
    if (s->has_fixed) {
      Eigen::VectorXd kv = vec_from(known_val, s->nknown);
      ok = min_quad_with_fixed_mg_solve(s->data, b, kv, x0, s->solver, tol, max_iter, s->mg, x, hist);
    } else {
      ok = min_quad_with_fixed_mg_solve(s->data, b, x0, s->solver, tol, max_iter, s->mg, x, hist);
    }
    to_colmajor(x, z);
  } else {
    Eigen::MatrixXd b = from_colmajor(RHS, n, k), x0 = from_colmajor(z0, n, k), x;
    if (s->has_fixed) {
      Eigen::MatrixXd kv = from_colmajor(known_val, s->nknown, k);
      ok = min_quad_with_fixed_mg_solve(s->data, b, kv, x0, s->solver, tol, max_iter, s->mg, x, hist);
    } else {
      ok = min_quad_with_fixed_mg_solve(s->data, b, x0, s->solver, tol, max_iter, s->mg, x, hist);
    }
    to_colmajor(x, z);
  }
  for (size_t i = 0; i < hist.size(); i++) r_his[i] = hist[i];
  *n_his = static_cast<int>(hist.size());
  return ok ? 1 : 0;
  ORC_CATCH(return -4)
}

// Two solves in a row that REUSE one r_his vector, as a caller's time-step loop would: the
// reference clears it at the start of every solve (src/min_quad_with_fixed_mg.cpp:105, :327).
// Returns r_his.size() after the first and after the second solve (k = 1).
int orc_solve_twice_same_rhis(void* h, const double* RHS, const double* known_val, const double* z0, double tol,
                              int max_iter, int* n_first, int* n_second) {
  ORC_TRY
  Ref* s = static_cast<Ref*>(h);
  const Eigen::Index n = s->data.n;
  std::vector<double> hist;
  Quiet q;
  Eigen::VectorXd b = vec_from(RHS, n), x0 = vec_from(z0, n), x;
  for (int rep = 0; rep < 2; rep++) {
    if (s->has_fixed) {
      Eigen::VectorXd kv = vec_from(known_val, s->nknown);
      min_quad_with_fixed_mg_solve(s->data, b, kv, x0, s->solver, tol, max_iter, s->mg, x, hist);
    } else {
      min_quad_with_fixed_mg_solve(s->data, b, x0, s->solver, tol, max_iter, s->mg, x, hist);
    }
    *(rep == 0 ? n_first : n_second) = static_cast<int>(hist.size());
  }
  return 0;
  ORC_CATCH(return -4)
}

// `cycles` iterations of the solve loop on the unknown-sized system (the body of
// src/min_quad_with_fixed_mg.cpp:330-347 without the early exit): timing aid
void orc_iterate(void* h, const double* bu, double* zu, int k, int cycles, double* r_his) {
  ORC_TRY
  Ref* s = static_cast<Ref*>(h);
  const Eigen::Index nu = s->mg[0].A.rows();
  Eigen::MatrixXd b = from_colmajor(bu, nu, k), x = from_colmajor(zu, nu, k);
  for (int it = 0; it < cycles; it++) {
    r_his[it] = (b - s->mg[0].A * x).norm();
    mg_VCycle(s->solver, b, 2, 2, 0, x, s->mg);
  }
  to_colmajor(x, zu);
  ORC_CATCH(return)
}

int orc_num_levels(void* h) { return static_cast<Ref*>(h)->nlev; }
int orc_num_unknown(void* h) { return static_cast<int>(static_cast<Ref*>(h)->data.unknown.size()); }
void orc_get_unknown(void* h, int* out) {
  Ref* s = static_cast<Ref*>(h);
  for (Eigen::Index i = 0; i < s->data.unknown.size(); i++) out[i] = s->data.unknown(i);
}
// the reference does not keep its keepIdx lists (min_quad_with_fixed_mg.cpp:190-204): -3
int orc_get_keep(void*, int, int*) { return -3; }

static const SpMat* pick(Ref* s, int lv, int which) {
  if (lv < 0 || lv >= s->nlev) return nullptr;
  mg_data& d = s->mg[static_cast<size_t>(lv)];
  switch (which) {
    case 0: return &d.A;
    case 1: return lv >= 1 ? &d.P : nullptr;
    case 2: return lv >= 1 ? &d.PT : nullptr;
    case 3: return &s->data.LHS;
    case 4: return s->has_fixed ? &s->data.Auk : nullptr;
  }
  return nullptr;
}
int orc_matrix_dims(void* h, int lv, int which, int* rows, int* cols, int* nnz) {
  const SpMat* m = pick(static_cast<Ref*>(h), lv, which);
  if (!m) return -1;
  *rows = static_cast<int>(m->rows());
  *cols = static_cast<int>(m->cols());
  *nnz = static_cast<int>(m->nonZeros());
  return 0;
}
int orc_matrix_copy(void* h, int lv, int which, int* colptr, int* rowidx, double* val) {
  const SpMat* m = pick(static_cast<Ref*>(h), lv, which);
  if (!m) return -1;
  std::memcpy(colptr, m->outerIndexPtr(), sizeof(int) * (static_cast<size_t>(m->cols()) + 1));
  if (m->nonZeros() > 0) {
    std::memcpy(rowidx, m->innerIndexPtr(), sizeof(int) * static_cast<size_t>(m->nonZeros()));
    std::memcpy(val, m->valuePtr(), sizeof(double) * static_cast<size_t>(m->nonZeros()));
  }
  return 0;
}
void orc_get_diag(void* h, int lv, double* out) {
  Ref* s = static_cast<Ref*>(h);
  to_colmajor(s->mg[static_cast<size_t>(lv)].A_diag, out);
}
int orc_coarse_bandwidth(void*) { return -1; }
// precompute calls the drop-in served by a numeric-only refresh (0 for the reference sources)
int orc_refresh_count(void) {
#ifdef SMG_HARNESS_USE_ADAPTER
  return smg_adapter_refresh_count();
#else
  return 0;
#endif
}
const char* orc_impl(void) {
#ifdef SMG_HARNESS_USE_ADAPTER
  return "adapter/smg_eigen_adapter.cpp + libsmg.so behind the reference's headers (ref_shim Eigen stand-in)";
#else
  return "reference sources (mg_VCycle.cpp, min_quad_with_fixed_mg.cpp) on the ref_shim Eigen stand-in";
#endif
}
}  // extern "C"
