/*
 * smg_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C, single-threaded restatement of the hot path of
 * HTDerekLiu/surface_multigrid_code:
 *     src/mg_VCycle.cpp            (V-cycle, relax, A, restrict, prolong, coarseSolve)
 *     src/min_quad_with_fixed_mg.cpp (precompute / solve, with and without fixed DOFs)
 * on raw CSC arrays.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * library (libsmg.so) never links or calls it.
 *
 * PINNING: the reference ships no tests / golden vectors for this path and genuine
 * Eigen 3.3.7 is not in this image (libigl's CMake downloads it:
 * libigl/cmake/LibiglDownloadExternal.cmake:67-74; no network).  This restatement is
 * pinned against (a) the reference's OWN two source files compiled unmodified on an
 * Eigen stand-in (oracle/_ref, `make ref`; tests/test_reference_sources.py: index
 * outputs, Galerkin values, relax / A / restrict / prolong bit-exact), (b) an
 * independent scipy restatement (tests/test_oracle.py) and (c) libigl's adjacent
 * known-answer tests for the problem generator.  Still unpinned: Eigen's own
 * arithmetic order, which both this file and the stand-in restate.
 *
 * Third-party arithmetic restated here (Eigen 3.3.7, not vendored in the
 * reference):
 *   - col-major sparse * dense  : result zeroed, then for every column j in
 *     ascending order, for every stored (i,j): res(i) += A(i,j) * x(j)
 *     (mul rounded, then add rounded; build with -ffp-contract=off).
 *   - sparse * sparse           : "conservative" product: structural zeros are
 *     kept; res(i,j) = sum over k (ascending, storage order of rhs column j) of
 *     lhs(i,k)*rhs(k,j); result columns sorted by row index.
 *   - transpose, setFromTriplets: explicit zeros kept, entries sorted.
 *   - SimplicialLDLT            : replaced by RCM + banded Cholesky (a different
 *     elimination order, so coarse solves agree to rounding, not bit-for-bit).
 *
 * Build:  gcc -O3 -ffp-contract=off -fPIC -shared (see oracle/Makefile).
 */
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* CSC matrix                                                                 */
/* ------------------------------------------------------------------------- */
typedef struct {
  int rows, cols;
  int *colptr; /* cols+1 */
  int *rowidx; /* nnz    */
  double *val; /* nnz    */
} orc_csc;

static orc_csc *csc_alloc(int rows, int cols, int nnz) {
  orc_csc *m = (orc_csc *)calloc(1, sizeof(orc_csc));
  m->rows = rows;
  m->cols = cols;
  m->colptr = (int *)calloc((size_t)cols + 1, sizeof(int));
  m->rowidx = (int *)malloc(sizeof(int) * (size_t)(nnz > 0 ? nnz : 1));
  m->val = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
  return m;
}

static void csc_free(orc_csc *m) {
  if (!m) return;
  free(m->colptr);
  free(m->rowidx);
  free(m->val);
  free(m);
}

static orc_csc *csc_copy_raw(int rows, int cols, const int *colptr, const int *rowidx,
                             const double *val) {
  int nnz = colptr[cols];
  orc_csc *m = csc_alloc(rows, cols, nnz);
  memcpy(m->colptr, colptr, sizeof(int) * ((size_t)cols + 1));
  memcpy(m->rowidx, rowidx, sizeof(int) * (size_t)nnz);
  memcpy(m->val, val, sizeof(double) * (size_t)nnz);
  return m;
}

static orc_csc *csc_clone(const orc_csc *a) {
  return csc_copy_raw(a->rows, a->cols, a->colptr, a->rowidx, a->val);
}

/* Eigen: SparseMatrix = other.transpose()  -> counting sort, explicit zeros kept,
 * row indices ascending inside every column of the result. */
static orc_csc *csc_transpose(const orc_csc *a) {
  int nnz = a->colptr[a->cols];
  orc_csc *t = csc_alloc(a->cols, a->rows, nnz);
  for (int p = 0; p < nnz; p++) t->colptr[a->rowidx[p] + 1]++;
  for (int i = 0; i < a->rows; i++) t->colptr[i + 1] += t->colptr[i];
  int *next = (int *)malloc(sizeof(int) * ((size_t)a->rows + 1));
  memcpy(next, t->colptr, sizeof(int) * ((size_t)a->rows + 1));
  for (int j = 0; j < a->cols; j++)
    for (int p = a->colptr[j]; p < a->colptr[j + 1]; p++) {
      int q = next[a->rowidx[p]]++;
      t->rowidx[q] = j;
      t->val[q] = a->val[p];
    }
  free(next);
  return t;
}

/* igl::setdiff(0..n-1, known)  (libigl/include/igl/setdiff.cpp:19-75):
 * sorted ascending complement; duplicates / order of `known` irrelevant. */
static int setdiff_range(int n, const int *known, int nk, int *unknown) {
  char *mark = (char *)calloc((size_t)n + 1, 1);
  for (int i = 0; i < nk; i++)
    if (known[i] >= 0 && known[i] < n) mark[known[i]] = 1;
  int c = 0;
  for (int i = 0; i < n; i++)
    if (!mark[i]) unknown[c++] = i;
  free(mark);
  return c;
}

/* igl::slice(X,R,C,Y) for sparse X (libigl/include/igl/slice.cpp:13-77):
 * Y(i,j) = X(R(i),C(j)); R and C may repeat indices; explicit zeros kept;
 * result built by setFromTriplets => columns sorted by row. R==NULL / C==NULL
 * means "all, in order" (the dim=1 / dim=2 overloads, slice.cpp:80-111). */
static orc_csc *csc_slice(const orc_csc *x, const int *R, int nr, const int *C, int nc) {
  int ym = R ? nr : x->rows, yn = C ? nc : x->cols;
  /* RI: list of output rows for every input row (CSR-like buckets, stable) */
  int *ri_ptr = (int *)calloc((size_t)x->rows + 2, sizeof(int));
  int *ri_idx = (int *)malloc(sizeof(int) * (size_t)(ym > 0 ? ym : 1));
  if (R) {
    for (int i = 0; i < ym; i++) ri_ptr[R[i] + 1]++;
    for (int i = 0; i < x->rows; i++) ri_ptr[i + 1] += ri_ptr[i];
    int *nx = (int *)malloc(sizeof(int) * ((size_t)x->rows + 1));
    memcpy(nx, ri_ptr, sizeof(int) * ((size_t)x->rows + 1));
    for (int i = 0; i < ym; i++) ri_idx[nx[R[i]]++] = i;
    free(nx);
  } else {
    for (int i = 0; i <= x->rows; i++) ri_ptr[i] = i;
    for (int i = 0; i < ym; i++) ri_idx[i] = i;
  }
  /* count entries per output column */
  orc_csc *y = NULL;
  int *cnt = (int *)calloc((size_t)yn + 1, sizeof(int));
  for (int j = 0; j < yn; j++) {
    int cj = C ? C[j] : j;
    int c = 0;
    for (int p = x->colptr[cj]; p < x->colptr[cj + 1]; p++) {
      int r = x->rowidx[p];
      c += ri_ptr[r + 1] - ri_ptr[r];
    }
    cnt[j + 1] = cnt[j] + c;
  }
  y = csc_alloc(ym, yn, cnt[yn]);
  memcpy(y->colptr, cnt, sizeof(int) * ((size_t)yn + 1));
  for (int j = 0; j < yn; j++) {
    int cj = C ? C[j] : j;
    int q = cnt[j];
    for (int p = x->colptr[cj]; p < x->colptr[cj + 1]; p++) {
      int r = x->rowidx[p];
      for (int t = ri_ptr[r]; t < ri_ptr[r + 1]; t++) {
        y->rowidx[q] = ri_idx[t];
        y->val[q] = x->val[p];
        q++;
      }
    }
    /* setFromTriplets sorts rows inside a column (stable insertion sort; the
     * segments are short and already sorted when R is ascending). */
    for (int a = cnt[j] + 1; a < q; a++) {
      int rr = y->rowidx[a];
      double vv = y->val[a];
      int b = a - 1;
      while (b >= cnt[j] && y->rowidx[b] > rr) {
        y->rowidx[b + 1] = y->rowidx[b];
        y->val[b + 1] = y->val[b];
        b--;
      }
      y->rowidx[b + 1] = rr;
      y->val[b + 1] = vv;
    }
  }
  free(cnt);
  free(ri_ptr);
  free(ri_idx);
  return y;
}

/* Eigen conservative_sparse_sparse_product (col-major x col-major -> col-major).
 * res(:,j) = sum_k lhs(:,k)*rhs(k,j), k in storage (ascending) order of rhs
 * column j; first touch assigns, later touches add; rows sorted afterwards. */
static orc_csc *csc_spgemm(const orc_csc *l, const orc_csc *r) {
  int rows = l->rows, cols = r->cols;
  int *mask = (int *)malloc(sizeof(int) * (size_t)(rows > 0 ? rows : 1));
  double *acc = (double *)malloc(sizeof(double) * (size_t)(rows > 0 ? rows : 1));
  int *touched = (int *)malloc(sizeof(int) * (size_t)(rows > 0 ? rows : 1));
  for (int i = 0; i < rows; i++) mask[i] = -1;
  /* pass 1: count */
  int *cp = (int *)calloc((size_t)cols + 1, sizeof(int));
  for (int j = 0; j < cols; j++) {
    int c = 0;
    for (int p = r->colptr[j]; p < r->colptr[j + 1]; p++) {
      int k = r->rowidx[p];
      for (int q = l->colptr[k]; q < l->colptr[k + 1]; q++) {
        int i = l->rowidx[q];
        if (mask[i] != j) {
          mask[i] = j;
          c++;
        }
      }
    }
    cp[j + 1] = cp[j] + c;
  }
  orc_csc *res = csc_alloc(rows, cols, cp[cols]);
  memcpy(res->colptr, cp, sizeof(int) * ((size_t)cols + 1));
  for (int i = 0; i < rows; i++) mask[i] = -1;
  for (int j = 0; j < cols; j++) {
    int nt = 0;
    for (int p = r->colptr[j]; p < r->colptr[j + 1]; p++) {
      int k = r->rowidx[p];
      double y = r->val[p];
      for (int q = l->colptr[k]; q < l->colptr[k + 1]; q++) {
        int i = l->rowidx[q];
        double x = l->val[q];
        if (mask[i] != j) {
          mask[i] = j;
          acc[i] = x * y;
          touched[nt++] = i;
        } else {
          acc[i] += x * y;
        }
      }
    }
    /* sort touched rows ascending (insertion sort; nt is small) */
    for (int a = 1; a < nt; a++) {
      int t = touched[a], b = a - 1;
      while (b >= 0 && touched[b] > t) {
        touched[b + 1] = touched[b];
        b--;
      }
      touched[b + 1] = t;
    }
    int base = cp[j];
    for (int a = 0; a < nt; a++) {
      res->rowidx[base + a] = touched[a];
      res->val[base + a] = acc[touched[a]];
    }
  }
  free(cp);
  free(mask);
  free(acc);
  free(touched);
  return res;
}

/* Eigen col-major sparse * dense (n x k, col-major): y = A*x. */
static void csc_spmv(const orc_csc *a, const double *x, int k, double *y) {
  for (int c = 0; c < k; c++) {
    const double *xc = x + (size_t)c * a->cols;
    double *yc = y + (size_t)c * a->rows;
    for (int i = 0; i < a->rows; i++) yc[i] = 0.0;
    for (int j = 0; j < a->cols; j++) {
      double xj = xc[j];
      for (int p = a->colptr[j]; p < a->colptr[j + 1]; p++) yc[a->rowidx[p]] += a->val[p] * xj;
    }
  }
}

/* ------------------------------------------------------------------------- */
/* coarse solver: RCM ordering + banded Cholesky (stands in for               */
/* Eigen::SimplicialLDLT, mg_VCycle.cpp:199, min_quad_with_fixed_mg.cpp:48,254)*/
/* ------------------------------------------------------------------------- */
typedef struct {
  int n, bw;
  int *perm;  /* new -> old */
  double *L;  /* n x (bw+1), row i holds L(i, i-bw .. i), diagonal at [bw] */
  double *tmp;
} orc_band;

static void band_free(orc_band *b) {
  if (!b) return;
  free(b->perm);
  free(b->L);
  free(b->tmp);
  free(b);
}

static int cmp_int_pair(const void *a, const void *b) {
  const int *x = (const int *)a, *y = (const int *)b;
  if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
  return x[1] < y[1] ? -1 : (x[1] > y[1]);
}

static orc_band *band_factor(const orc_csc *a) {
  int n = a->rows;
  orc_band *b = (orc_band *)calloc(1, sizeof(orc_band));
  b->n = n;
  b->perm = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  b->tmp = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  /* Cuthill-McKee BFS from a minimum-degree vertex, neighbours by degree */
  int *inv = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; i++) inv[i] = -1;
  int cnt = 0;
  int *pairs = (int *)malloc(sizeof(int) * 2 * 4096);
  int pcap = 4096;
  while (cnt < n) {
    int start = -1, best = 1 << 30;
    for (int i = 0; i < n; i++)
      if (inv[i] < 0) {
        int d = a->colptr[i + 1] - a->colptr[i];
        if (d < best) {
          best = d;
          start = i;
        }
      }
    inv[start] = cnt;
    b->perm[cnt++] = start;
    for (int head = cnt - 1; head < cnt; head++) {
      int v = b->perm[head];
      int np = 0;
      for (int p = a->colptr[v]; p < a->colptr[v + 1]; p++) {
        int w = a->rowidx[p];
        if (inv[w] < 0) {
          if (np >= pcap) {
            pcap *= 2;
            pairs = (int *)realloc(pairs, sizeof(int) * 2 * (size_t)pcap);
          }
          pairs[2 * np] = a->colptr[w + 1] - a->colptr[w];
          pairs[2 * np + 1] = w;
          np++;
          inv[w] = -2; /* queued */
        }
      }
      qsort(pairs, (size_t)np, 2 * sizeof(int), cmp_int_pair);
      for (int t = 0; t < np; t++) {
        inv[pairs[2 * t + 1]] = cnt;
        b->perm[cnt++] = pairs[2 * t + 1];
      }
    }
  }
  free(pairs);
  /* reverse */
  for (int i = 0; i < n / 2; i++) {
    int t = b->perm[i];
    b->perm[i] = b->perm[n - 1 - i];
    b->perm[n - 1 - i] = t;
  }
  for (int i = 0; i < n; i++) inv[b->perm[i]] = i;
  int bw = 0;
  for (int j = 0; j < n; j++)
    for (int p = a->colptr[j]; p < a->colptr[j + 1]; p++) {
      int d = inv[a->rowidx[p]] - inv[j];
      if (d < 0) d = -d;
      if (d > bw) bw = d;
    }
  b->bw = bw;
  size_t w = (size_t)bw + 1;
  b->L = (double *)calloc((size_t)(n > 0 ? n : 1) * w, sizeof(double));
  for (int j = 0; j < n; j++)
    for (int p = a->colptr[j]; p < a->colptr[j + 1]; p++) {
      int r = inv[a->rowidx[p]], c = inv[j];
      if (c <= r) b->L[(size_t)r * w + (size_t)(bw - (r - c))] += a->val[p];
    }
  free(inv);
  /* banded Cholesky, row by row */
  for (int i = 0; i < n; i++) {
    double *Li = b->L + (size_t)i * w;
    int j0 = i - bw < 0 ? 0 : i - bw;
    for (int j = j0; j <= i; j++) {
      double *Lj = b->L + (size_t)j * w;
      double s = Li[bw - (i - j)];
      int k0 = j - bw < 0 ? 0 : j - bw;
      if (k0 < j0) k0 = j0;
      for (int k = k0; k < j; k++) s -= Li[bw - (i - k)] * Lj[bw - (j - k)];
      if (j < i)
        Li[bw - (i - j)] = s / Lj[bw];
      else
        Li[bw] = sqrt(s);
    }
  }
  return b;
}

static void band_solve(const orc_band *b, const double *rhs, double *x) {
  int n = b->n, bw = b->bw;
  size_t w = (size_t)bw + 1;
  double *y = b->tmp;
  for (int i = 0; i < n; i++) {
    const double *Li = b->L + (size_t)i * w;
    double s = rhs[b->perm[i]];
    int j0 = i - bw < 0 ? 0 : i - bw;
    for (int j = j0; j < i; j++) s -= Li[bw - (i - j)] * y[j];
    y[i] = s / Li[bw];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = y[i];
    int j1 = i + bw >= n ? n - 1 : i + bw;
    for (int j = i + 1; j <= j1; j++) s -= b->L[(size_t)j * w + (size_t)(bw - (j - i))] * y[j];
    y[i] = s / b->L[(size_t)i * w + (size_t)bw];
  }
  for (int i = 0; i < n; i++) x[b->perm[i]] = y[i];
}

/* ------------------------------------------------------------------------- */
/* hierarchy (src/mg_data.h:11-44) and solver data                            */
/* (src/min_quad_with_fixed_mg.h:22-29)                                       */
/* ------------------------------------------------------------------------- */
typedef struct {
  orc_csc *P_full, *A, *P, *PT;
  double *A_diag;
  int *keep; /* kept columns of P at this level (fixed variant), or NULL */
  int nkeep;
} orc_level;

typedef struct {
  int nlev;
  orc_level *mg;
  int n;
  int *known, nknown;
  int *unknown, nunknown;
  orc_csc *LHS, *Auk;
  orc_band *solver;
  int has_fixed;
} orc_solver;

orc_solver *orc_create(int nlev) {
  orc_solver *s = (orc_solver *)calloc(1, sizeof(orc_solver));
  s->nlev = nlev;
  s->mg = (orc_level *)calloc((size_t)nlev, sizeof(orc_level));
  return s;
}

static void level_clear_derived(orc_level *l) {
  csc_free(l->A);
  l->A = NULL;
  free(l->A_diag);
  l->A_diag = NULL;
  free(l->keep);
  l->keep = NULL;
  l->nkeep = 0;
}

void orc_destroy(orc_solver *s) {
  if (!s) return;
  for (int i = 0; i < s->nlev; i++) {
    level_clear_derived(&s->mg[i]);
    csc_free(s->mg[i].P_full);
    csc_free(s->mg[i].P);
    csc_free(s->mg[i].PT);
  }
  free(s->mg);
  free(s->known);
  free(s->unknown);
  csc_free(s->LHS);
  csc_free(s->Auk);
  band_free(s->solver);
  free(s);
}

/* What mg_precompute leaves behind for level lv>=1 (src/mg_precompute.cpp:71-77):
 * P = P_full = the prolongation, PT = P^T. */
int orc_set_prolongation(orc_solver *s, int lv, int rows, int cols, const int *colptr,
                         const int *rowidx, const double *val) {
  if (lv < 1 || lv >= s->nlev) return -1;
  orc_level *l = &s->mg[lv];
  csc_free(l->P_full);
  csc_free(l->P);
  csc_free(l->PT);
  l->P_full = csc_copy_raw(rows, cols, colptr, rowidx, val);
  l->P = csc_clone(l->P_full);
  l->PT = csc_transpose(l->P_full);
  return 0;
}

static void finish_precompute(orc_solver *s) {
  /* src/min_quad_with_fixed_mg.cpp:31-48 / :237-254 */
  int last = s->nlev - 1;
  orc_csc *Ac = s->mg[last].A;
  for (int ii = 0; ii < Ac->rows; ii++) {
    int found = 0;
    for (int p = Ac->colptr[ii]; p < Ac->colptr[ii + 1]; p++)
      if (Ac->rowidx[p] == ii) {
        Ac->val[p] += 1e-12;
        found = 1;
        break;
      }
    if (!found) { /* coeffRef would insert; never happens for SPD input */
      fprintf(stderr, "orc: missing diagonal entry %d on coarsest level\n", ii);
    }
  }
  for (int lv = 0; lv < s->nlev; lv++) {
    orc_csc *A = s->mg[lv].A;
    free(s->mg[lv].A_diag);
    s->mg[lv].A_diag = (double *)calloc((size_t)(A->rows > 0 ? A->rows : 1), sizeof(double));
    for (int j = 0; j < A->cols; j++)
      for (int p = A->colptr[j]; p < A->colptr[j + 1]; p++)
        if (A->rowidx[p] == j) s->mg[lv].A_diag[j] = A->val[p];
  }
  band_free(s->solver);
  s->solver = band_factor(Ac);
}

/* min_quad_with_fixed_mg_precompute.  nknown < 0 selects the variant without
 * fixed values (src/min_quad_with_fixed_mg.cpp:3-51); otherwise the variant with
 * fixed values (:137-257). */
int orc_precompute(orc_solver *s, int n, const int *colptr, const int *rowidx, const double *val,
                   const int *known, int nknown) {
  if (s->nlev < 2) return -2; /* mg_precompute.cpp:39 TODO: single level unsupported */
  for (int lv = 0; lv < s->nlev; lv++) level_clear_derived(&s->mg[lv]);
  csc_free(s->LHS);
  csc_free(s->Auk);
  s->LHS = s->Auk = NULL;
  free(s->known);
  free(s->unknown);
  s->known = s->unknown = NULL;
  orc_csc *A = csc_copy_raw(n, n, colptr, rowidx, val);
  s->n = n;
  if (nknown < 0) {
    s->has_fixed = 0;
    s->nknown = 0;
    s->nunknown = n;
    s->LHS = A;
    s->mg[0].A = csc_clone(A);
    for (int lv = 1; lv < s->nlev; lv++) {
      orc_csc *t = csc_spgemm(s->mg[lv].PT, s->mg[lv - 1].A);
      s->mg[lv].A = csc_spgemm(t, s->mg[lv].P);
      csc_free(t);
    }
    finish_precompute(s);
    return 0;
  }
  s->has_fixed = 1;
  s->nknown = nknown;
  s->known = (int *)malloc(sizeof(int) * (size_t)(nknown > 0 ? nknown : 1));
  memcpy(s->known, known, sizeof(int) * (size_t)nknown);
  s->unknown = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  s->nunknown = setdiff_range(n, known, nknown, s->unknown); /* :156-158 */
  s->LHS = csc_slice(A, s->unknown, s->nunknown, s->unknown, s->nunknown); /* :167 */
  s->Auk = csc_slice(A, s->unknown, s->nunknown, s->known, s->nknown);     /* :170 */
  csc_free(A);
  /* :185  mg[1].P = P_full(unknown,:) */
  csc_free(s->mg[1].P);
  s->mg[1].P = csc_slice(s->mg[1].P_full, s->unknown, s->nunknown, NULL, 0);
  for (int lv = 1; lv < s->nlev; lv++) { /* :186-220 */
    orc_csc *P = s->mg[lv].P;
    int *keep = (int *)malloc(sizeof(int) * (size_t)(P->cols > 0 ? P->cols : 1));
    int nkeep = 0;
    for (int c = 0; c < P->cols; c++)
      for (int p = P->colptr[c]; p < P->colptr[c + 1]; p++)
        if (P->val[p] > 1e-15) {
          keep[nkeep++] = c;
          break;
        }
    if (nkeep < P->cols) {
      s->mg[lv].P = csc_slice(P, NULL, 0, keep, nkeep);
      csc_free(P);
      s->mg[lv].keep = keep;
      s->mg[lv].nkeep = nkeep;
      if (lv < s->nlev - 1) {
        csc_free(s->mg[lv + 1].P);
        s->mg[lv + 1].P = csc_slice(s->mg[lv + 1].P_full, keep, nkeep, NULL, 0);
      }
    } else {
      free(keep);
      break;
    }
  }
  s->mg[0].A = csc_clone(s->LHS); /* :223 */
  for (int lv = 1; lv < s->nlev; lv++) {
    csc_free(s->mg[lv].PT);
    s->mg[lv].PT = csc_transpose(s->mg[lv].P); /* :226 */
    orc_csc *t = csc_spgemm(s->mg[lv].PT, s->mg[lv - 1].A);
    s->mg[lv].A = csc_spgemm(t, s->mg[lv].P); /* :227 (left to right) */
    csc_free(t);
  }
  finish_precompute(s);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* V-cycle pieces (src/mg_VCycle.cpp)                                         */
/* ------------------------------------------------------------------------- */

/* relax, mg_VCycle.cpp:113-178: forward lexicographic Gauss-Seidel, reading
 * column i of the (symmetric) CSC matrix as row i, skipping the diagonal entry,
 * true division by A_diag; columns of u swept one after another per iteration. */
void orc_relax(const orc_solver *s, int lv, int iters, const double *B, double *u, int k) {
  const orc_csc *A = s->mg[lv].A;
  const double *d = s->mg[lv].A_diag;
  int n = A->rows;
  for (int iter = 0; iter < iters; iter++)
    for (int ri = 0; ri < k; ri++) {
      const double *b = B + (size_t)ri * n;
      double *x = u + (size_t)ri * n;
      for (int c = 0; c < n; c++) {
        double sum = 0;
        for (int p = A->colptr[c]; p < A->colptr[c + 1]; p++) {
          int r = A->rowidx[p];
          if (r != c) sum += A->val[p] * x[r];
        }
        x[c] = (b[c] - sum) / d[c];
      }
    }
}

/* A(), mg_VCycle.cpp:62-70 */
void orc_apply_A(const orc_solver *s, int lv, const double *u, double *Au, int k) {
  csc_spmv(s->mg[lv].A, u, k, Au);
}
/* restrict(), mg_VCycle.cpp:72-81: Rx = mg[lv+1].PT * x */
void orc_restrict(const orc_solver *s, int lv, const double *x, double *Rx, int k) {
  csc_spmv(s->mg[lv + 1].PT, x, k, Rx);
}
/* prolong(), mg_VCycle.cpp:83-92: Px = mg[lv+1].P * x */
void orc_prolong(const orc_solver *s, int lv, const double *x, double *Px, int k) {
  csc_spmv(s->mg[lv + 1].P, x, k, Px);
}
/* coarseSolve(), mg_VCycle.cpp:181-201: u = u + solver.solve(B) */
void orc_coarse_solve(const orc_solver *s, const double *B, double *u, int k) {
  int n = s->solver->n;
  double *d = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  for (int c = 0; c < k; c++) {
    band_solve(s->solver, B + (size_t)c * n, d);
    for (int i = 0; i < n; i++) u[(size_t)c * n + i] = u[(size_t)c * n + i] + d[i];
  }
  free(d);
}

/* mg_VCycle, mg_VCycle.cpp:3-59 */
void orc_vcycle(const orc_solver *s, const double *B, int pre, int post, int lv, double *u,
                int k) {
  if (lv == s->nlev - 1) {
    orc_coarse_solve(s, B, u, k);
    return;
  }
  int n = s->mg[lv].A->rows;
  int nc = s->mg[lv + 1].PT->rows;
  orc_relax(s, lv, pre, B, u, k);
  double *Au = (double *)malloc(sizeof(double) * (size_t)n * k + 8);
  double *r = (double *)malloc(sizeof(double) * (size_t)n * k + 8);
  orc_apply_A(s, lv, u, Au, k);
  for (size_t i = 0; i < (size_t)n * k; i++) r[i] = B[i] - Au[i];
  double *rc = (double *)malloc(sizeof(double) * (size_t)nc * k + 8);
  orc_restrict(s, lv, r, rc, k);
  double *uc = (double *)calloc((size_t)nc * k + 1, sizeof(double));
  orc_vcycle(s, rc, pre, post, lv + 1, uc, k);
  double *puc = Au; /* reuse */
  orc_prolong(s, lv, uc, puc, k);
  for (size_t i = 0; i < (size_t)n * k; i++) u[i] = u[i] + puc[i];
  orc_relax(s, lv, post, B, u, k);
  free(Au);
  free(r);
  free(rc);
  free(uc);
}

static double residual_norm(const orc_csc *A, const double *rhs, const double *z, int k,
                            double *tmp) {
  csc_spmv(A, z, k, tmp);
  double ss = 0;
  for (size_t i = 0; i < (size_t)A->rows * k; i++) {
    double d = rhs[i] - tmp[i];
    ss += d * d;
  }
  return sqrt(ss);
}

/* min_quad_with_fixed_mg_solve.
 *  free variant  : src/min_quad_with_fixed_mg.cpp:80-135
 *  fixed variant : src/min_quad_with_fixed_mg.cpp:288-361
 * RHS, z0, z are n x k col-major; known_val is nknown x k col-major.
 * Returns 1 when !(residual > tol) with the *last measured* residual. */
int orc_solve(const orc_solver *s, const double *RHS, const double *known_val, const double *z0,
              int k, double tol, int max_iter, double *z, double *r_his, int *n_his) {
  int n = s->n, nu = s->nunknown;
  double *zu = (double *)malloc(sizeof(double) * (size_t)nu * k + 8);
  double *bu = (double *)malloc(sizeof(double) * (size_t)nu * k + 8);
  double *tmp = (double *)malloc(sizeof(double) * (size_t)nu * k + 8);
  if (s->has_fixed) {
    for (int c = 0; c < k; c++)
      for (int i = 0; i < nu; i++) {
        zu[(size_t)c * nu + i] = z0[(size_t)c * n + s->unknown[i]];
        bu[(size_t)c * nu + i] = RHS[(size_t)c * n + s->unknown[i]];
      }
    if (s->nknown > 0) {
      csc_spmv(s->Auk, known_val, k, tmp);
      for (size_t i = 0; i < (size_t)nu * k; i++) bu[i] = bu[i] - tmp[i];
    } /* Auk has 0 columns -> product is exactly zero; x - 0 == x */
  } else {
    memcpy(zu, z0, sizeof(double) * (size_t)n * k);
    memcpy(bu, RHS, sizeof(double) * (size_t)n * k);
  }
  double residual = 0;
  int nh = 0;
  for (int iter = 0; iter < max_iter; iter++) {
    residual = residual_norm(s->mg[0].A, bu, zu, k, tmp);
    r_his[nh++] = residual;
    if (residual < tol) break;
    orc_vcycle(s, bu, 2, 2, 0, zu, k);
  }
  *n_his = nh;
  if (s->has_fixed) {
    for (int c = 0; c < k; c++) {
      for (int i = 0; i < nu; i++) z[(size_t)c * n + s->unknown[i]] = zu[(size_t)c * nu + i];
      for (int i = 0; i < s->nknown; i++)
        z[(size_t)c * n + s->known[i]] = known_val[(size_t)c * s->nknown + i];
    }
  } else {
    memcpy(z, zu, sizeof(double) * (size_t)n * k);
  }
  free(zu);
  free(bu);
  free(tmp);
  return residual > tol ? 0 : 1;
}

/* Timed variant for the CPU baseline: runs exactly `cycles` V(2,2)-cycles with
 * one residual measurement before each (same work per iteration as the solve
 * loop, no early exit). Works on the unknown-sized system directly. */
void orc_iterate(const orc_solver *s, const double *bu, double *zu, int k, int cycles,
                 double *r_his) {
  int nu = s->nunknown;
  double *tmp = (double *)malloc(sizeof(double) * (size_t)nu * k + 8);
  for (int it = 0; it < cycles; it++) {
    r_his[it] = residual_norm(s->mg[0].A, bu, zu, k, tmp);
    orc_vcycle(s, bu, 2, 2, 0, zu, k);
  }
  free(tmp);
}

/* ------------------------------------------------------------------------- */
/* Multi-threaded CPU variant (OpenMP) -- NOT the reference algorithm.          */
/* The reference path is single-threaded lexicographic Gauss-Seidel.  For        */
/* context only (bench.py reports it beside the reference arm, labelled), this   */
/* runs the same V(2,2) iteration the way a CPU would be used in parallel:       */
/* greedy multicolour Gauss-Seidel (rows of one colour in parallel) and          */
/* row-parallel products (A read column-as-row like the smoother does, P / PT    */
/* through each other's CSC).  Same fixed point, different sweep order.          */
/* ------------------------------------------------------------------------- */
typedef struct {
  int ncol;
  int *ptr;  /* ncol + 1 */
  int *rows; /* n, grouped by colour */
} orc_colouring;

static orc_colouring *colour_greedy(const orc_csc *A) {
  int n = A->rows;
  orc_colouring *c = (orc_colouring *)calloc(1, sizeof(orc_colouring));
  int *col = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  int *stamp = (int *)malloc(sizeof(int) * 64);
  int cap = 64, nc = 0;
  for (int i = 0; i < cap; i++) stamp[i] = -1;
  for (int v = 0; v < n; v++) col[v] = -1;
  for (int v = 0; v < n; v++) {
    for (int p = A->colptr[v]; p < A->colptr[v + 1]; p++) {
      int w = A->rowidx[p];
      if (w != v && col[w] >= 0) stamp[col[w]] = v;
    }
    int k = 0;
    while (k < nc && stamp[k] == v) k++;
    if (k == nc) {
      if (nc == cap) {
        cap *= 2;
        stamp = (int *)realloc(stamp, sizeof(int) * (size_t)cap);
        for (int i = nc; i < cap; i++) stamp[i] = -1;
      }
      nc++;
    }
    col[v] = k;
  }
  c->ncol = nc;
  c->ptr = (int *)calloc((size_t)nc + 1, sizeof(int));
  c->rows = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  for (int v = 0; v < n; v++) c->ptr[col[v] + 1]++;
  for (int k = 0; k < nc; k++) c->ptr[k + 1] += c->ptr[k];
  int *nx = (int *)malloc(sizeof(int) * (size_t)(nc > 0 ? nc : 1));
  for (int k = 0; k < nc; k++) nx[k] = c->ptr[k];
  for (int v = 0; v < n; v++) c->rows[nx[col[v]]++] = v;
  free(nx);
  free(col);
  free(stamp);
  return c;
}

static void colouring_free(orc_colouring *c) {
  if (!c) return;
  free(c->ptr);
  free(c->rows);
  free(c);
}

/* y = M^T-as-rows x: y[j] = sum over the stored entries of CSC column j of val * x[row] */
static void mt_gather_cols(const orc_csc *M, const double *x, int ldx, double *y, int ldy, int k) {
  for (int c = 0; c < k; c++) {
    const double *xc = x + (size_t)c * ldx;
    double *yc = y + (size_t)c * ldy;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < M->cols; j++) {
      double s = 0;
      for (int p = M->colptr[j]; p < M->colptr[j + 1]; p++) s += M->val[p] * xc[M->rowidx[p]];
      yc[j] = s;
    }
  }
}

static void mt_relax(const orc_csc *A, const double *d, const orc_colouring *col, int iters,
                     const double *B, double *u, int k) {
  int n = A->rows;
  for (int it = 0; it < iters; it++)
    for (int c = 0; c < k; c++) {
      const double *b = B + (size_t)c * n;
      double *x = u + (size_t)c * n;
      for (int q = 0; q < col->ncol; q++) {
#pragma omp parallel for schedule(static)
        for (int t = col->ptr[q]; t < col->ptr[q + 1]; t++) {
          int i = col->rows[t];
          double s = 0;
          for (int p = A->colptr[i]; p < A->colptr[i + 1]; p++) {
            int r = A->rowidx[p];
            if (r != i) s += A->val[p] * x[r];
          }
          x[i] = (b[i] - s) / d[i];
        }
      }
    }
}

static void mt_vcycle(const orc_solver *s, orc_colouring **cols, const double *B, int lv, double *u,
                      int k) {
  if (lv == s->nlev - 1) {
    orc_coarse_solve(s, B, u, k);
    return;
  }
  const orc_csc *A = s->mg[lv].A;
  int n = A->rows, nc = s->mg[lv + 1].PT->rows;
  mt_relax(A, s->mg[lv].A_diag, cols[lv], 2, B, u, k);
  double *r = (double *)malloc(sizeof(double) * (size_t)n * k + 8);
  mt_gather_cols(A, u, n, r, n, k);
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)n * k; i++) r[i] = B[i] - r[i];
  double *rc = (double *)malloc(sizeof(double) * (size_t)nc * k + 8);
  mt_gather_cols(s->mg[lv + 1].P, r, n, rc, nc, k); /* PT r: row c of PT = column c of P */
  double *uc = (double *)calloc((size_t)nc * k + 1, sizeof(double));
  mt_vcycle(s, cols, rc, lv + 1, uc, k);
  mt_gather_cols(s->mg[lv + 1].PT, uc, nc, r, n, k); /* P uc: row f of P = column f of PT */
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)n * k; i++) u[i] = u[i] + r[i];
  mt_relax(A, s->mg[lv].A_diag, cols[lv], 2, B, u, k);
  free(r);
  free(rc);
  free(uc);
}

/* `cycles` x (residual norm + multicolour V(2,2)) on the unknown-sized system with `threads`
 * OpenMP threads (<= 0: the runtime's default).  Returns the number of threads used. */
int orc_iterate_mt(const orc_solver *s, const double *bu, double *zu, int k, int cycles,
                   double *r_his, int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
  int used = omp_get_max_threads();
#else
  int used = 1;
  (void)threads;
#endif
  orc_colouring **cols = (orc_colouring **)calloc((size_t)s->nlev, sizeof(orc_colouring *));
  for (int lv = 0; lv + 1 < s->nlev; lv++) cols[lv] = colour_greedy(s->mg[lv].A);
  const orc_csc *A = s->mg[0].A;
  int n = A->rows;
  double *tmp = (double *)malloc(sizeof(double) * (size_t)n * k + 8);
  for (int it = 0; it < cycles; it++) {
    mt_gather_cols(A, zu, n, tmp, n, k);
    double ss = 0;
#pragma omp parallel for schedule(static) reduction(+ : ss)
    for (long i = 0; i < (long)n * k; i++) {
      double d = bu[i] - tmp[i];
      ss += d * d;
    }
    r_his[it] = sqrt(ss);
    mt_vcycle(s, cols, bu, 0, zu, k);
  }
  free(tmp);
  for (int lv = 0; lv < s->nlev; lv++) colouring_free(cols[lv]);
  free(cols);
  return used;
}

/* ------------------------------------------------------------------------- */
/* getters for tests                                                          */
/* ------------------------------------------------------------------------- */
int orc_num_levels(const orc_solver *s) { return s->nlev; }
int orc_num_unknown(const orc_solver *s) { return s->nunknown; }
void orc_get_unknown(const orc_solver *s, int *out) {
  if (s->has_fixed)
    memcpy(out, s->unknown, sizeof(int) * (size_t)s->nunknown);
  else
    for (int i = 0; i < s->n; i++) out[i] = i;
}
int orc_get_keep(const orc_solver *s, int lv, int *out) {
  if (!s->mg[lv].keep) return -1;
  if (out) memcpy(out, s->mg[lv].keep, sizeof(int) * (size_t)s->mg[lv].nkeep);
  return s->mg[lv].nkeep;
}
/* which: 0 = A, 1 = P, 2 = PT, 3 = LHS, 4 = Auk */
static const orc_csc *pick(const orc_solver *s, int lv, int which) {
  switch (which) {
    case 0: return s->mg[lv].A;
    case 1: return s->mg[lv].P;
    case 2: return s->mg[lv].PT;
    case 3: return s->LHS;
    case 4: return s->Auk;
  }
  return NULL;
}
int orc_matrix_dims(const orc_solver *s, int lv, int which, int *rows, int *cols, int *nnz) {
  const orc_csc *m = pick(s, lv, which);
  if (!m) return -1;
  *rows = m->rows;
  *cols = m->cols;
  *nnz = m->colptr[m->cols];
  return 0;
}
int orc_matrix_copy(const orc_solver *s, int lv, int which, int *colptr, int *rowidx,
                    double *val) {
  const orc_csc *m = pick(s, lv, which);
  if (!m) return -1;
  int nnz = m->colptr[m->cols];
  memcpy(colptr, m->colptr, sizeof(int) * ((size_t)m->cols + 1));
  memcpy(rowidx, m->rowidx, sizeof(int) * (size_t)nnz);
  memcpy(val, m->val, sizeof(double) * (size_t)nnz);
  return 0;
}
void orc_get_diag(const orc_solver *s, int lv, double *out) {
  memcpy(out, s->mg[lv].A_diag, sizeof(double) * (size_t)s->mg[lv].A->rows);
}
int orc_coarse_bandwidth(const orc_solver *s) { return s->solver ? s->solver->bw : -1; }
