// ref_shim_selftest.cpp -- C entry points over the Eigen stand-in (oracle/ref_shim) so that
// tests/test_ref_shim.py can check ITS arithmetic and patterns directly against scipy:
// sparse * sparse (conservative: structural zeros kept), sparse * dense, transpose,
// setFromTriplets (duplicates summed, explicit zeros kept), coeffRef (insertion), diagonal,
// and the SimplicialLDLT stand-in.  TEST INFRASTRUCTURE.
#include <Eigen/Sparse>

#include <cstring>
#include <vector>

namespace {
typedef Eigen::SparseMatrix<double> SpMat;
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
SpMat g_result;  // the last sparse result, read back with shim_result_*
}  // namespace

extern "C" {
int shim_result_nnz(void) { return static_cast<int>(g_result.nonZeros()); }
void shim_result_dims(int* rows, int* cols) {
  *rows = static_cast<int>(g_result.rows());
  *cols = static_cast<int>(g_result.cols());
}
void shim_result_copy(int* colptr, int* rowidx, double* val) {
  std::memcpy(colptr, g_result.outerIndexPtr(), sizeof(int) * (static_cast<size_t>(g_result.cols()) + 1));
  if (g_result.nonZeros() > 0) {
    std::memcpy(rowidx, g_result.innerIndexPtr(), sizeof(int) * static_cast<size_t>(g_result.nonZeros()));
    std::memcpy(val, g_result.valuePtr(), sizeof(double) * static_cast<size_t>(g_result.nonZeros()));
  }
}
// C = A * B ; with three operands C = A * B * D evaluated left to right like `PT * A * P`
void shim_spgemm(int ar, int ac, const int* ap, const int* ai, const double* av, int bc, const int* bp,
                 const int* bi, const double* bv) {
  g_result = from_csc(ar, ac, ap, ai, av) * from_csc(ac, bc, bp, bi, bv);
}
void shim_triple(int ar, int ac, const int* ap, const int* ai, const double* av, int bc, const int* bp,
                 const int* bi, const double* bv, int dc, const int* dp, const int* di, const double* dv) {
  g_result = from_csc(ar, ac, ap, ai, av) * from_csc(ac, bc, bp, bi, bv) * from_csc(bc, dc, dp, di, dv);
}
void shim_transpose(int ar, int ac, const int* ap, const int* ai, const double* av) {
  g_result = from_csc(ar, ac, ap, ai, av).transpose();
}
void shim_spmm(int ar, int ac, const int* ap, const int* ai, const double* av, const double* x, int k, double* y) {
  Eigen::MatrixXd X(ac, k);
  std::memcpy(X.data(), x, sizeof(double) * static_cast<size_t>(ac) * k);
  Eigen::MatrixXd Y = from_csc(ar, ac, ap, ai, av) * X;
  std::memcpy(y, Y.data(), sizeof(double) * static_cast<size_t>(ar) * k);
}
void shim_from_triplets(int rows, int cols, int n, const int* r, const int* c, const double* v) {
  std::vector<Eigen::Triplet<double>> t;
  for (int i = 0; i < n; i++) t.emplace_back(r[i], c[i], v[i]);
  g_result = SpMat(rows, cols);
  g_result.setFromTriplets(t.begin(), t.end());
}
void shim_coeffref_add(int ar, int ac, const int* ap, const int* ai, const double* av, int n, const int* r,
                       const int* c, const double* d, double* diag_out) {
  g_result = from_csc(ar, ac, ap, ai, av);
  for (int i = 0; i < n; i++) g_result.coeffRef(r[i], c[i]) += d[i];
  Eigen::VectorXd dg = g_result.diagonal();
  std::memcpy(diag_out, dg.data(), sizeof(double) * static_cast<size_t>(dg.size()));
}
int shim_ldlt_solve(int n, const int* ap, const int* ai, const double* av, const double* b, int k, double* x) {
  Eigen::SimplicialLDLT<SpMat> s;
  s.compute(from_csc(n, n, ap, ai, av));
  if (!s.ok()) return -1;
  Eigen::MatrixXd B(n, k);
  std::memcpy(B.data(), b, sizeof(double) * static_cast<size_t>(n) * k);
  Eigen::MatrixXd X = s.solve(B);
  std::memcpy(x, X.data(), sizeof(double) * static_cast<size_t>(n) * k);
  return 0;
}
}
