// igl/slice.h -- restatement of the overloads of igl::slice that the reference's hot path
// calls (libigl/include/igl/slice.cpp:13-153), for the Eigen stand-in of this directory.
// TEST INFRASTRUCTURE.  The vendored libigl headers themselves need far more of Eigen
// (MatrixBase expressions, DynamicSparseMatrix, LinSpaced, unique/sort) than the shim has.
#ifndef SMG_REF_SHIM_IGL_SLICE
#define SMG_REF_SHIM_IGL_SLICE
#include <Eigen/Core>
#include <Eigen/Sparse>

#include <vector>

namespace igl {
// Y = X(R, C), sparse.  Semantics of libigl's overload (slice.cpp:13-77): every stored entry
// X(r, c) - explicit zeros included - lands on every output position (i, j) with R(i) == r and
// C(j) == c; duplicates created by repeated indices are summed and every output column is sorted
// (that is what its setFromTriplets call does).  Implemented here with bucketed index lists.
template <typename TX, typename TY, typename DerivedR, typename DerivedC>
inline void slice(const Eigen::SparseMatrix<TX>& X, const Eigen::DenseBase<DerivedR>& R,
                  const Eigen::DenseBase<DerivedC>& C, Eigen::SparseMatrix<TY>& Y) {
  const Eigen::Index out_rows = R.size(), out_cols = C.size();
  Y.resize(out_rows, out_cols);
  if (out_rows == 0 || out_cols == 0) return;
  // targets_of_row[r] = output rows that read input row r, as one flat list with offsets
  std::vector<int> row_ofs(static_cast<size_t>(X.rows()) + 1, 0), row_tgt(static_cast<size_t>(out_rows));
  for (Eigen::Index i = 0; i < out_rows; i++) row_ofs[static_cast<size_t>(R(i)) + 1]++;
  for (size_t r = 0; r + 1 < row_ofs.size(); r++) row_ofs[r + 1] += row_ofs[r];
  {
    std::vector<int> fill(row_ofs.begin(), row_ofs.end() - 1);
    for (Eigen::Index i = 0; i < out_rows; i++) row_tgt[static_cast<size_t>(fill[static_cast<size_t>(R(i))]++)] = static_cast<int>(i);
  }
  std::vector<Eigen::Triplet<TY>> trip;
  for (Eigen::Index j = 0; j < out_cols; j++) {
    const Eigen::Index src_col = C(j);
    for (typename Eigen::SparseMatrix<TX>::InnerIterator e(X, src_col); e; ++e)
      for (int t = row_ofs[static_cast<size_t>(e.row())]; t < row_ofs[static_cast<size_t>(e.row()) + 1]; t++)
        trip.emplace_back(row_tgt[static_cast<size_t>(t)], static_cast<int>(j), static_cast<TY>(e.value()));
  }
  Y.setFromTriplets(trip.begin(), trip.end());
}

// Y = X(R, C), dense (slice.cpp:115-153)
template <typename DerivedX, typename DerivedR, typename DerivedC, typename DerivedY>
inline void slice(const Eigen::DenseBase<DerivedX>& X, const Eigen::DenseBase<DerivedR>& R,
                  const Eigen::DenseBase<DerivedC>& C, Eigen::PlainObjectBase<DerivedY>& Y) {
  Y.resize(R.size(), C.size());
  for (Eigen::Index j = 0; j < C.size(); j++)
    for (Eigen::Index i = 0; i < R.size(); i++) Y(i, j) = X(R(i), C(j));
}

// Y = X(R, :) for dim == 1, X(:, R) for dim == 2; sparse or dense (slice.cpp:79-113)
template <typename MatX, typename DerivedR, typename MatY>
inline void slice(const MatX& X, const Eigen::DenseBase<DerivedR>& R, const int dim, MatY& Y) {
  typedef Eigen::Matrix<typename DerivedR::Scalar, Eigen::Dynamic, 1> IndexVector;
  typedef typename DerivedR::Scalar I;
  if (dim == 1) {
    if (X.cols() == 0) {
      Y.resize(R.size(), 0);
      return;
    }
    const IndexVector all = IndexVector::LinSpaced(X.cols(), I(0), static_cast<I>(X.cols() - 1));
    slice(X, R, all, Y);
  } else if (dim == 2) {
    if (X.rows() == 0) {
      Y.resize(0, R.size());
      return;
    }
    const IndexVector all = IndexVector::LinSpaced(X.rows(), I(0), static_cast<I>(X.rows() - 1));
    slice(X, all, R, Y);
  }
}
}  // namespace igl
#endif
