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
// Y = X(R, C), sparse (slice.cpp:13-77): every stored entry of X, in storage order, goes to
// every (i, j) with R(i) == row and C(j) == col; setFromTriplets sorts and sums duplicates.
template <typename TX, typename TY, typename DerivedR, typename DerivedC>
inline void slice(const Eigen::SparseMatrix<TX>& X, const Eigen::DenseBase<DerivedR>& R,
                  const Eigen::DenseBase<DerivedC>& C, Eigen::SparseMatrix<TY>& Y) {
  const int xm = static_cast<int>(X.rows()), xn = static_cast<int>(X.cols());
  const int ym = static_cast<int>(R.size()), yn = static_cast<int>(C.size());
  if (ym == 0 || yn == 0) {
    Y.resize(ym, yn);
    return;
  }
  std::vector<std::vector<int>> RI(static_cast<size_t>(xm)), CI(static_cast<size_t>(xn));
  for (int i = 0; i < ym; i++) RI[static_cast<size_t>(R(i))].push_back(i);
  for (int i = 0; i < yn; i++) CI[static_cast<size_t>(C(i))].push_back(i);
  std::vector<Eigen::Triplet<TY>> entries;
  for (int k = 0; k < X.outerSize(); ++k)
    for (typename Eigen::SparseMatrix<TX>::InnerIterator it(X, k); it; ++it)
      for (int r : RI[static_cast<size_t>(it.row())])
        for (int c : CI[static_cast<size_t>(it.col())]) entries.emplace_back(r, c, it.value());
  Y.resize(ym, yn);
  Y.setFromTriplets(entries.begin(), entries.end());
}

// Y = X(R, C), dense (slice.cpp:115-153)
template <typename DerivedX, typename DerivedR, typename DerivedC, typename DerivedY>
inline void slice(const Eigen::DenseBase<DerivedX>& X, const Eigen::DenseBase<DerivedR>& R,
                  const Eigen::DenseBase<DerivedC>& C, Eigen::PlainObjectBase<DerivedY>& Y) {
  const int ym = static_cast<int>(R.size()), yn = static_cast<int>(C.size());
  if (ym == 0 || yn == 0) {
    Y.resize(ym, yn);
    return;
  }
  Y.resize(ym, yn);
  for (int i = 0; i < ym; i++)
    for (int j = 0; j < yn; j++) Y(i, j) = X(R(i), C(j));
}

// Y = X(R, :) (dim 1) or X(:, R) (dim 2), sparse or dense (slice.cpp:79-113)
template <typename MatX, typename DerivedR, typename MatY>
inline void slice(const MatX& X, const Eigen::DenseBase<DerivedR>& R, const int dim, MatY& Y) {
  Eigen::Matrix<typename DerivedR::Scalar, Eigen::Dynamic, 1> C;
  switch (dim) {
    case 1:
      if (X.cols() == 0) {
        Y.resize(R.size(), 0);
        return;
      }
      C = Eigen::Matrix<typename DerivedR::Scalar, Eigen::Dynamic, 1>::LinSpaced(X.cols(), 0, static_cast<typename DerivedR::Scalar>(X.cols() - 1));
      return slice(X, R, C, Y);
    case 2:
      if (X.rows() == 0) {
        Y.resize(0, R.size());
        return;
      }
      C = Eigen::Matrix<typename DerivedR::Scalar, Eigen::Dynamic, 1>::LinSpaced(X.rows(), 0, static_cast<typename DerivedR::Scalar>(X.rows() - 1));
      return slice(X, C, R, Y);
    default:
      return;
  }
}
}  // namespace igl
#endif
