// igl/slice_into.h -- restatement of the dense overloads the reference's hot path calls
// (libigl/include/igl/slice_into.cpp:51-119).  TEST INFRASTRUCTURE, see slice.h.
#ifndef SMG_REF_SHIM_IGL_SLICE_INTO
#define SMG_REF_SHIM_IGL_SLICE_INTO
#include <Eigen/Core>

namespace igl {
// Y(R(i), C(j)) = X(i, j), written in row-major order of X: when an index occurs more than
// once the last write wins (slice_into.cpp:51-82)
template <typename DerivedX, typename DerivedY, typename DerivedR, typename DerivedC>
inline void slice_into(const Eigen::MatrixBase<DerivedX>& X, const Eigen::MatrixBase<DerivedR>& R,
                       const Eigen::MatrixBase<DerivedC>& C, Eigen::PlainObjectBase<DerivedY>& Y) {
  for (Eigen::Index i = 0; i < X.rows(); i++)
    for (Eigen::Index j = 0; j < X.cols(); j++) Y(static_cast<Eigen::Index>(R(i)), static_cast<Eigen::Index>(C(j))) = X(i, j);
}

// rows (dim == 1) or columns (dim == 2) of Y (slice_into.cpp:84-117)
template <typename MatX, typename MatY, typename DerivedR>
inline void slice_into(const MatX& X, const Eigen::MatrixBase<DerivedR>& R, const int dim, MatY& Y) {
  typedef Eigen::Matrix<int, Eigen::Dynamic, 1> IndexVector;
  if (dim == 1 && X.cols() > 0) {
    const IndexVector all = IndexVector::LinSpaced(X.cols(), 0, static_cast<int>(X.cols() - 1));
    slice_into(X, R, all, Y);
  } else if (dim == 2 && X.rows() > 0) {
    const IndexVector all = IndexVector::LinSpaced(X.rows(), 0, static_cast<int>(X.rows() - 1));
    slice_into(X, all, R, Y);
  }
}
}  // namespace igl
#endif
