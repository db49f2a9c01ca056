// igl/slice_into.h -- restatement of the dense overloads the reference's hot path calls
// (libigl/include/igl/slice_into.cpp:51-119).  TEST INFRASTRUCTURE, see slice.h.
#ifndef SMG_REF_SHIM_IGL_SLICE_INTO
#define SMG_REF_SHIM_IGL_SLICE_INTO
#include <Eigen/Core>

namespace igl {
// Y(R(i), C(j)) = X(i, j): sequential writes, a repeated index is overwritten by its last
// occurrence (slice_into.cpp:51-82)
template <typename DerivedX, typename DerivedY, typename DerivedR, typename DerivedC>
inline void slice_into(const Eigen::MatrixBase<DerivedX>& X, const Eigen::MatrixBase<DerivedR>& R,
                       const Eigen::MatrixBase<DerivedC>& C, Eigen::PlainObjectBase<DerivedY>& Y) {
  const int xm = static_cast<int>(X.rows()), xn = static_cast<int>(X.cols());
  for (int i = 0; i < xm; i++)
    for (int j = 0; j < xn; j++) Y(int(R(i)), int(C(j))) = X(i, j);
}

// rows (dim 1) or columns (dim 2) of Y (slice_into.cpp:84-117)
template <typename MatX, typename MatY, typename DerivedR>
inline void slice_into(const MatX& X, const Eigen::MatrixBase<DerivedR>& R, const int dim, MatY& Y) {
  Eigen::Matrix<int, Eigen::Dynamic, 1> C;
  switch (dim) {
    case 1:
      if (X.cols() == 0) return;
      C = Eigen::Matrix<int, Eigen::Dynamic, 1>::LinSpaced(X.cols(), 0, static_cast<int>(X.cols() - 1));
      return slice_into(X, R, C, Y);
    case 2:
      if (X.rows() == 0) return;
      C = Eigen::Matrix<int, Eigen::Dynamic, 1>::LinSpaced(X.rows(), 0, static_cast<int>(X.rows() - 1));
      return slice_into(X, C, R, Y);
    default:
      return;
  }
}
}  // namespace igl
#endif
