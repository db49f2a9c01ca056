// igl/is_symmetric.h -- only used inside assert() by the reference's hot path, which its
// Release build (and ours: -DNDEBUG) compiles out.  TEST INFRASTRUCTURE, see slice.h.
#ifndef SMG_REF_SHIM_IGL_IS_SYMMETRIC
#define SMG_REF_SHIM_IGL_IS_SYMMETRIC
#include <Eigen/Sparse>

#include <cmath>
namespace igl {
template <typename T, typename E>
inline bool is_symmetric(const Eigen::SparseMatrix<T>& A, const E epsilon) {
  if (A.rows() != A.cols()) return false;
  Eigen::SparseMatrix<T> At = A.transpose();
  if (At.nonZeros() != A.nonZeros()) return false;
  for (Eigen::Index p = 0; p < A.nonZeros(); p++)
    if (A.innerIndexPtr()[p] != At.innerIndexPtr()[p] || std::fabs(A.valuePtr()[p] - At.valuePtr()[p]) > epsilon) return false;
  return true;
}
}  // namespace igl
#endif
