// igl/setdiff.h -- restatement of igl::setdiff (libigl/include/igl/setdiff.cpp:19-75):
// C = sorted unique values of A that do not occur in B, IA = index into A of (the first
// occurrence of) every value of C.  TEST INFRASTRUCTURE, see slice.h.
#ifndef SMG_REF_SHIM_IGL_SETDIFF
#define SMG_REF_SHIM_IGL_SETDIFF
#include <Eigen/Core>

#include <algorithm>
#include <vector>

namespace igl {
template <typename DerivedA, typename DerivedB, typename DerivedC, typename DerivedIA>
inline void setdiff(const Eigen::MatrixBase<DerivedA>& A, const Eigen::MatrixBase<DerivedB>& B,
                    Eigen::PlainObjectBase<DerivedC>& C, Eigen::PlainObjectBase<DerivedIA>& IA) {
  if (A.size() == 0) {
    C.resize(0, 1);
    IA.resize(0, 1);
    return;
  }
  typedef typename DerivedA::Scalar SA;
  std::vector<std::pair<SA, int>> sA;
  for (int i = 0; i < static_cast<int>(A.size()); i++) sA.emplace_back(A(i), i);
  std::stable_sort(sA.begin(), sA.end(), [](const std::pair<SA, int>& x, const std::pair<SA, int>& y) { return x.first < y.first; });
  sA.erase(std::unique(sA.begin(), sA.end(), [](const std::pair<SA, int>& x, const std::pair<SA, int>& y) { return x.first == y.first; }), sA.end());
  std::vector<typename DerivedB::Scalar> sB;
  for (int i = 0; i < static_cast<int>(B.size()); i++) sB.push_back(B(i));
  std::sort(sB.begin(), sB.end());
  std::vector<SA> vC;
  std::vector<int> vIA;
  for (const auto& a : sA)
    if (!std::binary_search(sB.begin(), sB.end(), a.first)) {
      vC.push_back(a.first);
      vIA.push_back(a.second);
    }
  C.resize(static_cast<Eigen::Index>(vC.size()), 1);
  IA.resize(static_cast<Eigen::Index>(vIA.size()), 1);
  for (size_t i = 0; i < vC.size(); i++) {
    C(static_cast<Eigen::Index>(i)) = vC[i];
    IA(static_cast<Eigen::Index>(i)) = vIA[i];
  }
}
}  // namespace igl
#endif
