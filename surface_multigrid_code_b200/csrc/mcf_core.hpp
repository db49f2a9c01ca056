// mcf_core.hpp -- per-face / per-vertex arithmetic of the mean-curvature-flow assembly
// (05_example_mean_curvature_flow/main.cpp:66-69: M = massmatrix(U, F, BARYCENTRIC),
// LHS = M - delta * L, RHS = M * U), shared by the CUDA kernels (kernels.cu) and by the host
// harness of the CPU tests (tests/native/mcf_host.cpp), so the arithmetic itself is checked
// without a GPU.  Restated from libigl (vendored by the reference):
//   squared_edge_lengths.cpp:39-41  l0 = |V[F1]-V[F2]|, l1 = |V[F2]-V[F0]|, l2 = |V[F0]-V[F1]|
//   doublearea.cpp (from lengths)   lengths sorted descending, Kahan's form of Heron's formula,
//                                   dblA = 2 * 0.25 * sqrt(arg), NaN replaced by 0
//   massmatrix_intrinsic.cpp:56-66  every corner of a face gets dblA / 6
// No FMA contraction (the reference build has none): products and sums rounded separately.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define SMG_HD __device__ __forceinline__
#define SMG_MUL(a, b) __dmul_rn((a), (b))
#define SMG_ADD(a, b) __dadd_rn((a), (b))
#define SMG_SUB(a, b) __dsub_rn((a), (b))
#else
#define SMG_HD inline
// the host harness is compiled with -ffp-contract=off
#define SMG_MUL(a, b) ((a) * (b))
#define SMG_ADD(a, b) ((a) + (b))
#define SMG_SUB(a, b) ((a) - (b))
#endif

namespace smg {

// |p - q| for rows p, q of the column-major nV x 3 array U
SMG_HD double mcf_edge_length(const double* U, int nV, int p, int q) {
  const double dx = SMG_SUB(U[p], U[q]);
  const double dy = SMG_SUB(U[p + nV], U[q + nV]);
  const double dz = SMG_SUB(U[p + 2 * nV], U[q + 2 * nV]);
  return sqrt(SMG_ADD(SMG_ADD(SMG_MUL(dx, dx), SMG_MUL(dy, dy)), SMG_MUL(dz, dz)));
}

// twice the area of the triangle (a, b, c)
SMG_HD double mcf_face_doublearea(const double* U, int nV, int a, int b, int c) {
  double l0 = mcf_edge_length(U, nV, b, c);
  double l1 = mcf_edge_length(U, nV, c, a);
  double l2 = mcf_edge_length(U, nV, a, b);
  double t;  // sort descending
  if (l0 < l1) { t = l0; l0 = l1; l1 = t; }
  if (l1 < l2) { t = l1; l1 = l2; l2 = t; }
  if (l0 < l1) { t = l0; l0 = l1; l1 = t; }
  const double f0 = SMG_ADD(l0, SMG_ADD(l1, l2));
  const double f1 = SMG_SUB(l2, SMG_SUB(l0, l1));
  const double f2 = SMG_ADD(l2, SMG_SUB(l0, l1));
  const double f3 = SMG_ADD(l0, SMG_SUB(l1, l2));
  const double arg = SMG_MUL(SMG_MUL(SMG_MUL(f0, f1), f2), f3);
  const double d = SMG_MUL(SMG_MUL(2.0, 0.25), sqrt(arg));
  return d == d ? d : 0.0;
}

// barycentric mass of a vertex: the corner contributions dblA / 6 summed in the order
// setFromTriplets sees them (all faces with the vertex at corner 0, then corner 1, then 2;
// faces ascending) -- `faces` lists the incident faces of the vertex in exactly that order
SMG_HD double mcf_vertex_mass(const double* dblA, const int* faces, int begin, int end) {
  double m = 0.0;
  for (int t = begin; t < end; t++) m = SMG_ADD(m, dblA[faces[t]] / 6.0);
  return m;
}

// one entry of LHS = M - delta * L (M diagonal): Eigen evaluates m - (delta * l)
SMG_HD double mcf_lhs_entry(double mass_or_zero, double delta, double l) {
  return SMG_SUB(mass_or_zero, SMG_MUL(delta, l));
}

}  // namespace smg
