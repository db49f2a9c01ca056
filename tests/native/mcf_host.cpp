// Host harness of surface_multigrid_code_b200/csrc/mcf_core.hpp for the CPU tests
// (tests/test_mcf_core.py): runs the SAME per-face / per-vertex arithmetic the CUDA kernels
// run, on the host, so it can be compared with a numpy restatement of libigl without a GPU.
// TEST INFRASTRUCTURE; never linked into libsmg.so.
#include "../../surface_multigrid_code_b200/csrc/mcf_core.hpp"

extern "C" {
// F: nF x 3 column-major; vf_ptr / vf_face: incident faces per vertex in assembly order
void mcf_host_assemble(int nV, int nF, const int* F, const double* U, const int* vf_ptr, const int* vf_face,
                       double* dblA, double* mass) {
  for (int f = 0; f < nF; f++) dblA[f] = smg::mcf_face_doublearea(U, nV, F[f], F[f + nF], F[f + 2 * nF]);
  for (int v = 0; v < nV; v++) mass[v] = smg::mcf_vertex_mass(dblA, vf_face, vf_ptr[v], vf_ptr[v + 1]);
}
double mcf_host_lhs_entry(double m, double delta, double l) { return smg::mcf_lhs_entry(m, delta, l); }
}
