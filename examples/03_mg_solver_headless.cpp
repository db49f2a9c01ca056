// Headless restatement of the solver calls of the reference's 03_mg_solver/main.cpp
// (lines 35-39 and 64-75; 04_mg_solver_nobd/main.cpp:100-105 is the same with a tolerance),
// written against the reference's OWN headers and linked with adapter/smg_eigen_adapter.cpp
// + libsmg.so instead of src/min_quad_with_fixed_mg.cpp + src/mg_VCycle.cpp.
//
// The viewer, mesh IO and mg_precompute (CPU hierarchy build, out of scope) are replaced by
// a flat problem file written by surface_multigrid_code_b200/meshgen.py::write_problem_file:
// it holds what main.cpp has in hand before the two solver calls (A, mg[lv].P_full, b, B,
// bval, z0).  Build (the image has no Eigen, so tests use oracle/ref_shim; with real
// Eigen 3.3.7 drop the first -I):
//   g++ -std=c++17 -I oracle/ref_shim -I /root/reference/src -I include
//       examples/03_mg_solver_headless.cpp adapter/smg_eigen_adapter.cpp
//       surface_multigrid_code_b200/libsmg.so -Wl,-rpath,$PWD/surface_multigrid_code_b200 -o 03_headless
//   ./03_headless problem.bin [tolerance [maxIter]]
#include <min_quad_with_fixed_mg.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

namespace {
template <class T>
bool read_vec(FILE* f, std::vector<T>& v, size_t n) {
  v.resize(n);
  return n == 0 || std::fread(v.data(), sizeof(T), n, f) == n;
}

bool read_sparse(FILE* f, Eigen::SparseMatrix<double>& M) {
  int hdr[3];
  if (std::fread(hdr, sizeof(int), 3, f) != 3) return false;
  M = Eigen::SparseMatrix<double>(hdr[0], hdr[1]);
  M.resizeNonZeros(hdr[2]);
  return std::fread(M.outerIndexPtr(), sizeof(int), hdr[1] + 1, f) == static_cast<size_t>(hdr[1] + 1) &&
         std::fread(M.innerIndexPtr(), sizeof(int), hdr[2], f) == static_cast<size_t>(hdr[2]) &&
         std::fread(M.valuePtr(), sizeof(double), hdr[2], f) == static_cast<size_t>(hdr[2]);
}
}  // namespace

int main(int argc, char* argv[]) {
  using namespace Eigen;
  using namespace std;
  if (argc < 2) {
    fprintf(stderr, "usage: %s problem.bin [tolerance [maxIter]]\n", argv[0]);
    return 2;
  }
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  int hdr[4];  // magic, number of levels, n, number of known
  if (fread(hdr, sizeof(int), 4, f) != 4 || hdr[0] != 0x534d4731) return 2;
  const int nLvs = hdr[1], n = hdr[2], nb = hdr[3];

  // what mg_precompute leaves behind (src/mg_precompute.cpp:71-77): mg[lv].P_full
  vector<mg_data> mg(nLvs);
  SparseMatrix<double> A;
  if (!read_sparse(f, A)) return 2;
  for (int lv = 1; lv < nLvs; lv++)
    if (!read_sparse(f, mg[lv].P_full)) return 2;
  VectorXi b(nb);
  VectorXd bval(nb), B(n), z0(n);
  vector<int> ib;
  vector<double> tmp;
  if (!read_vec(f, ib, nb)) return 2;
  for (int i = 0; i < nb; i++) b(i) = ib[i];
  if (!read_vec(f, tmp, nb)) return 2;
  for (int i = 0; i < nb; i++) bval(i) = tmp[i];
  if (!read_vec(f, tmp, n)) return 2;
  for (int i = 0; i < n; i++) B(i) = tmp[i];
  if (!read_vec(f, tmp, n)) return 2;
  for (int i = 0; i < n; i++) z0(i) = tmp[i];
  fclose(f);

  // ---- 03_mg_solver/main.cpp:66-75, verbatim in spirit ----
  vector<double> rHis;
  VectorXd z;
  min_quad_with_fixed_mg_data data;
  SimplicialLDLT<SparseMatrix<double>> solver;
  min_quad_with_fixed_mg_precompute(A, b, data, mg, solver);
  bool ok;
  if (argc >= 4)
    ok = min_quad_with_fixed_mg_solve(data, B, bval, z0, solver, atof(argv[2]), atoi(argv[3]), mg, z, rHis);
  else if (argc >= 3)
    ok = min_quad_with_fixed_mg_solve(data, B, bval, z0, solver, atof(argv[2]), mg, z, rHis);
  else
    ok = min_quad_with_fixed_mg_solve(data, B, bval, z0, solver, mg, z, rHis);

  double checksum = 0.0;
  for (int i = 0; i < n; i++) checksum += z(i) * ((i % 7) + 1);
  printf("converged %d iterations %d checksum %.17g\n", ok ? 1 : 0, (int)rHis.size(), checksum);
  for (size_t i = 0; i < rHis.size(); i++) printf("r_his %zu %.17g\n", i, rHis[i]);
  return 0;
}
