"""GPU checks at BASELINE.json's full size (configs[2]: 1 048 578-vertex sphere, 5 levels, FP64),
where the large-level kernel variants run (two rows per thread, next-phase L2 prefetch):
  * size-independent properties (tests/property_checks.py): linearity and symmetry of A,
    restrict = prolong^T, fused residual kernels, Gauss-Seidel fixed point and contraction, and
    a solve to 1e-10 whose result satisfies the system when recomputed on the host;
  * the order-exact (wavefront) mode against the CPU checker on the same inputs: operators to
    1e-12 (they are bit-identical on the small problems), residual history to 1e-6 with the same
    number of measurements.
"""
import numpy as np
import pytest

import property_checks as pc
from oracle import cpu_oracle
from oracle.cpu_oracle import Oracle
from surface_multigrid_code_b200 import meshgen as mg
from surface_multigrid_code_b200.solver import Solver

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big():
    return mg.sphere_problem(9, 5, tol=1e-10, max_iter=20, pad_three=True)


def test_properties_at_one_million_vertices(big):
    with Solver(device=0) as s:
        s.set_hierarchy(big.P).precompute(big.A, big.known)
        assert s.level_rows(0) == 1048572 and [s.level_stats(l)["phases"] for l in range(5)] == [4] * 5
        pc.check_operator_properties(s, big.nlev, np.random.default_rng(7), tol=1e-10)
        z, r_his = pc.check_solve_properties(s, big, 1e-10, 20)
        assert len(r_his) <= 18


def test_wavefront_mode_matches_the_cpu_checker_at_one_million_vertices(big):
    impl = "ref" if cpu_oracle.ref_available() else "port"
    ora = Oracle(big.P, impl=impl).precompute(big.A, big.known)
    rng = np.random.default_rng(8)

    def rel(a, b):
        return float(np.linalg.norm(a - b) / np.linalg.norm(b))

    with Solver(smoother="wavefront", device=0) as s:
        s.set_hierarchy(big.P).precompute(big.A, big.known)
        assert np.array_equal(s.unknown, ora.unknown)
        for lv in (0, 1):
            n = s.level_rows(lv)
            u, b = rng.standard_normal(n), rng.standard_normal(n)
            assert rel(s.apply_A(lv, u), ora.apply_A(lv, u)) < 1e-12
            assert rel(s.residual(lv, b, u), b - ora.apply_A(lv, u)) < 1e-12
            assert rel(s.restrict(lv, u), ora.restrict(lv, u)) < 1e-12
            x = rng.standard_normal(s.level_rows(lv + 1))
            assert rel(s.prolong(lv, x), ora.prolong(lv, x)) < 1e-12
            assert rel(s.relax(lv, 1, b, u), ora.relax(lv, 1, b, u.copy())) < 1e-12
            assert np.array_equal(s.diag(lv), ora.diag(lv))
        z, r_his, ok = s.solve(big.rhs, big.z0, big.known_val, 1e-10, 20)
        z_ref, r_ref, ok_ref = ora.solve(big.rhs, big.z0, big.known_val, 1e-10, 20)
        assert ok == ok_ref and len(r_his) == len(r_ref)
        assert np.allclose(r_his, r_ref, rtol=1e-6, atol=1e-14)
        assert rel(z, z_ref) < 1e-9


def test_hilbert_cube_four_million_vertices():
    """BASELINE configs[4]: hilbert_cube.obj upsampled three times (4 028 672 vertices, 346
    nearest-vertex constraints, 6 levels; irregular valence, two decimated-style levels with
    11-13 colours and wide rows): operator properties on every level and a solve to 1e-10 whose
    result satisfies the system when recomputed on the host."""
    import os

    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hilbert_cube_base.npz"))
    Pc = [mg.load_csc_keep_zeros(d, "Pc0"), mg.load_csc_keep_zeros(d, "Pc1")]
    pr = mg.upsampled_mesh_problem("hilbert_cube", d["V"], d["F"], d["known"], Pc, 3, tol=1e-10, max_iter=40)
    assert pr.n == 4028672 and pr.nlev == 6 and pr.known.size == 346
    with Solver(device=0) as s:
        s.set_hierarchy(pr.P).precompute(pr.A, pr.known)
        assert s.level_rows(0) == pr.n - 346
        pc.check_operator_properties(s, pr.nlev, np.random.default_rng(11), tol=1e-10)
        z, r_his, ok = s.solve(pr.rhs, pr.z0, pr.known_val, 1e-10, 40)
        assert ok and r_his[-1] < 1e-10 and len(r_his) <= 30
        assert np.all(np.diff(r_his[2:]) < 0)  # monotone after the first cycles
        A = pr.A.tocsr()
        unknown = np.setdiff1d(np.arange(pr.n), pr.known)
        assert np.linalg.norm((pr.rhs - A @ z)[unknown]) < 1e-9
        assert np.array_equal(z[pr.known], pr.known_val)
        assert s.solved_on_device
