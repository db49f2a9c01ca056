"""Pins the C restatement (oracle/smg_oracle.c) against THE REFERENCE'S OWN SOURCES:
oracle/_ref/libsmg_ref.so is /root/reference/src/mg_VCycle.cpp + min_quad_with_fixed_mg.cpp
compiled unmodified (make -C oracle ref) against the Eigen / igl stand-in of oracle/ref_shim
(genuine Eigen 3.3.7 is not vendored by the reference and not in this image).

Bars: index / topology outputs, Galerkin values, A_diag, relax, A, restrict, prolong: bit-exact.
Coarse solve / V-cycle / solve: rel 1e-9 (the stand-in factorises with RCM + an envelope
Cholesky, the restatement with RCM + banded Cholesky, the real reference with AMD +
SimplicialLDLT: three elimination orders); residual histories rel 1e-9; the number of residual
measurements and the return value: equal; residual histories rel 1e-9 with an absolute floor
of 1e-15.
"""
import numpy as np
import pytest

import golden_util
from oracle import cpu_oracle
from oracle.cpu_oracle import Oracle
from surface_multigrid_code_b200.solver import Solver

pytestmark = pytest.mark.skipif(not cpu_oracle.ref_available(),
                                reason="oracle/_ref/libsmg_ref.so not built (needs /root/reference)")

NAMES = ["sphere_pad", "sphere", "grid", "mcf", "block"]


def _rand(rng, n, k):
    return rng.standard_normal(n) if k == 1 else np.asfortranarray(rng.standard_normal((n, k)))


def _same_matrix(a, b):
    return (a.shape == b.shape and np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
            and np.array_equal(a.data, b.data))


@pytest.mark.parametrize("name", NAMES)
def test_precompute_outputs_bit_exact(problems, name):
    pr = problems[name]
    port = Oracle(pr.P).precompute(pr.A, pr.known)
    ref = Oracle(pr.P, impl="ref").precompute(pr.A, pr.known)
    assert np.array_equal(port.unknown, ref.unknown)
    for lv in range(pr.nlev):
        assert _same_matrix(port.matrix(lv, "A"), ref.matrix(lv, "A")), lv
        assert np.array_equal(port.diag(lv), ref.diag(lv)), lv
        if lv >= 1:
            assert _same_matrix(port.matrix(lv, "P"), ref.matrix(lv, "P")), lv
            assert _same_matrix(port.matrix(lv, "PT"), ref.matrix(lv, "PT")), lv
    assert _same_matrix(port.matrix(0, "LHS"), ref.matrix(0, "LHS"))
    if pr.known is not None:
        assert _same_matrix(port.matrix(0, "Auk"), ref.matrix(0, "Auk"))
    # the library's host planning reproduces the same index outputs (no CUDA needed)
    s = Solver(device="none").set_hierarchy(pr.P).precompute(pr.A, pr.known)
    assert np.array_equal(s.unknown, ref.unknown)
    for lv in range(pr.nlev):
        a, b = s.matrix(lv, "A", values=False), ref.matrix(lv, "A")
        assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)


@pytest.mark.parametrize("name", NAMES)
def test_operators_bit_exact(problems, name):
    pr = problems[name]
    port = Oracle(pr.P).precompute(pr.A, pr.known)
    ref = Oracle(pr.P, impl="ref").precompute(pr.A, pr.known)
    rng = np.random.default_rng(21)
    k = pr.k
    for lv in range(pr.nlev):
        n = ref.level_rows(lv)
        assert n == port.level_rows(lv)
        u, b = _rand(rng, n, k), _rand(rng, n, k)
        for iters in (1, 2):
            assert np.array_equal(port.relax(lv, iters, b, u.copy()), ref.relax(lv, iters, b, u.copy())), lv
        assert np.array_equal(port.apply_A(lv, u), ref.apply_A(lv, u)), lv
        if lv + 1 < pr.nlev:
            assert np.array_equal(port.restrict(lv, u), ref.restrict(lv, u)), lv
            x = _rand(rng, ref.level_rows(lv + 1), k)
            assert np.array_equal(port.prolong(lv, x), ref.prolong(lv, x)), lv
    nc = ref.level_rows(pr.nlev - 1)
    b, u = _rand(rng, nc, k), _rand(rng, nc, k)
    a, c = port.coarse_solve(b, u.copy()), ref.coarse_solve(b, u.copy())
    assert np.linalg.norm(a - c) <= 1e-9 * np.linalg.norm(c)


@pytest.mark.parametrize("name", NAMES)
def test_vcycle_and_solve_match(problems, name):
    pr = problems[name]
    port = Oracle(pr.P).precompute(pr.A, pr.known)
    ref = Oracle(pr.P, impl="ref").precompute(pr.A, pr.known)
    rng = np.random.default_rng(22)
    n0 = ref.level_rows(0)
    u, b = _rand(rng, n0, pr.k), _rand(rng, n0, pr.k)
    for lv0 in (0, 1):
        n = ref.level_rows(lv0)
        uu, bb = _rand(rng, n, pr.k), _rand(rng, n, pr.k)
        a, c = port.vcycle(lv0, bb, uu.copy()), ref.vcycle(lv0, bb, uu.copy())
        assert np.linalg.norm(a - c) <= 1e-9 * np.linalg.norm(c)
    for tol, max_iter in ((1e-3, 20), (1e-10, 30), (1e-30, 3)):  # default, converged, not converged
        z1, r1, ok1 = port.solve(pr.rhs, pr.z0, pr.known_val, tol, max_iter)
        z2, r2, ok2 = ref.solve(pr.rhs, pr.z0, pr.known_val, tol, max_iter)
        assert ok1 == ok2 and len(r1) == len(r2)
        # (rounding of the two coarse factorisations, ~1e-16 of the solution, is not small
        # against a residual of 1e-11: absolute floor)
        assert np.allclose(r1, r2, rtol=1e-9, atol=1e-15)
        assert np.linalg.norm(z1 - z2) <= 1e-9 * np.linalg.norm(z2)
    # quirk A.2.1: when the loop runs out, r_his has maxIter entries and the return value uses
    # the residual measured BEFORE the last cycle
    z2, r2, ok2 = ref.solve(pr.rhs, pr.z0, pr.known_val, 1e-30, 3)
    assert len(r2) == 3 and not ok2


@pytest.mark.parametrize("name", golden_util.NAMES)
def test_golden_fixtures_match_reference_sources(name):
    """the committed fixtures (generated by the scipy restatement, tests/golden/make_golden.py)
    are reproduced by the reference sources"""
    g = golden_util.load(name)
    ref = Oracle(g["P"], impl="ref").precompute(g["A"], g["known"])
    assert np.array_equal(ref.unknown, g["unknown"])
    z, r_his, ok = ref.solve(g["rhs"], g["z0"], g["known_val"], float(g["tol"]), int(g["max_iter"]))
    assert len(r_his) == len(g["r_his"]) and bool(ok) == bool(g["converged"])
    assert np.allclose(r_his, g["r_his"], rtol=1e-6, atol=1e-16)
    assert np.linalg.norm(z - g["z"]) <= 1e-8 * np.linalg.norm(g["z"])
