"""Drop-in test at the reference's own C++ API: ONE harness (oracle/ref_harness.cpp) calls
min_quad_with_fixed_mg_precompute / _solve, mg_VCycle, relax, A, restrict, prolong, coarseSolve
through the reference's unmodified headers; linked once with the reference's own source files
(oracle/_ref/libsmg_ref.so, CPU) and once with adapter/smg_eigen_adapter.cpp + libsmg.so
(oracle/_ref/libsmg_adapter.so, GPU).  Same calls, same arguments, compare what comes back.

SMG_SMOOTHER=0 (order-exact wavefront smoother): everything except the coarse direct solve is
bit-identical to the reference's code, including the host mirrors of mg[lv].A / A_diag / P / PT
and data.LHS / Auk that the adapter copies back; solves agree to 1e-9 with equal r_his length
and return value.  Default (multicolour) smoother: same solution to 1e-7.
"""
import os

import numpy as np
import pytest

from oracle import cpu_oracle
from oracle.cpu_oracle import Oracle

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not cpu_oracle.adapter_available(),
                                 reason="oracle/_ref/libsmg_adapter.so not built (make -C oracle ref)")]


def _rand(rng, n, k):
    return rng.standard_normal(n) if k == 1 else np.asfortranarray(rng.standard_normal((n, k)))


def _same_matrix(a, b):
    return (a.shape == b.shape and np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
            and np.array_equal(a.data, b.data))


@pytest.fixture
def smoother_env():
    old = os.environ.get("SMG_SMOOTHER")
    yield lambda v: os.environ.__setitem__("SMG_SMOOTHER", str(v))
    if old is None:
        os.environ.pop("SMG_SMOOTHER", None)
    else:
        os.environ["SMG_SMOOTHER"] = old


@pytest.mark.parametrize("name", ["sphere_pad", "grid", "mcf", "block"])
def test_adapter_is_a_drop_in_for_the_reference_sources(problems, name, smoother_env):
    pr = problems[name]
    smoother_env(0)
    ref = Oracle(pr.P, impl="ref").precompute(pr.A, pr.known)
    ada = Oracle(pr.P, impl="adapter").precompute(pr.A, pr.known)
    assert np.array_equal(ada.unknown, ref.unknown)
    for lv in range(pr.nlev):
        assert _same_matrix(ada.matrix(lv, "A"), ref.matrix(lv, "A")), lv
        assert np.array_equal(ada.diag(lv), ref.diag(lv)), lv
        if lv >= 1:
            assert _same_matrix(ada.matrix(lv, "P"), ref.matrix(lv, "P")), lv
            assert _same_matrix(ada.matrix(lv, "PT"), ref.matrix(lv, "PT")), lv
    assert _same_matrix(ada.matrix(0, "LHS"), ref.matrix(0, "LHS"))
    if pr.known is not None:
        assert _same_matrix(ada.matrix(0, "Auk"), ref.matrix(0, "Auk"))
    rng = np.random.default_rng(41)
    k = pr.k
    for lv in range(pr.nlev):
        n = ref.level_rows(lv)
        u, b = _rand(rng, n, k), _rand(rng, n, k)
        assert np.array_equal(ada.relax(lv, 2, b, u.copy()), ref.relax(lv, 2, b, u.copy())), lv
        assert np.array_equal(ada.apply_A(lv, u), ref.apply_A(lv, u)), lv
        if lv + 1 < pr.nlev:
            assert np.array_equal(ada.restrict(lv, u), ref.restrict(lv, u)), lv
            x = _rand(rng, ref.level_rows(lv + 1), k)
            assert np.array_equal(ada.prolong(lv, x), ref.prolong(lv, x)), lv
            a, c = ada.vcycle(lv, b, u.copy()), ref.vcycle(lv, b, u.copy())
            assert np.linalg.norm(a - c) <= 1e-9 * np.linalg.norm(c), lv
    nc = ref.level_rows(pr.nlev - 1)
    b, u = _rand(rng, nc, k), _rand(rng, nc, k)
    a, c = ada.coarse_solve(b, u.copy()), ref.coarse_solve(b, u.copy())
    assert np.linalg.norm(a - c) <= 1e-9 * np.linalg.norm(c)
    for tol, max_iter in ((1e-3, 20), (1e-10, 30), (1e-30, 3)):
        z1, r1, ok1 = ada.solve(pr.rhs, pr.z0, pr.known_val, tol, max_iter)
        z2, r2, ok2 = ref.solve(pr.rhs, pr.z0, pr.known_val, tol, max_iter)
        assert ok1 == ok2 and len(r1) == len(r2)
        assert np.allclose(r1, r2, rtol=1e-6, atol=1e-14)
        assert np.linalg.norm(z1 - z2) <= 1e-9 * np.linalg.norm(z2)


@pytest.mark.parametrize("name", ["sphere_pad", "mcf", "block"])
def test_adapter_default_smoother_reaches_the_reference_solution(problems, name, smoother_env):
    pr = problems[name]
    smoother_env(1)
    ref = Oracle(pr.P, impl="ref").precompute(pr.A, pr.known)
    ada = Oracle(pr.P, impl="adapter").precompute(pr.A, pr.known)
    z1, r1, ok1 = ada.solve(pr.rhs, pr.z0, pr.known_val, 1e-10, 40)
    z2, r2, ok2 = ref.solve(pr.rhs, pr.z0, pr.known_val, 1e-10, 40)
    assert ok1 and ok2 and abs(len(r1) - len(r2)) <= 2
    assert np.linalg.norm(z1 - z2) <= 1e-7 * np.linalg.norm(z2)


def test_repeated_precompute_with_new_values_is_a_numeric_refresh(problems, smoother_env):
    """05_example_mean_curvature_flow/main.cpp:74 calls min_quad_with_fixed_mg_precompute every
    time step with the same pattern and new values: the drop-in reuses its plan
    (smg_update_values) and still returns what the reference's code returns."""
    import scipy.sparse as sp

    pr = problems["mcf"]
    smoother_env(0)
    ref = Oracle(pr.P, impl="ref").precompute(pr.A, pr.known)
    ada = Oracle(pr.P, impl="adapter").precompute(pr.A, pr.known)
    before = ada.refresh_count()
    A2 = (pr.A + 0.37 * sp.identity(pr.A.shape[0], format="csc")).tocsc()
    A2.sort_indices()
    assert np.array_equal(A2.indptr, pr.A.tocsc().indptr)  # same pattern, new values
    ref.precompute(A2, pr.known)
    ada.precompute(A2, pr.known)
    assert ada.refresh_count() == before + 1
    for lv in range(pr.nlev):
        assert _same_matrix(ada.matrix(lv, "A"), ref.matrix(lv, "A")), lv
        assert np.array_equal(ada.diag(lv), ref.diag(lv)), lv
    z1, r1, ok1 = ada.solve(pr.rhs, pr.z0, pr.known_val, 1e-10, 30)
    z2, r2, ok2 = ref.solve(pr.rhs, pr.z0, pr.known_val, 1e-10, 30)
    assert ok1 == ok2 and len(r1) == len(r2)
    assert np.linalg.norm(z1 - z2) <= 1e-9 * np.linalg.norm(z2)
    # SMG_NO_REFRESH=1 forces the full precompute; same results
    os.environ["SMG_NO_REFRESH"] = "1"
    try:
        ada.precompute(pr.A, pr.known)
    finally:
        os.environ.pop("SMG_NO_REFRESH", None)
    assert ada.refresh_count() == before + 1
    ref.precompute(pr.A, pr.known)
    for lv in range(pr.nlev):
        assert _same_matrix(ada.matrix(lv, "A"), ref.matrix(lv, "A")), lv


def test_r_his_is_cleared_per_solve_and_handles_are_released(problems, smoother_env):
    """min_quad_with_fixed_mg.cpp:105/:327 clear r_his at the start of every solve: a caller that
    reuses one vector across time steps sees the history of the LAST solve only.  Also: destroying
    the caller's solver objects (smg_adapter_release) frees the device handle."""
    pr = problems["sphere_pad"]
    smoother_env(0)
    ref = Oracle(pr.P, impl="ref").precompute(pr.A, pr.known)
    ada = Oracle(pr.P, impl="adapter").precompute(pr.A, pr.known)
    assert ada.solve_twice_same_rhis(pr.rhs, pr.z0, pr.known_val, 1e-8, 30) == \
        ref.solve_twice_same_rhis(pr.rhs, pr.z0, pr.known_val, 1e-8, 30)
    live = ada.live_handles()
    assert live >= 1
    ada.__del__()
    assert Oracle(pr.P, impl="adapter").live_handles() == live - 1
