"""GPU tests of the communication-avoiding patch smoother (csrc/patch.hpp, patch_kernel):
a level smoothed in patches gives BIT-IDENTICAL results to the same level smoothed with one
kernel per colour phase (patch_rows = -1), because every row update does the same arithmetic
on the same inputs; and both agree with the CPU checker's colour-major sweep."""
import numpy as np
import pytest

from oracle.cpu_oracle import Oracle
from surface_multigrid_code_b200.solver import Solver

pytestmark = pytest.mark.gpu


def _pair(pr, patch_rows, **kw):
    a = Solver(smoother="multicolour", device=0, patch_rows=patch_rows, **kw).set_hierarchy(pr.P).precompute(pr.A, pr.known)
    b = Solver(smoother="multicolour", device=0, patch_rows=-1, **kw).set_hierarchy(pr.P).precompute(pr.A, pr.known)
    return a, b


@pytest.mark.parametrize("name", ["sphere_pad", "sphere", "grid", "mcf", "block"])
@pytest.mark.parametrize("patch_rows", [0, 40, 300])
@pytest.mark.parametrize("graph", [True, False])
def test_patched_vcycle_is_bit_identical_to_the_phase_kernels(problems, name, patch_rows, graph):
    pr = problems[name]
    a, b = _pair(pr, patch_rows, use_graph=graph)
    assert any(a.level_patched(lv) > 0 for lv in range(pr.nlev - 1))
    assert all(b.level_patched(lv) == 0 for lv in range(pr.nlev))
    rng = np.random.default_rng(5)
    for lv in range(pr.nlev - 1):
        n = a.level_rows(lv)
        for k in (1, 3, 5):
            B = rng.standard_normal(n) if k == 1 else np.asfortranarray(rng.standard_normal((n, k)))
            u = rng.standard_normal(n) if k == 1 else np.asfortranarray(rng.standard_normal((n, k)))
            ua, ub = a.vcycle(lv, B, u.copy()), b.vcycle(lv, B, u.copy())
            assert np.array_equal(ua, ub), (lv, k)
            # twice in a row: the second buffer of u is handed back correctly
            assert np.array_equal(a.vcycle(lv, B, ua.copy()), b.vcycle(lv, B, ub.copy())), (lv, k)
    a.close()
    b.close()


@pytest.mark.parametrize("name", ["sphere_pad", "grid", "mcf"])
def test_patched_solve_equals_unpatched_solve_and_reaches_the_reference_solution(problems, name):
    pr = problems[name]
    a, b = _pair(pr, 0)
    za, ra, oka = a.solve(pr.rhs, pr.z0, pr.known_val, 1e-10, 40)
    zb, rb, okb = b.solve(pr.rhs, pr.z0, pr.known_val, 1e-10, 40)
    assert oka and okb and np.array_equal(ra, rb) and np.array_equal(za, zb)
    ora = Oracle(pr.P).precompute(pr.A, pr.known)
    z_ref, r_ref, ok_ref = ora.solve(pr.rhs, pr.z0, pr.known_val, 1e-10, 40)
    assert ok_ref and abs(len(ra) - len(r_ref)) <= 2
    assert np.linalg.norm(za - z_ref) <= 1e-7 * np.linalg.norm(z_ref)
    a.close()
    b.close()


def test_other_sweep_counts_fall_back_to_the_phase_kernels(problems):
    """The patch layout is built for the handle's pre / post sweep counts; a V-cycle asked for
    with other counts runs one kernel per phase and still gives the same fixed point."""
    pr = problems["sphere_pad"]
    a, b = _pair(pr, 0)
    rng = np.random.default_rng(8)
    n = a.level_rows(0)
    B, u = rng.standard_normal(n), rng.standard_normal(n)
    for pre, post in ((1, 1), (3, 0), (0, 2)):
        assert np.array_equal(a.vcycle(0, B, u.copy(), pre, post), b.vcycle(0, B, u.copy(), pre, post))
    a.close()
    b.close()
    for pre, post in ((1, 1), (3, 0), (0, 2), (0, 0)):
        a, b = _pair(pr, 0, pre_relax=pre, post_relax=post)
        assert np.array_equal(a.vcycle(0, B, u.copy(), pre, post), b.vcycle(0, B, u.copy(), pre, post)), (pre, post)
        a.close()
        b.close()


def test_patch_values_follow_a_numeric_refresh(problems):
    pr = problems["mcf"]
    a, b = _pair(pr, 0)
    A2 = pr.A.copy()
    A2.data = A2.data * 1.25
    A2 = ((A2 + A2.T) * 0.5).tocsc()
    A2.sort_indices()
    a.update_values(A2.data)
    b.update_values(A2.data)
    za, ra, _ = a.solve(pr.rhs, pr.z0, None, pr.tol, pr.max_iter)
    zb, rb, _ = b.solve(pr.rhs, pr.z0, None, pr.tol, pr.max_iter)
    assert np.array_equal(ra, rb) and np.array_equal(za, zb)
    a.close()
    b.close()
