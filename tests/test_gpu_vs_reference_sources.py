"""GPU parity against THE REFERENCE'S OWN SOURCES (oracle/_ref/libsmg_ref.so: the reference's
mg_VCycle.cpp + min_quad_with_fixed_mg.cpp compiled unmodified on the Eigen stand-in of
oracle/ref_shim; built where /root/reference is mounted, shipped to the GPU box as a file).

Wavefront (order-exact) smoother: relax, A, residual, restrict, prolong and the Galerkin
operators computed on the GPU are BIT-IDENTICAL to what the reference's code computes; the
solve agrees to 1e-9 (coarse factorisations differ), with the same number of residual
measurements and the same return value.  Multicolour smoother: same solution to 1e-7.
"""
import numpy as np
import pytest

from oracle import cpu_oracle
from oracle.cpu_oracle import Oracle
from surface_multigrid_code_b200.solver import Solver

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not cpu_oracle.ref_available(), reason="oracle/_ref/libsmg_ref.so not built")]


def _rand(rng, n, k):
    return rng.standard_normal(n) if k == 1 else np.asfortranarray(rng.standard_normal((n, k)))


@pytest.mark.parametrize("name", ["sphere_pad", "grid", "mcf"])
def test_wavefront_mode_is_bit_identical_to_the_reference_code(problems, name):
    pr = problems[name]
    ref = Oracle(pr.P, impl="ref").precompute(pr.A, pr.known)
    rng = np.random.default_rng(31)
    k = pr.k
    with Solver(smoother="wavefront", device=0) as s:
        s.set_hierarchy(pr.P).precompute(pr.A, pr.known)
        assert np.array_equal(s.unknown, ref.unknown)
        for lv in range(pr.nlev):
            a, b = s.matrix(lv, "A"), ref.matrix(lv, "A")
            assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
            assert np.array_equal(a.data, b.data)
            assert np.array_equal(s.diag(lv), ref.diag(lv))
            n = s.level_rows(lv)
            u, rhs = _rand(rng, n, k), _rand(rng, n, k)
            assert np.array_equal(s.relax(lv, 2, rhs, u), ref.relax(lv, 2, rhs, u.copy()))
            assert np.array_equal(s.apply_A(lv, u), ref.apply_A(lv, u))
            assert np.array_equal(s.residual(lv, rhs, u), rhs - ref.apply_A(lv, u))
            if lv + 1 < pr.nlev:
                assert np.array_equal(s.restrict(lv, u), ref.restrict(lv, u))
                x = _rand(rng, s.level_rows(lv + 1), k)
                assert np.array_equal(s.prolong(lv, x), ref.prolong(lv, x))
        z, r_his, ok = s.solve(pr.rhs, pr.z0, pr.known_val, 1e-10, 30)
        z_ref, r_ref, ok_ref = ref.solve(pr.rhs, pr.z0, pr.known_val, 1e-10, 30)
        assert ok == ok_ref and len(r_his) == len(r_ref)
        assert np.allclose(r_his, r_ref, rtol=1e-6, atol=1e-14)
        assert np.linalg.norm(z - z_ref) <= 1e-9 * np.linalg.norm(z_ref)


@pytest.mark.parametrize("name", ["sphere_pad", "grid", "mcf"])
def test_multicolour_mode_reaches_the_reference_solution(problems, name):
    pr = problems[name]
    ref = Oracle(pr.P, impl="ref").precompute(pr.A, pr.known)
    z_ref, r_ref, ok_ref = ref.solve(pr.rhs, pr.z0, pr.known_val, 1e-10, 40)
    with Solver(device=0) as s:
        s.set_hierarchy(pr.P).precompute(pr.A, pr.known)
        z, r_his, ok = s.solve(pr.rhs, pr.z0, pr.known_val, 1e-10, 40)
    assert ok and ok_ref and abs(len(r_his) - len(r_ref)) <= 2
    assert np.linalg.norm(z - z_ref) <= 1e-7 * np.linalg.norm(z_ref)
