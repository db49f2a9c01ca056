"""Edge cases of min_quad_with_fixed_mg_precompute / _solve, checked three ways on the CPU: the
reference's own sources (oracle/_ref), the C restatement, and the library's host planning
(plan-only handle: index / topology outputs, no CUDA).  Covers what the example meshes do not
reach: column pruning that cascades over two levels (min_quad_with_fixed_mg.cpp:186-220), the
signed pruning threshold (:197), duplicate / unsorted / empty `known`, immediate convergence,
maxIter = 1, and several right-hand-side columns with fixed values.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import cpu_oracle
from oracle.cpu_oracle import Oracle
from surface_multigrid_code_b200.solver import Solver

pytestmark = pytest.mark.skipif(not cpu_oracle.ref_available(),
                                reason="oracle/_ref/libsmg_ref.so not built (needs /root/reference)")


def chain_hierarchy(n0=33, levels=3, negative_column=None):
    """1-D chain: A = tridiag(-1, 2.2, -1); P = linear interpolation onto every other node,
    stored with explicit zeros like the reference's 3-entries-per-row layout."""
    A = sp.diags([-np.ones(n0 - 1), 2.2 * np.ones(n0), -np.ones(n0 - 1)], [-1, 0, 1], format="csc")
    P = []
    nf = n0
    for _ in range(levels - 1):
        nc = (nf + 1) // 2
        rows, cols, vals = [], [], []
        for i in range(nf):
            if i % 2 == 0:
                ent = [(i // 2, 1.0), (min(i // 2 + 1, nc - 1), 0.0)]
            else:
                ent = [((i - 1) // 2, 0.5), ((i + 1) // 2, 0.5)]
            for c, v in ent:
                if (i, c) not in zip(rows, cols):
                    rows.append(i), cols.append(c), vals.append(v)
        M = sp.csc_matrix((nf, nc))
        coo = sp.coo_matrix((np.ones(len(vals)), (rows, cols)), shape=(nf, nc)).tocsc()  # pattern
        coo.sort_indices()
        lookup = {(r, c): v for r, c, v in zip(rows, cols, vals)}
        data = np.array([lookup[(r, c)] for c in range(nc) for r in coo.indices[coo.indptr[c]:coo.indptr[c + 1]]])
        M.indptr, M.indices, M.data = coo.indptr.copy(), coo.indices.copy(), data
        P.append(M)
        nf = nc
    if negative_column is not None:
        lv, c = negative_column
        M = P[lv]
        M.data[M.indptr[c]:M.indptr[c + 1]] = -np.abs(M.data[M.indptr[c]:M.indptr[c + 1]]) - 0.25
    return A, P


def _same_matrix(a, b):
    return (a.shape == b.shape and np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
            and np.array_equal(a.data, b.data))


def _three_ways(A, P, known):
    port = Oracle(P).precompute(A, known)
    ref = Oracle(P, impl="ref").precompute(A, known)
    plan = Solver(device="none").set_hierarchy(P).precompute(A, known)
    nlev = len(P) + 1
    assert np.array_equal(port.unknown, ref.unknown) and np.array_equal(plan.unknown, ref.unknown)
    for lv in range(nlev):
        assert _same_matrix(port.matrix(lv, "A"), ref.matrix(lv, "A")), lv
        a, b = plan.matrix(lv, "A", values=False), ref.matrix(lv, "A")
        assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices), lv
        assert np.array_equal(port.diag(lv), ref.diag(lv)), lv
        if lv >= 1:
            for which in ("P", "PT"):
                assert _same_matrix(port.matrix(lv, which), ref.matrix(lv, which)), (lv, which)
                assert _same_matrix(plan.matrix(lv, which), ref.matrix(lv, which)), (lv, which)
    if known is not None:
        assert _same_matrix(port.matrix(0, "Auk"), ref.matrix(0, "Auk"))
    return port, ref, plan


def _solves_agree(port, ref, rhs, z0, kv, tol, max_iter):
    z1, r1, ok1 = port.solve(rhs, z0, kv, tol, max_iter)
    z2, r2, ok2 = ref.solve(rhs, z0, kv, tol, max_iter)
    assert ok1 == ok2 and len(r1) == len(r2)
    assert np.allclose(r1, r2, rtol=1e-9, atol=1e-15)
    assert np.allclose(z1, z2, rtol=1e-9, atol=1e-13)
    return z2, r2, ok2


def test_pruning_cascades_over_two_levels():
    A, P = chain_hierarchy(33, 3)
    known = np.array([0, 1, 2, 3], dtype=np.int32)
    port, ref, plan = _three_ways(A, P, known)
    # level 1 loses the columns whose support is fixed (0 and 1), which removes two ROWS of the
    # next prolongation, which in turn empties its column 0
    assert ref.matrix(1, "P").shape == (29, 15) and ref.matrix(2, "P").shape == (15, 8)
    assert np.array_equal(plan.keep(1), np.arange(2, 17)) and np.array_equal(plan.keep(2), np.arange(1, 9))
    assert np.array_equal(port.keep(1), plan.keep(1)) and np.array_equal(port.keep(2), plan.keep(2))
    rng = np.random.default_rng(0)
    _solves_agree(port, ref, rng.standard_normal(33), rng.standard_normal(33), rng.standard_normal(4), 1e-10, 40)


def test_pruning_stops_at_the_first_level_that_keeps_everything():
    A, P = chain_hierarchy(33, 3)
    known = np.array([16], dtype=np.int32)  # an interior even node: every coarse column keeps support
    port, ref, plan = _three_ways(A, P, known)
    assert plan.keep(1) is None and plan.keep(2) is None
    assert ref.matrix(1, "P").shape == (32, 17)


def test_pruning_threshold_is_signed():
    """`it.value() > 1e-15` (cpp:197): a column with only negative weights is dropped"""
    A, P = chain_hierarchy(33, 3, negative_column=(0, 5))
    known = np.array([0], dtype=np.int32)
    port, ref, plan = _three_ways(A, P, known)
    assert 5 not in plan.keep(1) and ref.matrix(1, "P").shape[1] == len(plan.keep(1))


@pytest.mark.parametrize("known", [np.array([7, 2, 7, 30, 2], dtype=np.int32),  # duplicates, unsorted
                                   np.array([], dtype=np.int32)])               # fixed variant, nothing fixed
def test_duplicate_unsorted_and_empty_known(known):
    A, P = chain_hierarchy(33, 3)
    port, ref, plan = _three_ways(A, P, known)
    rng = np.random.default_rng(1)
    kv = rng.standard_normal(known.size)
    z, r_his, ok = _solves_agree(port, ref, rng.standard_normal(33), rng.standard_normal(33), kv, 1e-10, 40)
    # igl::slice_into writes sequentially: the LAST occurrence of a repeated index wins (cpp:355)
    for idx in np.unique(known):
        assert z[idx] == kv[np.flatnonzero(known == idx)[-1]]


def test_immediate_convergence_and_single_iteration():
    A, P = chain_hierarchy(33, 3)
    known = np.array([0, 32], dtype=np.int32)
    port, ref, plan = _three_ways(A, P, known)
    rng = np.random.default_rng(2)
    rhs, z0, kv = rng.standard_normal(33), rng.standard_normal(33), rng.standard_normal(2)
    # tolerance above the first residual: one measurement, no cycle, z = z0 with known overwritten
    z, r_his, ok = _solves_agree(port, ref, rhs, z0, kv, 1e6, 20)
    assert ok and len(r_his) == 1
    unknown = np.setdiff1d(np.arange(33), known)
    assert np.array_equal(z[unknown], z0[unknown]) and np.array_equal(z[known], kv)
    # maxIter = 1: one measurement, one cycle, "not converged" judged on the stale residual
    z, r_his, ok = _solves_agree(port, ref, rhs, z0, kv, 1e-12, 1)
    assert not ok and len(r_his) == 1 and not np.array_equal(z[unknown], z0[unknown])


def test_three_columns_with_fixed_values():
    A, P = chain_hierarchy(65, 4)
    known = np.array([0, 64, 31], dtype=np.int32)
    port, ref, plan = _three_ways(A, P, known)
    rng = np.random.default_rng(3)
    rhs, z0 = np.asfortranarray(rng.standard_normal((65, 3))), np.asfortranarray(rng.standard_normal((65, 3)))
    kv = np.asfortranarray(rng.standard_normal((3, 3)))
    z, r_his, ok = _solves_agree(port, ref, rhs, z0, kv, 1e-10, 40)
    assert ok and np.array_equal(z[known, :], kv)
    # the Frobenius norm over all columns drives the loop (cpp:332)
    r0 = np.linalg.norm((rhs - A @ np.where(np.isin(np.arange(65), known)[:, None], 0, z0) )[np.setdiff1d(np.arange(65), known)] -
                        (A[:, known] @ kv)[np.setdiff1d(np.arange(65), known)])
    assert abs(r_his[0] - r0) <= 1e-12 * r0
