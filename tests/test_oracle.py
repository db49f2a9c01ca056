"""CPU tests that pin the oracle (oracle/smg_oracle.c).

The reference has no tests or golden vectors for the hot path and cannot be built here
(Eigen absent), so the oracle is pinned against (1) an independent numpy/scipy
restatement (tests/scipy_restatement.py), (2) the committed fixtures generated from it
(tests/golden), (3) a direct sparse solve of the constrained system.
"""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import golden_util
from oracle.cpu_oracle import Oracle
from scipy_restatement import Hierarchy, gauss_seidel

NAMES = ["sphere_pad", "grid", "mcf", "block"]


@pytest.mark.parametrize("name", NAMES)
def test_precompute_matches_scipy(problems, name):
    pr = problems[name]
    ora = Oracle(pr.P).precompute(pr.A, pr.known)
    H = Hierarchy(pr.A, pr.P, pr.known)
    assert np.array_equal(ora.unknown, H.unknown)
    for lv in range(pr.nlev):
        a, b = ora.matrix(lv, "A"), H.A[lv]
        assert a.shape == b.shape
        scale = abs(b).max()
        assert abs(a - b).max() <= 1e-13 * scale
        assert np.allclose(ora.diag(lv), b.diagonal(), rtol=1e-13, atol=0)
        if lv >= 1:
            assert abs(ora.matrix(lv, "P") - H.P[lv - 1]).max() == 0
            assert abs(ora.matrix(lv, "PT") - H.P[lv - 1].T).max() == 0
    if pr.known is not None:
        for lv, keep in enumerate(H.keep, start=1):
            k2 = ora.keep(lv)
            assert (keep is None) == (k2 is None)
            if keep is not None:
                assert np.array_equal(keep, k2)


def test_explicit_zeros_of_P_propagate_into_coarse_patterns(problems):
    """Eigen's conservative products keep structural zeros: with the
    reference's 3-entries-per-row P the coarse patterns are supersets of P^T A P's
    numerical pattern."""
    pr = problems["sphere_pad"]
    ora = Oracle(pr.P).precompute(pr.A, pr.known)
    a1 = ora.matrix(1, "A")
    numeric = a1.copy()
    numeric.eliminate_zeros()
    assert a1.nnz > numeric.nnz
    assert ora.matrix(1, "P").nnz == 3 * ora.level_rows(0)


@pytest.mark.parametrize("name", NAMES)
def test_operators_match_scipy(problems, name):
    pr = problems[name]
    ora = Oracle(pr.P).precompute(pr.A, pr.known)
    H = Hierarchy(pr.A, pr.P, pr.known)
    rng = np.random.default_rng(0)
    k = pr.k
    for lv in range(pr.nlev):
        n = ora.level_rows(lv)
        u = rng.standard_normal(n) if k == 1 else rng.standard_normal((n, k))
        b = rng.standard_normal(n) if k == 1 else rng.standard_normal((n, k))
        assert np.allclose(ora.apply_A(lv, u), H.A[lv] @ u, rtol=1e-12, atol=1e-12)
        if n <= 1200:
            ref = gauss_seidel(H.A[lv], H.diag[lv], b, u, 2)
            assert np.allclose(ora.relax(lv, 2, b, u), ref, rtol=1e-11, atol=1e-11)
        if lv + 1 < pr.nlev:
            nc = ora.level_rows(lv + 1)
            uc = rng.standard_normal(nc) if k == 1 else rng.standard_normal((nc, k))
            assert np.allclose(ora.restrict(lv, u), H.P[lv].T @ u, rtol=1e-12, atol=1e-12)
            assert np.allclose(ora.prolong(lv, uc), H.P[lv] @ uc, rtol=1e-12, atol=1e-12)
    nc = ora.level_rows(pr.nlev - 1)
    b = rng.standard_normal(nc) if k == 1 else rng.standard_normal((nc, k))
    x = ora.coarse_solve(b, np.zeros_like(b))
    assert np.allclose(H.A[-1] @ x, b, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("name", golden_util.NAMES)
def test_solve_matches_golden(name):
    g = golden_util.load(name)
    ora = Oracle(g["P"]).precompute(g["A"], g["known"])
    assert np.array_equal(ora.unknown, g["unknown"])
    for lv in range(g["nlev"]):
        assert np.allclose(ora.diag(lv), g["diag"][lv], rtol=1e-12)
        a = ora.matrix(lv, "A")
        assert np.allclose(np.asarray(a.sum(axis=1)).ravel(), g["rowsum"][lv], rtol=1e-9, atol=1e-12)
    z, r_his, ok = ora.solve(g["rhs"], g["z0"], g["known_val"], g["tol"], g["max_iter"])
    assert ok == g["converged"] and len(r_his) == len(g["r_his"])
    assert np.allclose(r_his, g["r_his"], rtol=1e-6, atol=1e-14)
    assert np.linalg.norm(z - g["z"]) <= 1e-9 * np.linalg.norm(g["z"])


def test_solution_matches_direct_solve(problems):
    pr = problems["sphere_pad"]
    ora = Oracle(pr.P).precompute(pr.A, pr.known)
    z, r_his, ok = ora.solve(pr.rhs, pr.z0, pr.known_val, 1e-11, 40)
    assert ok
    unk = ora.unknown
    A = pr.A.tocsr()
    x = spla.spsolve(sp.csc_matrix(A[unk][:, unk]), pr.rhs[unk] - A[unk][:, pr.known] @ pr.known_val)
    assert np.linalg.norm(z[unk] - x) <= 1e-7 * np.linalg.norm(x)
    assert np.array_equal(z[pr.known], pr.known_val)


def test_solve_quirks(problems):
    """residual measured before each cycle; r_his length; strict `<`
    break and `!(residual > tol)` return with the stale residual."""
    pr = problems["sphere_pad"]
    ora = Oracle(pr.P).precompute(pr.A, pr.known)
    z, r_his, ok = ora.solve(pr.rhs, pr.z0, pr.known_val, 1e-30, 4)
    assert len(r_his) == 4 and not ok  # maxIter measurements, maxIter cycles
    z5, r5, _ = ora.solve(pr.rhs, pr.z0, pr.known_val, 1e-30, 5)
    assert np.array_equal(r5[:4], r_his)
    # z includes the last cycle although its residual was never measured
    unk = ora.unknown
    A = pr.A.tocsr()
    true_res = np.linalg.norm((pr.rhs - A @ z)[unk])
    assert true_res < r_his[-1] and abs(true_res - r5[4]) <= 1e-9 * r5[4]
    z1, r1, ok1 = ora.solve(pr.rhs, pr.z0, pr.known_val, 1e30, 5)
    assert len(r1) == 1 and ok1 and np.array_equal(z1[unk], pr.z0[unk])
    # tolerance equal to the measured residual: no break (strict <), but "converged"
    z2, r2, ok2 = ora.solve(pr.rhs, pr.z0, pr.known_val, float(r_his[0]), 1)
    assert len(r2) == 1 and ok2 and not np.array_equal(z2[unk], pr.z0[unk])


def test_known_order_and_duplicates(problems):
    """`known` stays in caller order (Auk columns), `unknown` is the sorted complement."""
    pr = problems["sphere"]
    known = np.array([5, 0, 3, 3, 17], dtype=np.int32)
    ora = Oracle(pr.P).precompute(pr.A, known)
    unk = ora.unknown
    assert np.array_equal(unk, np.setdiff1d(np.arange(pr.n), known))
    auk = ora.matrix(0, "Auk")
    A = pr.A.tocsr()
    assert abs(auk - A[unk][:, known]).max() == 0


def test_bench_cpu_arms_run_on_a_small_problem():
    """bench.py's CPU legs (cpu_baseline of the GPU line, the --impl reference line) on a small
    workload: they must run without a GPU and name what they timed."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench

    import argparse

    pr = bench.build_problem(argparse.Namespace(workload="sphere", subdiv=5, levels=3, max_iter=20, tol=None))
    cpu = bench.cpu_baseline(pr, 3)
    assert cpu["value"] > 0 and cpu["cores"] == 1 and cpu["kind"] in ("reference", "port")
    if cpu["kind"] == "reference":
        assert cpu["port_value"] > 0
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--subdiv", "5",
                          "--levels", "3", "--steps", "2", "--warmup", "1"], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == cpu["kind"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["cpu_parallel"]["value"] > 0 and line["cpu_parallel"]["kind"] == "variant"


def test_parallel_cpu_variant_reaches_the_same_solution(problems):
    """orc_iterate_mt (OpenMP multicolour Gauss-Seidel; context for the benchmark, not the
    reference algorithm) converges to the solution of the reference-order iteration"""
    pr = problems["sphere_pad"]
    ora = Oracle(pr.P).precompute(pr.A, pr.known)
    bu = np.ascontiguousarray(pr.rhs[ora.unknown])
    z_seq, r_seq = ora.iterate(bu, np.zeros_like(bu), 14)
    for threads in (1, 3):
        z_mt, r_mt, used = ora.iterate_mt(bu, np.zeros_like(bu), 14, threads)
        assert used >= 1 and r_mt[-1] < 1e-9 * r_mt[0]
        assert np.linalg.norm(z_mt - z_seq) <= 1e-8 * np.linalg.norm(z_seq)
