"""GPU tests of the row-partitioned (multi-GPU) path, include/smg.h "multi-GPU" block.

The ranks run as threads of this process and, on a single-GPU box, share device 0: the
partition, the peer-memory halo exchange kernels and the collective call semantics are
exactly those of one process per GPU (tests/dist_worker.py covers CUDA IPC between
processes when two devices are present).

Parity bars:
  * linear operators (A, residual, restrict, prolong, residual norm) on N ranks: equal to
    the oracle to summation-order rounding (rows are stored sorted by permuted column and
    the permutation depends on the partition): abs 1e-13 x magnitude;
  * exact mode (halo exchange after every colour): relax / V-cycle are the single-GPU
    multicolour smoother with the same colouring: rel 1e-12 against the 1-rank library;
  * hybrid mode (one exchange per sweep): same fixed point: final x rel 1e-7 against the
    oracle when solved to 1e-10, at most 4 extra V-cycles.
"""
import numpy as np
import pytest

from dist_util import run_ranks
from oracle.cpu_oracle import Oracle
from surface_multigrid_code_b200.solver import Solver

pytestmark = pytest.mark.gpu


def _rand(rng, n, k):
    return rng.standard_normal(n) if k == 1 else np.asfortranarray(rng.standard_normal((n, k)))


def _close(a, b, tol=1e-13):
    return np.allclose(a, b, rtol=0, atol=tol * max(1.0, float(np.abs(b).max())))


@pytest.mark.parametrize("world,dist_levels", [(2, 1), (2, 2), (3, 2), (4, 1)])
@pytest.mark.parametrize("name", ["sphere_pad", "grid", "mcf"])
def test_linear_operators_match_oracle(problems, name, world, dist_levels):
    pr = problems[name]
    ora = Oracle(pr.P).precompute(pr.A, pr.known)
    rng = np.random.default_rng(5)
    k = pr.k
    vecs = {lv: (_rand(rng, ora.level_rows(lv), k), _rand(rng, ora.level_rows(lv), k))
            for lv in range(pr.nlev)}

    def fn(rank, s):
        s.set_hierarchy(pr.P).precompute(pr.A, pr.known)
        s.barrier()
        out = {}
        for lv in range(pr.nlev):
            u, b = vecs[lv]
            out[lv, "A"] = s.apply_A(lv, u)
            out[lv, "res"] = s.residual(lv, b, u)
            out[lv, "norm"] = s.residual_norm(lv, b, u)
            if lv + 1 < pr.nlev:
                out[lv, "R"] = s.restrict(lv, u)
                out[lv, "P"] = s.prolong(lv, vecs[lv + 1][0])
        out["info"] = s.dist_info()
        return out

    res = run_ranks(world, fn, dist_levels=dist_levels)
    for rank, out in enumerate(res):
        assert out["info"]["dist_levels"] == dist_levels and out["info"]["exchanges"] > 0
        for lv in range(pr.nlev):
            u, b = vecs[lv]
            assert _close(out[lv, "A"], ora.apply_A(lv, u)), (rank, lv)
            assert _close(out[lv, "res"], b - ora.apply_A(lv, u)), (rank, lv)
            ref = np.linalg.norm(b - ora.apply_A(lv, u))
            assert abs(out[lv, "norm"] - ref) <= 1e-12 * ref
            if lv + 1 < pr.nlev:
                assert _close(out[lv, "R"], ora.restrict(lv, u)), (rank, lv)
                assert _close(out[lv, "P"], ora.prolong(lv, vecs[lv + 1][0])), (rank, lv)
    # every rank returns bit-identical complete vectors
    for out in res[1:]:
        for key, v in res[0].items():
            if key != "info":
                assert np.array_equal(np.asarray(v), np.asarray(out[key])), key


@pytest.mark.parametrize("world,dist_levels", [(2, 1), (3, 2)])
@pytest.mark.parametrize("name", ["sphere_pad", "grid", "mcf"])
def test_exact_mode_is_the_single_gpu_smoother(problems, name, world, dist_levels):
    pr = problems[name]
    rng = np.random.default_rng(9)
    k = pr.k
    with Solver(device=0) as one:
        one.set_hierarchy(pr.P).precompute(pr.A, pr.known)
        n0 = one.level_rows(0)
        u0, b0 = _rand(rng, n0, k), _rand(rng, n0, k)
        ref_relax = {lv: None for lv in range(pr.nlev)}
        vec = {}
        for lv in range(pr.nlev):
            n = one.level_rows(lv)
            vec[lv] = (_rand(rng, n, k), _rand(rng, n, k))
            ref_relax[lv] = one.relax(lv, 2, vec[lv][1], vec[lv][0])
        ref_v = one.vcycle(0, b0, u0)
        ref_z, ref_r, ref_ok = one.solve(pr.rhs, pr.z0, pr.known_val, 1e-10, 30)

    def fn(rank, s):
        s.set_hierarchy(pr.P).precompute(pr.A, pr.known)
        s.barrier()
        out = {lv: s.relax(lv, 2, vec[lv][1], vec[lv][0]) for lv in range(pr.nlev)}
        out["v"] = s.vcycle(0, b0, u0)
        out["solve"] = s.solve(pr.rhs, pr.z0, pr.known_val, 1e-10, 30)
        return out

    for out in run_ranks(world, fn, exact=True, dist_levels=dist_levels):
        for lv in range(pr.nlev):
            assert _close(out[lv], ref_relax[lv], 1e-12), lv
        assert _close(out["v"], ref_v, 1e-11)
        z, r_his, ok = out["solve"]
        assert ok == ref_ok and len(r_his) == len(ref_r)
        assert np.allclose(r_his, ref_r, rtol=1e-6, atol=1e-16)
        assert np.linalg.norm(z - ref_z) <= 1e-9 * np.linalg.norm(ref_z)


@pytest.mark.parametrize("world,halo", [(2, 0), (4, 0), (3, 2)])
@pytest.mark.parametrize("name", ["sphere_pad", "sphere", "grid", "mcf"])
def test_hybrid_solve_converges_to_the_oracle_solution(problems, name, world, halo):
    pr = problems[name]
    ora = Oracle(pr.P).precompute(pr.A, pr.known)
    z_ref, r_ref, ok_ref = ora.solve(pr.rhs, pr.z0, pr.known_val, 1e-10, 40)
    assert ok_ref

    def fn(rank, s):
        s.set_hierarchy(pr.P).precompute(pr.A, pr.known)
        s.barrier()
        assert np.array_equal(s.unknown, ora.unknown)
        return s.solve(pr.rhs, pr.z0, pr.known_val, 1e-10, 40)

    res = run_ranks(world, fn, exact=halo)
    for z, r_his, ok in res:
        assert ok and r_his[-1] < 1e-10
        assert len(r_his) <= len(r_ref) + (4 if halo == 0 else 8)
        assert np.linalg.norm(z - z_ref) <= 1e-7 * np.linalg.norm(z_ref)
    for z, r_his, ok in res[1:]:
        assert np.array_equal(z, res[0][0]) and np.array_equal(r_his, res[0][1])


def test_wavefront_smoother_partitioned(problems):
    """the order-exact smoother also runs partitioned (exact mode: a halo exchange after every
    wavefront level): same iterates as the oracle's lexicographic Gauss-Seidel to rounding of
    the reordered row sums"""
    pr = problems["sphere_pad"]
    ora = Oracle(pr.P).precompute(pr.A, pr.known)
    rng = np.random.default_rng(2)
    n0 = ora.level_rows(0)
    u, b = rng.standard_normal(n0), rng.standard_normal(n0)
    ref = ora.relax(0, 2, b, u.copy())  # the oracle relaxes in place

    def fn(rank, s):
        s.set_hierarchy(pr.P).precompute(pr.A, pr.known)
        s.barrier()
        return s.relax(0, 2, b, u)

    for out in run_ranks(2, fn, smoother="wavefront", exact=True):
        assert np.array_equal(out, ref)


def test_two_devices_peer_access(problems):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    pr = problems["sphere"]
    ora = Oracle(pr.P).precompute(pr.A, pr.known)
    z_ref, _, _ = ora.solve(pr.rhs, pr.z0, pr.known_val, 1e-10, 40)

    def fn(rank, s):
        s.set_hierarchy(pr.P).precompute(pr.A, pr.known)
        s.barrier()
        return s.solve(pr.rhs, pr.z0, pr.known_val, 1e-10, 40)

    for z, r_his, ok in run_ranks(2, fn, devices=[0, 1]):
        assert ok and np.linalg.norm(z - z_ref) <= 1e-7 * np.linalg.norm(z_ref)
