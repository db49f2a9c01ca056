"""CPU tests of the multi-GPU row partition (plan-only handles, no CUDA): ownership,
row ordering and the halo-exchange lists of include/smg.h's "multi-GPU" block are checked
against a direct numpy restatement of "which rows does rank d read that rank s owns".
The reference has no counterpart (single-threaded CPU code); the
partition is BASELINE.json's north_star row."""
import numpy as np
import pytest
import scipy.sparse as sp

from surface_multigrid_code_b200.solver import SmgError, Solver


def _plan(pr, rank, world, dist_levels=-1, min_rows=0, smoother="multicolour"):
    s = Solver(smoother=smoother, device="none")
    s.dist_init(rank, world).dist_options(False, dist_levels, min_rows)
    return s.set_hierarchy(pr.P).precompute(pr.A, pr.known)


def _needed(M, part_rows, part_cols, src, dst):
    """columns j owned by src that a row i owned by dst reads (M[i, j] != 0)."""
    M = sp.coo_matrix(M)
    nz = M.data != 0
    i, j = M.row[nz], M.col[nz]
    m = (part_rows[i] == dst) & (part_cols[j] == src)
    return np.unique(j[m])


@pytest.mark.parametrize("world", [2, 3, 4])
def test_partition_and_halo_lists(problems, world):
    pr = problems["sphere"]  # no explicit zeros: pattern == non-zero pattern
    plans = [_plan(pr, r, world, dist_levels=2) for r in range(world)]
    s = plans[0]
    info = s.dist_info()
    assert info["world"] == world and info["dist_levels"] == 2
    nlev = s.num_levels()
    parts = [s.dist_part(lv) for lv in range(nlev)]
    for lv in range(nlev):
        li = s.dist_level_info(lv)
        assert li["layout"] == (1 if lv < 2 else (2 if lv == 2 else 0))
        assert li["parts"] == (world if lv <= 2 else 1)
    # level 0: equal strips of the breadth-first order
    cnt = np.bincount(parts[0], minlength=world)
    assert cnt.sum() == s.level_rows(0) and cnt.max() - cnt.min() <= 1
    # every rank reports its own contiguous row range, together they tile the level
    ranges = sorted((p.dist_level_info(0)["own_begin"], p.dist_level_info(0)["own_end"]) for p in plans)
    assert ranges[0][0] == 0 and ranges[-1][1] == s.level_rows(0)
    assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    assert [p.dist_level_info(0)["own_rows"] for p in plans] == list(cnt)
    for lv in range(2):
        A = s.matrix(lv, "A", values=False)
        A.data[:] = 1.0
        P = plans[0].matrix(lv + 1, "P")  # rows: level lv, columns: level lv+1
        for src in range(world):
            for dst in range(world):
                if src == dst:
                    for w in ("halo_u", "halo_r", "gather"):
                        assert s.dist_exchange(lv, w, src, dst).size == 0
                    continue
                got = np.sort(s.dist_exchange(lv, "halo_u", src, dst))
                assert np.array_equal(got, _needed(A, parts[lv], parts[lv], src, dst))
                # restriction rows (coarse, owned by dst) read fine r owned by src: PT = P^T
                got = np.sort(s.dist_exchange(lv, "halo_r", src, dst))
                assert np.array_equal(got, _needed(P.T, parts[lv + 1], parts[lv], src, dst))
                got = np.sort(s.dist_exchange(lv, "gather", src, dst))
                assert np.array_equal(got, np.flatnonzero(parts[lv] == src))
                # every rank plans the same lists
                for p in plans[1:]:
                    assert np.array_equal(np.sort(p.dist_exchange(lv, "halo_u", src, dst)),
                                          np.sort(s.dist_exchange(lv, "halo_u", src, dst)))
    # prolongation rows of level 0 (owned by dst) read u of level 1 owned by src
    P1 = s.matrix(1, "P")
    for src in range(world):
        for dst in range(world):
            if src != dst:
                got = np.sort(s.dist_exchange(1, "halo_pu", src, dst))
                assert np.array_equal(got, _needed(P1, parts[0], parts[1], src, dst))
    # the split level (2) is replicated: only the all-gather of its right-hand side
    assert s.dist_exchange(2, "halo_u", 0, 1).size == 0
    got = np.sort(s.dist_exchange(2, "gather", 1, 0))
    assert np.array_equal(got, np.flatnonzero(parts[2] == 1))
    # a coarse row is owned by the rank of its heaviest fine row
    Pc = sp.csc_matrix(P1)
    heavy = np.array([Pc.indices[Pc.indptr[c]:Pc.indptr[c + 1]][np.argmax(np.abs(Pc.data[Pc.indptr[c]:Pc.indptr[c + 1]]))]
                      for c in range(Pc.shape[1])])
    assert np.array_equal(parts[1], parts[0][heavy])


def test_halo_is_small_and_index_outputs_do_not_depend_on_world(problems):
    pr = problems["sphere"]
    one = Solver(device="none").set_hierarchy(pr.P).precompute(pr.A, pr.known)
    two = _plan(pr, 1, 2)
    assert two.dist_info()["dist_levels"] == 1  # automatic: level 0 only on a small mesh
    li = two.dist_level_info(0)
    assert 0 < li["halo_u_recv"] < 0.2 * li["own_rows"]
    # the parity surface (unknown, patterns, colours) is that of the single-GPU plan
    assert np.array_equal(one.unknown, two.unknown)
    for lv in range(one.num_levels()):
        a, b = one.matrix(lv, "A", values=False), two.matrix(lv, "A", values=False)
        assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
        assert np.array_equal(one.phases(lv)[1], two.phases(lv)[1])


def test_padded_zero_entries_do_not_create_halo(problems):
    """explicit zeros of P (get_prolong.cpp:45-56) become structural zeros of A_l; they are
    not part of the compute pattern and must not be exchanged."""
    pr = problems["sphere_pad"]
    s = _plan(pr, 0, 2, dist_levels=2)
    part = s.dist_part(1)
    A1 = s.matrix(1, "A", values=False)
    A1.data[:] = 1.0
    with_zeros = _needed(A1, part, part, 1, 0)
    got = s.dist_exchange(1, "halo_u", 1, 0)
    assert got.size <= with_zeros.size and np.all(np.isin(got, with_zeros))


def test_dist_argument_checks(problems):
    s = Solver(device="none")
    with pytest.raises(SmgError):
        s.dist_init(2, 2)
    with pytest.raises(SmgError):
        s.dist_init(0, 0)
    s.dist_init(0, 2)
    with pytest.raises(SmgError):
        s.dist_handle()  # plan-only: nothing to export


@pytest.mark.parametrize("world", [1, 3])
def test_row_numbering_groups_rows_by_part_and_phase(problems, world):
    pr = problems["sphere"]
    s = _plan(pr, 1 % world, world, dist_levels=2) if world > 1 else Solver(device="none").set_hierarchy(pr.P).precompute(pr.A, pr.known)
    for lv in range(s.num_levels()):
        n = s.level_rows(lv)
        perm, gp = s.row_order(lv)
        assert np.array_equal(np.sort(perm), np.arange(n))  # a permutation
        nph, phase = s.phases(lv)
        part = s.dist_part(lv) if world > 1 else np.zeros(n, dtype=np.int32)
        li = s.dist_level_info(lv) if world > 1 else {"layout": 0, "parts": 1}
        W = li["parts"]
        assert gp[0] == 0 and gp[-1] == n and np.all(np.diff(gp) >= 0) and len(gp) - 1 == W * nph
        # rows of a group share (part, phase); groups are ordered (part, phase) on a partitioned
        # level, (phase, part) on the split level, (phase) otherwise
        for g in range(len(gp) - 1):
            rows = perm[gp[g]:gp[g + 1]]
            if li["layout"] == 1:
                want_part, want_phase = divmod(g, nph)
            elif li["layout"] == 2:
                want_phase, want_part = divmod(g, W)
            else:
                want_part, want_phase = 0, g
            assert np.all(phase[rows] == want_phase) and np.all(part[rows] == want_part), (lv, g)
        if li["layout"] == 1:  # this rank's rows are one contiguous range
            me = s.dist_info()["rank"]
            assert np.all(part[perm[li["own_begin"]:li["own_end"]]] == me)
            assert li["own_end"] - li["own_begin"] == int(np.sum(part == me))
