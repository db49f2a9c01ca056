"""CPU tests of the host side of libsmg.so: the C ABI loads and exports every symbol
include/smg.h declares; the index/topology planning (plan-only handles, no CUDA) is
bit-identical to the oracle; there is no CPU fallback."""
import os
import re

import numpy as np
import pytest

import golden_util
from oracle.cpu_oracle import Oracle
from surface_multigrid_code_b200 import _lib
from surface_multigrid_code_b200.solver import SmgError, Solver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "smg.h")).read()
    declared = set(re.findall(r"^(?:int|void|const char|int64_t)\s*\**\s*(smg_\w+)\s*\(", header, re.M))
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in smg.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes prototype"
    assert lib.smg_version() == 100
    assert lib.smg_status_string(2) == b"CUDA error or no device"


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(SmgError) as e:
        Solver()
    assert e.value.status == 2  # SMG_E_CUDA


def test_plan_only_handle_refuses_compute(problems):
    pr = problems["sphere_pad"]
    s = Solver(device="none").set_hierarchy(pr.P).precompute(pr.A, pr.known)
    for call in (lambda: s.solve(pr.rhs, pr.z0, pr.known_val),
                 lambda: s.relax(0, 1, np.zeros(s.level_rows(0)), np.zeros(s.level_rows(0))),
                 lambda: s.time_kernel("residual")):
        with pytest.raises(SmgError) as e:
            call()
        assert e.value.status == 5  # SMG_E_STATE


def _check_index_outputs(P, A, known, smoother):
    ora = Oracle(P).precompute(A, known)
    s = Solver(smoother=smoother, device="none").set_hierarchy(P).precompute(A, known)
    nlev = len(P) + 1
    assert s.num_levels() == nlev
    assert np.array_equal(s.unknown, ora.unknown)
    for lv in range(nlev):
        assert s.level_rows(lv) == ora.level_rows(lv)
        a, b = s.matrix(lv, "A", values=False), ora.matrix(lv, "A")
        assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
        if lv >= 1:
            ks, ko = s.keep(lv), ora.keep(lv)
            assert (ks is None) == (ko is None)
            if ks is not None:
                assert np.array_equal(ks, ko)
            for w in ("P", "PT"):
                a, b = s.matrix(lv, w), ora.matrix(lv, w)
                assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
                assert np.array_equal(a.data, b.data)
    if known is not None:
        for w in ("LHS", "Auk"):
            a, b = s.matrix(0, w, values=False), ora.matrix(0, w)
            assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
    return s, ora


@pytest.mark.parametrize("name", ["sphere_pad", "sphere", "grid", "mcf"])
@pytest.mark.parametrize("smoother", ["multicolour", "wavefront"])
def test_index_outputs_bit_exact(problems, name, smoother):
    pr = problems[name]
    _check_index_outputs(pr.P, pr.A, pr.known, smoother)


@pytest.mark.parametrize("name", golden_util.NAMES)
def test_index_outputs_golden(name):
    g = golden_util.load(name)
    s, _ = _check_index_outputs(g["P"], g["A"], g["known"], "multicolour")
    assert np.array_equal(s.unknown, g["unknown"])


def test_known_edge_cases(problems):
    pr = problems["sphere"]
    # empty known list (fixed variant with nothing fixed), repeated and unsorted indices
    for known in (np.zeros(0, dtype=np.int32), np.array([9, 2, 2, 40, 0], dtype=np.int32)):
        _check_index_outputs(pr.P, pr.A, known, "multicolour")
    with pytest.raises(SmgError):
        Solver(device="none").set_hierarchy(pr.P).precompute(pr.A, np.array([pr.n], dtype=np.int32))


@pytest.mark.parametrize("smoother", ["multicolour", "wavefront"])
def test_smoother_schedule_is_valid(problems, smoother):
    """Rows of one phase never touch each other through a non-zero-capable entry; the
    wavefront schedule additionally respects the lexicographic order."""
    pr = problems["grid"]
    s = Solver(smoother=smoother, device="none").set_hierarchy(pr.P).precompute(pr.A, pr.known)
    ora = Oracle(pr.P).precompute(pr.A, pr.known)
    for lv in range(pr.nlev):
        nph, phase = s.phases(lv)
        A = ora.matrix(lv, "A").tocoo()
        live = (A.data != 0) & (A.row != A.col)
        r, c = A.row[live], A.col[live]
        assert not np.any(phase[r] == phase[c])
        assert phase.min() == 0 and phase.max() == nph - 1
        if smoother == "wavefront":
            lower = c < r  # row r reads the already-updated u[c]
            assert np.all(phase[c[lower]] < phase[r[lower]])
        st = s.level_stats(lv)
        assert st["nnz"] <= st["nnz_ref"] and st["padded"] >= st["nnz"]


def test_hierarchy_validation(problems):
    pr = problems["sphere_pad"]
    with pytest.raises(SmgError) as e:
        Solver(device="none").set_hierarchy([])
    assert e.value.status == 3  # SMG_E_NLEVELS: the nLvs == 1 mis-solve is not replicated
    with pytest.raises(SmgError):
        Solver(device="none").set_hierarchy(pr.P).precompute(pr.A[:-1][:, :-1], None)
    with pytest.raises(SmgError):
        Solver(device="none").precompute(pr.A, pr.known)  # no hierarchy yet


def test_file_rendezvous_all_gathers_blobs(tmp_path):
    """smg_rendezvous_files (include/smg.h): the launcher-free way for N processes of the
    reference's examples to exchange their export blobs.  Ranks = threads here."""
    import ctypes as C
    import threading

    lib = _lib.load()
    world, nbytes = 4, 96
    mine = [bytes([r]) * nbytes for r in range(world)]
    got = [None] * world
    rcs = [None] * world

    def body(r):
        out = C.create_string_buffer(world * nbytes)
        rcs[r] = lib.smg_rendezvous_files(str(tmp_path).encode(), b"t0", r, world, mine[r], nbytes, out, 20000)
        got[r] = out.raw

    ts = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(60)
    assert rcs == [0] * world
    assert all(g == b"".join(mine) for g in got)
    # a rank that never shows up: timeout, not a hang
    out = C.create_string_buffer(2 * nbytes)
    assert lib.smg_rendezvous_files(str(tmp_path).encode(), b"t1", 0, 2, mine[0], nbytes, out, 200) == 10
    assert lib.smg_rendezvous_files(b"/nonexistent-dir", b"t", 0, 2, mine[0], nbytes, out, 200) == 1


def test_file_rendezvous_ignores_leftovers_of_a_dead_run(tmp_path):
    """A blob left behind by an earlier run with the same directory and tag (its writer is
    dead) must never be taken for a peer's blob; blobs are removed once every peer has them."""
    import ctypes as C
    import struct
    import subprocess
    import sys
    import threading
    import time

    lib = _lib.load()
    nbytes = 64
    p = subprocess.Popen([sys.executable, "-c", "pass"])
    p.wait()
    dead_pid = p.pid
    magic = 0x534D47525A563032
    stale = struct.pack("<Qqq", magic, dead_pid, 1) + b"\xee" * nbytes
    (tmp_path / "job.1").write_bytes(stale)
    (tmp_path / "job.1.ack").write_bytes(struct.pack("<Qqq", magic, dead_pid, 1) + b"\0" * 32)
    mine = [b"\x11" * nbytes, b"\x22" * nbytes]
    got, rcs = [None, None], [None, None]

    def body(r, delay):
        time.sleep(delay)
        out = C.create_string_buffer(2 * nbytes)
        rcs[r] = lib.smg_rendezvous_files(str(tmp_path).encode(), b"job", r, 2, mine[r], nbytes, out, 20000)
        got[r] = out.raw

    ts = [threading.Thread(target=body, args=(0, 0.0)), threading.Thread(target=body, args=(1, 0.3))]
    for t in ts:
        t.start()
    for t in ts:
        t.join(60)
    assert rcs == [0, 0]
    assert got[0] == mine[0] + mine[1] and got[1] == mine[0] + mine[1]
    assert not (tmp_path / "job.0").exists() and not (tmp_path / "job.1").exists()


def test_mcf_entry_points_need_a_device(problems):
    pr = problems["mcf"]
    s = Solver(device="none").set_hierarchy(pr.P).precompute(pr.A, None)
    with pytest.raises(SmgError) as e:
        s.mcf_setup(pr.F, pr.A)
    assert e.value.status == 5  # SMG_E_STATE: plan-only handles cannot compute
