"""CPU tests of the communication-avoiding patch schedule (csrc/patch.hpp): the planner lays a
whole relax call of a level out in patches with a shrinking halo, and smg_patch_plan proves
symbolically (verify_patches, integer work only) that every row update reads its neighbours at
exactly the version the phase-by-phase multicolour schedule would, that rows and coarse rows
are owned exactly once, and that entry lists / value sources equal the SELL rows."""
import numpy as np
import pytest

from surface_multigrid_code_b200 import meshgen as mg
from surface_multigrid_code_b200.solver import SmgError, Solver


@pytest.fixture(scope="module")
def plan_only():
    pr = mg.sphere_problem(5, 4, pad_three=True)
    s = Solver(device="none").set_hierarchy(pr.P).precompute(pr.A, pr.known)
    yield pr, s
    s.close()


@pytest.mark.parametrize("kind", ["down", "up"])
@pytest.mark.parametrize("iters", [0, 1, 2, 3])
@pytest.mark.parametrize("target", [32, 100, 5000])
def test_patch_schedule_is_equivalent_to_the_phase_schedule(plan_only, kind, iters, target):
    pr, s = plan_only
    for lv in range(pr.nlev - 1):
        st = s.patch_plan(lv, kind, iters, target, verify=True)
        n = s.level_rows(lv)
        assert st["owned"] == n
        assert st["patches"] >= max(1, n // max(target, 1))
        assert st["local"] >= n
        colours = s.level_stats(lv)["phases"]
        # every owned row is updated iters times; halo rows add redundant updates
        assert st["updates"] >= iters * n
        if st["patches"] == 1:  # no halo, no redundancy
            assert st["local"] == n and st["updates"] == iters * n
        if target >= n and n < 1000:
            assert st["patches"] == 1
        assert colours >= 1


def test_patch_plan_on_an_irregular_mesh_and_free_variant(problems):
    for name in ("grid", "mcf"):
        pr = problems[name]
        s = Solver(device="none").set_hierarchy(pr.P).precompute(pr.A, pr.known)
        for lv in range(pr.nlev - 1):
            for kind in ("down", "up"):
                st = s.patch_plan(lv, kind, 2, 64, verify=True)
                assert st["owned"] == s.level_rows(lv)
        s.close()


def test_patch_plan_respects_the_shared_memory_limit(plan_only):
    pr, s = plan_only
    big = s.patch_plan(0, "down", 2, 5000, verify=False)
    lim = big["max_blob_bytes"] // 3
    small = s.patch_plan(0, "down", 2, 5000, smem_limit=lim, verify=True)
    assert small["patches"] > big["patches"]
    assert small["max_blob_bytes"] + 8 * small["max_vec"] <= lim
    with pytest.raises(SmgError):
        s.patch_plan(0, "down", 2, 64, smem_limit=600)  # not even one row fits
    with pytest.raises(SmgError):
        s.patch_plan(pr.nlev - 1, "down", 2, 64)  # the coarsest level has no coarser one


def test_wavefront_schedules_are_not_patched(plan_only):
    pr, _ = plan_only
    s = Solver(smoother="wavefront", device="none").set_hierarchy(pr.P).precompute(pr.A, pr.known)
    # dozens of wavefront levels x sweeps exceed the phase budget of a patch
    with pytest.raises(SmgError):
        s.patch_plan(0, "down", 2, 128)
    s.close()
