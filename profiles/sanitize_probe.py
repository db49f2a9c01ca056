"""Small end-to-end exercise of every hot-path kernel for compute-sanitizer (memcheck / racecheck /
synccheck): precompute, both smoothers, patched and phase-by-phase V-cycles with and without PDL,
k = 1 and k = 3 (mean-curvature-flow step with device-side assembly), numeric refresh, the
device-side solve loop, and a 2-rank row partition (ranks as threads of this process).
    compute-sanitizer --tool memcheck python profiles/sanitize_probe.py"""
import os
import sys
import threading

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from surface_multigrid_code_b200 import meshgen as mg  # noqa: E402
from surface_multigrid_code_b200.solver import Solver  # noqa: E402

pr = mg.sphere_problem(4, 3, pad_three=True)
for smoother in ("multicolour", "wavefront"):
    for patch_rows in (0, -1):
        with Solver(smoother=smoother, device=0, patch_rows=patch_rows) as s:
            s.set_hierarchy(pr.P).precompute(pr.A, pr.known)
            z, r, ok = s.solve(pr.rhs, pr.z0, pr.known_val, 1e-10, 30)
            assert ok, (smoother, patch_rows, r)
            u = s.vcycle(0, pr.rhs[s.unknown], np.zeros(s.level_rows(0)))
            print(smoother, patch_rows, "cycles", len(r), "patched", [s.level_patched(l) for l in range(pr.nlev)])
V0, F0 = mg.octahedron()
Vs, Fs, P = mg.subdivision_hierarchy(V0, F0, 4, 3, project_sphere=True, pad_three=True)
Vs = mg.normalize_unit_area(Vs, Fs)
U = np.asfortranarray(Vs * (1.0 + 0.05 * np.random.default_rng(3).standard_normal((Vs.shape[0], 1))))
L0 = mg.cotmatrix(Vs, Fs).tocsc()
L0.sort_indices()
pm = mg.mcf_step_problem(Vs, Fs, P, U=U, L0=L0)
with Solver(device=0) as s:
    s.set_hierarchy(P).precompute(pm.A, None)
    s.mcf_setup(np.ascontiguousarray(Fs, dtype=np.int32), L0, 0.01)
    for _ in range(2):
        U, r, ok = s.mcf_step(U, 5e-7, 20)
        U = np.asfortranarray(U)
    s.update_values(pm.A.data * 1.1)
    s.solve(pm.rhs, pm.z0, None, 5e-7, 20)
    print("mcf ok", ok, len(r))
if os.environ.get("SMG_PROBE_DIST", "1") != "0":
    from dist_util import run_ranks  # ranks = threads sharing the device (host-synchronised exchanges)

    def fn(rank, s):
        s.set_hierarchy(pr.P).precompute(pr.A, pr.known)
        s.barrier()
        return s.solve(pr.rhs, pr.z0, pr.known_val, 1e-10, 40)

    for z, r, ok in run_ranks(2, fn):
        assert ok
    print("dist ok")
print("probe done")
