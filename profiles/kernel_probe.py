"""Launches only the level-`lv` hot-path kernels of the 1M-vertex workload (for ncu):
    ncu --set full -k regex:'sell_' ... python profiles/kernel_probe.py [--lv 0] [--reps 2]
Prints the CUDA-event time of each kernel when run without a profiler."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from surface_multigrid_code_b200 import meshgen as mg  # noqa: E402
from surface_multigrid_code_b200.solver import Solver  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--subdiv", type=int, default=9)
ap.add_argument("--levels", type=int, default=5)
ap.add_argument("--lv", type=int, default=0)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--k", type=int, default=1)
ap.add_argument("--smoother", default="multicolour")
ap.add_argument("--kernels", default="relax_sweep,residual,residual_norm,restrict,prolong_add")
args = ap.parse_args()
pr = mg.sphere_problem(args.subdiv, args.levels, pad_three=True)
s = Solver(smoother=args.smoother, device=0).set_hierarchy(pr.P).precompute(pr.A, pr.known)
s.solve(pr.rhs, pr.z0, pr.known_val, 1e-10, 2)  # finite data in the work vectors
print("level stats:", [s.level_stats(l) for l in range(pr.nlev)])
for name in args.kernels.split(","):
    ms, nl = s.time_kernel(name, args.lv, args.k, args.reps, True)
    print(f"{name:14s} lv{args.lv} k{args.k}: {ms*1e3:8.2f} us  ({nl} launches)")
s.close()
