"""Prints the device timeline of one solve-loop iteration of the 1M workload
(smg_trace_iteration): per kernel launch [start, end] in us, gaps, and per-level sums."""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from surface_multigrid_code_b200 import meshgen as mg  # noqa: E402
from surface_multigrid_code_b200.solver import Solver  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--subdiv", type=int, default=9)
ap.add_argument("--levels", type=int, default=5)
ap.add_argument("--smoother", default="multicolour")
ap.add_argument("--k", type=int, default=1)
ap.add_argument("--workload", default="sphere", choices=["sphere", "hilbert", "bunny", "ogre"])
ap.add_argument("--max-iter", type=int, default=20)
ap.add_argument("--tol", type=float, default=None)
args = ap.parse_args()
if args.workload == "sphere":
    pr = mg.sphere_problem(args.subdiv, args.levels, pad_three=True)
else:
    import bench
    if args.workload == "hilbert" and args.subdiv == 9:
        args.subdiv = 3
    pr = bench.build_problem(args)
s = Solver(smoother=args.smoother, device=0).set_hierarchy(pr.P).precompute(pr.A, pr.known)
s.solve(pr.rhs, pr.z0, pr.known_val, 1e-10, 3)
print("levels:", [(s.level_rows(l), s.level_stats(l)["phases"], s.level_patched(l)) for l in range(pr.nlev)])
ev = s.trace_iteration(args.k)
prev_end = 0.0
per = collections.OrderedDict()
for name, t0, t1 in ev:
    if "." in name.split()[-1] and not name.split()[-1].startswith("g"):  # a stage inside the previous launch
        print(f"{t0:9.2f} {t1:9.2f}  dur {t1 - t0:7.2f}              {name}")
        continue
    print(f"{t0:9.2f} {t1:9.2f}  dur {t1 - t0:7.2f}  gap {t0 - prev_end:6.2f}  {name}")
    key = name.rsplit(" g", 1)[0]
    per.setdefault(key, [0, 0.0])
    per[key][0] += 1
    per[key][1] += t1 - max(t0, prev_end)
    prev_end = max(prev_end, t1)
print("---- exclusive time per (level, kernel) ----")
for kname, (c, t) in per.items():
    print(f"{kname:28s} x{c:3d} {t:8.2f} us")
print(f"total {prev_end:.2f} us")
s.close()
