"""Host planning time of smg_precompute (plan-only handle, no CUDA) on the 1M-vertex workload,
single-threaded and with the default thread count:  python profiles/plan_timing.py"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    from surface_multigrid_code_b200 import meshgen as mg
    from surface_multigrid_code_b200.solver import Solver

    pr = mg.sphere_problem(9, 5, pad_three=True)
    best = 1e9
    for _ in range(3):
        s = Solver(device="none").set_hierarchy(pr.P)
        t = time.perf_counter()
        s.precompute(pr.A, pr.known)
        best = min(best, time.perf_counter() - t)
    print(f"threads={os.environ.get('SMG_PLAN_THREADS', 'default')} cores={os.cpu_count()} plan_s={best:.3f}")
else:
    for th in ("1", None):
        env = dict(os.environ)
        if th:
            env["SMG_PLAN_THREADS"] = th
        else:
            env.pop("SMG_PLAN_THREADS", None)
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, check=False)
