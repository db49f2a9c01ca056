#!/usr/bin/env python
"""bench.py -- V-cycles/s of the surface-multigrid hot path on B200.

Workload (BASELINE.json configs[2], SURVEY.md section 8d "config 3"): octahedron
subdivided 9x and projected to the sphere (1 048 578 vertices, nnz(A) = 7 340 034),
A = -cotmatrix, b = voronoi mass, 6 pinned vertices, 5-level V(2,2) hierarchy with the
reference's 3-entries-per-row prolongation layout, FP64.

A "step" of the GPU arm is one solve of that problem to tol 1e-10 (16 V-cycles) through the
device-pointer call smg_solve_device: the loop of min_quad_with_fixed_mg_solve
(src/min_quad_with_fixed_mg.cpp:330-347) as ONE graph launch, residual test on the device.
  value : V-cycles/s of those resident solves (inputs and result in HBM); every solve is timed
          with CUDA events on the library's stream, L2 is flushed (256 MB write) between solves,
          and inside a solve every iteration streams ~9x the L2.
  iteration_flushed : round 1's definition of `value`, kept for comparison: one iteration per
          step (residual norm with its host read-back + V(2,2) graph), L2 flushed before each.
  e2e   : V-cycles/s through the public host-buffer call (smg_solve): per solve the RHS
          and z0 are copied from pinned host memory, z and r_his come back.
  roofline : the fine-level Gauss-Seidel sweep (dominant kernel): ONE sweep after an L2 flush,
          algorithmic bytes / CUDA-event time, against MEASURED_PEAKS.json; the L2-assisted pair
          of sweeps and the in-situ timeline beside it.
  cpu_baseline : the reference's own sources (oracle/_ref, compiled unmodified against an Eigen
          stand-in) on ONE host thread, on the same problem; a step there is one iteration.
  --workload bunny | ogre | hilbert | mcf : BASELINE configs 1, 2, 5 and 4.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "V-cycles/sec (1M-vertex sphere Poisson, 5-level V(2,2), FP64)"  # the default workload


def metric_name(pr, args=None):
    if pr.n == 1048578 and pr.nlev == 5 and (args is None or args.workload == "sphere"):
        return METRIC
    wl = args.workload if args is not None else "sphere"
    return f"V-cycles/sec ({pr.n}-vertex {wl} Poisson, {pr.nlev}-level V(2,2), FP64)"

UNIT = "V-cycles/s"
FALLBACK_HBM_GBS = 6650.0


GOLDEN = os.path.join(ROOT, "tests", "golden")


def build_problem(args):
    """The Poisson workloads.  sphere: BASELINE configs[2] (default) and its 4M-vertex sibling;
    hilbert: configs[4], hilbert_cube.obj upsampled --subdiv times (3 = 4 028 672 vertices), 346
    nearest-vertex constraints, 3 subdivision levels + the original mesh + 2 stand-in coarsened
    levels (tests/golden/make_hilbert.py); bunny / ogre: configs[0-1], the reference's meshes with
    03_mg_solver's settings and the stand-in hierarchy of tests/golden/make_golden.py."""
    from surface_multigrid_code_b200 import meshgen as mg

    wl = args.workload
    if wl == "sphere":
        return mg.sphere_problem(args.subdiv, args.levels, tol=1e-10, max_iter=args.max_iter, pad_three=True)
    if wl == "hilbert":
        d = np.load(os.path.join(GOLDEN, "hilbert_cube_base.npz"))
        Pc = [mg.load_csc_keep_zeros(d, "Pc0"), mg.load_csc_keep_zeros(d, "Pc1")]
        return mg.upsampled_mesh_problem("hilbert_cube", d["V"], d["F"], d["known"], Pc, args.subdiv, tol=1e-10,
                                         max_iter=args.max_iter)
    if wl in ("bunny", "ogre"):
        name = {"bunny": "bunny_l3", "ogre": "ogre_l4"}[wl]
        d = np.load(os.path.join(GOLDEN, name + ".npz"))
        n, nlev = int(d["n"]), int(d["nlev"])
        import scipy.sparse as sp

        A = sp.csc_matrix((d["A_data"], d["A_indices"], d["A_indptr"]), shape=(n, n))
        P = [mg.load_csc_keep_zeros(d, f"P{l}") for l in range(nlev - 1)]
        return mg.Problem(wl, A, P, d["known"], d["known_val"], d["rhs"], d["z0"], args.tol or float(d["tol"]),
                          args.max_iter)
    raise SystemExit(f"bench.py: unknown workload {wl}")


def workload_config(pr, args, extra=None):
    cfg = {
        "workload": (f"sphere_subdiv{args.subdiv}_{pr.n}v_{pr.nlev}level_poisson_fp64" if args.workload == "sphere"
                     else f"{args.workload}_{pr.n}v_{pr.nlev}level_poisson_fp64"),
        "vertices": int(pr.n),
        "nnz_A": int(pr.A.nnz),
        "levels": int(pr.nlev),
        "rhs_columns": int(pr.k),
        "pre_post": [2, 2],
        "tol": pr.tol,
        "data_layout": "reference P layout (3 stored entries per row, explicit zeros)",
    }
    if extra:
        cfg.update(extra)
    return cfg


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------
# algorithmic bytes (int32 indices, fp64 values, k RHS columns)
# --------------------------------------------------------------------------------------
def bytes_gs_sweep(n, nnz, k=1):
    return 12 * nnz + 4 * (n + 1) + 8 * n + 8 * n * k + 16 * n * k


def bytes_residual(n, nnz, k=1):
    return 12 * nnz + 4 * (n + 1) + 8 * n * k + 8 * n * k + 8 * n * k


def bytes_residual_norm(n, nnz, k=1):
    return 12 * nnz + 4 * (n + 1) + 8 * n * k + 8 * n * k


def bytes_restrict(nf, nc, pnnz, k=1):
    return 12 * pnnz + 4 * (nc + 1) + 8 * nf * k + 8 * nc * k


def bytes_prolong_add(nf, nc, pnnz, k=1):
    return 12 * pnnz + 8 * nc * k + 16 * nf * k


def iteration_bytes(stats, k=1):
    """Algorithmic bytes of one solve-loop iteration (residual norm + V(2,2))."""
    tot = bytes_residual_norm(stats[0]["rows"], stats[0]["nnz"], k)
    for l in range(len(stats) - 1):
        n, nnz = stats[l]["rows"], stats[l]["nnz"]
        nc, pnnz = stats[l + 1]["rows"], stats[l + 1]["p_nnz"]
        tot += 4 * bytes_gs_sweep(n, nnz, k) + bytes_residual(n, nnz, k)
        tot += bytes_restrict(n, nc, pnnz, k) + bytes_prolong_add(n, nc, pnnz, k)
    nc = stats[-1]["rows"]
    # coarse direct solve: the kernel reads the packed lower-triangular 64x64 tiles of the dense
    # symmetric inverse (half the matrix), b, and reads + writes u
    nblk = (nc + 63) // 64
    tot += 8 * 64 * 64 * (nblk * (nblk + 1) // 2) + 24 * nc * k
    return tot


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [t.strip() for t in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                      "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


# --------------------------------------------------------------------------------------
# CPU arm (oracle): cpu_baseline of the GPU line and the whole --impl reference run
# --------------------------------------------------------------------------------------
def cpu_impl():
    """-> (Oracle impl, cpu_baseline.kind, description).  oracle/_ref/libsmg_ref.so = the
    reference's own mg_VCycle.cpp + min_quad_with_fixed_mg.cpp compiled unmodified against the
    Eigen stand-in of oracle/ref_shim (genuine Eigen is not in this image); built where
    /root/reference is mounted, it travels to the GPU box as a file.  Otherwise the C port."""
    from oracle import cpu_oracle

    if cpu_oracle.ref_available():
        return "ref", "reference", ("the reference's own src/mg_VCycle.cpp + src/min_quad_with_fixed_mg.cpp, "
                                    "compiled unmodified (-O3 -DNDEBUG, no -march, like its Release build) against "
                                    "an Eigen 3.3.7 stand-in (oracle/ref_shim: Eigen is not vendored by the "
                                    "reference and absent from this image; coarse solve = RCM + envelope "
                                    "Cholesky instead of SimplicialLDLT)")
    return "port", "port", "oracle/smg_oracle.c (C restatement of the reference path)"


def cpu_iterations(pr, warmup: int, steps: int, impl: str = "port"):
    """-> seconds per step.  One step = residual norm + V(2,2) on the unknown-sized system,
    single thread (the reference path has no threading)."""
    from oracle.cpu_oracle import Oracle

    ora = Oracle(pr.P, impl=impl).precompute(pr.A, pr.known)
    unknown = ora.unknown
    bu = np.ascontiguousarray(pr.rhs[unknown])
    zu = np.zeros_like(bu)
    for _ in range(warmup):
        zu, _ = ora.iterate(bu, zu, 1)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        zu, _ = ora.iterate(bu, zu, 1)
        times.append(time.perf_counter() - t0)
    return times


def cpu_baseline(pr, nc: int):
    """The `cpu_baseline` object of the GPU line: `nc` solve-loop iterations on one host core."""
    impl, kind, what = cpu_impl()
    times = cpu_iterations(pr, 1, nc, impl)
    cpu = {"value": nc / sum(times), "unit": UNIT, "cores": 1, "kind": kind,
           "sample": f"{nc} solve-loop iterations (residual norm + V(2,2)) of the same problem; {what}; "
                     "single thread (the reference path is single-threaded)",
           "host_cores_available": os.cpu_count()}
    if impl != "port":  # the C restatement beside it
        tp = cpu_iterations(pr, 1, max(nc // 2, 3), "port")
        cpu["port_value"] = len(tp) / sum(tp)
    return cpu


def cpu_parallel(pr, steps: int):
    """Context only, NOT the reference algorithm (which is single-threaded): the same solve-loop
    iteration with OpenMP multicolour Gauss-Seidel and row-parallel products on all host cores
    (oracle/smg_oracle.c::orc_iterate_mt).  Never lets a failure reach the reference line."""
    try:
        from oracle.cpu_oracle import Oracle

        ora = Oracle(pr.P).precompute(pr.A, pr.known)
        bu = np.ascontiguousarray(pr.rhs[ora.unknown])
        want = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        zu, _, threads = ora.iterate_mt(bu, np.zeros_like(bu), 1, want)  # (torchrun exports OMP_NUM_THREADS=1)
        t0 = time.perf_counter()
        zu, _, threads = ora.iterate_mt(bu, zu, steps, want)
        dt = time.perf_counter() - t0
        return {"value": steps / dt, "unit": UNIT, "cores": threads, "kind": "variant",
                "what": "multicolour Gauss-Seidel + row-parallel products with OpenMP on all host cores; "
                        "not the reference algorithm (lexicographic Gauss-Seidel, one thread)"}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": str(e)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pr = build_problem(args)
    impl, kind, what = cpu_impl()
    times = cpu_iterations(pr, args.warmup, args.steps, impl)
    total = sum(times)
    v = args.steps / total
    line = {
        "impl": "reference",
        "metric": metric_name(pr, args), "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(pr, args, {"parallelism": "host cpu, 1 thread"}),
        "cpu_baseline": {
            "value": v, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"{args.steps} solve-loop iterations (residual norm + V(2,2)) of the same problem; "
                      f"{what}; single thread: the reference path has no threading (mg_VCycle.cpp)",
        },
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cores_available": os.cpu_count(),
        "cpu_parallel": cpu_parallel(pr, max(3, min(args.steps, 10))),
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def run_gpu(args):
    import torch

    from surface_multigrid_code_b200.solver import Solver

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    pr = build_problem(args)
    s = Solver(smoother=args.smoother, device=local_rank, use_graph=not args.no_graph)
    # N > 1: ONE problem, its fine levels partitioned by rows over the N GPUs (halo exchange
    # through peer-mapped memory inside the V-cycle graph); --replicas: N independent problems
    partitioned = world > 1 and not args.replicas
    if partitioned:
        # staging: every rank receives up to n/world rows x 4 columns at 16 bytes per value from
        # every peer, double-buffered
        comm_mb = args.comm_mb or max(256, (pr.n * 4 * 16 * 2 * 5 // 4 >> 20) + 16)
        s.dist_init(rank, world, comm_mb << 20)
        s.dist_options(args.halo, args.dist_levels, args.dist_min_rows)
        s.dist_connect_torch()
    t0 = time.perf_counter()
    s.set_hierarchy(pr.P).precompute(pr.A, pr.known)
    t_pre = time.perf_counter() - t0
    stats = [s.level_stats(l) for l in range(pr.nlev)]
    n, k = pr.n, pr.k
    part_info = None
    own_frac = 1.0
    if partitioned:
        li = [s.dist_level_info(l) for l in range(pr.nlev)]
        own_frac = li[0]["own_rows"] / max(stats[0]["rows"], 1)
        part_info = {"partitioned_levels": s.dist_info()["dist_levels"],
                     "halo": {0: "hybrid (exchange per sweep)", 1: "exact (exchange per colour)",
                              2: "hybrid (exchange per relax call)"}[args.halo],
                     "rank0_rows_per_level": [x["own_rows"] for x in li],
                     "rank0_halo_u_rows_per_level": [x["halo_u_recv"] for x in li],
                     "transport": "peer-mapped device memory (CUDA IPC), stores over NVLink + epoch flags; "
                                  "no NCCL call on the data path"}

    # pinned host buffers for the end-to-end call
    h_rhs = torch.from_numpy(np.ascontiguousarray(pr.rhs)).pin_memory()
    h_z0 = torch.from_numpy(np.ascontiguousarray(pr.z0)).pin_memory()
    h_kv = torch.from_numpy(np.ascontiguousarray(pr.known_val)).pin_memory()
    h_z = torch.empty(n * k, dtype=torch.float64).pin_memory()
    flush = torch.empty(1 << 25, dtype=torch.float64, device="cuda")  # 256 MB > L2

    def solve_e2e():
        return s.solve_host_ptr(h_rhs.data_ptr(), h_kv.data_ptr(), h_z0.data_ptr(), h_z.data_ptr(),
                                k, pr.tol, pr.max_iter)

    # correctness gate + warm-up of the whole path (graph capture, allocations)
    r_his, ok = solve_e2e()
    if not ok:
        raise SystemExit(f"bench.py: solve did not converge: {r_his}")
    cycles_per_solve = len(r_his) - 1
    z = h_z.numpy()
    A = pr.A.tocsr()
    true_res = float(np.linalg.norm((pr.rhs - A @ z)[s.unknown]))
    if not true_res < 10 * pr.tol:
        raise SystemExit(f"bench.py: returned z does not satisfy the system: {true_res}")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: resident solve-loop iterations ---------------------------------------
    s.time_kernel("mg_iteration", 0, k, max(args.warmup, 3), True)
    clocks = ClockSampler(local_rank)
    clocks.start()
    time.sleep(1.5)  # nvidia-smi needs about a second before its first sample
    barrier()
    l0 = s.launch_count
    x0 = s.dist_info()["exchanges"] if partitioned else 0
    ms_iter, _ = s.time_kernel("mg_iteration", 0, k, args.steps, True)
    launches = s.launch_count - l0
    barrier()
    if part_info is not None:
        # (exchanges inside a replayed CUDA graph are counted once per capture: see gpu_launches)
        part_info["exchange_kernels_counted"] = s.dist_info()["exchanges"] - x0
    total_ms = ms_iter * args.steps

    # ---- e2e: host-buffer solves ---------------------------------------------------------
    ext = torch.cuda.ExternalStream(s.stream)
    e2e_ms, e2e_cycles = 0.0, 0
    n_solves = max(3, min(10, args.steps))
    with torch.cuda.stream(ext):
        for i in range(-2, n_solves):
            flush.fill_(1.0)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            r_his, ok = solve_e2e()
            ev1.record()
            ev1.synchronize()
            if i >= 0:
                e2e_ms += ev0.elapsed_time(ev1)
                e2e_cycles += len(r_his) - 1
    barrier()

    # ---- value: resident solves through the device-pointer call (smg_solve_device): the solve loop
    # as it runs in production (one graph launch per solve, residual test on the device, L2 state
    # carried from iteration to iteration: an iteration streams ~1.1 GB, nine times the L2); L2
    # flushed between solves; every solve timed with CUDA events on the library's stream.
    d_rhs = torch.from_numpy(np.ascontiguousarray(pr.rhs)).cuda()
    d_z0 = torch.from_numpy(np.ascontiguousarray(pr.z0)).cuda()
    d_kv = torch.from_numpy(np.ascontiguousarray(pr.known_val)).cuda()
    d_z = torch.empty(n * k, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    res_ms, res_cycles, res_launches = 0.0, 0, 0
    barrier()
    with torch.cuda.stream(ext):
        for i in range(-max(args.warmup, 3), args.steps):
            flush.fill_(1.0)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            lc0 = s.launch_count
            ev0.record()
            r_res, ok_res = s.solve_device(d_rhs.data_ptr(), d_kv.data_ptr(), d_z0.data_ptr(), d_z.data_ptr(), k,
                                           pr.tol, pr.max_iter)
            ev1.record()
            ev1.synchronize()
            if i >= 0:
                res_ms += ev0.elapsed_time(ev1)
                res_cycles += len(r_res) - 1
                res_launches += s.launch_count - lc0
    barrier()

    # ---- per-kernel roofline numbers (rank-local, level 0) ----------------------------------
    peak, peak_src = measured_peak()
    reps = 20
    kern = {}
    n0, nnz0 = stats[0]["rows"], stats[0]["nnz"]
    n1, p1 = stats[1]["rows"], stats[1]["p_nnz"]
    # partitioned: a rank streams its own rows only (time includes its halo exchanges)
    n0, nnz0 = int(n0 * own_frac), int(nnz0 * own_frac)
    for name, nbytes in (("relax_sweep", bytes_gs_sweep(n0, nnz0, k)),
                         ("residual", bytes_residual(n0, nnz0, k)),
                         ("residual_norm", bytes_residual_norm(n0, nnz0, k)),
                         ("restrict", bytes_restrict(n0, int(n1 * own_frac), int(p1 * own_frac), k)),
                         ("prolong_add", bytes_prolong_add(n0, n1, int(p1 * own_frac), k))):
        ms, nl = s.time_kernel(name, 0, k, reps, True)
        kern[name] = {"ms": ms, "launches": nl, "algorithmic_bytes": nbytes,
                      "gbs": nbytes / (ms * 1e-3) / 1e9, "frac": nbytes / (ms * 1e-3) / 1e9 / peak}
    ms_v, nl_v = s.time_kernel("vcycle", 0, k, reps, True)
    # the smoother as it runs inside a V-cycle: the two pre-smoothing sweeps back to back
    # (L2 flushed before the pair, not between), CUDA events on the library's stream
    ms_pre, nl_pre = s.time_kernel("relax_pre", 0, k, reps, True)
    # device timeline of one iteration (%globaltimer per kernel, PDL overlap resolved)
    in_situ = {}
    try:
        prev_end, acc = 0.0, {}
        for name, t0, t1 in s.trace_iteration(k):
            key = name.rsplit(" g", 1)[0]
            acc[key] = acc.get(key, 0.0) + (t1 - max(t0, prev_end))
            prev_end = max(prev_end, t1)
        gs_us = acc.get("L0 down gs_phase", 0.0) + acc.get("L0 up gs_phase", 0.0)
        in_situ = {"iteration_us": prev_end, "exclusive_us": acc,
                   "l0_gs_sweep_us": gs_us / 4.0,
                   "l0_gs_gbs": 4.0 * bytes_gs_sweep(n0, nnz0, k) / (gs_us * 1e-6) / 1e9 if gs_us else None,
                   "l0_residual_gbs": bytes_residual(n0, nnz0, k) / (acc["L0 down residual"] * 1e-6) / 1e9
                   if acc.get("L0 down residual") else None}
        if in_situ["l0_gs_gbs"]:
            in_situ["l0_gs_frac"] = in_situ["l0_gs_gbs"] / peak
    except Exception as e:  # profiling aid only
        in_situ = {"error": str(e)}
    clk = clocks.stop()
    traffic, traffic_pair, traffic_note = None, None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tj = json.load(fh)
        traffic, traffic_note = tj.get("relax_sweep_dram_bytes"), tj.get("note")
        traffic_pair = tj.get("relax_pair_dram_bytes_per_sweep")
    except Exception:
        pass
    gs = kern["relax_sweep"]
    pre_sweeps = max(nl_pre // max(gs["launches"], 1), 1)
    pre_gbs = pre_sweeps * gs["algorithmic_bytes"] / (ms_pre * 1e-3) / 1e9
    # `frac` is the HBM-honest figure: ONE sweep with L2 flushed before it (its DRAM traffic equals
    # its algorithmic bytes, profiles/).  The pair of pre-smoothing sweeps as a V-cycle runs them
    # and the in-situ device timeline are L2-assisted (the second sweep finds part of the 88 MB
    # matrix in the 126 MB L2) and are reported beside it, not as the roofline fraction.
    roofline = {
        "bound": "hbm", "kernel": "fine-level Gauss-Seidel sweep "
                                  f"({gs['launches']} colour launches of sell_gs_phase_multi_kernel per sweep)",
        "achieved": gs["gbs"], "peak": peak, "unit": "GB/s", "frac": gs["frac"], "traffic": traffic,
        "traffic_note": traffic_note,
        "peak_source": peak_src, "algorithmic_bytes_per_sweep": gs["algorithmic_bytes"],
        "ms_per_sweep": gs["ms"],
        "how": "one sweep (all colour launches), L2 flushed before every sweep (256 MB write), CUDA events on the "
               "library stream, mean of 20; algorithmic bytes = 12 nnz + 4 (n + 1) + 8 n + 24 n k (SURVEY.md 8d)",
        "warm_pair": {"achieved": pre_gbs, "frac": pre_gbs / peak, "ms_per_sweep": ms_pre / pre_sweeps,
                      "traffic_per_sweep": traffic_pair,
                      "how": f"{pre_sweeps} pre-smoothing sweeps back to back as inside a V-cycle, L2 flushed before "
                             "each pair only: L2-assisted, NOT an HBM roofline fraction"},
        "in_situ": in_situ,
        "iteration": {"algorithmic_bytes": iteration_bytes(stats, k), "ms": ms_iter,
                      "gbs": iteration_bytes(stats, k) / (ms_iter * 1e-3) / 1e9,
                      "frac": iteration_bytes(stats, k) / (ms_iter * 1e-3) / 1e9 / peak},
        "kernels": kern, "vcycle_ms": ms_v, "vcycle_launches": nl_v,
    }

    # ---- max over ranks -----------------------------------------------------------------------
    t = torch.tensor([total_ms, e2e_ms, res_ms], dtype=torch.float64, device="cuda")
    c = torch.tensor([float(e2e_cycles), float(res_cycles)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    total_ms_max, e2e_ms_max, res_ms_max = float(t[0]), float(t[1]), float(t[2])
    # replicas: every GPU runs its own problem; partitioned: the N GPUs share one
    jobs = 1 if partitioned else world
    # value: V-cycles/s of resident solves (a step = one smg_solve_device call); the round-1
    # definition (one iteration per step, L2 flushed before each, residual read back by the host)
    # is kept beside it as `iteration_flushed`
    value = float(c[1]) / world * jobs / (res_ms_max * 1e-3)
    iteration_flushed = {"value": jobs * args.steps / (total_ms_max * 1e-3), "unit": UNIT,
                         "ms_per_iteration": total_ms_max / args.steps, "iterations": args.steps,
                         "what": "one solve-loop iteration per step (residual norm with its host read-back + V(2,2) "
                                 "graph), L2 flushed (256 MB write) before every iteration: the `value` of round 1"}
    e2e_value = float(c[0]) / world * jobs / (e2e_ms_max * 1e-3)

    # the other way to use N GPUs: N independent problems, one per GPU (no exchange at all)
    replicas = None
    if partitioned and not args.no_replicas:
        barrier()
        with Solver(smoother=args.smoother, device=local_rank, use_graph=not args.no_graph) as s1:
            s1.set_hierarchy(pr.P).precompute(pr.A, pr.known)
            s1.time_kernel("mg_iteration", 0, k, 3, True)
            barrier()
            ms1, _ = s1.time_kernel("mg_iteration", 0, k, args.steps, True)
        t1 = torch.tensor([ms1], dtype=torch.float64, device="cuda")
        dist.all_reduce(t1, op=dist.ReduceOp.MAX)
        replicas = {"value": world * 1e3 / float(t1[0]), "unit": UNIT, "scaling": "weak",
                    "ms_per_step": float(t1[0]),
                    "what": f"{world} independent copies of the same problem, one per GPU, same timing rules"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(pr, args.cpu_steps)
    barrier()
    if rank == 0:
        line = {
            "metric": metric_name(pr, args), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": res_ms_max / args.steps,
            "higher_is_better": True, "scaling": "strong" if partitioned else "weak", "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(pr, args, {
                "parallelism": "single GPU" if world == 1 else (
                    f"one problem, fine levels row-partitioned over {world} GPUs, coarse levels replicated"
                    if partitioned else f"{world} independent replicas (one problem per GPU)"),
                "partition": part_info,
                "smoother": args.smoother, "cuda_graph": not args.no_graph,
                "step": f"one smg_solve_device call = {cycles_per_solve} V-cycles to tol {pr.tol} on device-resident buffers",
                "l2": "flushed between timed steps (256 MB write); inside a step every iteration streams ~9x the L2",
                "phases_per_level": [st["phases"] for st in stats],
                "rows_per_level": [st["rows"] for st in stats],
                "precompute_s": t_pre}),
            "e2e": {"value": e2e_value, "unit": UNIT,
                    # (partitioned: every rank copies 1/N of the rows of RHS / z0, NVLink completes them)
                    "h2d_bytes_per_step": int(8 * 2 * n * k) * (1 if partitioned else world) + int(8 * h_kv.numel()) * world,
                    "d2h_bytes_per_step": int(8 * n * k + 8 * (cycles_per_solve + 1)) * world,
                    "step": f"one smg_solve call = {cycles_per_solve} V-cycles to tol {pr.tol}",
                    "ms_per_solve": e2e_ms_max / n_solves, "solves": n_solves},
            "gpu_launches": int(res_launches),
            "iteration_flushed": iteration_flushed,
            "clocks": clk,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "replicas": replicas,
            "final_residual": float(r_his[-1]), "true_residual": true_res,
        }
        print(json.dumps(line), flush=True)
    barrier()  # nobody frees its comm buffer while a peer may still use it
    s.close()
    if dist is not None:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------
# BASELINE configs[3]: mean-curvature flow (05_example_mean_curvature_flow/main.cpp:57-79)
# --------------------------------------------------------------------------------------
def build_mcf(args):
    """The 1M-vertex sphere with a seeded radial perturbation (something to flow), its
    subdivision hierarchy, the cotangent matrix of the rest shape (main.cpp:41, computed once)."""
    from surface_multigrid_code_b200 import meshgen as mg

    V0, F0 = mg.octahedron()
    V, F, P = mg.subdivision_hierarchy(V0, F0, args.subdiv, args.levels, project_sphere=True, pad_three=True)
    rng = np.random.default_rng(0)
    V = V * (1.0 + 0.02 * rng.standard_normal((V.shape[0], 1)))
    V = mg.normalize_unit_area(V, F)
    L0 = mg.cotmatrix(V, F).tocsc()
    L0.sort_indices()
    return V, np.ascontiguousarray(F, dtype=np.int32), P, L0


def run_mcf(args):
    """One line for configs[3].  A step = one flow step: M = massmatrix(U), LHS = M - 0.01 L,
    RHS = M U, min_quad_with_fixed_mg_precompute(LHS), min_quad_with_fixed_mg_solve(RHS, U, 5e-7),
    normalize_unit_area.  GPU arm: smg_mcf_step (assembly, Galerkin refresh and coarse
    factorisation on the device; only U crosses the bus); value = V-cycles/s over the solves,
    per-step precompute and solve times beside it.  --impl reference: the same steps through the
    reference's own sources on one host thread (assembly in numpy, not timed)."""
    from surface_multigrid_code_b200 import meshgen as mg

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    delta, tol, max_iter = 0.01, 5e-7, 20
    V, F, P, L0 = build_mcf(args)
    n = V.shape[0]
    cfg = {"workload": f"mcf_sphere_subdiv{args.subdiv}_{n}v_{args.levels}level_k3_fp64", "vertices": int(n),
           "nnz_A": int(L0.nnz), "levels": args.levels, "rhs_columns": 3, "pre_post": [2, 2], "tol": tol,
           "delta": delta, "flow_steps": args.flow_steps,
           "data_layout": "reference P layout (3 stored entries per row, explicit zeros)"}
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle.cpu_oracle import Oracle

        impl, kind, what = cpu_impl()
        U = np.asfortranarray(V.copy())
        steps = max(1, min(args.flow_steps, args.steps if args.steps < 10 else 2))
        t_pre = t_solve = 0.0
        cycles = 0
        ora = Oracle(P, impl=impl)
        for _ in range(steps):
            pr = mg.mcf_step_problem(V, F, P, U=U, delta=delta, tol=tol, max_iter=max_iter, L0=L0)
            t0 = time.perf_counter()
            ora.precompute(pr.A, None)
            t1 = time.perf_counter()
            z, r_his, ok = ora.solve(pr.rhs, pr.z0, None, tol, max_iter)
            t2 = time.perf_counter()
            t_pre += t1 - t0
            t_solve += t2 - t1
            cycles += len(r_his) - 1
            U = np.asfortranarray(mg.normalize_unit_area(z, F))
        v = cycles / t_solve
        line = {"impl": "reference", "metric": f"V-cycles/sec (mean-curvature flow, {n} vertices, k = 3, FP64)",
                "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": 0,
                "ms_per_step": 1e3 * (t_pre + t_solve) / steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": dict(cfg, parallelism="host cpu, 1 thread"),
                "mcf": {"flow_steps_timed": steps, "precompute_ms_per_step": 1e3 * t_pre / steps,
                        "solve_ms_per_step": 1e3 * t_solve / steps, "vcycles_per_step": cycles / steps},
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": kind,
                                 "sample": f"{steps} flow steps (precompute + solve to {tol}); {what}"},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    import torch

    from surface_multigrid_code_b200.solver import Solver

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    s = Solver(smoother=args.smoother, device=local_rank, use_graph=not args.no_graph)
    if world > 1:
        s.dist_init(rank, world, (args.comm_mb or max(256, (n * 4 * 16 * 2 * 5 // 4 >> 20) + 16)) << 20)
        s.dist_options(args.halo, args.dist_levels, args.dist_min_rows)
        s.dist_connect_torch()
    pr0 = mg.mcf_step_problem(V, F, P, U=V, delta=delta, tol=tol, max_iter=max_iter, L0=L0)
    t0 = time.perf_counter()
    s.set_hierarchy(P).precompute(pr0.A, None)
    s.mcf_setup(F, L0, delta)
    t_setup = time.perf_counter() - t0

    def flow(nsteps, U):
        pre = sol = copy = 0.0
        cyc = 0
        wall0 = time.perf_counter()
        for _ in range(nsteps):
            z, r_his, ok = s.mcf_step(U, tol, max_iter)
            tm = s.timings()
            pre += tm["precompute_device_ms"]
            sol += tm["solve_ms"]
            copy += tm["d2h_ms"]
            cyc += len(r_his) - 1
            U = np.asfortranarray(mg.normalize_unit_area(z, F))
        return U, pre, sol, copy, cyc, time.perf_counter() - wall0

    U = np.asfortranarray(V.copy())
    flow(max(1, min(args.warmup, 2)), U)  # graph capture, allocations
    clocks = ClockSampler(local_rank)
    clocks.start()
    time.sleep(1.5)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = s.launch_count
    U, pre, sol, copy, cyc, wall = flow(args.flow_steps, U)
    launches = s.launch_count - l0
    torch.cuda.synchronize()
    clk = clocks.stop()
    t = torch.tensor([pre, sol, copy, wall], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    pre, sol, copy, wall = (float(x) for x in t)
    assert np.all(np.isfinite(U))
    if rank == 0:
        nst = args.flow_steps
        line = {"metric": f"V-cycles/sec (mean-curvature flow, {n} vertices, k = 3, FP64)",
                "value": cyc / (sol * 1e-3), "unit": UNIT, "n_gpus": world, "steps": nst,
                "warmup": max(1, min(args.warmup, 2)), "ms_per_step": (pre + sol + copy) / nst,
                "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": dict(cfg, parallelism="single GPU" if world == 1 else
                               f"fine levels row-partitioned over {world} GPUs", smoother=args.smoother,
                               setup_s=t_setup),
                "mcf": {"flow_steps_timed": nst, "precompute_ms_per_step": pre / nst, "solve_ms_per_step": sol / nst,
                        "d2h_ms_per_step": copy / nst, "vcycles_per_step": cyc / nst,
                        "wall_ms_per_step_incl_host_normalisation": 1e3 * wall / nst,
                        "what": "precompute = device-side assembly of M - delta L and M U, Galerkin products of all "
                                "levels, diagonals, dense coarse factorisation (smg_mcf_step, H2D of U included); "
                                "solve = the solve loop on the device (k = 3)"},
                "e2e": {"value": cyc / (1e-3 * (pre + sol + copy)), "unit": UNIT,
                        "h2d_bytes_per_step": 24 * n * world, "d2h_bytes_per_step": 24 * n * world,
                        "step": "one smg_mcf_step call with host U in / out (assembly + precompute + solve)"},
                "gpu_launches": int(launches), "clocks": clk, "roofline": None, "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
    s.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sphere", choices=["sphere", "hilbert", "bunny", "ogre", "mcf"],
                    help="sphere: BASELINE configs[2] (default); hilbert: configs[4]; bunny / ogre: configs[0-1]; "
                         "mcf: configs[3] (10 mean-curvature-flow steps, k = 3)")
    ap.add_argument("--subdiv", type=int, default=None,
                    help="sphere / mcf: octahedron subdivisions (default 9 = 1M vertices); hilbert: upsampling "
                         "steps (default 3 = 4M vertices)")
    ap.add_argument("--levels", type=int, default=5)
    ap.add_argument("--max-iter", type=int, default=None,
                    help="maxIter of the solve (reference default 20; the 4M meshes need more for 1e-10: default 40)")
    ap.add_argument("--tol", type=float, default=None, help="bunny / ogre: tolerance (default: the fixture's 1e-3)")
    ap.add_argument("--flow-steps", type=int, default=10, help="mcf: flow steps (05_example: one per key press)")
    ap.add_argument("--smoother", default="multicolour", choices=["multicolour", "wavefront"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-steps", type=int, default=20)
    ap.add_argument("--no-replicas", action="store_true", help="skip the extra replica-mode measurement")
    ap.add_argument("--replicas", action="store_true",
                    help="N > 1: N independent problems instead of one row-partitioned problem")
    ap.add_argument("--halo", type=int, default=2, choices=[0, 1, 2],
                    help="halo exchange per sweep (0), per colour (1, = single-GPU smoother), per relax call (2)")
    ap.add_argument("--dist-levels", type=int, default=-1)
    ap.add_argument("--dist-min-rows", type=int, default=0)
    ap.add_argument("--comm-mb", type=int, default=0, help="peer-mapped staging buffer per GPU (0 = auto)")
    args = ap.parse_args()
    if args.subdiv is None:
        args.subdiv = 3 if args.workload == "hilbert" else 9
    if args.max_iter is None:
        args.max_iter = 40 if args.workload == "hilbert" else 20
    if args.workload == "mcf":
        run_mcf(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
