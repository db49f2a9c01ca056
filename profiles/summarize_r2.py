"""Turns the outputs of profiles/run_profile_r2.sh / run_profile_r2b.sh (gpurun_out/) into the committed
round-2 summaries: profiles/r2_full.md, r2_traffic_nocc.md, traffic.json, r2_sanitizer.md, r2_bench.json,
r2_timeline.txt.  (The launch list is written by profiles/summarize.py r2.)"""
import csv
import json
import os

OUT, DST = "gpurun_out", "profiles"


def short(name):
    return name.split("(")[0].replace("void smg::<unnamed>::", "").replace("void unnamed>::", "")


def full(files):
    want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
    lines = ["# ncu --set full, r2 (`--clock-control none`; SMG_HOST_LOOP=1: ncu cannot profile kernel nodes of a graph "
             "with conditional nodes)", "",
             "Level-0 kernels of the 1M-vertex workload as `profiles/kernel_probe.py` launches them (L2 flushed before "
             "each timed launch group), then the patch / coarse / level-1 kernels of the first V-cycle of its solve. "
             "Per-launch times under ncu are serialised (no PDL overlap) and cold.", ""]
    for title, path in files:
        if not os.path.exists(path):
            continue
        rows = list(csv.reader(open(path)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        idx = [hdr.index(w) for w in want if w in hdr]
        lines += [f"## {title}", "", "| " + " | ".join(hdr[i].split(".")[0].replace("smsp__average_warps_issue_", "") +
                                                        f" [{units[i]}]" for i in idx) + " |", "|" + "---|" * len(idx)]
        for r in rows[2:]:
            cells = []
            for i in idx:
                v = short(r[i]) if hdr[i] == "Kernel Name" else r[i]
                try:
                    v = f"{float(v.replace(',', '')):.2f}" if "." in v else v
                except ValueError:
                    pass
                cells.append(v[:44])
            lines.append("| " + " | ".join(cells) + " |")
        lines.append("")
    open(f"{DST}/r2_full.md", "w").write("\n".join(lines) + "\n")


def traffic():
    path = f"{OUT}/traffic_nocc_r2.csv"
    if not os.path.exists(path):
        return
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    iid, im, iv, ig = hdr.index("ID"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    per = {}
    for r in rows[1:]:
        per.setdefault(int(r[iid]), {"grid": r[ig]})[r[im]] = float(r[iv].replace(",", ""))
    ids = sorted(per)
    b = lambda i: per[i]["dram__bytes_read.sum"] + per[i]["dram__bytes_write.sum"]
    # capture order (profiles/kernel_probe.py --kernels relax_sweep,relax_pre, one warm-up + one timed rep each, our
    # own 256 MB L2 flush before every rep): [0-3] sweep warm-up, [4-7] timed sweep, [8-15] pair warm-up, [16-23] pair
    cold = sum(b(i) for i in ids[4:8])
    pair = sum(b(i) for i in ids[16:24])
    second = sum(b(i) for i in ids[20:24])
    lines = ["# DRAM traffic of the fine-level Gauss-Seidel sweep, r2 (`ncu --cache-control none --clock-control none`)", "",
             "Without ncu's own cache flush between kernels, so the colour launches of a sweep share L2 as in production "
             "(each also prefetches the next colour's matrix chunk into L2); the library flushes L2 (256 MB write) "
             "before every timed group.", "",
             "| launch | grid | dram read [MB] | dram write [MB] | time under ncu [us] |", "|---|---|---:|---:|---:|"]
    names = (["sweep warm-up"] * 4 + ["COLD SWEEP"] * 4 + ["pair warm-up"] * 8 + ["pair: first sweep"] * 4 +
             ["pair: second sweep"] * 4)
    for n, i in zip(names, ids):
        lines.append(f"| {n} | {per[i]['grid']} | {per[i]['dram__bytes_read.sum'] / 1e6:.2f} | "
                     f"{per[i]['dram__bytes_write.sum'] / 1e6:.2f} | {per[i]['gpu__time_duration.sum'] / 1e3:.2f} |")
    lines += ["", f"cold single sweep: **{cold / 1e6:.1f} MB** (algorithmic 125.8 MB); pair of sweeps: {pair / 1e6:.1f} MB = "
                  f"{pair / 2e6:.1f} MB per sweep (second sweep alone {second / 1e6:.1f} MB: part of the 88 MB matrix is "
                  "still in the 126 MB L2)."]
    open(f"{DST}/r2_traffic_nocc.md", "w").write("\n".join(lines) + "\n")
    json.dump({"relax_sweep_dram_bytes": cold, "relax_pair_dram_bytes_per_sweep": pair / 2,
               "second_sweep_dram_bytes": second, "launches_in_sweep": 4, "source": "profiles/r2_traffic_nocc.md",
               "note": "dram__bytes_read.sum + dram__bytes_write.sum with ncu --cache-control none: one sweep after an "
                       "L2 flush (cold), and per sweep of two sweeps back to back (the second finds part of the matrix "
                       "in L2)"}, open(f"{DST}/traffic.json", "w"), indent=1)


def sanitizer():
    lines = ["# compute-sanitizer, r2 (profiles/sanitize_probe.py: both smoothers, patched and phase-by-phase V-cycles, "
             "k = 1 and k = 3, mean-curvature-flow step, numeric refresh, device-side solve loop, 2-rank partition)", ""]
    for tool in ("memcheck", "racecheck", "racecheck_nopdl", "synccheck"):
        p = f"{OUT}/sanitizer_{tool}_r2.log"
        if os.path.exists(p):
            tail = [l.rstrip() for l in open(p).read().splitlines() if l.strip()][-4:]
            lines += [f"## {tool}", "", "```"] + tail + ["```", ""]
    open(f"{DST}/r2_sanitizer.md", "w").write("\n".join(lines) + "\n")


full([("level 0 (1 048 572 rows)", f"{OUT}/prof_l0_r2.csv"),
      ("first V-cycle: level-1 phase kernels, patch launches (levels 2-3), coarse solve", f"{OUT}/prof_small_r2.csv")])
traffic()
sanitizer()
for src, dst in (("bench_r2.json", "r2_bench.json"), ("timeline_r2.txt", "r2_timeline.txt")):
    if os.path.exists(f"{OUT}/{src}"):
        open(f"{DST}/{dst}", "w").write(open(f"{OUT}/{src}").read())
print("done")
