"""Turns the ncu outputs of profiles/run_profile.sh into the committed summaries.
    python profiles/summarize.py <tag>     (reads gpurun_out/, writes profiles/<tag>_*.{md,json})"""
import collections
import csv
import json
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
OUT, DST = "gpurun_out", "profiles"


def launch_list():
    path = f"{OUT}/launches_{tag}.csv"
    if not os.path.exists(path):
        return None
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void smg::<unnamed>::", "")
        t = float(r[vi].replace(",", "")) / 1e3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    lines = [f"# ncu launch list, {tag}: `ncu --metrics gpu__time_duration.sum --clock-control none` over bench.py",
             "", "Per-launch times are cold-cache and serialised (no PDL overlap): compare SHARES.", "",
             "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append(f"| `{n}` | {c} | {t:.1f} | {100 * t / tot:.1f}% |")
    lines.append(f"| total | {sum(v[0] for v in agg.values())} | {tot:.1f} | 100% |")
    open(f"{DST}/{tag}_launches.md", "w").write("\n".join(lines) + "\n")
    return {n: {"launches": c, "us": t, "share": t / tot} for n, (c, t) in agg.items()}


def full_capture():
    rep = f"{OUT}/prof_{tag}.ncu-rep"
    if not os.path.exists(rep):
        return None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
    idx = [hdr.index(w) for w in want if w in hdr]
    lines = [f"# ncu --set full, {tag}: level-0 kernels of the 1M-vertex workload (profiles/kernel_probe.py)", "",
             "| " + " | ".join(hdr[i].split(".")[0] + f" [{units[i]}]" for i in idx) + " |",
             "|" + "---|" * len(idx)]
    out = []
    for r in rows[2:]:
        rec = {hdr[i]: r[i] for i in idx}
        rec["Kernel Name"] = rec["Kernel Name"].split("(")[0].replace("void smg::<unnamed>::", "")
        out.append(rec)
        lines.append("| " + " | ".join(str(rec[hdr[i]])[:48] for i in idx) + " |")
    open(f"{DST}/{tag}_full.md", "w").write("\n".join(lines) + "\n")
    return out


ll, fc = launch_list(), full_capture()


def to_bytes(text, unit):
    v = float(text.replace(",", ""))
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)


if fc:
    # DRAM traffic of one fine-level Gauss-Seidel sweep = the first four level-0 colour launches
    rep = f"{OUT}/prof_{tag}.ncu-rep"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ir, iw, ik, ig = (hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"),
                      hdr.index("Kernel Name"), hdr.index("Grid Size"))
    gs = [r for r in rows[2:] if "gs_phase" in r[ik]]
    big = max(int(r[ig].strip("()").split(",")[0]) for r in gs)
    lvl0 = [r for r in gs if int(r[ig].strip("()").split(",")[0]) >= big - 1][:4]
    sweep = sum(to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]) for r in lvl0)
    json.dump({"relax_sweep_dram_bytes": sweep, "launches_in_sweep": len(lvl0), "source": f"profiles/{tag}_full.md",
               "note": "dram__bytes_read.sum + dram__bytes_write.sum of the level-0 colour launches of one sweep "
                       "(ncu --set full, cold cache)"}, open(f"{DST}/traffic.json", "w"), indent=1)
for f in (f"bench_{tag}.json", f"timeline_{tag}.txt"):
    if os.path.exists(f"{OUT}/{f}"):
        open(f"{DST}/{tag}_{f.split('_')[0]}" + os.path.splitext(f)[1], "w").write(open(f"{OUT}/{f}").read())
print("launch list:", bool(ll), "full capture:", bool(fc))
