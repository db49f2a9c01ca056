"""Opcode histogram per kernel of libsmg.so (cuobjdump -sass) and the ptxas resource report:
evidence that the hot path is sm_100a code using TMA bulk copies (UBLKCP), mbarriers (SYNCS),
programmatic dependent launch (ACQBULK / PDL control), FP64 multiply + add kept separate on the
parity surface (DMUL / DADD, not DFMA), and true division (MUFU.RCP64H sequences).
    python profiles/sass_histogram.py > profiles/r2_sass.md      (no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "surface_multigrid_code_b200", "libsmg.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = m.group(1)
    elif cur and "REG:" in line:
        usage[cur] = dict(re.findall(r"(REG|STACK|SHARED):(\d+)", line))
        cur = None
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
WANT = ["UBLKCP", "UBLKPF", "SYNCS", "ACQBULK", "DFMA", "DMUL", "DADD", "MUFU.RCP64H", "LDG", "STG", "LDS", "STS",
        "BAR", "ATOM", "RED", "MEMBAR", "CCTL"]
hist = collections.OrderedDict()
name = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1)
        hist[name] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and name:
        op = m.group(1)
        hist[name]["_total"] += 1
        for w in WANT:
            if op == w or op.startswith(w + "."):
                hist[name][w] += 1
print(f"# SASS of libsmg.so ({', '.join(arch)}): opcode counts per kernel\n")
print("`cuobjdump -sass` / `-res-usage`; UBLKCP = cp.async.bulk (TMA bulk copy), UBLKPF = bulk L2 prefetch, SYNCS = "
      "mbarrier, DMUL + DADD (not DFMA) = products and sums rounded separately like the reference build.\n")
hot = [n for n in hist if re.search(r"sell_|patch_kernel|dense_sym|halo_exchange|solve_decide|galerkin|mcf_", n)]
print("| kernel | instr | regs | stack | " + " | ".join(WANT) + " |")
print("|---|---:|---:|---:|" + "---:|" * len(WANT))
for n in sorted(hot, key=demangle):
    d = demangle(n)
    d = re.sub(r"\((int|bool)\)", "", d)
    d = re.sub(r"smg::(\(anonymous namespace\)|<unnamed>)::", "", d).split("(")[0].replace("void ", "")
    u = usage.get(n, {})
    print(f"| `{d}` | {hist[n]['_total']} | {u.get('REG', '')} | {u.get('STACK', '')} | " +
          " | ".join(str(hist[n][w]) if hist[n][w] else "" for w in WANT) + " |")
