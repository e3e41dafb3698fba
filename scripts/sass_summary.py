#!/usr/bin/env python
"""Per-kernel SASS instruction counts of libpvd_b200.so (cuobjdump -sass; no GPU needed): which kernels use the tcgen05 tensor
cores (UTCHMMA / UTCQMMA), TMEM loads (LDTM) / stores (STTM), tcgen05 barriers (UTCBAR), TMA bulk copies (UBLKCP) or tensor-map TMA
(UTMALDG / UTMASTG), vector reductions (REDG / RED), multimem loads / stores through the NVSwitch (LDGMC / STGMC), setmaxnreg (USETMAXREG), plus registers
and shared memory per kernel (cuobjdump -res-usage).  Writes profiles/sass_summary.txt.

    python scripts/sass_summary.py [path/to/lib.so] [out.txt]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "aaai2023-pvd_b200", "pvd_b200", "libpvd_b200.so")
OUT = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "sass_summary.txt")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "REDG", "RED", "ATOMG", "LDGMC", "STGMC", "USETMAXREG", "HMMA", "LDG", "STG", "SHFL"]


def demangle(names):
    r = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True)
    return r.stdout.splitlines() if r.returncode == 0 else names


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*SHARED:(\d+)", line)
        if m and cur:
            usage[cur] = (int(m.group(1)), int(m.group(2)))
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        counts[cur]["_total"] += 1
        base = op.split(".")[0]
        if base in KEYS:
            counts[cur][base] += 1
    names = list(counts)
    pretty = demangle(names)
    lines = [f"SASS summary of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a); counts are static instructions per kernel",
             f"{'kernel':70s} {'regs':>4s} {'smem':>6s} {'instr':>6s}  " + " ".join(f"{k:>8s}" for k in KEYS)]
    total = collections.Counter()
    for n, p in sorted(zip(names, pretty), key=lambda t: t[1]):
        c = counts[n]
        short = re.sub(r"\(.*", "", p).replace("pvd::", "").replace("void ", "")
        reg, sh = usage.get(n, (0, 0))
        lines.append(f"{short[:70]:70s} {reg:4d} {sh:6d} {c['_total']:6d}  " + " ".join(f"{c[k]:8d}" for k in KEYS))
        total.update(c)
    lines.append(f"{'TOTAL':70s} {'':4s} {'':6s} {total['_total']:6d}  " + " ".join(f"{total[k]:8d}" for k in KEYS))
    open(OUT, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[-1:]))
    print("wrote", OUT, len(names), "kernels")


if __name__ == "__main__":
    main()
