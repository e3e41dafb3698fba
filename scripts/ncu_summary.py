#!/usr/bin/env python
"""Turns `ncu -i <rep> --page raw --csv` of the step's kernels into (1) profiles/ncu_traffic.json, the per-launch DRAM traffic
bench.py reports as roofline.traffic, and (2) a short text table.   usage: python scripts/ncu_summary.py RAW.csv [TAG] [--merge]

--merge keeps kernels already present in ncu_traffic.json that this capture does not contain (captures of different workloads);
--suffix=S appends S to every kernel name of this capture (a variant of a kernel that is already in the table)."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = {
    "duration_us": "gpu__time_duration.sum",
    "dram_read": "dram__bytes_read.sum",
    "dram_write": "dram__bytes_write.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "l1_hit_pct": "l1tex__t_sector_hit_rate.pct",
    "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "smem_dyn": "launch__shared_mem_per_block_dynamic",
    # L2-side work of the scattered field kernels, quoted against the MEASURED random-access peaks of scripts/micro/red_peak.cu
    "ld_sectors": "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "red_sectors": "lts__t_sectors_srcunit_tex_op_red.sum",
    "lts_sectors": "lts__t_sectors.sum",
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ns": 1e-3, "ms": 1e3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}


def main():
    raw = sys.argv[1]
    tag = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else os.path.basename(raw).split("_ncu")[0]
    merge = "--merge" in sys.argv
    suffix = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--suffix=")), "")
    rows = list(csv.reader(l for l in open(raw) if not l.startswith("==")))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ix = {k: hdr.index(v) for k, v in COLS.items() if v in hdr}
    kn = hdr.index("Kernel Name")

    def val(r, k):
        if k not in ix or r[ix[k]] in ("", "n/a"):
            return None
        return float(r[ix[k]].replace(",", "")) * UNIT.get(units[ix[k]], 1.0)

    kernels, lines = {}, []
    for r in body:
        name = re.sub(r"^(void\s+)?(pvd::)?", "", r[kn]).split("(")[0].split("<")[0] + suffix
        d = {k: val(r, k) for k in COLS}
        d["dram_bytes"] = (d["dram_read"] or 0.0) + (d["dram_write"] or 0.0)
        d["source"] = f"profiles/{os.path.basename(raw)}"
        kernels.setdefault(name, d)   # first launch of each kernel
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    out = {"how": "ncu --set full --clock-control none --import-source on, one launch per kernel, cold caches; per-kernel `source` names the raw "
                  "metric table under profiles/", "kernels": {}}
    if merge and os.path.exists(path):
        out["kernels"] = json.load(open(path))["kernels"]
    out["kernels"].update(kernels)
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    f = lambda v, p=1: "-" if v is None else f"{v:.{p}f}"
    lines.append(f"{'kernel':28s} {'us':>7s} {'DRAM rd MB':>10s} {'wr MB':>7s} {'DRAM%':>6s} {'L2hit%':>6s} {'L1hit%':>6s} {'SM%':>5s} {'tensor%':>7s} {'warps%':>6s} {'regs':>4s} {'grid':>6s}")
    for n, d in kernels.items():
        d = {k: v for k, v in d.items() if k != "source"}
        lines.append(f"{n:28s} {f(d['duration_us'], 2):>7s} {f((d['dram_read'] or 0) / 1e6, 2):>10s} {f((d['dram_write'] or 0) / 1e6, 2):>7s} {f(d['dram_pct']):>6s} "
                     f"{f(d['l2_hit_pct']):>6s} {f(d['l1_hit_pct']):>6s} {f(d['sm_pct']):>5s} {f(d['tensor_pct'], 2):>7s} {f(d['warps_active_pct']):>6s} "
                     f"{f(d['regs'], 0):>4s} {f(d['grid'], 0):>6s}")
    txt = "\n".join(lines)
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.txt"), "w").write(
        f"# {tag}: ncu --set full --clock-control none --import-source on, one launch of each kernel of the step (eager launches of bench.py,\n"
        f"# cold caches, serialised: durations are NOT bench numbers); raw metrics: profiles/{tag}_ncu_full_raw.csv\n" + txt + "\n")
    print(txt)


if __name__ == "__main__":
    main()
