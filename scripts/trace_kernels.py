"""In-kernel timeline of the hot-path kernels (diagnostic; never part of a bench number).

Builds the library with -DPVD_TRACE (libpvd_b200_trace.so: thread 0 of every CTA / lane 0 of every ray stamps clock64() at phase
boundaries into a device buffer), runs the bench workload eagerly for a few steps and prints where the per-CTA time goes.

    PVD_TRACE=1 python scripts/trace_kernels.py [--rays 4096] > profiles/r01_trace_timeline.txt
"""
import argparse
import ctypes as C
import os
import sys

os.environ["PVD_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "aaai2023-pvd_b200"))

import numpy as np
import torch


def pct(a, q):
    return float(np.percentile(a, q)) if len(a) else float("nan")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--levels", type=int, default=14)
    args = ap.parse_args()
    from pvd_b200 import _build
    _build.build_native()
    from pvd_b200 import _native as nv
    from pvd_b200 import synthetic as syn
    from pvd_b200.engine import HashTrainEngine
    from pvd_b200.fused import HashNeRFField
    import bench

    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    field = HashNeRFField(num_levels=args.levels, desired_resolution=2048).to(dev)
    _, bitfield, _ = syn.lego_bitfield()
    eng = HashTrainEngine(field, torch.from_numpy(bitfield), args.rays, loss_scale=65536.0, device=dev)
    eng.stage()
    host = bench.make_workload(args.rays, 20, seed=0, rank=0)
    devb = [(a.to(dev), b.to(dev), c.to(dev)) for a, b, c in host]
    for i in range(16):
        eng.rays_o, eng.rays_d, eng.gt = devb[i]
        eng.step(warmup=True)
    eng.finish_warmup()
    lib = nv.lib()
    REC = 8192
    bufs = {}
    for name in ("raymarch", "field_hash"):
        b = torch.zeros(REC, 16, dtype=torch.int64, device=dev)
        fn = getattr(lib, "pvd_debug_trace_" + name)
        fn.argtypes = [C.c_void_p, C.c_uint]
        nv.check(fn(C.c_void_p(b.data_ptr()), REC))
        bufs[name] = b
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for i in range(16, 20):
        for b in bufs.values():
            b.zero_()
        flush.fill_(1)
        eng.rays_o, eng.rays_d, eng.gt = devb[i]
        eng.step()
        torch.cuda.synchronize()
    clk = 1e-3 * torch.cuda.clock_rate() if hasattr(torch.cuda, "clock_rate") else 1.965  # cycles per ns
    ghz = 1.965
    us = lambda cyc: cyc / (ghz * 1e3)

    # ---------------- march
    m = bufs["raymarch"].cpu().numpy().astype(np.int64)[: args.rays]
    dur = us(m[:, 2] - m[:, 0])
    coarse = us(m[:, 1] - m[:, 0])
    heavy = m[:, 3] + m[:, 4] > 0
    g0 = m[:, 8].min()
    print(f"== k_march_count: {args.rays} rays, {int(heavy.sum())} pass the coarse test, {int((m[:, 6] > 0).sum())} have samples, "
          f"{int(m[:, 6].sum())} samples")
    print(f"   kernel span (globaltimer, first warp start -> last warp end): {(m[:, 9].max() - g0) / 1e3:.2f} us; "
          f"last warp START at {(m[:, 8].max() - g0) / 1e3:.2f} us")
    print(f"   coarse test per ray: median {pct(coarse, 50):.2f} us, p95 {pct(coarse, 95):.2f}, max {coarse.max():.2f}")
    for name, sel in (("rejected", ~heavy), ("marching", heavy)):
        d = dur[sel]
        print(f"   {name:9s}: n={len(d)}, duration median {pct(d, 50):.2f} us, p90 {pct(d, 90):.2f}, p99 {pct(d, 99):.2f}, max {d.max() if len(d) else 0:.2f}")
    h = m[heavy]
    if len(h):
        md = us(h[:, 2] - h[:, 1])
        print(f"   marching rays: groups median {pct(h[:, 3], 50):.0f} max {h[:, 3].max()}, general windows median {pct(h[:, 4], 50):.0f} max {h[:, 4].max()}, "
              f"walk iterations median {pct(h[:, 5], 50):.0f} max {h[:, 5].max()}")
        per_group = md / np.maximum(h[:, 3] + h[:, 4] / 4.0, 1)
        print(f"   time per 128-point group: median {pct(per_group, 50):.2f} us, p90 {pct(per_group, 90):.2f}")
        g = np.maximum(h[:, 3], 1)
        for nm, col, den in (("group: probe (pos, cell, load, exit)", 10, g), ("group: prepare (nxt, doubling)", 11, g), ("group: resolve x4", 12, g),
                             ("general window (each)", 13, np.maximum(h[:, 4], 1)), ("loop head (each iteration)", 14, np.maximum(h[:, 3] + h[:, 4], 1))):
            v = us(h[:, col] / den)
            print(f"     {nm:38s}: median {pct(v, 50):.3f} us  p90 {pct(v, 90):.3f}")
        order = np.argsort(-(h[:, 9]))[:5]
        for i in order:
            print(f"     late finisher: start {(h[i, 8] - g0) / 1e3:.2f} us end {(h[i, 9] - g0) / 1e3:.2f} us groups {h[i, 3]} general {h[i, 4]} iters {h[i, 5]} count {h[i, 6]} sm {h[i, 7]}")
        smid = m[:, 7]
        per_sm = np.bincount(smid[heavy], minlength=148)
        print(f"   marching rays per SM: min {per_sm.min()} median {np.median(per_sm):.0f} max {per_sm.max()}")

    # ---------------- composite backward
    cb = bufs["raymarch"].cpu().numpy().astype(np.int64)[4096:4096 + args.rays]
    cb = cb[cb[:, 0] > 0]
    if len(cb):
        g0 = cb[:, 8].min()
        act = cb[cb[:, 9] > 0]
        print(f"== k_composite_bwd: {len(cb)} rays traced, {len(act)} composited; first start -> last end {(act[:, 9].max() - g0) / 1e3:.2f} us; "
              f"last warp START {(cb[:, 8].max() - g0) / 1e3:.2f} us")
        hd = us(act[:, 1] - act[:, 0]); lp = us(act[:, 2] - act[:, 1])
        print(f"   header loads: median {pct(hd, 50):.2f} us p99 {pct(hd, 99):.2f}; sample loop: median {pct(lp, 50):.2f} us p99 {pct(lp, 99):.2f} max {lp.max():.2f}; "
              f"samples/ray median {pct(act[:, 6], 50):.0f} max {act[:, 6].max()}")
        for lo, hi in ((1, 32), (33, 64), (65, 96), (97, 128), (129, 160), (161, 192), (193, 256), (257, 1024)):
            sel = (act[:, 6] >= lo) & (act[:, 6] <= hi)
            if sel.any():
                print(f"     {lo:4d}-{hi:4d} samples: n={int(sel.sum()):4d} loop median {pct(lp[sel], 50):.2f} us max {lp[sel].max():.2f}")
        late = np.argsort(-act[:, 9])[:3]
        for i in late:
            print(f"     late: start {(act[i, 8] - g0) / 1e3:.2f} end {(act[i, 9] - g0) / 1e3:.2f} us cnt {act[i, 6]} header {us(act[i, 1] - act[i, 0]):.2f} loop {us(act[i, 2] - act[i, 1]):.2f}")
    # ---------------- hash field forward
    fall = bufs["field_hash"].cpu().numpy().astype(np.int64)
    f = fall[:2048]
    f = f[f[:, 0] > 0]
    g0 = f[:, 12].min()
    print(f"== k_hash_field_fwd: {len(f)} CTAs; span {(f[:, 14].max() - g0) / 1e3:.2f} us; last CTA start {(f[:, 12].max() - g0) / 1e3:.2f} us")
    names = ["setup(stage weights, tmem alloc)", "gather levels 0-3", "gather levels 4-7", "gather levels 8-11", "gather levels 12-15",
             "sigma_net.0", "sigma_net.1", "color_net.0 (+SH, exp)", "color_net.1", "color_net.2", "outputs"]
    for i, nm in enumerate(names):
        d = us(f[:, i + 1] - f[:, i])
        print(f"   {nm:34s}: median {pct(d, 50):6.2f} us  p90 {pct(d, 90):6.2f}  max {d.max():6.2f}")
    sm = f[:, 13]
    for sid in (0, 1, 73, 147):
        sel = np.where(sm == sid)[0]
        st = np.sort((f[sel, 12] - g0) / 1e3)
        en = np.sort((f[sel, 14] - g0) / 1e3)
        print(f"   SM {sid}: CTA starts {np.round(st, 2).tolist()} ends {np.round(en, 2).tolist()}")
    starts = np.sort((f[:, 12] - g0) / 1e3)
    print(f"   CTA start time percentiles: 25% {pct(starts, 25):.2f}  50% {pct(starts, 50):.2f}  75% {pct(starts, 75):.2f}  100% {starts.max():.2f} us")
    tot = us(f[:, 11] - f[:, 0])
    print(f"   total per CTA: median {pct(tot, 50):.2f} us, max {tot.max():.2f}")
    report_bwd(fall, us, pct)


def report_bwd(fall, us, pct):
    b = fall[2048:4096]
    b = b[b[:, 0] > 0]
    print(f"== k_hash_field_bwd: {len(b)} CTAs (timeline of the LAST tile of each CTA; start/setup/end are per CTA)")
    seq = [("setup (tmem alloc, weights)", 0, 1), ("[all earlier tiles]", 1, 2), ("fwd: sigma_net.0", 2, 6), ("fwd: sigma_net.1", 6, 7), ("fwd: color_net.0", 7, 8),
           ("fwd: color_net.1", 8, 9), ("fwd: color_net.2", 9, 10), ("bwd: color_net.2 (wgrad+dgrad)", 10, 3), ("bwd: color_net.1", 3, 4),
           ("bwd: color_net.0", 4, 5), ("bwd: sigma_net.1", 5, 11), ("bwd: sigma_net.0", 11, 12), ("dx store", 12, 13), ("wgrad flush", 13, 14)]
    for nm, i0, i1 in seq:
        d = us(b[:, i1] - b[:, i0])
        print(f"   {nm:34s}: median {pct(d, 50):6.2f} us  p90 {pct(d, 90):6.2f}  max {d.max():6.2f}")
    tot = us(b[:, 14] - b[:, 0])
    print(f"   total per CTA: median {pct(tot, 50):.2f} us, max {tot.max():.2f}")


if __name__ == "__main__":
    main()


