#!/usr/bin/env python
"""bench.py -- training rays/s (forward + backward) of the PVD volume-rendering hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload hash|vm|hash-vm|mlp-hash] [--levels 14] [--rays R]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Default workload (BASELINE.json configs[1], the one `metric` is quoted on): "hash" teacher training, 4096 rays per GPU per step x
1024 max steps, synthetic 800x800 Lego-shaped scene (pvd_b200/synthetic.py), random-init weights, fp16 tables + fp16 tensor-core MLP
with fp32 accumulation (the reference's own autocast precision), MSE loss, loss scale 65536 (GradScaler's default).
The other BASELINE configurations run through the same harness with --workload: vm (configs[2], TensoRF VM-48 teacher training),
hash-vm (configs[3], hash teacher -> vm student distillation at shared samples), mlp-hash (configs[4], NeRF-MLP teacher -> hash
student, 8192 rays).
A step = near/far -> march -> field query -> composite -> loss -> backward of all of it (and, for N > 1, one NCCL all-reduce
of the gradients); the optimizer step, data loading and density-grid upkeep are outside, as in SURVEY.md 8d.

One JSON line on rank 0 (keys: see the repo's DESIGN.md "measurement").
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "aaai2023-pvd_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "training rays/sec (fwd+bwd)"
L2_FLUSH_BYTES = 256 << 20


# ------------------------------------------------------------------------------------------------ helpers
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index=0, period=0.002):
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                # NVML's utilisation counter integrates over ~1/6 s, longer than the timed region: keep every sample taken
                # while the region runs (the sampler is started right before it and stopped right after)
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = get_reasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=1.0)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def make_workload(n_rays, n_batches, seed, rank):
    """Per-rank ray batches (different poses per rank) and random ground-truth colours; pinned host tensors."""
    from pvd_b200 import synthetic as syn
    batches = syn.make_ray_batches(n_batches, n_rays, seed=seed + 1000 * rank)
    g = torch.Generator().manual_seed(seed + 7 + rank)
    out = []
    for ro, rd in batches:
        gt = torch.rand(n_rays, 3, generator=g)
        packed = torch.stack([ro, rd, gt]).pin_memory()   # [3, N, 3]: what a loader hands over, one H2D copy per batch
        out.append((packed[0], packed[1], packed[2], packed))
    return out


WORKLOADS = {
    "hash": "hash (INGP L={L} T=2^19 F=2) teacher training, {R} rays/GPU x 1024 max steps, cuda_ray, synthetic 800x800 Lego-shaped scene, "
            "fwd+bwd (MSE), random-init weights",
    "vm": "vm (TensoRF VM-48: sigma 16 + colour 48 components, 300^3) teacher training, {R} rays/GPU x 1024 max steps, cuda_ray, synthetic "
          "800x800 Lego-shaped scene, fwd+bwd (MSE + L1 on the sigma planes), random-init weights",
    "hash-vm": "hash (L={L}) teacher -> vm (VM-48, 300^3) student distillation (main_distill_mutual stage 3: normL2 rgb + feature + colour + sigma "
               "losses), both networks queried at the SAME samples, {R} rays/GPU x 1024 max steps, synthetic Lego-shaped scene, random-init weights",
    "mlp-hash": "mlp (NeRF 8x256, PE 10) teacher -> hash (L={L}) student distillation (main_distill_mutual stage 3), both networks queried at the "
                "SAME samples, {R} rays/GPU x 1024 max steps, synthetic Lego-shaped scene, random-init weights",
    "hash-fp32": "hash (INGP L={L}) teacher training in FP32 END TO END (the reference without --fp16; north_star's fp32 bound; not a BASELINE config, "
                 "reported beside them), {R} rays/GPU x 1024 max steps, cuda_ray, synthetic Lego-shaped scene, fwd+bwd (MSE), random-init weights",
    "mlp": "mlp (NeRF 8x256, PE 10, skip) teacher training (main_just_train_tea.py --model_type mlp; not a BASELINE config, reported beside them), "
           "{R} rays/GPU x 1024 max steps, cuda_ray, synthetic 800x800 Lego-shaped scene, fwd+bwd (MSE), random-init weights",
}
DEFAULT_RAYS = {"hash": 4096, "vm": 4096, "hash-vm": 4096, "mlp-hash": 8192, "mlp": 4096, "hash-fp32": 4096}
ALL_WORKLOADS = ("hash", "vm", "hash-vm", "mlp-hash", "mlp", "hash-fp32")
ONE_GPU_ONLY = ("mlp", "hash-fp32")       # not BASELINE configs: reported beside them on one GPU
PAIR_RATES = (1.0, 0.002, 0.002, 0.002)   # main_distill_mutual.py:174-177
L1_REG = 1e-4                             # main_distill_mutual.py:178 / main_just_train_tea.py:170
MLP_FLOPS_PER_SAMPLE = 865280             # 2 * (63*256 + 5*256^2 + 319*256 + 256*28)
MLP_BWD_FLOPS_PER_SAMPLE = 2 * (28 * 256 + 6 * 256 * 256) + MLP_FLOPS_PER_SAMPLE   # data gradients (layers 7..1, hidden parts) + weight gradients
TAIL_FLOPS_FWD_BWD = 54500                # sigma_net + color_net, 18.2 kFLOP forward, x3 with both gradient GEMMs


# ------------------------------------------------------------------------------------------------ CPU baseline (oracle port)
def cpu_baseline(workload, levels, n_rays, budget_s=15.0):
    """The oracle's CPU training step (C kernels + torch-CPU networks, fp32) on a bounded sample of the same workload: whole steps
    of n_rays rays until `budget_s` seconds are spent (at least one)."""
    from oracle import cpu, field
    from pvd_b200 import synthetic as syn
    torch.set_num_threads(os.cpu_count() or 1)
    _, bitfield, _ = syn.lego_bitfield()
    torch.manual_seed(0)
    lin = lambda dims: [torch.nn.Linear(i, o, bias=False).weight for i, o in dims]
    TAIL = ((2 * levels, 64), (64, 16), (31, 64), (64, 64), (64, 3))

    def hash_model(train):
        offsets, pls = cpu.grid_offsets(3, levels, 16, 19, desired_resolution=2048)
        emb = torch.empty(int(offsets[-1]), 2).uniform_(-1e-4, 1e-4).requires_grad_(train)
        ws = lin(TAIL)
        return (lambda x, d: field.hash_field_forward(x, d, emb, offsets, pls, 16, ws)), [emb] + ws

    def vm_model():
        mk = lambda r, w: [(0.1 * torch.randn(1, r, 300, w)).requires_grad_(True) for _ in range(3)]
        sm, sv, cm, cv = mk(16, 300), mk(16, 1), mk(48, 300), mk(48, 1)
        bw, cw = lin(((144, 15),))[0], lin(((31, 64), (64, 64), (64, 3)))
        aabb = torch.tensor([-1.0, -1, -1, 1, 1, 1])
        fn = lambda x, d: field.vm_field_forward(x, d, sm, sv, cm, cv, bw, cw, aabb)
        return fn, sm + sv + cm + cv + [bw] + cw, (sm, sv)

    def mlp_model(train=False):
        dims = [(63, 256)] + [(256, 256)] * 3 + [(319, 256)] + [(256, 256)] * 2 + [(256, 28)]
        ls = [torch.nn.Linear(i, o) for i, o in dims]
        nw, nb, tw = [l.weight.detach() for l in ls], [l.bias.detach() for l in ls], [w.detach() for w in lin(((28, 64),) + TAIL[1:])]
        for t in (nw + nb + tw) if train else ():
            t.requires_grad_(True)
        fn = lambda x, d: field.mlp_field_forward(x, d, nw, nb, tw)
        return (fn, nw + nb + tw) if train else fn

    if workload in ("hash", "hash-fp32"):
        f_s, params = hash_model(True)
        one = lambda ro, rd, gt: field.render_train_step(ro, rd, bitfield, gt, lambda x, d: f_s(x, d)[:2])["loss"]
    elif workload == "vm":
        f_s, params, (sm, sv) = vm_model()
        one = lambda ro, rd, gt: field.render_train_step(ro, rd, bitfield, gt, lambda x, d: f_s(x, d)[:2])["loss"] + L1_REG * field.vm_density_loss(sm, sv)
    elif workload == "hash-vm":
        f_t, _ = hash_model(False)
        f_s, params, (sm, sv) = vm_model()
        one = lambda ro, rd, gt: field.pair_distill_step(ro, rd, bitfield, f_s, f_t, PAIR_RATES, l1_reg=L1_REG * field.vm_density_loss(sm, sv))["loss"]
    elif workload == "mlp":
        f_s, params = mlp_model(True)
        one = lambda ro, rd, gt: field.render_train_step(ro, rd, bitfield, gt, lambda x, d: f_s(x, d)[:2])["loss"]
    else:
        f_t = mlp_model()
        f_s, params = hash_model(True)
        one = lambda ro, rd, gt: field.pair_distill_step(ro, rd, bitfield, f_s, f_t, PAIR_RATES)["loss"]
    batches = syn.make_ray_batches(8, n_rays, seed=123)
    g = torch.Generator().manual_seed(5)
    done, i, t0 = 0, 0, time.perf_counter()
    while True:
        ro, rd = batches[i % len(batches)]
        i += 1
        gt = torch.rand(n_rays, 3, generator=g)
        for p_ in params:
            p_.grad = None
        one(ro, rd, gt).backward()
        done += n_rays
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": "rays/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{done} rays ({done // n_rays} steps of {n_rays}) of the same workload, fp32, oracle C kernels (1 thread) + torch-CPU "
                      f"networks ({torch.get_num_threads()} threads), {dt:.1f} s"}


# ------------------------------------------------------------------------------------------------ our arm
def build_engine(args, dev, bitfield):
    """The engine of the chosen workload, module-default random init under seed 0."""
    from pvd_b200.engine import HashTrainEngine, PairDistillEngine, VMTrainEngine
    from pvd_b200.fused import HashNeRFField
    torch.manual_seed(0)
    bf = torch.from_numpy(bitfield)
    kw = dict(loss_scale=65536.0, device=dev)
    w = args.workload
    # A TRAINED hash model gathers its fp32 master table in place (table_fp16=False): random 8-byte gathers run at the same sector rate
    # as 4-byte ones (scripts/micro/red_peak.cu: 231 vs 244 G loads/s), so the per-step fp32 -> fp16 re-cast of the whole table that an
    # fp16 shadow needs after every optimizer step (63 MB of traffic, the reference's grid.py:52) buys nothing.  The frozen teacher of a
    # distillation pair keeps its fp16 shadow (cast once, half the L2 footprint).
    if w == "hash":
        return HashTrainEngine(HashNeRFField(num_levels=args.levels, desired_resolution=2048, table_fp16=False).to(dev), bf, args.rays, **kw)
    if w == "hash-fp32":
        return HashTrainEngine(HashNeRFField(num_levels=args.levels, desired_resolution=2048, table_fp16=False, fp32=True).to(dev), bf, args.rays, **kw)
    if w == "vm":
        from pvd_b200.fused_vm import VMNeRFField
        return VMTrainEngine(VMNeRFField(resolution0=300).to(dev), bf, args.rays, l1_reg_weight=L1_REG, **kw)
    if w == "hash-vm":
        from pvd_b200.fused_vm import VMNeRFField
        tea = HashNeRFField(num_levels=args.levels, desired_resolution=2048, is_teacher=True).to(dev)
        return PairDistillEngine(tea, VMNeRFField(resolution0=300).to(dev), bf, args.rays, rates=PAIR_RATES, stage=3, l1_reg_weight=L1_REG, **kw)
    from pvd_b200.fused_mlp import MLPNeRFField
    if w == "mlp":
        from pvd_b200.engine import MLPTrainEngine
        return MLPTrainEngine(MLPNeRFField().to(dev), bf, args.rays, **kw)
    tea = MLPNeRFField().to(dev)
    return PairDistillEngine(tea, HashNeRFField(num_levels=args.levels, desired_resolution=2048, table_fp16=False).to(dev), bf, args.rays,
                             rates=PAIR_RATES, stage=3, **kw)


def phase_plan(eng):
    """[(phase name, launcher)] of one eager step, for per-kernel event timing (roofline pass only)."""
    import ctypes as C
    st = lambda: C.c_void_p(torch.cuda.current_stream(eng.dev).cuda_stream)
    rs = lambda: eng.sets[eng.cur]
    pair = hasattr(eng, "tea")
    plan = [("march", lambda: (eng._march_count(st(), rs()), eng._march_write(st(), rs(), eng.M)))]
    if pair:
        plan.append(("teacher_fwd", lambda: eng.tea.forward(st(), rs().xyzs, rs().dirs, eng.M, eng.sigmas_tea, eng.rgbs_tea, eng.feat_tea, eng.status)))
        plan.append(("field_fwd", lambda: eng.ops.forward(st(), rs().xyzs, rs().dirs, eng.M, eng.sigmas, eng.rgbs, eng.feat, eng.status)))
    else:
        plan.append(("field_fwd", lambda: eng._forward(st(), rs(), eng.M, eng.M)))
    plan.append(("composite", lambda: eng._loss_backward(st(), rs(), eng.M, eng.M)))
    gfeat = eng.grad_feat if pair else None
    nval = (lambda: None) if pair else (lambda: rs().counter)
    if eng.ops.kind == "hash" and eng.ops.dx_ws is not None:
        for name, ph in (("mlp_bwd", 1), ("scatter", 2)):
            plan.append((name, lambda ph=ph: eng.ops.backward(st(), rs().xyzs, rs().dirs, eng.grad_sigmas, eng.grad_rgbs, gfeat, eng.M, nval(),
                                                              eng.gw_ws, eng.status, phases=ph)))
    else:
        plan.append(("field_bwd", lambda: eng._field_backward(st(), rs(), eng.M, None)))
    return plan


def time_phases(eng, plan):
    """One step with events between phases (perturbs the step by a few event records)."""
    cur = torch.cuda.current_stream(eng.dev)
    eng._zeros.zero_()
    eng._clear_big(cur)
    cur.wait_stream(eng._side)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(len(plan) + 1)]
    e[0].record()
    for i, (_, fn) in enumerate(plan):
        fn()
        e[i + 1].record()
    torch.cuda.synchronize()
    return {name: e[i].elapsed_time(e[i + 1]) for i, (name, _) in enumerate(plan)}


def roofline_kernels(eng):
    """(kernel, phase, bound, algorithmic bytes or FLOPs per sample) for the field kernels of the step."""
    out = []
    pair = hasattr(eng, "tea")
    if pair:
        t = eng.tea
        if t.kind == "hash":
            out.append(("k_hash_field_fwd(teacher)", "teacher_fwd", "hbm", t.algorithmic_bytes()[0]))
        elif t.kind == "vm":
            out.append(("k_vm_field_fwd(teacher)", "teacher_fwd", "hbm", t.algorithmic_bytes()[0]))
        else:
            out.append(("k_mlp_field_fwd", "teacher_fwd", "tensor", MLP_FLOPS_PER_SAMPLE))
    o = eng.ops
    fb, bb = o.algorithmic_bytes()
    if o.kind == "mlp":
        out.append(("k_mlp_field_fwd", "field_fwd", "tensor", MLP_FLOPS_PER_SAMPLE))
        out.append(("k_mlp_trunk_bwd+k_mlp_wgrad", "field_bwd", "tensor", MLP_BWD_FLOPS_PER_SAMPLE))
    elif o.kind == "hash" and o.fp32:
        out.append(("k_hash_field_fwd_f32", "field_fwd", "hbm", fb))
        out.append(("k_hash_field_bwd_f32", "field_bwd", "hbm", bb))
    elif o.kind == "hash":
        out.append(("k_hash_field_fwd", "field_fwd", "hbm", fb))
        if o.dx_ws is not None:
            out.append(("k_hash_field_bwd", "mlp_bwd", "tensor", TAIL_FLOPS_FWD_BWD))   # the MLP backward GEMMs (forward recomputed)
            out.append(("k_hash_scatter", "scatter", "hbm", bb))                        # fp32 reductions + d(encoding) in
        else:
            out.append(("k_hash_field_bwd", "field_bwd", "hbm", bb))
    else:
        out.append(("k_vm_field_fwd", "field_fwd", "hbm", fb))
        two = getattr(o, "scatter_ws", None) is not None                           # MLP kernel + stand-alone scatter kernel
        out.append(("k_vm_field_bwd+k_vm_scatter" if two else "k_vm_field_bwd", "field_bwd", "hbm", fb + bb))   # re-gather + reductions
    return out


def l2_peaks():
    """MEASURED random-access rates of this pool's B200 (scripts/micro/red_peak.cu -> profiles/r02_red_peak.json): 32-byte-sector loads and
    reductions per second with every lane of a warp in a different sector -- what bounds the hash gather / scatter (their tables live in L2)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r02_red_peak.json")))
        return {"ld": d["ld_4B_Gops"] * 1e9, "red": max(d["red_f32_v2_Gops"], d["red_f32_v4_Gops"], d["red_f32_x1_Gops"]) * 1e9}
    except Exception:
        return None


def l2_roofline(kernel, ms):
    """{"bound": "l2_sector_ld" | "l2_atomic", achieved / peak in G sector-ops/s} for a hash field kernel, from the sector count of the
    committed ncu capture of the same workload and this run's event-timed duration; None when either is missing."""
    try:
        k = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["kernels"].get(kernel.split("(")[0])
        pk = l2_peaks()
        if not k or not pk:
            return None
        if "bwd" in kernel or "scatter" in kernel:
            n, peak, bound = k.get("red_sectors"), pk["red"], "l2_atomic"
        else:
            n, peak, bound = k.get("ld_sectors"), pk["ld"], "l2_sector_ld"
        if not n:
            return None
        ach = n / (ms * 1e-3)
        return {"bound": bound, "sector_ops_per_launch": n, "achieved_Gops": ach / 1e9, "peak_Gops": peak / 1e9, "frac": ach / peak,
                "peak_kind": "measured: scripts/micro/red_peak.cu (profiles/r02_red_peak.json)", "count_source": k.get("source")}
    except Exception:
        return None


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/ncu_traffic.json), or None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        parts = [t["kernels"].get(k.split("(")[0]) for k in kernel.split("+")]   # "a+b": a phase made of two launches
        if any(p is None for p in parts):
            return None, None
        return sum(p["dram_bytes"] for p in parts), " + ".join(sorted({p.get("source", "profiles/") for p in parts}))
    except Exception:
        return None, None


def _timed(fn_load, fn_step, flush, steps, world, dev):
    """K steps, each bracketed by CUDA events on the launching stream, L2 flushed between steps outside the brackets; returns the
    summed milliseconds, MAX over ranks."""
    import torch.distributed as dist
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    # The step is timed on the DEVICE.  With W ranks every step ends in an exchange that waits for the slowest rank, so a rank whose
    # HOST thread is descheduled for a millisecond (8 Python processes share the box) stalls the other ranks INSIDE their event
    # brackets.  The stream is therefore gated by a ~10 ms spin kernel while the host enqueues the whole timed region behind it, and
    # two untimed steps absorb the ranks' start skew: the K timed steps then run back to back from the queue.
    torch.cuda._sleep(20_000_000)
    for i in range(2):
        fn_load()
        flush.fill_(0xA5)
        fn_step()
    for i in range(steps):
        fn_load()
        flush.fill_(i & 0xFF)
        ev[i][0].record()
        fn_step()
        ev[i][1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    per = sorted(a.elapsed_time(b) for a, b in ev)
    _timed.last_percentiles = {"median": per[len(per) // 2], "p10": per[len(per) // 10], "p90": per[(9 * len(per)) // 10]}   # this rank's steps
    tt = torch.tensor([sum(per)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return float(tt.item())


def measure_ours(args, workload, n_rays, rank, world, local, headline):
    """Everything bench.py reports for ONE workload on this rank's GPU; a dict on every rank (times are max over ranks)."""
    import torch.distributed as dist
    from pvd_b200 import optim as pvd_optim
    from pvd_b200 import synthetic as syn

    dev = torch.device("cuda", local)
    _, bitfield, sha = syn.lego_bitfield()
    wargs = argparse.Namespace(**vars(args))
    wargs.workload, wargs.rays = workload, n_rays
    eng = build_engine(wargs, dev, bitfield)
    pair = hasattr(eng, "tea")
    steps = args.steps if headline else min(args.steps, args.secondary_steps)
    eng.stage()
    # the timed step is a whole fwd+bwd as a trainer with an EXTERNAL optimizer must run it: the parameters changed since the last
    # step, so the fp16 table shadow is re-cast and the weight tiles re-packed at the top of every step (the reference pays its
    # equivalent inside its forward, gridencoder/grid.py:52), and the small weight gradients leave the step in parameter shapes
    eng.restage_each_step = True
    eng.unpack_each_step = True
    n_b = 16 + args.warmup + steps
    host = make_workload(n_rays, min(n_b, 64), seed=0, rank=rank)
    devb = [(a.to(dev), b.to(dev), c.to(dev)) for a, b, c, _ in host]

    use_graph = False
    pipelined = False
    pipe = {"i": 0, "nb": 17}   # pipelined mode: index of the next replay (its parity selects the ray set that is computed on)

    def load(i, from_host=False):
        """Make batch i the input of the next step.  Pipelined: batch i goes into the set the NEXT replay marches (it is computed
        on one replay later), which is exactly one batch of look-ahead; the caller passes consecutive i.  A distillation step has
        no ground-truth colours (the teacher's rendering is the target): only the rays are copied."""
        ro, rd, gt = (host if from_host else devb)[i % len(devb)][:3]
        if pipelined:
            rs = eng.sets[(pipe["i"] + 1) & 1]
        elif use_graph or from_host:  # the captured graph reads the engine's static input buffers
            rs = eng.sets[eng.cur]
        else:
            eng.rays_o, eng.rays_d, eng.gt = ro, rd, gt
            return
        rs.rays_o.copy_(ro, non_blocking=True); rs.rays_d.copy_(rd, non_blocking=True)
        if not pair:
            rs.gt.copy_(gt, non_blocking=True)

    def load_next():
        load(pipe["nb"]); pipe["nb"] += 1

    def run_step():
        if pipelined:
            eng.replay_pipelined(pipe["i"])
            pipe["i"] += 1
        elif use_graph:
            eng.replay()
        else:
            eng.step()

    exchange = None
    big = eng.ops.big_grad()
    if world > 1:
        # one gradient exchange per step, INSIDE the step (engine._epilogue; in the pipelined graphs the next batch's march runs beside
        # it).  fp16 payload (default): the big parameter gradient is cast once (x 1/W, saturating) and summed in half precision -- the
        # precision the reference ACCUMULATES table gradients in (gridencoder.cu:299-305); on NVSwitch the sum is done by the switch
        # (multimem, csrc/collective.cu), else by NCCL.  The small MLP gradients stay fp32.  --grad-comm fp32: plain NCCL on fp32.
        from pvd_b200.dist import TableGradExchange
        if args.grad_comm == "fp32":
            class _Fp32Exchange:
                kind, why, world = "nccl-fp32", "", dist.get_world_size()
                def __call__(self):
                    dist.all_reduce(big); dist.all_reduce(eng.gw_ws)
            exchange = _Fp32Exchange()
        else:
            exchange = TableGradExchange(big, eng.gw_ws, mode={"fp16": "nccl", "multimem": "multimem", "p2p": "p2p", "auto": "auto"}[args.grad_comm],
                                         blocks=args.mm_blocks, unroll=args.mm_unroll)
        if rank == 0:
            print(f"[bench] {workload}: gradient exchange = {exchange.kind} {exchange.why}", file=sys.stderr)

    # ---- 16 sizing steps (the reference's mean_count warm-up), then W untimed steps at the steady-state M
    for i in range(16):
        load(i)
        eng.step(warmup=True)
    eng.finish_warmup()
    if exchange is not None and exchange.kind in ("multimem", "p2p"):
        # self-check of the in-switch reduction against NCCL on this step's real gradients; any rank's mismatch -> all fall back
        ref = (big * exchange.pre_scale).clamp(-65504, 65504).to(torch.float16)
        dist.all_reduce(ref)
        exchange()
        torch.cuda.synchronize()
        err = (exchange.payload.float() - ref.float()).abs().max() / (ref.float().abs().max() + 1e-20)
        bad = torch.tensor([0.0 if (float(err) < 2e-2 and exchange.barrier_error() == 0) else 1.0], device=dev)
        dist.all_reduce(bad)
        if float(bad.item()) > 0:
            exchange.kind, exchange.why = "nccl", f"{exchange.kind} self-check failed (rel err {float(err):.3g})"
        if rank == 0:
            print(f"[bench] custom exchange self-check: max rel err {float(err):.3g} -> using {exchange.kind}", file=sys.stderr)
        del ref
    eng.exchange = exchange
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    serial_ms = None

    def prime_pipeline():
        pipe["i"] = 0
        rs = eng.sets[0]   # prime: the next batch is marched eagerly, its compute is replay 0
        ro, rd, gt = devb[pipe["nb"] % len(devb)]
        pipe["nb"] += 1
        rs.rays_o.copy_(ro); rs.rays_d.copy_(rd); rs.gt.copy_(gt)
        eng.march(0)

    if not args.no_graph:
        try:  # CUDA graphs for the whole step: removes launch gaps and host work from the loop
            for rs in eng.sets:
                rs.reset_views()   # the graphs read the sets' own input buffers
            use_graph = True
            load(16)
            eng.capture()
            if not args.no_pipeline:
                # reference point: the serial single-graph step, timed the same way (reported as serial_ms_per_step)
                for i in range(args.warmup):
                    load_next(); run_step()
                n_ser = min(steps, 50)
                serial_ms = _timed(load_next, run_step, flush, n_ser, world, dev) / n_ser
                # steady state: the march of batch i+1 runs beside the field kernels / the exchange of batch i (two graphs, two ray sets)
                eng.capture_pipelined()
                pipelined = True
                prime_pipeline()
        except Exception as ex:  # noqa: BLE001
            print(f"[bench] CUDA graph capture failed ({ex!r}); running eagerly", file=sys.stderr)
            use_graph = pipelined = False
    for i in range(args.warmup):
        load_next()
        run_step()
    torch.cuda.synchronize()
    assert int(eng.status.item()) == 0, "tensor-core pipeline reported a timeout"

    sampler = ClockSampler(local)
    sampler.start()
    # ---- timed region: K steps, each bracketed by events, L2 flushed between steps (outside the brackets)
    t_ms = _timed(load_next, run_step, flush, steps, world, dev)
    step_pct = dict(_timed.last_percentiles)

    # ---- per-kernel pass (same steps again, events around the field kernels; used for the roofline only)
    plan = phase_plan(eng)
    kt = {name: 0.0 for name, _ in plan}
    S_total = 0
    reps = min(steps, 50)
    was_pipelined, pipelined = pipelined, False
    saved_exchange, eng.exchange = eng.exchange, None
    eng.cur = 0
    for i in range(reps):
        load(16 + args.warmup + i)
        flush.fill_(i & 0xFF)
        t = time_phases(eng, plan)
        for k in kt:
            kt[k] += t[k]
        S_total += min(int(eng.counter[0].item()), eng.M)
    for k in kt:
        kt[k] /= reps
    S_mean = S_total / reps
    pipelined = was_pipelined
    eng.exchange = saved_exchange

    # ---- end-to-end through the public API with HOST buffers: H2D of rays (+ gt), step, D2H of the loss.
    # Pipelined engine: both copies are NODES of the step's graphs (engine.capture_pipelined(host_io=True)): the H2D of batch i+1 heads
    # the march branch of step i, the D2H of the loss words ends it; the caller only writes the next batch into the pinned staging
    # buffer.  The CPU may run at most two replays ahead of the GPU (a staging buffer is re-used every second step).
    e2e_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    loss_dev = eng._loss_dev()   # pair: [total, 4 norms]; single model: 64 (loss, rays) slots, summed on the host
    host_io = pipelined
    nb = pipe["nb"]
    if host_io:
        eng.capture_pipelined(host_io=True)
        prime_pipeline()
        nb = pipe["nb"]
        loss_host = eng.host_loss
    else:
        loss_host = torch.empty(loss_dev.numel(), dtype=torch.float32).pin_memory()
    if not use_graph:
        eng.rays_o, eng.rays_d, eng.gt = (torch.empty(n_rays, 3, device=dev) for _ in range(3))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for i in range(steps):
        flush.fill_(i & 0xFF)
        e2e_ev[i][0].record()
        if host_io:
            if i >= 2:
                e2e_ev[i - 2][1].synchronize()          # the staging buffer's previous batch has been consumed
            eng.sets[(pipe["i"] + 1) & 1].host.copy_(host[nb % len(host)][3]); nb += 1
            eng.replay_pipelined(pipe["i"], host_io=True)
            pipe["i"] += 1
        else:
            load(nb, from_host=True); nb += 1   # pinned host rays (+ gt) -> device
            run_step()
            loss_host.copy_(loss_dev, non_blocking=True)
        e2e_ev[i][1].record()
    torch.cuda.synchronize()
    e2e_ms = sum(a.elapsed_time(b) for a, b in e2e_ev)
    tt = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_ms = float(tt.item())
    clocks = sampler.stop()
    assert int(eng.status.item()) == 0
    pipe["nb"] = nb

    # ---- the same step WITHOUT the per-step re-stage / unpack (round 1's definition of a step), for continuity
    extra = {}
    n_x = min(steps, 50)
    if use_graph and pipelined:
        try:
            if exchange is not None and not hasattr(exchange, "payload"):
                raise RuntimeError("iteration timing needs the fp16-payload exchange")
            eng.restage_each_step = eng.unpack_each_step = False
            eng.capture_pipelined()
            prime_pipeline()
            for i in range(3):
                load_next(); run_step()
            extra["ms_per_step_no_restage"] = _timed(load_next, run_step, flush, n_x, world, dev) / n_x
            # ---- the whole training ITERATION: step + exchange + fused AdamW (GradScaler check, fp32 masters, fp16 shadow, weight
            # tiles, gradient zeroing) in the same graphs; no re-stage, no gradient memset (csrc/optim.cu)
            opt = pvd_optim.for_engine(eng, lr=1e-3, lr2=1e-4, exchange=exchange if hasattr(exchange, "payload") else None)
            eng.attach_optimizer(opt)
            eng.capture_pipelined()
            prime_pipeline()
            for i in range(3):
                load_next(); run_step()
            it_ms = _timed(load_next, run_step, flush, n_x, world, dev) / n_x
            stt = opt.read_state()
            extra["iteration"] = {"ms": it_ms, "rays_per_s": n_rays * world / (it_ms * 1e-3), "optimizer_steps": int(stt.step), "skipped": int(stt.skipped),
                                  "what": "step + gradient exchange + fused AdamW (found_inf check, fp32 master update, fp16 table shadow, weight "
                                          "tiles re-packed, gradients zeroed): a complete training iteration per graph replay",
                                  "launches_per_iteration": eng.launches_per_step + opt.kernels_per_step}
            assert int(eng.status.item()) == 0
        except Exception as ex:  # noqa: BLE001
            extra["iteration_error"] = repr(ex)[:200]

    peaks, peak_kind = measured_peaks()
    rays_total = n_rays * world * steps
    value = rays_total / (t_ms * 1e-3)
    L = args.levels
    # roofline of every field kernel; the headline object is the dominant one (longest average launch)
    roofs = []
    for kernel, phase, bound, per_sample in roofline_kernels(eng):
        ms = kt[phase]
        if bound == "hbm":
            ach, peak, unit = S_mean * per_sample / (ms * 1e-3) / 1e9, peaks["hbm_gbs"], "GB/s"
        else:
            ach, peak, unit = S_mean * per_sample / (ms * 1e-3) / 1e12, peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]), "TFLOP/s"
        traffic, src = ncu_traffic(kernel)
        roofs.append({"bound": bound, "kernel": kernel, "ms": ms, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                      "traffic": traffic, "traffic_source": src, "peak_kind": peak_kind,
                      ("algorithmic_bytes_per_sample" if bound == "hbm" else "algorithmic_flops_per_sample"): per_sample})
        if "hash" in kernel:   # the tables are L2-resident: the kernel's real bound is the L2 sector rate, quoted against a measured peak
            l2 = l2_roofline(kernel, ms)
            if l2:
                roofs[-1]["l2"] = l2
    # the workload's bound: the table / plane traffic (HBM roofline) -- except mlp -> hash, where the teacher's GEMMs dominate
    want = "tensor" if workload in ("mlp-hash", "mlp") else "hbm"
    dom = max((r for r in roofs if r["bound"] == want), key=lambda r: r["ms"])
    n_extra = 2 + (1 if eng.ops.kind == "hash" else 0) + (3 if eng.ops.kind == "mlp" else 0)   # per step: weight pack, weight-gradient unpack (+ the table cast for hash)
    out = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": steps, "warmup": args.warmup,
        "ms_per_step": t_ms / steps, "ms_per_step_percentiles": step_pct, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if workload == "hash-fp32" else "f16", "data": "synthetic", "impl": "ours",
        "config": {"workload": WORKLOADS[workload].format(L=L, R=n_rays), "workload_key": workload,
                   "rays_per_gpu": n_rays, "levels": L, "samples_per_step": S_mean, "M_rows": eng.M,
                   "precision": ("fp32 end to end: fp32 table gather, fp32 MLPs on the CUDA cores, fp32 composite / gradients" if workload == "hash-fp32" else
                                 "fp16 tcgen05 MLP (fp32 accumulate); trained hash table gathered as fp32 master, frozen teacher table / vm planes as fp16 shadows; fp32 composite / gradients"),
                   "loss_scale": 65536,
                   "step": "fwd+bwd of a trainer with an external optimizer: fp16 table shadow re-cast + weight tiles re-packed at the top of "
                           "EVERY timed step, small weight gradients unpacked to parameter shapes at its end" + ("; gradient exchange inside" if world > 1 else ""),
                   "parallelism": (f"rays sharded over {world} GPU(s), one all-reduce of the gradients per step inside the step graph ({exchange.kind}: "
                                   f"{'fp32 NCCL' if exchange.kind == 'nccl-fp32' else 'fp16 payload, ' + ({'multimem': 'in-switch multimem reduction, barriers inside the kernel', 'p2p': 'two-shot reduction over NVLink peer pointers, one kernel with device-side barriers'}.get(exchange.kind, 'NCCL'))}); "
                                   "the next batch's march runs beside the exchange") if world > 1 else "single GPU",
                   "launch": ("two CUDA graphs (even/odd steps): the march of batch i+1 runs on a parallel branch of step i's graph "
                              "(one batch of look-ahead, double-buffered ray sets)") if pipelined else
                             ("one CUDA graph per step" if use_graph else "eager (one launch per kernel)"),
                   "l2": f"flushed between steps ({L2_FLUSH_BYTES >> 20} MiB write, outside the timed brackets)",
                   "timing": "CUDA events around every step on the launching stream, summed over the K steps, max over ranks; the host enqueues the whole "
                             "timed region behind a ~10 ms gate kernel and two untimed steps, so that the K steps run back to back from the queue",
                   "scene_bitfield_sha256": sha[:16]},
        "serial_ms_per_step": serial_ms,
        "kernel_ms": kt,
        "roofline": dom,
        "roofline_all": roofs,
        "e2e": {"value": rays_total / (e2e_ms * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": n_rays * (6 if pair else 9) * 4,
                "d2h_bytes_per_step": 4 * loss_dev.numel(),
                "how": ("H2D of the next batch and D2H of the loss words are nodes of the step's CUDA graphs (pinned staging buffers)"
                        if host_io else "copies issued on the step's stream around the step")},
        "gpu_launches": (eng.launches_per_step + n_extra) * steps,
        "clocks": clocks,
    }
    out.update(extra)
    del eng, flush, devb
    torch.cuda.empty_cache()
    return out


def run_ours(args, rank, world, local):
    import torch.distributed as dist

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # The cyclic garbage collector must not run while a stream is capturing: a collected engine of an EARLIER workload owns pinned host
    # buffers whose release records CUDA events, which invalidates the capture in progress.  Collect explicitly between workloads.
    gc.disable()
    head = measure_ours(args, args.workload, args.rays, rank, world, local, headline=True)
    gc.collect()
    torch.cuda.synchronize()
    # the other BASELINE configurations in the same line (shorter runs): configs[1]-[4] all measured by one invocation
    others = {}
    if args.all_workloads:
        for w in ALL_WORKLOADS:
            if w == args.workload or (w in ONE_GPU_ONLY and world > 1):
                continue
            try:
                r = measure_ours(args, w, DEFAULT_RAYS[w], rank, world, local, headline=False)
                others[w] = {k: r[k] for k in ("value", "unit", "ms_per_step", "serial_ms_per_step", "steps", "e2e", "roofline", "kernel_ms", "gpu_launches",
                                               "ms_per_step_no_restage", "iteration", "iteration_error") if k in r}
                others[w]["config"] = {k: r["config"][k] for k in ("workload", "rays_per_gpu", "samples_per_step", "M_rows")}
            except Exception as ex:  # noqa: BLE001  (a secondary workload must not take the headline line down)
                others[w] = {"error": repr(ex)[:300]}
            gc.collect()
            torch.cuda.synchronize()
            torch.cuda.empty_cache()
    strong = None
    if world > 1 and args.strong:
        # strong scaling beside the weak headline: the SAME total batch (args.rays) split over the ranks
        try:
            per = max(128, args.rays // world)
            r = measure_ours(args, args.workload, per, rank, world, local, headline=False)
            strong = {"rays_total": per * world, "rays_per_gpu": per, "value": r["value"], "unit": "rays/s", "ms_per_step": r["ms_per_step"],
                      "steps": r["steps"], "e2e": r["e2e"], "note": "same step, total batch fixed: per-GPU work shrinks with N, the gradient exchange does not"}
        except Exception as ex:  # noqa: BLE001
            strong = {"error": repr(ex)[:300]}
        gc.collect()
        torch.cuda.synchronize()
    if rank == 0:
        if strong is not None:
            head["strong_scaling"] = strong
        head["workloads"] = {args.workload: {k: head[k] for k in ("value", "unit", "ms_per_step", "serial_ms_per_step", "steps", "e2e", "roofline", "kernel_ms",
                                                                   "gpu_launches", "ms_per_step_no_restage", "iteration") if k in head}}
        head["workloads"][args.workload]["config"] = {k: head["config"][k] for k in ("workload", "rays_per_gpu", "samples_per_step", "M_rows")}
        head["workloads"].update(others)
        if world == 1 and not args.no_cpu_baseline:
            head["cpu_baseline"] = cpu_baseline(args.workload, args.levels, args.rays, budget_s=args.cpu_budget)
        print(json.dumps(head), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ reference arm
def build_reference(args, dev, ext, bitfield):
    from oracle import cpu, ref_pipeline as rp
    torch.manual_seed(0)
    bf = torch.from_numpy(bitfield).to(dev)

    def hash_net():
        offsets, pls = cpu.grid_offsets(3, args.levels, 16, 19, desired_resolution=2048)
        return rp.RefHashNetwork(ext, offsets, pls).to(dev)

    w = args.workload
    if w == "hash":
        return rp.RefTrainer(ext, hash_net(), bf)
    if w == "hash-fp32":
        return rp.RefTrainer(ext, hash_net(), bf, autocast=False)
    if w == "vm":
        return rp.RefTrainer(ext, rp.RefVmNetwork(ext).to(dev), bf, l1_reg_weight=L1_REG)
    if w == "hash-vm":
        return rp.RefPairTrainer(ext, rp.RefVmNetwork(ext).to(dev), hash_net(), bf, rates=PAIR_RATES, l1_reg_weight=L1_REG)
    if w == "mlp":
        return rp.RefTrainer(ext, rp.RefMlpNetwork(ext).to(dev), bf)
    return rp.RefPairTrainer(ext, hash_net(), rp.RefMlpNetwork(ext).to(dev), bf, rates=PAIR_RATES, l1_reg_weight=0.0)


def measure_reference(args, workload, n_rays, local, ext, steps):
    """The reference's own CUDA extensions + its Python flow (oracle/ref_pipeline.py) on ONE GPU for one workload."""
    from pvd_b200 import synthetic as syn
    dev = torch.device("cuda", local)
    _, bitfield, sha = syn.lego_bitfield()
    wargs = argparse.Namespace(**vars(args))
    wargs.workload, wargs.rays = workload, n_rays
    tr = build_reference(wargs, dev, ext, bitfield)
    pair = workload in ("hash-vm", "mlp-hash")
    host = make_workload(n_rays, min(16 + args.warmup + steps, 64), seed=0, rank=0)
    devb = [(a.to(dev), b.to(dev), c.to(dev)) for a, b, c, _ in host]
    for i in range(16):  # the reference's 16 warm-up iterations with a D2H sync each, then mean_count
        tr.step(*devb[i % len(devb)])
    tr.update_mean_count()
    for i in range(args.warmup):
        tr.step(*devb[(16 + i) % len(devb)])
    torch.cuda.synchronize()
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(steps):
        b = devb[(16 + args.warmup + i) % len(devb)]
        flush.fill_(i & 0xFF)
        ev[i][0].record()
        tr.step(*b)
        ev[i][1].record()
    torch.cuda.synchronize()
    t_ms = sum(a.elapsed_time(b) for a, b in ev)
    # e2e: host buffers in, loss out
    e2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(steps):
        ro, rd, gt = host[(16 + args.warmup + i) % len(host)][:3]
        flush.fill_(i & 0xFF)
        e2[i][0].record()
        loss = tr.step(ro.to(dev, non_blocking=True), rd.to(dev, non_blocking=True), None if pair else gt.to(dev, non_blocking=True))
        _ = loss.detach().to("cpu", non_blocking=True)
        e2[i][1].record()
    torch.cuda.synchronize()
    e2e_ms = sum(a.elapsed_time(b) for a, b in e2)
    # the whole iteration as the reference runs it (distill_mutual/utils.py:802-819): + GradScaler.unscale_/step with torch.optim.AdamW
    it_ms = None
    try:
        params = [p for p in tr.net.parameters() if p.requires_grad]
        opt = torch.optim.AdamW(params, lr=1e-3, betas=(0.9, 0.99), eps=1e-15)
        n_it = min(steps, 20)
        ev3 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_it)]
        inv = 1.0 / tr.loss_scale
        for i in range(n_it + 2):
            b = devb[(16 + i) % len(devb)]
            flush.fill_(i & 0xFF)
            if i >= 2:
                ev3[i - 2][0].record()
            tr.step(*b)
            grads = [p.grad for p in params if p.grad is not None]
            found = torch.zeros(1, device=dev)
            torch._amp_foreach_non_finite_check_and_unscale_(grads, found, torch.full((1,), inv, device=dev))   # GradScaler.unscale_
            opt.step()
            if i >= 2:
                ev3[i - 2][1].record()
        torch.cuda.synchronize()
        it_ms = sum(a.elapsed_time(b) for a, b in ev3) / n_it
    except Exception as ex:  # noqa: BLE001
        print(f"[bench] reference iteration timing failed: {ex!r}", file=sys.stderr)
    clocks = sampler.stop()
    rays_total = n_rays * steps
    cfg = {"workload": WORKLOADS[workload].format(L=args.levels, R=n_rays), "workload_key": workload, "rays_per_gpu": n_rays, "levels": args.levels,
           "M_rows": tr.mean_count + (128 - tr.mean_count % 128), "precision": ("fp32 (the reference without --fp16)" if workload == "hash-fp32" else "torch autocast fp16 (the reference's -O default), GradScaler-style loss scale 65536"),
           "l2": f"flushed between steps ({L2_FLUSH_BYTES >> 20} MiB write)", "scene_bitfield_sha256": sha[:16],
           "what": "unmodified reference CUDA extensions (oracle/_ref, sm_100a rebuild) + cuBLAS GEMMs via F.linear + F.grid_sample + torch "
                   "autograd, Python flow restated in oracle/ref_pipeline.py"}
    out = {"metric": METRIC, "value": rays_total / (t_ms * 1e-3), "unit": "rays/s", "n_gpus": 1, "steps": steps, "warmup": args.warmup,
           "ms_per_step": t_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
           "data": "synthetic", "impl": "reference", "config": cfg,
           "e2e": {"value": rays_total / (e2e_ms * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": n_rays * (6 if pair else 9) * 4, "d2h_bytes_per_step": 4},
           "clocks": clocks}
    if it_ms:
        out["iteration"] = {"ms": it_ms, "rays_per_s": n_rays / (it_ms * 1e-3), "what": "step + GradScaler.unscale_ + torch.optim.AdamW.step (foreach)"}
    del tr, devb, flush
    torch.cuda.empty_cache()
    return out


def run_reference(args, rank, world, local):
    """The reference's own CUDA extensions + its Python flow (oracle/ref_pipeline.py) on ONE GPU; if oracle/_ref cannot be
    loaded, the CPU oracle port on the host cores instead.  Under torchrun rank 0 alone runs: the reference has no multi-GPU path
    (tools/details.md:25), so at N > 1 this line is a ONE-GPU reference (`n_gpus` 1)."""
    if rank != 0:
        return
    why = "no CUDA device"
    try:
        from oracle import ref_pipeline
        ext = ref_pipeline.load_ext()
        have_ref = torch.cuda.is_available()
    except Exception as ex:  # noqa: BLE001
        ext, have_ref = None, False
        why = repr(ex)
    if not have_ref:
        cfg = {"workload": WORKLOADS[args.workload].format(L=args.levels, R=args.rays), "workload_key": args.workload,
               "rays_per_gpu": args.rays, "levels": args.levels}
        cb = cpu_baseline(args.workload, args.levels, args.rays, budget_s=args.cpu_budget if args.cpu_budget != 15.0 else 20.0)
        out = {"metric": METRIC, "value": cb["value"], "unit": "rays/s", "n_gpus": 0, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": 1e3 * args.rays / cb["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic", "impl": "reference", "config": cfg, "cpu_baseline": cb,
               "e2e": {"value": cb["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "note": "oracle/_ref not loadable (" + why[:80] + "); CPU oracle port timed instead"}
        print(json.dumps(out), flush=True)
        return
    torch.cuda.set_device(local)
    out = measure_reference(args, args.workload, args.rays, local, ext, args.steps)
    gc.collect()
    if world > 1:
        out["note"] = f"launched under torchrun with {world} ranks: ONE-GPU reference (the reference has no multi-GPU path), rank 0 only"
    out["workloads"] = {args.workload: {k: out[k] for k in ("value", "unit", "ms_per_step", "steps", "e2e", "iteration") if k in out}}
    if args.all_workloads:
        for w in ALL_WORKLOADS:
            if w == args.workload:
                continue
            try:
                r = measure_reference(args, w, DEFAULT_RAYS[w], local, ext, min(args.steps, args.secondary_steps))
                out["workloads"][w] = {k: r[k] for k in ("value", "unit", "ms_per_step", "steps", "e2e", "iteration") if k in r}
            except Exception as ex:  # noqa: BLE001
                out["workloads"][w] = {"error": repr(ex)[:300]}
                torch.cuda.empty_cache()
    cb = cpu_baseline(args.workload, args.levels, args.rays, budget_s=args.cpu_budget) if not args.no_cpu_baseline else None
    if cb:
        out["cpu_baseline"] = cb
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="hash", choices=sorted(WORKLOADS), help="hash = BASELINE configs[1] (the headline); vm = configs[2]; "
                    "hash-vm = configs[3] (distillation, teacher and student at shared samples); mlp-hash = configs[4]")
    ap.add_argument("--levels", type=int, default=14, help="hash levels: 14 = what PVD builds (network.py:47-51), 16 = BASELINE.json's text")
    ap.add_argument("--rays", type=int, default=None, help="rays per GPU per step (default 4096; 8192 for mlp-hash)")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--grad-comm", default="auto", choices=["auto", "p2p", "multimem", "fp16", "fp32"], help="big-gradient exchange for N > 1, fp16 payload unless fp32: auto (default) = p2p if torch symmetric memory sets up and the kernel passes its self-check against NCCL, else NCCL; p2p = two-shot reduction over NVLink peer pointers, one kernel with device-side barriers; multimem = the same shard reduced in the NVSwitch; fp16 = NCCL; fp32 = NCCL on the fp32 buffers")
    ap.add_argument("--mm-blocks", type=int, default=0, help="multimem exchange kernel: CTAs (0 = default)")
    ap.add_argument("--mm-unroll", type=int, default=4, help="multimem exchange kernel: 16-byte switch reductions in flight per thread (2 | 4 | 8)")
    ap.add_argument("--only", dest="all_workloads", action="store_false", help="measure only --workload (default: the other BASELINE configurations "
                    "are measured too, with --secondary-steps steps each, and reported under \"workloads\")")
    ap.add_argument("--secondary-steps", type=int, default=50)
    ap.add_argument("--no-strong", dest="strong", action="store_false", help="N > 1: skip the strong-scaling measurement (same total batch split over the ranks)")
    ap.add_argument("--no-graph", action="store_true", help="launch the step's kernels one by one instead of replaying a CUDA graph")
    ap.add_argument("--no-pipeline", action="store_true", help="serial step graph: do not overlap the next batch's march with the backward")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.rays is None:
        args.rays = DEFAULT_RAYS[args.workload]
    rank, world, local = dist_env()
    if args.impl == "reference":
        run_reference(args, rank, world, local)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU oracle timing)")
    from pvd_b200 import _native
    _native.lib()  # fail loudly if the extension is missing
    run_ours(args, rank, world, local)


if __name__ == "__main__":
    main()
