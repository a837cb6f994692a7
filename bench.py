#!/usr/bin/env python
"""bench.py -- options/s of the Fd1d American-option pricer (BASELINE.json metric).

A "step" is one pass of the hot path (set-up + the whole Crank-Nicolson time march + price interpolation, ONE
march launch) over one batch of synthetic options.

    python bench.py [--gpus N --steps K --warmup W]               our arm, BASELINE configs[1] (default)
    python bench.py --config {2,3,4,5} ...                         the other BASELINE configs
    python bench.py --impl reference [--gpus N --steps K ...]      the reference's CPU pricer on the host cores

Configs (seeds and sizes of SURVEY.md 8(d); --n/--x/--t/--seed/--precision override):
  2  synthetic 32768 American puts PER GPU, fp64, x = t = 1024, seed 42 + rank           weak scaling   (default)
  3  batch sweep 128 ... 32768 options at x = t = 512 / 1024, fp32 and fp64, seed 1000 + log2 n   one GPU
  4  ONE 1 048 576-option portfolio, fp64, x = t = 1024, seed 7, sharded over the ranks   strong scaling
  5  ONE 65 536-option portfolio, fp64, x = t = 4096, seed 11, sharded over the ranks      strong scaling
Strong scaling = contiguous blocks (kwfd1d.sharded.shard_bounds), every rank prices its block, the prices are
all-gathered (NCCL) INSIDE the timed region.  The default run also carries config 4 as the sub-record "strong" at
every N, and on rank 0 the same portfolio through ONE call of the multi-device C-ABI handle
(kw_fd1d_create_multi, "c_abi_multi_device").

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for value / e2e / roofline / cpu_baseline.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "kwinto-cuda_b200"))

import numpy as np  # noqa: E402

UNIT = "options/s"
CONFIGS = {
    2: dict(n=32768, x=1024, t=1024, seed=42, scaling="weak",
            what="synthetic {n} American puts/GPU, {prec}, x={x} t={t}, mt19937_64 seed {seed}+rank (BASELINE configs[1])"),
    3: dict(n=32768, x=1024, t=1024, seed=1015, scaling="weak",
            what="batch sweep 128-32768 options at x=512/1024, fp32 and fp64, seeds 1000+log2(n) (BASELINE configs[2]); "
                 "headline entry: {n} options, {prec}, x={x} t={t}"),
    4: dict(n=1048576, x=1024, t=1024, seed=7, scaling="strong",
            what="ONE {n}-option synthetic portfolio, {prec}, x={x} t={t}, mt19937_64 seed {seed}, sharded over the GPUs and "
                 "gathered (BASELINE configs[3])"),
    5: dict(n=65536, x=4096, t=4096, seed=11, scaling="strong",
            what="ONE {n}-option synthetic portfolio, {prec}, x={x} t={t}, mt19937_64 seed {seed}, sharded over the GPUs and "
                 "gathered (BASELINE configs[4])"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed steps (0 = 10, or 2 for the 1 s steps of configs 4 / 5)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--n", "--options-per-gpu", dest="n", type=int, default=0,
                    help="options per step (per GPU for the weak configs; use the long form under torchrun, whose parser claims --n)")
    ap.add_argument("--x", type=int, default=0)
    ap.add_argument("--t", type=int, default=0)
    ap.add_argument("--seed", type=int, default=-1)
    ap.add_argument("--layout", default="auto", choices=["auto", "reg", "soa"])
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"],
                    help="f32 = fp64 set-up + fp32 march (config 3 sweeps); the headline metric is f64")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the strong / multi-device / pageable sub-records")
    ap.add_argument("--cpu-sample", type=int, default=0, help="options in the CPU baseline sample (0 = auto)")
    ap.add_argument("--cpu-full", type=int, default=-1,
                    help="1: also time ONE full-size reference call (35 s at 32768 x 1024^2); default: config 2 at N = 1")
    a = ap.parse_args()
    c = CONFIGS[a.config]
    a.n = a.n or c["n"]
    a.x = a.x or c["x"]
    a.t = a.t or c["t"]
    a.seed = a.seed if a.seed >= 0 else c["seed"]
    a.scaling = c["scaling"]
    if a.steps == 0:
        a.steps = 2 if a.config in (4, 5) else 10
    return a


def workload_name(a):
    prec = "fp64" if a.precision == "f64" else "fp32 march (fp64 set-up)"
    return CONFIGS[a.config]["what"].format(n=a.n, x=a.x, t=a.t, seed=a.seed, prec=prec)


def metric_name(a):
    return "options/sec, %s Fd1d x=%d t=%d" % ("fp64" if a.precision == "f64" else "fp32", a.x, a.t)


def flops_per_option(x, t):
    # 11 flop per node-step (hoisted count), node-steps = xDim*(tDim-1)
    return 11.0 * x * (t - 1)


# ------------------------------------------------------------------ clocks (B200_PROFILING.md)
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            f = [s.strip() for s in line.split(",")]
            if len(f) < 9 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------ the reference CPU arm
def cpu_reference(opts, t, x, sample, warm=True):
    """options/s of the reference's own CPU pricer (oracle/_ref, its thread pool on all host cores;
    falls back to the C port of it with one pthread per core).  Returns (value, info)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle

    s = opts[:sample]
    if pyoracle.RefLib.available():
        ref = pyoracle.RefLib()
        cores, kind = ref.pool_size(), "reference"

        def run(o):
            p, err = ref.price(o, t, x)
            assert err == "", err
            return p
    else:
        orc = pyoracle.Oracle()
        cores, kind = orc.max_threads(), "port"

        def run(o):
            p, err = orc.fd1d(o, t, x, compress=True, nthreads=0)
            assert err == "", err
            return p
    if warm:
        run(s[: max(1, min(len(s), cores))])  # first call starts the pool / faults pages in
    t0 = time.perf_counter()
    p = run(s)
    dt = time.perf_counter() - t0
    return len(s) / dt, {"cores": cores, "kind": kind, "seconds": dt, "prices": p}


def cpu_sample_size(a, cores, seconds=12.0):
    if a.cpu_sample > 0:
        return min(a.n, a.cpu_sample)
    # ~30 ns per node-step per core -> aim at ~12 s of wall time
    per_core_rate = 1.0 / (30e-9 * a.x * (a.t - 1))
    return int(max(cores, min(a.n, seconds * per_core_rate * cores)))


def run_reference_arm(a, rank, world):
    if rank != 0:
        return
    from kwfd1d.synthetic import synthetic_options

    cores = os.cpu_count() or 1
    sample = max(cores, cpu_sample_size(a, cores) // max(1, a.steps + a.warmup) * 2)
    sample = min(sample, a.n)
    opts = synthetic_options(min(a.n, max(sample, 4096)), a.seed)
    for _ in range(a.warmup):
        cpu_reference(opts, a.t, a.x, max(1, min(sample, cores)), warm=False)
    t0 = time.perf_counter()
    info = None
    for _ in range(a.steps):
        _, info = cpu_reference(opts, a.t, a.x, sample, warm=False)
    dt = time.perf_counter() - t0
    value = sample * a.steps / dt
    line = {
        "impl": "reference", "metric": metric_name(a), "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": a.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "x": a.x, "t": a.t, "options_per_gpu": a.n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                         "sample": f"each step = the first {sample} options of the workload, priced by the "
                                   f"reference's thread-pool CPU Fd1d pricer on {info['cores']} host threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ our arm
class Ctx:
    """torch / distributed plumbing shared by the measurement legs"""

    def __init__(self, rank, world, local_rank):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self.stream = torch.cuda.current_stream()
        self.sp = self.stream.cuda_stream
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]


def make_pricer(a, local_rank, x=None, t=None, precision=None, variant=None, devices=None):
    import kwfd1d

    cfg = kwfd1d.Config(PRICER="FD1D-GPU")
    cfg.set("FD1D.T_GRID_SIZE", int(t or a.t))
    cfg.set("FD1D.X_GRID_SIZE", int(x or a.x))
    cfg.set("FD1D.GPU.DEVICE", local_rank)
    cfg.set("FD1D.GPU.LAYOUT", a.layout)
    cfg.set("FD1D.GPU.VARIANT", a.variant if variant is None else variant)
    cfg.set("FD1D.GPU.PRECISION", precision or a.precision)
    if devices:
        cfg.set("FD1D.GPU.DEVICES", devices)
    err, pricer = kwfd1d.PricerFactory.create(cfg)
    if err:
        raise SystemExit("bench.py: " + err)
    return pricer


def measure(cx, pricer, opts_local, n_total, steps, warmup, strong, clocks=False, pageable=True):
    """Device-resident leg (`value`), the march kernel's own time, and the host-API legs of one workload.
    opts_local: this rank's options (its block of the ONE portfolio when `strong`, its own batch otherwise).
    strong: the price vectors of all ranks are all-gathered inside every timed step."""
    import kwfd1d
    from kwfd1d.sharded import gather_prices, shard_bounds

    torch = cx.torch
    n = opts_local.shape[0]
    h_in = torch.empty(max(n, 1) * 56, dtype=torch.uint8).pin_memory()
    h_opts = h_in.numpy().view(kwfd1d.OPTION_DTYPE)[:n]
    h_opts[:] = opts_local
    d_opts = h_in.cuda(non_blocking=False)
    d_prices = torch.empty(max(n, 1), dtype=torch.float64, device="cuda")
    gathered = [None]

    def gather():
        if strong and cx.world > 1:
            gathered[0] = gather_prices(d_prices[:n], n_total, cx.world, cx.rank, cx.dist)

    def step():
        cx.flush.zero_()  # L2 flush between iterations (inside the timed region: ~40 us)
        e = pricer.price_device(d_opts.data_ptr(), n, d_prices.data_ptr(), cx.sp)
        if e:
            raise SystemExit("bench.py: " + e)
        gather()

    for _ in range(warmup):
        step()
    cx.barrier()
    sampler = ClockSampler(cx.local_rank) if (clocks and cx.rank == 0) else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cx.barrier()
    t_wall0 = time.time()
    ev0.record(cx.stream)
    for _ in range(steps):
        step()
    ev1.record(cx.stream)
    cx.barrier()
    t_wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    err = pricer.sync(cx.sp)
    if err:
        raise SystemExit("bench.py: " + err)
    out = {"clocks": sampler.stop(t_wall0, t_wall1) if sampler else None}
    # the dominant kernel's own duration (CUDA events around the launch on its stream) and the gather's
    kernel_ms, gather_ms = [], []
    for _ in range(min(steps, 5)):
        cx.flush.zero_()
        pricer.price_device(d_opts.data_ptr(), n, d_prices.data_ptr(), cx.sp)
        pricer.sync(cx.sp)
        kernel_ms.append(pricer.info()["last_kernel_ms"])
        if strong and cx.world > 1:
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            cx.barrier()
            g0.record(cx.stream)
            gather()
            g1.record(cx.stream)
            torch.cuda.synchronize()
            gather_ms.append(g0.elapsed_time(g1))
    out["kernel_ms"] = statistics.mean(kernel_ms)
    out["gather_ms"] = statistics.mean(gather_ms) if gather_ms else 0.0
    out["launches_per_step"] = pricer.info()["launches"]
    got_dev = d_prices[:n].cpu().numpy()
    out["prices"] = got_dev
    if strong and cx.world > 1:
        lo, hi = shard_bounds(n_total, cx.world, cx.rank)
        full = gathered[0].cpu().numpy()
        assert full.shape[0] == n_total and np.array_equal(full[lo:hi], got_dev), "gathered vector != local block"
        out["prices_full"] = full

    # ---- end-to-end through the public host API: host options in, host prices out, every step.  Strong scaling
    #      under N ranks: H2D of the rank's block, march, all-gather, D2H of the whole price vector on every rank.
    def e2e_pass(host_opts):
        if strong and cx.world > 1:
            d_opts.copy_(torch.from_numpy(host_opts.view(np.uint8).reshape(-1)), non_blocking=True)
            e = pricer.price_device(d_opts.data_ptr(), n, d_prices.data_ptr(), cx.sp)
            if e:
                raise SystemExit("bench.py: " + e)
            gather()
            e = pricer.sync(cx.sp)
            if e:
                raise SystemExit("bench.py: " + e)
            return gathered[0].cpu().numpy()
        e, p = pricer.price(host_opts)
        if e:
            raise SystemExit("bench.py: " + e)
        return p

    def e2e_leg(host_opts):
        for _ in range(min(warmup, 2)):
            e2e_pass(host_opts)
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            p = e2e_pass(host_opts)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return dt, p

    t_e2e, p = e2e_leg(h_opts)
    assert np.array_equal(p[:n] if not (strong and cx.world > 1) else p, out.get("prices_full", got_dev)), \
        "host API and device API disagree"
    t_page = 0.0
    if pageable:
        t_page, p2 = e2e_leg(np.array(opts_local, copy=True))  # an ordinary (pageable) numpy array
        assert np.array_equal(p2, p)
    ms_max, e2e_ms_max, page_ms_max = cx.max_over_ranks(ms, t_e2e * 1e3, t_page * 1e3)
    total = n_total if strong else n * cx.world
    out.update(n_local=n, total=total, ms_per_step=ms_max / steps, value=total * steps / (ms_max * 1e-3),
               e2e_ms_per_step=e2e_ms_max / steps, e2e_value=total * steps / (e2e_ms_max * 1e-3),
               e2e_pageable_ms_per_step=page_ms_max / steps,
               e2e_pageable_value=(total * steps / (page_ms_max * 1e-3)) if pageable else None)
    return out


def roofline_of(a, info, n_launch, kernel_ms, local_rank, x=None, t=None):
    import kwfd1d

    F = flops_per_option(x or a.x, t or a.t)
    achieved = F * n_launch / (kernel_ms * 1e-3) * 1e-12
    try:
        peak_meas, _ = kwfd1d.fp64_peak(local_rank)
    except Exception:
        peak_meas = None
    try:
        probe = kwfd1d.dfma_probe(local_rank)
    except Exception:
        probe = None
    sm_max = 1965.0
    try:
        sm_max = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("sm_max_mhz", sm_max)
    except Exception:
        pass
    peak_nominal = info["sm_count"] * 64 * 2 * sm_max * 1e6 * 1e-12
    peak = peak_meas or peak_nominal
    traffic, traffic_src = None, None
    try:
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        tj = json.load(open(tp))
        key = "x%d" % (x or a.x)
        ent = tj.get(key) or (tj if "dram_bytes_per_launch" in tj else None)
        if ent:
            traffic = ent.get("dram_bytes_per_launch")
            traffic_src = {"file": "profiles/traffic.json", "kernel": ent.get("kernel"), "variant": ent.get("variant"),
                           "options_per_launch": ent.get("options_per_launch"), "captured": ent.get("captured"),
                           "file_mtime": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime(os.path.getmtime(tp))),
                           "stale": ent.get("variant") not in (None, info["variant"])}
    except Exception:
        pass
    iw_pack = {237: 1, 137: 1, 239: 1, 138: 2, 38: 4, 1237: 1, 1138: 2, 1038: 4}
    if info["variant"] in iw_pack:
        kname = "fd1d_iw_kernel (Layout W, independent warps%s%s)" % (
            {1: "", 2: ", two PDEs per warp", 4: ", four PDEs per warp"}[iw_pack[info["variant"]]],
            ", fp32 march" if info["variant"] >= 1000 else "")
    elif info["threads_per_pde"] > 32 and info["variant"] in (331, 336, 431, 436):
        kname = "fd1d_wide_kernel (Layout W, %d warps per PDE)" % (info["threads_per_pde"] // 32)
    elif info["threads_per_pde"] == 32 and (x or a.x) > 256:
        kname = "fd1d_warp_kernel (Layout W)"
    else:
        kname = "fd1d_%s_kernel" % info["layout"]
    r = {
        "bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "traffic": traffic, "traffic_source": traffic_src,
        "kernel": "%s variant %d" % (kname, info["variant"]), "kernel_ms": kernel_ms,
        "algorithmic_flop_per_option": F, "options_per_launch": n_launch,
        "peak_source": ("measured here: DFMA throughput probe kw_fd1d_fp64_peak (MEASURED_PEAKS.json has no FP64 figure)"
                        if peak_meas else "nominal"),
        "peak_nominal": peak_nominal, "frac_nominal": achieved / peak_nominal, "algorithmic_bytes_per_option": 64,
    }
    if probe:
        # every DFMA of the march reads three distinct registers; such DFMAs issue every 3 cycles, not 2
        # (DESIGN.md "Roofline"): the measured ceiling for this instruction form, reported next to the peak
        r["peak_3source_dfma"] = probe["3reg_64warps"]
        r["frac_of_3source_dfma_peak"] = achieved / probe["3reg_64warps"]
    return r


def sweep_config3(a, cx):
    """BASELINE configs[2]: batch sweep 128 ... 32768 options at x = t = 512 / 1024, fp32 and fp64 (mirrors the
    reference's log/ sweeps), seeds 1000 + log2(n): device-resident value, march time, host-API e2e pinned and pageable."""
    from kwfd1d.synthetic import synthetic_options

    rows = []
    for x in (512, 1024):
        for prec in ("f64", "f32"):
            pricer = make_pricer(a, cx.local_rank, x=x, t=x, precision=prec)
            for lg in (7, 9, 11, 13, 15):
                n = 1 << lg
                opts = synthetic_options(n, 1000 + lg)
                m = measure(cx, pricer, opts, n, steps=max(3, min(10, a.steps)), warmup=3, strong=False)
                info = pricer.info()
                rows.append({"x": x, "t": x, "precision": prec, "n": n, "seed": 1000 + lg, "variant": info["variant"],
                             "value": m["value"], "kernel_ms": m["kernel_ms"], "e2e": m["e2e_value"],
                             "e2e_pageable": m["e2e_pageable_value"],
                             "frac_fp64_peak": None})
            pricer.close()
    return rows


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference_arm(a, rank, world)
        return

    import torch
    import torch.distributed as dist

    import kwfd1d
    from kwfd1d.sharded import shard_bounds
    from kwfd1d.synthetic import synthetic_options

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the Fd1d GPU pricer has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cx = Ctx(rank, world, local_rank)
    strong = a.scaling == "strong"

    # ---- the headline workload
    pricer = make_pricer(a, local_rank)
    if strong:
        full = synthetic_options(a.n, a.seed)  # ONE portfolio; every rank draws it and keeps its block
        lo, hi = shard_bounds(a.n, world, rank)
        opts = full[lo:hi]
        n_total = a.n
    else:
        opts = synthetic_options(a.n, a.seed + rank)
        n_total = a.n * world
    m = measure(cx, pricer, opts, n_total, a.steps, a.warmup, strong, clocks=True)
    info = pricer.info()

    extras = {}
    if not a.no_extras and a.config == 2:
        # config 4 as a strong-scaling sub-record at every N: ONE 1 M-option portfolio (seed 7), blocks, gather timed
        c4 = CONFIGS[4]
        full4 = synthetic_options(c4["n"], c4["seed"])
        lo, hi = shard_bounds(c4["n"], world, rank)
        m4 = measure(cx, pricer, full4[lo:hi], c4["n"], steps=2, warmup=1, strong=True, pageable=False)
        if rank == 0:
            extras["strong"] = {
                "workload": c4["what"].format(n=c4["n"], x=a.x, t=a.t, seed=c4["seed"], prec="fp64"),
                "scaling": "strong", "n_gpus": world, "options": c4["n"], "value": m4["value"], "unit": UNIT,
                "ms_per_step": m4["ms_per_step"], "kernel_ms_rank0": m4["kernel_ms"], "gather_ms": m4["gather_ms"],
                "gather": ("all_gather_into_tensor (NCCL) of every rank's padded price block, inside the timed region"
                           if world > 1 else "none (one rank)"),
                "e2e": {"value": m4["e2e_value"], "ms_per_step": m4["e2e_ms_per_step"],
                        "api": "pinned host block -> H2D -> kw_fd1d_price_device -> all-gather -> D2H of all prices on every rank"
                               if world > 1 else "PricerFactory.create(FD1D-GPU).price(host options) -> host prices"},
            }
        cx.barrier()
        # the same portfolio through ONE call of the multi-device C-ABI handle (single process, rank 0):
        # kw_fd1d_create_multi over every visible device.  The other ranks wait on the HOST (the rendezvous store) --
        # an NCCL barrier would spin a kernel on their GPUs, which rank 0's shards are about to use (measured: 2.4x
        # slower marches next to a spinning barrier kernel).
        ndev = torch.cuda.device_count()
        store = dist.distributed_c10d._get_default_store() if world > 1 else None
        if rank != 0 and store is not None:
            store.wait(["kw_c_abi_multi_device_done"])
        if rank == 0:
            try:
                mp = make_pricer(a, 0, devices=",".join(str(d) for d in range(ndev)))
                mp.price(full4[:4096 * ndev])  # warm-up: contexts, buffers
                best, p_multi = None, None
                for _ in range(2):
                    t0 = time.perf_counter()
                    err, p_multi = mp.price(full4)
                    dt = time.perf_counter() - t0
                    if err:
                        raise RuntimeError(err)
                    best = dt if best is None else min(best, dt)
                mi = mp.info()
                lo0, hi0 = shard_bounds(c4["n"], world, 0)
                extras["c_abi_multi_device"] = {
                    "what": "ONE kw_fd1d_price call on a kw_fd1d_create_multi handle: host options in, host prices out "
                            "(contiguous blocks, one host thread + stream + buffers per device, D2H into the caller's slices)",
                    "devices": mi["n_devices"], "devices_used": mi["devices_used"], "options": c4["n"],
                    "value": c4["n"] / best, "unit": UNIT, "ms": best * 1e3, "max_kernel_ms": mi["last_kernel_ms"],
                    "bit_identical_to_rank0_block": bool(np.array_equal(p_multi[lo0:hi0], m4["prices"])),
                }
                mp.close()
            except Exception as e:  # never lose the headline line over the extra record
                extras["c_abi_multi_device"] = {"error": str(e)}
            if store is not None:
                store.set("kw_c_abi_multi_device_done", "1")
        cx.barrier()
    sweep = sweep_config3(a, cx) if a.config == 3 and world == 1 else None

    if rank == 0:
        n_launch = opts.shape[0]
        roofline = roofline_of(a, info, n_launch, m["kernel_ms"], local_rank)
        if sweep:
            pk = roofline["peak"]
            for r in sweep:
                if r["precision"] == "f64":
                    r["frac_fp64_peak"] = flops_per_option(r["x"], r["t"]) * r["n"] / (r["kernel_ms"] * 1e-3) * 1e-12 / pk
        line = {
            "metric": metric_name(a), "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": a.scaling,
            "vs_baseline": None, "dtype": a.precision, "data": "synthetic",
            "config": {"workload": workload_name(a), "config": a.config, "x": a.x, "t": a.t,
                       "options_per_gpu": a.n if not strong else None, "options_total": n_total,
                       "layout": info["layout"], "variant": info["variant"],
                       "threads_per_pde": info["threads_per_pde"], "ctas_per_sm": info["ctas_per_sm"],
                       "regs_per_thread": info["regs_per_thread"], "smem_per_cta": info["smem_per_cta"],
                       "grid": info["grid"], "device": info["device_name"],
                       "carry_mode_histogram": info["mode_count"],
                       "l2": "256 MiB device buffer zeroed before every step (inside the timed region)",
                       "parallelism": (f"dp{world} (ONE portfolio in contiguous blocks; prices all-gathered over NCCL inside "
                                       f"the timed region)" if strong else
                                       f"dp{world} (options sharded, no data-path collective)")},
            "roofline": roofline,
            "e2e": {"value": m["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": 56 * n_launch,
                    "d2h_bytes_per_step": 8 * (n_total if (strong and world > 1) else n_launch) + 8,
                    "ms_per_step": m["e2e_ms_per_step"],
                    "api": ("pinned host block -> H2D -> kw_fd1d_price_device -> all_gather_into_tensor -> D2H of all prices"
                            if (strong and world > 1) else
                            "PricerFactory.create(FD1D-GPU).price(host options) -> host prices, pinned input, chain compression on")},
            "e2e_pageable": {"value": m["e2e_pageable_value"], "unit": UNIT, "ms_per_step": m["e2e_pageable_ms_per_step"],
                             "api": "the same call with an ordinary (pageable) host array, as std::vector<Option>::data() is"},
            "gpu_launches": m["launches_per_step"] * a.steps,
            "clocks": m["clocks"],
        }
        if strong:
            line["gather_ms"] = m["gather_ms"]
        line.update(extras)
        if sweep:
            line["sweep"] = sweep
        if world == 1 and not a.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sample = cpu_sample_size(a, cores)
            v, ci = cpu_reference(opts, a.t, a.x, sample)
            d = float(np.max(np.abs(ci["prices"] - m["prices"][:sample])))
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": ci["cores"], "kind": ci["kind"],
                "sample": f"first {sample} options of the same workload, one call after a warm-up call, "
                          f"{ci['seconds']:.1f} s on {ci['cores']} host threads",
                "max_abs_diff_vs_gpu": d,
            }
            want_full = a.cpu_full == 1 or (a.cpu_full < 0 and a.config == 2 and a.n <= 32768 and a.x <= 1024)
            if want_full and sample < opts.shape[0]:
                vf, cf = cpu_reference(opts, a.t, a.x, opts.shape[0], warm=False)
                line["cpu_baseline"]["full"] = {
                    "value": vf, "unit": UNIT, "options": int(opts.shape[0]), "seconds": cf["seconds"],
                    "what": "ONE reference call on the whole workload (no sub-sampling)",
                    "max_abs_diff_vs_gpu": float(np.max(np.abs(cf["prices"] - m["prices"]))),
                }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
