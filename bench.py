#!/usr/bin/env python
"""bench.py -- options/s of the Fd1d American-option pricer (BASELINE.json metric).

A "step" is one pass of the hot path (set-up + the whole Crank-Nicolson time march + price
interpolation, ONE kernel launch) over one batch of synthetic options.  Workload at every N:
BASELINE.json configs[1] per GPU -- 32768 synthetic American puts (std::mt19937_64 seed 42 + rank,
SURVEY.md 8(d) generator), fp64, x = t = 1024 -- i.e. weak scaling; value = all ranks' options /
max-over-ranks device time.

    python bench.py [--gpus N --steps K --warmup W]              our arm (CUDA, sm_100a)
    python bench.py --impl reference [--gpus N --steps K ...]     the reference's CPU pricer

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for the definitions of
value / e2e / roofline / cpu_baseline.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "kwinto-cuda_b200"))

import numpy as np  # noqa: E402

METRIC = "options/sec, fp64 Fd1d x=1024 t=1024"
UNIT = "options/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--options-per-gpu", dest="n", type=int, default=32768,
                    help="options per GPU per step (use the long form under torchrun, whose parser claims --n)")
    ap.add_argument("--x", type=int, default=1024)
    ap.add_argument("--t", type=int, default=1024)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--layout", default="auto", choices=["auto", "reg", "soa"])
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"],
                    help="f32 = fp64 set-up + fp32 march (config 3 sweeps); the headline metric is f64")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="options in the CPU baseline sample (0 = auto)")
    return ap.parse_args()


def workload_name(a):
    prec = "fp64" if a.precision == "f64" else "fp32 march (fp64 set-up)"
    return f"synthetic {a.n} American puts/GPU, {prec}, x={a.x} t={a.t}, mt19937_64 seed {a.seed}+rank (BASELINE configs[1])"


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
def cpu_reference(opts, a, sample, warm=True):
    """options/s of the reference's own CPU pricer (oracle/_ref, its thread pool on all host cores;
    falls back to the C port of it with one pthread per core).  Returns (value, info)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle

    s = opts[:sample]
    if pyoracle.RefLib.available():
        ref = pyoracle.RefLib()
        cores, kind = ref.pool_size(), "reference"

        def run(o):
            p, err = ref.price(o, a.t, a.x)
            assert err == "", err
            return p
    else:
        orc = pyoracle.Oracle()
        cores, kind = orc.max_threads(), "port"

        def run(o):
            p, err = orc.fd1d(o, a.t, a.x, compress=True, nthreads=0)
            assert err == "", err
            return p
    if warm:
        run(s[: max(1, min(len(s), cores))])  # first call starts the pool / faults pages in
    t0 = time.perf_counter()
    p = run(s)
    dt = time.perf_counter() - t0
    return len(s) / dt, {"cores": cores, "kind": kind, "seconds": dt, "prices": p}


def cpu_sample_size(a, cores):
    if a.cpu_sample > 0:
        return min(a.n, a.cpu_sample)
    # ~30 ns per node-step per core -> aim at ~12 s of wall time
    per_core_rate = 1.0 / (30e-9 * a.x * (a.t - 1))
    return int(max(cores, min(a.n, 12.0 * per_core_rate * cores)))


def run_reference_arm(a, rank, world):
    if rank != 0:
        return
    from kwfd1d.synthetic import synthetic_options

    cores = os.cpu_count() or 1
    sample = max(cores, cpu_sample_size(a, cores) // max(1, a.steps + a.warmup) * 2)
    sample = min(sample, a.n)
    opts = synthetic_options(a.n, a.seed)
    for _ in range(a.warmup):
        cpu_reference(opts, a, max(1, min(sample, cores)), warm=False)
    t0 = time.perf_counter()
    info = None
    for _ in range(a.steps):
        _, info = cpu_reference(opts, a, sample, warm=False)
    dt = time.perf_counter() - t0
    value = sample * a.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
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
    from kwfd1d.synthetic import synthetic_options

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the Fd1d GPU pricer has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cfg = kwfd1d.Config(PRICER="FD1D-GPU")
    cfg.set("FD1D.T_GRID_SIZE", a.t)
    cfg.set("FD1D.X_GRID_SIZE", a.x)
    cfg.set("FD1D.GPU.DEVICE", local_rank)
    cfg.set("FD1D.GPU.LAYOUT", a.layout)
    cfg.set("FD1D.GPU.VARIANT", a.variant)
    cfg.set("FD1D.GPU.PRECISION", a.precision)
    err, pricer = kwfd1d.PricerFactory.create(cfg)
    if err:
        raise SystemExit("bench.py: " + err)

    n = a.n
    opts = synthetic_options(n, a.seed + rank)
    # pinned host staging for the e2e leg (numpy views over pinned torch storage)
    h_in = torch.empty(n * 56, dtype=torch.uint8).pin_memory()
    h_opts = h_in.numpy().view(kwfd1d.OPTION_DTYPE)
    h_opts[:] = opts
    d_opts = h_in.cuda(non_blocking=False)
    d_prices = torch.empty(n, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        flush.zero_()  # L2 flush between iterations (inside the timed region: ~40 us of ~30 ms)
        e = pricer.price_device(d_opts.data_ptr(), n, d_prices.data_ptr(), sp)
        if e:
            raise SystemExit("bench.py: " + e)

    # ---- device-resident leg: `value`
    for _ in range(a.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    barrier()
    t_wall0 = time.time()
    ev0.record(stream)
    for _ in range(a.steps):
        step()
    ev1.record(stream)
    barrier()
    t_wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    err = pricer.sync(sp)
    if err:
        raise SystemExit("bench.py: " + err)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    # the dominant kernel's own duration, CUDA events around the launch on the launching stream
    for _ in range(min(a.steps, 5)):
        step()
        torch.cuda.synchronize()
        kernel_ms.append(pricer.info()["last_kernel_ms"])
    launches_per_step = pricer.info()["launches"]  # status reset + chain compression (4) + march
    got_dev = d_prices.cpu().numpy()

    # ---- end-to-end leg through the public host API: pinned host options in, host prices out
    e2e_steps = a.steps
    for _ in range(min(a.warmup, 2)):
        err, p = pricer.price(h_opts)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        err, p = pricer.price(h_opts)
        if err:
            raise SystemExit("bench.py: " + err)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    assert np.array_equal(p, got_dev), "host API and device API disagree"

    times = torch.tensor([ms, t_e2e * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = (float(v) for v in times.cpu())
    total = n * world
    value = total * a.steps / (ms_max * 1e-3)
    e2e_value = total * e2e_steps / (e2e_ms_max * 1e-3)

    if rank == 0:
        info = pricer.info()
        F = flops_per_option(a.x, a.t)
        k_ms = statistics.mean(kernel_ms)
        achieved = F * n / (k_ms * 1e-3) * 1e-12
        try:
            peak_meas, mhz_eff = kwfd1d.fp64_peak(local_rank)
        except Exception:
            peak_meas, mhz_eff = None, None
        try:
            probe = kwfd1d.dfma_probe(local_rank)
        except Exception:
            probe = None
        sm_max = (clocks or {}).get("sm_max_mhz") or 1965.0
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            sm_max = mp.get("sm_max_mhz", sm_max)
        except Exception:
            mp = None
        peak_nominal = info["sm_count"] * 64 * 2 * sm_max * 1e6 * 1e-12
        peak = peak_meas or peak_nominal
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("dram_bytes_per_launch")
        except Exception:
            pass
        roofline = {
            "bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": traffic,
            "kernel": (f"fd1d_warp_kernel (Layout W) variant {info['variant']}" if info["threads_per_pde"] == 32 and a.x > 256
                       else f"fd1d_{info['layout']}_kernel variant {info['variant']}"), "kernel_ms": k_ms,
            "algorithmic_flop_per_option": F, "options_per_launch": n,
            "peak_source": ("measured here: DFMA throughput probe kw_fd1d_fp64_peak (MEASURED_PEAKS.json has no "
                            "FP64 figure)" if peak_meas else "nominal"),
            "peak_nominal": peak_nominal, "frac_nominal": achieved / peak_nominal,
            "algorithmic_bytes_per_option": 64,
        }
        if probe:
            # every DFMA of the march reads three distinct registers; such DFMAs issue every 3 cycles, not 2
            # (DESIGN.md "Roofline"): the measured ceiling for this instruction form, reported next to the peak
            roofline["peak_3source_dfma"] = probe["3reg_64warps"]
            roofline["frac_of_3source_dfma_peak"] = achieved / probe["3reg_64warps"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": a.precision, "data": "synthetic",
            "config": {"workload": workload_name(a), "x": a.x, "t": a.t, "options_per_gpu": n,
                       "layout": info["layout"], "variant": info["variant"],
                       "threads_per_pde": info["threads_per_pde"], "ctas_per_sm": info["ctas_per_sm"],
                       "regs_per_thread": info["regs_per_thread"], "smem_per_cta": info["smem_per_cta"],
                       "grid": info["grid"], "device": info["device_name"],
                       "carry_mode_histogram": info["mode_count"],
                       "l2": "256 MiB device buffer zeroed before every step (inside the timed region)",
                       "parallelism": f"dp{world} (options sharded, no data-path collective)"},
            "roofline": roofline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 56 * n, "d2h_bytes_per_step": 8 * n + 8,
                    "ms_per_step": e2e_ms_max / e2e_steps,
                    "api": "PricerFactory.create(FD1D-GPU).price(host options) -> host prices, pinned input, "
                           "chain compression on"},
            "gpu_launches": launches_per_step * a.steps,
            "clocks": clocks,
        }
        if world == 1 and not a.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sample = cpu_sample_size(a, cores)
            v, ci = cpu_reference(opts, a, sample)
            d = float(np.max(np.abs(ci["prices"] - got_dev[:sample])))
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": ci["cores"], "kind": ci["kind"],
                "sample": f"first {sample} options of the same workload, one call after a warm-up call, "
                          f"{ci['seconds']:.1f} s on {ci['cores']} host threads",
                "max_abs_diff_vs_gpu": d,
            }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
