#!/usr/bin/env python
"""BASELINE.json configs[2..4] on one GPU (run on the B200 box):
  config 3: batch sweep n = 128 .. 32768 at x = t = 512 / 1024, fp32 and fp64 (mirrors the reference's log/ sweeps)
  config 4: 1M-option synthetic portfolio, fp64, 1024^2 (this GPU's shard when --shard-of N is given)
  config 5: fine grid 4096^2, a bounded sample of the 65536 options
For each line: host-API wall time of one steady-state price() (host options in, host prices out; median of
`reps` calls after one warm-up -- SURVEY.md 8(d)'s metric), the march kernel's own time (CUDA events), options/s
and the fraction of the measured FP64 peak by the 11-flop count.  Writes a markdown table to stdout."""
import argparse
import statistics
import sys
import time

sys.path.insert(0, "kwinto-cuda_b200")
import numpy as np  # noqa: E402

import kwfd1d  # noqa: E402
from kwfd1d.synthetic import synthetic_options  # noqa: E402


def run(n, x, t, prec, seed, reps, peak):
    cfg = kwfd1d.Config(PRICER="FD1D-GPU")
    cfg.set("FD1D.T_GRID_SIZE", t)
    cfg.set("FD1D.X_GRID_SIZE", x)
    cfg.set("FD1D.GPU.PRECISION", prec)
    err, p = kwfd1d.PricerFactory.create(cfg)
    assert err == "", err
    o = synthetic_options(n, seed)
    err, _ = p.price(o)
    assert err == "", err
    wall, kern = [], []
    for _ in range(reps):
        t0 = time.perf_counter()
        err, prices = p.price(o)
        wall.append(time.perf_counter() - t0)
        kern.append(p.info()["last_kernel_ms"])
    w, k = statistics.median(wall), statistics.median(kern)
    frac = 11.0 * x * (t - 1) * n / (k * 1e-3) * 1e-12 / peak
    info = p.info()
    print(f"| {n} | {x}x{t} | {prec} | {info['variant']} | {w * 1e3:.3f} | {k:.3f} | {n / w:,.0f} | {n / (k * 1e-3):,.0f} | "
          f"{100 * frac:.1f}% | {info['mode_count']} |", flush=True)
    p.close()
    return prices


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="3,4,5")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--n4", type=int, default=1 << 20)
    ap.add_argument("--n5", type=int, default=8192)
    a = ap.parse_args()
    peak, mhz = kwfd1d.fp64_peak(0)
    print(f"measured FP64 peak {peak:.2f} TFLOP/s ({mhz:.0f} MHz effective)\n")
    print("| options | grid | march | variant | price() wall ms | kernel ms | options/s (host API) | options/s (kernel) | % FP64 peak | carry modes |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    cfgs = a.configs.split(",")
    if "3" in cfgs:
        for x in (512, 1024):
            for prec in ("f64", "f32"):
                n = 128
                while n <= 32768:
                    run(n, x, x, prec, 1000 + int(np.log2(n)), a.reps, peak)
                    n *= 2
    if "4" in cfgs:
        run(a.n4, 1024, 1024, "f64", 7, 2, peak)
    if "5" in cfgs:
        run(a.n5, 4096, 4096, "f64", 11, 1, peak)


if __name__ == "__main__":
    main()
