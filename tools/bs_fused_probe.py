#!/usr/bin/env python
"""GPU probe, FD1D-BS at 1024^2: the fused marches (variants 252 and 251) against two solves -- wall clock of
the host API and march-kernel time."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kwinto-cuda_b200"))
import kwfd1d  # noqa: E402
from kwfd1d.synthetic import synthetic_options  # noqa: E402


def pricer(mode, t, x, **keys):
    cfg = kwfd1d.Config(PRICER=mode)
    cfg.set("FD1D.T_GRID_SIZE", t)
    cfg.set("FD1D.X_GRID_SIZE", x)
    for k, v in keys.items():
        cfg.set(k, v)
    err, p = kwfd1d.PricerFactory.create(cfg)
    assert err == "", err
    return p


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    opts = synthetic_options(n, 42)
    for variant in (233, 235):
        p = pricer("FD1D-GPU", 1024, 1024, **{"FD1D.GPU.VARIANT": variant})
        ms = []
        for _ in range(4):
            err, got = p.price(opts)
            assert err == ""
            ms.append(p.info()["last_kernel_ms"])
        print("FD1D variant %d kernel ms %s -> %.4f M options/s" % (variant, ["%.3f" % m for m in ms], n / min(ms) / 1e3),
              flush=True)
        p.close()
    res = {}
    for fused in (1, 4, 3):
        p = pricer("FD1D-BS-GPU", 1024, 1024, **{"FD1D.GPU.BS_FUSED": fused})
        wall, ms = [], []
        for _ in range(3):
            t0 = time.perf_counter()
            err, got = p.price(opts)
            wall.append(time.perf_counter() - t0)
            assert err == ""
            ms.append(p.info()["last_kernel_ms"])
        res[fused] = got
        print("FD1D-BS fused=%d variant %d  wall ms %s  last march kernel ms %s -> %.4f M options/s"
              % (fused, p.info()["variant"], ["%.2f" % (1e3 * w) for w in wall], ["%.3f" % m for m in ms],
                 n / min(wall) / 1e6), flush=True)
        p.close()
    for fused in (1, 4):
        p = pricer("FD1D-BS-GPU", 512, 512, **{"FD1D.GPU.BS_FUSED": fused})
        wall = []
        for _ in range(4):
            t0 = time.perf_counter()
            err, got = p.price(opts)
            wall.append(time.perf_counter() - t0)
            assert err == ""
        res[10 + fused] = got
        print("FD1D-BS 512x512 fused=%d variant %d  wall ms %s -> %.4f M options/s"
              % (fused, p.info()["variant"], ["%.2f" % (1e3 * w) for w in wall], n / min(wall) / 1e6), flush=True)
        p.close()
    print("512x512 fused vs two solves maxdiff %.2e" % float(np.max(np.abs(res[11] - res[14]))))
    print("FD1D-BS fused vs two solves maxdiff %.2e" % max(float(np.max(np.abs(res[1] - res[3]))), float(np.max(np.abs(res[1] - res[4])))))


if __name__ == "__main__":
    main()
