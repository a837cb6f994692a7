#!/usr/bin/env python
"""GPU probe: (1) start-up skew sweep of Layout W (KW_FD1D_SKEW_PERMILLE), march-kernel time at BASELINE
configs[1]; (2) FD1D-BS: fused march against two solves (wall clock of the host API and kernel time)."""
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
    base = None
    for x, t in ((1024, 1024), (512, 512)):
        for skew in (0, 250, 400, 500, 600):
            os.environ["KW_FD1D_SKEW_PERMILLE"] = str(skew)
            p = pricer("FD1D-GPU", t, x)
            ms = []
            for _ in range(4):
                err, got = p.price(opts)
                assert err == ""
                ms.append(p.info()["last_kernel_ms"])
            if skew == 0:
                base = got
            d = float(np.max(np.abs(got - base)))
            print("skew %4d permille  x=%d t=%d  variant %d  kernel ms %s  -> %.4f M options/s  maxdiff vs skew 0: %.1e"
                  % (skew, x, t, p.info()["variant"], ["%.3f" % m for m in ms], n / min(ms) / 1e3, d), flush=True)
            p.close()
    os.environ["KW_FD1D_SKEW_PERMILLE"] = "0"
    res = {}
    for fused in (1, 2):
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
    print("FD1D-BS fused vs two solves maxdiff %.2e" % float(np.max(np.abs(res[1] - res[2]))))


if __name__ == "__main__":
    main()
