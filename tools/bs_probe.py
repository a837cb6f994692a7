#!/usr/bin/env python
"""FD1D-BS throughput: fused march vs two solves: tools/bs_probe.py x t n"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kwinto-cuda_b200"))
import kwfd1d  # noqa: E402
from kwfd1d.synthetic import synthetic_options  # noqa: E402

x, t, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
o = synthetic_options(n, 42)
res = {}
for fused in (1, 0):
    cfg = kwfd1d.Config(PRICER="FD1D-BS-GPU")
    cfg.set("FD1D.T_GRID_SIZE", t)
    cfg.set("FD1D.X_GRID_SIZE", x)
    cfg.set("FD1D.GPU.BS_FUSED", fused)
    err, p = kwfd1d.PricerFactory.create(cfg)
    assert err == "", err
    p.price(o)
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        err, got = p.price(o)
        dt = time.perf_counter() - t0
        assert err == "", err
        best = dt if best is None else min(best, dt)
    res[fused] = got
    i = p.info()
    print("x=%d t=%d n=%d BS_FUSED=%d variant %d: %.2f ms = %.4f M options/s (march %.2f ms, %d launches)"
          % (x, t, n, fused, i["variant"], best * 1e3, n / best / 1e6, i["last_kernel_ms"], i["launches"]), flush=True)
print("max |fused - two solves| = %.2e" % float(np.max(np.abs(res[0] - res[1]))))
