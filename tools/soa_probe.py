#!/usr/bin/env python
"""Layout A (thread per PDE, SoA in HBM) against its HBM roofline: tools/soa_probe.py x t n1 n2 ..."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kwinto-cuda_b200"))
import kwfd1d  # noqa: E402
from kwfd1d.synthetic import synthetic_options  # noqa: E402

x, t = int(sys.argv[1]), int(sys.argv[2])
for n in [int(a) for a in sys.argv[3:]]:
    o = synthetic_options(n, 42)
    cfg = kwfd1d.Config(PRICER="FD1D-GPU")
    cfg.set("FD1D.T_GRID_SIZE", t)
    cfg.set("FD1D.X_GRID_SIZE", x)
    cfg.set("FD1D.GPU.LAYOUT", "soa")
    cfg.set("FD1D.GPU.COMPRESS", 0)
    err, p = kwfd1d.PricerFactory.create(cfg)
    assert err == "", err
    t0 = time.perf_counter()
    err, got = p.price(o)
    dt = time.perf_counter() - t0
    assert err == "", err
    ms = p.info()["last_kernel_ms"]
    traffic = 72.0 * x * (t - 1) * n  # bytes of the march
    print("Layout A x=%d t=%d n=%d: %.1f ms (march + set-up + value kernels) = %.1f k options/s, march traffic %.2f TB -> %.2f TB/s"
          % (x, t, n, ms, n / ms, traffic / 1e12, traffic / (ms * 1e-3) / 1e12), flush=True)
    p.close()
