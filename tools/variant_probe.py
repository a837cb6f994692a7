#!/usr/bin/env python
"""GPU probe: march-kernel time of kernel variants at one shape: tools/variant_probe.py x t n v1 v2 ..."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kwinto-cuda_b200"))
import kwfd1d  # noqa: E402
from kwfd1d.synthetic import synthetic_options  # noqa: E402

x, t, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
opts = synthetic_options(n, 42)
base = None
for v in [int(a) for a in sys.argv[4:]]:
    cfg = kwfd1d.Config(PRICER="FD1D-GPU")
    cfg.set("FD1D.T_GRID_SIZE", t)
    cfg.set("FD1D.X_GRID_SIZE", x)
    cfg.set("FD1D.GPU.VARIANT", abs(v))
    if v >= 1000:
        cfg.set("FD1D.GPU.PRECISION", "f32")
    err, p = kwfd1d.PricerFactory.create(cfg)
    assert err == "", err
    ms = []
    for _ in range(4):
        err, got = p.price(opts)
        assert err == "", err
        ms.append(p.info()["last_kernel_ms"])
    if base is None:
        base = got
    print("x=%d t=%d n=%d variant %d regs %d kernel ms %s -> %.4f M options/s  maxdiff vs first %.1e"
          % (x, t, n, p.info()["variant"], p.info()["regs_per_thread"], ["%.3f" % m for m in ms], n / min(ms) / 1e3,
             float(np.max(np.abs(got - base)))), flush=True)
    p.close()
