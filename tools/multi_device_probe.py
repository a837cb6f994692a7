#!/usr/bin/env python
"""One process, N devices through the C-ABI multi-device handle: tools/multi_device_probe.py [n_options] [x] [t]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kwinto-cuda_b200"))
import kwfd1d  # noqa: E402
from kwfd1d.synthetic import synthetic_options  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
x = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
t = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
ndev = kwfd1d.load_library().kw_fd1d_device_count()
o = synthetic_options(n, 7)
base = None
for k in sorted({1, 2, 4, 8, ndev}):
    if k > ndev:
        continue
    cfg = kwfd1d.Config(PRICER="FD1D-GPU")
    cfg.set("FD1D.T_GRID_SIZE", t)
    cfg.set("FD1D.X_GRID_SIZE", x)
    cfg.set("FD1D.GPU.DEVICES", ",".join(str(d) for d in range(k)))
    err, p = kwfd1d.PricerFactory.create(cfg)
    assert err == "", err
    p.price(o[: 8192 * k])
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        err, got = p.price(o)
        dt = time.perf_counter() - t0
        assert err == "", err
        best = dt if best is None else min(best, dt)
    if base is None:
        base = got
    i = p.info()
    print("devices %d used %d: %d options in %.1f ms = %.3f M options/s (max kernel %.1f ms) bit-identical to 1 device: %s"
          % (k, i["devices_used"], n, best * 1e3, n / best / 1e6, i["last_kernel_ms"], np.array_equal(got, base)), flush=True)
    p.close()
