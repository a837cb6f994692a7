#!/usr/bin/env python
"""GPU vs oracle per option on a stiff shape (few steps, fine grid): tools/parity_probe.py x t n [exact]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kwinto-cuda_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import kwfd1d  # noqa: E402
import pyoracle  # noqa: E402
from kwfd1d.synthetic import synthetic_options  # noqa: E402

x, t, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
o = synthetic_options(n, 32, call_every=2)
want, err = pyoracle.Oracle().fd1d(o, t, x, compress=False)
for variant in ([0] + [int(v) for v in sys.argv[4:]]):
    cfg = kwfd1d.Config(PRICER="FD1D-GPU")
    cfg.set("FD1D.T_GRID_SIZE", t)
    cfg.set("FD1D.X_GRID_SIZE", x)
    cfg.set("FD1D.GPU.COMPRESS", 0)
    cfg.set("FD1D.GPU.EXACT", 2)
    cfg.set("FD1D.GPU.VARIANT", variant)
    e, p = kwfd1d.PricerFactory.create(cfg)
    assert e == "", e
    e, got = p.price(o)
    print("variant", p.info()["variant"], "max |gpu - oracle| %.3e" % np.max(np.abs(got - want)))
    print(" per option:", " ".join("%.1e" % d for d in (got - want)))
np.save(os.path.join(ROOT, "gpurun_out", "parity_probe_gpu.npy"), got)
