#!/usr/bin/env python
"""One FD1D-BS price call (for ncu): tools/bs_once.py <fused 1|2> [n] [t] [x]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kwinto-cuda_b200"))
import kwfd1d  # noqa: E402
from kwfd1d.synthetic import synthetic_options  # noqa: E402

fused = int(sys.argv[1])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
t = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
x = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
cfg = kwfd1d.Config(PRICER="FD1D-BS-GPU")
cfg.set("FD1D.T_GRID_SIZE", t)
cfg.set("FD1D.X_GRID_SIZE", x)
cfg.set("FD1D.GPU.BS_FUSED", fused)
err, p = kwfd1d.PricerFactory.create(cfg)
assert err == "", err
err, got = p.price(synthetic_options(n, 42))
assert err == "", err
print(p.info())
