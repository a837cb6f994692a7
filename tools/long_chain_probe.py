#!/usr/bin/env python
"""One huge chain next to many short ones: tools/long_chain_probe.py [members] [x] [t].  Compares the march-kernel time with
the long-chain path (default) and without it (KW_FD1D_NO_LONG=1: the marching warp interpolates every option itself)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kwinto-cuda_b200"))
import kwfd1d  # noqa: E402
from kwfd1d.synthetic import synthetic_options  # noqa: E402

members = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
x = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
t = int(sys.argv[3]) if len(sys.argv) > 3 else 64
base = synthetic_options(1200, 17)
e = np.repeat(base[3:4], members)
rng = np.random.default_rng(1)
e["k"] = base[3]["k"] * rng.uniform(0.7, 1.4, size=members)
o = np.concatenate([base, e])
cfg = kwfd1d.Config(PRICER="FD1D-GPU")
cfg.set("FD1D.T_GRID_SIZE", t)
cfg.set("FD1D.X_GRID_SIZE", x)
err, p = kwfd1d.PricerFactory.create(cfg)
assert err == "", err
for _ in range(3):
    err, got = p.price(o)
    assert err == "", err
    i = p.info()
    print("x=%d t=%d: %d chains, one with %d options: variant %d, long chains %d, march + value kernels %.3f ms"
          % (x, t, i["last_n_pde"], members, i["variant"], i["long_chains"], i["last_kernel_ms"]), flush=True)
