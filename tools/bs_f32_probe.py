import sys, time
sys.path.insert(0, "kwinto-cuda_b200")
import numpy as np, kwfd1d
from kwfd1d.synthetic import synthetic_options
for x in (1024, 512):
    o = synthetic_options(32768, 42)
    for fused in (1, 0):
        cfg = kwfd1d.Config(PRICER="FD1D-BS-GPU"); cfg.set("FD1D.T_GRID_SIZE", x); cfg.set("FD1D.X_GRID_SIZE", x)
        cfg.set("FD1D.GPU.PRECISION", "f32"); cfg.set("FD1D.GPU.BS_FUSED", fused)
        err, p = kwfd1d.PricerFactory.create(cfg); assert err == "", err
        p.price(o); best = 1e9
        for _ in range(3):
            t0 = time.perf_counter(); err, got = p.price(o); best = min(best, time.perf_counter() - t0)
        print("fp32 FD1D-BS x=%d BS_FUSED=%d variant %d: %.2f ms = %.3f M options/s" % (x, fused, p.info()["variant"], best * 1e3, 32768 / best / 1e6), flush=True)
