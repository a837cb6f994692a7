#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/gpu_sharded_check.py : the 6000-option fixture sharded over N GPUs,
prices gathered over NCCL, compared with the reference's golden prices on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kwinto-cuda_b200"))
import kwfd1d  # noqa: E402
from kwfd1d.sharded import price_sharded  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = np.load(os.path.join(ROOT, "tests", "golden", "portfolio_fd1d.npz"))
cfg = kwfd1d.Config(PRICER="FD1D-GPU")
cfg.set("FD1D.T_GRID_SIZE", 1024)
cfg.set("FD1D.X_GRID_SIZE", 1024)
cfg.set("FD1D.GPU.DEVICE", local)
err, pricer = kwfd1d.PricerFactory.create(cfg)
assert err == "", err
err, prices = price_sharded(pricer.price, g["options"], world, rank, dist, device=f"cuda:{local}")
assert err == "", err
d = float(np.max(np.abs(prices - g["fd1d_1024"])))
print(f"rank {rank}/{world}: 6000 options sharded, max|gpu - reference| = {d:.3e}", flush=True)
assert d <= 1e-9
dist.destroy_process_group()
