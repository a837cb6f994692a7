#!/usr/bin/env python
"""Golden vectors for the fused FD1D-BS march (fd1d_warp_bs.cuh), from the UNMODIFIED reference.

Run in the build container (where /root/reference is mounted):

    make -C oracle ref && python tests/golden/make_golden_bs.py

Output: tests/golden/bs_fused.npz -- option batches that mix American / European, puts / calls and
chains with several members, priced by the reference's own "FD1D-BS" and "FD1D" pricers
(src/Pricer/kwFd1d_BlackScholes.cpp:15-43) through oracle/_ref/libkwref.so, raw f64.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "kwinto-cuda_b200"))
from pyoracle import RefLib  # noqa: E402
from kwfd1d.synthetic import synthetic_options  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def with_chain_members(o, every, seed):
    """Append copies of every `every`-th option with another strike / spot: same chain, new option."""
    extra = o[::every].copy()
    rng = np.random.default_rng(seed)
    extra["k"] *= 0.9 + 0.2 * rng.random(extra.shape[0])
    extra["s"] *= 0.95 + 0.1 * rng.random(extra.shape[0])
    return np.concatenate([o, extra])


def main():
    ref = RefLib()
    out = {}
    cases = [
        # key, options, tDim, xDim
        ("bs_1024", with_chain_members(synthetic_options(1200, 21, european_every=5, call_every=3), 4, 1), 1024, 1024),
        ("bs_700x200", with_chain_members(synthetic_options(160, 22, european_every=3, call_every=2), 3, 2), 200, 700),
        ("bs_513x64", synthetic_options(40, 23, european_every=4, call_every=2), 64, 513),
        # the two-chunk tile (256 < x <= 512), incl. the reference's default grid 512 x 512
        ("bs_512", with_chain_members(synthetic_options(1300, 24, european_every=6, call_every=4), 5, 3), 512, 512),
        ("bs_300x100", with_chain_members(synthetic_options(90, 25, european_every=3, call_every=2), 3, 4), 100, 300),
        # the wide tiles (round 2: fused march on the two- and four-warp kernels)
        ("bs_2048", with_chain_members(synthetic_options(200, 26, european_every=5, call_every=3), 4, 5), 2048, 2048),
        ("bs_4096x256", with_chain_members(synthetic_options(120, 27, european_every=4, call_every=3), 4, 6), 256, 4096),
        ("bs_1500x96", with_chain_members(synthetic_options(60, 28, european_every=3, call_every=2), 3, 7), 96, 1500),
        # x <= 256: four PDEs per warp (fd1d_iw.cuh, PACK = 4)
        ("bs_256", with_chain_members(synthetic_options(2400, 29, european_every=5, call_every=3), 5, 8), 256, 256),
        ("bs_200x80", with_chain_members(synthetic_options(70, 30, european_every=3, call_every=2), 3, 9), 80, 200),
    ]
    have = {}
    if os.path.exists(os.path.join(OUT, "bs_fused.npz")):  # cases already on file are kept (same seeds, same prices)
        with np.load(os.path.join(OUT, "bs_fused.npz")) as z:
            have = {k: z[k] for k in z.files}
    for key, o, t, x in cases:
        if key + "/fd1d" in have and np.array_equal(have[key + "/options"], o):
            for part in ("options", "grid", "fd1d_bs", "fd1d"):
                out[key + "/" + part] = have[key + "/" + part]
            continue
        out[key + "/options"] = o
        out[key + "/grid"] = np.array([t, x], dtype=np.int64)
        for mode, name in (("FD1D-BS", "fd1d_bs"), ("FD1D", "fd1d")):
            p, err = ref.price(o, t, x, mode=mode)
            assert err == "", err
            out[key + "/" + name] = p
        print(key, o.shape[0], "options", t, x)
    np.savez_compressed(os.path.join(OUT, "bs_fused.npz"), **out)
    print("bs_fused.npz", os.path.getsize(os.path.join(OUT, "bs_fused.npz")))


if __name__ == "__main__":
    main()
