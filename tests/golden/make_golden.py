#!/usr/bin/env python
"""Generate the committed golden vectors from the UNMODIFIED reference.

Run in the build container (where /root/reference is mounted):

    make -C oracle ref && python tests/golden/make_golden.py

Inputs : the reference's fixtures test/portfolio_{fd1d,qdfp}.csv.zst (decoded with pyarrow),
         its 12 known-answer options (test/kwPricer_test.cpp:15-29), and seeded synthetic
         batches (kwfd1d.synthetic).
Outputs: tests/golden/*.npz -- option arrays + the prices returned by the reference's own
         pricers through oracle/_ref/libkwref.so (PricerFactory::create -> Pricer::price),
         stored as raw f64 so the parity tests can compare to 1e-9 and the oracle bit-for-bit.
"""
import io
import os
import sys

import numpy as np
import pyarrow as pa
import pyarrow.csv as pacsv

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "kwinto-cuda_b200"))
from pyoracle import OPTION_DTYPE, RefLib  # noqa: E402
from kwfd1d.synthetic import synthetic_options  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

KAT = [  # ((t, k, z, r, q, s, e, w), want)  test/kwPricer_test.cpp:15-29
    ((1.0, 100., 0.2, 0.06, 0.02, 90., 0, -1), 10.627),
    ((1.0, 100., 0.2, 0.06, 0.02, 100., 0, -1), 5.885),
    ((1.0, 100., 0.2, 0.06, 0.02, 110., 0, -1), 2.987),
    ((1.0, 100., 0.2, 0.06, 0.02, 90., 0, +1), 4.668),
    ((1.0, 100., 0.2, 0.06, 0.02, 100., 0, +1), 9.729),
    ((1.0, 100., 0.2, 0.06, 0.02, 110., 0, +1), 16.633),
    ((1.0, 100., 0.2, 0.06, 0.08, 90., 1, -1), 13.988121682),
    ((1.0, 100., 0.2, 0.06, 0.08, 100., 1, -1), 8.409190396),
    ((1.0, 100., 0.2, 0.06, 0.08, 110., 1, -1), 4.6592955111),
    ((1.0, 100., 0.2, 0.06, 0.08, 90., 1, +1), 2.9472036256),
    ((1.0, 100., 0.2, 0.06, 0.08, 100., 1, +1), 6.8422540642),
    ((1.0, 100., 0.2, 0.06, 0.08, 110., 1, +1), 12.7940107108),
]


def load_fixture(name):
    """Decode test/<name>.csv.zst; columns as in src/Utils/kwPortfolio.cpp:21-58."""
    with pa.input_stream(os.path.join(REF, "test", name + ".csv.zst"), compression="zstd") as f:
        raw = f.read()
    tab = pacsv.read_csv(io.BytesIO(raw))
    n = tab.num_rows
    o = np.zeros(n, dtype=OPTION_DTYPE)
    o["t"] = tab["expiry"].to_numpy()
    o["s"] = tab["spot"].to_numpy().astype(np.float64)
    o["k"] = tab["strike"].to_numpy().astype(np.float64)
    o["z"] = tab["volatility"].to_numpy()
    o["r"] = tab["interest_rate"].to_numpy()
    o["q"] = tab["dividend_rate"].to_numpy()
    o["w"] = np.where(np.array(tab["parity"].to_pylist()) == "c", 1, -1)
    o["e"] = np.where(np.array(tab["exercise"].to_pylist()) == "a", 1, 0)
    return o, tab["price"].to_numpy().astype(np.float64)


def main():
    ref = RefLib()

    def price(o, t, x, mode="FD1D", **kw):
        p, err = ref.price(o, tdim=t, xdim=x, mode=mode, **kw)
        assert err == "", err
        return p

    # 1. known-answer tests, all three pricers, default 512x512
    o = np.array([k for k, _ in KAT], dtype=OPTION_DTYPE)
    np.savez_compressed(os.path.join(OUT, "kat.npz"), options=o, want=np.array([w for _, w in KAT]),
                        fd1d_512=price(o, 512, 512), bs=price(o, 512, 512, "BS"),
                        fd1d_bs_512=price(o, 512, 512, "FD1D-BS"), fd1d_1024=price(o, 1024, 1024))

    # 2. the two 6000-option fixtures at 512^2 and 1024^2
    for name in ("portfolio_fd1d", "portfolio_qdfp"):
        o, want = load_fixture(name)
        assert o.shape[0] == 6000
        np.savez_compressed(os.path.join(OUT, name + ".npz"), options=o, quantlib=want,
                            fd1d_512=price(o, 512, 512), fd1d_1024=price(o, 1024, 1024))

    # 3. seeded synthetic batches (sub-samples of the bench configs) + odd shapes
    syn = {}
    cases = [
        ("c2_1024", synthetic_options(512, 42), 1024, 1024, {}),            # head of config 2
        ("c3_512", synthetic_options(512, 1009), 512, 512, {}),             # config 3, n=512
        ("c4_1024", synthetic_options(256, 7), 1024, 1024, {}),             # head of config 4
        ("c5_4096", synthetic_options(16, 11), 4096, 4096, {}),             # head of config 5
        ("mix_512", synthetic_options(256, 5, european_every=3, call_every=2), 512, 512, {}),
        ("mix_1024", synthetic_options(128, 6, european_every=4, call_every=3), 1024, 1024, {}),
        ("odd_500x300", synthetic_options(64, 8, european_every=5, call_every=2), 300, 500, {}),
        ("odd_1000x100", synthetic_options(32, 9, call_every=4), 100, 1000, {}),
        ("odd_67x33", synthetic_options(32, 10, european_every=2, call_every=3), 33, 67, {}),
        ("dens_768", synthetic_options(32, 12, call_every=2), 256, 768, {"density": 0.1, "scale": 30.0}),
        ("tall_256x2048", synthetic_options(32, 13, call_every=5), 2048, 256, {}),
        ("wide_2048x256", synthetic_options(32, 14, call_every=5), 256, 2048, {}),
    ]
    for key, o, t, x, kw in cases:
        syn[key + "/options"] = o
        syn[key + "/grid"] = np.array([t, x], dtype=np.int64)
        syn[key + "/params"] = np.array([kw.get("density", 0.25), kw.get("scale", 50.0)])
        syn[key + "/fd1d"] = price(o, t, x, **kw)
    o = synthetic_options(96, 15, european_every=2, call_every=2)
    syn["bs_mix/options"] = o
    syn["bs_mix/grid"] = np.array([512, 512], dtype=np.int64)
    syn["bs_mix/params"] = np.array([0.25, 50.0])
    syn["bs_mix/fd1d"] = price(o, 512, 512)
    syn["bs_mix/fd1d_bs"] = price(o, 512, 512, "FD1D-BS")
    syn["bs_mix/bs"] = price(o, 512, 512, "BS")
    np.savez_compressed(os.path.join(OUT, "synthetic.npz"), **syn)

    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
