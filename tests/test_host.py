"""CPU-side tests: the C-ABI library loads and exports every declared symbol, the host mirror of
the reference interface behaves like the reference, multi-process sharding works (gloo, world 2).
No compute calls are made here -- the pricer has no CPU fallback."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g

    g.build()
    import kwfd1d

    return kwfd1d


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "kw_fd1d.h")).read()
    declared = set(re.findall(r"\b(kw_fd1d_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(built.EXPORTED_SYMBOLS)
    lib = built.load_library()
    for s in declared:
        assert getattr(lib, s) is not None
    assert b"sm_100a" in lib.kw_fd1d_version()


def test_option_layout_matches_reference(built):
    from kwfd1d.types import OPTION_DTYPE

    assert OPTION_DTYPE.itemsize == 56
    assert [OPTION_DTYPE.fields[k][1] for k in "tkzrqsew"] == [0, 8, 16, 24, 32, 40, 48, 49]
    import pyoracle

    assert pyoracle.OPTION_DTYPE == OPTION_DTYPE


def test_config_is_typed_like_the_reference(built):
    # src/Core/kwConfig.h:33-50: the map is chosen by the type of the default argument
    c = built.Config()
    c.set("FD1D.DENSITY", 1)  # stored as an integer ...
    assert c.get("FD1D.DENSITY", 0.25) == 0.25  # ... so the f64 lookup does not see it
    c.set("FD1D.DENSITY", 0.5)
    assert c.get("FD1D.DENSITY", 0.25) == 0.5
    c.set("PRICER", "FD1D-GPU")
    assert c.get("PRICER", "") == "FD1D-GPU"
    c.erase("PRICER")
    assert c.get("PRICER", "") == ""


def test_factory_errors(built):
    assert built.PricerFactory.create(built.Config())[0] == "PricerFactory: Missing PRICER key"
    assert built.PricerFactory.create(built.Config(PRICER="BS2"))[0] == "PricerFactory: Unknown PRICER = BS2"


def test_no_cpu_fallback(built):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    err, p = built.PricerFactory.create(built.Config(PRICER="FD1D-GPU"))
    assert p is None and "no CUDA device" in err and "no CPU fallback" in err
    with pytest.raises(RuntimeError):
        built.fp64_peak(0)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "kwinto-cuda_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(d, f)).read()
                assert "pyoracle" not in src and "kworacle" not in src and "libkwref" not in src, f


def test_shard_bounds(built):
    from kwfd1d.sharded import shard_bounds

    for n in (0, 1, 7, 32768, 1000003):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


WORKER = r"""
import os, sys
sys.path.insert(0, os.path.join(%(root)r, "kwinto-cuda_b200")); sys.path.insert(0, os.path.join(%(root)r, "oracle"))
import numpy as np, torch, torch.distributed as dist
from kwfd1d.sharded import price_sharded
from kwfd1d.synthetic import synthetic_options
import pyoracle
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
orc = pyoracle.Oracle()
def price_fn(o):            # the CHECKER stands in for the GPU pricer in this CPU test
    p, e = orc.fd1d(o, 24, 48, nthreads=1)
    return e, p
o = synthetic_options(101, 9, european_every=3)
err, full = price_sharded(price_fn, o, world, rank, dist)
want, _ = orc.fd1d(o, 24, 48, nthreads=1)
assert err == "" and np.array_equal(full, want), (rank, err)
o["k"][100] = 1e-9          # last rank's shard fails -> every rank sees the error
err, full = price_sharded(price_fn, o, world, rank, dist)
assert "not in range" in err, (rank, err)
dist.destroy_process_group()
print("ok", rank)
"""


def test_sharded_gather_world2_gloo(built, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


def test_stage_swizzle_is_a_conflict_free_permutation():
    # fd1d_warp.cuh: stage_swz -- the node-index swizzle of the set-up stage (Python copy of the one-line
    # formula, which is also checked against the source text): a permutation inside aligned groups of 16
    # doubles, conflict-free per half-warp for the set-up threads' writes (thread k, nodes 8k + i) and for the
    # owning warp's reads (lane l, nodes 8 * (NCH * l + c) + i)
    import os
    import re

    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kwinto-cuda_b200", "csrc",
                            "fd1d_warp.cuh")).read()
    m = re.search(r"int stage_swz\(int j\) \{ return (.*?); \}", src)
    assert m and m.group(1) == "j ^ ((j >> 4) & 7) ^ ((j >> 5) & 15)"

    def swz(j):
        return j ^ ((j >> 4) & 7) ^ ((j >> 5) & 15)

    for n in (512, 1024):
        assert sorted(swz(j) for j in range(n)) == list(range(n))
        assert all(swz(j) // 16 == j // 16 for j in range(n))
    for base in range(0, 128, 16):
        for i in range(8):
            banks = [swz(8 * k + i) % 16 for k in range(base, base + 16)]
            assert len(set(banks)) == 16
    for nch, worst_allowed in ((4, 1), (2, 2)):
        for half in (0, 16):
            for c in range(nch):
                for i in range(8):
                    banks = [swz((lane * nch + c) * 8 + i) % 16 for lane in range(half, half + 16)]
                    assert max(banks.count(b) for b in banks) <= worst_allowed


def test_gpu_pricer_header_compiles_against_the_reference_tree():
    """inside the reference, kwFd1dGpu.h is compiled with -DKW_WITH_REFERENCE against the
    reference's own Core/Pricer headers (kw::Pricer src/Pricer/kwPricer.h:12-22, factory shape
    src/Pricer/kwPricerFactory.h:15-41).  Pinned here where the reference tree is mounted (this container; the GPU
    box has no /root/reference)."""
    ref = "/root/reference/src"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not mounted")
    src = ('#include "kw/kwFd1dGpu.h"\n'
           "int main() { kw::Config c; kw::sPtr<kw::Pricer> p; auto e = kw::GpuPricerFactory::create(c, p);"
           " std::vector<kw::Option> a; std::vector<double> v; return (int)e.size(); }\n")
    r = subprocess.run(["g++", "-std=c++20", "-fsyntax-only", "-DKW_WITH_REFERENCE", "-I", ref,
                        "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "kwinto-cuda_b200", "host"),
                        "-x", "c++", "-"], input=src, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
